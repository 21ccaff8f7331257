// stage_args.hpp -- fills the POD argument structs of the stage kernels from a plan.
// Shared by the product (transform_engine.cpp, device pointers) and by the CPU emulation harness
// used in unit tests (tests/emu, host pointers), so both exercise the same wiring.
#pragma once
#include "index_plan.hpp"
#include "fast_stage_kernels.hpp"
#include "stage_kernels.hpp"

namespace spfft {
namespace b200 {

// Where the plan arrays live (device memory in the product).
template <typename T>
struct PlanPointers {
  const sb::cx<T>* twZ = nullptr;
  const sb::cx<T>* twY = nullptr;
  const sb::cx<T>* twX = nullptr;
  const sb::cx<T>* ftwX = nullptr;  // register-FFT stage twiddles, nullptr = generic kernels
  const sb::cx<T>* ftwY = nullptr;
  const sb::cx<T>* ftwZ = nullptr;
  const int* tileStart = nullptr;
  const int* entrySrc = nullptr;   // nullptr when TileMaps::identityOrder
  const int* entrySlot = nullptr;
  const int* bwdTileStart = nullptr;  // == tileStart unless TileMaps::hasDuplicates
  const int* bwdEntrySrc = nullptr;
  const int* bwdEntrySlot = nullptr;
  const int* xtStart = nullptr;
  const int* stickSlot = nullptr;
  const unsigned short* zInv = nullptr;  // gather-form inverse maps (nullptr: scatter form)
  const unsigned short* yInv = nullptr;
};

struct AxisPlans {
  sb::RadixPlan rpX, rpY, rpZ;
  int log2Vx = 0, log2Vy = 0, log2Vz = 0;
};

template <typename T>
inline sb::ZArgs<T> make_z_args(const IndexMaps& m, const TileMaps& t, const AxisPlans& ax,
                                const PlanPointers<T>& p, bool forward, sb::cx<T>* sticks,
                                const T* valuesIn, T* valuesOut, bool useScale) {
  sb::ZArgs<T> a{};
  a.nz = m.dimZ;
  a.log2V = ax.log2Vz;
  a.numTiles = t.numStickTiles;
  a.pitch = t.pitch;
  a.rp = ax.rpZ;
  a.tw = p.twZ;
  a.ftw = p.ftwZ;
  a.inv = p.zInv;
  if (forward) {
    a.tileStart = p.tileStart;
    a.entrySrc = p.entrySrc;
    a.entrySlot = p.entrySlot;
  } else {
    a.tileStart = p.bwdTileStart;
    a.entrySrc = p.bwdEntrySrc;
    a.entrySlot = p.bwdEntrySlot;
  }
  a.valuesIn = reinterpret_cast<const sb::cx<T>*>(valuesIn);
  a.valuesOut = reinterpret_cast<sb::cx<T>*>(valuesOut);
  a.sticks = sticks;
  a.symTile = t.symTile;
  a.symLane = t.symLane;
  a.useScale = useScale ? 1 : 0;
  // reference: T(1.0 / double(Nx*Ny*Nz)), src/execution/execution_gpu.cpp:58-59
  a.scale = static_cast<T>(1.0 / static_cast<double>(static_cast<unsigned long long>(m.dimX) *
                                                     static_cast<unsigned long long>(m.dimY) *
                                                     static_cast<unsigned long long>(m.dimZ)));
  return a;
}

template <typename T>
inline sb::YArgs<T> make_y_args(const IndexMaps& m, const TileMaps& t, const AxisPlans& ax,
                                const PlanPointers<T>& p, sb::cx<T>* sticks, sb::cx<T>* planes) {
  sb::YArgs<T> a{};
  a.ny = m.dimY;
  a.log2V = ax.log2Vy;
  a.nxf = m.dimXFreq;
  a.numPlanes = m.local_planes();
  a.numXTiles = t.numXTiles;
  a.pitch = t.pitch;
  a.zRowOffset = m.local_plane_offset();
  a.symmetry = m.type == SPFFT_TRANS_R2C ? 1 : 0;
  a.rp = ax.rpY;
  a.tw = p.twY;
  a.ftw = p.ftwY;
  a.inv = p.yInv;
  a.xtStart = p.xtStart;
  a.stickSlot = p.stickSlot;
  a.sticks = sticks;
  a.planes = planes;
  return a;
}

template <typename T>
inline sb::XArgs<T> make_x_args(const IndexMaps& m, const AxisPlans& ax, const PlanPointers<T>& p,
                                sb::cx<T>* planes, const void* spaceIn, void* spaceOut) {
  sb::XArgs<T> a{};
  a.nx = m.dimX;
  a.nxf = m.dimXFreq;
  a.ny = m.dimY;
  a.log2V = ax.log2Vx;
  a.numPlanes = m.local_planes();
  const int V = 1 << ax.log2Vx;
  a.r2c = m.type == SPFFT_TRANS_R2C ? 1 : 0;
  // the register-FFT real-row kernels transform two rows per lane (x_r2c_pair_tile)
  const int rowsPerTile = (a.r2c && p.ftwX) ? 2 * V : V;
  a.numRowTiles = (m.dimY + rowsPerTile - 1) / rowsPerTile;
  a.rp = ax.rpX;
  a.tw = p.twX;
  a.ftw = p.ftwX;
  a.planes = planes;
  a.spaceIn = spaceIn;
  a.spaceOut = spaceOut;
  return a;
}

}  // namespace b200
}  // namespace spfft
