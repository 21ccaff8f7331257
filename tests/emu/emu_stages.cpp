// emu_stages.cpp -- TEST HARNESS ONLY: runs the stage kernel *bodies* (spfft_b200/csrc/
// stage_kernels.hpp, fft_tile.hpp) block by block on the host, each phase as a loop over the
// emulated thread ids, with the same plan builder and argument wiring as the product.
//
// Purpose: unit-test tile/index arithmetic and the Stockham butterflies on a machine without a
// GPU (`pytest -m "not gpu"`). It is never linked into libspfft_b200.so, is not reachable from the
// SpFFT API, and is not a fallback: the product fails loudly without CUDA.
#define SB_EMULATE 1
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../spfft_b200/csrc/index_plan.hpp"
#include "../../spfft_b200/csrc/stage_args.hpp"
#include "../../spfft_b200/csrc/fast_stage_kernels.hpp"
#include "../../spfft_b200/csrc/fast3_stage_kernels.hpp"
#include "../../spfft_b200/csrc/stage_kernels.hpp"
#include "spfft/exceptions.hpp"

using namespace spfft::b200;

namespace {

// length dispatch of the register-FFT bodies (the product does the same in fast_launch.cuh)
#define EMU_DISPATCH(n, CALL) \
  switch (n) {                \
    case 32: CALL(32); break; \
    case 64: CALL(64); break; \
    case 128: CALL(128); break; \
    case 256: CALL(256); break; \
    case 512: CALL(512); break; \
    case 1024: CALL(1024); break; \
    default: throw spfft::InternalError(); \
  }

// N = 3 * 2^k register-FFT bodies (the product dispatches in fast3_launch.cuh)
#define EMU_DISPATCH3(n, CALL) \
  switch (n) {                 \
    case 96: CALL(96); break;  \
    case 192: CALL(192); break; \
    case 384: CALL(384); break; \
    case 768: CALL(768); break; \
    case 160: CALL(160); break; \
    case 320: CALL(320); break; \
    case 640: CALL(640); break; \
    default: throw spfft::InternalError(); \
  }

template <typename T>
void run_z(bool fwd, const sb::ZArgs<T>& a, int b, sb::Ctx ctx, sb::cx<T>* smem) {
  if (!a.ftw) {
    if (fwd) sb::z_forward_body<T>(a, b, ctx, smem); else sb::z_backward_body<T>(a, b, ctx, smem);
    return;
  }
  if (sb::is_fast3_length(a.nz)) {
    sb::Ctx c3{(1 << a.log2V) * (a.nz / fast_path_values_per_thread(a.nz))};
#define CALL(NN)                                                                                               \
  if (a.wireF32) { if (fwd) sb::z_forward_fast3<T, NN, true>(a, b, c3, smem); else sb::z_backward_fast3<T, NN, true>(a, b, c3, smem); } \
  else { if (fwd) sb::z_forward_fast3<T, NN>(a, b, c3, smem); else sb::z_backward_fast3<T, NN>(a, b, c3, smem); }
    EMU_DISPATCH3(a.nz, CALL)
#undef CALL
    return;
  }
  sb::Ctx c{(1 << sb::FastLanes<T>::log2V) * (a.nz / 8)};
#define CALL(NN)                                                                                               \
  if (a.wireF32) { if (fwd) sb::z_fast_any<T, NN, true, true>(a, b, c, smem); else sb::z_fast_any<T, NN, false, true>(a, b, c, smem); } \
  else { if (fwd) sb::z_fast_any<T, NN, true>(a, b, c, smem); else sb::z_fast_any<T, NN, false>(a, b, c, smem); }
  EMU_DISPATCH(a.nz, CALL)
#undef CALL
}
template <typename T>
void run_y(bool fwd, const sb::YArgs<T>& a, int b, sb::Ctx ctx, sb::cx<T>* smem) {
  if (!a.ftw) {
    if (fwd) sb::y_forward_body<T>(a, b, ctx, smem); else sb::y_backward_body<T>(a, b, ctx, smem);
    return;
  }
  if (sb::is_fast3_length(a.ny)) {
    sb::Ctx c3{(1 << a.log2V) * (a.ny / fast_path_values_per_thread(a.ny))};
#define CALL(NN)                                                                                               \
  if (a.wireF32) { if (fwd) sb::y_forward_fast3<T, NN, true>(a, b, c3, smem); else sb::y_backward_fast3<T, NN, true>(a, b, c3, smem); } \
  else { if (fwd) sb::y_forward_fast3<T, NN>(a, b, c3, smem); else sb::y_backward_fast3<T, NN>(a, b, c3, smem); }
    EMU_DISPATCH3(a.ny, CALL)
#undef CALL
    return;
  }
  sb::Ctx c{(1 << sb::FastLanes<T>::log2V) * (a.ny / 8)};
#define CALL(NN)                                                                                               \
  if (a.wireF32) { if (fwd) sb::y_forward_fast<T, NN, true>(a, b, c, smem); else sb::y_backward_fast<T, NN, true>(a, b, c, smem); } \
  else { if (fwd) sb::y_forward_fast<T, NN>(a, b, c, smem); else sb::y_backward_fast<T, NN>(a, b, c, smem); }
  EMU_DISPATCH(a.ny, CALL)
#undef CALL
}
template <typename T>
void run_x(bool fwd, const sb::XArgs<T>& a, int b, sb::Ctx ctx, sb::cx<T>* smem) {
  if (!a.ftw) {
    if (fwd) sb::x_forward_body<T>(a, b, ctx, smem); else sb::x_backward_body<T>(a, b, ctx, smem);
    return;
  }
  if (sb::is_fast3_length(a.nx)) {
    sb::Ctx c3{(1 << fast_path_log2_lanes_x(a.nx)) * (a.nx / fast_path_values_per_thread(a.nx))};
#define CALL(NN)                                                                                       \
  if (a.r2c) {                                                                                         \
    if (fwd) sb::x_r2c_fast3<T, NN, false>(a, b, c3, smem); else sb::x_r2c_fast3<T, NN, true>(a, b, c3, smem); \
  } else {                                                                                             \
    if (fwd) sb::x_c2c_fast3<T, NN, false>(a, b, c3, smem); else sb::x_c2c_fast3<T, NN, true>(a, b, c3, smem); \
  }
    EMU_DISPATCH3(a.nx, CALL)
#undef CALL
    return;
  }
  sb::Ctx c{(1 << fast_path_log2_lanes_x(a.nx)) * (a.nx / 8)};
#define CALL(NN)                                                                                     \
  if (a.r2c) {                                                                                       \
    if (fwd) sb::x_r2c_fast<T, NN, false>(a, b, c, smem); else sb::x_r2c_fast<T, NN, true>(a, b, c, smem); \
  } else {                                                                                           \
    if (fwd) sb::x_c2c_fast<T, NN, false>(a, b, c, smem); else sb::x_c2c_fast<T, NN, true>(a, b, c, smem); \
  }
  EMU_DISPATCH(a.nx, CALL)
#undef CALL
}

template <typename T>
int run(int type, int dimX, int dimY, int dimZ, int n, const int* triplets, int forward,
        const void* in, void* out, int scaling, int nthreads, int maxLog2V) {
  try {
    auto maps = make_local_index_maps(static_cast<SpfftTransformType>(type), dimX, dimY, dimZ, n,
                                      SPFFT_INDEX_TRIPLETS, triplets);
    AxisPlans ax;
    const long long smemLimit = 200 * 1024;
    const int cb = sizeof(sb::cx<T>);
    // maxLog2V == -2 forces the generic kernels everywhere (otherwise same choice as the product)
    const bool allowFast = maxLog2V != -2;
    const bool fastX = allowFast && fast_path_length(dimX, cb);
    const bool fastY = allowFast && fast_path_length(dimY, cb);
    const bool fastZ = allowFast && fast_path_length(dimZ, cb);
    const int fl = fast_path_log2_lanes(cb);
    ax.log2Vx = choose_log2_lanes(dimX, cb, smemLimit);
    ax.log2Vy = choose_log2_lanes(dimY, cb, smemLimit);
    ax.log2Vz = choose_log2_lanes(dimZ, cb, smemLimit);
    if (maxLog2V >= 0) {
      if (ax.log2Vx > maxLog2V) ax.log2Vx = maxLog2V;
      if (ax.log2Vy > maxLog2V) ax.log2Vy = maxLog2V;
      if (ax.log2Vz > maxLog2V) ax.log2Vz = maxLog2V;
    }
    // same x tile shape as the product, which keeps 8-row x tiles where its fused xy stage applies (wfft_xy.cu:
    // double precision C2C with dimX == dimY == 512; the emulation runs the separate y and x kernels there)
    const bool fusedShape = fastX && fastY && dimX == dimY && dimX == 512 && cb == 16 && type == SPFFT_TRANS_C2C;
    if (fastX) ax.log2Vx = fusedShape ? fl : fast_path_log2_lanes_x(dimX);
    if (fastY) ax.log2Vy = fast_path_log2_lanes(cb, dimY);
    if (fastZ) ax.log2Vz = fast_path_log2_lanes(cb, dimZ);
    ax.rpX = make_radix_plan(dimX);
    ax.rpY = make_radix_plan(dimY);
    ax.rpZ = make_radix_plan(dimZ);
    TileMaps t = build_tile_maps(*maps, ax.log2Vz, ax.log2Vy, fastZ, fastY);
    auto twX = make_roots<T>(dimX), twY = make_roots<T>(dimY), twZ = make_roots<T>(dimZ);
    auto ftwX = make_fast_twiddles<T>(dimX), ftwY = make_fast_twiddles<T>(dimY),
         ftwZ = make_fast_twiddles<T>(dimZ);
    PlanPointers<T> p;
    if (fastX) p.ftwX = ftwX.data();
    if (fastY) p.ftwY = ftwY.data();
    if (fastZ) p.ftwZ = ftwZ.data();
    p.twX = twX.data();
    p.twY = twY.data();
    p.twZ = twZ.data();
    p.tileStart = t.tileStart.data();
    p.entrySrc = t.identityOrder ? nullptr : t.entrySrc.data();
    p.entrySlot = t.entrySlot.data();
    if (t.hasDuplicates) {
      p.bwdTileStart = t.bwdTileStart.data();
      p.bwdEntrySrc = t.bwdEntrySrc.data();
      p.bwdEntrySlot = t.bwdEntrySlot.data();
    } else {
      p.bwdTileStart = p.tileStart;
      p.bwdEntrySrc = p.entrySrc;
      p.bwdEntrySlot = p.entrySlot;
    }
    if (!t.zInv.empty()) p.zInv = t.zInv.data();
    if (!t.yInv.empty()) p.yInv = t.yInv.data();
    p.xtStart = t.xtStart.data();
    p.stickSlot = t.stickSlot.data();

    const size_t stickElems = static_cast<size_t>(dimZ) * t.pitch;
    const size_t planeElems = static_cast<size_t>(dimZ) * dimY * maps->dimXFreq;
    // poison the intermediates: every stage must fully define what the next one reads
    std::vector<sb::cx<T>> sticks(stickElems + 1, sb::mk<T>(T(1e30), T(-1e30)));
    std::vector<sb::cx<T>> planes(planeElems + 1, sb::mk<T>(T(1e30), T(-1e30)));
    const int maxN = std::max(dimX, std::max(dimY, dimZ));
    std::vector<sb::cx<T>> smem(2 * static_cast<size_t>(maxN) * 32);
    sb::Ctx ctx{nthreads};

    if (!forward) {
      auto za = make_z_args<T>(*maps, t, ax, p, false, sticks.data(), static_cast<const T*>(in),
                               nullptr, false);
      for (int b = 0; b < za.numTiles; ++b) run_z<T>(false, za, b, ctx, smem.data());
      auto ya = make_y_args<T>(*maps, t, ax, p, sticks.data(), planes.data());
      for (int b = 0; b < ya.numXTiles * ya.numPlanes; ++b)
        run_y<T>(false, ya, b, ctx, smem.data());
      auto xa = make_x_args<T>(*maps, ax, p, planes.data(), nullptr, out);
      for (int b = 0; b < xa.numRowTiles * xa.numPlanes; ++b)
        run_x<T>(false, xa, b, ctx, smem.data());
    } else {
      auto xa = make_x_args<T>(*maps, ax, p, planes.data(), in, nullptr);
      for (int b = 0; b < xa.numRowTiles * xa.numPlanes; ++b)
        run_x<T>(true, xa, b, ctx, smem.data());
      auto ya = make_y_args<T>(*maps, t, ax, p, sticks.data(), planes.data());
      for (int b = 0; b < ya.numXTiles * ya.numPlanes; ++b)
        run_y<T>(true, ya, b, ctx, smem.data());
      auto za = make_z_args<T>(*maps, t, ax, p, true, sticks.data(), nullptr,
                               static_cast<T*>(out), scaling != 0);
      for (int b = 0; b < za.numTiles; ++b) run_z<T>(true, za, b, ctx, smem.data());
    }
    return 0;
  } catch (const spfft::GenericError& e) {
    return static_cast<int>(e.error_code());
  } catch (...) {
    return 1;
  }
}


// Whole DISTRIBUTED transform with all P ranks emulated in this process: every rank's stick buffer A
// and plane-side buffer Q are host arrays, so the peer-memory form of the exchange (z / y kernels
// storing straight into the owner's buffer, transform_engine.cpp) runs unchanged with peer[d] =
// rank d's array; peerMode == 0 runs the block form (what ncclSend / ncclRecv move) with memcpy.
template <typename T>
int run_distributed(int type, int dimX, int dimY, int dimZ, int P, const int* numLocal,
                    const int* const* triplets, const int* planesPerRank, int forward,
                    const void* const* in, void* const* out, int scaling, int nthreads, int peerMode,
                    int wireF32) {
  try {
    std::vector<std::shared_ptr<IndexMaps>> maps(P);
    std::vector<std::vector<long long>> counts(P, std::vector<long long>(6));
    std::vector<std::vector<int>> sticks(P);
    for (int r = 0; r < P; ++r) {
      maps[r] = make_local_index_maps(static_cast<SpfftTransformType>(type), dimX, dimY, dimZ, numLocal[r],
                                      SPFFT_INDEX_TRIPLETS, triplets[r]);
      counts[r] = {dimX, dimY, dimZ, planesPerRank[r], maps[r]->num_sticks(), numLocal[r]};
      sticks[r] = maps[r]->stickIndices;
    }
    for (int r = 0; r < P; ++r) finish_distributed_index_maps(*maps[r], r, counts, sticks);
    AxisPlans ax;
    const long long smemLimit = 200 * 1024;
    const int cb = sizeof(sb::cx<T>);
    const bool fastX = fast_path_length(dimX, cb), fastY = fast_path_length(dimY, cb),
               fastZ = fast_path_length(dimZ, cb);
    const int fl = fast_path_log2_lanes(cb);
    ax.log2Vx = fastX ? fast_path_log2_lanes_x(dimX) : choose_log2_lanes(dimX, cb, smemLimit);
    ax.log2Vy = fastY ? fast_path_log2_lanes(cb, dimY) : choose_log2_lanes(dimY, cb, smemLimit);
    ax.log2Vz = fastZ ? fast_path_log2_lanes(cb, dimZ) : choose_log2_lanes(dimZ, cb, smemLimit);
    (void)fl;
    ax.rpX = make_radix_plan(dimX);
    ax.rpY = make_radix_plan(dimY);
    ax.rpZ = make_radix_plan(dimZ);
    auto twX = make_roots<T>(dimX), twY = make_roots<T>(dimY), twZ = make_roots<T>(dimZ);
    auto ftwX = make_fast_twiddles<T>(dimX), ftwY = make_fast_twiddles<T>(dimY), ftwZ = make_fast_twiddles<T>(dimZ);
    std::vector<TileMaps> tiles(P);
    std::vector<ExchangePlan> exch(P);
    std::vector<PlanPointers<T>> ptrs(P);
    std::vector<std::vector<sb::cx<T>>> A(P), Q(P), planes(P);
    const sb::cx<T> poison = sb::mk<T>(T(1e30), T(-1e30));
    for (int r = 0; r < P; ++r) {
      tiles[r] = build_tile_maps(*maps[r], ax.log2Vz, ax.log2Vy, fastZ, fastY);
      exch[r] = build_exchange_plan(*maps[r], ax.log2Vz, ax.log2Vy, fastY);
      TileMaps& t = tiles[r];
      PlanPointers<T>& p = ptrs[r];
      if (fastX) p.ftwX = ftwX.data();
      if (fastY) p.ftwY = ftwY.data();
      if (fastZ) p.ftwZ = ftwZ.data();
      p.twX = twX.data();
      p.twY = twY.data();
      p.twZ = twZ.data();
      p.tileStart = t.tileStart.data();
      p.entrySrc = t.identityOrder ? nullptr : t.entrySrc.data();
      p.entrySlot = t.entrySlot.data();
      if (t.hasDuplicates) {
        p.bwdTileStart = t.bwdTileStart.data();
        p.bwdEntrySrc = t.bwdEntrySrc.data();
        p.bwdEntrySlot = t.bwdEntrySlot.data();
      } else {
        p.bwdTileStart = p.tileStart;
        p.bwdEntrySrc = p.entrySrc;
        p.bwdEntrySlot = p.entrySlot;
      }
      if (!t.zInv.empty()) p.zInv = t.zInv.data();
      // y stage over the sticks of ALL ranks (DevicePlan of a distributed transform)
      t.numXTiles = exch[r].numXTiles;
      p.xtStart = exch[r].xtStart.data();
      p.stickSlot = exch[r].stickSlot.data();
      if (!exch[r].yInv.empty()) p.yInv = exch[r].yInv.data();
      A[r].assign(static_cast<size_t>(dimZ) * t.pitch + 1, poison);
      Q[r].assign(static_cast<size_t>(exch[r].planeSideElements) + 1, poison);
      planes[r].assign(static_cast<size_t>(planesPerRank[r]) * dimY * maps[r]->dimXFreq + 1, poison);
    }
    const int maxN = std::max(dimX, std::max(dimY, dimZ));
    std::vector<sb::cx<T>> smem(2 * static_cast<size_t>(maxN) * 32);
    sb::Ctx ctx{nthreads};
    auto y_args = [&](int r, bool fwd) {
      auto ya = make_y_args<T>(*maps[r], tiles[r], ax, ptrs[r], A[r].data(), planes[r].data());
      const ExchangePlan& x = exch[r];
      if (fwd && peerMode) {
        for (int d = 0; d < P; ++d) ya.peer[d] = A[d].data();
        ya.stickRank = x.stickRank.data();
        ya.fwdBase = x.fwdBase.data();
        ya.tileFwdBase = x.tileFwdBase.data();
        ya.xtRotate = x.fwdTileRotate;
        if (!x.fwdTileOrder.empty()) ya.xtOrder = x.fwdTileOrder.data();
      }
      ya.sticks = Q[r].data();
      ya.srcBase = x.srcBase.data();
      ya.srcPitch = x.srcPitch.data();
      ya.tileBase = x.tileBase.data();
      ya.tilePitch = x.tilePitch.data();
      ya.zRowOffset = 0;
      ya.wireF32 = wireF32;
      return ya;
    };
    // block form of the exchange: rank r's block for d (stick side) <-> d's block from r (plane side)
    auto exchange_blocks = [&](bool toPlanes) {
      for (int r = 0; r < P; ++r)
        for (int d = 0; d < P; ++d) {
          const long long n = exch[r].stickCount[d];
          if (n != exch[d].planeCount[r]) throw spfft::InternalError();
          if (wireF32 && sizeof(T) == 8) {  // same element offsets, elements of cx<float>
            auto* a = reinterpret_cast<sb::cx<float>*>(A[r].data()) + exch[r].stickOffset[d];
            auto* q = reinterpret_cast<sb::cx<float>*>(Q[d].data()) + exch[d].planeOffset[r];
            if (toPlanes) std::copy(a, a + n, q); else std::copy(q, q + n, a);
            continue;
          }
          sb::cx<T>* a = A[r].data() + exch[r].stickOffset[d];
          sb::cx<T>* q = Q[d].data() + exch[d].planeOffset[r];
          if (toPlanes) std::copy(a, a + n, q); else std::copy(q, q + n, a);
        }
    };
    if (!forward) {
      for (int r = 0; r < P; ++r) {
        if (tiles[r].numStickTiles == 0) continue;
        auto za = make_z_args<T>(*maps[r], tiles[r], ax, ptrs[r], false, A[r].data(), static_cast<const T*>(in[r]),
                                 nullptr, false);
        if (peerMode) {
          for (int d = 0; d < P; ++d) za.peer[d] = Q[d].data();
          za.rowRank = exch[r].rowRank.data();
          za.rowOff = exch[r].rowOff.data();
        }
        za.wireF32 = wireF32;
        for (int b = 0; b < za.numTiles; ++b) run_z<T>(false, za, b, ctx, smem.data());
      }
      if (!peerMode) exchange_blocks(true);
      for (int r = 0; r < P; ++r) {
        if (planesPerRank[r] == 0) continue;
        auto ya = y_args(r, false);
        for (int b = 0; b < ya.numXTiles * ya.numPlanes; ++b) run_y<T>(false, ya, b, ctx, smem.data());
        auto xa = make_x_args<T>(*maps[r], ax, ptrs[r], planes[r].data(), nullptr, out[r]);
        for (int b = 0; b < xa.numRowTiles * xa.numPlanes; ++b) run_x<T>(false, xa, b, ctx, smem.data());
      }
    } else {
      for (int r = 0; r < P; ++r) {
        if (planesPerRank[r] == 0) continue;
        auto xa = make_x_args<T>(*maps[r], ax, ptrs[r], planes[r].data(), in[r], nullptr);
        for (int b = 0; b < xa.numRowTiles * xa.numPlanes; ++b) run_x<T>(true, xa, b, ctx, smem.data());
        if (exch[r].stickSlot.empty()) continue;
        auto ya = y_args(r, true);
        for (int b = 0; b < ya.numXTiles * ya.numPlanes; ++b) run_y<T>(true, ya, b, ctx, smem.data());
      }
      if (!peerMode) exchange_blocks(false);
      for (int r = 0; r < P; ++r) {
        if (tiles[r].numStickTiles == 0 || numLocal[r] == 0) continue;
        auto za = make_z_args<T>(*maps[r], tiles[r], ax, ptrs[r], true, A[r].data(), nullptr,
                                 static_cast<T*>(out[r]), scaling != 0);
        za.wireF32 = wireF32;
        for (int b = 0; b < za.numTiles; ++b) run_z<T>(true, za, b, ctx, smem.data());
      }
    }
    return 0;
  } catch (const spfft::GenericError& e) {
    return static_cast<int>(e.error_code());
  } catch (...) {
    return 1;
  }
}

}  // namespace

extern "C" {

// Distributed transform over P emulated ranks (see run_distributed). in / out: per-rank pointers
// (forward == 0: values -> slab; else slab -> values).
int sb_emu_transform_distributed(int isFloat, int type, int dimX, int dimY, int dimZ, int numRanks,
                                 const int* numLocal, const int* const* triplets, const int* planesPerRank,
                                 int forward, const void* const* in, void* const* out, int scaling,
                                 int nthreads, int peerMode, int wireF32) {
  return isFloat ? run_distributed<float>(type, dimX, dimY, dimZ, numRanks, numLocal, triplets, planesPerRank,
                                          forward, in, out, scaling, nthreads, peerMode, wireF32)
                 : run_distributed<double>(type, dimX, dimY, dimZ, numRanks, numLocal, triplets, planesPerRank,
                                           forward, in, out, scaling, nthreads, peerMode, wireF32);
}

// Whole local transform through the emulated stage kernels. forward == 0: `in` = values
// (2*n reals), `out` = space (z,y,x); forward != 0: `in` = space, `out` = values.
int sb_emu_transform(int isFloat, int type, int dimX, int dimY, int dimZ, int n,
                     const int* triplets, int forward, const void* in, void* out, int scaling,
                     int nthreads, int maxLog2V) {
  return isFloat ? run<float>(type, dimX, dimY, dimZ, n, triplets, forward, in, out, scaling,
                              nthreads, maxLog2V)
                 : run<double>(type, dimX, dimY, dimZ, n, triplets, forward, in, out, scaling,
                               nthreads, maxLog2V);
}

// Tile addresses of the x stage swizzle (fast_fft.hpp: SwzX) for every (n, lane): out[n*V + lane].
int sb_emu_swzx(int elemBytes, int log2V, int n, int* out) {
  const int V = 1 << log2V;
  for (int i = 0; i < n; ++i) {
    for (int lane = 0; lane < V; ++lane) {
      int a = -1;
#define SWZ(E, L) if (elemBytes == E && log2V == L) a = sb::SwzX<E>::at<L>(i, lane);
      SWZ(16, 0) SWZ(16, 1) SWZ(16, 2) SWZ(16, 3)
      SWZ(8, 0) SWZ(8, 1) SWZ(8, 2) SWZ(8, 3) SWZ(8, 4)
#undef SWZ
      out[i * V + lane] = a;
    }
  }
  return 0;
}

// A single batched 1-D transform through sb::tile_fft (lanes = 1 << log2V sequences of length n,
// tile layout n*V + lane), for butterfly-level tests. data is overwritten with the result.
int sb_emu_tile_fft(int isFloat, int n, int log2V, int backward, int swizzle, void* data,
                    int nthreads) {
  try {
    sb::RadixPlan rp = make_radix_plan(n);
    sb::Ctx ctx{nthreads};
    const size_t elems = static_cast<size_t>(n) << log2V;
    auto go = [&](auto tag) {
      using T = decltype(tag);
      auto tw = make_roots<T>(n);
      std::vector<sb::cx<T>> a(elems), b(elems);
      std::memcpy(a.data(), data, elems * sizeof(sb::cx<T>));
      sb::cx<T>* r;
      if (backward)
        r = swizzle ? sb::tile_fft<T, true, true>(a.data(), b.data(), rp, log2V, tw.data(), ctx)
                    : sb::tile_fft<T, true, false>(a.data(), b.data(), rp, log2V, tw.data(), ctx);
      else
        r = swizzle ? sb::tile_fft<T, false, true>(a.data(), b.data(), rp, log2V, tw.data(), ctx)
                    : sb::tile_fft<T, false, false>(a.data(), b.data(), rp, log2V, tw.data(), ctx);
      std::memcpy(data, r, elems * sizeof(sb::cx<T>));
    };
    if (isFloat)
      go(float{});
    else
      go(double{});
    return 0;
  } catch (...) {
    return 1;
  }
}
}

// Hand-out order of the fused xy stage (wfft_xy.cu): the dense decode enumerates exactly the valid items of
// xy_decode (fast_stage_kernels.hpp) in the same order; every B tile comes after all A tiles of its plane and
// every A tile after the B tiles of the plane that used its ring slot (ring > lag). Returns 0 or an error code.
template <int TILES>
static int check_xy_order(int numPlanes, int lag, int ring) {
  sb::XYArgs<double> a{};
  a.y.numPlanes = numPlanes;
  a.y.numXTiles = TILES;
  a.x.numRowTiles = TILES;
  a.lag = lag;
  a.ring = ring;
  const long long total = sb::xy_total_items<double, true>(a);
  int dense = 0;
  std::vector<int> aSeen(numPlanes, 0), bSeen(numPlanes, 0);
  for (long long item = 0; item < total; ++item) {
    const sb::XYItem it = sb::xy_decode<double, true>(a, (int)item);
    if (!it.valid) continue;
    const sb::XYItem d = sb::w_decode_dense<TILES>(dense, numPlanes, lag);
    if (d.roleA != it.roleA || d.plane != it.plane || d.tile != it.tile) return 1;
    if (d.plane < 0 || d.plane >= numPlanes || d.tile < 0 || d.tile >= TILES) return 2;
    if (d.roleA) {
      if (d.plane >= ring && bSeen[d.plane - ring] != TILES) return 3;  // slot not yet consumed
      ++aSeen[d.plane];
    } else {
      if (aSeen[d.plane] != TILES) return 4;  // plane not yet complete
      ++bSeen[d.plane];
    }
    ++dense;
  }
  if (dense != 2 * TILES * numPlanes) return 5;
  for (int p = 0; p < numPlanes; ++p)
    if (aSeen[p] != TILES || bSeen[p] != TILES) return 6;
  return 0;
}
// tiles: items per plane and role, 64 (double precision) or 32 (single precision: 16 columns / rows per item)
extern "C" int sb_emu_check_xy_order(int numPlanes, int lag, int ring, int tiles) {
  return tiles == 64 ? check_xy_order<64>(numPlanes, lag, ring) : (tiles == 32 ? check_xy_order<32>(numPlanes, lag, ring) : -1);
}
