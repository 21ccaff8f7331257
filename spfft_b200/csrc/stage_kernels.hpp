// stage_kernels.hpp -- bodies of the six generic stage kernels (any transform length).
//
//   backward:  z_backward  (decompress + stick symmetry + z-FFT, writes plane-major sticks)
//              y_backward  (gather sticks of an x-tile + plane symmetry + y-FFT, writes planes)
//              x_backward  (x-FFT rows, C2C or C2R, writes the space domain)
//   forward :  x_forward -> y_forward -> z_forward (+ compress, + scaling)
//
// They replace, fused, the reference GPU stages
//   decompress/compress   src/compression/gpu_kernels/compression_kernels.cu:40-150
//   stick/plane symmetry  src/symmetry/gpu_kernels/symmetry_kernels.cu:39-165
//   local transpose       src/transpose/gpu_kernels/local_transpose_kernels.cu:48-201
//   cuFFT z / xy plans    src/fft/transform_1d_gpu.hpp:52-141, transform_2d_gpu.hpp:51-140,
//                         transform_real_2d_gpu.hpp:54-257
// wired in the order of src/execution/execution_gpu.cpp:254-400.
//
// Intermediate layouts (all complex, element counts):
//   sticks  [z][pitch]           plane-major: row z holds the value of every local stick at z;
//                                pitch = numSticks rounded up to the z-tile width
//   planes  [zl][y][nxf]         x innermost (same as the reference GPU plane buffer,
//                                src/execution/execution_gpu.cpp:88-89)
//   space   [zl][y][nx]          complex (C2C) or real, unpadded (R2C)  -- user visible
#pragma once
#include <cstddef>
#include "cx.hpp"
#include "fft_tile.hpp"

namespace sb {

// Distributed transforms over NVLink peer memory: at most this many GPUs (one NVSwitch box)
constexpr int kMaxPeers = 8;

template <typename T>
struct ZArgs {
  int nz;         // dimZ = transform length
  int log2V;      // sticks per tile = 1 << log2V
  int numTiles;   // ceil(numSticks / V)
  int pitch;      // row pitch of the stick buffer
  RadixPlan rp;
  const cx<T>* tw;        // forward roots of unity, length nz
  const cx<T>* ftw;       // stage twiddles of the register FFT (fast_fft.hpp) or nullptr
  const int* tileStart;   // [numTiles+1]: range of sparse entries of every stick tile
  const int* entrySrc;    // [numEntries]: position in the user's value array; nullptr = identity
  const int* entrySlot;   // [numEntries]: z*V + lane inside the tile
  const cx<T>* valuesIn;  // backward input (user order)
  cx<T>* valuesOut;       // forward output (user order)
  cx<T>* sticks;          // [nz][pitch]
  int symTile, symLane;   // R2C: tile / lane of the (x=0,y=0) stick, -1 if not local
  int useScale;
  T scale;
  int pfDist;             // L2 prefetch distance in tiles (resident CTAs), 0 = off; set by the launcher
  // register-FFT kernels, values in stick order without duplicates: inverse map
  // [tile][thread][8] -> offset of the element's value from the tile's first entry, 0xFFFF = none
  const unsigned short* inv;
  // Distributed backward over peer memory (replaces pack + MPI_Alltoallv + unpack,
  // transpose_mpi_compact_buffered_gpu.cpp:163-224): the kernel stores row z of its sticks straight
  // into the plane-side buffer of the rank that owns plane z, at peer[rowRank[z]] + rowOff[z]
  // (peer[me] = the local buffer). rowRank == nullptr: local rows sticks + z*pitch.
  cx<T>* peer[kMaxPeers];
  const unsigned char* rowRank;  // [nz]
  const long long* rowOff;       // [nz] element offset of this rank's row inside the owner's buffer
  // Distributed double-precision transform with a single-precision wire format
  // (SPFFT_EXCH_*_FLOAT, reference: complex_conversion.cuh:35-54 inside the pack / unpack kernels of
  // compact_buffered_kernels.cu:50-52,114-115): the stick buffer holds cx<float>, same element offsets.
  int wireF32;
};

// Element conversion between the arithmetic type and the type of the exchanged buffers (W = cx<T>,
// or cx<float> for the single-precision wire format of a double-precision transform).
template <typename T, typename S>
SB_HD cx<T> from_wire(cx<S> v) {
  return mk<T>(static_cast<T>(v.x), static_cast<T>(v.y));
}
template <typename W, typename T>
SB_HD W to_wire(cx<T> v) {
  W r;
  r.x = static_cast<decltype(r.x)>(v.x);
  r.y = static_cast<decltype(r.y)>(v.y);
  return r;
}
// Compile-time choice of the exchanged element type: the register-FFT kernels are instantiated once per
// wire format (a run-time branch inside one kernel cost registers: +4 % on the y stage at 512^3).
template <typename T, bool WIRE>
struct WireElemOf {
  using type = cx<T>;
};
template <>
struct WireElemOf<double, true> {
  using type = cx<float>;
};
template <typename T, bool WIRE>
using WireElem = typename WireElemOf<T, WIRE>::type;

// Runs f(tag) with tag = cx<float>{} when the wire format is single precision, cx<T>{} otherwise
// (generic kernels only: one kernel serves both formats).
template <typename T, typename F>
SB_DEV void with_wire_type(int wireF32, F f) {
  if (sizeof(T) == 8 && wireF32)
    f(cx<float>{});
  else
    f(cx<T>{});
}

// Row z of the stick buffer a z kernel writes (backward) / reads (forward), elements of type W.
template <typename T, typename W = cx<T>>
SB_DEV W* z_row(const ZArgs<T>& a, int z) {
  if (a.rowRank) return reinterpret_cast<W*>(a.peer[a.rowRank[z]]) + a.rowOff[z];
  return reinterpret_cast<W*>(a.sticks) + (size_t)z * a.pitch;
}

template <typename T>
struct YArgs {
  int ny;          // dimY = transform length
  int log2V;       // x columns per tile
  int nxf;         // dimXFreq
  int numPlanes;   // local planes
  int numXTiles;   // ceil(nxf / V)
  int pitch;       // row pitch of the stick buffer
  int zRowOffset;  // row of the stick buffer that holds local plane 0
  int symmetry;    // R2C: hermitian fill of the x=0 column
  RadixPlan rp;
  const cx<T>* tw;
  const cx<T>* ftw;      // stage twiddles of the register FFT or nullptr (generic path)
  const int* xtStart;    // [numXTiles+1]: range of sticks per x tile (sticks are sorted by x*Ny+y)
  const int* stickSlot;  // [numSticks]: y*V + (x mod V)
  cx<T>* sticks;
  cx<T>* planes;         // [numPlanes][ny][nxf]
  int pfDist;            // L2 prefetch distance in blocks, 0 = off; set by the launcher
  // register-FFT kernels: inverse map [x tile][thread][8] -> offset of the element's stick from
  // the tile's first stick, 0xFFFF = none
  const unsigned short* inv;
  // distributed transforms: `sticks` is the plane-side exchange buffer and stick e of local plane
  // zl lives at sticks[srcBase[e] + zl*srcPitch[e]] (ExchangePlan, index_plan.hpp); nullptr: the
  // local plane-major layout sticks[(zl + zRowOffset)*pitch + e]
  const int* srcBase;
  const int* srcPitch;
  // per x tile: sticks contiguous at sticks[tileBase[xt] + zl*tilePitch[xt] + i] (tilePitch != 0)
  const int* tileBase;
  const int* tilePitch;
  // Distributed forward over peer memory: stick e of local plane zl is stored straight into the
  // stick buffer of its owner, at peer[stickRank[e]] + fwdBase[e] + zl*srcPitch[e] (row
  // planeOffset(me) + zl of the owner's plane-major buffer); single-source tiles at
  // peer[stickRank[first stick]] + tileFwdBase[xt] + zl*tilePitch[xt] + i. stickRank == nullptr:
  // `sticks` + srcBase / tileBase (local exchange buffer).
  cx<T>* peer[kMaxPeers];
  const unsigned char* stickRank;  // [numSticks over all ranks]
  const int* fwdBase;              // [numSticks over all ranks]
  const int* tileFwdBase;          // [numXTiles]
  // forward kernels visit the x tiles of a plane starting at tile xtRotate: with peer stores
  // every rank then targets a different destination at any time (no ingress hot spot)
  int xtRotate;
  // ... or in the order xtOrder[0 .. numXTiles) when set: the tiles of the destination ranks interleaved
  // (next rank first), so that the NVLink stores and the local stores of a plane overlap instead of coming
  // in two bursts, and every rank still addresses all destinations evenly
  const int* xtOrder;
  int wireF32;  // the exchanged buffers (`sticks`, `peer`) hold cx<float> (see ZArgs)
};

// x tile the forward y kernels work on in position `pos` of a plane's visiting order
template <typename T>
SB_DEV int y_forward_tile_at(const YArgs<T>& a, int pos) {
  if (a.xtOrder) return a.xtOrder[pos];
  return (pos + a.xtRotate) % a.numXTiles;
}

// Where stick e (global sorted list) of local plane zl lives for a distributed y kernel:
// backward reads the local plane-side buffer, forward writes the owner's stick buffer.
template <typename T, bool FWD, typename W = cx<T>>
SB_DEV W* y_dist_stick(const YArgs<T>& a, int e, int zl) {
  if (FWD && a.stickRank)
    return reinterpret_cast<W*>(a.peer[a.stickRank[e]]) + (size_t)a.fwdBase[e] + (size_t)zl * a.srcPitch[e];
  return reinterpret_cast<W*>(a.sticks) + (size_t)a.srcBase[e] + (size_t)zl * a.srcPitch[e];
}
// First stick of single-source tile xt (tilePitch[xt] != 0) at local plane zl.
template <typename T, bool FWD, typename W = cx<T>>
SB_DEV W* y_dist_tile(const YArgs<T>& a, int xt, int zl) {
  if (FWD && a.stickRank)
    return reinterpret_cast<W*>(a.peer[a.stickRank[a.xtStart[xt]]]) + (size_t)a.tileFwdBase[xt] +
           (size_t)zl * a.tilePitch[xt];
  return reinterpret_cast<W*>(a.sticks) + (size_t)a.tileBase[xt] + (size_t)zl * a.tilePitch[xt];
}

template <typename T>
struct XArgs {
  int nx, nxf, ny;
  int log2V;        // rows per tile
  int numPlanes;
  int numRowTiles;  // ceil(ny / V)
  int r2c;
  RadixPlan rp;
  const cx<T>* tw;
  const cx<T>* ftw;   // stage twiddles of the register FFT or nullptr (generic path)
  cx<T>* planes;      // [numPlanes][ny][nxf]
  const void* spaceIn;  // forward input
  void* spaceOut;       // backward output
  int pfDist;           // L2 prefetch distance in blocks, 0 = off; set by the launcher
};

// -------------------------------------------------------------------------------------------
// Batched multi-transform (reference: multi_transform_internal.hpp:50-176 runs the transforms one
// by one, phase-interleaved). Transforms that share one immutable plan (clones: the bands of a
// plane-wave code) run as ONE launch per stage with blockIdx.y = band; only the data pointers
// differ per band and travel in this table as a kernel parameter.
// -------------------------------------------------------------------------------------------
constexpr int kMaxBands = 32;  // per launch (1.5 KB of kernel parameters)

template <typename T>
struct BandTable {
  const cx<T>* valuesIn[kMaxBands];
  cx<T>* valuesOut[kMaxBands];
  cx<T>* sticks[kMaxBands];
  cx<T>* planes[kMaxBands];
  const void* spaceIn[kMaxBands];
  void* spaceOut[kMaxBands];
};

// The arguments of band `band`: the shared ones with this band's pointers (local transforms only:
// the distributed tables are cleared, which also lets the compiler drop those branches).
template <typename T>
SB_DEV ZArgs<T> band_args(const ZArgs<T>& a0, const BandTable<T>& b, int band) {
  ZArgs<T> a = a0;
  a.valuesIn = b.valuesIn[band];
  a.valuesOut = b.valuesOut[band];
  a.sticks = b.sticks[band];
  a.rowRank = nullptr;
  a.rowOff = nullptr;
  a.wireF32 = 0;
  return a;
}
template <typename T>
SB_DEV YArgs<T> band_args(const YArgs<T>& a0, const BandTable<T>& b, int band) {
  YArgs<T> a = a0;
  a.sticks = b.sticks[band];
  a.planes = b.planes[band];
  a.srcBase = nullptr;
  a.srcPitch = nullptr;
  a.tileBase = nullptr;
  a.tilePitch = nullptr;
  a.stickRank = nullptr;
  a.fwdBase = nullptr;
  a.tileFwdBase = nullptr;
  a.xtRotate = 0;
  a.xtOrder = nullptr;
  a.wireF32 = 0;
  return a;
}
template <typename T>
SB_DEV XArgs<T> band_args(const XArgs<T>& a0, const BandTable<T>& b, int band) {
  XArgs<T> a = a0;
  a.planes = b.planes[band];
  a.spaceIn = b.spaceIn[band];
  a.spaceOut = b.spaceOut[band];
  return a;
}

// Hermitian completion of one lane of a tile, low index first (reference semantics:
// src/symmetry/symmetry_host.hpp:47-58,73-90; GPU twin symmetry_kernels.cu:56-78,119-141).
template <typename T>
SB_DEV void hermitian_fill_lane(cx<T>* A, int n, int lane, int log2V, Ctx ctx) {
  (void)ctx;
  const int half = n / 2;
  SB_PHASE_BEGIN
  for (int i = 1 + tid; i <= half; i += nthr) {
    const cx<T> v = A[(i << log2V) + lane];
    if (nonzero(v)) A[((n - i) << log2V) + lane] = conj(v);
  }
  SB_PHASE_END
  SB_PHASE_BEGIN
  for (int i = half + 1 + tid; i < n; i += nthr) {
    const cx<T> v = A[(i << log2V) + lane];
    if (nonzero(v)) A[((n - i) << log2V) + lane] = conj(v);
  }
  SB_PHASE_END
}

// -------------------------------------------------------------------------------------------
// z stage
// -------------------------------------------------------------------------------------------
template <typename T, typename W>
SB_DEV void z_backward_body_w(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* smem) {
  const int V = 1 << a.log2V;
  const int n = a.nz << a.log2V;
  cx<T>* A = smem;
  cx<T>* B = smem + n;
  SB_PHASE_BEGIN
  for (int i = tid; i < n; i += nthr) A[i] = mk<T>(0, 0);
  SB_PHASE_END
  SB_PHASE_BEGIN
  const int e1 = a.tileStart[tile + 1];
  for (int e = a.tileStart[tile] + tid; e < e1; e += nthr) {
    const int src = a.entrySrc ? a.entrySrc[e] : e;
    A[a.entrySlot[e]] = a.valuesIn[src];
  }
  SB_PHASE_END
  if (tile == a.symTile) hermitian_fill_lane<T>(A, a.nz, a.symLane, a.log2V, ctx);
  cx<T>* R = tile_fft<T, true, false>(A, B, a.rp, a.log2V, a.tw, ctx);
  SB_PHASE_BEGIN
  for (int i = tid; i < n; i += nthr) {
    const int z = i >> a.log2V;
    const int lane = i & (V - 1);
    z_row<T, W>(a, z)[(size_t)tile * V + lane] = to_wire<W>(R[i]);
  }
  SB_PHASE_END
}
template <typename T>
SB_DEV void z_backward_body(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* smem) {
  with_wire_type<T>(a.wireF32, [&](auto w) { z_backward_body_w<T, decltype(w)>(a, tile, ctx, smem); });
}

template <typename T, typename W>
SB_DEV void z_forward_body_w(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* smem) {
  const int V = 1 << a.log2V;
  const int n = a.nz << a.log2V;
  cx<T>* A = smem;
  cx<T>* B = smem + n;
  SB_PHASE_BEGIN
  const W* in = reinterpret_cast<const W*>(a.sticks) + (size_t)tile * V;
  for (int i = tid; i < n; i += nthr) {
    const int z = i >> a.log2V;
    const int lane = i & (V - 1);
    A[i] = from_wire<T>(in[(size_t)z * a.pitch + lane]);
  }
  SB_PHASE_END
  cx<T>* R = tile_fft<T, false, false>(A, B, a.rp, a.log2V, a.tw, ctx);
  SB_PHASE_BEGIN
  const int e1 = a.tileStart[tile + 1];
  for (int e = a.tileStart[tile] + tid; e < e1; e += nthr) {
    const int dst = a.entrySrc ? a.entrySrc[e] : e;
    cx<T> v = R[a.entrySlot[e]];
    if (a.useScale) v = a.scale * v;
    a.valuesOut[dst] = v;
  }
  SB_PHASE_END
}
template <typename T>
SB_DEV void z_forward_body(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* smem) {
  with_wire_type<T>(a.wireF32, [&](auto w) { z_forward_body_w<T, decltype(w)>(a, tile, ctx, smem); });
}

// -------------------------------------------------------------------------------------------
// y stage
// -------------------------------------------------------------------------------------------
template <typename T, typename W>
SB_DEV void y_backward_body_w(const YArgs<T>& a, int block, Ctx ctx, cx<T>* smem) {
  const int V = 1 << a.log2V;
  const int n = a.ny << a.log2V;
  const int xt = block % a.numXTiles;
  const int zl = block / a.numXTiles;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  cx<T>* planeTile = a.planes + (size_t)zl * a.ny * a.nxf + (size_t)xt * V;
  const int lanesValid = (a.nxf - xt * V) < V ? (a.nxf - xt * V) : V;
  if (e0 == e1) {
    // empty x tile: the x stage still reads these columns -> store zeros, no transform
    SB_PHASE_BEGIN
    for (int i = tid; i < n; i += nthr) {
      const int y = i >> a.log2V;
      const int lane = i & (V - 1);
      if (lane < lanesValid) planeTile[(size_t)y * a.nxf + lane] = mk<T>(0, 0);
    }
    SB_PHASE_END
    return;
  }
  cx<T>* A = smem;
  cx<T>* B = smem + n;
  SB_PHASE_BEGIN
  for (int i = tid; i < n; i += nthr) A[i] = mk<T>(0, 0);
  SB_PHASE_END
  SB_PHASE_BEGIN
  const W* base = reinterpret_cast<const W*>(a.sticks);
  const W* row = base + (size_t)(zl + a.zRowOffset) * a.pitch;
  for (int e = e0 + tid; e < e1; e += nthr)
    A[a.stickSlot[e]] = from_wire<T>(a.srcBase ? base[(size_t)a.srcBase[e] + (size_t)zl * a.srcPitch[e]] : row[e]);
  SB_PHASE_END
  if (a.symmetry && xt == 0) hermitian_fill_lane<T>(A, a.ny, 0, a.log2V, ctx);
  cx<T>* R = tile_fft<T, true, false>(A, B, a.rp, a.log2V, a.tw, ctx);
  SB_PHASE_BEGIN
  for (int i = tid; i < n; i += nthr) {
    const int y = i >> a.log2V;
    const int lane = i & (V - 1);
    if (lane < lanesValid) planeTile[(size_t)y * a.nxf + lane] = R[i];
  }
  SB_PHASE_END
}
template <typename T>
SB_DEV void y_backward_body(const YArgs<T>& a, int block, Ctx ctx, cx<T>* smem) {
  with_wire_type<T>(a.wireF32, [&](auto w) { y_backward_body_w<T, decltype(w)>(a, block, ctx, smem); });
}

template <typename T, typename W>
SB_DEV void y_forward_body_w(const YArgs<T>& a, int block, Ctx ctx, cx<T>* smem) {
  const int V = 1 << a.log2V;
  const int n = a.ny << a.log2V;
  const int xt = y_forward_tile_at<T>(a, block % a.numXTiles);
  const int zl = block / a.numXTiles;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  if (e0 == e1) return;  // no stick needs these columns
  const cx<T>* planeTile = a.planes + (size_t)zl * a.ny * a.nxf + (size_t)xt * V;
  const int lanesValid = (a.nxf - xt * V) < V ? (a.nxf - xt * V) : V;
  cx<T>* A = smem;
  cx<T>* B = smem + n;
  SB_PHASE_BEGIN
  for (int i = tid; i < n; i += nthr) {
    const int y = i >> a.log2V;
    const int lane = i & (V - 1);
    A[i] = lane < lanesValid ? planeTile[(size_t)y * a.nxf + lane] : mk<T>(0, 0);
  }
  SB_PHASE_END
  cx<T>* R = tile_fft<T, false, false>(A, B, a.rp, a.log2V, a.tw, ctx);
  SB_PHASE_BEGIN
  W* row = reinterpret_cast<W*>(a.sticks) + (size_t)(zl + a.zRowOffset) * a.pitch;
  for (int e = e0 + tid; e < e1; e += nthr) {
    if (a.srcBase)
      *y_dist_stick<T, true, W>(a, e, zl) = to_wire<W>(R[a.stickSlot[e]]);
    else
      row[e] = to_wire<W>(R[a.stickSlot[e]]);
  }
  SB_PHASE_END
}
template <typename T>
SB_DEV void y_forward_body(const YArgs<T>& a, int block, Ctx ctx, cx<T>* smem) {
  with_wire_type<T>(a.wireF32, [&](auto w) { y_forward_body_w<T, decltype(w)>(a, block, ctx, smem); });
}

// -------------------------------------------------------------------------------------------
// x stage (rows are contiguous in global memory; tile = V consecutive rows, swizzled lanes)
// -------------------------------------------------------------------------------------------
template <typename T>
SB_DEV void x_backward_body(const XArgs<T>& a, int block, Ctx ctx, cx<T>* smem) {
  const int V = 1 << a.log2V;
  const int n = a.nx << a.log2V;
  const int rt = block % a.numRowTiles;
  const int zl = block / a.numRowTiles;
  const int y0 = rt * V;
  const int rows = (a.ny - y0) < V ? (a.ny - y0) : V;
  const size_t rowBase = (size_t)zl * a.ny + y0;
  cx<T>* A = smem;
  cx<T>* B = smem + n;
  if (a.r2c || rows < V) {
    // C2R rows are completed by conjugation below; zero first so that every slot is defined
    SB_PHASE_BEGIN
    for (int i = tid; i < n; i += nthr) A[i] = mk<T>(0, 0);
    SB_PHASE_END
  }
  SB_PHASE_BEGIN
  const int total = rows * a.nxf;
  const int mirrorMax = a.nx - a.nxf;  // x in [1, mirrorMax] has a distinct mirror nx - x
  for (int i = tid; i < total; i += nthr) {
    const int r = i / a.nxf;
    const int x = i - r * a.nxf;
    const cx<T> v = a.planes[(rowBase + r) * a.nxf + x];
    A[at<true>(x, r, a.log2V)] = v;
    if (a.r2c && x >= 1 && x <= mirrorMax) A[at<true>(a.nx - x, r, a.log2V)] = conj(v);
  }
  SB_PHASE_END
  cx<T>* R = tile_fft<T, true, true>(A, B, a.rp, a.log2V, a.tw, ctx);
  SB_PHASE_BEGIN
  const int total = rows * a.nx;
  if (a.r2c) {
    T* out = static_cast<T*>(a.spaceOut);
    for (int i = tid; i < total; i += nthr) {
      const int r = i / a.nx;
      const int x = i - r * a.nx;
      out[(rowBase + r) * a.nx + x] = R[at<true>(x, r, a.log2V)].x;
    }
  } else {
    cx<T>* out = static_cast<cx<T>*>(a.spaceOut);
    for (int i = tid; i < total; i += nthr) {
      const int r = i / a.nx;
      const int x = i - r * a.nx;
      out[(rowBase + r) * a.nx + x] = R[at<true>(x, r, a.log2V)];
    }
  }
  SB_PHASE_END
}

template <typename T>
SB_DEV void x_forward_body(const XArgs<T>& a, int block, Ctx ctx, cx<T>* smem) {
  const int V = 1 << a.log2V;
  const int n = a.nx << a.log2V;
  const int rt = block % a.numRowTiles;
  const int zl = block / a.numRowTiles;
  const int y0 = rt * V;
  const int rows = (a.ny - y0) < V ? (a.ny - y0) : V;
  const size_t rowBase = (size_t)zl * a.ny + y0;
  cx<T>* A = smem;
  cx<T>* B = smem + n;
  if (rows < V) {
    SB_PHASE_BEGIN
    for (int i = tid; i < n; i += nthr) A[i] = mk<T>(0, 0);
    SB_PHASE_END
  }
  SB_PHASE_BEGIN
  const int total = rows * a.nx;
  if (a.r2c) {
    const T* in = static_cast<const T*>(a.spaceIn);
    for (int i = tid; i < total; i += nthr) {
      const int r = i / a.nx;
      const int x = i - r * a.nx;
      A[at<true>(x, r, a.log2V)] = mk<T>(in[(rowBase + r) * a.nx + x], T(0));
    }
  } else {
    const cx<T>* in = static_cast<const cx<T>*>(a.spaceIn);
    for (int i = tid; i < total; i += nthr) {
      const int r = i / a.nx;
      const int x = i - r * a.nx;
      A[at<true>(x, r, a.log2V)] = in[(rowBase + r) * a.nx + x];
    }
  }
  SB_PHASE_END
  cx<T>* R = tile_fft<T, false, true>(A, B, a.rp, a.log2V, a.tw, ctx);
  SB_PHASE_BEGIN
  const int total = rows * a.nxf;
  for (int i = tid; i < total; i += nthr) {
    const int r = i / a.nxf;
    const int x = i - r * a.nxf;
    a.planes[(rowBase + r) * a.nxf + x] = R[at<true>(x, r, a.log2V)];
  }
  SB_PHASE_END
}

}  // namespace sb
