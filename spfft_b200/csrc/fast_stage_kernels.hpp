// fast_stage_kernels.hpp -- stage kernel bodies for power-of-two transform lengths, built on the
// register-resident FFT of fast_fft.hpp. Same global layouts, argument structs and semantics as the
// generic bodies in stage_kernels.hpp (which remain the path for every other length):
//
//   z / y stages ("row" thread mapping: lane = tid % V fastest, so every global access of a
//   quarter warp is one 128-byte row segment of the plane-major stick buffer / the xy planes):
//       sparse side  <->  swizzled shared tile  <->  registers  <->  dense side in global memory
//   x stage ("column" mapping: consecutive threads walk along x, rows are contiguous in global
//   memory): global -> registers -> 2 exchanges -> registers -> global, no transposition at all.
//
// One CTA = one tile of V = 8 lanes (128-byte tile rows in double, 64-byte in float), V*N/8 threads,
// ONE tile buffer of N*V complex values in shared memory.
#pragma once
#include "fast_fft.hpp"
#include "stage_kernels.hpp"

namespace sb {

// lanes per tile row: 8 for both precisions = 128-byte rows (double) / 64-byte rows (float).
// For float, 16 lanes (128-byte rows) means 1024-thread CTAs at one CTA per SM; 8 lanes keep the
// two-CTAs-per-SM structure of the double kernels and measured 15 % faster (316 -> 367 pairs/s at
// 512^3, profiles/r01_v3_float_lanes.log). SB_LOG2V_F64 / SB_LOG2V_F32: build-time overrides.
#ifndef SB_LOG2V_F64
#define SB_LOG2V_F64 3
#endif
#ifndef SB_LOG2V_F32
#define SB_LOG2V_F32 3
#endif
template <typename T>
struct FastLanes {
  static constexpr int log2V = sizeof(T) == 8 ? SB_LOG2V_F64 : SB_LOG2V_F32;
};
// Rows per tile of the stand-alone x stage kernels. Rows are contiguous in global memory, so the
// lane count does not touch coalescing there: it only sets the CTA size (V * threads-per-row
// threads, N*V elements of shared memory). Small CTAs = many tiles per SM at different phases:
// measured at 512^3 double (profiles/r01_v4_summary.md) 8 / 4 / 2 / 1 rows = 0.690 / 0.661 / 0.626 /
// 0.620 ms, at 384^3 0.301 / 0.290 / 0.282 / 0.431 ms -- best around 64-128 threads per CTA, so
// V = 128 / (N/8) for the power-of-two kernels and 64 / (N/24) for the 3*2^k ones, within 1..8.
// SB_LOG2VX: build-time override (experiments).
constexpr int x_lanes_log2(int n) {
  const int target = n % 5 == 0 ? 2560 / n : (n % 3 == 0 ? 1536 / n : 1024 / n);  // rows for the CTA size above
  return target >= 8 ? 3 : (target >= 4 ? 2 : (target >= 2 ? 1 : 0));
}
template <typename T, int N>
struct FastLanesX {
#ifdef SB_LOG2VX
  static constexpr int log2V = SB_LOG2VX;
#else
  static constexpr int log2V = x_lanes_log2(N);
#endif
};

// Cache policy of the dense-side global accesses of the stand-alone stage kernels (experiments:
// -DSB_MEM_YB_ST=Mem::Stream ...). Every stage streams 3-4 GB through the 126 MB L2 once.
#ifndef SB_MEM_YB_ST
#define SB_MEM_YB_ST Mem::Plain
#endif
#ifndef SB_MEM_YF_LD
#define SB_MEM_YF_LD Mem::Plain
#endif
#ifndef SB_MEM_X_LD
#define SB_MEM_X_LD Mem::Plain
#endif
#ifndef SB_MEM_X_ST
#define SB_MEM_X_ST Mem::Plain
#endif

// gather-form tiles (defined below)
// (W = element type of the stick buffer: cx<T>, or cx<float> for the single-precision wire format of a
// distributed double-precision transform; deduced from the row pointer)
template <typename T, int N, Mem STP, bool TWS = false, typename W = cx<T>>
SB_DEV void y_backward_gather(const YArgs<T>& a, int xt, const W* stickRow, cx<T>* plane,
                              int nextXt, const W* nextStickRow, Ctx ctx, cx<T>* S,
                              const cx<T>* tw = nullptr);
template <typename T, int N, Mem LDP, bool TWS = false, typename W = cx<T>>
SB_DEV void y_forward_gather(const YArgs<T>& a, int xt, const cx<T>* plane, W* stickRow,
                             int nextXt, const cx<T>* nextPlane, Ctx ctx, cx<T>* S,
                             const cx<T>* tw = nullptr);

// Stages 0 .. last-1 with their exchanges; on return the tile holds the input of the last stage.
// `vAll`: the caller's SB_REGS array. COL0 / COL select the thread mapping (column: consecutive
// threads along n; row: lanes fastest) of stage 0 with its exchange write / of all later phases:
// a change of mapping across an exchange is free, which is how the sparse-side kernels turn
// "contiguous along n" (sparse values) into "contiguous along lanes" (plane-major rows).
template <typename T, int N, int LOG2V, bool BWD, typename Swz, bool COL0, bool COL = COL0, bool TWS = false>
SB_DEV void fast_fft_head(cx<T>* vAll, cx<T>* S, const cx<T>* __restrict__ tw, Ctx ctx) {
  (void)ctx;
  using P = FastPlan<N>;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = P::T;
#define SB_FAST_IDS_(C)                               \
  cx<T>* v = SB_RP(vAll, 8);                          \
  const int lane = C ? tid / TT : (tid & (V - 1));    \
  const int j = C ? (tid & (TT - 1)) : (tid >> LOG2V); \
  (void)nthr;
#define SB_FAST_IDS SB_FAST_IDS_(COL)
#define SB_FAST_WRITE(STAGE)                                                          \
  {                                                                                   \
    constexpr int R = P::radix(STAGE);                                                \
    constexpr int M = 8 / R;                                                          \
    _Pragma("unroll") for (int i = 0; i < M; ++i) {                                   \
      _Pragma("unroll") for (int q = 0; q < R; ++q)                                   \
          S[Swz::template at<LOG2V>(fast_out_index<N, STAGE>(j, i, q), lane)] = v[i + M * q]; \
    }                                                                                 \
  }
#define SB_FAST_READ \
  _Pragma("unroll") for (int m = 0; m < 8; ++m) v[m] = S[Swz::template at<LOG2V>(j + TT * m, lane)];

  SB_MARK(ctx, 2);
  run_item_chores(ctx);
  if constexpr (P::numStages > 1) {
    SB_PHASE_BEGIN
    SB_FAST_IDS_(COL0)
    fast_stage<T, N, BWD, 0, TWS>(v, j, tw);
    SB_FAST_WRITE(0)
    SB_PHASE_END
  }
  SB_MARK(ctx, 3);
  if constexpr (P::numStages > 2) {
    SB_PHASE_BEGIN
    SB_FAST_IDS
    SB_FAST_READ
    fast_stage<T, N, BWD, 1, TWS>(v, j, tw);
    SB_PHASE_END
    SB_MARK(ctx, 4);
    SB_PHASE_BEGIN
    SB_FAST_IDS
    SB_FAST_WRITE(1)
    SB_PHASE_END
    SB_MARK(ctx, 5);
  }
  if constexpr (P::numStages > 3) {
    SB_PHASE_BEGIN
    SB_FAST_IDS
    SB_FAST_READ
    fast_stage<T, N, BWD, 2, TWS>(v, j, tw);
    SB_PHASE_END
    SB_PHASE_BEGIN
    SB_FAST_IDS
    SB_FAST_WRITE(2)
    SB_PHASE_END
  }
  static_assert(P::numStages <= 4, "N <= 4096");
}

// Last stage, to be called inside the caller's final phase: afterwards v[m] = X[j + T*m].
// The two halves exist separately for the persistent kernels, which put a barrier between the last
// read of the tile buffer and the rest (warps then run ahead into the next item without another
// CTA-wide barrier).
template <typename T, int N, int LOG2V, typename Swz>
SB_DEV void fast_fft_tail_read(cx<T>* v, const cx<T>* S, int j, int lane) {
  using P = FastPlan<N>;
  constexpr int TT = P::T;
  if constexpr (P::numStages > 1) {
    SB_FAST_READ
  }
}
template <typename T, int N, int LOG2V, bool BWD, typename Swz, bool TWS = false>
SB_DEV void fast_fft_tail(cx<T>* v, const cx<T>* S, const cx<T>* __restrict__ tw, int j, int lane) {
  using P = FastPlan<N>;
  fast_fft_tail_read<T, N, LOG2V, Swz>(v, S, j, lane);
  fast_stage<T, N, BWD, P::numStages - 1, TWS>(v, j, tw);
}

// Hermitian completion of one lane of a swizzled tile, low index first (same semantics as
// hermitian_fill_lane in stage_kernels.hpp; reference src/symmetry/symmetry_host.hpp:47-58,73-90).
template <typename T, int LOG2V, typename Swz>
SB_DEV void hermitian_fill_lane_swz(cx<T>* A, int n, int lane, Ctx ctx) {
  (void)ctx;
  const int half = n / 2;
  SB_PHASE_BEGIN
  for (int i = 1 + tid; i <= half; i += nthr) {
    const cx<T> val = A[Swz::template at<LOG2V>(i, lane)];
    if (nonzero(val)) A[Swz::template at<LOG2V>(n - i, lane)] = conj(val);
  }
  SB_PHASE_END
  SB_PHASE_BEGIN
  for (int i = half + 1 + tid; i < n; i += nthr) {
    const cx<T> val = A[Swz::template at<LOG2V>(i, lane)];
    if (nonzero(val)) A[Swz::template at<LOG2V>(n - i, lane)] = conj(val);
  }
  SB_PHASE_END
}

#define SB_ROW_IDS                    \
  cx<T>* v = SB_RP(vAll, 8);          \
  const int lane = tid & (V - 1);     \
  const int j = tid >> LOG2V;         \
  (void)nthr;

// -------------------------------------------------------------------------------------------
// z stage
// -------------------------------------------------------------------------------------------
template <typename T, int N, typename W = cx<T>>
SB_DEV void z_backward_fast(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  SB_REGS(cx<T>, vAll, 8);
  const int e0 = a.tileStart[tile], e1 = a.tileStart[tile + 1];  // in flight during the zero fill
  SB_PHASE_BEGIN
  if (a.pfDist > 0 && tile + a.pfDist < a.numTiles) {
    const int p0 = a.tileStart[tile + a.pfDist], p1 = a.tileStart[tile + a.pfDist + 1];
    prefetch_l2(a.entrySlot + p0, (size_t)(p1 - p0) * sizeof(int), tid, nthr);
    if (!a.entrySrc) prefetch_l2(a.valuesIn + p0, (size_t)(p1 - p0) * sizeof(cx<T>), tid, nthr);
  }
  for (int i = tid; i < N * V; i += nthr) S[i] = mk<T>(0, 0);
  SB_PHASE_END
  SB_PHASE_BEGIN
  // all loads of a batch are issued before the first shared-memory store (one DRAM round trip
  // per batch instead of one per entry)
  constexpr int U = 6;
  for (int base = e0 + tid; base < e1; base += U * nthr) {
    int slot[U];
    cx<T> val[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = base + u * nthr;
      if (e < e1) {
        slot[u] = a.entrySlot[e];
        val[u] = a.valuesIn[a.entrySrc ? a.entrySrc[e] : e];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * nthr < e1) S[SwzRow::template at<LOG2V>(slot[u] >> LOG2V, slot[u] & (V - 1))] = val[u];
    }
  }
  SB_PHASE_END
  if (tile == a.symTile) hermitian_fill_lane_swz<T, LOG2V, SwzRow>(S, N, a.symLane, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = S[SwzRow::template at<LOG2V>(j + TT * m, lane)];
  SB_PHASE_END
  fast_fft_head<T, N, LOG2V, true, SwzRow, false>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
  fast_fft_tail<T, N, LOG2V, true, SwzRow>(v, S, a.ftw, j, lane);
  const size_t col = (size_t)tile * V + lane;
#pragma unroll
  for (int m = 0; m < 8; ++m) z_row<T, W>(a, j + TT * m)[col] = to_wire<W>(v[m]);
  SB_PHASE_END_NOSYNC
}

template <typename T, int N, typename W = cx<T>>
SB_DEV void z_forward_fast(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  SB_REGS(cx<T>, vAll, 8);
  const int e0 = a.tileStart[tile], e1 = a.tileStart[tile + 1];
  SB_PHASE_BEGIN
  SB_ROW_IDS
  const W* in = reinterpret_cast<const W*>(a.sticks) + (size_t)tile * V + lane;
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = from_wire<T>(in[(size_t)(j + TT * m) * a.pitch]);
  if (a.pfDist > 0 && tile + a.pfDist < a.numTiles) {
    for (int r = tid; r < N; r += nthr)
      prefetch_l2_line(reinterpret_cast<const W*>(a.sticks) + (size_t)(tile + a.pfDist) * V + (size_t)r * a.pitch);
    const int p0 = a.tileStart[tile + a.pfDist], p1 = a.tileStart[tile + a.pfDist + 1];
    prefetch_l2(a.entrySlot + p0, (size_t)(p1 - p0) * sizeof(int), tid, nthr);
  }
  SB_PHASE_END
  fast_fft_head<T, N, LOG2V, false, SwzRow, false>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
  fast_fft_tail<T, N, LOG2V, false, SwzRow>(v, S, a.ftw, j, lane);
  SB_PHASE_END  // every thread has read its inputs of the last stage
  SB_PHASE_BEGIN
  SB_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m) S[SwzRow::template at<LOG2V>(j + TT * m, lane)] = v[m];
  SB_PHASE_END
  SB_PHASE_BEGIN
  constexpr int U = 6;
  for (int base = e0 + tid; base < e1; base += U * nthr) {
    int slot[U], dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = base + u * nthr;
      if (e < e1) {
        slot[u] = a.entrySlot[e];
        dst[u] = a.entrySrc ? a.entrySrc[e] : e;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * nthr < e1) {
        cx<T> val = S[SwzRow::template at<LOG2V>(slot[u] >> LOG2V, slot[u] & (V - 1))];
        if (a.useScale) val = a.scale * val;
        a.valuesOut[dst[u]] = val;
      }
    }
  }
  SB_PHASE_END_NOSYNC
}

// -------------------------------------------------------------------------------------------
// y stage: one tile = V consecutive x columns of one plane.
//   stickRow : row of the plane-major stick buffer that holds this plane (all local sticks)
//   plane    : the plane's [ny][nxf] array (the plane buffer, or a slot of the L2 scratch ring)
// -------------------------------------------------------------------------------------------
template <typename T, int N, Mem LDS, Mem STP, typename W = cx<T>>
SB_DEV void y_backward_tile(const YArgs<T>& a, int xt, const W* stickRow, cx<T>* plane,
                            int nextXt, const W* nextStickRow, Ctx ctx, cx<T>* S, int zl = 0) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  cx<T>* planeTile = plane + (size_t)xt * V;
  const int lanesValid = (a.nxf - xt * V) < V ? (a.nxf - xt * V) : V;
  if (e0 == e1) {
    // empty x tile: the x stage still reads these columns -> store zeros, no transform
    SB_PHASE_BEGIN
    for (int i = tid; i < N * V; i += nthr) {
      const int y = i >> LOG2V;
      const int lane = i & (V - 1);
      if (lane < lanesValid) st_g<STP>(planeTile + (size_t)y * a.nxf + lane, mk<T>(0, 0));
    }
    SB_PHASE_END_NOSYNC
    return;
  }
  SB_REGS(cx<T>, vAll, 8);
  SB_PHASE_BEGIN
  if (nextXt >= 0) {
    const int p0 = a.xtStart[nextXt], p1 = a.xtStart[nextXt + 1];
    prefetch_l2(nextStickRow + p0, (size_t)(p1 - p0) * sizeof(W), tid, nthr);
  }
  for (int i = tid; i < N * V; i += nthr) S[i] = mk<T>(0, 0);
  SB_PHASE_END
  SB_PHASE_BEGIN
  constexpr int U = 6;
  for (int base = e0 + tid; base < e1; base += U * nthr) {
    int slot[U];
    cx<T> val[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = base + u * nthr;
      if (e < e1) {
        slot[u] = a.stickSlot[e];
        val[u] = from_wire<T>(a.srcBase ? reinterpret_cast<const W*>(a.sticks)[(size_t)a.srcBase[e] + (size_t)zl * a.srcPitch[e]]
                                        : ld_g<LDS>(stickRow + e));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * nthr < e1) S[SwzRow::template at<LOG2V>(slot[u] >> LOG2V, slot[u] & (V - 1))] = val[u];
    }
  }
  SB_PHASE_END
  if (a.symmetry && xt == 0) hermitian_fill_lane_swz<T, LOG2V, SwzRow>(S, N, 0, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = S[SwzRow::template at<LOG2V>(j + TT * m, lane)];
  SB_PHASE_END
  fast_fft_head<T, N, LOG2V, true, SwzRow, false>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
  fast_fft_tail<T, N, LOG2V, true, SwzRow>(v, S, a.ftw, j, lane);
  if (lane < lanesValid) {
#pragma unroll
    for (int m = 0; m < 8; ++m) st_g<STP>(planeTile + (size_t)(j + TT * m) * a.nxf + lane, v[m]);
  }
  SB_PHASE_END_NOSYNC
}

template <typename T, int N, Mem LDP, Mem STS, typename W = cx<T>>
SB_DEV void y_forward_tile(const YArgs<T>& a, int xt, const cx<T>* plane, W* stickRow,
                           int nextXt, const cx<T>* nextPlane, Ctx ctx, cx<T>* S, int zl = 0) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  if (e0 == e1) return;  // no stick needs these columns
  const cx<T>* planeTile = plane + (size_t)xt * V;
  const int lanesValid = (a.nxf - xt * V) < V ? (a.nxf - xt * V) : V;
  SB_REGS(cx<T>, vAll, 8);
  SB_PHASE_BEGIN
  SB_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m)
    v[m] = lane < lanesValid ? ld_g<LDP>(planeTile + (size_t)(j + TT * m) * a.nxf + lane) : mk<T>(0, 0);
  if (nextXt >= 0) {
    for (int r = tid; r < N; r += nthr) prefetch_l2_line(nextPlane + (size_t)nextXt * V + (size_t)r * a.nxf);
  }
  SB_PHASE_END
  fast_fft_head<T, N, LOG2V, false, SwzRow, false>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
  fast_fft_tail<T, N, LOG2V, false, SwzRow>(v, S, a.ftw, j, lane);
  SB_PHASE_END
  SB_PHASE_BEGIN
  SB_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m) S[SwzRow::template at<LOG2V>(j + TT * m, lane)] = v[m];
  SB_PHASE_END
  SB_PHASE_BEGIN
  constexpr int U = 6;
  for (int base = e0 + tid; base < e1; base += U * nthr) {
    int slot[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * nthr < e1) slot[u] = a.stickSlot[base + u * nthr];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * nthr < e1) {
        const int e = base + u * nthr;
        W* dst = a.srcBase ? y_dist_stick<T, true, W>(a, e, zl) : stickRow + e;
        st_g<STS>(dst, to_wire<W>(S[SwzRow::template at<LOG2V>(slot[u] >> LOG2V, slot[u] & (V - 1))]));
      }
    }
  }
  SB_PHASE_END_NOSYNC
}

template <typename T, int N, typename W>
SB_DEV void y_backward_fast_w(const YArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  const int xt = block % a.numXTiles;
  const int zl = block / a.numXTiles;
  const W* sticks = reinterpret_cast<const W*>(a.sticks);
  int nextXt = -1;
  const W* nextRow = nullptr;
  if (a.pfDist > 0 && !a.srcBase && block + a.pfDist < a.numXTiles * a.numPlanes) {
    nextXt = (block + a.pfDist) % a.numXTiles;
    nextRow = sticks + (size_t)((block + a.pfDist) / a.numXTiles + a.zRowOffset) * a.pitch;
  }
  if (a.srcBase && a.inv && a.tilePitch[xt] != 0) {
    // distributed, all sticks of this tile from one rank: contiguous inside that rank's block.
    // The gather form indexes relative to the tile's first stick, so pass row - xtStart[xt].
    const W* row = sticks + (size_t)a.tileBase[xt] + (size_t)zl * a.tilePitch[xt] - a.xtStart[xt];
    y_backward_gather<T, N, SB_MEM_YB_ST>(a, xt, row, a.planes + (size_t)zl * N * a.nxf, -1, (const W*)nullptr, ctx, S);
  } else if (a.inv && !a.srcBase)
    y_backward_gather<T, N, SB_MEM_YB_ST>(a, xt, sticks + (size_t)(zl + a.zRowOffset) * a.pitch,
                                        a.planes + (size_t)zl * N * a.nxf, nextXt, nextRow, ctx, S);
  else
    y_backward_tile<T, N, Mem::Plain, SB_MEM_YB_ST>(a, xt, sticks + (size_t)(zl + a.zRowOffset) * a.pitch,
                                                  a.planes + (size_t)zl * N * a.nxf, nextXt, nextRow, ctx, S, zl);
}
template <typename T, int N, bool WIRE = false>
SB_DEV void y_backward_fast(const YArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  y_backward_fast_w<T, N, WireElem<T, WIRE>>(a, block, ctx, S);
}

template <typename T, int N, typename W>
SB_DEV void y_forward_fast_w(const YArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  const int xt = y_forward_tile_at<T>(a, block % a.numXTiles);
  const int zl = block / a.numXTiles;
  W* sticks = reinterpret_cast<W*>(a.sticks);
  int nextXt = -1;
  const cx<T>* nextPlane = nullptr;
  if (a.pfDist > 0 && block + a.pfDist < a.numXTiles * a.numPlanes) {
    nextXt = y_forward_tile_at<T>(a, (block + a.pfDist) % a.numXTiles);
    nextPlane = a.planes + (size_t)((block + a.pfDist) / a.numXTiles) * N * a.nxf;
  }
  if (a.srcBase && a.inv && a.tilePitch[xt] != 0) {
    W* row = y_dist_tile<T, true, W>(a, xt, zl) - a.xtStart[xt];
    y_forward_gather<T, N, SB_MEM_YF_LD>(a, xt, a.planes + (size_t)zl * N * a.nxf, row, nextXt, nextPlane, ctx, S);
  } else if (a.inv && !a.srcBase)
    y_forward_gather<T, N, SB_MEM_YF_LD>(a, xt, a.planes + (size_t)zl * N * a.nxf,
                                       sticks + (size_t)(zl + a.zRowOffset) * a.pitch, nextXt, nextPlane,
                                       ctx, S);
  else
    y_forward_tile<T, N, SB_MEM_YF_LD, Mem::Plain>(a, xt, a.planes + (size_t)zl * N * a.nxf,
                                                 sticks + (size_t)(zl + a.zRowOffset) * a.pitch, nextXt,
                                                 nextPlane, ctx, S, zl);
}
template <typename T, int N, bool WIRE = false>
SB_DEV void y_forward_fast(const YArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  y_forward_fast_w<T, N, WireElem<T, WIRE>>(a, block, ctx, S);
}


// -------------------------------------------------------------------------------------------
// Sparse side through an inverse map ("gather" form). Instead of zero-filling a shared tile,
// scattering the sparse entries into it and reading it back, every thread looks up where (if
// anywhere) each of its 8 elements lives in the sparse array and loads it straight into registers.
// The thread mapping of this first (backward) or last (forward) stage is "column" (consecutive
// threads walk along n, i.e. along consecutive sparse entries of one stick / one x column ->
// coalesced), all other stages and the dense side use the "row" mapping; the change of mapping
// rides on an exchange that is needed anyway. Shared-memory traffic per tile: the two exchanges
// only (as in the x stage) -- no zero fill, no scatter, no read-back.
//
// inv[((tile * threads) + tid) * 8 + m], tid = lane * T + j: offset of element n = j + T*m of the
// tile's lane-th sequence, relative to the tile's first sparse entry; 0xFFFF = not present (zero).
// -------------------------------------------------------------------------------------------
constexpr unsigned short kNoEntry = 0xFFFF;

struct Inv8 {
  unsigned short i[8];
};

SB_DEV Inv8 load_inv8(const unsigned short* p) {
  Inv8 r;
#if SB_ON_GPU
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  r.i[0] = q.x & 0xFFFF; r.i[1] = q.x >> 16;
  r.i[2] = q.y & 0xFFFF; r.i[3] = q.y >> 16;
  r.i[4] = q.z & 0xFFFF; r.i[5] = q.z >> 16;
  r.i[6] = q.w & 0xFFFF; r.i[7] = q.w >> 16;
#else
  for (int m = 0; m < 8; ++m) r.i[m] = p[m];
#endif
  return r;
}

// Hermitian completion in gather form (same result as the in-place low-index-first fill of the
// reference, src/symmetry/symmetry_host.hpp:47-58,73-90): element n of a sequence of length N,
// p = given value at n, q = given value at N-n (zero if absent).
template <typename T>
SB_DEV cx<T> hermitian_combine(int n, int N, cx<T> p, cx<T> q) {
  if (n == 0) return p;
  if (2 * n == N) return conj(p);
  if (2 * n < N) return nonzero(p) ? p : conj(q);
  return nonzero(q) ? conj(q) : p;
}

// Loads the 8 elements of thread (lane, j) (column mapping) from `sparse` through the inverse map
// of this tile (`invTile` = first entry of the tile's map). hermitianLane: complete that lane.
template <typename T, int N, typename W>
SB_DEV void gather_load(cx<T>* v, const W* sparse, const unsigned short* invTile, int tid, int j,
                        int lane, int hermitianLane) {
  constexpr int TT = FastPlan<N>::T;
  const Inv8 iv = load_inv8(invTile + (size_t)tid * 8);
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = iv.i[m] != kNoEntry ? from_wire<T>(sparse[iv.i[m]]) : mk<T>(0, 0);
  if (lane == hermitianLane) {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int n = j + TT * m;
      const int n2 = (N - n) & (N - 1);
      const unsigned short i2 = invTile[((size_t)lane * TT + (n2 & (TT - 1))) * 8 + n2 / TT];
      const cx<T> q = i2 != kNoEntry ? from_wire<T>(sparse[i2]) : mk<T>(0, 0);
      v[m] = hermitian_combine<T>(n, N, v[m], q);
    }
  }
}

#define SB_COLMAP_IDS                 \
  cx<T>* v = SB_RP(vAll, 8);          \
  const int lane = tid / TT;          \
  const int j = tid & (TT - 1);       \
  (void)nthr;

template <typename T, int N, typename W = cx<T>>
SB_DEV void z_backward_gather(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  constexpr int THREADS = V * TT;
  SB_REGS(cx<T>, vAll, 8);
  const int e0 = a.tileStart[tile];
  SB_PHASE_BEGIN
  SB_COLMAP_IDS
  gather_load<T, N>(v, a.valuesIn + e0, a.inv + (size_t)tile * THREADS * 8, tid, j, lane,
                    tile == a.symTile ? a.symLane : -1);
  if (a.pfDist > 0 && tile + a.pfDist < a.numTiles) {
    const int p0 = a.tileStart[tile + a.pfDist], p1 = a.tileStart[tile + a.pfDist + 1];
    prefetch_l2(a.valuesIn + p0, (size_t)(p1 - p0) * sizeof(cx<T>), tid, nthr);
    prefetch_l2(a.inv + (size_t)(tile + a.pfDist) * THREADS * 8, (size_t)THREADS * 16, tid, nthr);
  }
  SB_PHASE_END_NOSYNC
  fast_fft_head<T, N, LOG2V, true, SwzCol, true, false>(vAll, S, a.ftw, ctx);
#if SB_ON_GPU && SB_BULK_STORE
  {
    W* R = reinterpret_cast<W*>(S);  // the finished tile, natural layout [row][lane]
    SB_PHASE_BEGIN
    SB_ROW_IDS
    fast_fft_tail<T, N, LOG2V, true, SwzCol>(v, S, a.ftw, j, lane);
    SB_PHASE_END  // every thread has read its inputs of the last stage
    SB_PHASE_BEGIN
    SB_ROW_IDS
#pragma unroll
    for (int m = 0; m < 8; ++m) R[((j + TT * m) << LOG2V) + lane] = to_wire<W>(v[m]);
    bulk_store_fence();
    SB_PHASE_END
    SB_PHASE_BEGIN
    for (int r = tid; r < N; r += nthr)
      bulk_store_row(z_row<T, W>(a, r) + (size_t)tile * V, R + ((size_t)r << LOG2V), V * sizeof(W));
    bulk_store_commit_wait();
    SB_PHASE_END_NOSYNC
    return;
  }
#endif
  SB_PHASE_BEGIN
  SB_ROW_IDS
  fast_fft_tail<T, N, LOG2V, true, SwzCol>(v, S, a.ftw, j, lane);
  const size_t col = (size_t)tile * V + lane;
#pragma unroll
  for (int m = 0; m < 8; ++m) z_row<T, W>(a, j + TT * m)[col] = to_wire<W>(v[m]);
  SB_PHASE_END_NOSYNC
}

template <typename T, int N, typename W = cx<T>>
SB_DEV void z_forward_gather(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  constexpr int THREADS = V * TT;
  SB_REGS(cx<T>, vAll, 8);
  const int e0 = a.tileStart[tile];
  SB_PHASE_BEGIN
  SB_ROW_IDS
  const W* in = reinterpret_cast<const W*>(a.sticks) + (size_t)tile * V + lane;
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = from_wire<T>(in[(size_t)(j + TT * m) * a.pitch]);
  if (a.pfDist > 0 && tile + a.pfDist < a.numTiles) {
    for (int r = tid; r < N; r += nthr)
      prefetch_l2_line(reinterpret_cast<const W*>(a.sticks) + (size_t)(tile + a.pfDist) * V + (size_t)r * a.pitch);
    prefetch_l2(a.inv + (size_t)(tile + a.pfDist) * THREADS * 8, (size_t)THREADS * 16, tid, nthr);
  }
  SB_PHASE_END_NOSYNC
  fast_fft_head<T, N, LOG2V, false, SwzCol, false, false>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  SB_COLMAP_IDS
  fast_fft_tail<T, N, LOG2V, false, SwzCol>(v, S, a.ftw, j, lane);
  const Inv8 iv = load_inv8(a.inv + ((size_t)tile * THREADS + tid) * 8);
  cx<T>* out = a.valuesOut + e0;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    if (iv.i[m] != kNoEntry) out[iv.i[m]] = a.useScale ? a.scale * v[m] : v[m];
  }
  SB_PHASE_END_NOSYNC
}

// z stage entry: inverse-map (gather) form when the values are in stick order, scatter form otherwise
template <typename T, int N, bool FWD, bool WIRE = false>
SB_DEV void z_fast_any(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* S) {
  using W = WireElem<T, WIRE>;
  if (a.inv) {
    if (FWD)
      z_forward_gather<T, N, W>(a, tile, ctx, S);
    else
      z_backward_gather<T, N, W>(a, tile, ctx, S);
  } else {
    if (FWD)
      z_forward_fast<T, N, W>(a, tile, ctx, S);
    else
      z_backward_fast<T, N, W>(a, tile, ctx, S);
  }
}

template <typename T, int N, Mem STP, bool TWS, typename W>
SB_DEV void y_backward_gather(const YArgs<T>& a, int xt, const W* stickRow, cx<T>* plane,
                              int nextXt, const W* nextStickRow, Ctx ctx, cx<T>* S,
                              const cx<T>* tw) {
  if (!TWS) tw = a.ftw;
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  constexpr int THREADS = V * TT;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  cx<T>* planeTile = plane + (size_t)xt * V;
  const int lanesValid = (a.nxf - xt * V) < V ? (a.nxf - xt * V) : V;
  SB_REGS(cx<T>, vAll, 8);
  if (e0 == e1) {
    // empty x tile: the x stage still reads these columns -> store zeros, no transform
    // (barrier: persistent kernels publish thread 0's bookkeeping through it)
    run_item_chores(ctx);
    SB_PHASE_BEGIN
    SB_PHASE_END
    SB_PHASE_BEGIN
    for (int i = tid; i < N * V; i += nthr) {
      const int y = i >> LOG2V;
      const int lane = i & (V - 1);
      if (lane < lanesValid) st_g<STP>(planeTile + (size_t)y * a.nxf + lane, mk<T>(0, 0));
    }
    SB_PHASE_END_NOSYNC
    return;
  }
  SB_PHASE_BEGIN
  SB_COLMAP_IDS
  gather_load<T, N>(v, stickRow + e0, a.inv + (size_t)xt * THREADS * 8, tid, j, lane,
                    (a.symmetry && xt == 0) ? 0 : -1);
  if (nextXt >= 0) {
    const int p0 = a.xtStart[nextXt], p1 = a.xtStart[nextXt + 1];
    prefetch_l2(nextStickRow + p0, (size_t)(p1 - p0) * sizeof(W), tid, nthr);
  }
  SB_PHASE_END_NOSYNC
  fast_fft_head<T, N, LOG2V, true, SwzCol, true, false, TWS>(vAll, S, tw, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
  fast_fft_tail_read<T, N, LOG2V, SwzCol>(v, S, j, lane);
  SB_PHASE_END_IF(TWS)
  SB_MARK(ctx, 8);
  SB_PHASE_BEGIN
  SB_ROW_IDS
  fast_stage<T, N, true, FastPlan<N>::numStages - 1, TWS>(v, j, tw);
#if SB_ON_GPU && SB_BULK_STORE
  if (!TWS && ((a.nxf * sizeof(cx<T>)) & 15) == 0) {
    // (every thread is past its last read of the tile buffer only after a barrier)
    group_sync(ctx);
#pragma unroll
    for (int m = 0; m < 8; ++m) S[((j + TT * m) << LOG2V) + lane] = v[m];
    bulk_store_fence();
    group_sync(ctx);
    for (int r = tid; r < N; r += nthr)
      bulk_store_row(planeTile + (size_t)r * a.nxf, S + ((size_t)r << LOG2V), lanesValid * sizeof(cx<T>));
    bulk_store_commit_wait();
  } else
#endif
  if (lane < lanesValid) {
#pragma unroll
    for (int m = 0; m < 8; ++m) st_g<STP>(planeTile + (size_t)(j + TT * m) * a.nxf + lane, v[m]);
  }
  SB_PHASE_END_NOSYNC
}

template <typename T, int N, Mem LDP, bool TWS, typename W>
SB_DEV void y_forward_gather(const YArgs<T>& a, int xt, const cx<T>* plane, W* stickRow,
                             int nextXt, const cx<T>* nextPlane, Ctx ctx, cx<T>* S,
                             const cx<T>* tw) {
  if (!TWS) tw = a.ftw;
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  constexpr int THREADS = V * TT;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  if (e0 == e1) {  // no stick needs these columns
    // (persistent kernel: every item passes at least one barrier)
    run_item_chores(ctx);
    SB_PHASE_BEGIN
    (void)tid;
    (void)nthr;
    SB_PHASE_END_IF(TWS)
    return;
  }
  const cx<T>* planeTile = plane + (size_t)xt * V;
  const int lanesValid = (a.nxf - xt * V) < V ? (a.nxf - xt * V) : V;
  SB_REGS(cx<T>, vAll, 8);
  SB_PHASE_BEGIN
  SB_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m)
    v[m] = lane < lanesValid ? ld_g<LDP>(planeTile + (size_t)(j + TT * m) * a.nxf + lane) : mk<T>(0, 0);
  if (nextXt >= 0) {
    for (int r = tid; r < N; r += nthr) prefetch_l2_line(nextPlane + (size_t)nextXt * V + (size_t)r * a.nxf);
  }
  SB_PHASE_END_NOSYNC
  fast_fft_head<T, N, LOG2V, false, SwzCol, false, false, TWS>(vAll, S, tw, ctx);
  SB_PHASE_BEGIN
  SB_COLMAP_IDS
  fast_fft_tail_read<T, N, LOG2V, SwzCol>(v, S, j, lane);
  SB_PHASE_END_IF(TWS)
  SB_MARK(ctx, 8);
  SB_PHASE_BEGIN
  SB_COLMAP_IDS
  fast_stage<T, N, false, FastPlan<N>::numStages - 1, TWS>(v, j, tw);
  const Inv8 iv = load_inv8(a.inv + ((size_t)xt * THREADS + tid) * 8);
  W* out = stickRow + e0;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    if (iv.i[m] != kNoEntry) out[iv.i[m]] = to_wire<W>(v[m]);
  }
  SB_PHASE_END_NOSYNC
}
#undef SB_COLMAP_IDS

// -------------------------------------------------------------------------------------------
// x stage, complex rows (C2C). Tile = V consecutive rows y0 .. y0+V-1 of one plane;
// thread = (row lane, j). in / out: the plane's [ny][N] arrays (may be the same memory).
// -------------------------------------------------------------------------------------------
template <typename T, int N, bool BWD, Mem LD, Mem ST, bool TWS = false, int LOG2V = FastLanes<T>::log2V,
          typename Swz = SwzCol>
SB_DEV void x_c2c_tile(const cx<T>* in, cx<T>* out, int y0, int ny, const cx<T>* __restrict__ ftw,
                       const cx<T>* nextRows, Ctx ctx, cx<T>* S) {
  constexpr int TT = FastPlan<N>::T;
  SB_REGS(cx<T>, vAll, 8);
#define SB_COL_IDS                    \
  cx<T>* v = SB_RP(vAll, 8);          \
  const int lane = tid / TT;          \
  const int j = tid & (TT - 1);       \
  const bool valid = y0 + lane < ny;  \
  (void)nthr;
  SB_PHASE_BEGIN
  SB_COL_IDS
  const cx<T>* src = in + (size_t)(y0 + lane) * N + j;
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = valid ? ld_g<LD>(src + TT * m) : mk<T>(0, 0);
  if (nextRows) prefetch_l2(nextRows, sizeof(cx<T>) * N * (1 << LOG2V), tid, nthr);
  SB_PHASE_END_NOSYNC  // first use of the tile buffer is the exchange after stage 0
  fast_fft_head<T, N, LOG2V, BWD, Swz, true, true, TWS>(vAll, S, ftw, ctx);
  SB_PHASE_BEGIN
  SB_COL_IDS
  fast_fft_tail_read<T, N, LOG2V, Swz>(v, S, j, lane);
  SB_PHASE_END_IF(TWS)
  SB_MARK(ctx, 8);
  SB_PHASE_BEGIN
  SB_COL_IDS
  fast_stage<T, N, BWD, FastPlan<N>::numStages - 1, TWS>(v, j, ftw);
  if (valid) {
    cx<T>* dst = out + (size_t)(y0 + lane) * N + j;
#pragma unroll
    for (int m = 0; m < 8; ++m) st_g<ST>(dst + TT * m, v[m]);
  }
  SB_PHASE_END_NOSYNC
#undef SB_COL_IDS
}

// -------------------------------------------------------------------------------------------
// x stage, real rows (R2C / C2R): TWO real rows per complex register FFT. A lane of the tile holds
// the rows yA = y0 + 2*lane and yB = yA + 1 as z = a + i*b (the tile covers 2*V rows).
//   backward (C2R): the half spectra A, B (nxf = N/2+1 complex each) are completed by conjugation
//                   while loading and combined to Z = A + i*B; one backward FFT gives a = Re z,
//                   b = Im z (unpadded real rows of N). Like cuFFT's Z2D (reference:
//                   transform_real_2d_gpu.hpp:54-256) the imaginary parts of X[0] and X[N/2] do
//                   not enter the result.
//   forward (R2C) : z = a + i*b -> FFT -> A[k] = (Z[k] + conj(Z[N-k]))/2,
//                   B[k] = (Z[k] - conj(Z[N-k]))/(2i), k <= N/2; Z[N-k] comes from the tile buffer.
// Memory traffic is the minimum (half spectra + real rows) and a row costs half a complex FFT.
// -------------------------------------------------------------------------------------------
template <typename T>
SB_HD cx<T> pack_half_spectra(cx<T> A, cx<T> B, bool edge, bool mirrored) {
  if (edge) {  // x == 0 or x == N/2: real by symmetry
    A.y = T(0);
    B.y = T(0);
  }
  if (mirrored) {
    A = conj(A);
    B = conj(B);
  }
  return mk<T>(A.x - B.y, A.y + B.x);
}
// A[k], B[k] from Z[k] and Z[N-k]
template <typename T>
SB_HD void unpack_half_spectra(cx<T> zk, cx<T> zn, cx<T>& A, cx<T>& B) {
  const cx<T> c = conj(zn);
  A = T(0.5) * (zk + c);
  const cx<T> d = zk - c;
  B = mk<T>(T(0.5) * d.y, T(-0.5) * d.x);
}

template <typename T, int N, bool BWD>
SB_DEV void x_r2c_pair_tile(const XArgs<T>& a, size_t planeRow0, int y0, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanesX<T, N>::log2V;
  using SwzCol = SwzX<sizeof(cx<T>)>;  // (shadows the 8-lane fold: this kernel has its own lane count)
  constexpr int TT = FastPlan<N>::T;
  constexpr int NXF = N / 2 + 1;
  SB_REGS(cx<T>, vAll, 8);
#define SB_COL_IDS                         \
  cx<T>* v = SB_RP(vAll, 8);               \
  const int lane = tid / TT;               \
  const int j = tid & (TT - 1);            \
  const int yA = y0 + 2 * lane;            \
  const bool validA = yA < a.ny;           \
  const bool validB = yA + 1 < a.ny;       \
  (void)nthr;
  SB_PHASE_BEGIN
  SB_COL_IDS
  if (BWD) {
    const cx<T>* srcA = a.planes + (planeRow0 + yA) * NXF;
    const cx<T>* srcB = srcA + NXF;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int n = j + TT * m;
      const bool hi = n >= NXF;
      const int x = hi ? N - n : n;
      const cx<T> A = validA ? srcA[x] : mk<T>(0, 0);
      const cx<T> B = validB ? srcB[x] : mk<T>(0, 0);
      v[m] = pack_half_spectra<T>(A, B, x == 0 || 2 * x == N, hi);
    }
  } else {
    const T* srcA = static_cast<const T*>(a.spaceIn) + (planeRow0 + yA) * N;
    const T* srcB = srcA + N;
#pragma unroll
    for (int m = 0; m < 8; ++m)
      v[m] = mk<T>(validA ? srcA[j + TT * m] : T(0), validB ? srcB[j + TT * m] : T(0));
  }
  SB_PHASE_END_NOSYNC
  fast_fft_head<T, N, LOG2V, BWD, SwzCol, true, true>(vAll, S, a.ftw, ctx);
  if (BWD) {
    SB_PHASE_BEGIN
    SB_COL_IDS
    fast_fft_tail<T, N, LOG2V, BWD, SwzCol>(v, S, a.ftw, j, lane);
    T* dstA = static_cast<T*>(a.spaceOut) + (planeRow0 + yA) * N;
    T* dstB = dstA + N;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if (validA) dstA[j + TT * m] = v[m].x;
      if (validB) dstB[j + TT * m] = v[m].y;
    }
    SB_PHASE_END_NOSYNC
  } else {
    SB_PHASE_BEGIN
    SB_COL_IDS
    (void)validA;
    (void)validB;
    fast_fft_tail<T, N, LOG2V, BWD, SwzCol>(v, S, a.ftw, j, lane);
    SB_PHASE_END  // every thread has read its inputs of the last stage
    SB_PHASE_BEGIN
    SB_COL_IDS
    (void)validA;
    (void)validB;
#pragma unroll
    for (int m = 0; m < 8; ++m) S[SwzCol::template at<LOG2V>(j + TT * m, lane)] = v[m];
    SB_PHASE_END
    SB_PHASE_BEGIN
    SB_COL_IDS
    cx<T>* dstA = a.planes + (planeRow0 + yA) * NXF;
    cx<T>* dstB = dstA + NXF;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int k = j + TT * m;
      if (k < NXF) {
        cx<T> A, B;
        unpack_half_spectra<T>(v[m], S[SwzCol::template at<LOG2V>((N - k) & (N - 1), lane)], A, B);
        if (validA) dstA[k] = A;
        if (validB) dstB[k] = B;
      }
    }
    SB_PHASE_END_NOSYNC
  }
#undef SB_COL_IDS
}

template <typename T, int N, bool BWD>
SB_DEV void x_r2c_fast(const XArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  constexpr int V = 1 << FastLanesX<T, N>::log2V;
  const int rt = block % a.numRowTiles;  // numRowTiles = ceil(ny / 2V), stage_args.hpp
  const int zl = block / a.numRowTiles;
  x_r2c_pair_tile<T, N, BWD>(a, (size_t)zl * a.ny, rt * 2 * V, ctx, S);
}

template <typename T, int N, bool BWD>
SB_DEV void x_c2c_fast(const XArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  constexpr int V = 1 << FastLanesX<T, N>::log2V;
  const int rt = block % a.numRowTiles;
  const int zl = block / a.numRowTiles;
  const size_t planeOff = (size_t)zl * a.ny * N;
  const cx<T>* src = (BWD ? a.planes : static_cast<const cx<T>*>(a.spaceIn)) + planeOff;
  cx<T>* dst = (BWD ? static_cast<cx<T>*>(a.spaceOut) : a.planes) + planeOff;
  x_c2c_tile<T, N, BWD, SB_MEM_X_LD, SB_MEM_X_ST, false, FastLanesX<T, N>::log2V, SwzX<sizeof(cx<T>)>>(
      src, dst, rt * V, a.ny, a.ftw, nullptr, ctx, S);
}

// -------------------------------------------------------------------------------------------
// Fused xy stage (wfft_xy.cu): y tiles and x tiles of every plane as items of ONE persistent kernel, the
// y<->x hand-off going through a small ring of scratch planes that stays resident in the 126 MB L2
// instead of a full-size plane buffer in HBM (replaces the two passes of the reference's
// cufftMakePlanMany 2-D plans, src/fft/transform_2d_gpu.hpp:51-140). Argument struct and item order:
//
// Item order (handed out by an atomic counter): for step u = 0 .. P+lag-1:
//     the A tiles of plane u (if u < P), then the B tiles of plane u-lag (if u >= lag)
//   backward: A = y tile (sticks -> scratch),  B = x tile (scratch -> space domain)
//   forward : A = x tile (space -> scratch),   B = y tile (scratch -> sticks)
// B tiles of plane p wait until all A tiles of p are done; A tiles of plane p wait until all B
// tiles of plane p-ring are done (slot reuse). Dependencies only point to earlier items, and items
// are started in order, so waiting never deadlocks.
// -------------------------------------------------------------------------------------------
template <typename T>
struct XYArgs {
  YArgs<T> y;      // stick side: sticks, pitch, zRowOffset, xtStart, stickSlot, symmetry, ftw (y)
  XArgs<T> x;      // space side: spaceIn / spaceOut, ny, numRowTiles, ftw (x)
  cx<T>* scratch;  // [ring][ny][nx]
  int ring;
  int lag;
  int* counters;   // [0] work counter, [1+p] A tiles done of plane p, [1+P+p] B tiles done
};

struct XYItem {
  int plane, tile;
  bool roleA, valid;
};

template <typename T, bool BWD>
SB_HD int xy_tiles_a(const XYArgs<T>& a) {
  return BWD ? a.y.numXTiles : a.x.numRowTiles;
}
template <typename T, bool BWD>
SB_HD int xy_tiles_b(const XYArgs<T>& a) {
  return BWD ? a.x.numRowTiles : a.y.numXTiles;
}
template <typename T, bool BWD>
SB_HD long long xy_total_items(const XYArgs<T>& a) {
  return (long long)(a.y.numPlanes + a.lag) * (xy_tiles_a<T, BWD>(a) + xy_tiles_b<T, BWD>(a));
}
template <typename T, bool BWD>
SB_HD XYItem xy_decode(const XYArgs<T>& a, int item) {
  const int nA = xy_tiles_a<T, BWD>(a), nB = xy_tiles_b<T, BWD>(a);
  const int u = item / (nA + nB);
  const int r = item - u * (nA + nB);
  XYItem it;
  it.roleA = r < nA;
  it.plane = it.roleA ? u : u - a.lag;
  it.tile = it.roleA ? r : r - nA;
  it.valid = it.roleA ? (u < a.y.numPlanes) : (it.plane >= 0);
  return it;
}

// Items in hand-out order, valid ones only (dense index 0 .. 2 * 64 * P - 1): for step u = 0 .. P + lag - 1 the
// 64 A tiles of plane u (if u < P), then the 64 B tiles of plane u - lag (if u >= lag) -- the order of xy_decode
// (fast_stage_kernels.hpp) without its empty slots.
constexpr int kWTiles = 64;  // tiles of 8 columns / rows per plane (N = 512)
// (TILES: 64, or 32 for the single-precision kernels whose items cover 16 columns / rows)
template <int TILES = kWTiles>
SB_HD XYItem w_decode_dense(int idx, int P, int lag) {
  XYItem it;
  it.valid = true;
  const int lead = (lag < P ? lag : P) * TILES;  // steps with A tiles only
  if (idx < lead) {
    it.roleA = true;
    it.plane = idx / TILES;
    it.tile = idx % TILES;
    return it;
  }
  idx -= lead;
  const int both = P > lag ? (P - lag) * 2 * TILES : 0;  // steps with A and B tiles
  if (idx < both) {
    const int u = lag + idx / (2 * TILES);
    const int r = idx % (2 * TILES);
    it.roleA = r < TILES;
    it.plane = it.roleA ? u : u - lag;
    it.tile = r % TILES;
    return it;
  }
  idx -= both;
  it.roleA = false;  // steps with B tiles only
  it.plane = (P > lag ? P - lag : 0) + idx / TILES;
  it.tile = idx % TILES;
  return it;
}

#undef SB_ROW_IDS
#undef SB_FAST_IDS
#undef SB_FAST_IDS_
#undef SB_FAST_WRITE
#undef SB_FAST_READ

}  // namespace sb
