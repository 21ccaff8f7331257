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
// One CTA = one tile of V lanes (V = 8 double / 16 float = 128 bytes per tile row), V*N/8 threads,
// ONE tile buffer of N*V complex values in shared memory.
#pragma once
#include "fast_fft.hpp"
#include "stage_kernels.hpp"

namespace sb {

template <typename T>
struct FastLanes {
  static constexpr int log2V = sizeof(T) == 8 ? 3 : 4;
};

// Stages 0 .. last-1 with their exchanges; on return the tile holds the input of the last stage.
// `vAll`: the caller's SB_REGS array. COL selects the thread mapping.
template <typename T, int N, int LOG2V, bool BWD, typename Swz, bool COL>
SB_DEV void fast_fft_head(cx<T>* vAll, cx<T>* S, const cx<T>* __restrict__ tw, Ctx ctx) {
  (void)ctx;
  using P = FastPlan<N>;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = P::T;
#define SB_FAST_IDS                                   \
  cx<T>* v = SB_RP(vAll, 8);                          \
  const int lane = COL ? tid / TT : (tid & (V - 1));  \
  const int j = COL ? (tid & (TT - 1)) : (tid >> LOG2V); \
  (void)nthr;
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

  if constexpr (P::numStages > 1) {
    SB_PHASE_BEGIN
    SB_FAST_IDS
    fast_stage<T, N, BWD, 0>(v, j, tw);
    SB_FAST_WRITE(0)
    SB_PHASE_END
  }
  if constexpr (P::numStages > 2) {
    SB_PHASE_BEGIN
    SB_FAST_IDS
    SB_FAST_READ
    fast_stage<T, N, BWD, 1>(v, j, tw);
    SB_PHASE_END
    SB_PHASE_BEGIN
    SB_FAST_IDS
    SB_FAST_WRITE(1)
    SB_PHASE_END
  }
  if constexpr (P::numStages > 3) {
    SB_PHASE_BEGIN
    SB_FAST_IDS
    SB_FAST_READ
    fast_stage<T, N, BWD, 2>(v, j, tw);
    SB_PHASE_END
    SB_PHASE_BEGIN
    SB_FAST_IDS
    SB_FAST_WRITE(2)
    SB_PHASE_END
  }
  static_assert(P::numStages <= 4, "N <= 4096");
}

// Last stage, to be called inside the caller's final phase: afterwards v[m] = X[j + T*m].
template <typename T, int N, int LOG2V, bool BWD, typename Swz>
SB_DEV void fast_fft_tail(cx<T>* v, const cx<T>* S, const cx<T>* __restrict__ tw, int j, int lane) {
  using P = FastPlan<N>;
  constexpr int TT = P::T;
  if constexpr (P::numStages > 1) {
    SB_FAST_READ
  }
  fast_stage<T, N, BWD, P::numStages - 1>(v, j, tw);
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
template <typename T, int N>
SB_DEV void z_backward_fast(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  SB_REGS(cx<T>, vAll, 8);
  SB_PHASE_BEGIN
  for (int i = tid; i < N * V; i += nthr) S[i] = mk<T>(0, 0);
  SB_PHASE_END
  SB_PHASE_BEGIN
  const int e1 = a.tileStart[tile + 1];
  for (int e = a.tileStart[tile] + tid; e < e1; e += nthr) {
    const int src = a.entrySrc ? a.entrySrc[e] : e;
    const int slot = a.entrySlot[e];
    S[SwzRow::at<LOG2V>(slot >> LOG2V, slot & (V - 1))] = a.valuesIn[src];
  }
  SB_PHASE_END
  if (tile == a.symTile) hermitian_fill_lane_swz<T, LOG2V, SwzRow>(S, N, a.symLane, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = S[SwzRow::at<LOG2V>(j + TT * m, lane)];
  SB_PHASE_END
  fast_fft_head<T, N, LOG2V, true, SwzRow, false>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
  fast_fft_tail<T, N, LOG2V, true, SwzRow>(v, S, a.ftw, j, lane);
  cx<T>* out = a.sticks + (size_t)tile * V + lane;
#pragma unroll
  for (int m = 0; m < 8; ++m) out[(size_t)(j + TT * m) * a.pitch] = v[m];
  SB_PHASE_END_NOSYNC
}

template <typename T, int N>
SB_DEV void z_forward_fast(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  SB_REGS(cx<T>, vAll, 8);
  SB_PHASE_BEGIN
  SB_ROW_IDS
  const cx<T>* in = a.sticks + (size_t)tile * V + lane;
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = in[(size_t)(j + TT * m) * a.pitch];
  SB_PHASE_END
  fast_fft_head<T, N, LOG2V, false, SwzRow, false>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
  fast_fft_tail<T, N, LOG2V, false, SwzRow>(v, S, a.ftw, j, lane);
  SB_PHASE_END  // every thread has read its inputs of the last stage
  SB_PHASE_BEGIN
  SB_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m) S[SwzRow::at<LOG2V>(j + TT * m, lane)] = v[m];
  SB_PHASE_END
  SB_PHASE_BEGIN
  const int e1 = a.tileStart[tile + 1];
  for (int e = a.tileStart[tile] + tid; e < e1; e += nthr) {
    const int dst = a.entrySrc ? a.entrySrc[e] : e;
    const int slot = a.entrySlot[e];
    cx<T> val = S[SwzRow::at<LOG2V>(slot >> LOG2V, slot & (V - 1))];
    if (a.useScale) val = a.scale * val;
    a.valuesOut[dst] = val;
  }
  SB_PHASE_END_NOSYNC
}

// -------------------------------------------------------------------------------------------
// y stage
// -------------------------------------------------------------------------------------------
template <typename T, int N>
SB_DEV void y_backward_fast(const YArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  const int xt = block % a.numXTiles;
  const int zl = block / a.numXTiles;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  cx<T>* planeTile = a.planes + (size_t)zl * N * a.nxf + (size_t)xt * V;
  const int lanesValid = (a.nxf - xt * V) < V ? (a.nxf - xt * V) : V;
  if (e0 == e1) {
    SB_PHASE_BEGIN
    for (int i = tid; i < N * V; i += nthr) {
      const int y = i >> LOG2V;
      const int lane = i & (V - 1);
      if (lane < lanesValid) planeTile[(size_t)y * a.nxf + lane] = mk<T>(0, 0);
    }
    SB_PHASE_END_NOSYNC
    return;
  }
  SB_REGS(cx<T>, vAll, 8);
  SB_PHASE_BEGIN
  for (int i = tid; i < N * V; i += nthr) S[i] = mk<T>(0, 0);
  SB_PHASE_END
  SB_PHASE_BEGIN
  const cx<T>* row = a.sticks + (size_t)(zl + a.zRowOffset) * a.pitch;
  for (int e = e0 + tid; e < e1; e += nthr) {
    const int slot = a.stickSlot[e];
    S[SwzRow::at<LOG2V>(slot >> LOG2V, slot & (V - 1))] = row[e];
  }
  SB_PHASE_END
  if (a.symmetry && xt == 0) hermitian_fill_lane_swz<T, LOG2V, SwzRow>(S, N, 0, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = S[SwzRow::at<LOG2V>(j + TT * m, lane)];
  SB_PHASE_END
  fast_fft_head<T, N, LOG2V, true, SwzRow, false>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
  fast_fft_tail<T, N, LOG2V, true, SwzRow>(v, S, a.ftw, j, lane);
  if (lane < lanesValid) {
#pragma unroll
    for (int m = 0; m < 8; ++m) planeTile[(size_t)(j + TT * m) * a.nxf + lane] = v[m];
  }
  SB_PHASE_END_NOSYNC
}

template <typename T, int N>
SB_DEV void y_forward_fast(const YArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  const int xt = block % a.numXTiles;
  const int zl = block / a.numXTiles;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  if (e0 == e1) return;
  const cx<T>* planeTile = a.planes + (size_t)zl * N * a.nxf + (size_t)xt * V;
  const int lanesValid = (a.nxf - xt * V) < V ? (a.nxf - xt * V) : V;
  SB_REGS(cx<T>, vAll, 8);
  SB_PHASE_BEGIN
  SB_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m)
    v[m] = lane < lanesValid ? planeTile[(size_t)(j + TT * m) * a.nxf + lane] : mk<T>(0, 0);
  SB_PHASE_END
  fast_fft_head<T, N, LOG2V, false, SwzRow, false>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  SB_ROW_IDS
  fast_fft_tail<T, N, LOG2V, false, SwzRow>(v, S, a.ftw, j, lane);
  SB_PHASE_END
  SB_PHASE_BEGIN
  SB_ROW_IDS
#pragma unroll
  for (int m = 0; m < 8; ++m) S[SwzRow::at<LOG2V>(j + TT * m, lane)] = v[m];
  SB_PHASE_END
  SB_PHASE_BEGIN
  cx<T>* row = a.sticks + (size_t)(zl + a.zRowOffset) * a.pitch;
  for (int e = e0 + tid; e < e1; e += nthr) {
    const int slot = a.stickSlot[e];
    row[e] = S[SwzRow::at<LOG2V>(slot >> LOG2V, slot & (V - 1))];
  }
  SB_PHASE_END_NOSYNC
}

// -------------------------------------------------------------------------------------------
// x stage, complex rows (C2C). Tile = V consecutive rows; thread = (row lane, j).
// -------------------------------------------------------------------------------------------
template <typename T, int N, bool BWD>
SB_DEV void x_c2c_fast(const XArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanes<T>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = FastPlan<N>::T;
  const int rt = block % a.numRowTiles;
  const int zl = block / a.numRowTiles;
  const int y0 = rt * V;
  const size_t rowBase = (size_t)zl * a.ny + y0;
  const cx<T>* src = BWD ? a.planes : static_cast<const cx<T>*>(a.spaceIn);
  cx<T>* dst = BWD ? static_cast<cx<T>*>(a.spaceOut) : a.planes;
  SB_REGS(cx<T>, vAll, 8);
#define SB_COL_IDS                      \
  cx<T>* v = SB_RP(vAll, 8);            \
  const int lane = tid / TT;            \
  const int j = tid & (TT - 1);         \
  const bool valid = y0 + lane < a.ny;  \
  (void)nthr;
  SB_PHASE_BEGIN
  SB_COL_IDS
  const cx<T>* in = src + (rowBase + lane) * N + j;
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = valid ? in[TT * m] : mk<T>(0, 0);
  SB_PHASE_END
  fast_fft_head<T, N, LOG2V, BWD, SwzCol, true>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  SB_COL_IDS
  fast_fft_tail<T, N, LOG2V, BWD, SwzCol>(v, S, a.ftw, j, lane);
  if (valid) {
    cx<T>* out = dst + (rowBase + lane) * N + j;
#pragma unroll
    for (int m = 0; m < 8; ++m) out[TT * m] = v[m];
  }
  SB_PHASE_END_NOSYNC
#undef SB_COL_IDS
}

#undef SB_ROW_IDS
#undef SB_FAST_IDS
#undef SB_FAST_WRITE
#undef SB_FAST_READ

}  // namespace sb
