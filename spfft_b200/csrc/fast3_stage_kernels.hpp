// fast3_stage_kernels.hpp -- register-FFT stage kernel bodies for transform lengths N = 3 * 2^k
// (96, 192, 384, 768: the 2^a*3 grid sizes plane-wave codes pick between the powers of two, e.g.
// the 192^3 band batches of BASELINE.json config 5) and, with G = 5 groups instead of 3 (a radix-5
// step, 40 values per thread), N = 5 * 2^k (160, 320, 640). Same layouts, argument structs and
// semantics as fast_stage_kernels.hpp (power-of-two lengths) and stage_kernels.hpp (any length).
// The comments below describe G = 3; read "3" as G and "24" as 8*G.
//
// Every thread keeps 24 complex values = 3 groups of 8 in registers; a transform of length
// N = 3*M is done by T = M/8 threads:
//   * the radix-3 step runs entirely inside a thread (no exchange): thread j holds the elements
//     j + T*m, m = 0..23, and M = 8*T, so x[n2], x[n2+M], x[n2+2M] share a thread;
//   * the three length-M sub-transforms are the power-of-two register FFT of fast_fft.hpp
//     (FastPlan<M>, ONE shared-memory exchange for M <= 64, two for M <= 512), each in its own
//     third of the tile buffer.
// Two index splittings, chosen per kernel so that the side of the kernel that walks along n with
// consecutive threads (sparse values, stick segments) sees natural order:
//   DIT3 (natural in):  n = M*n1 + n2, k = k1 + 3*k2
//       X[k1 + 3*k2] = sum_n2 w_M^(n2*k2) * { w_N^(n2*k1) * sum_n1 w_3^(n1*k1) x[M*n1 + n2] }
//       register 8*k1 + m of thread j ends up holding X[k1 + 3*(j + T*m)]
//   DIF3 (natural out): n = n1 + 3*n2, k = M*k1 + k2
//       X[M*k1 + k2] = sum_n1 w_3^(n1*k1) * w_N^(n1*k2) * { sum_n2 w_M^(n2*k2) x[n1 + 3*n2] }
//       register 8*n1 + m of thread j must be loaded with x[n1 + 3*(j + T*m)]
// Twiddle table of a length-N plan: [r-1][k] = exp(-2*pi*i*r*k/N) for r = 1,2, k < M, followed by
// the stage twiddles of FastPlan<M> (index_plan.cpp: make_fast_twiddles).
//
// Replaces the cuFFT plans of the reference for these lengths (src/fft/transform_1d_gpu.hpp:52-141,
// transform_2d_gpu.hpp:51-140, transform_real_2d_gpu.hpp:54-256) together with the compression,
// symmetry and transpose kernels fused around them (see stage_kernels.hpp).
#pragma once
#include <type_traits>

#include "fast_stage_kernels.hpp"

namespace sb {

// G = 3 or 5 groups: N = G * 2^k (the same scheme with a radix-5 step serves 160, 320, 640)
constexpr int fast_group_count(int n) { return n % 5 == 0 ? 5 : 3; }

template <int N>
struct Fast3Plan {
  static constexpr int G = fast_group_count(N);
  static_assert(N % G == 0, "N = 3 * 2^k or 5 * 2^k");
  static constexpr int M = N / G;
  using Sub = FastPlan<M>;
  static constexpr int T = M / 8;             // threads per transform
  static constexpr int VPT = 8 * G;           // values per thread
  static constexpr int subTw = (G - 1) * M;   // offset of the sub-transform stage twiddles
  static_assert(Sub::numStages >= 2 && Sub::numStages <= 3, "32 <= M <= 512");
};

constexpr bool is_fast3_length(int n) { return n % 3 == 0 || n % 5 == 0; }

// Lanes per z / y tile of this kernel family: the common choice (FastLanes), except single precision at
// N = 96, 192 where 16 lanes (128-byte rows, still only 64 / 128 threads per CTA) measured 3-6 % faster
// (profiles/r01_v4_float_lanes16.log: 64 bands at 192^3 single 9548 -> 10130 pairs/s; at 384 it loses).
constexpr int fast3_lanes_log2(int n, int complexBytes, int common) {
  return (complexBytes == 8 && n % 3 == 0 && n <= 192) ? 4 : common;
}
template <typename T, int N>
struct Fast3Lanes {
  static constexpr int log2V = fast3_lanes_log2(N, (int)sizeof(cx<T>), FastLanes<T>::log2V);
};

struct LaneJ {
  int lane, j;
};
// column mapping: consecutive threads walk along n; row mapping: lanes fastest
template <int LOG2V, int TT, bool COL>
SB_HD LaneJ fast_ids(int tid) {
  LaneJ r;
  if (COL) {
    r.lane = tid / TT;
    r.j = tid % TT;
  } else {
    r.lane = tid & ((1 << LOG2V) - 1);
    r.j = tid >> LOG2V;
  }
  return r;
}

// w^r, r = 1 .. G-1, of one element: table [r-1][k], or w and its powers (SB_TW_DERIVE: one table read
// per element; the L1 data pipe is the bottleneck unit, the fp64 pipe has room)
template <typename T, int N, bool BWD>
SB_DEV void group_twiddles(cx<T>* w, int k, const cx<T>* __restrict__ tw) {
  using P = Fast3Plan<N>;
#if SB_TW_DERIVE
  w[1] = ld_ro(tw + k);
  w[2] = w[1] * w[1];
  if (P::G > 3) {
    w[3] = w[2] * w[1];
    w[4] = w[2] * w[2];
  }
#else
#pragma unroll
  for (int r = 1; r < P::G; ++r) w[r] = ld_ro(tw + (r - 1) * P::M + k);
#endif
  if (BWD) {
#pragma unroll
    for (int r = 1; r < P::G; ++r) w[r] = conj(w[r]);
  }
}
// radix-G step + twiddles in front of the sub-transforms (DIT)
template <typename T, int N, bool BWD>
SB_DEV void dit3_front(cx<T>* v, int j, const cx<T>* __restrict__ tw) {
  using P = Fast3Plan<N>;
  constexpr int G = P::G;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    cx<T> a[G], w[G];
#pragma unroll
    for (int g = 0; g < G; ++g) a[g] = v[8 * g + m];
    Butterfly<T, BWD, G>::run(a);
    group_twiddles<T, N, BWD>(w, j + P::T * m, tw);
    v[m] = a[0];
#pragma unroll
    for (int g = 1; g < G; ++g) v[8 * g + m] = a[g] * w[g];
  }
}
// twiddles + radix-G step behind the sub-transforms (DIF)
template <typename T, int N, bool BWD>
SB_DEV void dif3_back(cx<T>* v, int j, const cx<T>* __restrict__ tw) {
  using P = Fast3Plan<N>;
  constexpr int G = P::G;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    cx<T> a[G], w[G];
    group_twiddles<T, N, BWD>(w, j + P::T * m, tw);
    a[0] = v[m];
#pragma unroll
    for (int g = 1; g < G; ++g) a[g] = v[8 * g + m] * w[g];
    Butterfly<T, BWD, G>::run(a);
#pragma unroll
    for (int g = 0; g < G; ++g) v[8 * g + m] = a[g];
  }
}

// exchange of one sub-transform through its third of the tile buffer
template <typename T, int M, int LOG2V, typename Swz, int STAGE>
SB_DEV void sub_write(const cx<T>* v, cx<T>* Sg, int j, int lane) {
  using P = FastPlan<M>;
  constexpr int R = P::radix(STAGE);
  constexpr int Q = 8 / R;
#pragma unroll
  for (int i = 0; i < Q; ++i) {
#pragma unroll
    for (int q = 0; q < R; ++q)
      Sg[Swz::template at<LOG2V>(fast_out_index<M, STAGE>(j, i, q), lane)] = v[i + Q * q];
  }
}
template <typename T, int M, int LOG2V, typename Swz>
SB_DEV void sub_read(cx<T>* v, const cx<T>* Sg, int j, int lane) {
  constexpr int TT = FastPlan<M>::T;
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = Sg[Swz::template at<LOG2V>(j + TT * m, lane)];
}

// [radix-3 step,] stages 0 .. last-1 of the three sub-transforms with their exchanges. On return
// the tile buffer holds the inputs of the last stage. COL0 / COL: thread mapping of stage 0 with
// its exchange write / of all later phases (a change of mapping across an exchange is free).
template <typename T, int N, int LOG2V, bool BWD, typename Swz, bool COL0, bool COL, bool DIT>
SB_DEV void fast3_head(cx<T>* vAll, cx<T>* S, const cx<T>* __restrict__ tw, Ctx ctx) {
  (void)ctx;
  using P3 = Fast3Plan<N>;
  using P = typename P3::Sub;
  constexpr int M = P3::M;
  constexpr int G = P3::G;
  constexpr int VPT = 8 * G;
  (void)VPT;
  constexpr int V = 1 << LOG2V;
  const cx<T>* stw = tw + P3::subTw;
  SB_PHASE_BEGIN
  (void)nthr;
  cx<T>* v = SB_RP(vAll, VPT);
  const LaneJ id = fast_ids<LOG2V, P3::T, COL0>(tid);
  if (DIT) dit3_front<T, N, BWD>(v, id.j, tw);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    fast_stage<T, M, BWD, 0>(v + 8 * g, id.j, stw);
    sub_write<T, M, LOG2V, Swz, 0>(v + 8 * g, S + g * M * V, id.j, id.lane);
  }
  SB_PHASE_END
  if constexpr (P::numStages > 2) {
    SB_PHASE_BEGIN
    (void)nthr;
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, P3::T, COL>(tid);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      sub_read<T, M, LOG2V, Swz>(v + 8 * g, S + g * M * V, id.j, id.lane);
      fast_stage<T, M, BWD, 1>(v + 8 * g, id.j, stw);
    }
    SB_PHASE_END
    SB_PHASE_BEGIN
    (void)nthr;
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, P3::T, COL>(tid);
#pragma unroll
    for (int g = 0; g < G; ++g) sub_write<T, M, LOG2V, Swz, 1>(v + 8 * g, S + g * M * V, id.j, id.lane);
    SB_PHASE_END
  }
}

// Last stage of the sub-transforms [+ radix-3 step], inside the caller's final phase.
//   DIT: v[8*k1 + m] = X[k1 + 3*(j + T*m)];   DIF: v[m] = X[j + T*m], m = 0..23
template <typename T, int N, int LOG2V, bool BWD, typename Swz, bool DIT>
SB_DEV void fast3_tail(cx<T>* v, const cx<T>* S, const cx<T>* __restrict__ tw, int j, int lane) {
  using P3 = Fast3Plan<N>;
  using P = typename P3::Sub;
  constexpr int M = P3::M;
  constexpr int G = P3::G;
  constexpr int VPT = 8 * G;
  (void)VPT;
  constexpr int V = 1 << LOG2V;
  const cx<T>* stw = tw + P3::subTw;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    sub_read<T, M, LOG2V, Swz>(v + 8 * g, S + g * M * V, j, lane);
    fast_stage<T, M, BWD, P::numStages - 1>(v + 8 * g, j, stw);
  }
  if (!DIT) dif3_back<T, N, BWD>(v, j, tw);
}

// Inverse-map ("gather") load of the VPT elements j + TT*m of thread (lane, j), column mapping;
// layout of the map as in fast_stage_kernels.hpp with VPT entries per thread.
template <typename T, int N, int TT, int VPT, typename W>
SB_DEV void gather_load_v(cx<T>* v, const W* sparse, const unsigned short* invTile, int tid, int j,
                          int lane, int hermitianLane) {
  static_assert(VPT % 8 == 0, "whole 16-byte map loads");
#pragma unroll
  for (int c = 0; c < VPT / 8; ++c) {
    const Inv8 iv = load_inv8(invTile + (size_t)tid * VPT + 8 * c);
#pragma unroll
    for (int m = 0; m < 8; ++m) v[8 * c + m] = iv.i[m] != kNoEntry ? from_wire<T>(sparse[iv.i[m]]) : mk<T>(0, 0);
  }
  if (lane == hermitianLane) {
#pragma unroll
    for (int m = 0; m < VPT; ++m) {
      const int n = j + TT * m;
      const int n2 = n == 0 ? 0 : N - n;
      const unsigned short i2 = invTile[((size_t)lane * TT + (n2 % TT)) * VPT + n2 / TT];
      const cx<T> q = i2 != kNoEntry ? from_wire<T>(sparse[i2]) : mk<T>(0, 0);
      v[m] = hermitian_combine<T>(n, N, v[m], q);
    }
  }
}
template <typename T, int VPT, typename W>
SB_DEV void gather_store_v(const cx<T>* v, W* sparse, const unsigned short* invThread, bool useScale,
                           T scale) {
#pragma unroll
  for (int c = 0; c < VPT / 8; ++c) {
    const Inv8 iv = load_inv8(invThread + 8 * c);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if (iv.i[m] != kNoEntry) sparse[iv.i[m]] = to_wire<W>(useScale ? scale * v[8 * c + m] : v[8 * c + m]);
    }
  }
}

// Scatter form of the sparse side (values not in stick order, duplicates, mixed-source tiles of a
// distributed transform): the natural-order tile S[n*V + lane] (SwzRow) is filled / drained
// entry by entry as in z_backward_fast / y_backward_tile.
template <typename T, int LOG2V, typename Load>
SB_DEV void scatter_into_tile(cx<T>* S, int elems, const int* slotOf, int e0, int e1, Load load, Ctx ctx) {
  (void)ctx;
  constexpr int V = 1 << LOG2V;
  SB_PHASE_BEGIN
  for (int i = tid; i < elems; i += nthr) S[i] = mk<T>(0, 0);
  SB_PHASE_END
  SB_PHASE_BEGIN
  constexpr int U = 4;
  for (int base = e0 + tid; base < e1; base += U * nthr) {
    int slot[U];
    cx<T> val[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = base + u * nthr;
      if (e < e1) {
        slot[u] = slotOf[e];
        val[u] = load(e);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * nthr < e1) S[SwzRow::template at<LOG2V>(slot[u] >> LOG2V, slot[u] & (V - 1))] = val[u];
    }
  }
  SB_PHASE_END
}

// -------------------------------------------------------------------------------------------
// z stage
// -------------------------------------------------------------------------------------------
template <typename T, int N, bool GATHER, typename W>
SB_DEV void z_backward_fast3_impl(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = Fast3Lanes<T, N>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = Fast3Plan<N>::T;
  constexpr int G = Fast3Plan<N>::G;
  constexpr int VPT = 8 * G;
  (void)G;
  constexpr int THREADS = V * TT;
  using Swz = typename std::conditional<GATHER, SwzCol, SwzRow>::type;
  SB_REGS(cx<T>, vAll, VPT);
  const int e0 = a.tileStart[tile], e1 = a.tileStart[tile + 1];
  if (GATHER) {
    SB_PHASE_BEGIN
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, TT, true>(tid);
    gather_load_v<T, N, TT, VPT>(v, a.valuesIn + e0, a.inv + (size_t)tile * THREADS * VPT, tid, id.j, id.lane,
                                tile == a.symTile ? a.symLane : -1);
    if (a.pfDist > 0 && tile + a.pfDist < a.numTiles) {
      const int p0 = a.tileStart[tile + a.pfDist], p1 = a.tileStart[tile + a.pfDist + 1];
      prefetch_l2(a.valuesIn + p0, (size_t)(p1 - p0) * sizeof(cx<T>), tid, nthr);
      prefetch_l2(a.inv + (size_t)(tile + a.pfDist) * THREADS * VPT, (size_t)THREADS * 2 * VPT, tid, nthr);
    }
    SB_PHASE_END_NOSYNC
  } else {
    scatter_into_tile<T, LOG2V>(
        S, N * V, a.entrySlot, e0, e1, [&](int e) { return a.valuesIn[a.entrySrc ? a.entrySrc[e] : e]; }, ctx);
    if (tile == a.symTile) hermitian_fill_lane_swz<T, LOG2V, SwzRow>(S, N, a.symLane, ctx);
    SB_PHASE_BEGIN
    (void)nthr;
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, TT, false>(tid);
#pragma unroll
    for (int m = 0; m < VPT; ++m) v[m] = S[SwzRow::template at<LOG2V>(id.j + TT * m, id.lane)];
    SB_PHASE_END
  }
  fast3_head<T, N, LOG2V, true, Swz, GATHER, false, true>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  (void)nthr;
  cx<T>* v = SB_RP(vAll, VPT);
  const LaneJ id = fast_ids<LOG2V, TT, false>(tid);
  fast3_tail<T, N, LOG2V, true, Swz, true>(v, S, a.ftw, id.j, id.lane);
  const size_t col = (size_t)tile * V + id.lane;
#pragma unroll
  for (int k1 = 0; k1 < G; ++k1) {
#pragma unroll
    for (int m = 0; m < 8; ++m) z_row<T, W>(a, k1 + G * (id.j + TT * m))[col] = to_wire<W>(v[8 * k1 + m]);
  }
  SB_PHASE_END_NOSYNC
}

template <typename T, int N, bool GATHER, typename W>
SB_DEV void z_forward_fast3_impl(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = Fast3Lanes<T, N>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = Fast3Plan<N>::T;
  constexpr int G = Fast3Plan<N>::G;
  constexpr int VPT = 8 * G;
  (void)G;
  constexpr int THREADS = V * TT;
  using Swz = typename std::conditional<GATHER, SwzCol, SwzRow>::type;
  SB_REGS(cx<T>, vAll, VPT);
  const int e0 = a.tileStart[tile], e1 = a.tileStart[tile + 1];
  SB_PHASE_BEGIN
  cx<T>* v = SB_RP(vAll, VPT);
  const LaneJ id = fast_ids<LOG2V, TT, false>(tid);
  const W* in = reinterpret_cast<const W*>(a.sticks) + (size_t)tile * V + id.lane;
#pragma unroll
  for (int n1 = 0; n1 < G; ++n1) {
#pragma unroll
    for (int m = 0; m < 8; ++m) v[8 * n1 + m] = from_wire<T>(in[(size_t)(n1 + G * (id.j + TT * m)) * a.pitch]);
  }
  if (a.pfDist > 0 && tile + a.pfDist < a.numTiles) {
    for (int r = tid; r < N; r += nthr)
      prefetch_l2_line(reinterpret_cast<const W*>(a.sticks) + (size_t)(tile + a.pfDist) * V + (size_t)r * a.pitch);
    if (GATHER) prefetch_l2(a.inv + (size_t)(tile + a.pfDist) * THREADS * VPT, (size_t)THREADS * 2 * VPT, tid, nthr);
  }
  SB_PHASE_END_NOSYNC
  fast3_head<T, N, LOG2V, false, Swz, false, GATHER, false>(vAll, S, a.ftw, ctx);
  if (GATHER) {
    SB_PHASE_BEGIN
    (void)nthr;
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, TT, true>(tid);
    fast3_tail<T, N, LOG2V, false, Swz, false>(v, S, a.ftw, id.j, id.lane);
    gather_store_v<T, VPT>(v, a.valuesOut + e0, a.inv + ((size_t)tile * THREADS + tid) * VPT, a.useScale != 0,
                          a.scale);
    SB_PHASE_END_NOSYNC
  } else {
    SB_PHASE_BEGIN
    (void)nthr;
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, TT, false>(tid);
    fast3_tail<T, N, LOG2V, false, Swz, false>(v, S, a.ftw, id.j, id.lane);
    SB_PHASE_END  // every thread has read its inputs of the last stage
    SB_PHASE_BEGIN
    (void)nthr;
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, TT, false>(tid);
#pragma unroll
    for (int m = 0; m < VPT; ++m) S[SwzRow::template at<LOG2V>(id.j + TT * m, id.lane)] = v[m];
    SB_PHASE_END
    SB_PHASE_BEGIN
    for (int e = e0 + tid; e < e1; e += nthr) {
      const int slot = a.entrySlot[e];
      cx<T> val = S[SwzRow::template at<LOG2V>(slot >> LOG2V, slot & (V - 1))];
      if (a.useScale) val = a.scale * val;
      a.valuesOut[a.entrySrc ? a.entrySrc[e] : e] = val;
    }
    SB_PHASE_END_NOSYNC
  }
}

template <typename T, int N, bool WIRE = false>
SB_DEV void z_backward_fast3(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* S) {
  if (a.inv)
    z_backward_fast3_impl<T, N, true, WireElem<T, WIRE>>(a, tile, ctx, S);
  else
    z_backward_fast3_impl<T, N, false, WireElem<T, WIRE>>(a, tile, ctx, S);
}
template <typename T, int N, bool WIRE = false>
SB_DEV void z_forward_fast3(const ZArgs<T>& a, int tile, Ctx ctx, cx<T>* S) {
  if (a.inv)
    z_forward_fast3_impl<T, N, true, WireElem<T, WIRE>>(a, tile, ctx, S);
  else
    z_forward_fast3_impl<T, N, false, WireElem<T, WIRE>>(a, tile, ctx, S);
}

// -------------------------------------------------------------------------------------------
// y stage: one tile = V consecutive x columns of one plane.
//   GATHER: the sticks of the tile are contiguous at stickRow[e0 .. e1) (local plane-major row, or
//           one source rank's block of a distributed transform) and a.inv holds the inverse map
//   else  : entry by entry (distributed tiles with sticks from several ranks, or no inverse map)
// -------------------------------------------------------------------------------------------
template <typename T, int N, bool GATHER, typename W>
SB_DEV void y_backward_fast3_impl(const YArgs<T>& a, int xt, int zl, const W* stickRow, int nextXt,
                                  const W* nextStickRow, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = Fast3Lanes<T, N>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = Fast3Plan<N>::T;
  constexpr int G = Fast3Plan<N>::G;
  constexpr int VPT = 8 * G;
  (void)G;
  constexpr int THREADS = V * TT;
  using Swz = typename std::conditional<GATHER, SwzCol, SwzRow>::type;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  cx<T>* planeTile = a.planes + (size_t)zl * N * a.nxf + (size_t)xt * V;
  const int lanesValid = (a.nxf - xt * V) < V ? (a.nxf - xt * V) : V;
  if (e0 == e1) {
    // empty x tile: the x stage still reads these columns -> store zeros, no transform
    SB_PHASE_BEGIN
    for (int i = tid; i < N * V; i += nthr) {
      const int y = i >> LOG2V;
      const int lane = i & (V - 1);
      if (lane < lanesValid) planeTile[(size_t)y * a.nxf + lane] = mk<T>(0, 0);
    }
    SB_PHASE_END_NOSYNC
    return;
  }
  SB_REGS(cx<T>, vAll, VPT);
  if (GATHER) {
    SB_PHASE_BEGIN
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, TT, true>(tid);
    gather_load_v<T, N, TT, VPT>(v, stickRow + e0, a.inv + (size_t)xt * THREADS * VPT, tid, id.j, id.lane,
                                (a.symmetry && xt == 0) ? 0 : -1);
    if (nextXt >= 0) {
      const int p0 = a.xtStart[nextXt], p1 = a.xtStart[nextXt + 1];
      prefetch_l2(nextStickRow + p0, (size_t)(p1 - p0) * sizeof(W), tid, nthr);
    }
    SB_PHASE_END_NOSYNC
  } else {
    scatter_into_tile<T, LOG2V>(
        S, N * V, a.stickSlot, e0, e1,
        [&](int e) {
          return from_wire<T>(a.srcBase ? reinterpret_cast<const W*>(a.sticks)[(size_t)a.srcBase[e] + (size_t)zl * a.srcPitch[e]]
                                        : stickRow[e]);
        },
        ctx);
    if (a.symmetry && xt == 0) hermitian_fill_lane_swz<T, LOG2V, SwzRow>(S, N, 0, ctx);
    SB_PHASE_BEGIN
    (void)nthr;
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, TT, false>(tid);
#pragma unroll
    for (int m = 0; m < VPT; ++m) v[m] = S[SwzRow::template at<LOG2V>(id.j + TT * m, id.lane)];
    SB_PHASE_END
  }
  fast3_head<T, N, LOG2V, true, Swz, GATHER, false, true>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  (void)nthr;
  cx<T>* v = SB_RP(vAll, VPT);
  const LaneJ id = fast_ids<LOG2V, TT, false>(tid);
  fast3_tail<T, N, LOG2V, true, Swz, true>(v, S, a.ftw, id.j, id.lane);
  if (id.lane < lanesValid) {
#pragma unroll
    for (int k1 = 0; k1 < G; ++k1) {
#pragma unroll
      for (int m = 0; m < 8; ++m)
        planeTile[(size_t)(k1 + G * (id.j + TT * m)) * a.nxf + id.lane] = v[8 * k1 + m];
    }
  }
  SB_PHASE_END_NOSYNC
}

// stickRow (GATHER): where stick e of this plane is stored (local row, or the owner's buffer)
template <typename T, int N, bool GATHER, typename W>
SB_DEV void y_forward_fast3_impl(const YArgs<T>& a, int xt, int zl, W* stickRow, int nextXt,
                                 const cx<T>* nextPlane, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = Fast3Lanes<T, N>::log2V;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = Fast3Plan<N>::T;
  constexpr int G = Fast3Plan<N>::G;
  constexpr int VPT = 8 * G;
  (void)G;
  constexpr int THREADS = V * TT;
  using Swz = typename std::conditional<GATHER, SwzCol, SwzRow>::type;
  const int e0 = a.xtStart[xt], e1 = a.xtStart[xt + 1];
  if (e0 == e1) return;  // no stick needs these columns
  const cx<T>* planeTile = a.planes + (size_t)zl * N * a.nxf + (size_t)xt * V;
  const int lanesValid = (a.nxf - xt * V) < V ? (a.nxf - xt * V) : V;
  SB_REGS(cx<T>, vAll, VPT);
  SB_PHASE_BEGIN
  cx<T>* v = SB_RP(vAll, VPT);
  const LaneJ id = fast_ids<LOG2V, TT, false>(tid);
#pragma unroll
  for (int n1 = 0; n1 < G; ++n1) {
#pragma unroll
    for (int m = 0; m < 8; ++m)
      v[8 * n1 + m] = id.lane < lanesValid ? planeTile[(size_t)(n1 + G * (id.j + TT * m)) * a.nxf + id.lane]
                                           : mk<T>(0, 0);
  }
  if (nextXt >= 0) {
    for (int r = tid; r < N; r += nthr) prefetch_l2_line(nextPlane + (size_t)nextXt * V + (size_t)r * a.nxf);
  }
  SB_PHASE_END_NOSYNC
  fast3_head<T, N, LOG2V, false, Swz, false, GATHER, false>(vAll, S, a.ftw, ctx);
  if (GATHER) {
    SB_PHASE_BEGIN
    (void)nthr;
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, TT, true>(tid);
    fast3_tail<T, N, LOG2V, false, Swz, false>(v, S, a.ftw, id.j, id.lane);
    gather_store_v<T, VPT>(v, stickRow + e0, a.inv + ((size_t)xt * THREADS + tid) * VPT, false, T(1));
    SB_PHASE_END_NOSYNC
  } else {
    SB_PHASE_BEGIN
    (void)nthr;
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, TT, false>(tid);
    fast3_tail<T, N, LOG2V, false, Swz, false>(v, S, a.ftw, id.j, id.lane);
    SB_PHASE_END
    SB_PHASE_BEGIN
    (void)nthr;
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, TT, false>(tid);
#pragma unroll
    for (int m = 0; m < VPT; ++m) S[SwzRow::template at<LOG2V>(id.j + TT * m, id.lane)] = v[m];
    SB_PHASE_END
    SB_PHASE_BEGIN
    for (int e = e0 + tid; e < e1; e += nthr) {
      const int slot = a.stickSlot[e];
      W* dst = a.srcBase ? y_dist_stick<T, true, W>(a, e, zl) : stickRow + e;
      *dst = to_wire<W>(S[SwzRow::template at<LOG2V>(slot >> LOG2V, slot & (V - 1))]);
    }
    SB_PHASE_END_NOSYNC
  }
}

template <typename T, int N, typename W>
SB_DEV void y_backward_fast3_w(const YArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  const int xt = block % a.numXTiles;
  const int zl = block / a.numXTiles;
  const W* sticks = reinterpret_cast<const W*>(a.sticks);
  int nextXt = -1;
  const W* nextRow = nullptr;
  if (a.pfDist > 0 && !a.srcBase && block + a.pfDist < a.numXTiles * a.numPlanes) {
    nextXt = (block + a.pfDist) % a.numXTiles;
    nextRow = sticks + (size_t)((block + a.pfDist) / a.numXTiles + a.zRowOffset) * a.pitch;
  }
  const W* localRow = sticks + (size_t)(zl + a.zRowOffset) * a.pitch;
  if (a.srcBase && a.inv && a.tilePitch[xt] != 0) {
    // distributed, all sticks of this tile from one rank: contiguous inside that rank's block
    const W* row = sticks + (size_t)a.tileBase[xt] + (size_t)zl * a.tilePitch[xt] - a.xtStart[xt];
    y_backward_fast3_impl<T, N, true, W>(a, xt, zl, row, -1, nullptr, ctx, S);
  } else if (a.inv && !a.srcBase) {
    y_backward_fast3_impl<T, N, true, W>(a, xt, zl, localRow, nextXt, nextRow, ctx, S);
  } else {
    y_backward_fast3_impl<T, N, false, W>(a, xt, zl, localRow, -1, nullptr, ctx, S);
  }
}
template <typename T, int N, bool WIRE = false>
SB_DEV void y_backward_fast3(const YArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  y_backward_fast3_w<T, N, WireElem<T, WIRE>>(a, block, ctx, S);
}

template <typename T, int N, typename W>
SB_DEV void y_forward_fast3_w(const YArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  const int xt = y_forward_tile_at<T>(a, block % a.numXTiles);
  const int zl = block / a.numXTiles;
  int nextXt = -1;
  const cx<T>* nextPlane = nullptr;
  if (a.pfDist > 0 && block + a.pfDist < a.numXTiles * a.numPlanes) {
    nextXt = y_forward_tile_at<T>(a, (block + a.pfDist) % a.numXTiles);
    nextPlane = a.planes + (size_t)((block + a.pfDist) / a.numXTiles) * N * a.nxf;
  }
  W* localRow = reinterpret_cast<W*>(a.sticks) + (size_t)(zl + a.zRowOffset) * a.pitch;
  if (a.srcBase && a.inv && a.tilePitch[xt] != 0) {
    W* row = y_dist_tile<T, true, W>(a, xt, zl) - a.xtStart[xt];
    y_forward_fast3_impl<T, N, true, W>(a, xt, zl, row, nextXt, nextPlane, ctx, S);
  } else if (a.inv && !a.srcBase) {
    y_forward_fast3_impl<T, N, true, W>(a, xt, zl, localRow, nextXt, nextPlane, ctx, S);
  } else {
    y_forward_fast3_impl<T, N, false, W>(a, xt, zl, localRow, nextXt, nextPlane, ctx, S);
  }
}
template <typename T, int N, bool WIRE = false>
SB_DEV void y_forward_fast3(const YArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  y_forward_fast3_w<T, N, WireElem<T, WIRE>>(a, block, ctx, S);
}

// -------------------------------------------------------------------------------------------
// x stage. Tile = V consecutive rows of one plane, thread = (row lane, j), column mapping on both
// sides. SB_X3_MODE selects how the side that is not in natural order reaches global memory:
//   0: DIT3, natural loads; a thread stores the 3 consecutive elements 3*(j + T*m) + {0,1,2}
//      (stride-3 across threads: half-filled 32-byte sectors on their way to L2)
//   1: DIF3, natural stores; the loads are the stride-3 ones (the three loads of a thread hit the
//      same lines, L1 serves two of them)
//   2: DIT3 + one more pass through the tile buffer to store in natural order
// -------------------------------------------------------------------------------------------
// Measured (profiles/r01_v4_fast3.md): mode 1 is the fastest (x stage at 384^3 double 0.377 -> 0.299 ms,
// at 192^3 single 0.0355 -> 0.0246 ms); mode 2 gains little over mode 0.
#ifndef SB_X3_MODE
#define SB_X3_MODE 1
#endif

// registers of a finished DIT3 transform -> natural order (v[m] = X[j + T*m]) through the tile
template <typename T, int N, int LOG2V>
SB_DEV void dit3_to_natural(cx<T>* vAll, cx<T>* S, Ctx ctx) {
  (void)ctx;
  using SwzCol = SwzX<sizeof(cx<T>)>;
  constexpr int TT = Fast3Plan<N>::T;
  constexpr int G = Fast3Plan<N>::G;
  constexpr int VPT = 8 * G;
  (void)G;
  SB_PHASE_BEGIN  // (the caller's last phase ended with a barrier: every thread is done reading S)
  (void)nthr;
  cx<T>* v = SB_RP(vAll, VPT);
  const LaneJ id = fast_ids<LOG2V, TT, true>(tid);
#pragma unroll
  for (int k1 = 0; k1 < G; ++k1) {
#pragma unroll
    for (int m = 0; m < 8; ++m) S[SwzCol::template at<LOG2V>(k1 + G * (id.j + TT * m), id.lane)] = v[8 * k1 + m];
  }
  SB_PHASE_END
  SB_PHASE_BEGIN
  (void)nthr;
  cx<T>* v = SB_RP(vAll, VPT);
  const LaneJ id = fast_ids<LOG2V, TT, true>(tid);
#pragma unroll
  for (int m = 0; m < VPT; ++m) v[m] = S[SwzCol::template at<LOG2V>(id.j + TT * m, id.lane)];
  SB_PHASE_END_NOSYNC
}

template <typename T, int N, bool BWD>
SB_DEV void x_c2c_fast3(const XArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanesX<T, N>::log2V;
  using SwzCol = SwzX<sizeof(cx<T>)>;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = Fast3Plan<N>::T;
  constexpr int G = Fast3Plan<N>::G;
  constexpr int VPT = 8 * G;
  (void)G;
  constexpr bool DIT = SB_X3_MODE != 1;
  const int rt = block % a.numRowTiles;
  const int zl = block / a.numRowTiles;
  const size_t planeOff = (size_t)zl * a.ny * N;
  const cx<T>* in = (BWD ? a.planes : static_cast<const cx<T>*>(a.spaceIn)) + planeOff;
  cx<T>* out = (BWD ? static_cast<cx<T>*>(a.spaceOut) : a.planes) + planeOff;
  const int y0 = rt * V;
  SB_REGS(cx<T>, vAll, VPT);
  SB_PHASE_BEGIN
  (void)nthr;
  cx<T>* v = SB_RP(vAll, VPT);
  const LaneJ id = fast_ids<LOG2V, TT, true>(tid);
  const bool valid = y0 + id.lane < a.ny;
  const cx<T>* src = in + (size_t)(y0 + id.lane) * N;
  if (DIT) {
#pragma unroll
    for (int m = 0; m < VPT; ++m) v[m] = valid ? src[id.j + TT * m] : mk<T>(0, 0);
  } else {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
#pragma unroll
      for (int n1 = 0; n1 < G; ++n1) v[8 * n1 + m] = valid ? src[n1 + G * (id.j + TT * m)] : mk<T>(0, 0);
    }
  }
  SB_PHASE_END_NOSYNC
  fast3_head<T, N, LOG2V, BWD, SwzCol, true, true, DIT>(vAll, S, a.ftw, ctx);
  SB_PHASE_BEGIN
  (void)nthr;
  cx<T>* v = SB_RP(vAll, VPT);
  const LaneJ id = fast_ids<LOG2V, TT, true>(tid);
  fast3_tail<T, N, LOG2V, BWD, SwzCol, DIT>(v, S, a.ftw, id.j, id.lane);
  if (SB_X3_MODE == 0 && y0 + id.lane < a.ny) {
    cx<T>* dst = out + (size_t)(y0 + id.lane) * N;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
#pragma unroll
      for (int k1 = 0; k1 < G; ++k1) dst[k1 + G * (id.j + TT * m)] = v[8 * k1 + m];
    }
  }
  SB_PHASE_END_IF(SB_X3_MODE == 2)
  if (SB_X3_MODE != 0) {
    if (SB_X3_MODE == 2) dit3_to_natural<T, N, LOG2V>(vAll, S, ctx);
    SB_PHASE_BEGIN
    (void)nthr;
    cx<T>* v = SB_RP(vAll, VPT);
    const LaneJ id = fast_ids<LOG2V, TT, true>(tid);
    if (y0 + id.lane < a.ny) {
      cx<T>* dst = out + (size_t)(y0 + id.lane) * N + id.j;
#pragma unroll
      for (int m = 0; m < VPT; ++m) dst[TT * m] = v[m];
    }
    SB_PHASE_END_NOSYNC
  }
}

// real rows, two per lane (same contract and packing as x_r2c_pair_tile in fast_stage_kernels.hpp)
template <typename T, int N, bool BWD>
SB_DEV void x_r2c_fast3(const XArgs<T>& a, int block, Ctx ctx, cx<T>* S) {
  constexpr int LOG2V = FastLanesX<T, N>::log2V;
  using SwzCol = SwzX<sizeof(cx<T>)>;
  constexpr int V = 1 << LOG2V;
  constexpr int TT = Fast3Plan<N>::T;
  constexpr int G = Fast3Plan<N>::G;
  constexpr int VPT = 8 * G;
  (void)G;
  constexpr int NXF = N / 2 + 1;
  constexpr bool DIT = SB_X3_MODE != 1;
  const int rt = block % a.numRowTiles;  // numRowTiles = ceil(ny / 2V), stage_args.hpp
  const int zl = block / a.numRowTiles;
  const int y0 = rt * 2 * V;
  const size_t planeRow0 = (size_t)zl * a.ny;
  SB_REGS(cx<T>, vAll, VPT);
#define SB_PAIR_IDS                                          \
  cx<T>* v = SB_RP(vAll, VPT);                                \
  const LaneJ id = fast_ids<LOG2V, TT, true>(tid);           \
  const int yA = y0 + 2 * id.lane;                           \
  const bool validA = yA < a.ny;                             \
  const bool validB = yA + 1 < a.ny;                         \
  (void)nthr;
  // element held in register r before the transform / after it
#define SB_IN_INDEX(r) (DIT ? id.j + TT * (r) : (r) / 8 + G * (id.j + TT * ((r) % 8)))
#define SB_OUT_INDEX(r) (DIT ? (r) / 8 + G * (id.j + TT * ((r) % 8)) : id.j + TT * (r))
  SB_PHASE_BEGIN
  SB_PAIR_IDS
  if (BWD) {
    const cx<T>* srcA = a.planes + (planeRow0 + yA) * NXF;
    const cx<T>* srcB = srcA + NXF;
#pragma unroll
    for (int r = 0; r < VPT; ++r) {
      const int n = SB_IN_INDEX(r);
      const bool hi = n >= NXF;
      const int x = hi ? N - n : n;
      const cx<T> A = validA ? srcA[x] : mk<T>(0, 0);
      const cx<T> B = validB ? srcB[x] : mk<T>(0, 0);
      v[r] = pack_half_spectra<T>(A, B, x == 0 || 2 * x == N, hi);
    }
  } else {
    const T* srcA = static_cast<const T*>(a.spaceIn) + (planeRow0 + yA) * N;
    const T* srcB = srcA + N;
#pragma unroll
    for (int r = 0; r < VPT; ++r) {
      const int n = SB_IN_INDEX(r);
      v[r] = mk<T>(validA ? srcA[n] : T(0), validB ? srcB[n] : T(0));
    }
  }
  SB_PHASE_END_NOSYNC
  fast3_head<T, N, LOG2V, BWD, SwzCol, true, true, DIT>(vAll, S, a.ftw, ctx);
  if (BWD) {
    SB_PHASE_BEGIN
    SB_PAIR_IDS
    fast3_tail<T, N, LOG2V, BWD, SwzCol, DIT>(v, S, a.ftw, id.j, id.lane);
    if (SB_X3_MODE != 2) {
      T* dstA = static_cast<T*>(a.spaceOut) + (planeRow0 + yA) * N;
      T* dstB = dstA + N;
#pragma unroll
      for (int r = 0; r < VPT; ++r) {
        const int k = SB_OUT_INDEX(r);
        if (validA) dstA[k] = v[r].x;
        if (validB) dstB[k] = v[r].y;
      }
    }
    SB_PHASE_END_IF(SB_X3_MODE == 2)
    if (SB_X3_MODE == 2) {
      dit3_to_natural<T, N, LOG2V>(vAll, S, ctx);
      SB_PHASE_BEGIN
      SB_PAIR_IDS
      T* dstA = static_cast<T*>(a.spaceOut) + (planeRow0 + yA) * N;
      T* dstB = dstA + N;
#pragma unroll
      for (int r = 0; r < VPT; ++r) {
        if (validA) dstA[id.j + TT * r] = v[r].x;
        if (validB) dstB[id.j + TT * r] = v[r].y;
      }
      SB_PHASE_END_NOSYNC
    }
  } else {
    SB_PHASE_BEGIN
    SB_PAIR_IDS
    (void)validA;
    (void)validB;
    fast3_tail<T, N, LOG2V, BWD, SwzCol, DIT>(v, S, a.ftw, id.j, id.lane);
    SB_PHASE_END  // every thread has read its inputs of the last stage
    SB_PHASE_BEGIN
    SB_PAIR_IDS
    (void)validA;
    (void)validB;
#pragma unroll
    for (int r = 0; r < VPT; ++r) S[SwzCol::template at<LOG2V>(SB_OUT_INDEX(r), id.lane)] = v[r];
    SB_PHASE_END
    SB_PHASE_BEGIN
    SB_PAIR_IDS
    cx<T>* dstA = a.planes + (planeRow0 + yA) * NXF;
    cx<T>* dstB = dstA + NXF;
#pragma unroll
    for (int r = 0; r < VPT; ++r) {
      const int k = id.j + TT * r;
      if (k < NXF) {
        const cx<T> zk = DIT ? S[SwzCol::template at<LOG2V>(k, id.lane)] : v[r];
        cx<T> A, B;
        unpack_half_spectra<T>(zk, S[SwzCol::template at<LOG2V>(k == 0 ? 0 : N - k, id.lane)], A, B);
        if (validA) dstA[k] = A;
        if (validB) dstB[k] = B;
      }
    }
    SB_PHASE_END_NOSYNC
  }
#undef SB_PAIR_IDS
#undef SB_IN_INDEX
#undef SB_OUT_INDEX
}

}  // namespace sb
