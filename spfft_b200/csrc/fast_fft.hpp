// fast_fft.hpp -- register-resident Stockham FFT for power-of-two lengths (8 <= N <= 4096).
//
// Every thread keeps 8 complex values in registers; a transform of length N is done by N/8
// threads in ceil(log8 N) radix stages (a leading radix-2 or radix-4 stage when log2 N is not a
// multiple of 3, radix 8 otherwise) with ONE shared-memory exchange between consecutive stages.
// The first stage takes its input from wherever the caller loaded it (global memory or a
// shared-memory tile), the last stage leaves the result in registers for the caller to store, so
// a tile makes 2 shared-memory round trips for N = 512 instead of the 5 of the generic
// tile_fft (fft_tile.hpp).
//
// Index algebra (Stockham autosort, decimation in time). Stage s has radix R, stride
// NS = product of the earlier radices, and N/R butterflies b:
//     inputs   n  = b + r*N/R                (r = 0..R-1), twiddled by w_{NS*R}^{r*(b mod NS)}
//     outputs  n' = (b - k)*R + k + q*NS     (q = 0..R-1, k = b mod NS)
// With T = N/8 threads per transform, thread j owns the butterflies b = j + i*T (i < 8/R), i.e.
// it always holds the elements n = j + T*m (m = 0..7) at the start of a stage, and after the last
// stage register m holds output element j + T*m.
//
// Replaces the cuFFT plans of the reference (src/fft/transform_1d_gpu.hpp:52-141,
// src/fft/transform_2d_gpu.hpp:51-140): unnormalised DFT, sign + backward / - forward.
#pragma once
#include "cx.hpp"
#include "fft_tile.hpp"

// Phase timeline of the persistent fused xy kernel (experiments only, -DSB_XY_TRACE): thread 0 of
// the first CTAs stores clock64() at numbered points of its first items (tools/xy_trace.py).
#if SB_ON_GPU && defined(SB_XY_TRACE)
namespace sb {
constexpr int kTraceCtas = 8, kTraceItems = 96, kTraceMarks = 16;
static __device__ long long g_xy_trace[kTraceCtas * kTraceItems * kTraceMarks];  // one per TU; fast_xy.cu reads its own
__device__ __forceinline__ void trace_mark(int item, int id) {
  if (threadIdx.x == 0 && blockIdx.x < kTraceCtas && item >= 0 && item < kTraceItems)
    g_xy_trace[((size_t)blockIdx.x * kTraceItems + item) * kTraceMarks + id] = clock64();
}
}  // namespace sb
#define SB_MARK(ctx, id) ::sb::trace_mark((ctx).traceItem, id)
#else
#define SB_MARK(ctx, id)
#endif

// 1: derive w^3, w^5, w^6, w^7 of a radix-8 stage from w^1, w^2, w^4 instead of reading them
#ifndef SB_TW_DERIVE
#define SB_TW_DERIVE 1
#endif

namespace sb {

constexpr int ilog2_c(int n) { return n <= 1 ? 0 : 1 + ilog2_c(n >> 1); }

template <int N>
struct FastPlan {
  static_assert(N >= 8 && (N & (N - 1)) == 0, "power of two >= 8");
  static constexpr int log2N = ilog2_c(N);
  static constexpr int R0 = (log2N % 3 == 0) ? 8 : ((log2N % 3 == 1) ? 2 : 4);
  static constexpr int numStages = (log2N + 2) / 3;
  static constexpr int T = N / 8;  // threads per transform
  static constexpr int radix(int s) { return s == 0 ? R0 : 8; }
  static constexpr int ns(int s) { return s == 0 ? 1 : R0 << (3 * (s - 1)); }
  // stage twiddle table: for every stage s >= 1, 7*ns(s) entries laid out [r-1][k]
  static constexpr int tw_offset(int s) { return s <= 1 ? 0 : tw_offset(s - 1) + 7 * ns(s - 1); }
  static constexpr int tw_size() { return tw_offset(numStages); }
};

// number of entries of the stage twiddle table of a length-n plan (host side, runtime n)
inline int fast_tw_size(int n) {
  int log2n = 0;
  while ((1 << log2n) < n) ++log2n;
  const int r0 = (log2n % 3 == 0) ? 8 : ((log2n % 3 == 1) ? 2 : 4);
  const int stages = (log2n + 2) / 3;
  int total = 0, ns = r0;
  for (int s = 1; s < stages; ++s) {
    total += 7 * ns;
    ns *= 8;
  }
  return total;
}

// read-only 16/8-byte load of a twiddle
template <typename T>
SB_DEV cx<T> ld_ro(const cx<T>* p) {
#if SB_ON_GPU
  if constexpr (sizeof(T) == 8) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return mk<T>(v.x, v.y);
  } else {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    return mk<T>(v.x, v.y);
  }
#else
  return *p;
#endif
}

// Global loads / stores with a cache policy:
//   Plain  : default (L1 + L2)
//   Stream : read-once / write-once data (ld.global.cs / st.global.cs, evict first)
//   L2Only : data handed from one CTA to another through L2 (ld.global.cg / st.global.cg): never
//            served from a (non-coherent) L1
enum class Mem { Plain, Stream, L2Only };

template <Mem M, typename T>
SB_DEV cx<T> ld_g(const cx<T>* p) {
#if SB_ON_GPU
  if constexpr (M == Mem::Plain) {
    return *p;
  } else if constexpr (sizeof(T) == 8) {
    const double2 v = M == Mem::Stream ? __ldcs(reinterpret_cast<const double2*>(p))
                                       : __ldcg(reinterpret_cast<const double2*>(p));
    return mk<T>(v.x, v.y);
  } else {
    const float2 v = M == Mem::Stream ? __ldcs(reinterpret_cast<const float2*>(p))
                                      : __ldcg(reinterpret_cast<const float2*>(p));
    return mk<T>(v.x, v.y);
  }
#else
  return *p;
#endif
}

template <Mem M, typename T>
SB_DEV void st_g(cx<T>* p, cx<T> v) {
#if SB_ON_GPU
  if constexpr (M == Mem::Plain) {
    *p = v;
  } else if constexpr (sizeof(T) == 8) {
    if constexpr (M == Mem::Stream)
      __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
    else
      __stcg(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
  } else {
    if constexpr (M == Mem::Stream)
      __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
    else
      __stcg(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
  }
#else
  *p = v;
#endif
}

// Ask L2 to fetch [base, base+bytes) (no registers or shared memory held, no wait): issued for the
// inputs of the tile that will run on this SM next, so that its demand loads find the data in L2
// instead of paying the DRAM latency with only 2 CTAs per SM to hide it.
SB_DEV void prefetch_l2(const void* base, size_t bytes, int tid, int nthr) {
#if SB_ON_GPU
  const char* p = static_cast<const char*>(base);
  const size_t mis = reinterpret_cast<size_t>(p) & 127;
  p -= mis;
  bytes += mis;
  for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)nthr * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
#else
  (void)base; (void)bytes; (void)tid; (void)nthr;
#endif
}
// Tell L2 that [base, base+bytes) (128-byte aligned lines) will not be read again before it is
// overwritten: dirty lines of the y<->x hand-off are dropped instead of written back to HBM.
SB_DEV void discard_l2(const void* base, size_t bytes, int tid, int nthr) {
#if SB_ON_GPU
  const char* p = static_cast<const char*>(base);
  for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)nthr * 128)
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p + off) : "memory");
#else
  (void)base; (void)bytes; (void)tid; (void)nthr;
#endif
}
SB_DEV void prefetch_l2_line(const void* p) {
#if SB_ON_GPU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// Thread 0's bookkeeping of a persistent kernel (ItemChores, cx.hpp): called right after the loads
// of the tile were issued, so its two L2 round trips overlap the load latency.
SB_DEV void run_item_chores(const Ctx& ctx) {
#if SB_ON_GPU
  if (ctx.chores && threadIdx.x == 0) {
    const ItemChores c = *ctx.chores;
    const int claimed = atomicAdd(c.claimCounter, 1);
    int ready = 1;
    if (c.depCounter) {
      int v;
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(c.depCounter) : "memory");
      ready = v >= c.depNeed;
    }
    *c.claimOut = claimed;
    *c.readyOut = ready;
  }
#else
  (void)ctx;
#endif
}

// Twiddles + butterflies of stage S on the 8 registers of thread j.
// TWS: the twiddle table lives in shared memory (persistent kernels whose L1 is invalidated by
// their acquire / release operations) instead of global memory read through L1.
template <typename T, int N, bool BWD, int S, bool TWS = false>
SB_DEV void fast_stage(cx<T>* v, int j, const cx<T>* __restrict__ tw) {
  using P = FastPlan<N>;
  constexpr int R = P::radix(S);
  constexpr int NS = P::ns(S);
  constexpr int M = 8 / R;
  if (S > 0) {  // R == 8 here
    const int k = j & (NS - 1);
    const cx<T>* t = tw + P::tw_offset(S) + k;
#if SB_TW_DERIVE
    // w^1, w^2, w^4 from the table, the other four powers by one or two multiplications: the L1 /
    // shared-memory data pipe is the busiest unit of every stage kernel (ncu: 70-91 %) and the seven
    // table reads of a thread cost as many wavefronts as its eight data loads; the fp64 pipe has room.
    // Error of a derived power: ~2 ulp, far inside the 1e-12 budget of a length-N transform.
    cx<T> w[8];
    w[1] = TWS ? t[0] : ld_ro(t);
    w[2] = TWS ? t[NS] : ld_ro(t + NS);
    w[4] = TWS ? t[3 * NS] : ld_ro(t + 3 * NS);
    w[3] = w[1] * w[2];
    w[5] = w[4] * w[1];
    w[6] = w[4] * w[2];
    w[7] = w[4] * w[3];
#pragma unroll
    for (int r = 1; r < 8; ++r) v[r] = v[r] * (BWD ? conj(w[r]) : w[r]);
#else
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      const cx<T> w = TWS ? t[(r - 1) * NS] : ld_ro(t + (r - 1) * NS);
      v[r] = v[r] * (BWD ? conj(w) : w);
    }
#endif
  }
#pragma unroll
  for (int i = 0; i < M; ++i) {
    cx<T> a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = v[i + M * r];
    Butterfly<T, BWD, R>::run(a);
#pragma unroll
    for (int r = 0; r < R; ++r) v[i + M * r] = a[r];
  }
}

// Where register (i + M*q) of thread j goes in the exchange after stage S.
template <int N, int S>
SB_HD int fast_out_index(int j, int i, int q) {
  using P = FastPlan<N>;
  constexpr int R = P::radix(S);
  constexpr int NS = P::ns(S);
  const int b = j + i * P::T;
  const int k = b & (NS - 1);
  return (b - k) * R + k + q * NS;
}

// XOR swizzles of the lane (16/8-byte slot inside a 128-byte tile row).
//   SwzRow : lanes are the fastest thread index (a quarter/half warp covers a whole row), any
//            function of n is conflict free for the exchanges; n & (V-1) additionally makes
//            accesses that walk along n with a fixed lane (sparse scatter / gather) conflict free.
//   SwzCol : consecutive threads walk along n with a fixed lane in every stage
//            (n = j + T*m reads, (b-k)*R + k + q*NS writes): the varying bits of n are 3-bit
//            groups [0,3), [3,6), [6,9) ..., so fold them down onto the slot bits.
// 4-lane tiles of 16-byte elements (64-byte rows, experiment builds): two consecutive rows form one
// 128-byte line of 8 slots, permuted as a whole (conflict free for N = 256, 512 in every thread
// mapping; searched exhaustively like the 8-lane folds)
SB_HD int swz_pair_rows(int n, int lane) {
  return ((n >> 1) << 3) + ((((n & 1) << 2) | lane) ^ ((n ^ (n >> 1) ^ (n >> 2) ^ (n >> 4)) & 7));
}
struct SwzRow {
  template <int LOG2V>
  static SB_HD int at(int n, int lane) {
    if (LOG2V == 2) return swz_pair_rows(n, lane);
    return (n << LOG2V) + (lane ^ (n & ((1 << LOG2V) - 1)));
  }
};
struct SwzCol {
  // checked exhaustively by tools/check_swizzle.py: conflict free for N >= 64 (16-byte elements,
  // 8 lanes) and N >= 128 (8-byte elements, 16 lanes)
  template <int LOG2V>
  static SB_HD int at(int n, int lane) {
    if (LOG2V == 2) return swz_pair_rows(n, lane);
    const int f = LOG2V == 3 ? (n ^ (n >> 3) ^ (n >> 6) ^ (n >> 9))
                             : (n ^ (n >> 3) ^ (n >> 4) ^ (n >> 6) ^ (n >> 7) ^ (n >> 8));
    return (n << LOG2V) + (lane ^ (f & ((1 << LOG2V) - 1)));
  }
};

// Dense-side stores through the tile buffer and one bulk copy (cp.async.bulk, async proxy) per tile row
// instead of STG.128 per element: a 128-bit global store moves 64 bytes per wavefront of the L1 data pipe,
// a shared-memory store 128, and the bulk copy none (profiles/r01_v4_summary.md). Experiment switch:
// measured SLOWER (z backward 0.55 -> 0.64 ms, y backward 0.77 -> 0.89 ms at 512^3, r01_v4_bulk_store.log):
// the two extra barriers and the wait on the bulk group cost more than the pipe work they save.
#ifndef SB_BULK_STORE
#define SB_BULK_STORE 0
#endif
#if SB_ON_GPU
SB_DEV void bulk_store_row(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes)
               : "memory");
}
SB_DEV void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SB_DEV void bulk_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
#endif

// Programmatic dependent launch: every stage kernel starts with this. `wait` returns once the preceding
// kernel of the stream has completed (no-op for a normal launch); `launch_dependents` lets the NEXT stage
// kernel's CTAs become resident while this grid drains, where they block in their own `wait`. Removes the
// launch gap and the ramp-up between the six dependent kernels of a transform (what bounds 64^3-128^3).
#if SB_ON_GPU
SB_DEV void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

// x stage (column mapping on both sides of every exchange), any lane count V <= 128 bytes / element:
// 128 / ELEM consecutive slots form one 128-byte line = 128 / (ELEM * V) consecutive rows, and the
// slot inside the line is XOR-permuted by a fold of n. Folds found by exhaustive search with the
// model of tools/check_swizzle.py: conflict free for N >= 64 in every exchange of the power-of-two
// plans (and of the 3*2^k plans, whose sub-transforms are power-of-two plans of length N/3).
// Every fold must keep the map a bijection of the tile (tests/test_emu.py::test_x_stage_swizzle).
template <int ELEM>
struct SwzX {
  template <int LOG2V>
  static SB_HD int at(int n, int lane) {
    constexpr int SLOTS = 128 / ELEM;
    constexpr int V = 1 << LOG2V;
    static_assert(V <= SLOTS, "at most 128 bytes per tile row");
    constexpr int ROWS = SLOTS / V;  // rows per line
    int f;
    if (ELEM == 16)
      f = V == 1 ? (n >> 3) : (V == 8 ? (n ^ (n >> 3) ^ (n >> 6) ^ (n >> 9)) : (n ^ (n >> 3)));
    else
      f = V == 1 ? ((n >> 1) ^ (n >> 3)) : V == 2 ? (n ^ (n >> 2) ^ (n >> 6))
                 : (V == 4 ? (n ^ (n >> 1) ^ (n >> 4))
                           : (V == 8 ? (n ^ (n >> 1) ^ (n >> 3)) : (n ^ (n >> 3) ^ (n >> 4) ^ (n >> 6) ^ (n >> 7) ^ (n >> 8))));
    const int slot = ((n % ROWS) << LOG2V) | lane;
    return (n / ROWS) * SLOTS + (slot ^ (f & (SLOTS - 1)));
  }
};

}  // namespace sb
