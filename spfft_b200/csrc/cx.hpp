// cx.hpp -- minimal complex value type + the host/device portability macros used by the stage
// kernel bodies.
//
// The stage bodies (fft_tile.hpp, stage_kernels.hpp) are written once and compiled two ways:
//   * by nvcc for sm_100a: a "phase" is the code between two __syncthreads(), `tid` is threadIdx.x
//   * by g++ for the CPU-side unit tests (tests/emu): a "phase" is a loop over all thread ids.
// The second build exists only so that index arithmetic can be unit-tested without a GPU; it is
// never linked into libspfft_b200.so and is not a fallback path.
#pragma once
#if !defined(__CUDACC__) || defined(SB_EMULATE)
#include <vector>
#endif

#if defined(__CUDACC__) && !defined(SB_EMULATE)
#define SB_HD __host__ __device__ __forceinline__
#define SB_DEV __device__ __forceinline__
#define SB_ON_GPU 1
// A "group" is the set of threads that work on one tile: the whole CTA (Ctx{blockDim.x}) or, in
// the pipelined persistent kernels, `nthreads` consecutive threads starting at `tidBase` that
// synchronise on their own named barrier `barId`.
#define SB_PHASE_BEGIN                               \
  {                                                  \
    const int tid = (int)threadIdx.x - ctx.tidBase;  \
    const int nthr = ctx.nthreads;
#define SB_PHASE_END \
  }                  \
  ::sb::group_sync(ctx);
// like SB_PHASE_END but without the barrier (last phase of a kernel)
#define SB_PHASE_END_NOSYNC }
// barrier only if the (compile-time) condition holds
#define SB_PHASE_END_IF(c) \
  }                        \
  if (c) ::sb::group_sync(ctx);
// per-thread registers that live across phases
#define SB_REGS(type, name, n) type name[n]
#define SB_RP(name, n) (name)
#else
#define SB_HD inline
#define SB_DEV inline
#define SB_ON_GPU 0
#define SB_PHASE_BEGIN                                \
  for (int tid = 0; tid < ctx.nthreads; ++tid) {      \
    const int nthr = ctx.nthreads;
#define SB_PHASE_END }
#define SB_PHASE_END_NOSYNC }
#define SB_PHASE_END_IF(c) }
#define SB_REGS(type, name, n) std::vector<type> name##_store((size_t)ctx.nthreads * (n)); type* name = name##_store.data()
#define SB_RP(name, n) ((name) + (size_t)tid * (n))
#endif

namespace sb {

template <typename T>
struct alignas(2 * sizeof(T)) cx {
  T x, y;
};

template <typename T>
SB_HD cx<T> mk(T x, T y) {
  cx<T> r;
  r.x = x;
  r.y = y;
  return r;
}
template <typename T>
SB_HD cx<T> operator+(cx<T> a, cx<T> b) {
  return mk<T>(a.x + b.x, a.y + b.y);
}
template <typename T>
SB_HD cx<T> operator-(cx<T> a, cx<T> b) {
  return mk<T>(a.x - b.x, a.y - b.y);
}
template <typename T>
SB_HD cx<T> operator*(cx<T> a, cx<T> b) {
  return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T>
SB_HD cx<T> operator*(T s, cx<T> a) {
  return mk<T>(s * a.x, s * a.y);
}
template <typename T>
SB_HD cx<T> conj(cx<T> a) {
  return mk<T>(a.x, -a.y);
}
// multiply by +i (BWD) or -i (forward): the sign of the DFT exponent
template <bool BWD, typename T>
SB_HD cx<T> mul_si(cx<T> a) {
  return BWD ? mk<T>(-a.y, a.x) : mk<T>(a.y, -a.x);
}
template <typename T>
SB_HD bool nonzero(cx<T> a) {
  return a.x != T(0) || a.y != T(0);
}

// Execution context: empty on the GPU, carries the emulated block size on the CPU.
// Bookkeeping one thread of a persistent kernel does while the loads of its tile are in flight
// (fast_xy.cu): claim a later work item and look at the next item's dependency counter.
struct ItemChores {
  int* claimCounter;      // *claimOut = atomicAdd(claimCounter, 1)
  int* claimOut;
  const int* depCounter;  // *readyOut = depCounter ? (acquire-load(*depCounter) >= depNeed) : 1
  int depNeed;
  int* readyOut;
};

struct Ctx {
  int nthreads;
  int traceItem = -1;                    // experiments only (SB_XY_TRACE)
  const ItemChores* chores = nullptr;    // persistent kernels only
  int tidBase = 0;                       // first thread of the group (pipelined kernels)
  int barId = 0;                         // named barrier of the group, 0 = __syncthreads()
};

#if SB_ON_GPU
__device__ __forceinline__ void group_sync(const Ctx& ctx) {
  if (ctx.barId == 0)
    __syncthreads();
  else
    asm volatile("bar.sync %0, %1;" ::"r"(ctx.barId), "r"(ctx.nthreads) : "memory");
}
#endif

}  // namespace sb
