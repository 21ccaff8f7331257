// fast_pipe.cu -- pipelined persistent xy stage for sm_100a: TMA-staged tiles, warp groups,
// hand-off through L2 (C2C, dimX == dimY == N power of two, double precision).
//
// One CTA per SM, NG groups of V*N/8 threads, NG+1 tile buffers in shared memory. Work items
// (y tiles and x tiles of all local planes, in the dependency order of xy_decode) are claimed with
// an atomic counter; local tile q of a CTA lives in buffer q % (NG+1) and is transformed by group
// q % NG. While NG tiles are being transformed, the spare buffer is filled by ONE bulk copy
// (cp.async.bulk global -> shared, completion on an mbarrier): every tile's input is a single
// contiguous block (fast_pipe_kernels.hpp), so the memory side needs no registers, no load
// instructions and no waiting in the compute warps. The thread that frees a buffer (after the last
// shared-memory read of its tile) claims the next item and issues its copy; dependencies between
// y and x tiles of a plane are counters in global memory (release by the last warp of a tile,
// acquire by the issuing thread); the hand-off planes live in a ring of scratch slots that stays
// in the 126 MB L2 and whose consumed lines are discarded instead of written back.
//
// Replaces the separate y and x kernels (fast_y.cu / fast_x.cu) and the first fused kernel
// (fast_xy.cu) where it applies.
#include <cstdint>

#include "fast_launch.cuh"
#include "fast_pipe_kernels.hpp"
#include "launch.h"

namespace sb {

// ---- PTX wrappers: mbarrier, bulk copy, proxy fence ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// global -> shared bulk copy (TMA), bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T, int N>
struct PipeCfg {
  static constexpr int V = 1 << FastLanes<T>::log2V;
  static constexpr int GT = V * (N / 8);                      // threads per group (one tile)
  static constexpr int NG = GT >= 512 ? 2 : (GT >= 256 ? 4 : 8);  // groups per CTA
  static constexpr int NB = NG + 1;                           // tile buffers
  static constexpr int threads = GT * NG;
  static constexpr size_t tileBytes = sizeof(cx<T>) * (size_t)N * V;
  static constexpr size_t twBytes = sizeof(cx<T>) * (FastPlan<N>::tw_size() > 0 ? FastPlan<N>::tw_size() : 1);
  static constexpr size_t smem = tileBytes * NB + twBytes;
  static constexpr bool supported = sizeof(T) == 8 && N >= 128 && threads <= 1024 && smem <= 225 * 1024;
};

template <typename T, int N, bool FWD>
__global__ void __launch_bounds__(PipeCfg<T, N>::threads, 1) k_xy_pipe(const __grid_constant__ XYArgs<T> a) {
  using C = PipeCfg<T, N>;
  constexpr bool BWD = !FWD;
  constexpr int V = C::V, GT = C::GT, NG = C::NG, NB = C::NB;
  constexpr int WARPS = GT / 32;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  cx<T>* tws = S + (size_t)NB * N * V;
  __shared__ uint64_t full[NB];          // bulk copy of the buffer's tile has landed
  __shared__ volatile int sPend[NB];     // 1: item claimed, copy not issued (dependency not met yet)
  __shared__ volatile int sTile[NB];     // local tile index the buffer was (re)filled for
  __shared__ int sArrive[NG][2];         // warps of the group that finished their tile (by parity)

  const int group = (int)threadIdx.x / GT;
  const int gtid = (int)threadIdx.x - group * GT;
  const int P = a.y.numPlanes;
  const int nA = xy_tiles_a<T, BWD>(a), nB = xy_tiles_b<T, BWD>(a);
  const int total = (int)xy_total_items<T, BWD>(a);
  int* aDone = a.counters + 1;
  int* bDone = a.counters + 1 + P;

  // dependency of an item: A tile -> its scratch slot was consumed, B tile -> its plane is complete
  auto ready = [&](const XYItem& it) -> bool {
    if (!it.valid) return true;
    if (it.roleA) return it.plane < a.ring || ld_acquire_gpu_s32(&bDone[it.plane - a.ring]) >= nB;
    return ld_acquire_gpu_s32(&aDone[it.plane]) >= nA;
  };
  // start the bulk copy of `item` into buffer b (the caller has seen its dependency satisfied)
  auto issue = [&](int b, int item) {
    unsigned bytes = 0;
    const XYItem it = xy_decode<T, BWD>(a, item);
    const cx<T>* src = it.valid ? pipe_item_source<T, N, BWD>(a, it, &bytes) : nullptr;
    if (bytes == 0) {
      mbar_arrive(&full[b]);
    } else {
      fence_proxy_async();  // generic-proxy accesses (this CTA's exchanges, other SMs' stores) before the copy
      mbar_arrive_expect_tx(&full[b], bytes);
      bulk_load(S + (size_t)b * N * V, src, bytes, &full[b]);
    }
  };
  // Static schedule: local tile q of this CTA is item q * gridDim.x + blockIdx.x. Items grow with
  // q, so the first tile past the end of the schedule ends a group's loop and nobody ever waits
  // for a later one.
  auto item_of = [&](int q) -> int { return q * (int)gridDim.x + (int)blockIdx.x; };
  // Stage local tile q into buffer b (now free) if its dependency is already met, else leave it
  // pending for the consuming group. sTile tells the phases of a buffer apart: groups run
  // independently, so a group may reach the wait for tile q while an EARLIER tile of the same
  // buffer has not even landed, which the mbarrier parity alone cannot distinguish.
  auto refill = [&](int b, int q) {
    const int item = item_of(q);
    if (item >= total) return;
    sTile[b] = q;
    __threadfence_block();
    const XYItem it = xy_decode<T, BWD>(a, item);
    // this CTA's tile after next: its HBM-resident input -> L2
    {
      const int ahead = item_of(q + 2);
      if (ahead < total) {
        const XYItem pf = xy_decode<T, BWD>(a, ahead);
        if (pf.valid && pf.roleA) {
          unsigned pb = 0;
          const cx<T>* ps = pipe_item_source<T, N, BWD>(a, pf, &pb);
          if (pb) bulk_prefetch_l2(ps, pb);
        }
      }
    }
    if (ready(it)) {
      issue(b, item);
    } else {
      __threadfence_block();
      sPend[b] = 1;
    }
  };

  if (threadIdx.x == 0) {
    for (int b = 0; b < NB; ++b) {
      mbar_init(&full[b], 1);
      sPend[b] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < NG * 2) (&sArrive[0][0])[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < FastPlan<N>::tw_size(); i += blockDim.x) tws[i] = a.x.ftw[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int b = 0; b < NB; ++b) refill(b, b);
  }
  __syncthreads();

  Ctx ctx{GT};
  ctx.tidBase = group * GT;
  ctx.barId = 1 + group;
  for (int q = group, use = 0;; q += NG, ++use) {
    const int item = item_of(q);
    if (item >= total) break;
    const int b = q % NB;
    const uint32_t parity = (uint32_t)(q / NB) & 1u;
    ctx.traceItem = use;
    SB_MARK(ctx, 0);
    // wait for the tile; thread 0 of the group issues the copy itself if it was left pending
    if (gtid == 0) {
      while (!(mbar_try_wait(&full[b], parity) && sTile[b] == q)) {
        if (sPend[b] && sTile[b] == q) {
          __threadfence_block();
          const XYItem pit = xy_decode<T, BWD>(a, item);
          if (ready(pit)) {
            sPend[b] = 0;
            issue(b, item);
          } else {
            __nanosleep(100);
          }
        }
      }
    } else {
      while (!(mbar_try_wait(&full[b], parity) && sTile[b] == q)) {
      }
    }
    SB_MARK(ctx, 1);
    const XYItem it = xy_decode<T, BWD>(a, item);
#ifdef SB_XY_TRACE
    if (gtid == 0 && blockIdx.x < kTraceCtas && use < kTraceItems && group == 0)
      g_xy_trace[((size_t)blockIdx.x * kTraceItems + use) * kTraceMarks + 15] =
          it.valid ? (it.roleA ? 1 : 2) + 4 * (long long)it.plane + 4096LL * it.tile : 0;
    if (group != 0) ctx.traceItem = -1;
#endif
    cx<T>* B = S + (size_t)b * N * V;
    auto onFree = [&]() { refill(b, q + NB); };
    if (it.valid) {
      pipe_run_item<T, N, BWD>(a, it, B, tws, ctx, onFree);
      SB_MARK(ctx, 6);
      // completion: last warp of the group publishes the item (and drops consumed hand-off lines)
      {
        const int lane = threadIdx.x & 31;
        int* arr = &sArrive[group][use & 1];
        __syncwarp();
        int last = 0;
        if (lane == 0) {
          __threadfence_block();
          last = atomicAdd(arr, 1) == WARPS - 1;
          __threadfence_block();
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          if (lane == 0) *arr = 0;
          if (!it.roleA) {
            unsigned bytes = 0;
            const cx<T>* src = pipe_item_source<T, N, BWD>(a, it, &bytes);
            if (bytes) discard_l2(src, bytes, lane, 32);  // both B-tile inputs are whole 128-byte lines
            __syncwarp();
          }
          if (lane == 0) {
            __threadfence();
            atomicAdd(it.roleA ? &aDone[it.plane] : &bDone[it.plane], 1);
          }
        }
      }
      SB_MARK(ctx, 7);
    } else {
      // padding item of the schedule (no tile): free the buffer again (after every thread of the
      // group has observed the buffer's mbarrier phase)
      group_sync(ctx);
      if (gtid == 0) refill(b, q + NB);
    }
  }
}

template <typename T, int N>
static int launch_xy_pipe_n(int forward, const XYArgs<T>& a, cudaStream_t s) {
  using C = PipeCfg<T, N>;
  if constexpr (!C::supported) {
    return (int)cudaErrorInvalidValue;
  } else {
    const long long total = forward ? xy_total_items<T, false>(a) : xy_total_items<T, true>(a);
    if (total <= 0) return 0;
    if (total > 0x3fffffffLL) return (int)cudaErrorInvalidConfiguration;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    auto kf = k_xy_pipe<T, N, true>;
    auto kb = k_xy_pipe<T, N, false>;
    e = cudaFuncSetAttribute(forward ? kf : kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem);
    if (e != cudaSuccess) return (int)e;
    long long grid = sms;
    if (grid * C::NG > total) grid = (total + C::NG - 1) / C::NG;
    e = cudaMemsetAsync(a.counters, 0, sizeof(int) * (1 + 2 * (size_t)a.y.numPlanes), s);
    if (e != cudaSuccess) return (int)e;
    if (forward)
      kf<<<(unsigned)grid, C::threads, C::smem, s>>>(a);
    else
      kb<<<(unsigned)grid, C::threads, C::smem, s>>>(a);
    return (int)cudaGetLastError();
  }
}

template <typename T>
static int launch_xy_pipe(int forward, const XYArgs<T>& a, cudaStream_t s) {
#define CALL(NN) return launch_xy_pipe_n<T, NN>(forward, a, s)
  SB_FAST_DISPATCH(a.x.nx, CALL)
#undef CALL
  return (int)cudaErrorInvalidValue;
}

template <typename T>
static bool pipe_supported(int n) {
  switch (n) {
    case 128: return PipeCfg<T, 128>::supported;
    case 256: return PipeCfg<T, 256>::supported;
    case 512: return PipeCfg<T, 512>::supported;
    case 1024: return PipeCfg<T, 1024>::supported;
    default: return false;
  }
}
template <typename T>
static int pipe_groups(int n) {
  switch (n) {
    case 128: return PipeCfg<T, 128>::NG + 1;
    case 256: return PipeCfg<T, 256>::NG + 1;
    case 512: return PipeCfg<T, 512>::NG + 1;
    case 1024: return PipeCfg<T, 1024>::NG + 1;
    default: return 0;
  }
}

}  // namespace sb

extern "C" {

int sb_xy_pipe_config(int isFloat, int n, int numPlanes, int* ring, int* lag, int* numCounters) {
  if (isFloat || !sb::pipe_supported<double>(n)) return (int)cudaErrorInvalidValue;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  const int lanes = 8;
  const int perStep = 2 * ((n + lanes - 1) / lanes);
  // items claimed but not finished: every CTA holds one per buffer, plus one being claimed
  const int window = sms * (sb::pipe_groups<double>(n) + 1);
  int l = (window + perStep - 1) / perStep + 1;
  int r = 2 * l + 2;
  if (const char* env = getenv("SPFFT_B200_XY_LAG")) l = atoi(env) > 0 ? atoi(env) : l;
  if (const char* env = getenv("SPFFT_B200_XY_RING")) r = atoi(env) > l ? atoi(env) : l + 1;
  if (r <= l) r = l + 1;
  if (numPlanes <= r) r = numPlanes > 0 ? numPlanes : 1;
  *ring = r;
  *lag = l;
  *numCounters = 1 + 2 * (numPlanes > 0 ? numPlanes : 0);
  return 0;
}

#ifdef SB_XY_TRACE
/* experiment builds only (tools/xy_trace.py) */
__attribute__((visibility("default"))) int sb_pipe_trace_read(long long* host, int maxEntries) {
  const int n = sb::kTraceCtas * sb::kTraceItems * sb::kTraceMarks;
  if (maxEntries < n) return -1;
  if (cudaMemcpyFromSymbol(host, sb::g_xy_trace, sizeof(long long) * n) != cudaSuccess) return -2;
  return n;
}
#endif

int sb_launch_xy_pipe_f64(int forward, const sb::XYArgs<double>* a, void* stream) {
  sb_note_launches(1);
  return sb::launch_xy_pipe<double>(forward, *a, static_cast<cudaStream_t>(stream));
}
}
