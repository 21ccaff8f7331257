// fast_xy.cu -- fused xy stage: persistent kernel over y tiles and x tiles with the hand-off held
// in an L2-resident scratch ring (see fast_stage_kernels.hpp), sm_100a.
#include "fast_launch.cuh"
#include "launch.h"

namespace sb {

template <typename T, int N>
constexpr size_t xy_smem() {
  return FastCfg<T, N>::smem + sizeof(cx<T>) * (FastPlan<N>::tw_size() > 0 ? FastPlan<N>::tw_size() : 1);
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Item completion without a CTA-wide barrier. Every warp, after its last global store of the item,
// arrives on a shared-memory counter; the warp that arrives last (all stores of the CTA are then
// ordered before it at CTA scope) optionally drops the consumed hand-off lines from L2, and
// publishes the item with one gpu-scope fence + atomic -- the cumulative release of a
// cooperative-groups grid barrier, but paid by one warp while the others already run the next item.
template <typename DiscardFn>
__device__ __forceinline__ void warp_arrive_and_signal(int* sArrive, int numWarps, int* doneCounter,
                                                       DiscardFn discard) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    __threadfence_block();
    last = atomicAdd(sArrive, 1) == numWarps - 1;
    __threadfence_block();
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (last) {
    if (lane == 0) *sArrive = 0;  // next use: two items later, after barriers every warp passes
    discard(lane);
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      atomicAdd(doneCounter, 1);
    }
  }
}

// Is the dependency of `it` satisfied? A tile: its scratch slot was consumed (B tiles of
// plane - ring done); B tile: the hand-off plane is complete (A tiles of its plane done).
template <typename T>
__device__ __forceinline__ bool xy_item_ready(const XYArgs<T>& a, const XYItem& it, const int* aDone,
                                              const int* bDone, int nA, int nB) {
  if (!it.valid) return true;
  if (it.roleA) return it.plane < a.ring || ld_acquire_gpu(&bDone[it.plane - a.ring]) >= nB;
  return ld_acquire_gpu(&aDone[it.plane]) >= nA;
}

template <typename T, int N, bool FWD>
__global__ void __launch_bounds__(FastCfg<T, N>::threads, FastCfg<T, N>::minBlocks)
    k_xy_fused(const __grid_constant__ XYArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  // stage twiddles in shared memory: the acquire loads / release fences below invalidate L1
  // (CCTL.IVALL) on every item, which would send all twiddle reads to L2
  cx<T>* tws = S + (size_t)N * FastCfg<T, N>::V;
  for (int i = threadIdx.x; i < FastPlan<N>::tw_size(); i += blockDim.x) tws[i] = a.x.ftw[i];
  __shared__ int sQ[2];       // item claimed two iterations ahead
  __shared__ int sReady[2];   // dependency of the NEXT item already seen satisfied (early poll)
  __shared__ int sArrive[2];  // warps that finished the current item (by item parity)
  constexpr bool BWD = !FWD;
  constexpr int V = FastCfg<T, N>::V;
  constexpr int WARPS = FastCfg<T, N>::threads / 32;
  const int P = a.y.numPlanes;
  const int nA = xy_tiles_a<T, BWD>(a), nB = xy_tiles_b<T, BWD>(a);
  const long long total = xy_total_items<T, BWD>(a);
  int* aDone = a.counters + 1;
  int* bDone = a.counters + 1 + P;
  if (threadIdx.x == 0) {
    sQ[0] = atomicAdd(&a.counters[0], 1);
    sQ[1] = atomicAdd(&a.counters[0], 1);
    sReady[0] = sReady[1] = 0;
    sArrive[0] = sArrive[1] = 0;
  }
  __syncthreads();
  int cur = sQ[0], nxt = sQ[1];
  __syncthreads();
  for (int k = 0; cur < total; ++k) {
    const XYItem it = xy_decode<T, BWD>(a, cur);
    XYItem nx;
    nx.valid = false;
    if (nxt < total) nx = xy_decode<T, BWD>(a, nxt);
    Ctx ctx{FastCfg<T, N>::threads};
    ctx.traceItem = k;
    SB_MARK(ctx, 0);
#ifdef SB_XY_TRACE
    if (threadIdx.x == 0 && blockIdx.x < kTraceCtas && k < kTraceItems)
      g_xy_trace[((size_t)blockIdx.x * kTraceItems + k) * kTraceMarks + 15] =
          it.valid ? (it.roleA ? 1 : 2) + 4 * (long long)it.plane + 4096LL * it.tile : 0;
#endif
    // Dependency of this item: normally seen satisfied one item ago by thread 0 (sReady, ordered
    // by the barriers of the previous item); otherwise every warp polls on its own. For A tiles
    // only the final stores depend on it, but they are ordered after this point anyway.
    if (it.valid && !sReady[k & 1]) {
      if ((threadIdx.x & 31) == 0) {
        while (!xy_item_ready<T>(a, it, aDone, bDone, nA, nB)) __nanosleep(64);
      }
      __syncwarp();
    }
    SB_MARK(ctx, 1);
    // Thread 0 claims the item after next and looks at the next item's dependency while the
    // loads of this item are in flight (run_item_chores, called by the tile right after its loads)
    ItemChores chores;
    chores.claimCounter = &a.counters[0];
    chores.claimOut = &sQ[k & 1];
    chores.readyOut = &sReady[(k + 1) & 1];
    chores.depCounter = nullptr;
    chores.depNeed = 0;
    if (nx.valid) {
      if (nx.roleA) {
        if (nx.plane >= a.ring) {
          chores.depCounter = &bDone[nx.plane - a.ring];
          chores.depNeed = nB;
        }
      } else {
        chores.depCounter = &aDone[nx.plane];
        chores.depNeed = nA;
      }
    }
    ctx.chores = &chores;
    if (it.valid) {
      xy_run_item<T, N, BWD>(a, it, nx, tws, ctx, S);
      SB_MARK(ctx, 6);
      if (it.roleA) {
        warp_arrive_and_signal(&sArrive[k & 1], WARPS, &aDone[it.plane], [](int) {});
      } else {
        // the consumed hand-off lines are dropped from L2 instead of being written back to HBM
        const cx<T>* slot = a.scratch + (size_t)(it.plane % a.ring) * N * N;
        const int tile = it.tile;
        warp_arrive_and_signal(&sArrive[k & 1], WARPS, &bDone[it.plane], [&](int lane) {
          if (BWD) {  // x tile: V whole rows, contiguous
            discard_l2(slot + (size_t)tile * V * N, sizeof(cx<T>) * V * N, lane, 32);
          } else if (sizeof(cx<T>) * V >= 128) {  // y tile: one 128-byte column segment per row
            // (64-byte float segments share their line with the neighbour tile: no discard)
            for (int y = lane; y < N; y += 32)
              discard_l2(slot + (size_t)y * N + (size_t)tile * V, sizeof(cx<T>) * V, 0, 1);
          }
        });
      }
      SB_MARK(ctx, 7);
    } else {
      run_item_chores(ctx);
      __syncthreads();
    }
    // every item passes at least one barrier after thread 0 wrote sQ / sReady
    cur = nxt;
    nxt = sQ[k & 1];
  }
}

template <typename T, int N>
static int xy_blocks_per_sm(int* out) {
  using C = FastCfg<T, N>;
  if constexpr (C::threads > 1024) {
    return (int)cudaErrorInvalidValue;
  } else {
    cudaError_t e = cudaFuncSetAttribute(k_xy_fused<T, N, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xy_smem<T, N>());
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_xy_fused<T, N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)xy_smem<T, N>());
    if (e != cudaSuccess) return (int)e;
    int b0 = 0, b1 = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_xy_fused<T, N, false>, C::threads, xy_smem<T, N>());
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_xy_fused<T, N, true>, C::threads, xy_smem<T, N>());
    if (e != cudaSuccess) return (int)e;
    *out = b0 < b1 ? b0 : b1;
    return *out > 0 ? 0 : (int)cudaErrorInvalidConfiguration;
  }
}

template <typename T>
static int xy_blocks_per_sm_n(int n, int* out) {
#define CALL(NN) return xy_blocks_per_sm<T, NN>(out)
  SB_FAST_DISPATCH(n, CALL)
#undef CALL
  return (int)cudaErrorInvalidValue;
}

static int num_sms(int* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev);
}

template <typename T, int N>
static int launch_xy_n(int forward, const XYArgs<T>& a, cudaStream_t s) {
  using C = FastCfg<T, N>;
  if constexpr (C::threads > 1024) {
    return (int)cudaErrorInvalidValue;
  } else {
    int perSm = 0, sms = 0;
    int err = xy_blocks_per_sm<T, N>(&perSm);
    if (err) return err;
    err = num_sms(&sms);
    if (err) return err;
    const long long total = forward ? xy_total_items<T, false>(a) : xy_total_items<T, true>(a);
    if (total <= 0) return 0;
    if (total > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
    long long grid = (long long)perSm * sms;
    if (grid > total) grid = total;
    cudaError_t e = cudaMemsetAsync(a.counters, 0, sizeof(int) * (1 + 2 * (size_t)a.y.numPlanes), s);
    if (e != cudaSuccess) return (int)e;
    if (forward)
      k_xy_fused<T, N, true><<<(unsigned)grid, C::threads, xy_smem<T, N>(), s>>>(a);
    else
      k_xy_fused<T, N, false><<<(unsigned)grid, C::threads, xy_smem<T, N>(), s>>>(a);
    return (int)cudaGetLastError();
  }
}

template <typename T>
static int launch_xy(int forward, const XYArgs<T>& a, cudaStream_t s) {
#define CALL(NN) return launch_xy_n<T, NN>(forward, a, s)
  SB_FAST_DISPATCH(a.x.nx, CALL)
#undef CALL
  return (int)cudaErrorInvalidValue;
}

}  // namespace sb

extern "C" {

int sb_xy_fused_config(int isFloat, int n, int numPlanes, int* ring, int* lag, int* numCounters) {
  int perSm = 0, sms = 0;
  int err = isFloat ? sb::xy_blocks_per_sm_n<float>(n, &perSm) : sb::xy_blocks_per_sm_n<double>(n, &perSm);
  if (err) return err;
  err = sb::num_sms(&sms);
  if (err) return err;
  const int lanes = isFloat ? 16 : 8;
  const int perStep = 2 * ((n + lanes - 1) / lanes);
  const int resident = perSm * sms;
  int l = (resident + perStep - 1) / perStep + 1;
  int r = 2 * l + 2;
  if (r < 6) r = 6;
  // debug knobs: SPFFT_B200_XY_LAG / SPFFT_B200_XY_RING override the defaults (ring > lag)
  if (const char* e = getenv("SPFFT_B200_XY_LAG")) l = atoi(e) > 0 ? atoi(e) : l;
  if (const char* e = getenv("SPFFT_B200_XY_RING")) r = atoi(e) > l ? atoi(e) : l + 1;
  if (numPlanes <= r) {
    r = numPlanes > 0 ? numPlanes : 1;  // every plane has its own slot: no reuse waits
  }
  *ring = r;
  *lag = l;
  *numCounters = 1 + 2 * (numPlanes > 0 ? numPlanes : 0);
  return 0;
}

#ifdef SB_XY_TRACE
/* experiment builds only (tools/xy_trace.py) */
__attribute__((visibility("default"))) int sb_xy_trace_read(long long* host, int maxEntries) {
  const int n = sb::kTraceCtas * sb::kTraceItems * sb::kTraceMarks;
  if (maxEntries < n) return -1;
  if (cudaMemcpyFromSymbol(host, sb::g_xy_trace, sizeof(long long) * n) != cudaSuccess) return -2;
  return n;
}
#endif

int sb_launch_xy_f64(int forward, const sb::XYArgs<double>* a, void* stream) {
  sb_note_launches(1);
  return sb::launch_xy<double>(forward, *a, static_cast<cudaStream_t>(stream));
}
int sb_launch_xy_f32(int forward, const sb::XYArgs<float>* a, void* stream) {
  sb_note_launches(1);
  return sb::launch_xy<float>(forward, *a, static_cast<cudaStream_t>(stream));
}
}
