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

// All global writes of this tile (every thread's) become visible before the counter moves:
// barrier (CTA-scope ordering), then one cumulative gpu-scope fence + atomic by thread 0 -- the
// pattern of a cooperative-groups grid barrier.
__device__ __forceinline__ void cta_signal(int* p) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(p, 1);
  }
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
  __shared__ int sQ[2];
  constexpr bool BWD = !FWD;
  constexpr int V = FastCfg<T, N>::V;
  const int P = a.y.numPlanes;
  const int nA = xy_tiles_a<T, BWD>(a), nB = xy_tiles_b<T, BWD>(a);
  const long long total = xy_total_items<T, BWD>(a);
  int* aDone = a.counters + 1;
  int* bDone = a.counters + 1 + P;
  // work items are claimed two ahead: the next item is known while the current one runs, so its
  // input can be prefetched into L2, and the claim's round trip is off the critical path
  if (threadIdx.x == 0) {
    sQ[0] = atomicAdd(&a.counters[0], 1);
    sQ[1] = atomicAdd(&a.counters[0], 1);
  }
  __syncthreads();
  int cur = sQ[0], nxt = sQ[1];
  __syncthreads();
  for (int k = 0; cur < total; ++k) {
    if (threadIdx.x == 0) sQ[k & 1] = atomicAdd(&a.counters[0], 1);
    const XYItem it = xy_decode<T, BWD>(a, cur);
    XYItem nx;
    nx.valid = false;
    if (nxt < total) nx = xy_decode<T, BWD>(a, nxt);
    if (it.valid) {
      if (it.roleA) {
        // slot reuse: the B tiles of plane - ring must have read the slot before this tile's final
        // store phase. Thread 0 waits here; the barriers inside the tile order every thread's
        // stores after it (no extra barrier).
        if (it.plane >= a.ring && threadIdx.x == 0) {
          while (ld_acquire_gpu(&bDone[it.plane - a.ring]) < nB) __nanosleep(64);
        }
        xy_run_item<T, N, BWD>(a, it, nx, tws, Ctx{FastCfg<T, N>::threads}, S);
        cta_signal(&aDone[it.plane]);
      } else {
        // the first thing a B tile does is read the hand-off plane: every warp waits on its own
        // (one polling lane per warp) instead of a CTA-wide barrier
        if ((threadIdx.x & 31) == 0) {
          while (ld_acquire_gpu(&aDone[it.plane]) < nA) __nanosleep(64);
        }
        __syncwarp();
        xy_run_item<T, N, BWD>(a, it, nx, tws, Ctx{FastCfg<T, N>::threads}, S);
        // all loads of the tile are complete (their values were consumed before the tile's first
        // barrier): drop the dirty hand-off lines from L2 instead of writing them back to HBM
        {
          const cx<T>* slot = a.scratch + (size_t)(it.plane % a.ring) * N * N;
          if (BWD) {  // x tile: V whole rows, contiguous
            discard_l2(slot + (size_t)it.tile * V * N, sizeof(cx<T>) * V * N, threadIdx.x, blockDim.x);
          } else {    // y tile: one 128-byte column segment per row
            for (int y = threadIdx.x; y < N; y += blockDim.x)
              discard_l2(slot + (size_t)y * N + (size_t)it.tile * V, sizeof(cx<T>) * V, 0, 1);
          }
        }
        cta_signal(&bDone[it.plane]);
      }
    } else {
      __syncthreads();
    }
    // (cta_signal's barrier doubles as the end-of-item barrier: sQ[k & 1] is visible, S is free)
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

int sb_launch_xy_f64(int forward, const sb::XYArgs<double>* a, void* stream) {
  sb_note_launches(1);
  return sb::launch_xy<double>(forward, *a, static_cast<cudaStream_t>(stream));
}
int sb_launch_xy_f32(int forward, const sb::XYArgs<float>* a, void* stream) {
  sb_note_launches(1);
  return sb::launch_xy<float>(forward, *a, static_cast<cudaStream_t>(stream));
}
}
