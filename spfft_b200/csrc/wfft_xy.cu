// wfft_xy.cu -- fused xy stage on the warp FFT (wfft.hpp, wfft_kernels.cuh): y tiles and x tiles of
// every plane are items of ONE persistent kernel, the y <-> x hand-off plane lives in a small ring
// of scratch planes that stays resident in the 126 MB L2 (written once, read once, then dropped with
// discard.global.L2), so the stage moves only its algorithmic bytes through HBM:
//     backward:  sticks (sparse rows, gathered through the inverse map) -> y-FFT -> ring -> x-FFT -> space
//     forward :  space -> x-FFT -> ring -> y-FFT -> sticks (scattered through the inverse map)
// C2C, double precision, dimX == dimY == 512, one process-local slab of planes.
// Replaces the two passes of the reference's 2-D cuFFT plans (src/fft/transform_2d_gpu.hpp:51-140)
// and its transposing unpack / pack kernels (src/transpose/gpu_kernels/local_transpose_kernels.cu).
//
// Schedule: items in the order of xy_decode (fast_stage_kernels.hpp): for step u, the A tiles of
// plane u, then the B tiles of plane u - lag. Item i belongs to CTA i mod gridDim (static, no
// claim counter); the grid is launched cooperatively, so every CTA is resident and waiting on an
// earlier item cannot deadlock. A B tile waits until all A tiles of its plane are complete, an A
// tile until the B tiles of the plane that used its ring slot before are complete. Completion of
// item k is published while item k+1 of the same CTA runs (after the CTA barrier that item needs
// anyway), and the dependency of item k+1 is polled while the loads of item k are in flight, so in
// the steady state no warp ever waits on a flag.
#include <cstdlib>

#include "fast_launch.cuh"
#include "launch.h"
#include "wfft_kernels.cuh"

namespace sb {

namespace {

template <typename T>
struct WxyCounters {
  int* aDone;
  int* bDone;
  int nA, nB, ring;
  __device__ __forceinline__ bool ready(const XYItem& it) const {
    if (it.roleA) return it.plane < ring || w_ld_acquire(&bDone[it.plane - ring]) >= nB;
    return w_ld_acquire(&aDone[it.plane]) >= nA;
  }
};

// first valid item of this CTA's sequence at or after `i`
template <typename T, bool BWD>
__device__ __forceinline__ long long w_next_valid(const XYArgs<T>& a, long long i, long long total, XYItem& it) {
  while (i < total) {
    it = xy_decode<T, BWD>(a, (int)i);
    if (it.valid) break;
    i += gridDim.x;
  }
  return i;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// backward: A = y tile (8 columns of one plane: gather -> FFT -> transposed into S -> TMA store into
// the ring), B = x tile (8 rows of one plane, one per warp: ring -> FFT -> space domain)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kWThreads, 2)
    k_wxy_bwd(const __grid_constant__ XYArgs<T> a, const __grid_constant__ TensorMap ringMap,
              const __grid_constant__ WTw4<T> twp) {
  constexpr int N = kWN;
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  __shared__ int sReady[2];
  const int tid = threadIdx.x;
  const int w = tid >> 5, L = tid & 31;
  const WAddr ad = w_addr(w, L);
  __shared__ __align__(16) cx<T> sTw[4 * 32];
  w_stage_twiddles<T>(sTw, twp);
  __syncthreads();
  const int P = a.y.numPlanes;
  WxyCounters<T> dep;
  dep.aDone = a.counters + 1;
  dep.bDone = a.counters + 1 + P;
  dep.nA = a.y.numXTiles;
  dep.nB = N / kWWarps;
  dep.ring = a.ring;
  const long long total = xy_total_items<T, true>(a);
  const size_t planeElems = (size_t)N * N;

  int* pend = nullptr;   // counter of the previous item, not yet published (CTA-uniform)
  bool pendTma = false;  // ... whose output is a bulk tensor store issued by thread 0
  XYItem it, nx;
  long long cur = w_next_valid<T, true>(a, blockIdx.x, total, it);
  for (int k = 0; cur < total; ++k) {
    const long long nxt = w_next_valid<T, true>(a, cur + gridDim.x, total, nx);
    // ---- dependency of this item: normally seen satisfied one item ago
    if (k == 0 || !sReady[k & 1]) {
      __syncthreads();
      if (tid == 0) {
        if (pendTma) w_tma_wait_all();
        if (pend) {
          __threadfence();
          atomicAdd(pend, 1);
        }
        while (!dep.ready(it)) __nanosleep(100);
      }
      pend = nullptr;
      pendTma = false;
      __syncthreads();
    }
    // ---- loads
    cx<T> v[16];
    const int slot = it.plane % a.ring;
    if (it.roleA) {
      const int e0 = a.y.xtStart[it.tile];
      const cx<T>* row = a.y.sticks + (size_t)(it.plane + a.y.zRowOffset) * a.y.pitch + e0;
      const WInv16 iv = w_load_inv(a.y.inv, it.tile, w, L);
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        v[m] = mk<T>(0, 0);
        if (iv.i[m] != kWNone) v[m] = row[iv.i[m]];
      }
    } else {
      const cx<T>* src = a.scratch + (size_t)slot * planeElems + (size_t)(it.tile * kWWarps + w) * N + L;
#pragma unroll
      for (int m = 0; m < 16; ++m) v[m] = w_ldcg(src + 32 * m);
    }
    if (tid == 0) sReady[(k + 1) & 1] = nxt < total ? (dep.ready(nx) ? 1 : 0) : 1;
    w512_head<T, true>(v, L);
    if (!it.roleA) {
      // the consumed hand-off row (8 KB, fully read by now) is dropped from L2, not written back
      const char* rowBytes = reinterpret_cast<const char*>(a.scratch + (size_t)slot * planeElems +
                                                           (size_t)(it.tile * kWWarps + w) * N);
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(rowBytes + (size_t)L * 128) : "memory");
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(rowBytes + (size_t)(L + 32) * 128) : "memory");
    }
    // ---- S is free once the previous tile's tensor store has read it; publish the previous item
    if (tid == 0 && pendTma) w_tma_wait_all();
    __syncthreads();
    if (tid == 0 && pend) {
      __threadfence();
      atomicAdd(pend, 1);
    }
    w512_exchange<T>(v, S, ad);
    w512_tail<T, true>(v, sTw, L);
    if (it.roleA) {
      w512_col_store<T>(v, S, ad);
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        tma_store_3d(&ringMap, it.tile * 16, 0, slot, S);
        tma_store_3d(&ringMap, it.tile * 16, 256, slot, S + 256 * 8);
        tma_store_commit();
      }
      pend = &dep.aDone[it.plane];
      pendTma = true;
    } else {
      cx<T>* dst = static_cast<cx<T>*>(a.x.spaceOut) + (size_t)it.plane * planeElems +
                   (size_t)(it.tile * kWWarps + w) * N + L;
#pragma unroll
      for (int m = 0; m < 16; ++m) dst[32 * m] = v[m];
      pend = &dep.bDone[it.plane];
      pendTma = false;
    }
    cur = nxt;
    it = nx;
  }
  // publish the CTA's last item (tiles of other CTAs may wait for it); the bulk store of the last y
  // tile must have read S before the CTA exits
  __syncthreads();
  if (tid == 0) {
    if (pendTma) w_tma_wait_all();
    if (pend) {
      __threadfence();
      atomicAdd(pend, 1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward: A = x tile (8 rows: space -> FFT -> ring), B = y tile (TMA load of 8 ring columns into S ->
// FFT -> scatter into the stick rows)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kWThreads, 2)
    k_wxy_fwd(const __grid_constant__ XYArgs<T> a, const __grid_constant__ TensorMap ringMap,
              const __grid_constant__ WTw4<T> twp) {
  constexpr int N = kWN;
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  __shared__ int sReady[2];
  __shared__ __align__(8) uint64_t full;
  const int tid = threadIdx.x;
  const int w = tid >> 5, L = tid & 31;
  const WAddr ad = w_addr(w, L);
  __shared__ __align__(16) cx<T> sTw[4 * 32];
  w_stage_twiddles<T>(sTw, twp);
  __syncthreads();
  if (tid == 0) {
    mbar_init(&full, 1);
    mbar_fence_init();
  }
  const int P = a.y.numPlanes;
  WxyCounters<T> dep;
  dep.aDone = a.counters + 1;
  dep.bDone = a.counters + 1 + P;
  dep.nA = N / kWWarps;
  dep.nB = a.y.numXTiles;
  dep.ring = a.ring;
  const long long total = xy_total_items<T, false>(a);
  const size_t planeElems = (size_t)N * N;

  int* pend = nullptr;
  uint32_t phase = 0;
  bool preloaded = false;  // the tile of the current (B) item is already on its way into S
  XYItem it, nx;
  long long cur = w_next_valid<T, false>(a, blockIdx.x, total, it);
  __syncthreads();
  for (int k = 0; cur < total; ++k) {
    const long long nxt = w_next_valid<T, false>(a, cur + gridDim.x, total, nx);
    if (k == 0 || !sReady[k & 1]) {
      __syncthreads();
      if (tid == 0) {
        if (pend) {
          __threadfence();
          atomicAdd(pend, 1);
        }
        while (!dep.ready(it)) __nanosleep(100);
      }
      pend = nullptr;
      __syncthreads();
    }
    cx<T> v[16];
    const int slot = it.plane % a.ring;
    if (it.roleA) {
      const cx<T>* src = static_cast<const cx<T>*>(a.x.spaceIn) + (size_t)it.plane * planeElems +
                         (size_t)(it.tile * kWWarps + w) * N + L;
#pragma unroll
      for (int m = 0; m < 16; ++m) v[m] = src[32 * m];
      if (tid == 0) sReady[(k + 1) & 1] = nxt < total ? (dep.ready(nx) ? 1 : 0) : 1;
    } else {
      // every warp is past the barrier of the previous item, i.e. past its last access of S
      if (!preloaded && tid == 0) {
        w_fence_proxy_async();
        mbar_expect_tx(&full, (uint32_t)kWTileBytes);
        tma_load_3d(S, &ringMap, it.tile * 16, 0, slot, &full);
        tma_load_3d(S + 256 * 8, &ringMap, it.tile * 16, 256, slot, &full);
      }
      if (tid == 0) sReady[(k + 1) & 1] = nxt < total ? (dep.ready(nx) ? 1 : 0) : 1;
      mbar_wait(&full, phase);
      phase ^= 1;
      w512_col_load<T>(v, S, ad);
      __syncwarp();
      // the consumed column segments (one 128-byte line per row) are dropped from L2
      const char* tileBytes = reinterpret_cast<const char*>(a.scratch + (size_t)slot * planeElems + (size_t)it.tile * kWWarps);
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(tileBytes + (size_t)tid * N * sizeof(cx<T>)) : "memory");
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(tileBytes + (size_t)(tid + 256) * N * sizeof(cx<T>)) : "memory");
    }
    w512_head<T, false>(v, L);
    w512_exchange<T>(v, S, ad);
    // ---- every warp is done with S; all stores of the previous item were issued before this point
    __syncthreads();
    preloaded = nxt < total && !nx.roleA && sReady[(k + 1) & 1];
    if (tid == 0) {
      if (pend) {
        __threadfence();
        atomicAdd(pend, 1);
      }
      if (preloaded) {
        w_fence_proxy_async();
        const int nslot = nx.plane % a.ring;
        mbar_expect_tx(&full, (uint32_t)kWTileBytes);
        tma_load_3d(S, &ringMap, nx.tile * 16, 0, nslot, &full);
        tma_load_3d(S + 256 * 8, &ringMap, nx.tile * 16, 256, nslot, &full);
      }
    }
    w512_tail<T, false>(v, sTw, L);
    if (it.roleA) {
      cx<T>* dst = a.scratch + (size_t)slot * planeElems + (size_t)(it.tile * kWWarps + w) * N + L;
#pragma unroll
      for (int m = 0; m < 16; ++m) w_stcg(dst + 32 * m, v[m]);
      pend = &dep.aDone[it.plane];
    } else {
      const int e0 = a.y.xtStart[it.tile];
      cx<T>* row = a.y.sticks + (size_t)(it.plane + a.y.zRowOffset) * a.y.pitch + e0;
      const WInv16 iv = w_load_inv(a.y.inv, it.tile, w, L);
#pragma unroll
      for (int m = 0; m < 16; ++m)
        if (iv.i[m] != kWNone) row[iv.i[m]] = v[m];
      pend = &dep.bDone[it.plane];
    }
    cur = nxt;
    it = nx;
  }
  // publish the CTA's last item (tiles of other CTAs may wait for it)
  __syncthreads();
  if (tid == 0 && pend) {
    __threadfence();
    atomicAdd(pend, 1);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace {

int wxy_grid(int* gridOut) {
  static int cached = 0;
  if (cached > 0) {
    *gridOut = cached;
    return 0;
  }
  cudaError_t e = cudaFuncSetAttribute(k_wxy_bwd<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWTileBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_wxy_fwd<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWTileBytes);
  if (e != cudaSuccess) return (int)e;
  int b0 = 0, b1 = 0, dev = 0, sms = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_wxy_bwd<double>, kWThreads, kWTileBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_wxy_fwd<double>, kWThreads, kWTileBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  int coop = 0;
  e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (e != cudaSuccess) return (int)e;
  const int perSm = b0 < b1 ? b0 : b1;
  if (perSm <= 0 || !coop) return (int)cudaErrorInvalidConfiguration;
  cached = perSm * sms;
  *gridOut = cached;
  return 0;
}

template <typename Kernel>
int wxy_launch(Kernel kernel, int grid, const XYArgs<double>& a, const TensorMap& map, const WTw4<double>& tw,
               cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kWThreads);
  cfg.dynamicSmemBytes = kWTileBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kernel, a, map, tw);
}

}  // namespace
}  // namespace sb

extern "C" {

int sb_wxy_config(int isFloat, int n, int numPlanes, int* ring, int* lag, int* numCounters) {
  if (isFloat || n != sb::kWN) return (int)cudaErrorInvalidValue;
  int grid = 0;
  const int err = sb::wxy_grid(&grid);
  if (err) return err;
  const int perStep = 2 * (n / sb::kWWarps);
  int l = (grid + perStep - 1) / perStep + 1;
  int r = 2 * l + 2;
  if (const char* e = getenv("SPFFT_B200_XY_LAG")) l = atoi(e) > 0 ? atoi(e) : l;
  if (const char* e = getenv("SPFFT_B200_XY_RING")) r = atoi(e) > l ? atoi(e) : l + 1;
  if (numPlanes <= r) r = numPlanes > 0 ? numPlanes : 1;  // every plane has its own slot: no reuse waits
  *ring = r;
  *lag = l;
  *numCounters = 1 + 2 * (numPlanes > 0 ? numPlanes : 0);
  return 0;
}

int sb_launch_wxy_f64(int forward, const sb::XYArgs<double>* args, void* stream) {
  using namespace sb;
  const XYArgs<double>& a = *args;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a.y.numPlanes <= 0) return 0;
  if (a.x.nx != kWN || a.y.ny != kWN || a.y.nxf != kWN || !a.y.inv || a.y.srcBase) return (int)cudaErrorInvalidValue;
  int grid = 0;
  int err = wxy_grid(&grid);
  if (err) return err;
  const long long total = forward ? xy_total_items<double, false>(a) : xy_total_items<double, true>(a);
  if (total > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  if (grid > total) grid = (int)total;
  static const WTw4<double> tw = [] {
    WTw4<double> t;
    wfft_lane_twiddles<double>(kWN, 32, &t.w[0][0]);
    return t;
  }();
  TensorMap map;
  err = make_tile_map(&map, a.scratch, sizeof(cx<double>), kWN, kWN, kWN, a.ring, (long long)kWN * kWN, 8, 256);
  if (err) return err;
  cudaError_t e = cudaMemsetAsync(a.counters, 0, sizeof(int) * (1 + 2 * (size_t)a.y.numPlanes), s);
  if (e != cudaSuccess) return (int)e;
  sb_note_launches(1);
  return forward ? wxy_launch(k_wxy_fwd<double>, grid, a, map, tw, s) : wxy_launch(k_wxy_bwd<double>, grid, a, map, tw, s);
}
}
