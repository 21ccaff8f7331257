// wfft_z.cu -- z stage on the warp FFT (wfft.hpp, wfft_kernels.cuh), dimZ == 512; double precision, and single
// precision with two sticks per warp (WUnit: an item = 16 sticks = two tiles of the index plan, C2C only);
// values given in stick order (inverse-map form, index_plan.cpp):
//   backward: every warp gathers the values of ONE stick through the inverse map straight into
//             registers (absent = 0; stick (0,0) of an R2C transform completed by conjugation on the
//             fly), z-FFT, the 8 sticks of the tile meet in S as [512 z][8 sticks] and leave as ONE
//             bulk tensor store into the plane-major stick buffer [z][pitch]
//   forward : ONE bulk tensor load of the tile [512 z][8 sticks], z-FFT per warp, values scaled and
//             stored through the inverse map
// Replaces decompress / compress (src/compression/gpu_kernels/compression_kernels.cu:40-150), the
// stick symmetry kernel (src/symmetry/gpu_kernels/symmetry_kernels.cu:39-88) and the z cuFFT plan
// (src/fft/transform_1d_gpu.hpp:52-141) of the reference.
#include "fast_launch.cuh"
#include "launch.h"
#include "wfft_kernels.cuh"

namespace sb {

template <typename T>
__global__ void __launch_bounds__(kWThreads, 2)
    k_wz_bwd(const __grid_constant__ ZArgs<T> a, const __grid_constant__ TensorMap stickMap,
             const __grid_constant__ WTw4<T> twp) {
  constexpr int N = kWN;
  using Sc = typename WUnit<T>::Sc;
  constexpr int kPer = WUnit<T>::kPer;  // sticks per warp; an item = 8 kPer sticks = kPer tiles of the index plan
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  cx<Sc>* S = reinterpret_cast<cx<Sc>*>(smemRaw);
  const int tid = threadIdx.x;
  const int w = tid >> 5, L = tid & 31;
  const WAddr ad = w_addr<8>(w, L);
  __shared__ __align__(16) cx<Sc> sTw[4 * 32];
  w_stage_twiddles(sTw, twp);
  __syncthreads();
  // inverse-map entries and value range of the thread's NEXT tile, fetched one tile ahead (cp.async: no registers):
  // the gather of a tile then costs one HBM round trip instead of two dependent ones
  __shared__ __align__(16) uint4 sInv[2][2 * kPer][kWThreads];
  pdl_prologue();
  const int numItems = (a.numTiles + kPer - 1) / kPer;
  const int planLane = w_plan_lane<kPer>(w);
  int tile = blockIdx.x;
  // (single precision, odd number of plan tiles: the upper half of the last item has no sticks)
  bool have = tile < numItems && w_plan_tile<kPer>(tile, w) < a.numTiles;
  int e0 = have ? a.tileStart[w_plan_tile<kPer>(tile, w)] : 0;
  if (have) w_inv_prefetch<kPer>(sInv[0], a.inv, w_plan_tile<kPer>(tile, w), planLane, L, tid);
  for (int k = 0; tile < numItems; tile += gridDim.x, ++k) {
    const cx<T>* vals = a.valuesIn + e0;
    w_cp_async_wait();
    cx<Sc> v[16];
    if (have) {
      const WInvUnit<kPer> iv = w_inv_read<kPer>(sInv[k & 1], tid);
      w_gather_cs(v, vals, iv);
    } else {
#pragma unroll
      for (int m = 0; m < 16; ++m) v[m] = w_zero<Sc>();
    }
    const int next = tile + gridDim.x;
    const bool haveNext = next < numItems && w_plan_tile<kPer>(next, w) < a.numTiles;
    if (haveNext) {
      w_inv_prefetch<kPer>(sInv[(k + 1) & 1], a.inv, w_plan_tile<kPer>(next, w), planLane, L, tid);
      e0 = a.tileStart[w_plan_tile<kPer>(next, w)];
    }
    have = haveNext;
    if constexpr (kPer == 1) {
      // (double precision only; single-precision R2C transforms keep the round-1 z kernels)
      if (tile == a.symTile && w == a.symLane) {
        // hermitian completion of stick (0,0) (reference: symmetry_host.hpp:47-58): element n also
        // looks at the given value at N - n
        const unsigned short* invCol = a.inv + ((size_t)tile * 512 + (size_t)w * 64) * 8;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int n = L + 32 * m;
          const int n2 = (N - n) & (N - 1);
          const unsigned short i2 = invCol[(size_t)(n2 & 63) * 8 + (n2 >> 6)];
          const cx<T> q = i2 != kWNone ? vals[i2] : mk<T>(0, 0);
          v[m] = hermitian_combine<T>(n, N, v[m], q);
        }
      }
    }
    w512_head<Sc, true>(v, L);
    if (tid == 0) tma_store_wait_read();  // the previous tile's store has read S
    __syncthreads();
    w512_exchange<Sc, 8>(v, S, ad);
    w512_tail<Sc, true>(v, sTw, L);
    w512_col_store<Sc, 8>(v, S, ad);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tma_store_3d(&stickMap, tile * 16, 0, 0, S);
      tma_store_3d(&stickMap, tile * 16, 256, 0, S + 256 * 8);
      tma_store_commit();
    }
  }
  if (tid == 0) tma_store_wait_read();
}

template <typename T>
__global__ void __launch_bounds__(kWThreads, 2)
    k_wz_fwd(const __grid_constant__ ZArgs<T> a, const __grid_constant__ TensorMap stickMap,
             const __grid_constant__ WTw4<T> twp) {
  using Sc = typename WUnit<T>::Sc;
  constexpr int kPer = WUnit<T>::kPer;
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  cx<Sc>* S = reinterpret_cast<cx<Sc>*>(smemRaw);
  __shared__ __align__(8) uint64_t full;
  const int tid = threadIdx.x;
  const int w = tid >> 5, L = tid & 31;
  const WAddr ad = w_addr<8>(w, L);
  __shared__ __align__(16) cx<Sc> sTw[4 * 32];
  w_stage_twiddles(sTw, twp);
  __syncthreads();
  if (tid == 0) {
    mbar_init(&full, 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_prologue();
  uint32_t phase = 0;
  const int numItems = (a.numTiles + kPer - 1) / kPer;
  const int planLane = w_plan_lane<kPer>(w);
  int tile = blockIdx.x;
  if (tile < numItems && tid == 0) {
    mbar_expect_tx(&full, (uint32_t)kWTileBytes);
    tma_load_3d(S, &stickMap, tile * 16, 0, 0, &full);
    tma_load_3d(S + 256 * 8, &stickMap, tile * 16, 256, 0, &full);
  }
  for (; tile < numItems; tile += gridDim.x) {
    const int planTile = w_plan_tile<kPer>(tile, w);
    const bool have = planTile < a.numTiles;  // (single precision: the last item may cover one plan tile only)
    WInvUnit<kPer> iv;
    cx<T>* out = a.valuesOut;
    if (have) {
      iv = w_inv_load<kPer>(a.inv, planTile, planLane, L);
      out += a.tileStart[planTile];
    }
    mbar_wait(&full, phase);
    phase ^= 1;
    cx<Sc> v[16];
    w512_col_load<Sc, 8>(v, S, ad);
    __syncwarp();
    w512_head<Sc, false>(v, L);
    w512_exchange<Sc, 8>(v, S, ad);
    __syncthreads();  // every warp is done with S
    const int next = tile + gridDim.x;
    if (next < numItems && tid == 0) {
      mbar_expect_tx(&full, (uint32_t)kWTileBytes);
      tma_load_3d(S, &stickMap, next * 16, 0, 0, &full);
      tma_load_3d(S + 256 * 8, &stickMap, next * 16, 256, 0, &full);
    }
    w512_tail<Sc, false>(v, sTw, L);
    if (a.useScale) {
      const Sc sc = Sc((double)a.scale);
#pragma unroll
      for (int m = 0; m < 16; ++m) v[m] = sc * v[m];
    }
    if (have) w_scatter_plain(out, v, iv);
  }
}

namespace {
template <typename T>
int wz_grid(int* gridOut) {
  static int cached = 0;
  if (cached > 0) {
    *gridOut = cached;
    return 0;
  }
  cudaError_t e = cudaFuncSetAttribute(k_wz_bwd<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWTileBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_wz_fwd<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWTileBytes);
  if (e != cudaSuccess) return (int)e;
  int b0 = 0, b1 = 0, dev = 0, sms = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_wz_bwd<T>, kWThreads, kWTileBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_wz_fwd<T>, kWThreads, kWTileBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  const int perSm = b0 < b1 ? b0 : b1;
  if (perSm <= 0) return (int)cudaErrorInvalidConfiguration;
  cached = perSm * sms;
  *gridOut = cached;
  return 0;
}

template <typename T>
int wz_run(int forward, const ZArgs<T>& a, cudaStream_t s) {
  constexpr int kPer = WUnit<T>::kPer;
  if (a.numTiles <= 0) return 0;
  if (a.nz != kWN || !a.inv || a.rowRank || a.wireF32 || a.log2V != 3) return (int)cudaErrorInvalidValue;
  // single precision: C2C (no stick to complete), stick rows of whole 16-byte units
  if (kPer == 2 && (a.symTile >= 0 || (a.pitch & 1) || (reinterpret_cast<size_t>(a.sticks) & 15))) return (int)cudaErrorInvalidValue;
  int grid = 0;
  int err = wz_grid<T>(&grid);
  if (err) return err;
  const int numItems = (a.numTiles + kPer - 1) / kPer;
  if (grid > numItems) grid = numItems;
  static const WTw4<T> tw = [] {
    WTw4<T> t;
    wfft_lane_twiddles<T>(kWN, 32, &t.w[0][0]);
    return t;
  }();
  // the stick buffer [z][pitch] as rows of 16-byte units: box = 8 units x 256 rows in both precisions
  TensorMap map;
  err = make_tile_map(&map, a.sticks, 16, a.pitch / kPer, kWN, a.pitch / kPer, 1, 0, 8, 256);
  if (err) return err;
  sb_note_launches(1);
  if (forward) return launch_stage_kernel(k_wz_fwd<T>, dim3((unsigned)grid), kWThreads, kWTileBytes, s, a, map, tw);
  return launch_stage_kernel(k_wz_bwd<T>, dim3((unsigned)grid), kWThreads, kWTileBytes, s, a, map, tw);
}
}  // namespace
}  // namespace sb

extern "C" {

/* Is the warp-FFT z stage applicable (nz == 512, inverse-map form, local row layout)? */
int sb_wz_available(int /*isFloat*/, int nz) { return nz == sb::kWN; }

int sb_launch_wz_f64(int forward, const sb::ZArgs<double>* args, void* stream) {
  return sb::wz_run<double>(forward, *args, static_cast<cudaStream_t>(stream));
}
int sb_launch_wz_f32(int forward, const sb::ZArgs<float>* args, void* stream) {
  return sb::wz_run<float>(forward, *args, static_cast<cudaStream_t>(stream));
}
}
