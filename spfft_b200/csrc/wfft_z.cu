// wfft_z.cu -- z stage on the warp FFT (wfft.hpp, wfft_kernels.cuh), double precision, dimZ == 512,
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
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  const int tid = threadIdx.x;
  const int w = tid >> 5, L = tid & 31;
  const WAddr ad = w_addr<8>(w, L);
  __shared__ __align__(16) cx<T> sTw[4 * 32];
  w_stage_twiddles<T>(sTw, twp);
  __syncthreads();
  // inverse-map entries and value range of the thread's NEXT tile, fetched one tile ahead (cp.async: no registers):
  // the gather of a tile then costs one HBM round trip instead of two dependent ones
  __shared__ __align__(16) uint4 sInv[2][2][kWThreads];
  pdl_prologue();
  int tile = blockIdx.x;
  int e0 = tile < a.numTiles ? a.tileStart[tile] : 0;
  if (tile < a.numTiles) {
    const unsigned short* p = w_inv_ptr(a.inv, tile, w, L);
    w_cp_async16(&sInv[0][0][tid], p);
    w_cp_async16(&sInv[0][1][tid], p + 32 * 8);
  }
  for (int k = 0; tile < a.numTiles; tile += gridDim.x, ++k) {
    const cx<T>* vals = a.valuesIn + e0;
    w_cp_async_wait();
    const WInv16 iv = w_unpack_inv(sInv[k & 1][0][tid], sInv[k & 1][1][tid]);
    cx<T> v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      v[m] = mk<T>(0, 0);
      if (iv.i[m] != kWNone) v[m] = w_ldcs(vals + iv.i[m]);
    }
    const int next = tile + gridDim.x;
    if (next < a.numTiles) {
      const unsigned short* p = w_inv_ptr(a.inv, next, w, L);
      w_cp_async16(&sInv[(k + 1) & 1][0][tid], p);
      w_cp_async16(&sInv[(k + 1) & 1][1][tid], p + 32 * 8);
      e0 = a.tileStart[next];
    }
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
    w512_head<T, true>(v, L);
    if (tid == 0) tma_store_wait_read();  // the previous tile's store has read S
    __syncthreads();
    w512_exchange<T, 8>(v, S, ad);
    w512_tail<T, true>(v, sTw, L);
    w512_col_store<T, 8>(v, S, ad);
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
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  __shared__ __align__(8) uint64_t full;
  const int tid = threadIdx.x;
  const int w = tid >> 5, L = tid & 31;
  const WAddr ad = w_addr<8>(w, L);
  __shared__ __align__(16) cx<T> sTw[4 * 32];
  w_stage_twiddles<T>(sTw, twp);
  __syncthreads();
  if (tid == 0) {
    mbar_init(&full, 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_prologue();
  uint32_t phase = 0;
  int tile = blockIdx.x;
  if (tile < a.numTiles && tid == 0) {
    mbar_expect_tx(&full, (uint32_t)kWTileBytes);
    tma_load_3d(S, &stickMap, tile * 16, 0, 0, &full);
    tma_load_3d(S + 256 * 8, &stickMap, tile * 16, 256, 0, &full);
  }
  for (; tile < a.numTiles; tile += gridDim.x) {
    const WInv16 iv = w_load_inv(a.inv, tile, w, L);
    cx<T>* out = a.valuesOut + a.tileStart[tile];
    mbar_wait(&full, phase);
    phase ^= 1;
    cx<T> v[16];
    w512_col_load<T, 8>(v, S, ad);
    __syncwarp();
    w512_head<T, false>(v, L);
    w512_exchange<T, 8>(v, S, ad);
    __syncthreads();  // every warp is done with S
    const int next = tile + gridDim.x;
    if (next < a.numTiles && tid == 0) {
      mbar_expect_tx(&full, (uint32_t)kWTileBytes);
      tma_load_3d(S, &stickMap, next * 16, 0, 0, &full);
      tma_load_3d(S + 256 * 8, &stickMap, next * 16, 256, 0, &full);
    }
    w512_tail<T, false>(v, sTw, L);
    if (a.useScale) {
#pragma unroll
      for (int m = 0; m < 16; ++m) v[m] = a.scale * v[m];
    }
#pragma unroll
    for (int m = 0; m < 16; ++m)
      if (iv.i[m] != kWNone) out[iv.i[m]] = v[m];
  }
}

namespace {
int wz_grid(int* gridOut) {
  static int cached = 0;
  if (cached > 0) {
    *gridOut = cached;
    return 0;
  }
  cudaError_t e = cudaFuncSetAttribute(k_wz_bwd<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWTileBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_wz_fwd<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWTileBytes);
  if (e != cudaSuccess) return (int)e;
  int b0 = 0, b1 = 0, dev = 0, sms = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_wz_bwd<double>, kWThreads, kWTileBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_wz_fwd<double>, kWThreads, kWTileBytes);
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
}  // namespace
}  // namespace sb

extern "C" {

/* Is the warp-FFT z stage applicable (double, nz == 512, inverse-map form, local row layout)? */
int sb_wz_available(int isFloat, int nz) { return !isFloat && nz == sb::kWN; }

int sb_launch_wz_f64(int forward, const sb::ZArgs<double>* args, void* stream) {
  using namespace sb;
  const ZArgs<double>& a = *args;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a.numTiles <= 0) return 0;
  if (a.nz != kWN || !a.inv || a.rowRank || a.wireF32 || a.log2V != 3) return (int)cudaErrorInvalidValue;
  int grid = 0;
  int err = wz_grid(&grid);
  if (err) return err;
  if (grid > a.numTiles) grid = a.numTiles;
  static const WTw4<double> tw = [] {
    WTw4<double> t;
    wfft_lane_twiddles<double>(kWN, 32, &t.w[0][0]);
    return t;
  }();
  TensorMap map;
  err = make_tile_map(&map, a.sticks, sizeof(cx<double>), a.pitch, kWN, a.pitch, 1, 0, 8, 256);
  if (err) return err;
  sb_note_launches(1);
  if (forward) return launch_stage_kernel(k_wz_fwd<double>, dim3((unsigned)grid), kWThreads, kWTileBytes, s, a, map, tw);
  return launch_stage_kernel(k_wz_bwd<double>, dim3((unsigned)grid), kWThreads, kWTileBytes, s, a, map, tw);
}
}
