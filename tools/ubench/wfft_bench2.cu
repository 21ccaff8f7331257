// wfft_bench2.cu -- SM-side cost of the warp-FFT tile structures (experiments only, not part of the library).
// Every kernel transforms `vplanes` virtual 512 x 512 planes; plane p lives at physical plane p % inMod of the
// input array and p % outMod of the output array, so that either side can be made L2-resident (mod = 8: 32 MB)
// or HBM-resident (mod = vplanes) independently -- the two regimes of the fused xy stage, whose hand-off side
// stays in L2 while the other side streams through HBM.
//   x<WARPS,MINB>      : autonomous warps, one row each: LDG -> FFT -> STG (private 8 KB exchange region)
//   xt<WARPS>          : autonomous warps with per-warp bulk-copy pipelines: cp.async.bulk row -> smem (2 buffers
//                        per warp, next row prefetched), FFT in place, bulk store
//   yb<W,MINB> / yf<W,MINB>: W warps per CTA = tile [512][W columns] (128 / 64 / 32-byte tile rows):
//                        contiguous columns in -> FFT -> tile out through a bulk tensor store (yb), and the mirror (yf)
// Usage: wfft_bench2 [vplanes] [reps]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tma_util.hpp"
#include "wfft.hpp"

using namespace sb;
using T = double;
constexpr int N = 512;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

struct Tw4 {
  cx<T> w[4][32];
};

// twiddles staged in shared memory, re-read per transform (see wfft_kernels.cuh)
#define STAGE_TW                                                                      \
  __shared__ __align__(16) cx<T> sTw[128];                                            \
  for (int i_ = threadIdx.x; i_ < 128; i_ += blockDim.x) sTw[i_] = twp.w[i_ >> 5][i_ & 31]; \
  __syncthreads();
__device__ __forceinline__ LaneTw<T> lane_tw(const cx<T>* sTw, int L) {
  LaneTw<T> r;
  r.w1 = sTw[L];
  r.w2 = sTw[32 + L];
  r.w4 = sTw[64 + L];
  r.w8 = sTw[96 + L];
  return r;
}
template <bool BWD>
__device__ __forceinline__ void tail(cx<T>* v, const cx<T>* sTw, int L) {
  const LaneTw<T> tw = lane_tw(sTw, L);
  twiddle_dft16<T, BWD>(v, tw);
}

template <bool BWD>
__device__ __forceinline__ void head(cx<T>* v, int L) {
  using P = WPlan<T, 512>;
  P::template stage_a_local<BWD>(v, L);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const cx<T> recv = shfl_xor_cx<T>(v[8 + i], 16);
    P::stage_a_combine(v[i], v[8 + i], recv, L);
  }
}

// private exchange region of 512 elements, XOR swizzle of WPlan
__device__ __forceinline__ void exchange_private(cx<T>* v, cx<T>* X, int L) {
  using P = WPlan<T, 512>;
#pragma unroll
  for (int i = 0; i < 16; ++i) X[P::xw(L, i)] = v[i];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = X[P::xr(L, r)];
  __syncwarp();
}

// column w of a tile [512][W]
template <int W>
__device__ __forceinline__ cx<T>* wcolw(cx<T>* S, int s, int w) {
  if constexpr (W == 8) return S + (s << 3) + (w ^ (s & 7));
  if constexpr (W == 4) return S + (s << 2) + (w ^ ((s >> 1) & 3));
  return S + (s << 1) + (w ^ ((s >> 2) & 1));
}
template <int W>
__device__ __forceinline__ void exchange_col(cx<T>* v, cx<T>* S, int w, int L) {
  using P = WPlan<T, 512>;
#pragma unroll
  for (int i = 0; i < 16; ++i) *wcolw<W>(S, P::xw(L, i), w) = v[i];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = *wcolw<W>(S, P::xr(L, r), w);
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    k_x(const cx<T>* in, cx<T>* out, const __grid_constant__ Tw4 twp, int vplanes, int inMod, int outMod) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, L = threadIdx.x & 31;
  cx<T>* X = reinterpret_cast<cx<T>*>(smem) + warp * N;
  STAGE_TW
  const long long rows = (long long)vplanes * N;
  for (long long row = (long long)blockIdx.x * WARPS + warp; row < rows; row += (long long)gridDim.x * WARPS) {
    const int p = (int)(row / N), y = (int)(row % N);
    const cx<T>* src = in + ((size_t)(p % inMod) * N + y) * N + L;
    cx<T> v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = src[32 * m];
    head<true>(v, L);
    exchange_private(v, X, L);
    tail<true>(v, sTw, L);
    cx<T>* dst = out + ((size_t)(p % outMod) * N + y) * N + L;
#pragma unroll
    for (int m = 0; m < 16; ++m) dst[32 * m] = v[m];
  }
}

// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes) : "memory");
}

template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    k_xt(const cx<T>* in, cx<T>* out, const __grid_constant__ Tw4 twp, int vplanes, int inMod, int outMod) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[WARPS][2];
  const int warp = threadIdx.x >> 5, L = threadIdx.x & 31;
  cx<T>* B0 = reinterpret_cast<cx<T>*>(smem) + (size_t)warp * 2 * N;
  STAGE_TW
  if (L == 0) {
    mbar_init(&bars[warp][0], 1);
    mbar_init(&bars[warp][1], 1);
    mbar_fence_init();
  }
  __syncwarp();
  const long long rows = (long long)vplanes * N;
  const long long stride = (long long)gridDim.x * WARPS;
  long long row = (long long)blockIdx.x * WARPS + warp;
  auto src_of = [&](long long r) { return in + ((size_t)((int)(r / N) % inMod) * N + (int)(r % N)) * N; };
  auto dst_of = [&](long long r) { return out + ((size_t)((int)(r / N) % outMod) * N + (int)(r % N)) * N; };
  uint32_t ph[2] = {0, 0};
  if (row < rows && L == 0) {
    mbar_expect_tx(&bars[warp][0], N * sizeof(cx<T>));
    bulk_load(B0, src_of(row), N * sizeof(cx<T>), &bars[warp][0]);
  }
  for (int k = 0; row < rows; row += stride, ++k) {
    const int b = k & 1;
    cx<T>* B = B0 + b * N;
    mbar_wait(&bars[warp][b], ph[b]);
    ph[b] ^= 1;
    cx<T> v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = B[L + 32 * m];
    __syncwarp();
    head<true>(v, L);
    // the other buffer: its store (row k-1) has been read by now -> prefetch row k+1 into it
    if (L == 0) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (row + stride < rows) {
        mbar_expect_tx(&bars[warp][b ^ 1], N * sizeof(cx<T>));
        bulk_load(B0 + (b ^ 1) * N, src_of(row + stride), N * sizeof(cx<T>), &bars[warp][b ^ 1]);
      }
    }
    exchange_private(v, B, L);
    tail<true>(v, sTw, L);
#pragma unroll
    for (int m = 0; m < 16; ++m) B[L + 32 * m] = v[m];
    fence_async_smem();
    __syncwarp();
    if (L == 0) {
      bulk_store(dst_of(row), B, N * sizeof(cx<T>));
      tma_store_commit();
    }
  }
  if (L == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------
// yb: contiguous columns in ([p][x][y]), tile out ([p][y][x]) through TMA; W columns per tile
template <int W, int MINB>
__global__ void __launch_bounds__(W * 32, MINB)
    k_yb(const cx<T>* in, const __grid_constant__ TensorMap map, const __grid_constant__ Tw4 twp, int vplanes, int inMod, int outMod) {
  extern __shared__ __align__(1024) unsigned char smem[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smem);
  const int w = threadIdx.x >> 5, L = threadIdx.x & 31;
  STAGE_TW
  const int tilesPerPlane = N / W;
  const long long tiles = (long long)vplanes * tilesPerPlane;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int p = (int)(tile / tilesPerPlane), xt = (int)(tile % tilesPerPlane);
    const cx<T>* src = in + ((size_t)(p % inMod) * N + (size_t)xt * W + w) * N + L;
    cx<T> v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = src[32 * m];
    head<true>(v, L);
    if (threadIdx.x == 0) tma_store_wait_read();
    __syncthreads();
    exchange_col<W>(v, S, w, L);
    tail<true>(v, sTw, L);
#pragma unroll
    for (int m = 0; m < 16; ++m) *wcolw<W>(S, L + 32 * m, w) = v[m];
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_store_3d(&map, xt * W * 2, 0, p % outMod, S);
      tma_store_3d(&map, xt * W * 2, 256, p % outMod, S + 256 * W);
      tma_store_commit();
    }
  }
  if (threadIdx.x == 0) tma_store_wait_read();
}

// yf: tile in through TMA, contiguous columns out
template <int W, int MINB>
__global__ void __launch_bounds__(W * 32, MINB)
    k_yf(const __grid_constant__ TensorMap map, cx<T>* out, const __grid_constant__ Tw4 twp, int vplanes, int inMod, int outMod) {
  extern __shared__ __align__(1024) unsigned char smem[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smem);
  __shared__ __align__(8) uint64_t full;
  const int w = threadIdx.x >> 5, L = threadIdx.x & 31;
  STAGE_TW
  if (threadIdx.x == 0) {
    mbar_init(&full, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int tilesPerPlane = N / W;
  const long long tiles = (long long)vplanes * tilesPerPlane;
  constexpr uint32_t kBytes = N * W * sizeof(cx<T>);
  uint32_t phase = 0;
  long long tile = blockIdx.x;
  if (tile < tiles && threadIdx.x == 0) {
    const int p = (int)(tile / tilesPerPlane), xt = (int)(tile % tilesPerPlane);
    mbar_expect_tx(&full, kBytes);
    tma_load_3d(S, &map, xt * W * 2, 0, p % inMod, &full);
    tma_load_3d(S + 256 * W, &map, xt * W * 2, 256, p % inMod, &full);
  }
  for (; tile < tiles; tile += gridDim.x) {
    const int p = (int)(tile / tilesPerPlane), xt = (int)(tile % tilesPerPlane);
    mbar_wait(&full, phase);
    phase ^= 1;
    cx<T> v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = *wcolw<W>(S, L + 32 * m, w);
    __syncwarp();
    head<false>(v, L);
    exchange_col<W>(v, S, w, L);
    __syncthreads();
    const long long next = tile + gridDim.x;
    if (next < tiles && threadIdx.x == 0) {
      const int p2 = (int)(next / tilesPerPlane), xt2 = (int)(next % tilesPerPlane);
      mbar_expect_tx(&full, kBytes);
      tma_load_3d(S, &map, xt2 * W * 2, 0, p2 % inMod, &full);
      tma_load_3d(S + 256 * W, &map, xt2 * W * 2, 256, p2 % inMod, &full);
    }
    tail<false>(v, sTw, L);
    cx<T>* dst = out + ((size_t)(p % outMod) * N + (size_t)xt * W + w) * N + L;
#pragma unroll
    for (int m = 0; m < 16; ++m) dst[32 * m] = v[m];
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename F>
float time_ms(F&& f, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 2; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms / reps;
}

static int g_sms = 148;
static int g_vplanes = 256;
static int g_reps = 5;

static void report(const char* name, int inMod, int outMod, float ms, int perSm) {
  const double items = (double)g_vplanes * 64;  // units of 8 transforms
  const double cyc = ms * 1e-3 * 1.965e9 * g_sms / items;
  printf("%-22s in %s out %s  CTAs/SM %d  %.4f ms  %7.0f SM-cycles per 8 transforms  (%.3f ms per 512 planes)\n", name,
         inMod < g_vplanes ? "L2 " : "HBM", outMod < g_vplanes ? "L2 " : "HBM", perSm, ms, cyc, ms * 512.0 / g_vplanes);
  fflush(stdout);
}

template <typename K>
int occupancy(K kern, int threads, size_t smem) {
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int b = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, threads, smem));
  return b;
}

int main(int argc, char** argv) {
  g_vplanes = argc > 1 ? atoi(argv[1]) : 256;
  g_reps = argc > 2 ? atoi(argv[2]) : 5;
  CK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0));
  const int vp = g_vplanes;
  const size_t elems = (size_t)vp * N * N;
  cx<T>*a, *b;
  CK(cudaMalloc(&a, elems * sizeof(cx<T>)));
  CK(cudaMalloc(&b, elems * sizeof(cx<T>)));
  {
    std::vector<cx<T>> h((size_t)N * N);
    srand(1);
    for (auto& v : h) v = mk<T>(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
    for (int p = 0; p < vp; ++p) CK(cudaMemcpy(a + (size_t)p * N * N, h.data(), h.size() * sizeof(cx<T>), cudaMemcpyHostToDevice));
    CK(cudaMemset(b, 0, elems * sizeof(cx<T>)));
  }
  Tw4 tw;
  wfft_lane_twiddles<T>(N, 32, &tw.w[0][0]);
  const int mods[4][2] = {{8, 8}, {vp, 8}, {8, vp}, {vp, vp}};

  auto run_x = [&](auto kern, const char* name, int warps) {
    const size_t smem = (size_t)warps * N * sizeof(cx<T>);
    const int per = occupancy(kern, warps * 32, smem);
    for (auto& m : mods) {
      const int im = m[0], om = m[1];
      auto f = [&] { kern<<<g_sms * per, warps * 32, smem>>>(a, b, tw, vp, im, om); };
      report(name, im, om, time_ms(f, g_reps), per);
    }
  };
  run_x(k_x<8, 2>, "x ldg 8w x2", 8);
  run_x(k_x<4, 4>, "x ldg 4w x4", 4);
  run_x(k_x<2, 8>, "x ldg 2w x8", 2);
  run_x(k_x<12, 1>, "x ldg 12w x1 (168r)", 12);
  run_x(k_x<6, 2>, "x ldg 6w x2 (168r)", 6);
  run_x(k_x<4, 3>, "x ldg 4w x3 (168r)", 4);
  {
    auto run_xt = [&](auto kern, const char* name, int warps) {
      const size_t smem = (size_t)warps * 2 * N * sizeof(cx<T>);
      const int per = occupancy(kern, warps * 32, smem);
      for (auto& m : mods) {
        const int im = m[0], om = m[1];
        auto f = [&] { kern<<<g_sms * per, warps * 32, smem>>>(a, b, tw, vp, im, om); };
        report(name, im, om, time_ms(f, g_reps), per);
      }
    };
    run_xt(k_xt<12, 1>, "x bulk 12w", 12);
    run_xt(k_xt<8, 1>, "x bulk 8w", 8);
    run_xt(k_xt<6, 2>, "x bulk 6w x2", 6);
    run_xt(k_xt<4, 3>, "x bulk 4w x3", 4);
  }
  auto run_y = [&](auto kb, auto kf, const char* name, int W) {
    const size_t smem = (size_t)N * W * sizeof(cx<T>);
    TensorMap mapA, mapB;
    if (make_tile_map(&mapA, a, sizeof(cx<T>), N, N, N, vp, (long long)N * N, W, 256) ||
        make_tile_map(&mapB, b, sizeof(cx<T>), N, N, N, vp, (long long)N * N, W, 256)) {
      printf("%s: tensor map failed\n", name);
      return;
    }
    const int perB = occupancy(kb, W * 32, smem), perF = occupancy(kf, W * 32, smem);
    char nm[64];
    for (auto& m : mods) {
      const int im = m[0], om = m[1];
      auto fb = [&] { kb<<<g_sms * perB, W * 32, smem>>>(a, mapB, tw, vp, im, om); };
      snprintf(nm, sizeof nm, "yb %s", name);
      report(nm, im, om, time_ms(fb, g_reps), perB);
      auto ff = [&] { kf<<<g_sms * perF, W * 32, smem>>>(mapA, b, tw, vp, im, om); };
      snprintf(nm, sizeof nm, "yf %s", name);
      report(nm, im, om, time_ms(ff, g_reps), perF);
    }
  };
  run_y(k_yb<8, 2>, k_yf<8, 2>, "W8 x2", 8);
  run_y(k_yb<4, 4>, k_yf<4, 4>, "W4 x4", 4);
  run_y(k_yb<2, 8>, k_yf<2, 8>, "W2 x8", 2);
  return 0;
}
