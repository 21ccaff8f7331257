// wfft_bench.cu -- micro-benchmark of the warp-autonomous FFT (spfft_b200/csrc/wfft.hpp) in the three
// access patterns the stage kernels need, on synthetic 512 x 512 planes (experiments only):
//   x : rows in, rows out (contiguous both sides)                      -- x stage
//   yf: TMA tile [512 y][8 x] in (128B-swizzled), contiguous columns out -- y / z forward dense side
//   yb: contiguous columns in, tile [512 y][8 x] out through TMA        -- y / z backward dense side
// Usage: wfft_bench [planes] [reps]
#include <cuda_runtime.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tma_util.hpp"
#include "wfft.hpp"

using namespace sb;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int WARPS = 8;

template <typename T, int N>
__device__ __forceinline__ void load_tw(cx<T>* tws, const cx<T>* tw) {
  for (int i = threadIdx.x; i < WPlan<T, N>::TW; i += blockDim.x) tws[i] = tw[i];
}

// ---- x: one row per warp (N = 512) -------------------------------------------------------------
template <typename T, int N, bool BWD, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_x(const cx<T>* in, cx<T>* out, const cx<T>* tw, int rows) {
  using P = WPlan<T, N>;
  extern __shared__ __align__(1024) unsigned char smem[];
  cx<T>* X = reinterpret_cast<cx<T>*>(smem);
  cx<T>* tws = X + WARPS * N * P::PER_WARP;
  load_tw<T, N>(tws, tw);
  __syncthreads();
  const int warp = threadIdx.x >> 5, L = threadIdx.x & 31;
  cx<T>* Xw = X + warp * N * P::PER_WARP;
  const int rowsPerWarp = P::PER_WARP;
  for (long long item = (long long)blockIdx.x * WARPS + warp; item * rowsPerWarp < rows; item += (long long)gridDim.x * WARPS) {
    const long long row = item * rowsPerWarp + (L / P::LANES);
    const int l = L % P::LANES;
    const cx<T>* src = in + row * N + l;
    cx<T> v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = src[P::LANES * m];
    warp_fft<T, N, BWD>(v, Xw, tws, L);
    cx<T>* dst = out + row * N + l;
#pragma unroll
    for (int m = 0; m < 16; ++m) dst[P::LANES * m] = v[m];
  }
}

// ---- yf: TMA tile in, warp per tile column, contiguous out --------------------------------------
// in: planes [z][y][x] through the tensor map; out: [z][x][y] (column x of plane z contiguous)
template <typename T, int N, bool BWD, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_yf(const __grid_constant__ TensorMap map, cx<T>* out, const cx<T>* tw,
                                                        int planes) {
  using P = WPlan<T, N>;
  static_assert(N == 512, "one warp per column");
  extern __shared__ __align__(1024) unsigned char smem[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smem);  // [N][8] swizzled tile = exchange buffer
  cx<T>* tws = S + N * 8;
  __shared__ __align__(8) uint64_t full;
  load_tw<T, N>(tws, tw);
  if (threadIdx.x == 0) {
    mbar_init(&full, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, L = threadIdx.x & 31;
  const int tilesPerPlane = N / 8;
  const long long tiles = (long long)planes * tilesPerPlane;
  uint32_t phase = 0;
  long long tile = blockIdx.x;
  if (tile < tiles && threadIdx.x == 0) {
    mbar_expect_tx(&full, N * 8 * sizeof(cx<T>));
    const int z = (int)(tile / tilesPerPlane), xt = (int)(tile % tilesPerPlane);
    tma_load_3d(S, &map, xt * 16, 0, z, &full);
    tma_load_3d(S + 256 * 8, &map, xt * 16, 256, z, &full);
  }
  for (; tile < tiles; tile += gridDim.x) {
    const int z = (int)(tile / tilesPerPlane), xt = (int)(tile % tilesPerPlane);
    mbar_wait(&full, phase);
    phase ^= 1;
    cx<T> v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int n = L + 32 * m;
      v[m] = S[n * 8 + (w ^ (n & 7))];
    }
    __syncwarp();
    // transform with the exchange in the warp's own slots of the tile
    P::template stage_a_local<BWD>(v, L);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const cx<T> recv = shfl_xor_cx<T>(v[8 + i], 16);
      P::stage_a_combine(v[i], v[8 + i], recv, L);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int s = P::xw(L, i);
      S[s * 8 + (w ^ (s & 7))] = v[i];
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int s = P::xr(L, r);
      v[r] = S[s * 8 + (w ^ (s & 7))];
    }
    __syncthreads();  // every warp is done with the tile buffer
    const long long next = tile + gridDim.x;
    if (next < tiles && threadIdx.x == 0) {
      mbar_expect_tx(&full, N * 8 * sizeof(cx<T>));
      const int z2 = (int)(next / tilesPerPlane), xt2 = (int)(next % tilesPerPlane);
      tma_load_3d(S, &map, xt2 * 16, 0, z2, &full);
      tma_load_3d(S + 256 * 8, &map, xt2 * 16, 256, z2, &full);
    }
    P::template stage_b<BWD>(v, L, tws);
    cx<T>* dst = out + ((size_t)z * N + (size_t)xt * 8 + w) * N + L;
#pragma unroll
    for (int m = 0; m < 16; ++m) dst[32 * m] = v[m];
  }
}

// ---- yb: contiguous columns in, tile out through TMA ---------------------------------------------
template <typename T, int N, bool BWD, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_yb(const cx<T>* in, const __grid_constant__ TensorMap map, const cx<T>* tw,
                                                        int planes) {
  using P = WPlan<T, N>;
  static_assert(N == 512, "one warp per column");
  extern __shared__ __align__(1024) unsigned char smem[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smem);
  cx<T>* tws = S + N * 8;
  load_tw<T, N>(tws, tw);
  __syncthreads();
  const int w = threadIdx.x >> 5, L = threadIdx.x & 31;
  const int tilesPerPlane = N / 8;
  const long long tiles = (long long)planes * tilesPerPlane;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int z = (int)(tile / tilesPerPlane), xt = (int)(tile % tilesPerPlane);
    const cx<T>* src = in + ((size_t)z * N + (size_t)xt * 8 + w) * N + L;
    cx<T> v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = src[32 * m];
    P::template stage_a_local<BWD>(v, L);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const cx<T> recv = shfl_xor_cx<T>(v[8 + i], 16);
      P::stage_a_combine(v[i], v[8 + i], recv, L);
    }
    // the previous tile's TMA store must have finished reading the buffer
    if (threadIdx.x == 0) tma_store_wait_read();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int s = P::xw(L, i);
      S[s * 8 + (w ^ (s & 7))] = v[i];
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int s = P::xr(L, r);
      v[r] = S[s * 8 + (w ^ (s & 7))];
    }
    __syncwarp();
    P::template stage_b<BWD>(v, L, tws);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int n = L + 32 * m;
      S[n * 8 + (w ^ (n & 7))] = v[m];
    }
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_store_3d(&map, xt * 16, 0, z, S);
      tma_store_3d(&map, xt * 16, 256, z, S + 256 * 8);
      tma_store_commit();
    }
  }
  if (threadIdx.x == 0) tma_store_wait_read();
}

// ---- host ----------------------------------------------------------------------------------------
template <typename T>
std::vector<cx<T>> host_tw(int n) {
  std::vector<cx<T>> tw(15 * (n == 512 ? 32 : 16));
  const int lanes = n == 512 ? 32 : 16;
  const long double pi2 = 6.283185307179586476925286766559005768L;
  for (int r = 1; r < 16; ++r)
    for (int L = 0; L < lanes; ++L) {
      const long double a = -pi2 * (long double)((r * L) % n) / (long double)n;
      tw[(r - 1) * lanes + L] = mk<T>((T)cosl(a), (T)sinl(a));
    }
  return tw;
}

template <typename F>
float time_ms(F&& f, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

using cd = std::complex<double>;
static double check_rows(const std::vector<cd>& x, const std::vector<cd>& y, int n, bool bwd, int rowsToCheck, long long strideIn,
                         long long strideOut, long long elemIn, long long elemOut) {
  // x row r: x[r*strideIn + k*elemIn]; y likewise
  double worst = 0;
  for (int r = 0; r < rowsToCheck; ++r) {
    double num = 0, den = 0;
    for (int k = 0; k < n; ++k) {
      std::complex<long double> acc = 0;
      for (int j = 0; j < n; ++j) {
        const long double a = (bwd ? 1.0L : -1.0L) * 6.283185307179586476925286766559005768L * (long double)((long long)j * k % n) / n;
        acc += std::complex<long double>(x[r * strideIn + j * elemIn]) * std::complex<long double>(cosl(a), sinl(a));
      }
      const cd got = y[r * strideOut + k * elemOut];
      num += std::norm(got - cd((double)acc.real(), (double)acc.imag()));
      den += std::norm(cd((double)acc.real(), (double)acc.imag()));
    }
    worst = std::max(worst, std::sqrt(num / den));
  }
  return worst;
}

int main(int argc, char** argv) {
  const int planes = argc > 1 ? atoi(argv[1]) : 512;
  const int reps = argc > 2 ? atoi(argv[2]) : 10;
  constexpr int N = 512;
  using T = double;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t elems = (size_t)planes * N * N;
  cx<T>*a, *b;
  CK(cudaMalloc(&a, elems * sizeof(cx<T>)));
  CK(cudaMalloc(&b, elems * sizeof(cx<T>)));
  std::vector<cd> h(elems > (size_t)4 * N * N ? (size_t)4 * N * N : elems);
  srand(1);
  for (auto& v : h) v = cd(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
  for (size_t off = 0; off < elems; off += h.size())
    CK(cudaMemcpy(a + off, h.data(), std::min(h.size(), elems - off) * sizeof(cd), cudaMemcpyHostToDevice));
  auto tw = host_tw<T>(N);
  cx<T>* dtw;
  CK(cudaMalloc(&dtw, tw.size() * sizeof(cx<T>)));
  CK(cudaMemcpy(dtw, tw.data(), tw.size() * sizeof(cx<T>), cudaMemcpyHostToDevice));
  const double gb = 2.0 * elems * sizeof(cx<T>) / 1e9;
  std::vector<cd> out(h.size());

  // ---------------- x ----------------
  {
    const size_t smem = (size_t)WARPS * N * sizeof(cx<T>) + tw.size() * sizeof(cx<T>);
    auto kern = k_x<T, N, true, 2>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int rows = planes * N;
    for (int mult : {2, 4, 8}) {
      auto run = [&] { kern<<<sms * mult / 1 / 2 * 2 / 2, WARPS * 32, smem>>>(a, b, dtw, rows); };
      (void)run;
    }
    auto run = [&] { kern<<<sms * 2, WARPS * 32, smem>>>(a, b, dtw, rows); };
    run();
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), b, out.size() * sizeof(cd), cudaMemcpyDeviceToHost));
    printf("x  bwd check (4 rows): rel-L2 %.3e\n", check_rows(h, out, N, true, 4, N, N, 1, 1));
    const float ms = time_ms(run, reps);
    printf("x  bwd: %d planes  %.4f ms  %.1f GB/s  (%.4f ms per 512 planes)\n", planes, ms, gb / (ms * 1e-3), ms * 512.0 / planes);
    auto kernf = k_x<T, N, false, 2>;
    CK(cudaFuncSetAttribute(kernf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto runf = [&] { kernf<<<sms * 2, WARPS * 32, smem>>>(a, b, dtw, rows); };
    runf();
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), b, out.size() * sizeof(cd), cudaMemcpyDeviceToHost));
    printf("x  fwd check (4 rows): rel-L2 %.3e\n", check_rows(h, out, N, false, 4, N, N, 1, 1));
  }
  // ---------------- yf ----------------
  {
    TensorMap map;
    if (make_tile_map(&map, a, sizeof(cx<T>), N, N, N, planes, (long long)N * N, 8, 256) != 0) {
      printf("tensor map failed\n");
      return 1;
    }
    const size_t smem = (size_t)N * 8 * sizeof(cx<T>) + tw.size() * sizeof(cx<T>);
    auto kern = k_yf<T, N, false, 2>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto run = [&] { kern<<<sms * 2, WARPS * 32, smem>>>(map, b, dtw, planes); };
    run();
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), b, out.size() * sizeof(cd), cudaMemcpyDeviceToHost));
    // input column x of plane 0: h[y*N + x]; output: out[x*N + k]
    printf("yf fwd check (4 cols): rel-L2 %.3e\n", check_rows(h, out, N, false, 4, 1, N, N, 1));
    const float ms = time_ms(run, reps);
    printf("yf fwd: %d planes  %.4f ms  %.1f GB/s  (%.4f ms per 512 planes)\n", planes, ms, gb / (ms * 1e-3), ms * 512.0 / planes);
  }
  // ---------------- yb ----------------
  {
    TensorMap map;
    if (make_tile_map(&map, b, sizeof(cx<T>), N, N, N, planes, (long long)N * N, 8, 256) != 0) {
      printf("tensor map failed\n");
      return 1;
    }
    const size_t smem = (size_t)N * 8 * sizeof(cx<T>) + tw.size() * sizeof(cx<T>);
    auto kern = k_yb<T, N, true, 2>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto run = [&] { kern<<<sms * 2, WARPS * 32, smem>>>(a, map, dtw, planes); };
    CK(cudaMemset(b, 0, elems * sizeof(cx<T>)));
    run();
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), b, out.size() * sizeof(cd), cudaMemcpyDeviceToHost));
    // input column c of plane 0 is contiguous: h[c*N + j]; output element k of column c: out[k*N + c]
    printf("yb bwd check (4 cols): rel-L2 %.3e\n", check_rows(h, out, N, true, 4, N, 1, 1, N));
    const float ms = time_ms(run, reps);
    printf("yb bwd: %d planes  %.4f ms  %.1f GB/s  (%.4f ms per 512 planes)\n", planes, ms, gb / (ms * 1e-3), ms * 512.0 / planes);
  }
  return 0;
}
