// pipe_overlap.cu -- do the fp64 pipe and the shared-memory (LSU / MIO) pipe of an SM overlap? (experiment)
//   A: 16 warps per SM, all fp64 (16 independent DFMA chains per lane)
//   B: 16 warps per SM, all LDS.128 / STS.128 (conflict free)
//   C: 8 warps of A-work + 8 warps of B-work in the same CTAs
// If the pipes are independent, C takes max(A, B) / 2; if the instructions share one dispatch path, (A + B) / 2.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void fp64_work(double* acc, int iters, double a, double b) {
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = fma(acc[k], a, b);
  }
}
__device__ __forceinline__ void lsu_work(double2* s, int iters, int lane, double2& keep) {
  const unsigned base = (unsigned)__cvta_generic_to_shared(s + lane);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(base + 512u * k), "d"(keep.x), "d"(keep.y) : "memory");
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(keep.x), "=d"(keep.y) : "r"(base + 512u * ((k + 3) & 7)) : "memory");
    }
  }
}
// mode 0: fp64 only, 1: lsu only, 2: warps split half / half
__global__ void __launch_bounds__(256, 2) k(double* out, int iters, int mode, double a, double b) {
  extern __shared__ double2 sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) acc[k2] = threadIdx.x + k2;
  double2 keep = make_double2(lane, warp);
  const bool doFp = mode == 0 || (mode == 2 && (warp & 1) == 0);
  const bool doLs = mode == 1 || (mode == 2 && (warp & 1) == 1);
  if (doFp) fp64_work(acc, iters, a, b);
  if (doLs) lsu_work(sm + warp * 256, iters / 4, lane, keep);  // a quarter of the iterations: similar duration as A
  double r = keep.x + keep.y;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) r += acc[k2];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * sms * 2 * 256));
  const size_t smem = 8 * 256 * sizeof(double2);
  const int iters = 4000;
  for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<<<sms * 2, 256, smem>>>(out, iters, mode, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k<<<sms * 2, 256, smem>>>(out, iters, mode, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double cyc = ms * 1e-3 * 1.965e9;
    const char* nm[3] = {"A fp64 only (16 warps/SM)", "B LDS/STS only (16 warps/SM)", "C 8 warps fp64 + 8 warps LDS/STS"};
    // per SM: fp64 warp-instr = warps * iters * 16; lsu warp-instr = warps * iters * 16 (8 STS + 8 LDS, 4 wavefronts each)
    printf("%-36s %.3f ms  %.0f cycles", nm[mode], ms, cyc);
    if (mode == 0) printf("  -> %.2f cycles per fp64 warp instruction per SM", cyc / (16.0 * iters * 16));
    if (mode == 1) printf("  -> %.2f cycles per 128-bit shared access per SM (4 wavefronts)", cyc / (16.0 * (iters / 4) * 16));
    printf("\n");
  }
  return 0;
}
