// fast_launch.cuh -- kernel entry points + length dispatch of the register-FFT stage kernels.
// Included by fast_z.cu / fast_y.cu / fast_x.cu (one translation unit per stage so that the
// instantiations compile in parallel); kernels.cu calls the launch_*_fast functions.
#pragma once
#include <cuda_runtime.h>

#include "fast_stage_kernels.hpp"

namespace sb {

template <typename T, int N>
struct FastCfg {
  static constexpr int V = 1 << FastLanes<T>::log2V;
  static constexpr int threads = V * (N / 8);
  static constexpr int minBlocks = threads >= 1024 ? 1 : (1024 / threads > 16 ? 16 : 1024 / threads);
  static constexpr size_t smem = (size_t)N * V * sizeof(cx<T>);
};


template <typename Kernel, typename Args>
int launch_fast(Kernel kernel, const Args& args, long long blocks, int threads, size_t smemBytes,
                cudaStream_t stream) {
  if (blocks <= 0) return 0;
  if (blocks > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  if (smemBytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smemBytes);
    if (e != cudaSuccess) return (int)e;
  }
  kernel<<<(unsigned)blocks, threads, smemBytes, stream>>>(args);
  return (int)cudaGetLastError();
}

// lengths with instantiated kernels
#define SB_FAST_DISPATCH(n, CALL)            \
  switch (n) {                               \
    case 32: CALL(32); break;                \
    case 64: CALL(64); break;                \
    case 128: CALL(128); break;              \
    case 256: CALL(256); break;              \
    case 512: CALL(512); break;              \
    case 1024: CALL(1024); break;            \
    default: return (int)cudaErrorInvalidValue; \
  }

// returns a cudaError_t value; cudaErrorInvalidValue when the length has no fast kernel
template <typename T>
int launch_z_fast(int forward, const ZArgs<T>& a, cudaStream_t s);
template <typename T>
int launch_y_fast(int forward, const YArgs<T>& a, cudaStream_t s);
template <typename T>
int launch_x_fast(int forward, const XArgs<T>& a, cudaStream_t s);

}  // namespace sb
