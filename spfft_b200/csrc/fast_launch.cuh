// fast_launch.cuh -- kernel entry points + length dispatch of the register-FFT stage kernels.
// Included by fast_z.cu / fast_y.cu / fast_x.cu (one translation unit per stage so that the
// instantiations compile in parallel); kernels.cu calls the launch_*_fast functions.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "fast_stage_kernels.hpp"

namespace sb {

template <typename T, int N>
struct FastCfg {
  static constexpr int V = 1 << FastLanes<T>::log2V;
  static constexpr int threads = V * (N / 8);
  static constexpr int minBlocks = threads >= 1024 ? 1 : (1024 / threads > 16 ? 16 : 1024 / threads);
  static constexpr size_t smem = (size_t)N * V * sizeof(cx<T>);
};

// the stand-alone x stage kernels have their own lane count (FastLanesX)
template <typename T, int N>
struct FastCfgX {
  static constexpr int V = 1 << FastLanesX<T, N>::log2V;
  static constexpr int threads = V * (N / 8);
  static constexpr int minBlocks = threads >= 1024 ? 1 : (1024 / threads > 16 ? 16 : 1024 / threads);
  static constexpr size_t smem = (size_t)N * V * sizeof(cx<T>);
};

// Programmatic dependent launch of the stage kernels (pdl_prologue, fast_fft.hpp); SPFFT_B200_PDL=0 disables.
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("SPFFT_B200_PDL");
    return !(e && atoi(e) == 0);
  }();
  return on;
}
template <typename Kernel, typename... A>
int launch_stage_kernel(Kernel kernel, dim3 grid, int threads, size_t smemBytes, cudaStream_t stream, const A&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, kernel, args...);
}

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
  return launch_stage_kernel(kernel, dim3((unsigned)blocks), threads, smemBytes, stream, args);
}

// one launch over `bands` transforms of the same plan: grid (blocks, bands)
template <typename Kernel, typename Args, typename Table>
int launch_bands(Kernel kernel, const Args& args, const Table& table, long long blocks, int bands,
                 int threads, size_t smemBytes, cudaStream_t stream) {
  if (blocks <= 0 || bands <= 0) return 0;
  if (blocks > 0x7fffffffLL || bands > 65535) return (int)cudaErrorInvalidConfiguration;
  if (smemBytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smemBytes);
    if (e != cudaSuccess) return (int)e;
  }
  return launch_stage_kernel(kernel, dim3((unsigned)blocks, (unsigned)bands), threads, smemBytes, stream, args, table);
}

// Debug / tuning knob: environment variable SPFFT_B200_TUNE (integer bit mask, default 1).
//   bit 0: L2 prefetch of the next tile's inputs
inline int tune_flags() {
  static const int flags = [] {
    const char* e = getenv("SPFFT_B200_TUNE");
    return e ? atoi(e) : 1;
  }();
  return flags;
}

// CTAs resident on the device for a kernel that fits `perSm` per SM = how far ahead (in tiles) the
// tile that will run next on "this" slot is.
inline int resident_ctas(int perSm) {
  static const int sms = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
    return v;
  }();
  return sms * perSm;
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
