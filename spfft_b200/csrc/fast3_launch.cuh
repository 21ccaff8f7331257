// fast3_launch.cuh -- launch configuration + length dispatch of the N = 3 * 2^k register-FFT
// kernels (fast3_stage_kernels.hpp). Included by fast3_z.cu / fast3_y.cu / fast3_x.cu.
#pragma once
#include "fast3_stage_kernels.hpp"
#include "fast_launch.cuh"

namespace sb {

// One CTA = one tile of V lanes, V * N/24 threads (64 for N = 192) with 24 complex values each.
// Register budget: 128 per thread (512 threads per SM) for double up to N = 192, 168 (384 threads)
// for the three-stage lengths whose live ranges span more phases (ptxas -v: spills otherwise);
// float 80 / 128. The small CTAs keep the barriers cheap (2 warps at N = 192) and many tiles at
// different phases resident per SM.
template <typename T, int N>
struct Fast3Cfg {
  static constexpr int V = 1 << FastLanes<T>::log2V;
  static constexpr int threads = V * Fast3Plan<N>::T;
  static constexpr int perSmThreads = sizeof(T) == 8 ? (N <= 192 ? 512 : 384) : (N <= 192 ? 768 : 512);
  static constexpr int minBlocks = perSmThreads / threads > 16 ? 16 : (perSmThreads / threads < 1 ? 1 : perSmThreads / threads);
  static constexpr size_t smem = (size_t)N * V * sizeof(cx<T>);
};

#define SB_FAST3_DISPATCH(n, CALL)              \
  switch (n) {                                  \
    case 96: CALL(96); break;                   \
    case 192: CALL(192); break;                 \
    case 384: CALL(384); break;                 \
    case 768: CALL(768); break;                 \
    default: return (int)cudaErrorInvalidValue; \
  }

template <typename T>
int launch_z_fast3(int forward, const ZArgs<T>& a, cudaStream_t s);
template <typename T>
int launch_y_fast3(int forward, const YArgs<T>& a, cudaStream_t s);
template <typename T>
int launch_x_fast3(int forward, const XArgs<T>& a, cudaStream_t s);

}  // namespace sb
