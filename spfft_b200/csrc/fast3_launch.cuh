// fast3_launch.cuh -- launch configuration + length dispatch of the N = 3 * 2^k register-FFT
// kernels (fast3_stage_kernels.hpp). Included by fast3_z.cu / fast3_y.cu / fast3_x.cu.
#pragma once
#include "fast3_stage_kernels.hpp"
#include "fast_launch.cuh"

namespace sb {

// One CTA = one tile of V lanes, V * N/24 threads (64 for N = 192) with 24 complex values each.
// Register budget: 168 per thread (384 threads per SM) for double, float 80 (N <= 192) / 128. The small CTAs keep the barriers cheap (2 warps at N = 192) and many tiles at
// different phases resident per SM.
// threads per SM the register cap of the kernels for N <= 192 is derived from. Measured at 192^3
// (profiles/r01_v4_fast3.md): double 384 (168 registers, no spills) is 35 % faster than 512 (128
// registers, 16-80 bytes of spills).
#ifndef SB_F3_PSM_F64
#define SB_F3_PSM_F64 384
#endif
#ifndef SB_F3_PSM_F32
#define SB_F3_PSM_F32 768
#endif
template <typename T, int N>
struct Fast3Cfg {
  static constexpr int V = 1 << Fast3Lanes<T, N>::log2V;
  static constexpr int threads = V * Fast3Plan<N>::T;
  // (5 * 2^k: 40 values per thread -> no register cap for double, 168 for float)
  static constexpr int perSmThreads = N % 5 == 0 ? (sizeof(T) == 8 ? 256 : 384)
                                                 : (sizeof(T) == 8 ? (N <= 192 ? SB_F3_PSM_F64 : 384) : (N <= 192 ? SB_F3_PSM_F32 : 512));
  static constexpr int minBlocks = perSmThreads / threads > 16 ? 16 : (perSmThreads / threads < 1 ? 1 : perSmThreads / threads);
  static constexpr size_t smem = (size_t)N * V * sizeof(cx<T>);
};

template <typename T, int N>
struct Fast3CfgX {
  static constexpr int V = 1 << FastLanesX<T, N>::log2V;
  static constexpr int threads = V * Fast3Plan<N>::T;
  static constexpr int perSm = Fast3Cfg<T, N>::perSmThreads / threads;
  static constexpr int minBlocks = perSm > 16 ? 16 : (perSm < 1 ? 1 : perSm);
  static constexpr size_t smem = (size_t)N * V * sizeof(cx<T>);
};

#define SB_FAST3_DISPATCH(n, CALL)              \
  switch (n) {                                  \
    case 96: CALL(96); break;                   \
    case 192: CALL(192); break;                 \
    case 384: CALL(384); break;                 \
    case 768: CALL(768); break;                 \
    case 160: CALL(160); break;                 \
    case 320: CALL(320); break;                 \
    case 640: CALL(640); break;                 \
    default: return (int)cudaErrorInvalidValue; \
  }

template <typename T>
int launch_z_fast3(int forward, const ZArgs<T>& a, cudaStream_t s);
template <typename T>
int launch_y_fast3(int forward, const YArgs<T>& a, cudaStream_t s);
template <typename T>
int launch_x_fast3(int forward, const XArgs<T>& a, cudaStream_t s);

}  // namespace sb
