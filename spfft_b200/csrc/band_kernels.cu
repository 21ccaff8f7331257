// band_kernels.cu -- batched multi-transform kernels (sm_100a): the stage kernels of kernels.cu /
// fast_*.cu / fast3_*.cu with blockIdx.y = band, for transforms that share one plan (clones).
// One launch per stage covers up to kMaxBands transforms; per band only the data pointers differ
// (BandTable, stage_kernels.hpp). Instantiated for the lengths where a single transform cannot
// fill the GPU (power-of-two axes up to 256, 3*2^k up to 192, any length on the generic kernels);
// larger transforms overlap well enough on their own streams.
// Replaces the phase-interleaved loop of the reference (multi_transform_internal.hpp:50-176).
#include <cuda_runtime.h>

#include "fast3_launch.cuh"
#include "fast_launch.cuh"
#include "launch.h"

namespace sb {

constexpr int kBandThreads = 256;  // generic kernels (same as kernels.cu)

#define SB_SMEM(T) \
  extern __shared__ __align__(16) unsigned char smemRaw[]; \
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);

// ---- generic kernels ---------------------------------------------------------------------------
template <typename T, bool FWD>
__global__ void __launch_bounds__(kBandThreads) k_z_stage_b(const __grid_constant__ ZArgs<T> a0, const __grid_constant__ BandTable<T> bt) {
  pdl_prologue();
  SB_SMEM(T)
  const ZArgs<T> a = band_args(a0, bt, (int)blockIdx.y);
  if (FWD) z_forward_body<T>(a, (int)blockIdx.x, Ctx{kBandThreads}, S);
  else z_backward_body<T>(a, (int)blockIdx.x, Ctx{kBandThreads}, S);
}
template <typename T, bool FWD>
__global__ void __launch_bounds__(kBandThreads) k_y_stage_b(const __grid_constant__ YArgs<T> a0, const __grid_constant__ BandTable<T> bt) {
  pdl_prologue();
  SB_SMEM(T)
  const YArgs<T> a = band_args(a0, bt, (int)blockIdx.y);
  if (FWD) y_forward_body<T>(a, (int)blockIdx.x, Ctx{kBandThreads}, S);
  else y_backward_body<T>(a, (int)blockIdx.x, Ctx{kBandThreads}, S);
}
template <typename T, bool FWD>
__global__ void __launch_bounds__(kBandThreads) k_x_stage_b(const __grid_constant__ XArgs<T> a0, const __grid_constant__ BandTable<T> bt) {
  pdl_prologue();
  SB_SMEM(T)
  const XArgs<T> a = band_args(a0, bt, (int)blockIdx.y);
  if (FWD) x_forward_body<T>(a, (int)blockIdx.x, Ctx{kBandThreads}, S);
  else x_backward_body<T>(a, (int)blockIdx.x, Ctx{kBandThreads}, S);
}

// ---- power-of-two register-FFT kernels ---------------------------------------------------------
template <typename T, int N, bool FWD>
__global__ void __launch_bounds__(FastCfg<T, N>::threads, FastCfg<T, N>::minBlocks)
    k_z_fast_b(const __grid_constant__ ZArgs<T> a0, const __grid_constant__ BandTable<T> bt) {
  pdl_prologue();
  SB_SMEM(T)
  const ZArgs<T> a = band_args(a0, bt, (int)blockIdx.y);
  z_fast_any<T, N, FWD>(a, (int)blockIdx.x, Ctx{FastCfg<T, N>::threads}, S);
}
template <typename T, int N, bool FWD>
__global__ void __launch_bounds__(FastCfg<T, N>::threads, FastCfg<T, N>::minBlocks)
    k_y_fast_b(const __grid_constant__ YArgs<T> a0, const __grid_constant__ BandTable<T> bt) {
  pdl_prologue();
  SB_SMEM(T)
  const YArgs<T> a = band_args(a0, bt, (int)blockIdx.y);
  if (FWD) y_forward_fast<T, N>(a, (int)blockIdx.x, Ctx{FastCfg<T, N>::threads}, S);
  else y_backward_fast<T, N>(a, (int)blockIdx.x, Ctx{FastCfg<T, N>::threads}, S);
}
template <typename T, int N, bool FWD, bool REAL>
__global__ void __launch_bounds__(FastCfgX<T, N>::threads, FastCfgX<T, N>::minBlocks)
    k_x_fast_b(const __grid_constant__ XArgs<T> a0, const __grid_constant__ BandTable<T> bt) {
  pdl_prologue();
  SB_SMEM(T)
  const XArgs<T> a = band_args(a0, bt, (int)blockIdx.y);
  if (REAL) x_r2c_fast<T, N, !FWD>(a, (int)blockIdx.x, Ctx{FastCfgX<T, N>::threads}, S);
  else x_c2c_fast<T, N, !FWD>(a, (int)blockIdx.x, Ctx{FastCfgX<T, N>::threads}, S);
}

// ---- 3 * 2^k register-FFT kernels --------------------------------------------------------------
template <typename T, int N, bool FWD>
__global__ void __launch_bounds__(Fast3Cfg<T, N>::threads, Fast3Cfg<T, N>::minBlocks)
    k_z_fast3_b(const __grid_constant__ ZArgs<T> a0, const __grid_constant__ BandTable<T> bt) {
  pdl_prologue();
  SB_SMEM(T)
  const ZArgs<T> a = band_args(a0, bt, (int)blockIdx.y);
  if (FWD) z_forward_fast3<T, N>(a, (int)blockIdx.x, Ctx{Fast3Cfg<T, N>::threads}, S);
  else z_backward_fast3<T, N>(a, (int)blockIdx.x, Ctx{Fast3Cfg<T, N>::threads}, S);
}
template <typename T, int N, bool FWD>
__global__ void __launch_bounds__(Fast3Cfg<T, N>::threads, Fast3Cfg<T, N>::minBlocks)
    k_y_fast3_b(const __grid_constant__ YArgs<T> a0, const __grid_constant__ BandTable<T> bt) {
  pdl_prologue();
  SB_SMEM(T)
  const YArgs<T> a = band_args(a0, bt, (int)blockIdx.y);
  if (FWD) y_forward_fast3<T, N>(a, (int)blockIdx.x, Ctx{Fast3Cfg<T, N>::threads}, S);
  else y_backward_fast3<T, N>(a, (int)blockIdx.x, Ctx{Fast3Cfg<T, N>::threads}, S);
}
template <typename T, int N, bool FWD, bool REAL>
__global__ void __launch_bounds__(Fast3CfgX<T, N>::threads, Fast3CfgX<T, N>::minBlocks)
    k_x_fast3_b(const __grid_constant__ XArgs<T> a0, const __grid_constant__ BandTable<T> bt) {
  pdl_prologue();
  SB_SMEM(T)
  const XArgs<T> a = band_args(a0, bt, (int)blockIdx.y);
  if (REAL) x_r2c_fast3<T, N, !FWD>(a, (int)blockIdx.x, Ctx{Fast3CfgX<T, N>::threads}, S);
  else x_c2c_fast3<T, N, !FWD>(a, (int)blockIdx.x, Ctx{Fast3CfgX<T, N>::threads}, S);
}

// lengths with batched register-FFT kernels
#define SB_BAND_DISPATCH(n, POW2, THREE)          \
  switch (n) {                                    \
    case 32: POW2(32); break;                     \
    case 64: POW2(64); break;                     \
    case 128: POW2(128); break;                   \
    case 256: POW2(256); break;                   \
    case 96: THREE(96); break;                    \
    case 192: THREE(192); break;                  \
    default: return (int)cudaErrorInvalidValue;   \
  }

static bool band_length(int n) { return n == 32 || n == 64 || n == 128 || n == 256 || n == 96 || n == 192; }

template <typename T>
static int launch_z_bands(int fwd, const ZArgs<T>& a0, const BandTable<T>& bt, int bands, cudaStream_t s) {
  ZArgs<T> a = a0;
  if (!a.ftw) {
    const size_t smem = 2 * ((size_t)a.nz << a.log2V) * sizeof(cx<T>);
    return fwd ? launch_bands(k_z_stage_b<T, true>, a, bt, a.numTiles, bands, kBandThreads, smem, s)
               : launch_bands(k_z_stage_b<T, false>, a, bt, a.numTiles, bands, kBandThreads, smem, s);
  }
#define POW2(NN)                                                                                              \
  {                                                                                                           \
    using C = FastCfg<T, NN>;                                                                                 \
    a.pfDist = 0;                                                                                             \
    return fwd ? launch_bands(k_z_fast_b<T, NN, true>, a, bt, a.numTiles, bands, C::threads, C::smem, s)      \
               : launch_bands(k_z_fast_b<T, NN, false>, a, bt, a.numTiles, bands, C::threads, C::smem, s);    \
  }
#define THREE(NN)                                                                                             \
  {                                                                                                           \
    using C = Fast3Cfg<T, NN>;                                                                                \
    a.pfDist = 0;                                                                                             \
    return fwd ? launch_bands(k_z_fast3_b<T, NN, true>, a, bt, a.numTiles, bands, C::threads, C::smem, s)     \
               : launch_bands(k_z_fast3_b<T, NN, false>, a, bt, a.numTiles, bands, C::threads, C::smem, s);   \
  }
  SB_BAND_DISPATCH(a.nz, POW2, THREE)
#undef POW2
#undef THREE
  return (int)cudaErrorInvalidValue;
}

template <typename T>
static int launch_y_bands(int fwd, const YArgs<T>& a0, const BandTable<T>& bt, int bands, cudaStream_t s) {
  YArgs<T> a = a0;
  const long long blocks = (long long)a.numXTiles * a.numPlanes;
  if (!a.ftw) {
    const size_t smem = 2 * ((size_t)a.ny << a.log2V) * sizeof(cx<T>);
    return fwd ? launch_bands(k_y_stage_b<T, true>, a, bt, blocks, bands, kBandThreads, smem, s)
               : launch_bands(k_y_stage_b<T, false>, a, bt, blocks, bands, kBandThreads, smem, s);
  }
#define POW2(NN)                                                                                          \
  {                                                                                                       \
    using C = FastCfg<T, NN>;                                                                             \
    a.pfDist = 0;                                                                                         \
    return fwd ? launch_bands(k_y_fast_b<T, NN, true>, a, bt, blocks, bands, C::threads, C::smem, s)      \
               : launch_bands(k_y_fast_b<T, NN, false>, a, bt, blocks, bands, C::threads, C::smem, s);    \
  }
#define THREE(NN)                                                                                         \
  {                                                                                                       \
    using C = Fast3Cfg<T, NN>;                                                                            \
    a.pfDist = 0;                                                                                         \
    return fwd ? launch_bands(k_y_fast3_b<T, NN, true>, a, bt, blocks, bands, C::threads, C::smem, s)     \
               : launch_bands(k_y_fast3_b<T, NN, false>, a, bt, blocks, bands, C::threads, C::smem, s);   \
  }
  SB_BAND_DISPATCH(a.ny, POW2, THREE)
#undef POW2
#undef THREE
  return (int)cudaErrorInvalidValue;
}

template <typename T>
static int launch_x_bands(int fwd, const XArgs<T>& a, const BandTable<T>& bt, int bands, cudaStream_t s) {
  const long long blocks = (long long)a.numRowTiles * a.numPlanes;
  if (!a.ftw) {
    const size_t smem = 2 * ((size_t)a.nx << a.log2V) * sizeof(cx<T>);
    return fwd ? launch_bands(k_x_stage_b<T, true>, a, bt, blocks, bands, kBandThreads, smem, s)
               : launch_bands(k_x_stage_b<T, false>, a, bt, blocks, bands, kBandThreads, smem, s);
  }
#define SB_X_CALL(KERNEL, C)                                                                                     \
  if (a.r2c)                                                                                                     \
    return fwd ? launch_bands(KERNEL<T, NN_, true, true>, a, bt, blocks, bands, C::threads, C::smem, s)          \
               : launch_bands(KERNEL<T, NN_, false, true>, a, bt, blocks, bands, C::threads, C::smem, s);        \
  return fwd ? launch_bands(KERNEL<T, NN_, true, false>, a, bt, blocks, bands, C::threads, C::smem, s)           \
             : launch_bands(KERNEL<T, NN_, false, false>, a, bt, blocks, bands, C::threads, C::smem, s);
#define POW2(NN)                 \
  {                              \
    constexpr int NN_ = NN;      \
    using C = FastCfgX<T, NN>;   \
    SB_X_CALL(k_x_fast_b, C)     \
  }
#define THREE(NN)                \
  {                              \
    constexpr int NN_ = NN;      \
    using C = Fast3CfgX<T, NN>;  \
    SB_X_CALL(k_x_fast3_b, C)    \
  }
  SB_BAND_DISPATCH(a.nx, POW2, THREE)
#undef POW2
#undef THREE
#undef SB_X_CALL
  return (int)cudaErrorInvalidValue;
}

}  // namespace sb

extern "C" {
int sb_band_kernel_available(int n, int registerFft) { return registerFft ? (sb::band_length(n) ? 1 : 0) : 1; }
int sb_launch_z_bands_f64(int f, const sb::ZArgs<double>* a, const sb::BandTable<double>* t, int bands, void* s) {
  sb_note_launches(1);
  return sb::launch_z_bands<double>(f, *a, *t, bands, static_cast<cudaStream_t>(s));
}
int sb_launch_z_bands_f32(int f, const sb::ZArgs<float>* a, const sb::BandTable<float>* t, int bands, void* s) {
  sb_note_launches(1);
  return sb::launch_z_bands<float>(f, *a, *t, bands, static_cast<cudaStream_t>(s));
}
int sb_launch_y_bands_f64(int f, const sb::YArgs<double>* a, const sb::BandTable<double>* t, int bands, void* s) {
  sb_note_launches(1);
  return sb::launch_y_bands<double>(f, *a, *t, bands, static_cast<cudaStream_t>(s));
}
int sb_launch_y_bands_f32(int f, const sb::YArgs<float>* a, const sb::BandTable<float>* t, int bands, void* s) {
  sb_note_launches(1);
  return sb::launch_y_bands<float>(f, *a, *t, bands, static_cast<cudaStream_t>(s));
}
int sb_launch_x_bands_f64(int f, const sb::XArgs<double>* a, const sb::BandTable<double>* t, int bands, void* s) {
  sb_note_launches(1);
  return sb::launch_x_bands<double>(f, *a, *t, bands, static_cast<cudaStream_t>(s));
}
int sb_launch_x_bands_f32(int f, const sb::XArgs<float>* a, const sb::BandTable<float>* t, int bands, void* s) {
  sb_note_launches(1);
  return sb::launch_x_bands<float>(f, *a, *t, bands, static_cast<cudaStream_t>(s));
}
}
