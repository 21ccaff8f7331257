// fast_x.cu -- x stage on complex rows (C2C) and real rows (R2C / C2R) on the register FFT, sm_100a.
#include "fast_launch.cuh"

namespace sb {

template <typename T, int N, bool FWD>
__global__ void __launch_bounds__(FastCfgX<T, N>::threads, FastCfgX<T, N>::minBlocks)
    k_x_fast(const __grid_constant__ XArgs<T> a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  x_c2c_fast<T, N, !FWD>(a, (int)blockIdx.x, Ctx{FastCfgX<T, N>::threads}, S);
}

// real rows (R2C forward / C2R backward)
template <typename T, int N, bool FWD>
__global__ void __launch_bounds__(FastCfgX<T, N>::threads, FastCfgX<T, N>::minBlocks)
    k_x_real_fast(const __grid_constant__ XArgs<T> a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  x_r2c_fast<T, N, !FWD>(a, (int)blockIdx.x, Ctx{FastCfgX<T, N>::threads}, S);
}

template <typename T, int N>
static int launch_x_n(int forward, const XArgs<T>& a, cudaStream_t s) {
  using C = FastCfgX<T, N>;
  if constexpr (C::threads > 1024) {
    return (int)cudaErrorInvalidValue;
  } else {
    const long long blocks = (long long)a.numRowTiles * a.numPlanes;
    if (a.r2c)
      return forward ? launch_fast(k_x_real_fast<T, N, true>, a, blocks, C::threads, C::smem, s)
                     : launch_fast(k_x_real_fast<T, N, false>, a, blocks, C::threads, C::smem, s);
    return forward ? launch_fast(k_x_fast<T, N, true>, a, blocks, C::threads, C::smem, s)
                   : launch_fast(k_x_fast<T, N, false>, a, blocks, C::threads, C::smem, s);
  }
}

template <typename T>
int launch_x_fast(int forward, const XArgs<T>& a, cudaStream_t s) {
#define CALL(NN) return launch_x_n<T, NN>(forward, a, s)
  SB_FAST_DISPATCH(a.nx, CALL)
#undef CALL
  return (int)cudaErrorInvalidValue;
}
template int launch_x_fast<double>(int, const XArgs<double>&, cudaStream_t);
template int launch_x_fast<float>(int, const XArgs<float>&, cudaStream_t);

}  // namespace sb
