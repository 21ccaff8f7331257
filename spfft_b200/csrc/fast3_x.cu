// fast3_x.cu -- x stage on complex rows (C2C) and real rows (R2C / C2R) for N = 3 * 2^k, sm_100a.
#include "fast3_launch.cuh"

namespace sb {

template <typename T, int N, bool FWD>
__global__ void __launch_bounds__(Fast3CfgX<T, N>::threads, Fast3CfgX<T, N>::minBlocks)
    k_x_fast3(const __grid_constant__ XArgs<T> a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  x_c2c_fast3<T, N, !FWD>(a, (int)blockIdx.x, Ctx{Fast3CfgX<T, N>::threads}, S);
}

template <typename T, int N, bool FWD>
__global__ void __launch_bounds__(Fast3CfgX<T, N>::threads, Fast3CfgX<T, N>::minBlocks)
    k_x_real_fast3(const __grid_constant__ XArgs<T> a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  x_r2c_fast3<T, N, !FWD>(a, (int)blockIdx.x, Ctx{Fast3CfgX<T, N>::threads}, S);
}

template <typename T, int N>
static int launch_x3_n(int forward, const XArgs<T>& a, cudaStream_t s) {
  using C = Fast3CfgX<T, N>;
  const long long blocks = (long long)a.numRowTiles * a.numPlanes;
  if (a.r2c)
    return forward ? launch_fast(k_x_real_fast3<T, N, true>, a, blocks, C::threads, C::smem, s)
                   : launch_fast(k_x_real_fast3<T, N, false>, a, blocks, C::threads, C::smem, s);
  return forward ? launch_fast(k_x_fast3<T, N, true>, a, blocks, C::threads, C::smem, s)
                 : launch_fast(k_x_fast3<T, N, false>, a, blocks, C::threads, C::smem, s);
}

template <typename T>
int launch_x_fast3(int forward, const XArgs<T>& a, cudaStream_t s) {
#define CALL(NN) return launch_x3_n<T, NN>(forward, a, s)
  SB_FAST3_DISPATCH(a.nx, CALL)
#undef CALL
  return (int)cudaErrorInvalidValue;
}
template int launch_x_fast3<double>(int, const XArgs<double>&, cudaStream_t);
template int launch_x_fast3<float>(int, const XArgs<float>&, cudaStream_t);

}  // namespace sb
