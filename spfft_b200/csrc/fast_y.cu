// fast_y.cu -- y stage (plane-major sticks <-> xy planes) on the register FFT, sm_100a.
#include "fast_launch.cuh"

namespace sb {

template <typename T, int N, bool FWD, bool WIRE = false>
__global__ void __launch_bounds__(FastCfg<T, N>::threads, FastCfg<T, N>::minBlocks)
    k_y_fast(const __grid_constant__ YArgs<T> a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  if (FWD)
    y_forward_fast<T, N, WIRE>(a, (int)blockIdx.x, Ctx{FastCfg<T, N>::threads}, S);
  else
    y_backward_fast<T, N, WIRE>(a, (int)blockIdx.x, Ctx{FastCfg<T, N>::threads}, S);
}

template <typename T, int N>
static int launch_y_n(int forward, const YArgs<T>& a0, cudaStream_t s) {
  using C = FastCfg<T, N>;
  YArgs<T> a = a0;
  a.pfDist = (tune_flags() & 1) ? resident_ctas(C::minBlocks) : 0;
  if constexpr (C::threads > 1024) {
    return (int)cudaErrorInvalidValue;
  } else {
    if (a.wireF32) {  // single-precision wire format of a distributed double-precision transform
      if constexpr (sizeof(T) == 8) {
        return forward ? launch_fast(k_y_fast<T, N, true, true>, a, (long long)a.numXTiles * a.numPlanes, C::threads, C::smem, s)
                       : launch_fast(k_y_fast<T, N, false, true>, a, (long long)a.numXTiles * a.numPlanes, C::threads, C::smem, s);
      } else {
        return (int)cudaErrorInvalidValue;
      }
    }
    return forward ? launch_fast(k_y_fast<T, N, true>, a, (long long)a.numXTiles * a.numPlanes, C::threads, C::smem, s)
                   : launch_fast(k_y_fast<T, N, false>, a, (long long)a.numXTiles * a.numPlanes, C::threads, C::smem, s);
  }
}

template <typename T>
int launch_y_fast(int forward, const YArgs<T>& a, cudaStream_t s) {
#define CALL(NN) return launch_y_n<T, NN>(forward, a, s)
  SB_FAST_DISPATCH(a.ny, CALL)
#undef CALL
  return (int)cudaErrorInvalidValue;
}
template int launch_y_fast<double>(int, const YArgs<double>&, cudaStream_t);
template int launch_y_fast<float>(int, const YArgs<float>&, cudaStream_t);

}  // namespace sb
