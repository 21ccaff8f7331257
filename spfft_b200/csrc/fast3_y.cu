// fast3_y.cu -- y stage (plane-major sticks <-> xy planes) for N = 3 * 2^k, sm_100a.
#include "fast3_launch.cuh"

namespace sb {

template <typename T, int N, bool FWD, bool WIRE = false>
__global__ void __launch_bounds__(Fast3Cfg<T, N>::threads, Fast3Cfg<T, N>::minBlocks)
    k_y_fast3(const __grid_constant__ YArgs<T> a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  const Ctx ctx{Fast3Cfg<T, N>::threads};
  if (FWD)
    y_forward_fast3<T, N, WIRE>(a, (int)blockIdx.x, ctx, S);
  else
    y_backward_fast3<T, N, WIRE>(a, (int)blockIdx.x, ctx, S);
}

template <typename T, int N>
static int launch_y3_n(int forward, const YArgs<T>& a0, cudaStream_t s) {
  using C = Fast3Cfg<T, N>;
  YArgs<T> a = a0;
  a.pfDist = (tune_flags() & 1) ? resident_ctas(C::minBlocks) : 0;
  const long long blocks = (long long)a.numXTiles * a.numPlanes;
  if (a.wireF32) {  // single-precision wire format of a distributed double-precision transform
    if constexpr (sizeof(T) == 8) {
      return forward ? launch_fast(k_y_fast3<T, N, true, true>, a, blocks, C::threads, C::smem, s)
                     : launch_fast(k_y_fast3<T, N, false, true>, a, blocks, C::threads, C::smem, s);
    } else {
      return (int)cudaErrorInvalidValue;
    }
  }
  return forward ? launch_fast(k_y_fast3<T, N, true>, a, blocks, C::threads, C::smem, s)
                 : launch_fast(k_y_fast3<T, N, false>, a, blocks, C::threads, C::smem, s);
}

template <typename T>
int launch_y_fast3(int forward, const YArgs<T>& a, cudaStream_t s) {
#define CALL(NN) return launch_y3_n<T, NN>(forward, a, s)
  SB_FAST3_DISPATCH(a.ny, CALL)
#undef CALL
  return (int)cudaErrorInvalidValue;
}
template int launch_y_fast3<double>(int, const YArgs<double>&, cudaStream_t);
template int launch_y_fast3<float>(int, const YArgs<float>&, cudaStream_t);

}  // namespace sb
