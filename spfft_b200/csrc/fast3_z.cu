// fast3_z.cu -- z stage (sparse values <-> plane-major sticks) for N = 3 * 2^k, sm_100a.
#include "fast3_launch.cuh"

namespace sb {

template <typename T, int N, bool FWD, bool WIRE = false>
__global__ void __launch_bounds__(Fast3Cfg<T, N>::threads, Fast3Cfg<T, N>::minBlocks)
    k_z_fast3(const __grid_constant__ ZArgs<T> a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  const Ctx ctx{Fast3Cfg<T, N>::threads};
  if (FWD)
    z_forward_fast3<T, N, WIRE>(a, (int)blockIdx.x, ctx, S);
  else
    z_backward_fast3<T, N, WIRE>(a, (int)blockIdx.x, ctx, S);
}

template <typename T, int N>
static int launch_z3_n(int forward, const ZArgs<T>& a0, cudaStream_t s) {
  using C = Fast3Cfg<T, N>;
  ZArgs<T> a = a0;
  a.pfDist = (tune_flags() & 1) ? resident_ctas(C::minBlocks) : 0;
  if (a.wireF32) {  // single-precision wire format of a distributed double-precision transform
    if constexpr (sizeof(T) == 8) {
      return forward ? launch_fast(k_z_fast3<T, N, true, true>, a, a.numTiles, C::threads, C::smem, s)
                     : launch_fast(k_z_fast3<T, N, false, true>, a, a.numTiles, C::threads, C::smem, s);
    } else {
      return (int)cudaErrorInvalidValue;
    }
  }
  return forward ? launch_fast(k_z_fast3<T, N, true>, a, a.numTiles, C::threads, C::smem, s)
                 : launch_fast(k_z_fast3<T, N, false>, a, a.numTiles, C::threads, C::smem, s);
}

template <typename T>
int launch_z_fast3(int forward, const ZArgs<T>& a, cudaStream_t s) {
#define CALL(NN) return launch_z3_n<T, NN>(forward, a, s)
  SB_FAST3_DISPATCH(a.nz, CALL)
#undef CALL
  return (int)cudaErrorInvalidValue;
}
template int launch_z_fast3<double>(int, const ZArgs<double>&, cudaStream_t);
template int launch_z_fast3<float>(int, const ZArgs<float>&, cudaStream_t);

}  // namespace sb
