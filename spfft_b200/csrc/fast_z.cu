// fast_z.cu -- z stage (sparse values <-> plane-major sticks) on the register FFT, sm_100a.
#include "fast_launch.cuh"

namespace sb {

template <typename T, int N, bool FWD, bool WIRE = false>
__global__ void __launch_bounds__(FastCfg<T, N>::threads, FastCfg<T, N>::minBlocks)
    k_z_fast(const __grid_constant__ ZArgs<T> a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smemRaw);
  const Ctx ctx{FastCfg<T, N>::threads};
  z_fast_any<T, N, FWD, WIRE>(a, (int)blockIdx.x, ctx, S);
}

template <typename T, int N>
static int launch_z_n(int forward, const ZArgs<T>& a0, cudaStream_t s) {
  using C = FastCfg<T, N>;
  ZArgs<T> a = a0;
  a.pfDist = (tune_flags() & 1) ? resident_ctas(C::minBlocks) : 0;
  if constexpr (C::threads > 1024) {
    return (int)cudaErrorInvalidValue;
  } else {
    if (a.wireF32) {  // single-precision wire format of a distributed double-precision transform
      if constexpr (sizeof(T) == 8) {
        return forward ? launch_fast(k_z_fast<T, N, true, true>, a, a.numTiles, C::threads, C::smem, s)
                       : launch_fast(k_z_fast<T, N, false, true>, a, a.numTiles, C::threads, C::smem, s);
      } else {
        return (int)cudaErrorInvalidValue;
      }
    }
    return forward ? launch_fast(k_z_fast<T, N, true>, a, a.numTiles, C::threads, C::smem, s)
                   : launch_fast(k_z_fast<T, N, false>, a, a.numTiles, C::threads, C::smem, s);
  }
}

template <typename T>
int launch_z_fast(int forward, const ZArgs<T>& a, cudaStream_t s) {
#define CALL(NN) return launch_z_n<T, NN>(forward, a, s)
  SB_FAST_DISPATCH(a.nz, CALL)
#undef CALL
  return (int)cudaErrorInvalidValue;
}
template int launch_z_fast<double>(int, const ZArgs<double>&, cudaStream_t);
template int launch_z_fast<float>(int, const ZArgs<float>&, cudaStream_t);

}  // namespace sb
