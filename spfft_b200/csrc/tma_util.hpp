// tma_util.hpp -- tensor maps (TMA descriptors) for the 8-lane tiles of the stick / plane buffers,
// created through the driver entry point (no link-time dependency on libcuda), and the device-side
// mbarrier / cp.async.bulk.tensor primitives of the warp-FFT stage kernels (wfft_*.cu).
//
// A tile = 8 consecutive complex columns (128 bytes in double precision) x N rows of a row-major
// 2-D array [rows][pitch] (the plane-major stick buffer, or one xy plane [y][x]); batches of such
// arrays (planes) are the third tensor dimension. With CU_TENSOR_MAP_SWIZZLE_128B the 16-byte chunk c
// of tile row r lands at chunk c ^ (r & 7) of the 128-byte shared-memory row r -- exactly the layout
// in which one warp per column reads / writes its column without bank conflicts (wfft.hpp).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace sb {

using TensorMap = CUtensorMap;

// rows x pitch elements of `elemBytes`-byte complex numbers, `batch` arrays `batchStride` elements apart;
// box = `boxCols` complex columns x `boxRows` rows x 1. Returns a cudaError_t value.
inline int make_tile_map(TensorMap* map, void* base, int elemBytes, long long cols, long long rows,
                         long long pitch, long long batch, long long batchStride, int boxCols,
                         int boxRows) {
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                          const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                          CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<Fn>(fn);
  }();
  if (!encode) return (int)cudaErrorNotSupported;
  // element type: the real scalar (double / float); a complex number = 2 scalars
  const CUtensorMapDataType dt = elemBytes == 16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const int sb = elemBytes / 2;
  const cuuint64_t dims[3] = {(cuuint64_t)(2 * cols), (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
  const cuuint64_t strides[2] = {(cuuint64_t)(pitch * elemBytes), (cuuint64_t)((batch > 1 ? batchStride : pitch * rows) * elemBytes)};
  const cuuint32_t box[3] = {(cuuint32_t)(2 * boxCols), (cuuint32_t)boxRows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const int rowBytes = 2 * boxCols * sb;
  const CUtensorMapSwizzle swz = rowBytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : rowBytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : rowBytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUresult r = encode(map, dt, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared tile (box of the map at element coordinates (c0 scalars, row, batch)), completes on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const TensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_addr(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(bar))
      : "memory");
}
// L2 eviction policies for the cache-hint forms of the bulk copies
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_store_3d_hint(const TensorMap* map, int c0, int c1, int c2, const void* src, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;" ::"l"(map), "r"(c0),
               "r"(c1), "r"(c2), "r"(smem_addr(src)), "l"(policy)
               : "memory");
}
// shared -> global tile
__device__ __forceinline__ void tma_store_3d(const TensorMap* map, int c0, int c1, int c2, const void* src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0),
               "r"(c1), "r"(c2), "r"(smem_addr(src))
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

}  // namespace sb
