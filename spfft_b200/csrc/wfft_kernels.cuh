// wfft_kernels.cuh -- device building blocks of the warp-FFT stage kernels (wfft_xy.cu, wfft_z.cu),
// sm_100a, transform length 512, double precision and (two transforms per warp, WUnit) single precision.
//
// One warp = one transform (wfft.hpp: 16 values per lane, radix 32 = in-lane DFT16 + one shuffle step,
// ONE exchange through shared memory that never leaves the warp, radix 16 with lane twiddles held in
// registers). One CTA = 8 warps = one tile of 8 transforms and ONE 64 KB tile buffer S laid out
// [512 rows][8 x 16-byte chunks] with the 128-byte XOR pattern of a TMA tensor map
// (CU_TENSOR_MAP_SWIZZLE_128B): chunk c of row r sits at chunk c ^ (r & 7). Warp w owns "column" w of
// S: it is its private exchange region AND the place where the transposed side of a y / z tile is
// handed to / taken from the TMA engine, so the only CTA-wide synchronisation of a tile is around the
// bulk tensor copy itself.
//
// Replaces the cuFFT plans of the reference (src/fft/transform_1d_gpu.hpp:52-141,
// src/fft/transform_2d_gpu.hpp:51-140) together with its transpose / compression kernels
// (src/transpose/gpu_kernels/local_transpose_kernels.cu:48-201,
// src/compression/gpu_kernels/compression_kernels.cu:40-150).
#pragma once
#include <cuda_runtime.h>

#include "stage_kernels.hpp"
#include "tma_util.hpp"
#include "wfft.hpp"

namespace sb {

constexpr int kWThreads = 256;
constexpr size_t kWTileBytes = (size_t)kWN * kWWarps * sizeof(cx<double>);  // 64 KB

// w[i][L] = w_512^(2^i * L), forward sign; travels as a kernel parameter (2 KB)
template <typename T>
struct WTw4 {
  cx<T> w[4][32];
};

// What a 16-byte unit of a tile holds. Double precision: one complex number, a warp = one transform. Single
// precision: TWO complex numbers of two neighbouring transforms (columns x, x + 1 of a y tile, rows y, y + 1 of an
// x tile, sticks s, s + 1 of a z tile) computed together on the packed type f2 (wfft.hpp), a warp = two transforms.
// Tiles, sub-tiles, tensor maps and exchange addresses are the same bytes in both precisions; an item of the
// single-precision kernels covers 16 columns / rows / sticks, i.e. two tiles of the index plan.
template <typename T>
struct WUnit;
template <>
struct WUnit<double> {
  using Sc = double;
  static constexpr int kPer = 1;
};
template <>
struct WUnit<float> {
  using Sc = f2;
  static constexpr int kPer = 2;
};
// tile of the index plan (8 columns / sticks) and lane inside it of the FIRST transform of warp w of item `tile`
template <int PER>
__device__ __forceinline__ int w_plan_tile(int tile, int w) {
  return PER == 1 ? tile : 2 * tile + (w >> 2);
}
template <int PER>
__device__ __forceinline__ int w_plan_lane(int w) {
  return PER == 1 ? w : (2 * w) & 7;
}

// The lane twiddles live in shared memory (2 KB, staged once per CTA) and are re-read for every transform:
// as registers they would cost 16 of the 128, and read straight from the parameter bank (lane-dependent
// index) every access replays once per lane (ncu: LDC + short-scoreboard stalls, 20 % of all samples).
template <typename T>
__device__ __forceinline__ void w_stage_twiddles(cx<T>* sTw, const WTw4<T>& twp) {
  for (int i = threadIdx.x; i < 4 * 32; i += blockDim.x) sTw[i] = twp.w[i >> 5][i & 31];
}
// single precision: both halves of the packed twiddle are the same number
__device__ __forceinline__ void w_stage_twiddles(cx<f2>* sTw, const WTw4<float>& twp) {
  for (int i = threadIdx.x; i < 4 * 32; i += blockDim.x) {
    const cx<float> t = twp.w[i >> 5][i & 31];
    sTw[i] = mk<f2>(f2_make(t.x, t.x), f2_make(t.y, t.y));
  }
}
template <typename T>
__device__ __forceinline__ LaneTw<T> w_lane_twiddles(const cx<T>* sTw, int L) {
  LaneTw<T> t;
  t.w1 = sTw[L];
  t.w2 = sTw[32 + L];
  t.w4 = sTw[64 + L];
  t.w8 = sTw[96 + L];
  return t;
}

__device__ __forceinline__ void w_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void w_tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ cx<T> w_ldcg(const cx<T>* p) {
  const double2 q = __ldcg(reinterpret_cast<const double2*>(p));
  return mk<T>(q.x, q.y);
}
// store with an L2 eviction policy (createpolicy, tma_util.hpp)
template <typename T>
__device__ __forceinline__ void w_st_hint(cx<T>* p, cx<T> v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(policy) : "memory");
}
template <typename T>
__device__ __forceinline__ cx<T> w_ldcs(const cx<T>* p) {
  const double2 q = __ldcs(reinterpret_cast<const double2*>(p));
  return mk<T>(q.x, q.y);
}
template <typename T>
__device__ __forceinline__ void w_stcs(cx<T>* p, cx<T> v) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
}

// single precision: 8-byte accesses
__device__ __forceinline__ cx<float> w_ldcg(const cx<float>* p) {
  const float2 q = __ldcg(reinterpret_cast<const float2*>(p));
  return mk<float>(q.x, q.y);
}
__device__ __forceinline__ cx<float> w_ldcs(const cx<float>* p) {
  const float2 q = __ldcs(reinterpret_cast<const float2*>(p));
  return mk<float>(q.x, q.y);
}
__device__ __forceinline__ void w_stcs(cx<float>* p, cx<float> v) { __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y)); }
__device__ __forceinline__ void w_st_hint(cx<float>* p, cx<float> v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(policy) : "memory");
}
template <typename Sc>
__device__ __forceinline__ cx<Sc> w_zero() {
  return mk<Sc>(Sc(0.0), Sc(0.0));
}

// Element `off` of the warp's row(s) of a dense array (rows of n elements; single precision: the warp's two rows
// are adjacent) <-> unit
__device__ __forceinline__ cx<double> w_row_ldcg(const cx<double>* p, int) { return w_ldcg(p); }
__device__ __forceinline__ cx<f2> w_row_ldcg(const cx<float>* p, int n) { return unit_pack(w_ldcg(p), w_ldcg(p + n)); }
__device__ __forceinline__ void w_row_stcs(cx<double>* p, int, cx<double> v) { w_stcs(p, v); }
__device__ __forceinline__ void w_row_stcs(cx<float>* p, int n, cx<f2> v) {
  w_stcs(p, unit_lo(v));
  w_stcs(p + n, unit_hi(v));
}
__device__ __forceinline__ void w_row_st_hint(cx<double>* p, int, cx<double> v, uint64_t policy) { w_st_hint(p, v, policy); }
__device__ __forceinline__ void w_row_st_hint(cx<float>* p, int n, cx<f2> v, uint64_t policy) {
  w_st_hint(p, unit_lo(v), policy);
  w_st_hint(p + n, unit_hi(v), policy);
}

// 16-byte asynchronous copy global -> shared (LDGSTS): no register holds the data while it is in flight
__device__ __forceinline__ void w_cp_async16(void* smemDst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smemDst)), "l"(src) : "memory");
}
__device__ __forceinline__ void w_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// completion counter += 1 with release semantics at gpu scope (cumulative over what the thread observed
// through the preceding CTA / group barrier)
__device__ __forceinline__ void w_red_release(int* counter) {
  asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(counter) : "memory");
}

// Stage A of the length-512 plan: v[m] = x[L + 32 m] in.
template <typename T, bool BWD>
__device__ __forceinline__ void w512_head(cx<T>* v, int L) {
  using P = WPlan<T, 512>;
  P::template stage_a_local<BWD>(v, L);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const cx<T> recv = shfl_xor_cx<T>(v[8 + i], 16);
    P::stage_a_combine(v[i], v[8 + i], recv, L);
  }
}
template <int W>
__device__ __forceinline__ void w_group_sync(int g) {
  if constexpr (W == 8) {
    __syncthreads();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(W * 32) : "memory");
  }
}

template <typename T>
__device__ __forceinline__ cx<T>* w_at(cx<T>* S, unsigned byteOff) {
  return reinterpret_cast<cx<T>*>(reinterpret_cast<char*>(S) + byteOff);
}
// The exchange inside the warp's column of the sub-tile (private to the warp).
template <typename T, int W>
__device__ __forceinline__ void w512_exchange(cx<T>* v, cx<T>* S, const WAddr& ad) {
  static_assert(sizeof(cx<T>) == 16, "tiles are made of 16-byte units (cx<double> / cx<f2>)");
#pragma unroll
  for (int i = 0; i < 16; ++i) *w_at(S, w_xw_off<W>(ad, i)) = v[i];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = *w_at(S, w_xr_off<W>(ad, r));
  __syncwarp();
}
// The same exchange in a FLAT private region of 512 elements (a row that a bulk copy landed in natural order):
// slot(n) = n ^ ((n >> 5) & 7) (WPlan<T, 512>::slot), i.e. natural element L + 32 m at (L << 4) + m * 512 bytes,
// exchange write at (xwFlat ^ (q << 4)) + (i >> 3) * 256, exchange read at ((L << 4) ^ (q << 4)) + r * 512.
template <typename T>
__device__ __forceinline__ void w512_exchange_flat(cx<T>* v, cx<T>* R, int L) {
  const unsigned nat = (unsigned)L << 4;
  const unsigned j = L & 15, h = L >> 4;
  const unsigned xwFlat = ((32u * j + 8u * h) << 4) | ((j & 7) << 4);
#pragma unroll
  for (int i = 0; i < 16; ++i) *w_at(R, (xwFlat ^ ((unsigned)(i & 7) << 4)) + (unsigned)(i >> 3) * 256u) = v[i];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = *w_at(R, (nat ^ ((unsigned)(r & 7) << 4)) + (unsigned)r * 512u);
  __syncwarp();
}
__device__ __forceinline__ void w512_flat_load(cx<double>* v, const cx<double>* R, int L) {
#pragma unroll
  for (int m = 0; m < 16; ++m) v[m] = R[L + 32 * m];
}
// single precision: the region holds the warp's two rows one after the other, as the bulk copy delivered them
__device__ __forceinline__ void w512_flat_load(cx<f2>* v, const cx<f2>* R, int L) {
  const cx<float>* r = reinterpret_cast<const cx<float>*>(R);
#pragma unroll
  for (int m = 0; m < 16; ++m) v[m] = unit_pack(r[L + 32 * m], r[kWN + L + 32 * m]);
}
// global -> shared bulk copy (contiguous bytes, multiple of 16), completes on `bar`; L2 policy of the source lines
__device__ __forceinline__ void w_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
               : "memory");
}

// natural-order access of the warp's column (tile side: the layout bulk tensor copies move to / from global memory,
// so a single-precision unit is stored as its two complex numbers lie in memory)
template <typename T, int W>
__device__ __forceinline__ void w512_col_load(cx<T>* v, cx<T>* S, const WAddr& ad) {
#pragma unroll
  for (int m = 0; m < 16; ++m) v[m] = unit_transpose<T>(*w_at(S, w_nat_off<W>(ad, m)));
}
template <typename T, int W>
__device__ __forceinline__ void w512_col_store(const cx<T>* v, cx<T>* S, const WAddr& ad) {
#pragma unroll
  for (int m = 0; m < 16; ++m) *w_at(S, w_nat_off<W>(ad, m)) = unit_transpose<T>(v[m]);
}
// Stage B: afterwards v[q] = X[L + 32 q].
template <typename T, bool BWD>
__device__ __forceinline__ void w512_tail(cx<T>* v, const cx<T>* sTw, int L) {
  const LaneTw<T> tw = w_lane_twiddles<T>(sTw, L);
  twiddle_dft16<T, BWD>(v, tw);
}

// The 16 inverse-map entries of lane L of column w of a tile (maps of index_plan.cpp, layout
// [tile][lane_of_tile * 64 + (n & 63)][n >> 6], n = position along the transform): entry m belongs to
// n = L + 32 m. 0xFFFF = no sparse element at n.
struct WInv16 {
  unsigned short i[16];
};
__device__ __forceinline__ const unsigned short* w_inv_ptr(const unsigned short* inv, long long tile, int w, int L) {
  return inv + ((size_t)tile * 512 + (size_t)w * 64 + L) * 8;  // second half: + 32 * 8
}
__device__ __forceinline__ WInv16 w_unpack_inv(uint4 q0, uint4 q1) {
  WInv16 r;
  r.i[0] = q0.x & 0xFFFF;  r.i[2] = q0.x >> 16;
  r.i[4] = q0.y & 0xFFFF;  r.i[6] = q0.y >> 16;
  r.i[8] = q0.z & 0xFFFF;  r.i[10] = q0.z >> 16;
  r.i[12] = q0.w & 0xFFFF; r.i[14] = q0.w >> 16;
  r.i[1] = q1.x & 0xFFFF;  r.i[3] = q1.x >> 16;
  r.i[5] = q1.y & 0xFFFF;  r.i[7] = q1.y >> 16;
  r.i[9] = q1.z & 0xFFFF;  r.i[11] = q1.z >> 16;
  r.i[13] = q1.w & 0xFFFF; r.i[15] = q1.w >> 16;
  return r;
}
__device__ __forceinline__ WInv16 w_load_inv(const unsigned short* inv, long long tile, int w, int L) {
  const unsigned short* p = w_inv_ptr(inv, tile, w, L);
  return w_unpack_inv(__ldg(reinterpret_cast<const uint4*>(p)), __ldg(reinterpret_cast<const uint4*>(p + 32 * 8)));
}
constexpr unsigned short kWNone = 0xFFFF;
// Inverse-map entries of a thread's part held in shared memory (cp.async, one part ahead): PER columns x two
// 16-byte halves per thread. The second column of a single-precision unit is the next lane of the same plan tile.
template <int PER>
__device__ __forceinline__ void w_inv_prefetch(uint4 (*sInvPart)[kWThreads], const unsigned short* inv, long long planTile,
                                               int planLane, int L, int tid) {
#pragma unroll
  for (int c = 0; c < PER; ++c) {
    const unsigned short* p = w_inv_ptr(inv, planTile, planLane + c, L);
    w_cp_async16(&sInvPart[2 * c][tid], p);
    w_cp_async16(&sInvPart[2 * c + 1][tid], p + 32 * 8);
  }
}
template <int PER>
struct WInvUnit {
  WInv16 c[PER];
};
template <int PER>
__device__ __forceinline__ WInvUnit<PER> w_inv_read(const uint4 (*sInvPart)[kWThreads], int tid) {
  WInvUnit<PER> r;
#pragma unroll
  for (int c = 0; c < PER; ++c) r.c[c] = w_unpack_inv(sInvPart[2 * c][tid], sInvPart[2 * c + 1][tid]);
  return r;
}
template <int PER>
__device__ __forceinline__ WInvUnit<PER> w_inv_load(const unsigned short* inv, long long planTile, int planLane, int L) {
  WInvUnit<PER> r;
#pragma unroll
  for (int c = 0; c < PER; ++c) r.c[c] = w_load_inv(inv, planTile, planLane + c, L);
  return r;
}
// gather / scatter of a unit's 16 values through the inverse map (base = first element of the plan tile's block)
__device__ __forceinline__ void w_gather_cs(cx<double>* v, const cx<double>* base, const WInvUnit<1>& iv) {
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    v[m] = mk<double>(0, 0);
    if (iv.c[0].i[m] != kWNone) v[m] = w_ldcs(base + iv.c[0].i[m]);  // read once: evict first
  }
}
__device__ __forceinline__ void w_gather_cs(cx<f2>* v, const cx<float>* base, const WInvUnit<2>& iv) {
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    cx<float> c0 = mk<float>(0.f, 0.f), c1 = mk<float>(0.f, 0.f);
    if (iv.c[0].i[m] != kWNone) c0 = w_ldcs(base + iv.c[0].i[m]);
    if (iv.c[1].i[m] != kWNone) c1 = w_ldcs(base + iv.c[1].i[m]);
    v[m] = unit_pack(c0, c1);
  }
}
__device__ __forceinline__ void w_scatter_cs(cx<double>* base, const cx<double>* v, const WInvUnit<1>& iv) {
#pragma unroll
  for (int m = 0; m < 16; ++m)
    if (iv.c[0].i[m] != kWNone) w_stcs(base + iv.c[0].i[m], v[m]);  // written once: evict first
}
__device__ __forceinline__ void w_scatter_cs(cx<float>* base, const cx<f2>* v, const WInvUnit<2>& iv) {
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    if (iv.c[0].i[m] != kWNone) w_stcs(base + iv.c[0].i[m], unit_lo(v[m]));
    if (iv.c[1].i[m] != kWNone) w_stcs(base + iv.c[1].i[m], unit_hi(v[m]));
  }
}
__device__ __forceinline__ void w_scatter_plain(cx<double>* base, const cx<double>* v, const WInvUnit<1>& iv) {
#pragma unroll
  for (int m = 0; m < 16; ++m)
    if (iv.c[0].i[m] != kWNone) base[iv.c[0].i[m]] = v[m];
}
__device__ __forceinline__ void w_scatter_plain(cx<float>* base, const cx<f2>* v, const WInvUnit<2>& iv) {
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    if (iv.c[0].i[m] != kWNone) base[iv.c[0].i[m]] = unit_lo(v[m]);
    if (iv.c[1].i[m] != kWNone) base[iv.c[1].i[m]] = unit_hi(v[m]);
  }
}

}  // namespace sb
