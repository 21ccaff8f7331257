// wfft.hpp -- warp-autonomous register FFT: 16 complex values per lane, ONE shared-memory exchange
// per transform that never leaves the warp (no CTA barrier anywhere inside a transform).
//
//   N = 512 : one warp (32 lanes x 16 values) per transform
//             stage A = radix 32 = radix-16 DFT inside the lane  +  radix-2 across the lane pair
//             (L, L^16) with the compile-time twiddles w32^q, exchanged with 8 shuffles per lane;
//             exchange through the warp's private 8 KB (double) region of shared memory;
//             stage B = radix 16 with the per-lane twiddles w512^(r*L).
//   N = 256 : half a warp (16 lanes x 16 values) per transform, two transforms per warp
//             stage A = radix 16, exchange, stage B = radix 16 with w256^(r*j).
//
// Input and output are both in natural "lane-strided" order: lane L (of the T lanes of a transform)
// holds x[L + T*m] in register m before and X[L + T*m] after the transform. Consecutive lanes hold
// consecutive elements, so global loads / stores of rows are coalesced and a transposed store into
// an 8-lane tile [n][lane] with the 128-byte XOR pattern of a TMA tensor map is conflict free.
//
// Compared with the 8-values-per-thread plan of fast_fft.hpp (two CTA-wide exchanges, three barriers
// per tile) this halves the shared-memory wavefronts per point and removes every CTA barrier from
// the transform itself; what a tile still needs is ONE barrier / mbarrier wait on the side where
// eight transforms share 128-byte global segments (z / y stages).
//
// Index algebra (Stockham autosort, as in fast_fft.hpp): stage with radix R, stride NS, butterfly b:
//   inputs n = b + r*N/R, twiddled by w_{NS*R}^{r*(b mod NS)};  outputs (b-k)*R + k + q*NS, k = b mod NS.
// N = 512: stage A (R = 32, NS = 1): b = j = L & 15, input r' = 2m + h with h = L >> 4, i.e. element
//   j + 16*(2m + h) = L + 32 m; output slot 32 j + k'. Stage B (R = 16, NS = 32): b = L, inputs L + 32 r,
//   twiddle w512^(r L), outputs L + 32 q.
//
// The arithmetic bodies are SB_HD (host + device) so that tests/emu can run the identical index
// algebra lane by lane on the CPU; only the exchange primitives (shuffle, __syncwarp) are device code.
//
// Replaces the cuFFT plans of the reference (src/fft/transform_1d_gpu.hpp:52-141,
// src/fft/transform_2d_gpu.hpp:51-140): unnormalised DFT, sign + backward / - forward.
#pragma once
#include <cmath>
#include <cstring>

#include "cx.hpp"
#include "fft_tile.hpp"

namespace sb {

// v * (c + s*i*sn), s = +1 (BWD) / -1 (forward)
template <bool BWD, typename T>
SB_HD cx<T> mul_w(cx<T> v, T c, T sn) {
  const T si = BWD ? sn : -sn;
  return mk<T>(v.x * c - v.y * si, v.x * si + v.y * c);
}

// -v if `flip` (lane-dependent), v otherwise: an integer XOR of the sign bits on the GPU (ALU pipe, no
// branch, no select) instead of a negation on the fp64 pipe behind a divergent branch.
template <typename T>
SB_HD T flip_sign(T x, unsigned mask /* 0 or 0x80000000 */) {
#if SB_ON_GPU
  if constexpr (sizeof(T) == 8) {
    return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
  } else {
    return __int_as_float(__float_as_int(x) ^ (int)mask);
  }
#else
  return mask ? -x : x;
#endif
}
template <typename T>
SB_HD cx<T> flip_sign(cx<T> v, unsigned mask) {
  return mk<T>(flip_sign<T>(v.x, mask), flip_sign<T>(v.y, mask));
}

// --------------------------------------------------------------------------------------------
// f2: TWO single-precision values in one 64-bit register, element-wise arithmetic on the packed fp32 pipe of
// sm_100 (FADD2 / FMUL2 / FFMA2: one instruction per pair). A warp that computes on cx<f2> runs two single-
// precision transforms at the instruction count of one double-precision transform, and a cx<f2> is a 16-byte
// unit like cx<double>, so the tile geometry below (TMA swizzle, exchange addresses) serves both precisions.
// Half `lo` belongs to the first transform of the pair, `hi` to the second. On the CPU (tests/emu) the same
// type is a plain pair of floats.
// --------------------------------------------------------------------------------------------
struct f2 {
#if SB_ON_GPU
  unsigned long long r;
#else
  float a, b;
#endif
  f2() = default;
  SB_HD explicit f2(double c);
};
#if SB_ON_GPU
SB_HD f2 f2_make(float lo, float hi) {
  f2 v;
#ifdef __CUDA_ARCH__
  asm("mov.b64 %0, {%1, %2};" : "=l"(v.r) : "f"(lo), "f"(hi));  // (ptxas sees a register pair, not integer arithmetic)
#else
  unsigned int ul, uh;
  memcpy(&ul, &lo, 4);
  memcpy(&uh, &hi, 4);
  v.r = (unsigned long long)ul | ((unsigned long long)uh << 32);
#endif
  return v;
}
SB_HD float f2_lo(f2 v) {
#ifdef __CUDA_ARCH__
  float lo;
  asm("{ .reg .b32 t; mov.b64 {%0, t}, %1; }" : "=f"(lo) : "l"(v.r));
  return lo;
#else
  const unsigned int u = (unsigned)(v.r & 0xffffffffULL);
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
SB_HD float f2_hi(f2 v) {
#ifdef __CUDA_ARCH__
  float hi;
  asm("{ .reg .b32 t; mov.b64 {t, %0}, %1; }" : "=f"(hi) : "l"(v.r));
  return hi;
#else
  const unsigned int u = (unsigned)(v.r >> 32);
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
#else
SB_HD f2 f2_make(float lo, float hi) {
  f2 v;
  v.a = lo;
  v.b = hi;
  return v;
}
SB_HD float f2_lo(f2 v) { return v.a; }
SB_HD float f2_hi(f2 v) { return v.b; }
#endif
SB_HD f2::f2(double c) { *this = f2_make((float)c, (float)c); }
#if defined(__CUDA_ARCH__)
SB_HD f2 operator+(f2 x, f2 y) {
  f2 v;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(v.r) : "l"(x.r), "l"(y.r));
  return v;
}
SB_HD f2 operator-(f2 x, f2 y) {
  f2 v;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(v.r) : "l"(x.r), "l"(y.r));
  return v;
}
SB_HD f2 operator*(f2 x, f2 y) {
  f2 v;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(v.r) : "l"(x.r), "l"(y.r));
  return v;
}
// x * y + z
SB_HD f2 f2_fma(f2 x, f2 y, f2 z) {
  f2 v;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v.r) : "l"(x.r), "l"(y.r), "l"(z.r));
  return v;
}
#else
SB_HD f2 operator+(f2 x, f2 y) { return f2_make(f2_lo(x) + f2_lo(y), f2_hi(x) + f2_hi(y)); }
SB_HD f2 operator-(f2 x, f2 y) { return f2_make(f2_lo(x) - f2_lo(y), f2_hi(x) - f2_hi(y)); }
SB_HD f2 operator*(f2 x, f2 y) { return f2_make(f2_lo(x) * f2_lo(y), f2_hi(x) * f2_hi(y)); }
SB_HD f2 f2_fma(f2 x, f2 y, f2 z) {
  return f2_make(fmaf(f2_lo(x), f2_lo(y), f2_lo(z)), fmaf(f2_hi(x), f2_hi(y), f2_hi(z)));
}
#endif
// (ptxas folds the two negations into the operand modifiers of the packed instruction that consumes them)
SB_HD f2 operator-(f2 x) { return f2_make(-f2_lo(x), -f2_hi(x)); }
// the products of the butterflies with ONE fused multiply-add per component pair (the generic cx<T> operators
// rely on the compiler contracting a * b - c * d, which it cannot do across the packed instructions)
SB_HD cx<f2> operator*(cx<f2> a, cx<f2> b) {
  return mk<f2>(f2_fma(a.x, b.x, -(a.y * b.y)), f2_fma(a.x, b.y, a.y * b.x));
}
template <>
SB_HD cx<f2> mul_w<true, f2>(cx<f2> v, f2 c, f2 sn) {
  return mk<f2>(f2_fma(v.x, c, -(v.y * sn)), f2_fma(v.x, sn, v.y * c));
}
template <>
SB_HD cx<f2> mul_w<false, f2>(cx<f2> v, f2 c, f2 sn) {
  return mk<f2>(f2_fma(v.x, c, v.y * sn), f2_fma(v.y, c, -(v.x * sn)));
}
template <>
SB_HD f2 flip_sign<f2>(f2 x, unsigned mask) {
#if SB_ON_GPU
  f2 v;
  v.r = x.r ^ (((unsigned long long)mask << 32) | (unsigned long long)mask);
  return v;
#else
  return mask ? -x : x;
#endif
}
// the two complex numbers of a unit as they lie in memory, (re0, im0, re1, im1), <-> the packed form
// ((re0, re1), (im0, im1)); the permutation is its own inverse. Identity for double precision.
template <typename T>
SB_HD cx<T> unit_transpose(cx<T> v) {
  return v;
}
template <>
SB_HD cx<f2> unit_transpose<f2>(cx<f2> v) {
  return mk<f2>(f2_make(f2_lo(v.x), f2_lo(v.y)), f2_make(f2_hi(v.x), f2_hi(v.y)));
}
SB_HD cx<f2> unit_pack(cx<float> c0, cx<float> c1) { return mk<f2>(f2_make(c0.x, c1.x), f2_make(c0.y, c1.y)); }
SB_HD cx<float> unit_lo(cx<f2> v) { return mk<float>(f2_lo(v.x), f2_lo(v.y)); }
SB_HD cx<float> unit_hi(cx<f2> v) { return mk<float>(f2_hi(v.x), f2_hi(v.y)); }

// 16-point DFT in registers, natural order in and out (4 x 4 decomposition, 8 radix-4 butterflies).
template <typename T, bool BWD>
SB_HD void dft16(cx<T>* v) {
  const T h = T(0.70710678118654752440084436210485);   // cos(pi/4)
  const T c1 = T(0.92387953251128675612818318939679);  // cos(pi/8)
  const T s1 = T(0.38268343236508977172845998403040);  // sin(pi/8)
  // A[n0][k1] = sum_n1 x[n0 + 4 n1] w4^(n1 k1), stored at v[n0 + 4 k1]
#pragma unroll
  for (int n0 = 0; n0 < 4; ++n0) dft4<T, BWD>(v[n0], v[n0 + 4], v[n0 + 8], v[n0 + 12]);
  // A[n0][k1] *= w16^(n0 k1)
  v[1 + 4 * 1] = mul_w<BWD, T>(v[1 + 4 * 1], c1, s1);    // e = 1
  v[2 + 4 * 1] = mul_w<BWD, T>(v[2 + 4 * 1], h, h);      // e = 2
  v[3 + 4 * 1] = mul_w<BWD, T>(v[3 + 4 * 1], s1, c1);    // e = 3
  v[1 + 4 * 2] = mul_w<BWD, T>(v[1 + 4 * 2], h, h);      // e = 2
  v[2 + 4 * 2] = mul_si<BWD, T>(v[2 + 4 * 2]);           // e = 4
  v[3 + 4 * 2] = mul_w<BWD, T>(v[3 + 4 * 2], -h, h);     // e = 6
  v[1 + 4 * 3] = mul_w<BWD, T>(v[1 + 4 * 3], s1, c1);    // e = 3
  v[2 + 4 * 3] = mul_w<BWD, T>(v[2 + 4 * 3], -h, h);     // e = 6
  v[3 + 4 * 3] = mul_w<BWD, T>(v[3 + 4 * 3], -c1, -s1);  // e = 9
  // X[k1 + 4 k2] = sum_n0 A[n0][k1] w4^(n0 k2): now v[4 k1 + k2] = X[k1 + 4 k2]
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4<T, BWD>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // natural order (register renaming only: every index is a compile-time constant)
  cx<T> t;
#define SB_SWAP16(a, b) t = v[a]; v[a] = v[b]; v[b] = t;
  SB_SWAP16(1, 4) SB_SWAP16(2, 8) SB_SWAP16(3, 12) SB_SWAP16(6, 9) SB_SWAP16(7, 13) SB_SWAP16(11, 14)
#undef SB_SWAP16
}

// cos / sin of 2*pi*q/32, q = 0..15 (radix-2 step across the lane pair of the N = 512 plan)
template <typename T>
SB_HD void w32_const(int q, T& c, T& s) {
  constexpr double C[16] = {1.0,
                            0.98078528040323044912618223613424,
                            0.92387953251128675612818318939679,
                            0.83146961230254523707878837761791,
                            0.70710678118654752440084436210485,
                            0.55557023301960222474283081394853,
                            0.38268343236508977172845998403040,
                            0.19509032201612826784828486847702,
                            0.0,
                            -0.19509032201612826784828486847702,
                            -0.38268343236508977172845998403040,
                            -0.55557023301960222474283081394853,
                            -0.70710678118654752440084436210485,
                            -0.83146961230254523707878837761791,
                            -0.92387953251128675612818318939679,
                            -0.98078528040323044912618223613424};
  c = T(C[q]);
  s = T(C[q >= 8 ? q - 8 : 8 - q]);  // sin(t) = cos(t - pi/2) = cos(2 pi |q - 8| / 32)
}

template <typename T, int N>
struct WPlan;

// --------------------------------------------------------------------------------------------
// N = 512
// --------------------------------------------------------------------------------------------
template <typename T>
struct WPlan<T, 512> {
  static constexpr int N = 512;
  static constexpr int LANES = 32;          // lanes per transform
  static constexpr int PER_WARP = 1;        // transforms per warp
  static constexpr int TW = 15 * 32;        // stage-B twiddle table entries: tw[(r-1)*32 + L] = w512^(r L)
  // slot of element n'' of the exchange, XOR-swizzled so that the stage-A writes (32 j + k', lanes
  // along j) and the stage-B reads (L + 32 r, lanes along L) are both bank-conflict free
  static SB_HD int slot(int n) {
    return sizeof(T) == 8 ? (n ^ ((n >> 5) & 7)) : (n ^ ((n >> 5) & 15));
  }
  // exchange write: register i of lane L after stage A
  static SB_HD int xw(int L, int i) {
    const int j = L & 15, h = L >> 4;
    const int kp = (i & 7) + 8 * h + ((i & 8) << 1);
    return slot(32 * j + kp);
  }
  // exchange read: input r of stage B
  static SB_HD int xr(int L, int r) { return slot(L + 32 * r); }

  // Stage A, part 1 (inside the lane): DFT16 over the lane's 16 inputs. Lanes of the upper half warp
  // (h = 1) produce their outputs rotated by 8 registers (odd inputs negated) and multiplied by
  // w32^q', so that afterwards EVERY lane keeps registers 0..7 and sends registers 8..15.
  template <bool BWD>
  static SB_HD void stage_a_local(cx<T>* v, int L) {
    const int h = L >> 4;
    const unsigned mask = (unsigned)h << 31;
#pragma unroll
    for (int m = 1; m < 16; m += 2) v[m] = flip_sign<T>(v[m], mask);
    dft16<T, BWD>(v);
    if (h) {
      // register q holds E1[q'] with q' = (q + 8) & 15: multiply by w32^(q')
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int qp = (q + 8) & 15;
        if (qp == 0) continue;
        T c, s;
        w32_const<T>(qp, c, s);
        v[q] = qp == 8 ? mul_si<BWD, T>(v[q]) : mul_w<BWD, T>(v[q], c, s);
      }
    }
  }
  // Stage A, part 2: `recv` = register 8 + i of the partner lane L ^ 16. In place: register i gets
  // k' = i + 8 h, register 8 + i gets k' = i + 8 h + 16.
  static SB_HD void stage_a_combine(cx<T>& keep, cx<T>& hi, cx<T> recv, int L) {
    const cx<T> lo = keep + recv;
    const cx<T> d = keep - recv;  // h = 0: E0 - T;  h = 1: T - E0 = -(E0 - T)
    keep = lo;
    hi = flip_sign<T>(d, (unsigned)(L >> 4) << 31);
  }
  // Stage B: twiddles + DFT16; afterwards v[q] = X[L + 32 q]. tw: forward table (conjugated for BWD)
  template <bool BWD, typename TWP>
  static SB_HD void stage_b(cx<T>* v, int L, TWP tw) {
#pragma unroll
    for (int r = 1; r < 16; ++r) {
      const cx<T> w = tw[(r - 1) * 32 + L];
      v[r] = v[r] * (BWD ? conj(w) : w);
    }
    dft16<T, BWD>(v);
  }
};

// --------------------------------------------------------------------------------------------
// N = 256: half a warp per transform
// --------------------------------------------------------------------------------------------
template <typename T>
struct WPlan<T, 256> {
  static constexpr int N = 256;
  static constexpr int LANES = 16;
  static constexpr int PER_WARP = 2;
  static constexpr int TW = 15 * 16;  // tw[(r-1)*16 + j] = w256^(r j)
  // n'' in 0..255 of transform t (0/1) of the warp: writes 16 j + q (lanes along j), reads j + 16 r
  static SB_HD int slot(int n) {
    return sizeof(T) == 8 ? (n ^ ((n >> 4) & 7)) : (n ^ ((n >> 4) & 15));
  }
  static SB_HD int xw(int L, int i) { return ((L >> 4) << 8) + slot(16 * (L & 15) + i); }
  static SB_HD int xr(int L, int r) { return ((L >> 4) << 8) + slot((L & 15) + 16 * r); }
  template <bool BWD>
  static SB_HD void stage_a_local(cx<T>* v, int) {
    dft16<T, BWD>(v);
  }
  template <bool BWD, typename TWP>
  static SB_HD void stage_b(cx<T>* v, int L, TWP tw) {
    const int j = L & 15;
#pragma unroll
    for (int r = 1; r < 16; ++r) {
      const cx<T> w = tw[(r - 1) * 16 + j];
      v[r] = v[r] * (BWD ? conj(w) : w);
    }
    dft16<T, BWD>(v);
  }
};

// Lane twiddles kept in registers for the whole kernel: w^1, w^2, w^4, w^8 of w = w_N^(lane index),
// forward sign. The other eleven powers of a radix-16 stage are products of these (11 complex
// multiplications on the fp64 / fp32 pipe instead of 15 table reads through the L1 data pipe, which is
// the busier unit of the stage kernels).
template <typename T>
struct LaneTw {
  cx<T> w1, w2, w4, w8;
};

// v[r] *= w^r (r = 1..15), conjugated for the backward transform, then the 16-point DFT.
template <typename T, bool BWD>
SB_HD void twiddle_dft16(cx<T>* v, const LaneTw<T>& t) {
  const cx<T> w1 = BWD ? conj(t.w1) : t.w1;
  const cx<T> w2 = BWD ? conj(t.w2) : t.w2;
  const cx<T> w4 = BWD ? conj(t.w4) : t.w4;
  const cx<T> w8 = BWD ? conj(t.w8) : t.w8;
  const cx<T> w3 = w2 * w1;
  v[1] = v[1] * w1;
  v[2] = v[2] * w2;
  v[3] = v[3] * w3;
  v[4] = v[4] * w4;
  const cx<T> w5 = w4 * w1, w6 = w4 * w2, w7 = w4 * w3;
  v[5] = v[5] * w5;
  v[6] = v[6] * w6;
  v[7] = v[7] * w7;
  v[8] = v[8] * w8;
  v[9] = v[9] * (w8 * w1);
  v[10] = v[10] * (w8 * w2);
  v[11] = v[11] * (w8 * w3);
  v[12] = v[12] * (w8 * w4);
  v[13] = v[13] * (w8 * w5);
  v[14] = v[14] * (w8 * w6);
  v[15] = v[15] * (w8 * w7);
  dft16<T, BWD>(v);
}

// host side: the four lane twiddles w_n^(2^i * lane), i = 0..3, lane = 0..lanes-1, forward sign
template <typename T>
inline void wfft_lane_twiddles(int n, int lanes, cx<T>* out /* [4][lanes] */) {
  const long double pi2 = 6.283185307179586476925286766559005768L;
  for (int i = 0; i < 4; ++i)
    for (int L = 0; L < lanes; ++L) {
      const long double a = -pi2 * (long double)(((1 << i) * L) % n) / (long double)n;
      out[i * lanes + L] = mk<T>((T)cosl(a), (T)sinl(a));
    }
}

// --------------------------------------------------------------------------------------------
// Tile geometry of the stage kernels built on the length-512 plan (wfft_kernels.cuh): 16-byte units, i.e. one
// complex double or two complex floats (f2 above).
// Host + device so that tests/emu can check the address algebra against WPlan<T, 512>::xw / xr.
// --------------------------------------------------------------------------------------------
constexpr int kWN = 512;    // transform length of this kernel family
constexpr int kWWarps = 8;  // transforms per tile
// Sub-tiles. The 8 warps of a CTA form 8 / W independent groups of W warps; a group owns W adjacent columns of
// the CTA's 8-column tile and its own sub-tile buffer [512 rows][W x 16-byte chunks] (TMA swizzle of the row
// width: 128B / 64B / 32B), synchronises on its own named barrier and never waits for the other groups: with
// W = 2 (pairs) four groups per CTA drift through load / fp64 / exchange / store phases independently, which is
// what lets the SM overlap them (one 8-warp tile per CTA ran in lock step: ncu barrier stalls 29 % of all samples).
template <int W>
struct WGeom {
  static_assert(W == 2 || W == 4 || W == 8, "columns per group");
  static constexpr int kGroups = kWWarps / W;
  static constexpr int kGroupThreads = W * 32;
  static constexpr unsigned kRowBytes = 16u * W;
  static constexpr size_t kSubBytes = (size_t)kWN * kRowBytes;
  static constexpr int kLog2RowsPer128 = W == 8 ? 0 : (W == 4 ? 1 : 2);
  // chunk permutation of row s: chunk c sits at c ^ fold(s)
  static SB_HD constexpr unsigned fold(unsigned s) { return (s >> kLog2RowsPer128) & (W - 1); }
  // xor pattern of the low three slot bits q
  static SB_HD constexpr unsigned pat(unsigned q) { return (q * kRowBytes) | (fold(q) << 4); }
};
// Byte offsets inside the group's sub-tile of everything lane L of its warp wl touches, in a form that costs ONE
// xor per access (RB = row bytes = 16 W):
//   natural element n = L + 32 m of column wl        : nat + m * 32 RB
//   exchange write, register i (slot xw(L, i))       : (xwBase ^ pat(i & 7)) + (i >> 3) * 16 RB
//   exchange read, input r (slot xr(L, r))           : (nat ^ pat(r & 7)) + r * 32 RB
// (slot s of column wl lives at byte s * RB + ((wl ^ fold(s)) << 4); the slots of WPlan<T, 512> differ from
// lane-constant bases only in their low three bits q, which enter the address as the xor pattern pat(q).)
struct WAddr {
  unsigned nat, xwBase;
};
template <int W>
SB_HD WAddr w_addr(int wl, int L) {
  using G = WGeom<W>;
  WAddr a;
  a.nat = ((unsigned)L * G::kRowBytes) | ((unsigned)(wl ^ G::fold(L & 7)) << 4);
  const unsigned j = L & 15, h = L >> 4, q = j & 7;
  a.xwBase = ((32u * j + 8u * h) * G::kRowBytes) | (q * G::kRowBytes) | ((unsigned)(wl ^ G::fold(q)) << 4);
  return a;
}
// byte offsets of the three access patterns (see above)
template <int W>
SB_HD unsigned w_xw_off(const WAddr& ad, int i) {
  return (ad.xwBase ^ WGeom<W>::pat(i & 7)) + (unsigned)(i >> 3) * 16u * WGeom<W>::kRowBytes;
}
template <int W>
SB_HD unsigned w_xr_off(const WAddr& ad, int r) {
  return (ad.nat ^ WGeom<W>::pat(r & 7)) + (unsigned)r * 32u * WGeom<W>::kRowBytes;
}
template <int W>
SB_HD unsigned w_nat_off(const WAddr& ad, int m) {
  return ad.nat + (unsigned)m * 32u * WGeom<W>::kRowBytes;
}

// number of table entries / fill (host side, long double roots like make_fast_twiddles)
inline int wfft_tw_size(int n) { return n == 512 ? 15 * 32 : (n == 256 ? 15 * 16 : 0); }
inline bool wfft_length(int n) { return n == 512 || n == 256; }

#if SB_ON_GPU
// complex shuffle
template <typename T>
SB_DEV cx<T> shfl_xor_cx(cx<T> v, int mask) {
  return mk<T>(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}
template <>
SB_DEV cx<f2> shfl_xor_cx<f2>(cx<f2> v, int mask) {
  cx<f2> o;
  o.x.r = __shfl_xor_sync(0xffffffffu, v.x.r, mask);
  o.y.r = __shfl_xor_sync(0xffffffffu, v.y.r, mask);
  return o;
}

// The whole transform of one warp: v[m] = x[L + T m] in, v[q] = X[L + T q] out. `X` = the warp's private
// exchange region (N * PER_WARP elements), `tw` = stage-B table (shared or global memory).
template <typename T, int N, bool BWD, typename TWP>
SB_DEV void warp_fft(cx<T>* v, cx<T>* X, TWP tw, int L) {
  using P = WPlan<T, N>;
  P::template stage_a_local<BWD>(v, L);
  if constexpr (N == 512) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const cx<T> recv = shfl_xor_cx<T>(v[8 + i], 16);
      P::stage_a_combine(v[i], v[8 + i], recv, L);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) X[P::xw(L, i)] = v[i];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = X[P::xr(L, r)];
  __syncwarp();  // the region may be rewritten by the warp's next transform
  P::template stage_b<BWD>(v, L, tw);
}
#endif

}  // namespace sb
