// emu_wfft.cpp -- CPU emulation of the warp-autonomous FFT (spfft_b200/csrc/wfft.hpp): the SAME
// arithmetic bodies and index functions as the device code, with the warp's lanes as a loop and the
// shuffle / shared-memory exchange as plain arrays. Test infrastructure only (tests/test_emu.py).
#include <cmath>
#include <complex>
#include <vector>

#include "wfft.hpp"

namespace {

template <typename T, int N>
std::vector<sb::cx<T>> make_tw() {
  using P = sb::WPlan<T, N>;
  std::vector<sb::cx<T>> tw(P::TW);
  const long double pi2 = 6.283185307179586476925286766559005768L;
  for (int r = 1; r < 16; ++r)
    for (int L = 0; L < P::LANES; ++L) {
      const long double a = -pi2 * (long double)((r * L) % N) / (long double)N;
      tw[(r - 1) * P::LANES + L] = sb::mk<T>((T)cosl(a), (T)sinl(a));
    }
  return tw;
}

// one warp: `count` = PER_WARP transforms, x / out = [count][N] interleaved
template <typename T, int N, bool BWD>
int run_warp(const T* x, T* out) {
  using P = sb::WPlan<T, N>;
  const auto tw = make_tw<T, N>();
  std::vector<sb::cx<T>> v(32 * 16), X(N * P::PER_WARP);
  std::vector<int> written(N * P::PER_WARP, 0);
  const sb::cx<T>* in = reinterpret_cast<const sb::cx<T>*>(x);
  for (int L = 0; L < 32; ++L) {
    const int t = L / P::LANES, l = L % P::LANES;
    for (int m = 0; m < 16; ++m) v[L * 16 + m] = in[t * N + l + P::LANES * m];
    P::template stage_a_local<BWD>(&v[L * 16], L);
  }
  if (N == 512) {
    std::vector<sb::cx<T>> sent(32 * 8);
    for (int L = 0; L < 32; ++L)
      for (int i = 0; i < 8; ++i) sent[L * 8 + i] = v[L * 16 + 8 + i];
    for (int L = 0; L < 32; ++L)
      for (int i = 0; i < 8; ++i) {
        if constexpr (N == 512) P::stage_a_combine(v[L * 16 + i], v[L * 16 + 8 + i], sent[(L ^ 16) * 8 + i], L);
      }
  }
  for (int L = 0; L < 32; ++L)
    for (int i = 0; i < 16; ++i) {
      const int s = P::xw(L, i);
      if (s < 0 || s >= N * P::PER_WARP || written[s]) return 1;  // not a bijection
      written[s] = 1;
      X[s] = v[L * 16 + i];
    }
  for (int L = 0; L < 32; ++L) {
    for (int r = 0; r < 16; ++r) v[L * 16 + r] = X[P::xr(L, r)];
    P::template stage_b<BWD>(&v[L * 16], L, tw.data());
  }
  sb::cx<T>* o = reinterpret_cast<sb::cx<T>*>(out);
  for (int L = 0; L < 32; ++L) {
    const int t = L / P::LANES, l = L % P::LANES;
    for (int q = 0; q < 16; ++q) o[t * N + l + P::LANES * q] = v[L * 16 + q];
  }
  return 0;
}

// bank conflicts of the exchange: worst number of distinct 128-byte-bank rows hit by one wavefront
// (a quarter warp for 16-byte elements, half a warp for 8-byte elements), over all instructions
template <typename T, int N>
int worst_conflict() {
  using P = sb::WPlan<T, N>;
  const int per = 128 / (int)sizeof(sb::cx<T>);  // lanes per wavefront = slots per 128 bytes
  int worst = 1;
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < 16; ++i)
      for (int g = 0; g < 32 / per; ++g) {
        std::vector<int> cnt(per, 0);
        for (int l = 0; l < per; ++l) {
          const int L = g * per + l;
          const int s = pass ? P::xr(L, i) : P::xw(L, i);
          worst = std::max(worst, ++cnt[s % per]);
        }
      }
  return worst;
}

// The transform as the stage kernels run it (wfft_kernels.cuh): exchange through the column of a [512][W] sub-tile
// addressed in xor form (w_addr / w_xw_off / w_xr_off / w_nat_off), stage B with the four lane twiddles
// w, w^2, w^4, w^8 and derived powers (twiddle_dft16). One warp = column `wl` of the sub-tile; the other columns
// are poisoned and must stay untouched. Returns 0, or a code for: 1 overlapping writes, 2 foreign column touched.
inline bool operator!=(sb::f2 a, sb::f2 b) { return sb::f2_lo(a) != sb::f2_lo(b) || sb::f2_hi(a) != sb::f2_hi(b); }

template <typename T, int W, bool BWD>
int run_warp_tile(int wl, const sb::cx<T>* in, sb::cx<T>* o) {
  using P = sb::WPlan<T, 512>;
  using G = sb::WGeom<W>;
  constexpr int N = 512;
  std::vector<sb::cx<T>> tw4(4 * 32);
  sb::wfft_lane_twiddles<T>(N, 32, tw4.data());
  const size_t elems = G::kSubBytes / sizeof(sb::cx<T>);
  std::vector<sb::cx<T>> S(elems, sb::mk<T>(T(7e30), T(-7e30)));
  std::vector<int> owner(elems, -1);
  auto at = [&](unsigned byteOff) -> size_t { return byteOff / sizeof(sb::cx<T>); };
  std::vector<sb::cx<T>> v(32 * 16);
  // tile side in: the natural-order column, as a bulk tensor load would have placed it
  for (int L = 0; L < 32; ++L) {
    const sb::WAddr ad = sb::w_addr<W>(wl, L);
    for (int m = 0; m < 16; ++m) {
      const size_t e = at(sb::w_nat_off<W>(ad, m));
      const int n = L + 32 * m;
      // element n of column wl must sit at chunk wl ^ fold(n) of row n (the TMA swizzle of the row width)
      if (e != (size_t)n * W + (wl ^ G::fold(n))) return 3;
      S[e] = in[n];
    }
  }
  for (int L = 0; L < 32; ++L) {
    const sb::WAddr ad = sb::w_addr<W>(wl, L);
    for (int m = 0; m < 16; ++m) v[L * 16 + m] = S[at(sb::w_nat_off<W>(ad, m))];
    P::template stage_a_local<BWD>(&v[L * 16], L);
  }
  std::vector<sb::cx<T>> sent(32 * 8);
  for (int L = 0; L < 32; ++L)
    for (int i = 0; i < 8; ++i) sent[L * 8 + i] = v[L * 16 + 8 + i];
  for (int L = 0; L < 32; ++L)
    for (int i = 0; i < 8; ++i) P::stage_a_combine(v[L * 16 + i], v[L * 16 + 8 + i], sent[(L ^ 16) * 8 + i], L);
  std::vector<int> written(elems, 0);
  for (int L = 0; L < 32; ++L) {
    const sb::WAddr ad = sb::w_addr<W>(wl, L);
    for (int i = 0; i < 16; ++i) {
      const size_t e = at(sb::w_xw_off<W>(ad, i));
      const int s = P::xw(L, i);
      if (e != (size_t)s * W + (wl ^ G::fold(s))) return 4;  // xor form == slot of the plan in column wl
      if (written[e]) return 1;
      written[e] = 1;
      S[e] = v[L * 16 + i];
    }
  }
  for (int L = 0; L < 32; ++L) {
    const sb::WAddr ad = sb::w_addr<W>(wl, L);
    for (int r = 0; r < 16; ++r) {
      const size_t e = at(sb::w_xr_off<W>(ad, r));
      const int s = P::xr(L, r);
      if (e != (size_t)s * W + (wl ^ G::fold(s))) return 5;
      v[L * 16 + r] = S[e];
    }
    sb::LaneTw<T> t;
    t.w1 = tw4[L];
    t.w2 = tw4[32 + L];
    t.w4 = tw4[64 + L];
    t.w8 = tw4[96 + L];
    sb::twiddle_dft16<T, BWD>(&v[L * 16], t);
  }
  // every element outside column wl is still poisoned
  for (size_t e = 0; e < elems; ++e) {
    const size_t row = e / W, chunk = e % W;
    const bool mine = (chunk ^ G::fold((unsigned)row)) == (size_t)wl;
    if (!mine && S[e].x != T(7e30)) return 2;
  }
  for (int L = 0; L < 32; ++L)
    for (int q = 0; q < 16; ++q) o[L + 32 * q] = v[L * 16 + q];
  return 0;
}

// worst number of lanes of a quarter warp (8 lanes: one 128-byte wavefront of a 128-bit access) that fall into
// the same 16-byte bank group, over the three access patterns of the stage kernels
template <int W>
int worst_tile_conflict(int wl) {
  int worst = 1;
  for (int pass = 0; pass < 3; ++pass)
    for (int i = 0; i < 16; ++i)
      for (int g = 0; g < 4; ++g) {
        int cnt[8] = {0};
        for (int l = 0; l < 8; ++l) {
          const int L = g * 8 + l;
          const sb::WAddr ad = sb::w_addr<W>(wl, L);
          const unsigned off = pass == 0 ? sb::w_xw_off<W>(ad, i) : (pass == 1 ? sb::w_xr_off<W>(ad, i) : sb::w_nat_off<W>(ad, i));
          worst = std::max(worst, ++cnt[(off >> 4) & 7]);
        }
      }
  return worst;
}

}  // namespace

extern "C" {
// the product form: column wl of a [512][W] sub-tile, xor-form addresses, derived twiddles (double precision)
int emu_wfft_tile_f64(int W, int wl, int backward, const double* x, double* out) {
  const sb::cx<double>* in = reinterpret_cast<const sb::cx<double>*>(x);
  sb::cx<double>* o = reinterpret_cast<sb::cx<double>*>(out);
#define SB_TILE(WW) \
  if (W == WW) return backward ? run_warp_tile<double, WW, true>(wl, in, o) : run_warp_tile<double, WW, false>(wl, in, o);
  SB_TILE(2) SB_TILE(4) SB_TILE(8)
#undef SB_TILE
  return -1;
}
// the same with the packed single-precision pair type (sb::f2): x / out = [2][512] interleaved complex floats, the
// two transforms a warp of the single-precision stage kernels runs at once; also checks the memory <-> packed
// permutation of a unit (unit_transpose / unit_pack / unit_lo / unit_hi)
int emu_wfft_tile_f32x2(int W, int wl, int backward, const float* x, float* out) {
  const sb::cx<float>* xc = reinterpret_cast<const sb::cx<float>*>(x);
  sb::cx<float>* oc = reinterpret_cast<sb::cx<float>*>(out);
  std::vector<sb::cx<sb::f2>> in(512), o(512);
  for (int n = 0; n < 512; ++n) {
    in[n] = sb::unit_pack(xc[n], xc[512 + n]);
    // a unit as it lies in memory: (re0, im0, re1, im1)
    const sb::cx<sb::f2> mem = sb::mk<sb::f2>(sb::f2_make(xc[n].x, xc[n].y), sb::f2_make(xc[512 + n].x, xc[512 + n].y));
    const sb::cx<sb::f2> t = sb::unit_transpose<sb::f2>(mem);
    if (t.x != in[n].x || t.y != in[n].y) return 6;
    const sb::cx<sb::f2> back = sb::unit_transpose<sb::f2>(t);
    if (back.x != mem.x || back.y != mem.y) return 6;
  }
  int err = -1;
#define SB_TILE(WW) \
  if (W == WW) err = backward ? run_warp_tile<sb::f2, WW, true>(wl, in.data(), o.data()) : run_warp_tile<sb::f2, WW, false>(wl, in.data(), o.data());
  SB_TILE(2) SB_TILE(4) SB_TILE(8)
#undef SB_TILE
  if (err) return err;
  for (int n = 0; n < 512; ++n) {
    oc[n] = sb::unit_lo(o[n]);
    oc[512 + n] = sb::unit_hi(o[n]);
  }
  return 0;
}
int emu_wfft_tile_conflicts(int W, int wl) {
  if (W == 2) return worst_tile_conflict<2>(wl);
  if (W == 4) return worst_tile_conflict<4>(wl);
  if (W == 8) return worst_tile_conflict<8>(wl);
  return -1;
}
// dir: 0 forward (sign -), 1 backward (sign +). Returns 0 on success.
int emu_wfft_f64(int n, int backward, const double* x, double* out) {
  if (n == 512) return backward ? run_warp<double, 512, true>(x, out) : run_warp<double, 512, false>(x, out);
  if (n == 256) return backward ? run_warp<double, 256, true>(x, out) : run_warp<double, 256, false>(x, out);
  return 2;
}
int emu_wfft_f32(int n, int backward, const float* x, float* out) {
  if (n == 512) return backward ? run_warp<float, 512, true>(x, out) : run_warp<float, 512, false>(x, out);
  if (n == 256) return backward ? run_warp<float, 256, true>(x, out) : run_warp<float, 256, false>(x, out);
  return 2;
}
int emu_wfft_conflicts(int n, int isFloat) {
  if (n == 512) return isFloat ? worst_conflict<float, 512>() : worst_conflict<double, 512>();
  if (n == 256) return isFloat ? worst_conflict<float, 256>() : worst_conflict<double, 256>();
  return -1;
}
}
