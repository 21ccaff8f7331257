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

}  // namespace

extern "C" {
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
