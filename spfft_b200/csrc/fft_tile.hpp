// fft_tile.hpp -- batched shared-memory Stockham FFT over a "lane-interleaved" tile.
//
// Tile layout: element n of sequence `lane` lives at  n*V + lane  (V = lanes per tile, a power of
// two chosen so that V complex numbers are 128 bytes when the length allows). Consecutive threads
// own consecutive lanes, so every shared-memory access of a quarter warp touches one 128-byte row
// -> conflict free for any radix or stride, and a tile row is exactly one coalesced global segment
// of the plane-major stick buffer / the xy-plane buffer (see DESIGN.md "Data layout").
// With SWZ the lane is XOR-swizzled by the low bits of n, which additionally makes the
// row <-> lane transposition done by the x-stage loads/stores conflict free.
//
// This replaces the cuFFT plans the reference builds in src/fft/transform_1d_gpu.hpp:52-141 and
// src/fft/transform_2d_gpu.hpp:51-140 (unnormalised DFT, sign + backward / - forward,
// docs/source/details.rst:6-13).
#pragma once
#include "cx.hpp"

namespace sb {

constexpr int kMaxPasses = 24;

struct RadixPlan {
  int n;                   // transform length (product of radix[0..numPasses))
  int numPasses;
  int radix[kMaxPasses];
};

template <bool SWZ>
SB_HD int at(int n, int lane, int log2V) {
  const int V = 1 << log2V;
  return SWZ ? ((n << log2V) + (lane ^ (n & (V - 1)))) : ((n << log2V) + lane);
}

// forward root table: tw[k] = exp(-2*pi*i*k/N); the backward transform uses the conjugate
template <bool BWD, typename T>
SB_DEV cx<T> ldtw(const cx<T>* __restrict__ tw, int idx) {
  const cx<T> w = tw[idx];
  return BWD ? conj(w) : w;
}

// ---------------------------------------------------------------------------------------------
// in-register DFTs of length R:  v[q] <- sum_r v[r] * exp(s*2*pi*i*q*r/R),  s = +1 (BWD) / -1
// ---------------------------------------------------------------------------------------------
template <typename T, bool BWD>
SB_DEV void dft2(cx<T>& a, cx<T>& b) {
  const cx<T> t = a - b;
  a = a + b;
  b = t;
}

template <typename T, bool BWD>
SB_DEV void dft4(cx<T>& v0, cx<T>& v1, cx<T>& v2, cx<T>& v3) {
  const cx<T> a0 = v0 + v2, a1 = v0 - v2, a2 = v1 + v3;
  const cx<T> a3 = mul_si<BWD>(v1 - v3);
  v0 = a0 + a2;
  v1 = a1 + a3;
  v2 = a0 - a2;
  v3 = a1 - a3;
}

template <typename T, bool BWD, int R>
struct Butterfly;

template <typename T, bool BWD>
struct Butterfly<T, BWD, 2> {
  static SB_DEV void run(cx<T>* v) { dft2<T, BWD>(v[0], v[1]); }
};

template <typename T, bool BWD>
struct Butterfly<T, BWD, 3> {
  static SB_DEV void run(cx<T>* v) {
    const T h = T(0.86602540378443864676372317075294);  // sin(pi/3)
    const cx<T> t = v[1] + v[2];
    const cx<T> u = mul_si<BWD>(h * (v[1] - v[2]));
    const cx<T> m = v[0] - T(0.5) * t;
    v[0] = v[0] + t;
    v[1] = m + u;
    v[2] = m - u;
  }
};

template <typename T, bool BWD>
struct Butterfly<T, BWD, 4> {
  static SB_DEV void run(cx<T>* v) { dft4<T, BWD>(v[0], v[1], v[2], v[3]); }
};

template <typename T, bool BWD>
struct Butterfly<T, BWD, 5> {
  static SB_DEV void run(cx<T>* v) {
    const T c1 = T(0.30901699437494742410229341718282);   // cos(2pi/5)
    const T c2 = T(-0.80901699437494742410229341718282);  // cos(4pi/5)
    const T s1 = T(0.95105651629515357211643933337938);   // sin(2pi/5)
    const T s2 = T(0.58778525229247312916870595463907);   // sin(4pi/5)
    const cx<T> t1 = v[1] + v[4], t2 = v[2] + v[3], t3 = v[1] - v[4], t4 = v[2] - v[3];
    const cx<T> m1 = v[0] + c1 * t1 + c2 * t2;
    const cx<T> m2 = v[0] + c2 * t1 + c1 * t2;
    const cx<T> n1 = mul_si<BWD>(s1 * t3 + s2 * t4);
    const cx<T> n2 = mul_si<BWD>(s2 * t3 - s1 * t4);
    v[0] = v[0] + t1 + t2;
    v[1] = m1 + n1;
    v[4] = m1 - n1;
    v[2] = m2 + n2;
    v[3] = m2 - n2;
  }
};

template <typename T, bool BWD>
struct Butterfly<T, BWD, 8> {
  static SB_DEV void run(cx<T>* v) {
    const T h = T(0.70710678118654752440084436210485);  // sqrt(1/2)
    // split into even/odd output halves (decimation in frequency)
    cx<T> a0 = v[0] + v[4], b0 = v[0] - v[4];
    cx<T> a1 = v[1] + v[5], b1 = v[1] - v[5];
    cx<T> a2 = v[2] + v[6], b2 = v[2] - v[6];
    cx<T> a3 = v[3] + v[7], b3 = v[3] - v[7];
    // b_k *= w8^k, w8 = exp(s*2*pi*i/8) = h*(1 + s*i)
    {
      const cx<T> ib1 = mul_si<BWD>(b1);
      b1 = h * (b1 + ib1);          // (1 + s i)/sqrt2
      b2 = mul_si<BWD>(b2);         // s i
      const cx<T> ib3 = mul_si<BWD>(b3);
      b3 = h * (ib3 - b3);          // (-1 + s i)/sqrt2
    }
    dft4<T, BWD>(a0, a1, a2, a3);
    dft4<T, BWD>(b0, b1, b2, b3);
    v[0] = a0;
    v[2] = a1;
    v[4] = a2;
    v[6] = a3;
    v[1] = b0;
    v[3] = b1;
    v[5] = b2;
    v[7] = b3;
  }
};

// ---------------------------------------------------------------------------------------------
// One Stockham pass (autosort, out of place src -> dst), radix R held in registers.
// Butterfly j (0 <= j < N/R), k = j mod ns:
//   reads  src[j + r*N/R] * w_N^{r*k*N/(ns*R)},   writes dst[(j-k)*R + k + q*ns]
// ---------------------------------------------------------------------------------------------
template <typename T, bool BWD, bool SWZ, int R>
SB_DEV void pass_radix(const cx<T>* src, cx<T>* dst, int N, int log2V, int ns,
                       const cx<T>* __restrict__ tw, int tid, int nthr) {
  const int nb = N / R;
  const int twStep = N / (ns * R);
  const int V = 1 << log2V;
  const int items = nb << log2V;
  for (int item = tid; item < items; item += nthr) {
    const int lane = item & (V - 1);
    const int j = item >> log2V;
    const int k = j % ns;
    cx<T> v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = src[at<SWZ>(j + r * nb, lane, log2V)];
    if (ns > 1) {
      const int e = k * twStep;
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = v[r] * ldtw<BWD>(tw, r * e);
    }
    Butterfly<T, BWD, R>::run(v);
    const int o = (j - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) dst[at<SWZ>(o + q * ns, lane, log2V)] = v[q];
  }
}

// Any radix (7, 11, 13, large primes ...): one output per work item, O(N*R) per pass.
template <typename T, bool BWD, bool SWZ>
SB_DEV void pass_generic(const cx<T>* src, cx<T>* dst, int N, int log2V, int ns, int R,
                         const cx<T>* __restrict__ tw, int tid, int nthr) {
  const int nb = N / R;
  const int twStep = N / (ns * R);
  const int rootStep = N / R;
  const int V = 1 << log2V;
  const int items = N << log2V;
  for (int item = tid; item < items; item += nthr) {
    const int lane = item & (V - 1);
    const int rest = item >> log2V;
    const int q = rest % R;
    const int j = rest / R;
    const int k = j % ns;
    const int base = k * twStep + q * rootStep;  // < N
    cx<T> acc = src[at<SWZ>(j, lane, log2V)];
    int idx = 0;
    for (int r = 1; r < R; ++r) {
      idx += base;
      if (idx >= N) idx -= N;
      acc = acc + src[at<SWZ>(j + r * nb, lane, log2V)] * ldtw<BWD>(tw, idx);
    }
    dst[at<SWZ>((j - k) * R + k + q * ns, lane, log2V)] = acc;
  }
}

// Full transform of all lanes of a tile. `a` holds the input (complete and synchronised),
// `b` is scratch of the same size. Returns the buffer that holds the result (synchronised).
template <typename T, bool BWD, bool SWZ>
SB_DEV cx<T>* tile_fft(cx<T>* a, cx<T>* b, const RadixPlan& rp, int log2V,
                       const cx<T>* __restrict__ tw, Ctx ctx) {
  (void)ctx;
  const int N = rp.n;
  int ns = 1;
  for (int p = 0; p < rp.numPasses; ++p) {
    const int R = rp.radix[p];
    SB_PHASE_BEGIN
    switch (R) {
      case 2: pass_radix<T, BWD, SWZ, 2>(a, b, N, log2V, ns, tw, tid, nthr); break;
      case 3: pass_radix<T, BWD, SWZ, 3>(a, b, N, log2V, ns, tw, tid, nthr); break;
      case 4: pass_radix<T, BWD, SWZ, 4>(a, b, N, log2V, ns, tw, tid, nthr); break;
      case 5: pass_radix<T, BWD, SWZ, 5>(a, b, N, log2V, ns, tw, tid, nthr); break;
      case 8: pass_radix<T, BWD, SWZ, 8>(a, b, N, log2V, ns, tw, tid, nthr); break;
      default: pass_generic<T, BWD, SWZ>(a, b, N, log2V, ns, R, tw, tid, nthr); break;
    }
    SB_PHASE_END
    ns *= R;
    cx<T>* t = a;
    a = b;
    b = t;
  }
  return a;
}

}  // namespace sb
