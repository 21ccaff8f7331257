/*
 * fftw3_shim.cpp -- the arithmetic behind oracle/fftw3_shim/fftw3.h.
 *
 * TEST INFRASTRUCTURE ONLY (oracle / CPU baseline). Not part of the product library.
 *
 * The reference keeps all of its host FFT arithmetic in FFTW 3.x (unpinned version,
 * CMakeLists.txt:220 `find_package(FFTW REQUIRED)`), which is absent from /root/reference and
 * from this image. This file restates the *published* definition FFTW implements
 * (FFTW 3.3 manual, "What FFTW Really Computes"):
 *
 *     Y[k] = sum_{j=0}^{n-1} X[j] * exp(sign * 2*pi*i * j*k / n),   sign = -1 forward, +1 backward
 *
 * with the rank-1 "many" data layout (element j of transform t at in[t*idist + j*istride]).
 * r2c keeps k = 0..n/2; c2r consumes k = 0..n/2 and assumes hermitian symmetry (the imaginary
 * parts of the k=0 and, for even n, k=n/2 inputs are ignored).
 *
 * Algorithm: Stockham autosort mixed radix (4, 2, 3, 5, generic odd radix <= 31), Bluestein
 * chirp-z for sizes with a larger prime factor. Twiddles are generated in long double.
 * Validated against numpy.fft in tests/test_oracle_shim.py.
 */
#include "fftw3.h"

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace {

template <typename T>
using cpx = std::complex<T>;

inline std::vector<int> factorize(int n) {
  std::vector<int> f;
  while (n % 4 == 0) {
    f.push_back(4);
    n /= 4;
  }
  while (n % 2 == 0) {
    f.push_back(2);
    n /= 2;
  }
  for (int p = 3; (long long)p * p <= n; p += 2) {
    while (n % p == 0) {
      f.push_back(p);
      n /= p;
    }
  }
  if (n > 1) f.push_back(n);
  return f;
}

template <typename T>
struct Engine {
  int n = 0;
  int sign = -1;
  bool bluestein = false;
  std::vector<int> radices;
  std::vector<cpx<T>> roots;  // roots[k] = exp(sign*2*pi*i*k/n)

  // Bluestein data
  int m = 0;
  std::unique_ptr<Engine<T>> fwdM, bwdM;
  std::vector<cpx<T>> chirp;     // exp(sign*pi*i*k^2/n)
  std::vector<cpx<T>> chirpFft;  // FFT_m of conj-chirp kernel

  Engine(int n_, int sign_) : n(n_), sign(sign_) {
    radices = factorize(n);
    int maxPrime = 1;
    for (int r : radices) maxPrime = r > maxPrime ? r : maxPrime;
    roots.resize(n);
    const long double twoPi = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < n; ++k) {
      long double a = twoPi * (long double)k / (long double)n;
      roots[k] = cpx<T>((T)cosl(a), (T)(sign * sinl(a)));
    }
    if (maxPrime > 31) {
      bluestein = true;
      m = 1;
      while (m < 2 * n - 1) m *= 2;
      fwdM.reset(new Engine<T>(m, -1));
      bwdM.reset(new Engine<T>(m, +1));
      chirp.resize(n);
      for (int k = 0; k < n; ++k) {
        long long k2 = ((long long)k * k) % (2LL * n);
        long double a = twoPi * (long double)k2 / (2.0L * n);
        chirp[k] = cpx<T>((T)cosl(a), (T)(sign * sinl(a)));
      }
      std::vector<cpx<T>> b(m, cpx<T>(0, 0)), tmp(m);
      b[0] = std::conj(chirp[0]);
      for (int k = 1; k < n; ++k) b[k] = b[m - k] = std::conj(chirp[k]);
      fwdM->run(b.data(), tmp.data());
      chirpFft = b;
    }
  }

  // in-place on buf (length n); scratch length >= n (or 2*m for bluestein)
  void run(cpx<T>* buf, cpx<T>* scratch) const {
    if (n <= 1) return;
    if (bluestein) {
      cpx<T>* a = scratch;
      cpx<T>* s2 = scratch + m;
      for (int k = 0; k < n; ++k) a[k] = buf[k] * chirp[k];
      for (int k = n; k < m; ++k) a[k] = cpx<T>(0, 0);
      fwdM->run(a, s2);
      for (int k = 0; k < m; ++k) a[k] *= chirpFft[k];
      bwdM->run(a, s2);
      const T inv = (T)1 / (T)m;
      for (int k = 0; k < n; ++k) buf[k] = a[k] * chirp[k] * inv;
      return;
    }
    cpx<T>* src = buf;
    cpx<T>* dst = scratch;
    int ns = 1;
    for (int R : radices) {
      pass(src, dst, ns, R);
      ns *= R;
      std::swap(src, dst);
    }
    if (src != buf) std::memcpy((void*)buf, (void*)src, sizeof(cpx<T>) * n);
  }

  size_t scratch_size() const { return bluestein ? (size_t)2 * m + (size_t)m : (size_t)n; }

  void pass(const cpx<T>* src, cpx<T>* dst, int ns, int R) const {
    const int nr = n / R;            // butterflies
    const int blocks = nr / ns;      // outer blocks
    const int twStep = n / (ns * R); // root index step per (r*k)
    if (R == 4) {
      for (int jb = 0; jb < blocks; ++jb) {
        for (int k = 0; k < ns; ++k) {
          const int j = jb * ns + k;
          cpx<T> v0 = src[j];
          cpx<T> v1 = src[j + nr] * roots[k * twStep];
          cpx<T> v2 = src[j + 2 * nr] * roots[2 * k * twStep];
          cpx<T> v3 = src[j + 3 * nr] * roots[3 * k * twStep];
          cpx<T> a0 = v0 + v2, a1 = v0 - v2, a2 = v1 + v3, a3 = v1 - v3;
          // multiply a3 by sign*i
          cpx<T> a3i = sign < 0 ? cpx<T>(a3.imag(), -a3.real()) : cpx<T>(-a3.imag(), a3.real());
          cpx<T>* o = dst + jb * ns * 4 + k;
          o[0] = a0 + a2;
          o[ns] = a1 + a3i;
          o[2 * ns] = a0 - a2;
          o[3 * ns] = a1 - a3i;
        }
      }
    } else if (R == 2) {
      for (int jb = 0; jb < blocks; ++jb) {
        for (int k = 0; k < ns; ++k) {
          const int j = jb * ns + k;
          cpx<T> v0 = src[j];
          cpx<T> v1 = src[j + nr] * roots[k * twStep];
          cpx<T>* o = dst + jb * ns * 2 + k;
          o[0] = v0 + v1;
          o[ns] = v0 - v1;
        }
      }
    } else {
      cpx<T> v[32];
      const int rootStepR = n / R;  // roots[q*rootStepR] = w_R^q
      for (int jb = 0; jb < blocks; ++jb) {
        for (int k = 0; k < ns; ++k) {
          const int j = jb * ns + k;
          for (int r = 0; r < R; ++r) v[r] = src[j + r * nr] * roots[r * k * twStep];
          cpx<T>* o = dst + jb * ns * R + k;
          for (int q = 0; q < R; ++q) {
            cpx<T> acc = v[0];
            for (int r = 1; r < R; ++r) acc += v[r] * roots[((q * r) % R) * rootStepR];
            o[q * ns] = acc;
          }
        }
      }
    }
  }
};

enum class Kind { C2C, R2C, C2R };

template <typename T>
struct Plan {
  Kind kind;
  int n, howmany, istride, idist, ostride, odist, sign;
  void* in;
  void* out;
  std::unique_ptr<Engine<T>> eng;

  void exec(void* inV, void* outV) const {
    if (n <= 0 || howmany <= 0) return;
    std::vector<cpx<T>> buf(n);
    std::vector<cpx<T>> scratch(eng->scratch_size());
    if (kind == Kind::C2C) {
      const cpx<T>* I = reinterpret_cast<const cpx<T>*>(inV);
      cpx<T>* O = reinterpret_cast<cpx<T>*>(outV);
      for (int t = 0; t < howmany; ++t) {
        const cpx<T>* ip = I + (ptrdiff_t)t * idist;
        for (int j = 0; j < n; ++j) buf[j] = ip[(ptrdiff_t)j * istride];
        eng->run(buf.data(), scratch.data());
        cpx<T>* op = O + (ptrdiff_t)t * odist;
        for (int j = 0; j < n; ++j) op[(ptrdiff_t)j * ostride] = buf[j];
      }
    } else if (kind == Kind::R2C) {
      const T* I = reinterpret_cast<const T*>(inV);
      cpx<T>* O = reinterpret_cast<cpx<T>*>(outV);
      for (int t = 0; t < howmany; ++t) {
        const T* ip = I + (ptrdiff_t)t * idist;
        for (int j = 0; j < n; ++j) buf[j] = cpx<T>(ip[(ptrdiff_t)j * istride], 0);
        eng->run(buf.data(), scratch.data());
        cpx<T>* op = O + (ptrdiff_t)t * odist;
        for (int j = 0; j <= n / 2; ++j) op[(ptrdiff_t)j * ostride] = buf[j];
      }
    } else {
      const cpx<T>* I = reinterpret_cast<const cpx<T>*>(inV);
      T* O = reinterpret_cast<T*>(outV);
      for (int t = 0; t < howmany; ++t) {
        const cpx<T>* ip = I + (ptrdiff_t)t * idist;
        buf[0] = cpx<T>(ip[0].real(), 0);
        for (int j = 1; j <= n / 2; ++j) {
          cpx<T> v = ip[(ptrdiff_t)j * istride];
          if (2 * j == n) {
            buf[j] = cpx<T>(v.real(), 0);
          } else {
            buf[j] = v;
            buf[n - j] = std::conj(v);
          }
        }
        eng->run(buf.data(), scratch.data());
        T* op = O + (ptrdiff_t)t * odist;
        for (int j = 0; j < n; ++j) op[(ptrdiff_t)j * ostride] = buf[j].real();
      }
    }
  }
};

template <typename T>
Plan<T>* make_plan(Kind kind, int rank, const int* n, int howmany, void* in, int istride,
                   int idist, void* out, int ostride, int odist, int sign) {
  if (rank != 1 || !n || n[0] < 0) return nullptr;
  Plan<T>* p = new Plan<T>();
  p->kind = kind;
  p->n = n[0];
  p->howmany = howmany;
  p->istride = istride;
  p->idist = idist;
  p->ostride = ostride;
  p->odist = odist;
  p->sign = sign;
  p->in = in;
  p->out = out;
  p->eng.reset(new Engine<T>(n[0] > 0 ? n[0] : 1, sign));
  return p;
}

}  // namespace

struct shim_plan_s {
  bool isFloat;
  void* impl;
};

static shim_plan_s* wrap(bool isFloat, void* impl) {
  if (!impl) return nullptr;
  shim_plan_s* s = new shim_plan_s();
  s->isFloat = isFloat;
  s->impl = impl;
  return s;
}

extern "C" {

// ---------------- double ----------------
fftw_plan fftw_plan_many_dft(int rank, const int* n, int howmany, fftw_complex* in, const int*,
                             int istride, int idist, fftw_complex* out, const int*, int ostride,
                             int odist, int sign, unsigned) {
  return wrap(false, make_plan<double>(Kind::C2C, rank, n, howmany, in, istride, idist, out,
                                       ostride, odist, sign));
}
fftw_plan fftw_plan_dft_1d(int n, fftw_complex* in, fftw_complex* out, int sign, unsigned flags) {
  return fftw_plan_many_dft(1, &n, 1, in, nullptr, 1, n, out, nullptr, 1, n, sign, flags);
}
fftw_plan fftw_plan_many_dft_r2c(int rank, const int* n, int howmany, double* in, const int*,
                                 int istride, int idist, fftw_complex* out, const int*, int ostride,
                                 int odist, unsigned) {
  return wrap(false, make_plan<double>(Kind::R2C, rank, n, howmany, in, istride, idist, out,
                                       ostride, odist, FFTW_FORWARD));
}
fftw_plan fftw_plan_many_dft_c2r(int rank, const int* n, int howmany, fftw_complex* in, const int*,
                                 int istride, int idist, double* out, const int*, int ostride,
                                 int odist, unsigned) {
  return wrap(false, make_plan<double>(Kind::C2R, rank, n, howmany, in, istride, idist, out,
                                       ostride, odist, FFTW_BACKWARD));
}
void fftw_execute(const fftw_plan p) {
  auto* pl = static_cast<Plan<double>*>(p->impl);
  pl->exec(pl->in, pl->out);
}
void fftw_execute_dft(const fftw_plan p, fftw_complex* in, fftw_complex* out) {
  static_cast<Plan<double>*>(p->impl)->exec(in, out);
}
void fftw_execute_dft_r2c(const fftw_plan p, double* in, fftw_complex* out) {
  static_cast<Plan<double>*>(p->impl)->exec(in, out);
}
void fftw_execute_dft_c2r(const fftw_plan p, fftw_complex* in, double* out) {
  static_cast<Plan<double>*>(p->impl)->exec(in, out);
}
void fftw_destroy_plan(fftw_plan p) {
  if (!p) return;
  delete static_cast<Plan<double>*>(p->impl);
  delete p;
}
int fftw_alignment_of(double* p) { return (int)(reinterpret_cast<uintptr_t>(p) % 16); }

// ---------------- float ----------------
fftwf_plan fftwf_plan_many_dft(int rank, const int* n, int howmany, fftwf_complex* in, const int*,
                               int istride, int idist, fftwf_complex* out, const int*, int ostride,
                               int odist, int sign, unsigned) {
  return wrap(true, make_plan<float>(Kind::C2C, rank, n, howmany, in, istride, idist, out, ostride,
                                     odist, sign));
}
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign,
                             unsigned flags) {
  return fftwf_plan_many_dft(1, &n, 1, in, nullptr, 1, n, out, nullptr, 1, n, sign, flags);
}
fftwf_plan fftwf_plan_many_dft_r2c(int rank, const int* n, int howmany, float* in, const int*,
                                   int istride, int idist, fftwf_complex* out, const int*,
                                   int ostride, int odist, unsigned) {
  return wrap(true, make_plan<float>(Kind::R2C, rank, n, howmany, in, istride, idist, out, ostride,
                                     odist, FFTW_FORWARD));
}
fftwf_plan fftwf_plan_many_dft_c2r(int rank, const int* n, int howmany, fftwf_complex* in,
                                   const int*, int istride, int idist, float* out, const int*,
                                   int ostride, int odist, unsigned) {
  return wrap(true, make_plan<float>(Kind::C2R, rank, n, howmany, in, istride, idist, out, ostride,
                                     odist, FFTW_BACKWARD));
}
void fftwf_execute(const fftwf_plan p) {
  auto* pl = static_cast<Plan<float>*>(p->impl);
  pl->exec(pl->in, pl->out);
}
void fftwf_execute_dft(const fftwf_plan p, fftwf_complex* in, fftwf_complex* out) {
  static_cast<Plan<float>*>(p->impl)->exec(in, out);
}
void fftwf_execute_dft_r2c(const fftwf_plan p, float* in, fftwf_complex* out) {
  static_cast<Plan<float>*>(p->impl)->exec(in, out);
}
void fftwf_execute_dft_c2r(const fftwf_plan p, fftwf_complex* in, float* out) {
  static_cast<Plan<float>*>(p->impl)->exec(in, out);
}
void fftwf_destroy_plan(fftwf_plan p) {
  if (!p) return;
  delete static_cast<Plan<float>*>(p->impl);
  delete p;
}
int fftwf_alignment_of(float* p) { return (int)(reinterpret_cast<uintptr_t>(p) % 16); }

}  // extern "C"
