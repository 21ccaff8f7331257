/*
 * fftw3.h -- minimal FFTW-3 API surface for building the reference host pipeline in this image.
 *
 * TEST INFRASTRUCTURE ONLY (oracle). FFTW itself is not installed here and cannot be fetched
 * (no network), so the unmodified reference host sources under /root/reference/src are compiled
 * against this shim (see oracle/Makefile). Only the entry points the reference actually calls
 * are provided (call sites: src/fft/fftw_plan_1d.hpp:77-139,196-215, src/fft/fftw_interface.hpp):
 *
 *   fftw[f]_plan_many_dft, fftw[f]_plan_many_dft_r2c, fftw[f]_plan_many_dft_c2r,
 *   fftw[f]_plan_dft_1d, fftw[f]_execute, fftw[f]_execute_dft[_r2c|_c2r],
 *   fftw[f]_destroy_plan, fftw[f]_alignment_of
 *
 * Semantics follow the published FFTW 3.3 manual: unnormalised DFT, sign -1 = FORWARD,
 * +1 = BACKWARD, rank-1 "many" interface with istride/ostride/idist/odist/howmany, r2c output
 * and c2r input of n/2+1 complex numbers. The arithmetic lives in fftw3_shim.cpp.
 */
#ifndef SPFFT_B200_ORACLE_FFTW3_SHIM_H
#define SPFFT_B200_ORACLE_FFTW3_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

typedef double fftw_complex[2];
typedef float fftwf_complex[2];

struct shim_plan_s;
typedef struct shim_plan_s* fftw_plan;
typedef struct shim_plan_s* fftwf_plan;

/* double precision */
fftw_plan fftw_plan_dft_1d(int n, fftw_complex* in, fftw_complex* out, int sign, unsigned flags);
fftw_plan fftw_plan_many_dft(int rank, const int* n, int howmany, fftw_complex* in,
                             const int* inembed, int istride, int idist, fftw_complex* out,
                             const int* onembed, int ostride, int odist, int sign, unsigned flags);
fftw_plan fftw_plan_many_dft_r2c(int rank, const int* n, int howmany, double* in,
                                 const int* inembed, int istride, int idist, fftw_complex* out,
                                 const int* onembed, int ostride, int odist, unsigned flags);
fftw_plan fftw_plan_many_dft_c2r(int rank, const int* n, int howmany, fftw_complex* in,
                                 const int* inembed, int istride, int idist, double* out,
                                 const int* onembed, int ostride, int odist, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_execute_dft(const fftw_plan p, fftw_complex* in, fftw_complex* out);
void fftw_execute_dft_r2c(const fftw_plan p, double* in, fftw_complex* out);
void fftw_execute_dft_c2r(const fftw_plan p, fftw_complex* in, double* out);
void fftw_destroy_plan(fftw_plan p);
int fftw_alignment_of(double* p);

/* single precision */
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign,
                             unsigned flags);
fftwf_plan fftwf_plan_many_dft(int rank, const int* n, int howmany, fftwf_complex* in,
                               const int* inembed, int istride, int idist, fftwf_complex* out,
                               const int* onembed, int ostride, int odist, int sign,
                               unsigned flags);
fftwf_plan fftwf_plan_many_dft_r2c(int rank, const int* n, int howmany, float* in,
                                   const int* inembed, int istride, int idist, fftwf_complex* out,
                                   const int* onembed, int ostride, int odist, unsigned flags);
fftwf_plan fftwf_plan_many_dft_c2r(int rank, const int* n, int howmany, fftwf_complex* in,
                                   const int* inembed, int istride, int idist, float* out,
                                   const int* onembed, int ostride, int odist, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_execute_dft(const fftwf_plan p, fftwf_complex* in, fftwf_complex* out);
void fftwf_execute_dft_r2c(const fftwf_plan p, float* in, fftwf_complex* out);
void fftwf_execute_dft_c2r(const fftwf_plan p, fftwf_complex* in, float* out);
void fftwf_destroy_plan(fftwf_plan p);
int fftwf_alignment_of(float* p);

#ifdef __cplusplus
}
#endif
#endif
