/* caller_c.c -- a plain C consumer of the SpFFT C ABI (include/spfft/spfft.h), written the way a
 * user of the reference writes one (cf. the flow of the reference's examples/example.c: grid ->
 * transform -> destroy grid -> backward into the internal buffer -> backward / forward with external
 * buffers), but on SPFFT_PU_GPU and checked against a direct O(N^2) DFT.
 *
 * Build (tests/test_dropin_build.py): gcc -std=c99 -Iinclude caller_c.c -Lspfft_b200/lib -lspfft_b200 -lm
 * Exit code 0 = all checks passed; prints "CALLER_C PASS max_err=<e>". */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "spfft/spfft.h"

#define CHECK(call)                                                         \
  do {                                                                      \
    SpfftError e_ = (call);                                                 \
    if (e_ != SPFFT_SUCCESS) {                                              \
      fprintf(stderr, "%s failed with SpfftError %d\n", #call, (int)e_);    \
      return 10 + (int)e_;                                                  \
    }                                                                       \
  } while (0)

int main(void) {
  const int nx = 4, ny = 3, nz = 5;
  const int n = nx * ny * nz;
  const double twoPi = 6.283185307179586476925286766559;
  double* freq = (double*)malloc(2 * sizeof(double) * n);
  double* back = (double*)malloc(2 * sizeof(double) * n);
  double* space = (double*)malloc(2 * sizeof(double) * n);
  double* ref = (double*)calloc(2 * (size_t)n, sizeof(double));
  int* idx = (int*)malloc(3 * sizeof(int) * n);
  int c = 0;
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      for (int z = 0; z < nz; ++z, ++c) {
        freq[2 * c] = 0.25 * c - 3.0;
        freq[2 * c + 1] = 1.0 - 0.125 * c;
        idx[3 * c] = x;
        idx[3 * c + 1] = y;
        idx[3 * c + 2] = z;
      }
  /* direct backward DFT: space[z][y][x] = sum f(kx,ky,kz) exp(+2 pi i (x kx/nx + y ky/ny + z kz/nz)) */
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x) {
        double re = 0, im = 0;
        for (int k = 0; k < n; ++k) {
          const double a = twoPi * ((double)x * idx[3 * k] / nx + (double)y * idx[3 * k + 1] / ny +
                                    (double)z * idx[3 * k + 2] / nz);
          re += freq[2 * k] * cos(a) - freq[2 * k + 1] * sin(a);
          im += freq[2 * k] * sin(a) + freq[2 * k + 1] * cos(a);
        }
        ref[2 * ((z * ny + y) * nx + x)] = re;
        ref[2 * ((z * ny + y) * nx + x) + 1] = im;
      }

  SpfftGrid grid;
  CHECK(spfft_grid_create(&grid, nx, ny, nz, nx * ny, SPFFT_PU_GPU, -1));
  SpfftTransform t;
  CHECK(spfft_transform_create(&t, grid, SPFFT_PU_GPU, SPFFT_TRANS_C2C, nx, ny, nz, nz, n,
                               SPFFT_INDEX_TRIPLETS, idx));
  CHECK(spfft_grid_destroy(grid)); /* a transform keeps its grid alive */
  int v = 0;
  CHECK(spfft_transform_dim_x(t, &v));
  if (v != nx) return 2;
  CHECK(spfft_transform_num_local_elements(t, &v));
  if (v != n) return 3;

  /* option A: internal space-domain buffer, host view */
  double* internal = NULL;
  CHECK(spfft_transform_get_space_domain(t, SPFFT_PU_HOST, &internal));
  CHECK(spfft_transform_backward(t, freq, SPFFT_PU_HOST));
  double err = 0;
  for (int i = 0; i < 2 * n; ++i) err = fmax(err, fabs(internal[i] - ref[i]));

  /* option B: external buffers (host pointers), then the round trip with full scaling */
  CHECK(spfft_transform_backward_ptr(t, freq, space));
  for (int i = 0; i < 2 * n; ++i) err = fmax(err, fabs(space[i] - ref[i]));
  CHECK(spfft_transform_forward_ptr(t, space, back, SPFFT_FULL_SCALING));
  for (int i = 0; i < 2 * n; ++i) err = fmax(err, fabs(back[i] - freq[i]));

  /* clone + multi transform */
  SpfftTransform t2;
  CHECK(spfft_transform_clone(t, &t2));
  SpfftTransform both[2];
  const double* ins[2];
  double* outs[2];
  double* space2 = (double*)malloc(2 * sizeof(double) * n);
  both[0] = t;
  both[1] = t2;
  ins[0] = freq;
  ins[1] = freq;
  outs[0] = space;
  outs[1] = space2;
  CHECK(spfft_multi_transform_backward_ptr(2, both, ins, outs));
  for (int i = 0; i < 2 * n; ++i) err = fmax(err, fabs(space2[i] - ref[i]));

  /* errors come back as codes, never as crashes */
  if (spfft_transform_backward(NULL, freq, SPFFT_PU_HOST) != SPFFT_INVALID_HANDLE_ERROR) return 4;
  CHECK(spfft_transform_destroy(t2));
  CHECK(spfft_transform_destroy(t));
  printf("CALLER_C %s max_err=%.3e\n", err < 1e-10 ? "PASS" : "FAIL", err);
  free(freq); free(back); free(space); free(space2); free(ref); free(idx);
  return err < 1e-10 ? 0 : 1;
}
