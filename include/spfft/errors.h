/* spfft/errors.h -- C error codes. Order and values follow the reference
 * (include/spfft/errors.h:37-125); codes that cannot occur in this build (MPI, FFTW, host
 * execution) are kept so that numeric values stay stable for callers. */
#ifndef SPFFT_ERRORS_H
#define SPFFT_ERRORS_H
#include "spfft/config.h"

enum SpfftError {
  SPFFT_SUCCESS,                      /* 0 */
  SPFFT_UNKNOWN_ERROR,                /* 1 */
  SPFFT_INVALID_HANDLE_ERROR,         /* 2  null Grid / Transform handle */
  SPFFT_OVERFLOW_ERROR,               /* 3 */
  SPFFT_ALLOCATION_ERROR,             /* 4  host allocation */
  SPFFT_INVALID_PARAMETER_ERROR,      /* 5 */
  SPFFT_DUPLICATE_INDICES_ERROR,      /* 6  same z-stick on two ranks */
  SPFFT_INVALID_INDICES_ERROR,        /* 7  index outside the grid */
  SPFFT_MPI_SUPPORT_ERROR,            /* 8 */
  SPFFT_MPI_ERROR,                    /* 9  also reported for NCCL failures */
  SPFFT_MPI_PARAMETER_MISMATCH_ERROR, /* 10 */
  SPFFT_HOST_EXECUTION_ERROR,         /* 11 */
  SPFFT_FFTW_ERROR,                   /* 12 (the reference's InternalError maps here) */
  SPFFT_GPU_ERROR,                    /* 13 */
  SPFFT_GPU_PRECEDING_ERROR,          /* 14 sticky CUDA error found at entry */
  SPFFT_GPU_SUPPORT_ERROR,            /* 15 */
  SPFFT_GPU_ALLOCATION_ERROR,         /* 16 */
  SPFFT_GPU_LAUNCH_ERROR,             /* 17 */
  SPFFT_GPU_NO_DEVICE_ERROR,          /* 18 */
  SPFFT_GPU_INVALID_VALUE_ERROR,      /* 19 */
  SPFFT_GPU_INVALID_DEVICE_PTR_ERROR, /* 20 */
  SPFFT_GPU_COPY_ERROR,               /* 21 */
  SPFFT_GPU_FFT_ERROR                 /* 22 */
};

#ifndef __cplusplus
typedef enum SpfftError SpfftError;
#endif
#endif
