// spfft/exceptions.hpp -- C++ exception types of the SpFFT API (header only).
// Class names, hierarchy, messages and error_code() mapping follow the reference
// (include/spfft/exceptions.hpp:40-300), including its quirk that InternalError reports
// SPFFT_FFTW_ERROR (:170-175).
#ifndef SPFFT_EXCEPTIONS_H
#define SPFFT_EXCEPTIONS_H
#include <stdexcept>
#include "spfft/config.h"
#include "spfft/errors.h"

namespace spfft {

class SPFFT_EXPORT GenericError : public std::exception {
public:
  const char* what() const noexcept override { return "SpFFT: Generic error"; }
  virtual SpfftError error_code() const noexcept { return SPFFT_UNKNOWN_ERROR; }
};

#define SPFFT_DEFINE_ERROR(NAME, BASE, CODE, MESSAGE)                        \
  class SPFFT_EXPORT NAME : public BASE {                                    \
  public:                                                                    \
    const char* what() const noexcept override { return MESSAGE; }           \
    SpfftError error_code() const noexcept override { return CODE; }         \
  };

SPFFT_DEFINE_ERROR(OverflowError, GenericError, SPFFT_OVERFLOW_ERROR, "SpFFT: Overflow error")
SPFFT_DEFINE_ERROR(HostAllocationError, GenericError, SPFFT_ALLOCATION_ERROR,
                   "SpFFT: Host allocation error")
SPFFT_DEFINE_ERROR(InvalidParameterError, GenericError, SPFFT_INVALID_PARAMETER_ERROR,
                   "SpFFT: Invalid parameter error")
SPFFT_DEFINE_ERROR(DuplicateIndicesError, GenericError, SPFFT_DUPLICATE_INDICES_ERROR,
                   "SpFFT: Duplicate indices error")
SPFFT_DEFINE_ERROR(InvalidIndicesError, GenericError, SPFFT_INVALID_INDICES_ERROR,
                   "SpFFT: Invalid indices error")
SPFFT_DEFINE_ERROR(MPISupportError, GenericError, SPFFT_MPI_SUPPORT_ERROR,
                   "SpFFT: Not compiled with MPI support error")
SPFFT_DEFINE_ERROR(MPIError, GenericError, SPFFT_MPI_ERROR, "SpFFT: MPI error")
SPFFT_DEFINE_ERROR(MPIParameterMismatchError, GenericError, SPFFT_MPI_PARAMETER_MISMATCH_ERROR,
                   "SpFFT: Mismatched parameters between MPI ranks")
SPFFT_DEFINE_ERROR(HostExecutionError, GenericError, SPFFT_HOST_EXECUTION_ERROR,
                   "SpFFT: Host execution error")
SPFFT_DEFINE_ERROR(FFTWError, GenericError, SPFFT_FFTW_ERROR, "SpFFT: FFTW error")
SPFFT_DEFINE_ERROR(InternalError, GenericError, SPFFT_FFTW_ERROR, "SpFFT: Internal error")
SPFFT_DEFINE_ERROR(GPUError, GenericError, SPFFT_GPU_ERROR, "SpFFT: GPU error")
SPFFT_DEFINE_ERROR(GPUSupportError, GPUError, SPFFT_GPU_SUPPORT_ERROR,
                   "SpFFT: Not compiled with GPU support")
SPFFT_DEFINE_ERROR(GPUPrecedingError, GPUError, SPFFT_GPU_PRECEDING_ERROR,
                   "SpFFT: Detected error from preceding gpu calls.")
SPFFT_DEFINE_ERROR(GPUAllocationError, GPUError, SPFFT_GPU_ALLOCATION_ERROR,
                   "SpFFT: GPU allocation error")
SPFFT_DEFINE_ERROR(GPULaunchError, GPUError, SPFFT_GPU_LAUNCH_ERROR, "SpFFT: GPU launch error")
SPFFT_DEFINE_ERROR(GPUNoDeviceError, GPUError, SPFFT_GPU_NO_DEVICE_ERROR,
                   "SpFFT: no GPU available")
SPFFT_DEFINE_ERROR(GPUInvalidValueError, GPUError, SPFFT_GPU_INVALID_VALUE_ERROR,
                   "SpFFT: GPU call with invalid value")
SPFFT_DEFINE_ERROR(GPUInvalidDevicePointerError, GPUError, SPFFT_GPU_INVALID_DEVICE_PTR_ERROR,
                   "SpFFT: Invalid GPU pointer")
SPFFT_DEFINE_ERROR(GPUCopyError, GPUError, SPFFT_GPU_COPY_ERROR, "SpFFT: GPU Memory copy error")
SPFFT_DEFINE_ERROR(GPUFFTError, GPUError, SPFFT_GPU_FFT_ERROR, "SpFFT: GPU FFT error")

#undef SPFFT_DEFINE_ERROR

}  // namespace spfft
#endif
