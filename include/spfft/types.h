/* spfft/types.h -- enumerations of the SpFFT API. Names and numeric values are those of the
 * reference (include/spfft/types.h:33-117) so that existing callers bind unchanged. */
#ifndef SPFFT_TYPES_H
#define SPFFT_TYPES_H
#include "spfft/config.h"

/* Stick<->slab exchange flavour. The B200 build has one exchange (per-destination blocks over
 * NVLink); every value is accepted, the *_FLOAT ones select an fp32 wire format. */
enum SpfftExchangeType {
  SPFFT_EXCH_DEFAULT,
  SPFFT_EXCH_BUFFERED,
  SPFFT_EXCH_BUFFERED_FLOAT,
  SPFFT_EXCH_COMPACT_BUFFERED,
  SPFFT_EXCH_COMPACT_BUFFERED_FLOAT,
  SPFFT_EXCH_UNBUFFERED
};

/* Where data lives / where a transform executes (bit flags). */
enum SpfftProcessingUnitType { SPFFT_PU_HOST = 1, SPFFT_PU_GPU = 2 };

/* Format of the sparse frequency indices: (x, y, z) triplets. */
enum SpfftIndexFormatType { SPFFT_INDEX_TRIPLETS };

/* Complex-to-complex, or real space domain with hermitian frequency domain. */
enum SpfftTransformType { SPFFT_TRANS_C2C, SPFFT_TRANS_R2C };

/* Forward transforms may be scaled by 1/(Nx*Ny*Nz). */
enum SpfftScalingType { SPFFT_NO_SCALING, SPFFT_FULL_SCALING };

/* Synchronous: calls return after the GPU work completed. Asynchronous: work is only ordered
 * with respect to the default stream. */
enum SpfftExecType { SPFFT_EXEC_SYNCHRONOUS, SPFFT_EXEC_ASYNCHRONOUS };

/* Handle of an NCCL based communicator (spfft/b200_ext.h): stands where the reference's API takes
 * an MPI_Comm. */
typedef void* SpfftB200Comm;

#ifndef __cplusplus
typedef enum SpfftExchangeType SpfftExchangeType;
typedef enum SpfftProcessingUnitType SpfftProcessingUnitType;
typedef enum SpfftTransformType SpfftTransformType;
typedef enum SpfftIndexFormatType SpfftIndexFormatType;
typedef enum SpfftScalingType SpfftScalingType;
typedef enum SpfftExecType SpfftExecType;
#endif
#endif
