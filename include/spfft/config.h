/* spfft/config.h -- build configuration of the B200-native drop-in.
 * Stands where the reference's CMake-generated header stands (template:
 * /root/reference/include/spfft/config.h.in). This build is CUDA-only (sm_100a), always carries
 * the single-precision API, has no MPI (the distributed entry points live in spfft/b200_ext.h and
 * use NCCL) and no host (CPU) execution path. */
#ifndef SPFFT_CONFIG_H
#define SPFFT_CONFIG_H
#define SPFFT_CUDA
#define SPFFT_SINGLE_PRECISION
#define SPFFT_B200 1
#include "spfft/spfft_export.h"
#endif
