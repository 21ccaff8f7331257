/* spfft/grid.h -- C API, double. See the .inc files in spfft/detail for the documented declarations. */
#ifndef SPFFT_GRID_H
#define SPFFT_GRID_H
#include "spfft/config.h"
#include "spfft/errors.h"
#include "spfft/types.h"

#ifdef __cplusplus
extern "C" {
#endif
typedef void* SpfftGrid;
#define SPFFT_FN(name) spfft_##name
#define SPFFT_GRID_T SpfftGrid
#define SPFFT_TRANSFORM_T SpfftTransform
#define SPFFT_REAL double
#include "spfft/detail/grid_api.inc"
#undef SPFFT_FN
#undef SPFFT_GRID_T
#undef SPFFT_TRANSFORM_T
#undef SPFFT_REAL
#ifdef __cplusplus
}
#endif
#endif
