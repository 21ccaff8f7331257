/* spfft/transform_float.h -- C API, float. See the .inc files in spfft/detail for the documented declarations. */
#ifndef SPFFT_TRANSFORM_FLOAT_H
#define SPFFT_TRANSFORM_FLOAT_H
#include "spfft/config.h"
#include "spfft/errors.h"
#include "spfft/types.h"
#include "spfft/grid_float.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef void* SpfftFloatTransform;
#define SPFFT_FN(name) spfft_float_##name
#define SPFFT_GRID_T SpfftFloatGrid
#define SPFFT_TRANSFORM_T SpfftFloatTransform
#define SPFFT_REAL float
#include "spfft/detail/transform_api.inc"
#undef SPFFT_FN
#undef SPFFT_GRID_T
#undef SPFFT_TRANSFORM_T
#undef SPFFT_REAL
#ifdef __cplusplus
}
#endif
#endif
