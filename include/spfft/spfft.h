/* spfft/spfft.h -- umbrella header of the C API (reference include/spfft/spfft.h). */
#ifndef SPFFT_SPFFT_H
#define SPFFT_SPFFT_H
#include "spfft/config.h"
#include "spfft/grid.h"
#include "spfft/grid_float.h"
#include "spfft/multi_transform.h"
#include "spfft/multi_transform_float.h"
#include "spfft/transform.h"
#include "spfft/transform_float.h"
#include "spfft/b200_ext.h"
#endif
