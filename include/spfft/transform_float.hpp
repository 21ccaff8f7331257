// spfft/transform_float.hpp -- C++ API (float). The class declarations live in spfft/detail/classes.inc.
#ifndef SPFFT_TRANSFORM_FLOAT_HPP
#define SPFFT_TRANSFORM_FLOAT_HPP
#include <memory>
#include "spfft/config.h"
#include "spfft/types.h"
#include "spfft/detail/fwd.hpp"
#ifndef SPFFT_DETAIL_CLASSES_FLOAT
#define SPFFT_DETAIL_CLASSES_FLOAT
namespace spfft {
#define SPFFT_GRID_CLASS GridFloat
#define SPFFT_TRANSFORM_CLASS TransformFloat
#define SPFFT_REAL float
#include "spfft/detail/classes.inc"
#undef SPFFT_GRID_CLASS
#undef SPFFT_TRANSFORM_CLASS
#undef SPFFT_REAL
}  // namespace spfft
#endif
#endif
