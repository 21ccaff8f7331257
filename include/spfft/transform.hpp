// spfft/transform.hpp -- C++ API (double). The class declarations live in spfft/detail/classes.inc.
#ifndef SPFFT_TRANSFORM_HPP
#define SPFFT_TRANSFORM_HPP
#include <memory>
#include "spfft/config.h"
#include "spfft/types.h"
#include "spfft/detail/fwd.hpp"
#ifndef SPFFT_DETAIL_CLASSES_DOUBLE
#define SPFFT_DETAIL_CLASSES_DOUBLE
namespace spfft {
#define SPFFT_GRID_CLASS Grid
#define SPFFT_TRANSFORM_CLASS Transform
#define SPFFT_REAL double
#include "spfft/detail/classes.inc"
#undef SPFFT_GRID_CLASS
#undef SPFFT_TRANSFORM_CLASS
#undef SPFFT_REAL
}  // namespace spfft
#endif
#endif
