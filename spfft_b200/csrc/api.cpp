// api.cpp -- the drop-in boundary: spfft::Grid / Transform / GridFloat / TransformFloat,
// multi_transform_* and every extern "C" entry point of include/spfft/*.h, instantiated for both
// precisions from api_impl.inc.
#include <memory>
#include <vector>

#include "index_plan.hpp"
#include "launch.h"
#include "spfft/spfft.h"
#include "spfft/spfft.hpp"
#include "transform_engine.hpp"

// ---- double ----
#define SPFFT_GRID_CLASS Grid
#define SPFFT_TRANSFORM_CLASS Transform
#define SPFFT_REAL double
#define SPFFT_FN(name) spfft_##name
#define SPFFT_EXT_FN(name) spfft_b200_##name
#define SPFFT_GRID_T SpfftGrid
#define SPFFT_TRANSFORM_T SpfftTransform
#include "api_impl.inc"
#undef SPFFT_GRID_CLASS
#undef SPFFT_TRANSFORM_CLASS
#undef SPFFT_REAL
#undef SPFFT_FN
#undef SPFFT_EXT_FN
#undef SPFFT_GRID_T
#undef SPFFT_TRANSFORM_T

// ---- float ----
#define SPFFT_GRID_CLASS GridFloat
#define SPFFT_TRANSFORM_CLASS TransformFloat
#define SPFFT_REAL float
#define SPFFT_FN(name) spfft_float_##name
#define SPFFT_EXT_FN(name) spfft_b200_float_##name
#define SPFFT_GRID_T SpfftFloatGrid
#define SPFFT_TRANSFORM_T SpfftFloatTransform
#include "api_impl.inc"
#undef SPFFT_GRID_CLASS
#undef SPFFT_TRANSFORM_CLASS
#undef SPFFT_REAL
#undef SPFFT_FN
#undef SPFFT_EXT_FN
#undef SPFFT_GRID_T
#undef SPFFT_TRANSFORM_T

// ---- precision independent extensions ----
extern "C" {

SpfftError spfft_b200_convert_index_triplets(int hermitianSymmetry, int dimX, int dimY, int dimZ,
                                             int numValues, const int* triplets, int* valueIndices,
                                             int* stickIndices, int* numSticks) {
  try {
    if (dimX < 0 || dimY < 0 || dimZ < 0 || numValues < 0 || (!triplets && numValues > 0))
      throw spfft::InvalidParameterError();
    std::vector<int> vi, si;
    spfft::b200::convert_index_triplets(hermitianSymmetry != 0, dimX, dimY, dimZ, numValues,
                                        triplets, vi, si);
    if (valueIndices) std::copy(vi.begin(), vi.end(), valueIndices);
    if (stickIndices) std::copy(si.begin(), si.end(), stickIndices);
    if (numSticks) *numSticks = static_cast<int>(si.size());
  } catch (const spfft::GenericError& e) {
    return e.error_code();
  } catch (...) {
    return SPFFT_UNKNOWN_ERROR;
  }
  return SPFFT_SUCCESS;
}

SpfftError spfft_b200_kernel_launch_count(long long int* count) {
  if (!count) return SPFFT_INVALID_PARAMETER_ERROR;
  *count = sb_launch_count();
  return SPFFT_SUCCESS;
}

}  // extern "C"
