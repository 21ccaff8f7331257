// spfft/multi_transform.hpp -- run several independent transforms together (double).
// Reference: include/spfft/multi_transform.hpp:51-97. Transforms must not share a Grid.
#ifndef SPFFT_MULTI_TRANSFORM_HPP
#define SPFFT_MULTI_TRANSFORM_HPP
#include "spfft/config.h"
#include "spfft/transform.hpp"
#include "spfft/types.h"
namespace spfft {
SPFFT_EXPORT void multi_transform_forward(int numTransforms, Transform* transforms,
                                          const SpfftProcessingUnitType* inputLocations,
                                          double* const* outputPointers,
                                          const SpfftScalingType* scalingTypes);
SPFFT_EXPORT void multi_transform_forward(int numTransforms, Transform* transforms,
                                          const double* const* inputPointers,
                                          double* const* outputPointers,
                                          const SpfftScalingType* scalingTypes);
SPFFT_EXPORT void multi_transform_backward(int numTransforms, Transform* transforms,
                                           const double* const* inputPointers,
                                           const SpfftProcessingUnitType* outputLocations);
SPFFT_EXPORT void multi_transform_backward(int numTransforms, Transform* transforms,
                                           const double* const* inputPointers,
                                           double* const* outputPointers);
}  // namespace spfft
#endif
