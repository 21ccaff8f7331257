// spfft/multi_transform.hpp -- run several independent transforms together (float).
// Reference: include/spfft/multi_transform_float.hpp. Transforms must not share a Grid.
#ifndef SPFFT_MULTI_TRANSFORM_FLOAT_HPP
#define SPFFT_MULTI_TRANSFORM_FLOAT_HPP
#include "spfft/config.h"
#include "spfft/transform_float.hpp"
#include "spfft/types.h"
namespace spfft {
SPFFT_EXPORT void multi_transform_forward(int numTransforms, TransformFloat* transforms,
                                          const SpfftProcessingUnitType* inputLocations,
                                          float* const* outputPointers,
                                          const SpfftScalingType* scalingTypes);
SPFFT_EXPORT void multi_transform_forward(int numTransforms, TransformFloat* transforms,
                                          const float* const* inputPointers,
                                          float* const* outputPointers,
                                          const SpfftScalingType* scalingTypes);
SPFFT_EXPORT void multi_transform_backward(int numTransforms, TransformFloat* transforms,
                                           const float* const* inputPointers,
                                           const SpfftProcessingUnitType* outputLocations);
SPFFT_EXPORT void multi_transform_backward(int numTransforms, TransformFloat* transforms,
                                           const float* const* inputPointers,
                                           float* const* outputPointers);
}  // namespace spfft
#endif
