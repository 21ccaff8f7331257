// Forward declarations of the implementation types behind the public handles.
#ifndef SPFFT_DETAIL_FWD_HPP
#define SPFFT_DETAIL_FWD_HPP
#include "spfft/config.h"
namespace spfft {
namespace b200 {
template <typename T>
class SPFFT_NO_EXPORT GridResources;
template <typename T>
class SPFFT_NO_EXPORT TransformEngine;
}  // namespace b200
}  // namespace spfft
#endif
