/*
 * ref_indices_wrapper.cpp -- TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Exposes the reference's OWN header-only index conversion
 * (spfft::convert_index_triplets, /root/reference/src/compression/indices.hpp:120-186) through a
 * C symbol so that tests can pin our restatements (oracle/spfft_oracle.py and the product's
 * plan builder) bit-for-bit against it. The reference header is #included from where it lies;
 * nothing is copied. Built only when /root/reference is present (oracle/Makefile).
 */
#include <cstring>
#include <vector>
#include "compression/indices.hpp"

extern "C" __attribute__((visibility("default")))
int spfft_ref_convert_index_triplets(int hermitian, int dimX, int dimY, int dimZ, int numValues,
                                     const int* triplets, int* valueIndices, int* stickIndices,
                                     int* numSticks) {
  try {
    auto res = spfft::convert_index_triplets(hermitian != 0, dimX, dimY, dimZ, numValues, triplets,
                                             triplets + 1, triplets + 2, 3);
    if (valueIndices && !res.first.empty())
      std::memcpy(valueIndices, res.first.data(), sizeof(int) * res.first.size());
    if (stickIndices && !res.second.empty())
      std::memcpy(stickIndices, res.second.data(), sizeof(int) * res.second.size());
    *numSticks = (int)res.second.size();
    return 0;
  } catch (const spfft::GenericError& e) {
    return (int)e.error_code();
  } catch (...) {
    return 1;
  }
}
