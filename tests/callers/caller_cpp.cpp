// caller_cpp.cpp -- a C++ consumer of the spfft::Grid / spfft::Transform API (include/spfft/spfft.hpp)
// written as a user of the reference writes one: shared Grid, R2C and C2C transforms, centered
// indices, float twin, exceptions. Checked against a direct O(N^2) DFT.
//
// Build (tests/test_dropin_build.py): g++ -std=c++17 -Iinclude caller_cpp.cpp -Lspfft_b200/lib -lspfft_b200
// Exit code 0 = all checks passed; prints "CALLER_CPP PASS max_err=<e>".
#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>

#include "spfft/spfft.hpp"

int main() {
  const int nx = 6, ny = 4, nz = 3;
  const double twoPi = 6.283185307179586476925286766559;
  // centered indices: x in [-2, 3], y in [-1, 2], z in [-1, 1]; keep a sparse subset
  std::vector<int> idx;
  std::vector<std::complex<double>> freq;
  for (int x = -2; x <= 3; ++x)
    for (int y = -1; y <= 2; ++y)
      for (int z = -1; z <= 1; ++z)
        if ((x + 2 * y + 3 * z) % 3 != 0) {
          idx.insert(idx.end(), {x, y, z});
          freq.emplace_back(0.5 * x - y, 0.25 * z + 0.1 * x * y);
        }
  const int n = static_cast<int>(freq.size());
  std::vector<std::complex<double>> ref(static_cast<size_t>(nx) * ny * nz);
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x) {
        std::complex<double> s = 0;
        for (int k = 0; k < n; ++k) {
          const double a = twoPi * (double(x) * idx[3 * k] / nx + double(y) * idx[3 * k + 1] / ny +
                                    double(z) * idx[3 * k + 2] / nz);
          s += freq[k] * std::complex<double>(std::cos(a), std::sin(a));
        }
        ref[(static_cast<size_t>(z) * ny + y) * nx + x] = s;
      }
  double err = 0;
  try {
    spfft::Grid grid(nx, ny, nz, nx * ny, SPFFT_PU_GPU, -1);
    spfft::Transform t = grid.create_transform(SPFFT_PU_GPU, SPFFT_TRANS_C2C, nx, ny, nz, nz, n,
                                               SPFFT_INDEX_TRIPLETS, idx.data());
    if (t.dim_x() != nx || t.dim_y() != ny || t.dim_z() != nz || t.num_local_elements() != n ||
        t.local_z_length() != nz || t.local_z_offset() != 0 || t.processing_unit() != SPFFT_PU_GPU)
      return 2;
    std::vector<std::complex<double>> space(ref.size()), back(freq.size());
    t.backward(reinterpret_cast<const double*>(freq.data()), reinterpret_cast<double*>(space.data()));
    for (size_t i = 0; i < ref.size(); ++i) err = std::fmax(err, std::abs(space[i] - ref[i]));
    t.forward(reinterpret_cast<const double*>(space.data()), reinterpret_cast<double*>(back.data()),
              SPFFT_FULL_SCALING);
    for (size_t i = 0; i < freq.size(); ++i) err = std::fmax(err, std::abs(back[i] - freq[i]));

    // an independent transform (own grid) and the float twin
    spfft::Transform ti(-1, SPFFT_PU_GPU, SPFFT_TRANS_C2C, nx, ny, nz, n, SPFFT_INDEX_TRIPLETS, idx.data());
    ti.backward(reinterpret_cast<const double*>(freq.data()), SPFFT_PU_HOST);
    const auto* internal = reinterpret_cast<const std::complex<double>*>(ti.space_domain_data(SPFFT_PU_HOST));
    for (size_t i = 0; i < ref.size(); ++i) err = std::fmax(err, std::abs(internal[i] - ref[i]));

    spfft::TransformFloat tf(-1, SPFFT_PU_GPU, SPFFT_TRANS_C2C, nx, ny, nz, n, SPFFT_INDEX_TRIPLETS, idx.data());
    std::vector<std::complex<float>> freqF(freq.begin(), freq.end()), spaceF(ref.size());
    tf.backward(reinterpret_cast<const float*>(freqF.data()), reinterpret_cast<float*>(spaceF.data()));
    double errF = 0;
    for (size_t i = 0; i < ref.size(); ++i) errF = std::fmax(errF, std::abs(std::complex<double>(spaceF[i]) - ref[i]));
    if (errF > 1e-4) return 3;

    // out-of-range index -> InvalidIndicesError with the reference's error code
    std::vector<int> bad = {nx, 0, 0};
    bool thrown = false;
    try {
      spfft::Transform tb(-1, SPFFT_PU_GPU, SPFFT_TRANS_C2C, nx, ny, nz, 1, SPFFT_INDEX_TRIPLETS, bad.data());
    } catch (const spfft::InvalidIndicesError& e) {
      thrown = e.error_code() == SPFFT_INVALID_INDICES_ERROR;
    }
    if (!thrown) return 4;
  } catch (const spfft::GenericError& e) {
    std::fprintf(stderr, "SpFFT error %d: %s\n", static_cast<int>(e.error_code()), e.what());
    return 10;
  }
  std::printf("CALLER_CPP %s max_err=%.3e\n", err < 1e-10 ? "PASS" : "FAIL", err);
  return err < 1e-10 ? 0 : 1;
}
