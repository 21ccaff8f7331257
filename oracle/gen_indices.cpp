/*
 * gen_indices.cpp -- TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Restates the random sparse-index fixture used by the reference's own tests so that our parity
 * tests run on the very same index sets and values:
 *   - create_value_indices    tests/test_util/generate_indices.hpp:38-85
 *   - center_indices          tests/test_util/generate_indices.hpp:87-99
 *   - calculate_num_local_xy_planes  tests/test_util/generate_indices.hpp:102-136
 *   - value draw order        tests/test_util/test_transform.hpp:80-92 (C2C: re then im per
 *                             element, ranks in order) and :221-229 (R2C: real part only)
 * std::mt19937 is fully specified, but std::uniform_real_distribution / discrete_distribution
 * are implementation defined, so this must be compiled with the same libstdc++ the reference's
 * tests would use here -- hence C++ rather than numpy.
 *
 * C interface (ctypes): two-call pattern, first with null buffers to obtain sizes.
 */
#include <algorithm>
#include <cstring>
#include <numeric>
#include <random>
#include <utility>
#include <vector>

namespace {

struct Fixture {
  std::vector<std::vector<int>> tripletsPerRank;
  std::vector<std::vector<double>> valuesPerRank;  // interleaved re,im
};

Fixture make_fixture(unsigned seed, int numRanks, const double* stickDistribution,
                     double stickFraction, double fillFraction, int dimX, int dimY, int dimZ,
                     bool hermitian, bool realValuesOnly) {
  std::mt19937 gen(seed);
  std::uniform_real_distribution<double> uni(0.0, 1.0);
  std::discrete_distribution<int> pickRank(stickDistribution, stickDistribution + numRanks);

  std::vector<std::vector<std::pair<int, int>>> xyPerRank(numRanks);
  const int xEnd = hermitian ? dimX / 2 + 1 : dimX;
  const int yHalf = hermitian ? dimY / 2 + 1 : dimY;
  for (int x = 0; x < xEnd; ++x) {
    for (int y = 0; y < dimY; ++y) {
      // the short-circuit order matters: no random draw for the skipped half of the x=0 plane
      if (!(x == 0 && y >= yHalf) && uni(gen) < stickFraction) {
        if (!hermitian || x != 0 || y < yHalf) {
          const int r = pickRank(gen);
          xyPerRank[r].emplace_back(x, y);
        }
      }
    }
  }

  Fixture f;
  f.tripletsPerRank.resize(numRanks);
  f.valuesPerRank.resize(numRanks);
  const int zHalf = hermitian ? dimZ / 2 + 1 : dimZ;
  for (int r = 0; r < numRanks; ++r) {
    for (const auto& xy : xyPerRank[r]) {
      for (int z = 0; z < dimZ; ++z) {
        if (!(hermitian && xy.first == 0 && xy.second == 0 && z >= zHalf) &&
            uni(gen) < fillFraction) {
          f.tripletsPerRank[r].push_back(xy.first);
          f.tripletsPerRank[r].push_back(xy.second);
          f.tripletsPerRank[r].push_back(z);
        }
      }
    }
  }
  // values: drawn rank by rank, element by element, from the SAME generator afterwards
  for (int r = 0; r < numRanks; ++r) {
    const size_t n = f.tripletsPerRank[r].size() / 3;
    f.valuesPerRank[r].resize(2 * n);
    for (size_t i = 0; i < n; ++i) {
      f.valuesPerRank[r][2 * i] = uni(gen);
      f.valuesPerRank[r][2 * i + 1] = realValuesOnly ? 0.0 : uni(gen);
    }
  }
  return f;
}

}  // namespace

extern "C" {

// Returns the number of elements of `rank`; fills triplets (3*n ints) / values (2*n doubles) when
// the buffers are non-null. center != 0 applies center_indices to the triplets.
long long spfft_oracle_gen_fixture(unsigned seed, int numRanks, const double* stickDistribution,
                                   double stickFraction, double fillFraction, int dimX, int dimY,
                                   int dimZ, int hermitian, int realValuesOnly, int center,
                                   int rank, int* triplets, double* values) {
  Fixture f = make_fixture(seed, numRanks, stickDistribution, stickFraction, fillFraction, dimX,
                           dimY, dimZ, hermitian != 0, realValuesOnly != 0);
  std::vector<int>& t = f.tripletsPerRank[rank];
  if (center) {
    const int px = dimX / 2 + 1, py = dimY / 2 + 1, pz = dimZ / 2 + 1;
    for (size_t i = 0; i < t.size(); i += 3) {
      if (t[i] >= px) t[i] -= dimX;
      if (t[i + 1] >= py) t[i + 1] -= dimY;
      if (t[i + 2] >= pz) t[i + 2] -= dimZ;
    }
  }
  if (triplets && !t.empty()) std::memcpy(triplets, t.data(), sizeof(int) * t.size());
  if (values && !t.empty())
    std::memcpy(values, f.valuesPerRank[rank].data(), sizeof(double) * f.valuesPerRank[rank].size());
  return (long long)(t.size() / 3);
}

// calculate_num_local_xy_planes for every rank (out has numRanks entries)
void spfft_oracle_plane_split(int dimZ, int numRanks, const double* planeDistribution, int* out) {
  const double sum = std::accumulate(planeDistribution, planeDistribution + numRanks, 0.0);
  std::vector<int> n(numRanks);
  for (int i = 0; i < numRanks; ++i) n[i] = (int)(planeDistribution[i] / sum * dimZ);
  int missing = dimZ - std::accumulate(n.begin(), n.end(), 0);
  for (auto& v : n) {
    if (v > 0 && missing > 0) {
      v += missing;
      missing = 0;
      break;
    }
    if (missing < 0) {
      v -= std::min(v, -missing);
      missing += v;
      if (missing >= 0) {
        missing = 0;
        break;
      }
    }
  }
  if (missing > 0) n[0] = missing;
  for (int i = 0; i < numRanks; ++i) out[i] = n[i];
}

}  // extern "C"
