// index_plan.cpp -- see index_plan.hpp.
#include "index_plan.hpp"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>

#include "fast3_stage_kernels.hpp"
#include "fast_stage_kernels.hpp"
#include "spfft/exceptions.hpp"

namespace spfft {
namespace b200 {

namespace {
inline int storage_index(int dim, int idx) { return idx < 0 ? idx + dim : idx; }
}  // namespace

// Same results as the reference's std::map based routine (indices.hpp:120-186), computed with a
// presence table + prefix sum: O(Ne + Nx*Ny) instead of O(Ne log Ns).
void convert_index_triplets(bool hermitianSymmetry, int dimX, int dimY, int dimZ, int numValues,
                            const int* triplets, std::vector<int>& valueIndices,
                            std::vector<int>& stickIndices) {
  valueIndices.clear();
  stickIndices.clear();
  if (static_cast<unsigned long long>(numValues) >
      static_cast<unsigned long long>(dimX) * static_cast<unsigned long long>(dimY) *
          static_cast<unsigned long long>(dimZ)) {
    throw InvalidParameterError();  // indices.hpp:124-127
  }
  if (numValues == 0) return;

  // "centered" is a property of the whole index set, decided over all three coordinates
  // (indices.hpp:129-135)
  bool centered = false;
  for (long long i = 0; i < 3LL * numValues; ++i) {
    if (triplets[i] < 0) {
      centered = true;
      break;
    }
  }
  const int maxX = ((hermitianSymmetry || centered) ? dimX / 2 + 1 : dimX) - 1;
  const int maxY = (centered ? dimY / 2 + 1 : dimY) - 1;
  const int maxZ = (centered ? dimZ / 2 + 1 : dimZ) - 1;
  const int minX = hermitianSymmetry ? 0 : maxX - dimX + 1;
  const int minY = maxY - dimY + 1;
  const int minZ = maxZ - dimZ + 1;
  for (long long i = 0; i < numValues; ++i) {
    const int x = triplets[3 * i], y = triplets[3 * i + 1], z = triplets[3 * i + 2];
    if (x < minX || x > maxX || y < minY || y > maxY || z < minZ || z > maxZ)
      throw InvalidIndicesError();  // indices.hpp:145-149
  }

  const long long planeSize = static_cast<long long>(dimX) * dimY;
  valueIndices.resize(numValues);
  if (planeSize <= (1LL << 28)) {
    // presence table -> rank of every occupied (x,y) key
    std::vector<int> rankOfKey(static_cast<size_t>(planeSize), 0);
    for (long long i = 0; i < numValues; ++i) {
      const int x = storage_index(dimX, triplets[3 * i]);
      const int y = storage_index(dimY, triplets[3 * i + 1]);
      rankOfKey[static_cast<size_t>(x) * dimY + y] = 1;
    }
    int count = 0;
    for (long long k = 0; k < planeSize; ++k) {
      if (rankOfKey[k]) {
        rankOfKey[k] = count++;
        stickIndices.push_back(static_cast<int>(k));
      }
    }
    for (long long i = 0; i < numValues; ++i) {
      const int x = storage_index(dimX, triplets[3 * i]);
      const int y = storage_index(dimY, triplets[3 * i + 1]);
      const int z = storage_index(dimZ, triplets[3 * i + 2]);
      valueIndices[i] = rankOfKey[static_cast<size_t>(x) * dimY + y] * dimZ + z;
    }
  } else {
    // very large xy extent: sort the keys instead of tabulating the plane
    std::vector<int> keys(numValues);
    for (long long i = 0; i < numValues; ++i) {
      keys[i] = storage_index(dimX, triplets[3 * i]) * dimY +
                storage_index(dimY, triplets[3 * i + 1]);
    }
    stickIndices = keys;
    std::sort(stickIndices.begin(), stickIndices.end());
    stickIndices.erase(std::unique(stickIndices.begin(), stickIndices.end()), stickIndices.end());
    for (long long i = 0; i < numValues; ++i) {
      const int rank = static_cast<int>(
          std::lower_bound(stickIndices.begin(), stickIndices.end(), keys[i]) -
          stickIndices.begin());
      valueIndices[i] = rank * dimZ + storage_index(dimZ, triplets[3 * i + 2]);
    }
  }
}

void check_stick_duplicates(const std::vector<std::vector<int>>& sticksPerRank) {
  std::vector<int> all;
  for (const auto& s : sticksPerRank) all.insert(all.end(), s.begin(), s.end());
  std::sort(all.begin(), all.end());
  if (std::adjacent_find(all.begin(), all.end()) != all.end()) throw DuplicateIndicesError();
}

std::shared_ptr<IndexMaps> make_local_index_maps(SpfftTransformType type, int dimX, int dimY,
                                                 int dimZ, int numLocalElements,
                                                 SpfftIndexFormatType indexFormat,
                                                 const int* indices) {
  if (indexFormat != SPFFT_INDEX_TRIPLETS) throw InternalError();  // parameters.cpp:158-160
  auto m = std::make_shared<IndexMaps>();
  m->type = type;
  m->dimX = dimX;
  m->dimY = dimY;
  m->dimZ = dimZ;
  m->dimXFreq = type == SPFFT_TRANS_R2C ? dimX / 2 + 1 : dimX;
  convert_index_triplets(type == SPFFT_TRANS_R2C, dimX, dimY, dimZ, numLocalElements, indices,
                         m->valueIndices, m->stickIndices);
  m->sticksPerRank.assign(1, m->stickIndices);
  check_stick_duplicates(m->sticksPerRank);
  m->zeroZeroStickIndex = 0;
  for (int key : m->stickIndices) {
    if (key == 0) break;
    ++m->zeroZeroStickIndex;
  }
  m->commRank = 0;
  m->commSize = 1;
  m->numPlanesPerRank.assign(1, dimZ);
  m->planeOffsetPerRank.assign(1, 0);
  m->numGlobalElements = numLocalElements;
  return m;
}

void finish_distributed_index_maps(IndexMaps& m, int commRank,
                                   const std::vector<std::vector<long long>>& countsPerRank,
                                   std::vector<std::vector<int>> sticksPerRank) {
  const int size = static_cast<int>(countsPerRank.size());
  m.commRank = commRank;
  m.commSize = size;
  m.sticksPerRank = std::move(sticksPerRank);
  check_stick_duplicates(m.sticksPerRank);  // parameters.cpp:76
  long long sticksTotal = 0, planesTotal = 0;
  m.numGlobalElements = 0;
  m.numPlanesPerRank.clear();
  m.planeOffsetPerRank.clear();
  for (const auto& c : countsPerRank) {
    // dimensions must match on all ranks (parameters.cpp:95-99)
    if (c[0] != m.dimX || c[1] != m.dimY || c[2] != m.dimZ) throw MPIParameterMismatchError();
    planesTotal += c[3];
    sticksTotal += c[4];
  }
  if (sticksTotal > static_cast<long long>(m.dimX) * m.dimY) throw MPIParameterMismatchError();
  if (planesTotal != m.dimZ) throw MPIParameterMismatchError();
  int offset = 0;
  for (const auto& c : countsPerRank) {
    m.numPlanesPerRank.push_back(static_cast<int>(c[3]));
    m.planeOffsetPerRank.push_back(offset);
    offset += static_cast<int>(c[3]);
    m.numGlobalElements += c[5];
  }
}

ExchangePlan build_exchange_plan(const IndexMaps& m, int log2Vz, int log2Vy, bool fastY) {
  ExchangePlan x;
  const int P = m.commSize, me = m.commRank;
  const int Vz = 1 << log2Vz, Vy = 1 << log2Vy;
  x.commSize = P;
  x.commRank = me;
  x.pitchPerRank.resize(P);
  for (int r = 0; r < P; ++r) {
    const int ns = static_cast<int>(m.sticksPerRank[r].size());
    x.pitchPerRank[r] = (ns + Vz - 1) / Vz * Vz;
  }
  const long long myPitch = x.pitchPerRank[me];
  const long long myPlanes = m.numPlanesPerRank[me];
  x.stickOffset.resize(P);
  x.stickCount.resize(P);
  x.planeOffset.resize(P);
  x.planeCount.resize(P);
  long long q = 0;
  for (int r = 0; r < P; ++r) {
    x.stickOffset[r] = static_cast<long long>(m.planeOffsetPerRank[r]) * myPitch;
    x.stickCount[r] = static_cast<long long>(m.numPlanesPerRank[r]) * myPitch;
    x.planeOffset[r] = q;
    x.planeCount[r] = myPlanes * x.pitchPerRank[r];
    q += x.planeCount[r];
  }
  x.planeSideElements = q;
  if (q > 0x7fffffffLL) throw OverflowError();
  // peer-memory form, backward: where row z of MY sticks lives inside the plane-side buffer of the
  // rank d that owns plane z (block of source `me` inside d's buffer, row z - planeOffset(d))
  if (P <= 255) {
    x.rowRank.assign(m.dimZ, 0);
    x.rowOff.assign(m.dimZ, 0);
    for (int d = 0; d < P; ++d) {
      long long blockOfMe = 0;  // offset of my block inside d's buffer
      for (int r = 0; r < me; ++r)
        blockOfMe += static_cast<long long>(m.numPlanesPerRank[d]) * x.pitchPerRank[r];
      for (int zl = 0; zl < m.numPlanesPerRank[d]; ++zl) {
        const int z = m.planeOffsetPerRank[d] + zl;
        if (z < 0 || z >= m.dimZ) throw InternalError();
        x.rowRank[z] = static_cast<unsigned char>(d);
        x.rowOff[z] = blockOfMe + static_cast<long long>(zl) * myPitch;
      }
    }
  }

  // all sticks of all ranks, ordered by key: (key, rank, local index)
  struct Entry {
    int key, rank, idx;
  };
  std::vector<Entry> all;
  for (int r = 0; r < P; ++r) {
    const auto& v = m.sticksPerRank[r];
    for (int i = 0; i < static_cast<int>(v.size()); ++i) all.push_back(Entry{v[i], r, i});
  }
  std::sort(all.begin(), all.end(), [](const Entry& a, const Entry& b) { return a.key < b.key; });
  x.numXTiles = (m.dimXFreq + Vy - 1) / Vy;
  x.xtStart.assign(x.numXTiles + 1, 0);
  x.stickSlot.resize(all.size());
  x.srcBase.resize(all.size());
  x.srcPitch.resize(all.size());
  for (size_t e = 0; e < all.size(); ++e) {
    const int xx = all[e].key / m.dimY;
    const int yy = all[e].key - xx * m.dimY;
    ++x.xtStart[(xx >> log2Vy) + 1];
    x.stickSlot[e] = (yy << log2Vy) + (xx & (Vy - 1));
    x.srcBase[e] = static_cast<int>(x.planeOffset[all[e].rank] + all[e].idx);
    x.srcPitch[e] = x.pitchPerRank[all[e].rank];
  }
  // peer-memory form, forward: row planeOffset(me) + zl of the owner's plane-major stick buffer
  if (P <= 255) {
    x.stickRank.resize(all.size());
    x.fwdBase.resize(all.size());
    for (size_t e = 0; e < all.size(); ++e) {
      const long long base = static_cast<long long>(m.planeOffsetPerRank[me]) *
                                 x.pitchPerRank[all[e].rank] + all[e].idx;
      if (base + myPlanes * x.pitchPerRank[all[e].rank] > 0x7fffffffLL) throw OverflowError();
      x.stickRank[e] = static_cast<unsigned char>(all[e].rank);
      x.fwdBase[e] = static_cast<int>(base);
    }
  }
  for (int k = 0; k < x.numXTiles; ++k) x.xtStart[k + 1] += x.xtStart[k];

  // single-source tiles + inverse map for the gather-form y kernels
  x.tileBase.assign(x.numXTiles, 0);
  x.tilePitch.assign(x.numXTiles, 0);
  x.tileFwdBase.assign(x.numXTiles, 0);
  for (int t = 0; t < x.numXTiles; ++t) {
    const int e0 = x.xtStart[t], e1 = x.xtStart[t + 1];
    if (e0 == e1) continue;
    bool single = true;
    for (int e = e0 + 1; e < e1 && single; ++e)
      single = all[e].rank == all[e0].rank && all[e].idx == all[e0].idx + (e - e0);
    if (single) {
      x.tileBase[t] = x.srcBase[e0];
      x.tilePitch[t] = x.srcPitch[e0];
      if (!x.fwdBase.empty()) x.tileFwdBase[t] = x.fwdBase[e0];
    }
  }
  // forward visiting order: start at the first x tile whose sticks belong to the next rank
  if (!x.stickRank.empty() && P > 1) {
    const int next = (me + 1) % P;
    for (int t = 0; t < x.numXTiles; ++t) {
      if (x.xtStart[t] < x.xtStart[t + 1] && x.stickRank[x.xtStart[t]] == next) {
        x.fwdTileRotate = t;
        break;
      }
    }
  }
  if (!x.stickRank.empty() && P > 1) {
    std::vector<std::vector<int>> byOwner(P);
    std::vector<int> empty;
    for (int t = 0; t < x.numXTiles; ++t) {
      if (x.xtStart[t] < x.xtStart[t + 1])
        byOwner[x.stickRank[x.xtStart[t]]].push_back(t);
      else
        empty.push_back(t);
    }
    size_t longest = 0;
    for (const auto& v : byOwner) longest = std::max(longest, v.size());
    x.fwdTileOrder.reserve(x.numXTiles);
    for (size_t i = 0; i < longest; ++i)
      for (int k = 1; k <= P; ++k) {
        const auto& v = byOwner[(me + k) % P];
        if (i < v.size()) x.fwdTileOrder.push_back(v[i]);
      }
    x.fwdTileOrder.insert(x.fwdTileOrder.end(), empty.begin(), empty.end());
  }
  if (fastY && !all.empty()) {
    const int ny = m.dimY;
    const int vpt = fast_path_values_per_thread(ny);
    const int T = ny / vpt;
    const size_t perTile = static_cast<size_t>(Vy) * T * vpt;
    x.yInv.assign(perTile * x.numXTiles, 0xFFFF);
    for (size_t e = 0; e < all.size(); ++e) {
      const int xx = all[e].key / ny;
      const int yy = all[e].key - xx * ny;
      const int tile = xx >> log2Vy;
      const int lane = xx & (Vy - 1);
      const size_t tid = static_cast<size_t>(lane) * T + (yy % T);
      x.yInv[perTile * tile + tid * vpt + yy / T] = static_cast<unsigned short>(e - x.xtStart[tile]);
    }
  }
  return x;
}

TileMaps build_tile_maps(const IndexMaps& maps, int log2Vz, int log2Vy, bool fastZ, bool fastY) {
  TileMaps t;
  t.log2Vz = log2Vz;
  t.log2Vy = log2Vy;
  const int Vz = 1 << log2Vz, Vy = 1 << log2Vy;
  const int ns = maps.num_sticks();
  const int ne = maps.num_values();
  const int nz = maps.dimZ;
  t.numStickTiles = (ns + Vz - 1) / Vz;
  t.pitch = t.numStickTiles * Vz;

  // ---- z stage: counting sort of the sparse entries by stick tile, keeping user order inside
  t.tileStart.assign(t.numStickTiles + 1, 0);
  for (int i = 0; i < ne; ++i) {
    const int stick = maps.valueIndices[i] / nz;
    ++t.tileStart[(stick >> log2Vz) + 1];
  }
  for (int k = 0; k < t.numStickTiles; ++k) t.tileStart[k + 1] += t.tileStart[k];
  t.entrySrc.resize(ne);
  t.entrySlot.resize(ne);
  {
    std::vector<int> cursor(t.tileStart.begin(), t.tileStart.end() - 1);
    bool identity = true;
    for (int i = 0; i < ne; ++i) {
      const int vi = maps.valueIndices[i];
      const int stick = vi / nz;
      const int z = vi - stick * nz;
      const int p = cursor[stick >> log2Vz]++;
      t.entrySrc[p] = i;
      t.entrySlot[p] = (z << log2Vz) + (stick & (Vz - 1));
      identity = identity && (p == i);
    }
    t.identityOrder = identity;
  }

  // ---- duplicates (legal for the reference, Appendix B.8 of SURVEY.md): last writer wins
  {
    std::vector<int> stamp(static_cast<size_t>(nz) * Vz, -1);
    std::vector<char> dropped;
    for (int tile = 0; tile < t.numStickTiles; ++tile) {
      for (int p = t.tileStart[tile + 1] - 1; p >= t.tileStart[tile]; --p) {
        int& s = stamp[t.entrySlot[p]];
        if (s == tile) {
          if (dropped.empty()) dropped.assign(ne, 0);
          dropped[p] = 1;
        } else {
          s = tile;
        }
      }
    }
    if (!dropped.empty()) {
      t.hasDuplicates = true;
      t.bwdTileStart.assign(t.numStickTiles + 1, 0);
      for (int tile = 0; tile < t.numStickTiles; ++tile) {
        t.bwdTileStart[tile] = static_cast<int>(t.bwdEntrySrc.size());
        for (int p = t.tileStart[tile]; p < t.tileStart[tile + 1]; ++p) {
          if (!dropped[p]) {
            t.bwdEntrySrc.push_back(t.entrySrc[p]);
            t.bwdEntrySlot.push_back(t.entrySlot[p]);
          }
        }
      }
      t.bwdTileStart[t.numStickTiles] = static_cast<int>(t.bwdEntrySrc.size());
    }
  }

  if (maps.type == SPFFT_TRANS_R2C && maps.zeroZeroStickIndex < ns) {
    t.symTile = maps.zeroZeroStickIndex >> log2Vz;
    t.symLane = maps.zeroZeroStickIndex & (Vz - 1);
  }

  // ---- y stage
  t.numXTiles = (maps.dimXFreq + Vy - 1) / Vy;
  t.xtStart.assign(t.numXTiles + 1, 0);
  t.stickSlot.resize(ns);
  for (int s = 0; s < ns; ++s) {
    const int key = maps.stickIndices[s];
    const int x = key / maps.dimY;
    const int y = key - x * maps.dimY;
    ++t.xtStart[(x >> log2Vy) + 1];
    t.stickSlot[s] = (y << log2Vy) + (x & (Vy - 1));
  }
  for (int k = 0; k < t.numXTiles; ++k) t.xtStart[k + 1] += t.xtStart[k];

  // ---- inverse maps of the register-FFT kernels
  if (fastZ && t.identityOrder && !t.hasDuplicates && ne > 0) {
    const int vpt = fast_path_values_per_thread(nz);
    const int T = nz / vpt;
    const size_t perTile = static_cast<size_t>(Vz) * T * vpt;
    t.zInv.assign(perTile * t.numStickTiles, 0xFFFF);
    for (int p = 0; p < ne; ++p) {
      const int vi = maps.valueIndices[p];  // identity order: entry p == value p
      const int stick = vi / nz;
      const int z = vi - stick * nz;
      const int tile = stick >> log2Vz;
      const int lane = stick & (Vz - 1);
      const size_t tid = static_cast<size_t>(lane) * T + (z % T);
      t.zInv[perTile * tile + tid * vpt + z / T] = static_cast<unsigned short>(p - t.tileStart[tile]);
    }
  }
  if (fastY && ns > 0) {
    const int ny = maps.dimY;
    const int vpt = fast_path_values_per_thread(ny);
    const int T = ny / vpt;
    const size_t perTile = static_cast<size_t>(Vy) * T * vpt;
    t.yInv.assign(perTile * t.numXTiles, 0xFFFF);
    for (int s = 0; s < ns; ++s) {
      const int key = maps.stickIndices[s];
      const int x = key / ny;
      const int y = key - x * ny;
      const int tile = x >> log2Vy;
      const int lane = x & (Vy - 1);
      const size_t tid = static_cast<size_t>(lane) * T + (y % T);
      t.yInv[perTile * tile + tid * vpt + y / T] = static_cast<unsigned short>(s - t.xtStart[tile]);
    }
  }
  return t;
}

sb::RadixPlan make_radix_plan(int n) {
  sb::RadixPlan rp;
  rp.n = n;
  rp.numPasses = 0;
  for (int& r : rp.radix) r = 1;
  if (n <= 1) return rp;
  int m = n;
  auto push = [&](int r) {
    if (rp.numPasses >= sb::kMaxPasses) throw InvalidParameterError();
    rp.radix[rp.numPasses++] = r;
  };
  while (m % 8 == 0) {
    push(8);
    m /= 8;
  }
  if (m % 4 == 0) {
    push(4);
    m /= 4;
  }
  if (m % 2 == 0) {
    push(2);
    m /= 2;
  }
  while (m % 3 == 0) {
    push(3);
    m /= 3;
  }
  while (m % 5 == 0) {
    push(5);
    m /= 5;
  }
  for (int p = 7; static_cast<long long>(p) * p <= m; p += 2) {
    while (m % p == 0) {
      push(p);
      m /= p;
    }
  }
  if (m > 1) push(m);
  return rp;
}

template <typename T>
std::vector<sb::cx<T>> make_roots(int n) {
  std::vector<sb::cx<T>> w(static_cast<size_t>(n > 0 ? n : 1));
  const long double twoPi = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < n; ++k) {
    // exact octant symmetries are not needed: long double keeps the rounding error below T's ulp
    const long double a = twoPi * static_cast<long double>(k) / static_cast<long double>(n);
    w[k].x = static_cast<T>(cosl(a));
    w[k].y = static_cast<T>(-sinl(a));
  }
  if (n <= 0) {
    w[0].x = T(1);
    w[0].y = T(0);
  }
  return w;
}
template std::vector<sb::cx<float>> make_roots<float>(int);
template std::vector<sb::cx<double>> make_roots<double>(int);

bool fast_path_length(int n, int complexBytes) {
  if (n % 5 == 0) {  // 5 * 2^k: 160, 320, 640 (fast3_stage_kernels.hpp, 40 values per thread)
    const int m = n / 5;
    return m >= 32 && m <= 128 && (m & (m - 1)) == 0;
  }
  if (n % 3 == 0) {  // 3 * 2^k, sub-transforms of 32 .. 256 (fast3_stage_kernels.hpp)
    const int m = n / 3;
    return m >= 32 && m <= 256 && (m & (m - 1)) == 0;
  }
  if (n < 32 || (n & (n - 1)) != 0) return false;
  const int lanes = 1 << fast_path_log2_lanes(complexBytes, n);
  return lanes * (n / 8) <= 1024;  // threads per CTA
}

int fast_path_log2_lanes_x(int n) {
#ifdef SB_LOG2VX
  (void)n;
  return SB_LOG2VX;
#else
  return sb::x_lanes_log2(n);
#endif
}

int fast_path_values_per_thread(int n) { return n % 5 == 0 ? 40 : (n % 3 == 0 ? 24 : 8); }

int fast_path_log2_lanes(int complexBytes, int n) {
  const int common = complexBytes == 16 ? sb::FastLanes<double>::log2V : sb::FastLanes<float>::log2V;
  return (n > 0 && sb::is_fast3_length(n)) ? sb::fast3_lanes_log2(n, complexBytes, common) : common;
}

template <typename T>
std::vector<sb::cx<T>> make_fast_twiddles(int n) {
  std::vector<sb::cx<T>> w;
  const long double twoPiL = 6.283185307179586476925286766559005768L;
  if (n > 0 && (n % 3 == 0 || n % 5 == 0)) {
    // N = G*M: [r-1][k] = exp(-2*pi*i*r*k/N), r = 1 .. G-1, k < M, then the stage twiddles of length M
    const int groups = n % 5 == 0 ? 5 : 3;
    const int m = n / groups;
    for (int r = 1; r < groups; ++r) {
      for (int k = 0; k < m; ++k) {
        const long double a = twoPiL * static_cast<long double>(r * k) / static_cast<long double>(n);
        sb::cx<T> v;
        v.x = static_cast<T>(cosl(a));
        v.y = static_cast<T>(-sinl(a));
        w.push_back(v);
      }
    }
    const std::vector<sb::cx<T>> sub = make_fast_twiddles<T>(m);
    w.insert(w.end(), sub.begin(), sub.end());
    return w;
  }
  int log2n = 0;
  while ((1 << log2n) < n) ++log2n;
  const int r0 = (log2n % 3 == 0) ? 8 : ((log2n % 3 == 1) ? 2 : 4);
  const int stages = (log2n + 2) / 3;
  const long double twoPi = 6.283185307179586476925286766559005768L;
  int ns = r0;
  for (int s = 1; s < stages; ++s) {
    for (int r = 1; r < 8; ++r) {
      for (int k = 0; k < ns; ++k) {
        const long double a = twoPi * static_cast<long double>(r * k) / static_cast<long double>(8 * ns);
        sb::cx<T> v;
        v.x = static_cast<T>(cosl(a));
        v.y = static_cast<T>(-sinl(a));
        w.push_back(v);
      }
    }
    ns *= 8;
  }
  if (w.empty()) w.push_back(sb::mk<T>(T(1), T(0)));
  return w;
}
template std::vector<sb::cx<float>> make_fast_twiddles<float>(int);
template std::vector<sb::cx<double>> make_fast_twiddles<double>(int);

int choose_log2_lanes(int n, int complexBytes, long long smemLimit) {
  int log2V = 0;
  while ((complexBytes << (log2V + 1)) <= 128) ++log2V;  // at most 128 bytes per tile row
  for (; log2V >= 0; --log2V) {
    const long long bytes = 2LL * n * (static_cast<long long>(complexBytes) << log2V);
    if (bytes <= smemLimit) return log2V;
  }
  return -1;
}

}  // namespace b200
}  // namespace spfft
