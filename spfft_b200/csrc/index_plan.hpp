// index_plan.hpp -- plan-time (host) integer work: the reference-defined index maps and the tile
// maps the B200 kernels consume.
//
// Reference-defined (bit-exact contract):
//   convert_index_triplets   src/compression/indices.hpp:120-186
//   check_stick_duplicates   src/compression/indices.hpp:105-117
//   Parameters (local)       src/parameters/parameters.cpp:143-180
// Derived (ours): per-stick-tile sparse entry lists, per-x-tile stick ranges, radix plans,
// twiddle tables.
#pragma once
#include <memory>
#include <vector>

#include "fft_tile.hpp"
#include "spfft/types.h"

namespace spfft {
namespace b200 {

// Immutable per-transform parameters, shared by clones (the reference's `Parameters`).
struct IndexMaps {
  SpfftTransformType type = SPFFT_TRANS_C2C;
  int dimX = 0, dimY = 0, dimZ = 0, dimXFreq = 0;
  std::vector<int> valueIndices;  // [Ne]  stick*dimZ + z      (reference freqValueIndices_)
  std::vector<int> stickIndices;  // [Ns]  x*dimY + y, ascending (reference stickIndicesPerRank_[rank])
  int zeroZeroStickIndex = 0;     // position of key 0 or Ns (parameters.cpp:173-179)
  // decomposition (single rank: all planes local)
  int commRank = 0, commSize = 1;
  std::vector<int> numPlanesPerRank;       // [size]
  std::vector<int> planeOffsetPerRank;     // [size]
  std::vector<std::vector<int>> sticksPerRank;  // [size][Ns(r)]; sticksPerRank[rank] == stickIndices
  long long numGlobalElements = 0;

  int num_sticks() const { return static_cast<int>(stickIndices.size()); }
  int num_values() const { return static_cast<int>(valueIndices.size()); }
  int local_planes() const { return numPlanesPerRank[commRank]; }
  int local_plane_offset() const { return planeOffsetPerRank[commRank]; }
};

// Throws InvalidParameterError / InvalidIndicesError exactly where the reference does.
void convert_index_triplets(bool hermitianSymmetry, int dimX, int dimY, int dimZ, int numValues,
                            const int* triplets, std::vector<int>& valueIndices,
                            std::vector<int>& stickIndices);

// Throws DuplicateIndicesError if a stick key appears twice over all ranks.
void check_stick_duplicates(const std::vector<std::vector<int>>& sticksPerRank);

// Single-rank parameters (parameters.cpp:143-180).
std::shared_ptr<IndexMaps> make_local_index_maps(SpfftTransformType type, int dimX, int dimY,
                                                 int dimZ, int numLocalElements,
                                                 SpfftIndexFormatType indexFormat,
                                                 const int* indices);

// Distributed parameters (parameters.cpp:43-140). `countsPerRank` holds, for every rank, the six
// values the reference all-gathers (dimX, dimY, dimZ, numLocalXYPlanes, numLocalZSticks,
// numLocalElements); `sticksPerRank` the all-gathered stick lists. Throws
// MPIParameterMismatchError / DuplicateIndicesError exactly where the reference does.
// `local` must already hold this rank's valueIndices / stickIndices.
void finish_distributed_index_maps(IndexMaps& local, int commRank,
                                   const std::vector<std::vector<long long>>& countsPerRank,
                                   std::vector<std::vector<int>> sticksPerRank);

// The stick <-> slab exchange of one rank (the compact wire format of
// transpose_mpi_compact_buffered_host.cpp:83-175, re-blocked for the plane-major stick buffer):
//   stick side  A : [dimZ][pitch(rank)]            rows of destination r are one contiguous block
//   plane side  Q : for every source r a block [localPlanes][pitch(r)], blocks back to back
// backward: send A blocks -> receive into Q; forward: send Q blocks -> receive into A.
struct ExchangePlan {
  int commSize = 1, commRank = 0;
  std::vector<int> pitchPerRank;                 // stick row pitch of every rank (tile padded)
  std::vector<long long> stickOffset, stickCount;  // per peer: block inside A (elements)
  std::vector<long long> planeOffset, planeCount;  // per peer: block inside Q (elements)
  long long planeSideElements = 0;               // size of Q
  // y stage tables over ALL sticks of all ranks, sorted by x*dimY + y
  std::vector<int> xtStart;    // [numXTiles+1]
  std::vector<int> stickSlot;  // [NsTotal] y*Vy + (x mod Vy)
  std::vector<int> srcBase;    // [NsTotal] offset of the stick's value at local plane 0 inside Q
  std::vector<int> srcPitch;   // [NsTotal] distance between consecutive local planes
  // x tiles whose sticks all come from ONE rank (the usual case: ranks own contiguous x ranges)
  // are contiguous inside that rank's block: first stick at tileBase + plane*tilePitch, so the
  // inverse-map (gather) form of the y kernels applies; tilePitch == 0 marks a mixed tile
  std::vector<int> tileBase, tilePitch;  // [numXTiles]
  std::vector<unsigned short> yInv;      // inverse map over the global sorted stick list (fast y only)
  int numXTiles = 0;
  // Peer-memory form of the exchange (the kernels store straight into the destination rank's
  // buffer over NVLink, no send/recv):
  //   backward: row z of this rank's sticks -> buffer Q of rank rowRank[z] at element rowOff[z]
  //   forward : stick e of local plane zl   -> buffer A of rank stickRank[e] at element
  //             fwdBase[e] + zl*srcPitch[e]; single-source tiles at tileFwdBase[t] + zl*tilePitch[t]
  std::vector<unsigned char> rowRank;   // [dimZ]
  std::vector<long long> rowOff;        // [dimZ]
  std::vector<unsigned char> stickRank; // [NsTotal]
  std::vector<int> fwdBase;             // [NsTotal]
  std::vector<int> tileFwdBase;         // [numXTiles]
  int fwdTileRotate = 0;                // first x tile owned by the next rank (forward visiting order)
  // forward visiting order of the x tiles: destination ranks interleaved round robin, starting with the
  // next rank (a tile belongs to the owner of its first stick; tiles without sticks go last)
  std::vector<int> fwdTileOrder;        // [numXTiles]
};

ExchangePlan build_exchange_plan(const IndexMaps& maps, int log2Vz, int log2Vy, bool fastY = false);

// What the stage kernels read. All arrays are host vectors here; TransformEngine uploads them.
struct TileMaps {
  int log2Vz = 0, log2Vy = 0;
  // z stage: stick tiles of Vz consecutive sticks
  int numStickTiles = 0;
  int pitch = 0;                   // numStickTiles * Vz
  bool identityOrder = true;       // entrySrc[p] == p for all p -> kernels skip the indirection
  std::vector<int> tileStart;      // [numStickTiles+1]
  std::vector<int> entrySrc;       // [Ne] position in the user's value array, grouped by tile
  std::vector<int> entrySlot;      // [Ne] z*Vz + (stick mod Vz)
  // backward-only variant when the user passed duplicate triplets: only the LAST duplicate
  // scatters (what the reference host loop leaves behind, compression_host.hpp:88-91)
  bool hasDuplicates = false;
  std::vector<int> bwdTileStart, bwdEntrySrc, bwdEntrySlot;
  // (x=0,y=0) stick for the R2C stick symmetry
  int symTile = -1, symLane = -1;
  // y stage: x tiles of Vy consecutive x columns; sticks are sorted by x so each tile owns a range
  int numXTiles = 0;
  std::vector<int> xtStart;    // [numXTiles+1] (stick index ranges)
  std::vector<int> stickSlot;  // [Ns] y*Vy + (x mod Vy)
  // inverse maps of the register-FFT kernels (fast_stage_kernels.hpp, "gather form"), built only
  // for axes on the fast path: [tile][thread = lane*T + j][m] -> offset of element j + T*m of
  // the lane-th sequence from the tile's first entry / stick, 0xFFFF = absent.
  // zInv needs the values in stick order (identityOrder) and no duplicates.
  std::vector<unsigned short> zInv, yInv;
};

TileMaps build_tile_maps(const IndexMaps& maps, int log2Vz, int log2Vy, bool fastZ = false,
                         bool fastY = false);

// Radix schedule for the Stockham tile FFT: 8s, then 4 / 2, then 3s, 5s, then remaining primes.
sb::RadixPlan make_radix_plan(int n);

// Forward roots of unity exp(-2*pi*i*k/n), k = 0..n-1, rounded from long double.
template <typename T>
std::vector<sb::cx<T>> make_roots(int n);

// Register-FFT fast path: power-of-two lengths whose tile (8 lanes, n/8 threads per lane) fits one
// CTA (fast_fft.hpp), and 3 * 2^k lengths 96 .. 768 (fast3_stage_kernels.hpp, n/24 threads per lane).
bool fast_path_length(int n, int complexBytes);
// complex values a thread of the fast path holds (8, or 24 for 3 * 2^k): the inverse maps are laid
// out [tile][thread = lane*(n/vpt) + j][vpt]
int fast_path_values_per_thread(int n);
int fast_path_log2_lanes(int complexBytes, int n = 0);  // z / y tiles of an axis of length n on the fast path
int fast_path_log2_lanes_x(int n);  // log2 rows per tile of the stand-alone x stage kernels of length n
// Stage twiddles of the register FFT: for every stage s >= 1 (radix 8, stride ns) the entries
// [r-1][k] = exp(-2*pi*i*r*k/(8*ns)), r = 1..7, k < ns; rounded from long double.
template <typename T>
std::vector<sb::cx<T>> make_fast_twiddles(int n);

// Largest lane count (power of two, at most 128 bytes of complex<T> per tile row) such that two
// tile buffers of n rows fit into smemLimit bytes. Returns -1 if even one lane does not fit.
int choose_log2_lanes(int n, int complexBytes, long long smemLimit);

}  // namespace b200
}  // namespace spfft
