// api.cpp -- the drop-in boundary: spfft::Grid / Transform / GridFloat / TransformFloat,
// multi_transform_* and every extern "C" entry point of include/spfft/*.h, instantiated for both
// precisions from api_impl.inc.
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "index_plan.hpp"
#include "launch.h"
#include "spfft/spfft.h"
#include "spfft/spfft.hpp"
#include "transform_engine.hpp"

namespace spfft {
namespace b200 {
// SpfftB200Comm handles are heap allocated shared_ptr<Communicator>
using CommHandle = std::shared_ptr<Communicator>;
static std::shared_ptr<Communicator> communicator_from_handle(SpfftB200Comm comm) {
  if (!comm) throw InvalidParameterError();
  return *static_cast<CommHandle*>(comm);
}
}  // namespace b200
}  // namespace spfft

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

SpfftError spfft_b200_nccl_unique_id(char* id128) {
  try {
    if (!id128) throw spfft::InvalidParameterError();
    const spfft::b200::NcclUniqueId id = spfft::b200::Communicator::unique_id();
    std::memcpy(id128, id.internal, 128);
  } catch (const spfft::GenericError& e) {
    return e.error_code();
  } catch (...) {
    return SPFFT_UNKNOWN_ERROR;
  }
  return SPFFT_SUCCESS;
}

SpfftError spfft_b200_comm_create(SpfftB200Comm* comm, int numRanks, int rank, const char* id128) {
  try {
    if (!comm || !id128) throw spfft::InvalidParameterError();
    spfft::b200::NcclUniqueId id;
    std::memcpy(id.internal, id128, 128);
    *comm = new spfft::b200::CommHandle(
        std::make_shared<spfft::b200::Communicator>(numRanks, rank, id));
  } catch (const spfft::GenericError& e) {
    return e.error_code();
  } catch (...) {
    return SPFFT_UNKNOWN_ERROR;
  }
  return SPFFT_SUCCESS;
}

SpfftError spfft_b200_comm_destroy(SpfftB200Comm comm) {
  if (!comm) return SPFFT_INVALID_HANDLE_ERROR;
  delete static_cast<spfft::b200::CommHandle*>(comm);  // grids keep the communicator alive
  return SPFFT_SUCCESS;
}

SpfftError spfft_b200_comm_size(SpfftB200Comm comm, int* size) {
  if (!comm) return SPFFT_INVALID_HANDLE_ERROR;
  *size = (*static_cast<spfft::b200::CommHandle*>(comm))->size();
  return SPFFT_SUCCESS;
}

SpfftError spfft_b200_comm_rank(SpfftB200Comm comm, int* rank) {
  if (!comm) return SPFFT_INVALID_HANDLE_ERROR;
  *rank = (*static_cast<spfft::b200::CommHandle*>(comm))->rank();
  return SPFFT_SUCCESS;
}

// the exchange plan rank `commRank` would build (host only)
static spfft::b200::ExchangePlan host_exchange_plan(int transformType, int isFloat, int dimX, int dimY,
                                                    int dimZ, int commSize, int commRank,
                                                    const int* numSticksPerRank,
                                                    const int* sticksAllRanks, const int* planesPerRank,
                                                    int* log2Vy) {
  using namespace spfft::b200;
  if (commSize < 1 || commRank < 0 || commRank >= commSize || dimX <= 0 || dimY <= 0 || dimZ <= 0)
    throw spfft::InvalidParameterError();
  IndexMaps m;
  m.type = static_cast<SpfftTransformType>(transformType);
  m.dimX = dimX;
  m.dimY = dimY;
  m.dimZ = dimZ;
  m.dimXFreq = m.type == SPFFT_TRANS_R2C ? dimX / 2 + 1 : dimX;
  std::vector<std::vector<long long>> counts(commSize, std::vector<long long>(6, 0));
  std::vector<std::vector<int>> sticks(commSize);
  const int* cursor = sticksAllRanks;
  for (int r = 0; r < commSize; ++r) {
    counts[r][0] = dimX;
    counts[r][1] = dimY;
    counts[r][2] = dimZ;
    counts[r][3] = planesPerRank[r];
    counts[r][4] = numSticksPerRank[r];
    sticks[r].assign(cursor, cursor + numSticksPerRank[r]);
    cursor += numSticksPerRank[r];
  }
  m.stickIndices = sticks[commRank];
  finish_distributed_index_maps(m, commRank, counts, std::move(sticks));
  AxisPlans ax;
  bool fx, fy, fz;
  choose_tile_lanes(m, isFloat ? 8 : 16, 232448, ax, fx, fy, fz);
  if (log2Vy) *log2Vy = ax.log2Vy;
  return build_exchange_plan(m, ax.log2Vz, ax.log2Vy);
}

SpfftError spfft_b200_exchange_plan(int transformType, int isFloat, int dimX, int dimY, int dimZ,
                                    int commSize, int commRank, const int* numSticksPerRank,
                                    const int* sticksAllRanks, const int* planesPerRank,
                                    int* pitchPerRank, long long* stickOffset,
                                    long long* stickCount, long long* planeOffset,
                                    long long* planeCount, int* numXTiles, int* log2Vy, int* xtStart,
                                    int* stickSlot, int* srcBase, int* srcPitch) {
  try {
    const spfft::b200::ExchangePlan x =
        host_exchange_plan(transformType, isFloat, dimX, dimY, dimZ, commSize, commRank,
                           numSticksPerRank, sticksAllRanks, planesPerRank, log2Vy);
    std::copy(x.pitchPerRank.begin(), x.pitchPerRank.end(), pitchPerRank);
    std::copy(x.stickOffset.begin(), x.stickOffset.end(), stickOffset);
    std::copy(x.stickCount.begin(), x.stickCount.end(), stickCount);
    std::copy(x.planeOffset.begin(), x.planeOffset.end(), planeOffset);
    std::copy(x.planeCount.begin(), x.planeCount.end(), planeCount);
    *numXTiles = x.numXTiles;
    std::copy(x.xtStart.begin(), x.xtStart.end(), xtStart);
    std::copy(x.stickSlot.begin(), x.stickSlot.end(), stickSlot);
    std::copy(x.srcBase.begin(), x.srcBase.end(), srcBase);
    std::copy(x.srcPitch.begin(), x.srcPitch.end(), srcPitch);
  } catch (const spfft::GenericError& e) {
    return e.error_code();
  } catch (...) {
    return SPFFT_UNKNOWN_ERROR;
  }
  return SPFFT_SUCCESS;
}

SpfftError spfft_b200_exchange_plan_peer(int transformType, int isFloat, int dimX, int dimY, int dimZ,
                                         int commSize, int commRank, const int* numSticksPerRank,
                                         const int* sticksAllRanks, const int* planesPerRank,
                                         int* rowRank, long long* rowOff, int* stickRank, int* fwdBase,
                                         int* fwdTileRotate) {
  try {
    const spfft::b200::ExchangePlan x =
        host_exchange_plan(transformType, isFloat, dimX, dimY, dimZ, commSize, commRank,
                           numSticksPerRank, sticksAllRanks, planesPerRank, nullptr);
    if (x.rowRank.empty() && dimZ > 0) throw spfft::InvalidParameterError();  // more than 255 ranks
    std::copy(x.rowRank.begin(), x.rowRank.end(), rowRank);
    std::copy(x.rowOff.begin(), x.rowOff.end(), rowOff);
    std::copy(x.stickRank.begin(), x.stickRank.end(), stickRank);
    std::copy(x.fwdBase.begin(), x.fwdBase.end(), fwdBase);
    *fwdTileRotate = x.fwdTileRotate;
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
