// transform_engine.cpp -- see transform_engine.hpp.
#include "transform_engine.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "launch.h"

namespace spfft {
namespace b200 {

namespace {

inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// widest stick tile any stage kernel uses; the stick pitch is rounded up to a tile
constexpr int kMaxTileLanes = 32;

template <typename T>
struct Launch;
template <>
struct Launch<double> {
  static int zb(int f, const sb::ZArgs<double>& a, const sb::BandTable<double>& t, int n, void* s) { return sb_launch_z_bands_f64(f, &a, &t, n, s); }
  static int yb(int f, const sb::YArgs<double>& a, const sb::BandTable<double>& t, int n, void* s) { return sb_launch_y_bands_f64(f, &a, &t, n, s); }
  static int xb(int f, const sb::XArgs<double>& a, const sb::BandTable<double>& t, int n, void* s) { return sb_launch_x_bands_f64(f, &a, &t, n, s); }
  static int z(int f, const sb::ZArgs<double>& a, void* s) { return sb_launch_z_f64(f, &a, s); }
  static int y(int f, const sb::YArgs<double>& a, void* s) { return sb_launch_y_f64(f, &a, s); }
  static int x(int f, const sb::XArgs<double>& a, void* s) { return sb_launch_x_f64(f, &a, s); }
  static int wxy(int f, const sb::XYArgs<double>& a, void* s) { return sb_launch_wxy_f64(f, &a, s); }
  static int wz(int f, const sb::ZArgs<double>& a, void* s) { return sb_launch_wz_f64(f, &a, s); }
};
template <>
struct Launch<float> {
  static int zb(int f, const sb::ZArgs<float>& a, const sb::BandTable<float>& t, int n, void* s) { return sb_launch_z_bands_f32(f, &a, &t, n, s); }
  static int yb(int f, const sb::YArgs<float>& a, const sb::BandTable<float>& t, int n, void* s) { return sb_launch_y_bands_f32(f, &a, &t, n, s); }
  static int xb(int f, const sb::XArgs<float>& a, const sb::BandTable<float>& t, int n, void* s) { return sb_launch_x_bands_f32(f, &a, &t, n, s); }
  static int z(int f, const sb::ZArgs<float>& a, void* s) { return sb_launch_z_f32(f, &a, s); }
  static int y(int f, const sb::YArgs<float>& a, void* s) { return sb_launch_y_f32(f, &a, s); }
  static int x(int f, const sb::XArgs<float>& a, void* s) { return sb_launch_x_f32(f, &a, s); }
  static int wxy(int f, const sb::XYArgs<float>& a, void* s) { return sb_launch_wxy_f32(f, &a, s); }
  static int wz(int f, const sb::ZArgs<float>& a, void* s) { return sb_launch_wz_f32(f, &a, s); }
};

inline void check_launch(int err) {
  if (err != 0) {
    cudaGetLastError();
    check_gpu(static_cast<cudaError_t>(err));
  }
}

template <typename U>
const U* upload(std::vector<DeviceBuffer>& storage, size_t& total, const std::vector<U>& host) {
  if (host.empty()) return nullptr;
  storage.emplace_back(host.size() * sizeof(U));
  total += host.size() * sizeof(U);
  check_gpu(cudaMemcpy(storage.back().get(), host.data(), host.size() * sizeof(U),
                       cudaMemcpyHostToDevice));
  return storage.back().template as<const U>();
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// GridResources
// ---------------------------------------------------------------------------------------------
template <typename T>
size_t GridResources<T>::stick_capacity(int maxDimZ, int maxSticks) {
  return static_cast<size_t>(maxDimZ) * round_up(static_cast<size_t>(maxSticks), kMaxTileLanes);
}

template <typename T>
GridResources<T>::GridResources(int maxDimX, int maxDimY, int maxDimZ, int maxNumLocalZSticks,
                                SpfftProcessingUnitType processingUnit, int maxNumThreads)
    : maxDimX_(maxDimX),
      maxDimY_(maxDimY),
      maxDimZ_(maxDimZ),
      maxSticks_(maxNumLocalZSticks),
      maxPlanes_(maxDimZ),
      pu_(processingUnit),
      numThreads_(maxNumThreads) {
  // grid_internal.cpp:60-67
  if (maxDimX <= 0 || maxDimY <= 0 || maxDimZ <= 0 || maxNumLocalZSticks < 0)
    throw InvalidParameterError();
  if (!(processingUnit & (SPFFT_PU_HOST | SPFFT_PU_GPU))) throw InvalidParameterError();
  // This build executes on the GPU only. A grid / transform requested for SPFFT_PU_HOST is served by the same
  // device kernels with its data staged through pinned host memory (callers written against the reference's host
  // path, e.g. its examples, keep working); the device buffers are therefore allocated for every grid.
  if (numThreads_ < 1) numThreads_ = 1;  // host threads are not used; kept for the getter
  check_gpu(cudaGetDevice(&deviceId_));  // grid_internal.cpp:83
  allocate();
}

template <typename T>
GridResources<T>::GridResources(int maxDimX, int maxDimY, int maxDimZ, int maxNumLocalZSticks,
                                int maxLocalZLength, SpfftProcessingUnitType processingUnit,
                                int maxNumThreads, std::shared_ptr<Communicator> comm,
                                SpfftExchangeType exchangeType)
    : maxDimX_(maxDimX),
      maxDimY_(maxDimY),
      maxDimZ_(maxDimZ),
      maxSticks_(maxNumLocalZSticks),
      maxPlanes_(maxLocalZLength),
      pu_(processingUnit),
      numThreads_(maxNumThreads),
      comm_(std::move(comm)),
      exchangeType_(exchangeType) {
  // grid_internal.cpp:121-148
  if (!comm_) throw InvalidParameterError();
  if (static_cast<long long>(maxDimX) * maxDimY * maxLocalZLength > 0x7fffffffLL) throw OverflowError();
  if (static_cast<long long>(maxNumLocalZSticks) * maxDimZ > 0x7fffffffLL) throw OverflowError();
  if (maxDimX <= 0 || maxDimY <= 0 || maxDimZ <= 0 || maxNumLocalZSticks < 0 || maxLocalZLength < 0)
    throw InvalidParameterError();
  if (!(processingUnit & (SPFFT_PU_HOST | SPFFT_PU_GPU))) throw InvalidParameterError();
  if (exchangeType < SPFFT_EXCH_DEFAULT || exchangeType > SPFFT_EXCH_UNBUFFERED)
    throw InvalidParameterError();
  if (numThreads_ < 1) numThreads_ = 1;
  deviceId_ = comm_->device_id();
  // parameters must agree on all ranks (grid_internal.cpp:150-168)
  {
    const int mine[2] = {static_cast<int>(exchangeType), static_cast<int>(processingUnit)};
    const std::vector<int> all = comm_->all_gather_ints(mine, 2);
    for (int r = 0; r < comm_->size(); ++r)
      if (all[2 * r] != mine[0] || all[2 * r + 1] != mine[1]) throw MPIParameterMismatchError();
  }
  if (exchangeType_ == SPFFT_EXCH_DEFAULT) exchangeType_ = SPFFT_EXCH_COMPACT_BUFFERED;
  DeviceGuard guard(deviceId_);
  allocate();
  map_peers();
}

template <typename T>
GridResources<T>::GridResources(const GridResources& o)
    : maxDimX_(o.maxDimX_),
      maxDimY_(o.maxDimY_),
      maxDimZ_(o.maxDimZ_),
      maxSticks_(o.maxSticks_),
      maxPlanes_(o.maxPlanes_),
      pu_(o.pu_),
      deviceId_(o.deviceId_),
      numThreads_(o.numThreads_),
      comm_(o.comm_),
      exchangeType_(o.exchangeType_) {
  DeviceGuard guard(deviceId_);
  allocate();
  map_peers();  // collective for distributed grids, like every grid / transform creation
}

template <typename T>
void GridResources<T>::allocate() {
  // local slab volume (all planes for a local grid)
  const size_t volume =
      static_cast<size_t>(maxDimX_) * static_cast<size_t>(maxDimY_) * static_cast<size_t>(maxPlanes_);
  const size_t c = 2 * sizeof(T);
  const size_t stickElems = stick_capacity(maxDimZ_, maxSticks_);
  // A: plane-major sticks [z][pitch]; also the real space domain of R2C transforms
  a_.allocate(std::max(c * stickElems, sizeof(T) * volume));
  // B: xy planes / complex space domain / staged compressed values (Ne <= Ns*Nz, <= volume locally)
  b_.allocate(c * std::max(volume, local() ? size_t(0) : stickElems));
  if (!local()) {
    // Q: plane-side exchange buffer, one block [maxPlanes][pitch(r)] per source rank
    const size_t sticksAllRanks = static_cast<size_t>(maxDimX_) * static_cast<size_t>(maxDimY_) +
                                  static_cast<size_t>(kMaxTileLanes) * comm_->size();
    q_.allocate(c * static_cast<size_t>(maxPlanes_) * sticksAllRanks);
  }
}

template <typename T>
void GridResources<T>::map_peers() {
  // Peer-memory exchange: map arrays A and Q and the barrier flags of every rank (CUDA IPC).
  // SPFFT_B200_P2P=0 keeps the NCCL send/recv exchange. Collective; every rank takes the same
  // decision (PeerWindow votes).
  peerOk_ = false;
  if (local()) return;
  const char* env = std::getenv("SPFFT_B200_P2P");
  int want = !(env && std::atoi(env) == 0) && comm_->size() <= sb::kMaxPeers ? 1 : 0;
  if (want) {
    // second copy of the buffers the peers store into (array_a(1) / array_q(1), see transform_engine.hpp); a rank
    // that cannot allocate it votes against the peer exchange (everybody then uses NCCL)
    try {
      a2_.allocate(a_.bytes());
      q2_.allocate(q_.bytes());
    } catch (const GPUAllocationError&) {
      want = 0;
    }
  }
  {
    const std::vector<int> votes = comm_->all_gather_ints(&want, 1);
    for (int v : votes) want = want && v;
  }
  if (!want) {
    a2_.allocate(0);
    q2_.allocate(0);
    return;
  }
  flags_.allocate(sizeof(int) * sb::kMaxPeers);
  check_gpu(cudaMemset(flags_.get(), 0, flags_.bytes()));
  check_gpu(cudaDeviceSynchronize());
  peerA_.open(*comm_, a_.get());
  peerQ_.open(*comm_, q_.get());
  peerA2_.open(*comm_, a2_.get());
  peerQ2_.open(*comm_, q2_.get());
  peerFlags_.open(*comm_, flags_.get());
  peerOk_ = peerA_.mapped() && peerQ_.mapped() && peerA2_.mapped() && peerQ2_.mapped() && peerFlags_.mapped();
  if (!peerOk_) {
    a2_.allocate(0);
    q2_.allocate(0);
  }
  barrierEpoch_ = 0;
  exchangeCalls_[0] = exchangeCalls_[1] = 0;
}

template <typename T>
void GridResources<T>::enqueue_peer_barrier(cudaStream_t stream) {
  if (!peerOk_) throw InternalError();
  int* flags[sb::kMaxPeers] = {};
  for (int r = 0; r < comm_->size(); ++r) flags[r] = static_cast<int*>(peerFlags_.ptr(r));
  ++barrierEpoch_;
  check_launch(sb_launch_peer_barrier(flags, comm_->size(), comm_->rank(), barrierEpoch_, stream));
}

template <typename T>
void* GridResources<T>::host_space(size_t bytes) {
  std::lock_guard<std::mutex> lock(hostMutex_);
  if (host_.bytes() < bytes) {
    DeviceGuard guard(deviceId_);
    const size_t volume = static_cast<size_t>(maxDimX_) * static_cast<size_t>(maxDimY_) *
                          static_cast<size_t>(maxPlanes_);
    host_.allocate(std::max(bytes, 2 * sizeof(T) * volume));
  }
  return host_.get();
}

template <typename T>
void* GridResources<T>::scratch(size_t bytes) {
  std::lock_guard<std::mutex> lock(hostMutex_);
  if (scratch_.bytes() < bytes) {
    DeviceGuard guard(deviceId_);
    check_gpu(cudaDeviceSynchronize());  // nothing may still be using the old allocation
    scratch_.allocate(bytes);
  }
  return scratch_.get();
}

template <typename T>
int* GridResources<T>::counters(size_t count) {
  std::lock_guard<std::mutex> lock(hostMutex_);
  if (counters_.bytes() < count * sizeof(int)) {
    DeviceGuard guard(deviceId_);
    check_gpu(cudaDeviceSynchronize());
    counters_.allocate(count * sizeof(int));
  }
  return counters_.template as<int>();
}

// ---------------------------------------------------------------------------------------------
// DevicePlan
// ---------------------------------------------------------------------------------------------
void choose_tile_lanes(const IndexMaps& m, int cb, long long smemLimit, AxisPlans& ax, bool& fastX,
                       bool& fastY, bool& fastZ) {
  // per axis: register-FFT kernels for power-of-two lengths, generic tile kernels otherwise
  fastX = fast_path_length(m.dimX, cb);  // complex rows and real rows (x_r2c_tile)
  fastY = fast_path_length(m.dimY, cb);
  fastZ = fast_path_length(m.dimZ, cb);
  const int fastLanes = fast_path_log2_lanes(cb);
  // (the stand-alone x kernels pick their own row count; build_device_plan sets it once it knows
  // that the opt-in fused xy kernels, which share one tile shape between y and x, are not used)
  ax.log2Vx = fastX ? fastLanes : choose_log2_lanes(m.dimX, cb, smemLimit);
  ax.log2Vy = fastY ? fast_path_log2_lanes(cb, m.dimY) : choose_log2_lanes(m.dimY, cb, smemLimit);
  ax.log2Vz = fastZ ? fast_path_log2_lanes(cb, m.dimZ) : choose_log2_lanes(m.dimZ, cb, smemLimit);
  // a transform length whose two tile buffers exceed shared memory is not supported
  if (ax.log2Vx < 0 || ax.log2Vy < 0 || ax.log2Vz < 0) throw InvalidParameterError();
}

void throw_error_code(int code) {
  switch (code) {
    case SPFFT_SUCCESS: return;
    case SPFFT_OVERFLOW_ERROR: throw OverflowError();
    case SPFFT_ALLOCATION_ERROR: throw HostAllocationError();
    case SPFFT_INVALID_PARAMETER_ERROR: throw InvalidParameterError();
    case SPFFT_DUPLICATE_INDICES_ERROR: throw DuplicateIndicesError();
    case SPFFT_INVALID_INDICES_ERROR: throw InvalidIndicesError();
    case SPFFT_MPI_ERROR: throw MPIError();
    case SPFFT_MPI_PARAMETER_MISMATCH_ERROR: throw MPIParameterMismatchError();
    case SPFFT_FFTW_ERROR: throw InternalError();
    case SPFFT_GPU_ALLOCATION_ERROR: throw GPUAllocationError();
    case SPFFT_GPU_LAUNCH_ERROR: throw GPULaunchError();
    case SPFFT_GPU_INVALID_VALUE_ERROR: throw GPUInvalidValueError();
    case SPFFT_GPU_ERROR: throw GPUError();
    default: throw GenericError();
  }
}

void agree_on_error(Communicator& comm, int localCode) {
  // rank-local failures of a collective construction become the same error on EVERY rank (the
  // lowest failing rank's code) instead of leaving the other ranks waiting in the next collective
  const std::vector<int> all = comm.all_gather_ints(&localCode, 1);
  for (int c : all)
    if (c != SPFFT_SUCCESS) throw_error_code(c);
}

namespace {
template <typename F>
int error_code_of(F&& f) {
  try {
    f();
  } catch (const GenericError& e) {
    return static_cast<int>(e.error_code());
  } catch (const std::bad_alloc&) {
    return static_cast<int>(SPFFT_ALLOCATION_ERROR);
  } catch (...) {
    return static_cast<int>(SPFFT_UNKNOWN_ERROR);
  }
  return static_cast<int>(SPFFT_SUCCESS);
}
}  // namespace

std::shared_ptr<IndexMaps> make_distributed_index_maps(Communicator& comm, SpfftTransformType type,
                                                       int dimX, int dimY, int dimZ, int localZLength,
                                                       int numLocalElements,
                                                       SpfftIndexFormatType indexFormat,
                                                       const int* indices) {
  // local part exactly as for a single rank (index conversion, zeroZeroStickIndex); a rank whose
  // indices are invalid still takes part in the collectives below and reports its error code there
  std::shared_ptr<IndexMaps> m;
  const int localError = error_code_of(
      [&] { m = make_local_index_maps(type, dimX, dimY, dimZ, numLocalElements, indexFormat, indices); });
  const int P = comm.size();
  const int mine[7] = {dimX, dimY, dimZ, localZLength, m ? m->num_sticks() : 0, numLocalElements, localError};
  const std::vector<int> all7 = comm.all_gather_ints(mine, 7);  // parameters.cpp:89
  for (int r = 0; r < P; ++r) throw_error_code(all7[7 * r + 6]);
  std::vector<int> all(static_cast<size_t>(6) * P);
  for (int r = 0; r < P; ++r)
    for (int k = 0; k < 6; ++k) all[6 * r + k] = all7[7 * r + k];
  std::vector<std::vector<long long>> counts(P, std::vector<long long>(6));
  int maxSticks = 0;
  for (int r = 0; r < P; ++r) {
    for (int k = 0; k < 6; ++k) counts[r][k] = all[6 * r + k];
    maxSticks = std::max(maxSticks, all[6 * r + 4]);
  }
  // stick lists of all ranks (the reference sends them point to point, indices.hpp:58-102)
  std::vector<std::vector<int>> sticks(P);
  if (maxSticks > 0) {
    std::vector<int> padded(static_cast<size_t>(maxSticks), -1);
    std::copy(m->stickIndices.begin(), m->stickIndices.end(), padded.begin());
    const std::vector<int> gathered = comm.all_gather_ints(padded.data(), maxSticks);
    for (int r = 0; r < P; ++r)
      sticks[r].assign(gathered.begin() + static_cast<size_t>(r) * maxSticks,
                       gathered.begin() + static_cast<size_t>(r) * maxSticks + all[6 * r + 4]);
  }
  agree_on_error(comm, error_code_of([&] { finish_distributed_index_maps(*m, comm.rank(), counts, std::move(sticks)); }));
  return m;
}

template <typename T>
std::shared_ptr<DevicePlan<T>> build_device_plan(const IndexMaps& m, long long smemLimit) {
  auto plan = std::make_shared<DevicePlan<T>>();
  if (m.dimX == 0 || m.dimY == 0 || m.dimZ == 0) return plan;
  AxisPlans& ax = plan->axes;
  const int cb = static_cast<int>(sizeof(sb::cx<T>));
  bool fastX = false, fastY = false, fastZ = false;
  choose_tile_lanes(m, cb, smemLimit, ax, fastX, fastY, fastZ);
  ax.rpX = make_radix_plan(m.dimX);
  ax.rpY = make_radix_plan(m.dimY);
  ax.rpZ = make_radix_plan(m.dimZ);
  // Warp-FFT kernels (default where they exist; SPFFT_B200_WFFT=0 selects the round-1 kernels, bit 0 =
  // fused xy stage, bit 1 = z stage): one warp per transform, tiles staged by TMA, y <-> x hand-off in L2.
  const char* wEnv = std::getenv("SPFFT_B200_WFFT");
  const int wMask = wEnv ? std::atoi(wEnv) : 15;
  // (bit 2: the fused xy stage also for distributed transforms, whose y tiles read / write the exchange buffers;
  // bit 3: single precision, two transforms per warp on the packed fp32 pipe -- local C2C transforms)
  const bool single = sizeof(T) == 4;
  const bool wSingleOk = !single || ((wMask & 8) && m.commSize == 1 && m.type == SPFFT_TRANS_C2C);
  if (!plan->fusedXY && (wMask & 1) && wSingleOk && fastX && fastY && m.dimX == m.dimY &&
      m.type == SPFFT_TRANS_C2C && ax.log2Vy == 3 &&
      (m.commSize == 1 ? m.num_sticks() > 0 : ((wMask & 4) && m.local_planes() > 0))) {
    const int err = sb_wxy_config(single ? 1 : 0, m.dimX, m.local_planes(), &plan->xyRing, &plan->xyLag, &plan->xyCounters);
    plan->fusedXY = plan->wfftXY = err == 0;
  }
  if (!plan->fusedXY && fastX) ax.log2Vx = fast_path_log2_lanes_x(m.dimX);
  TileMaps t = build_tile_maps(m, ax.log2Vz, ax.log2Vy, fastZ, fastY);
  plan->wfftZ = (wMask & 2) && wSingleOk && fastZ && ax.log2Vz == 3 && sb_wz_available(single ? 1 : 0, m.dimZ) && !t.zInv.empty();
  plan->numStickTiles = t.numStickTiles;
  plan->pitch = t.pitch;
  plan->numXTiles = t.numXTiles;
  plan->symTile = t.symTile;
  plan->symLane = t.symLane;

  PlanPointers<T>& p = plan->ptrs;
  auto& st = plan->storage;
  size_t& total = plan->deviceBytes;
  p.twX = upload(st, total, make_roots<T>(m.dimX));
  p.twY = upload(st, total, make_roots<T>(m.dimY));
  p.twZ = upload(st, total, make_roots<T>(m.dimZ));
  if (fastX) p.ftwX = upload(st, total, make_fast_twiddles<T>(m.dimX));
  if (fastY) p.ftwY = upload(st, total, make_fast_twiddles<T>(m.dimY));
  if (fastZ) p.ftwZ = upload(st, total, make_fast_twiddles<T>(m.dimZ));
  p.tileStart = upload(st, total, t.tileStart);
  p.entrySrc = t.identityOrder ? nullptr : upload(st, total, t.entrySrc);
  p.zInv = upload(st, total, t.zInv);
  if (m.commSize == 1) p.yInv = upload(st, total, t.yInv);
  // the gather-form z kernels do not read entrySlot
  if (!p.zInv) p.entrySlot = upload(st, total, t.entrySlot);
  if (t.hasDuplicates) {
    p.bwdTileStart = upload(st, total, t.bwdTileStart);
    p.bwdEntrySrc = upload(st, total, t.bwdEntrySrc);
    p.bwdEntrySlot = upload(st, total, t.bwdEntrySlot);
  } else {
    p.bwdTileStart = p.tileStart;
    p.bwdEntrySrc = p.entrySrc;
    p.bwdEntrySlot = p.entrySlot;
  }
  if (m.commSize > 1) {
    // y stage over ALL ranks' sticks, read from / written to the plane-side exchange buffer
    plan->distributed = true;
    plan->exchange = build_exchange_plan(m, ax.log2Vz, ax.log2Vy, fastY);
    plan->numXTiles = plan->exchange.numXTiles;
    p.xtStart = upload(st, total, plan->exchange.xtStart);
    p.stickSlot = upload(st, total, plan->exchange.stickSlot);
    plan->srcBase = upload(st, total, plan->exchange.srcBase);
    plan->srcPitch = upload(st, total, plan->exchange.srcPitch);
    plan->tileBase = upload(st, total, plan->exchange.tileBase);
    plan->tilePitch = upload(st, total, plan->exchange.tilePitch);
    // inverse-map form for the tiles whose sticks are contiguous in one source block
    p.yInv = upload(st, total, plan->exchange.yInv);
    plan->rowRank = upload(st, total, plan->exchange.rowRank);
    plan->rowOff = upload(st, total, plan->exchange.rowOff);
    plan->stickRank = upload(st, total, plan->exchange.stickRank);
    plan->fwdBase = upload(st, total, plan->exchange.fwdBase);
    plan->tileFwdBase = upload(st, total, plan->exchange.tileFwdBase);
    // SPFFT_B200_FWD_ORDER=1: destination ranks interleaved tile by tile instead of the rotated order
    // (experiment: no difference on 2 GPUs, 0.689 ms for the forward exchange kernel either way;
    // the rotated order is the one measured on 4 and 8 GPUs, so it stays the default)
    const char* ord = std::getenv("SPFFT_B200_FWD_ORDER");
    if (ord && std::atoi(ord) == 1) plan->fwdTileOrder = upload(st, total, plan->exchange.fwdTileOrder);
    // the fused xy stage needs the inverse map over all ranks' sticks
    if (plan->exchange.yInv.empty() || plan->exchange.numXTiles * 8 != m.dimX) plan->fusedXY = plan->wfftXY = false;
  } else {
    p.xtStart = upload(st, total, t.xtStart);
    p.stickSlot = upload(st, total, t.stickSlot);
  }
  return plan;
}

template std::shared_ptr<DevicePlan<double>> build_device_plan<double>(const IndexMaps&, long long);
template std::shared_ptr<DevicePlan<float>> build_device_plan<float>(const IndexMaps&, long long);

// ---------------------------------------------------------------------------------------------
// TransformEngine
// ---------------------------------------------------------------------------------------------
template <typename T>
TransformEngine<T>::TransformEngine(SpfftProcessingUnitType executionUnit,
                                    std::shared_ptr<GridResources<T>> grid,
                                    std::shared_ptr<IndexMaps> maps,
                                    std::shared_ptr<DevicePlan<T>> plan)
    : executionUnit_(executionUnit),
      grid_(std::move(grid)),
      maps_(std::move(maps)),
      plan_(std::move(plan)) {
  if (!grid_) throw InvalidParameterError();
  auto construct = [&] {
  // transform_internal.cpp:52-80
  if (maps_->local_planes() > grid_->max_num_local_xy_planes()) throw InvalidParameterError();
  if (grid_->local() && maps_->dimZ != maps_->local_planes()) throw InvalidParameterError();
  if (!grid_->local() && (grid_->communicator()->size() != maps_->commSize ||
                          grid_->communicator()->rank() != maps_->commRank))
    throw InternalError();  // transform_internal.cpp:81-84
  if (grid_->local() && maps_->commSize != 1) throw InternalError();
  if (maps_->num_sticks() > grid_->max_num_local_z_columns()) throw InvalidParameterError();
  if (maps_->dimX > grid_->max_dim_x() || maps_->dimY > grid_->max_dim_y() ||
      maps_->dimZ > grid_->max_dim_z())
    throw InvalidParameterError();
  if (!(executionUnit & grid_->processing_unit())) throw InvalidParameterError();
  if (executionUnit != SPFFT_PU_HOST && executionUnit != SPFFT_PU_GPU)
    throw InvalidParameterError();
  // (SPFFT_PU_HOST: executed by the device kernels, data staged through host memory -- there is no CPU path;
  //  such a transform only accepts / exposes host locations, like the reference's host transform)

  DeviceGuard guard(grid_->device_id());
  if (!plan_) {
    long long smem = 0;
    check_gpu(static_cast<cudaError_t>(sb_max_dynamic_smem(&smem)));
    plan_ = build_device_plan<T>(*maps_, smem);
  }
  // the grid limits are expressed in sticks; the padded pitch must fit as well
  if (sizeof(sb::cx<T>) * static_cast<size_t>(maps_->dimZ) * static_cast<size_t>(plan_->pitch) >
      grid_->bytes_a())
    throw InvalidParameterError();
  if (plan_->distributed &&
      sizeof(sb::cx<T>) * static_cast<size_t>(plan_->exchange.planeSideElements) > grid_->bytes_q())
    throw InvalidParameterError();
  // array B also stages the frequency values of host-pointer calls; duplicate triplets can make
  // numLocalElements larger than every grid-derived size (indices.hpp:124 only bounds it by the volume)
  if (sizeof(sb::cx<T>) * static_cast<size_t>(maps_->num_values()) > grid_->bytes_b())
    throw InvalidParameterError();
  };
  // distributed: every rank reports, all ranks throw the same error (plan construction can fail on
  // one rank only: table overflow, shared-memory limit, buffer sizes)
  if (grid_->local())
    construct();
  else
    agree_on_error(*grid_->communicator(), error_code_of(construct));
  stream_.reset(new Stream());
  startEvent_.reset(new Event());
  endEvent_.reset(new Event());
}

template <typename T>
TransformEngine<T>::~TransformEngine() {
  // outstanding asynchronous work must not outlive the buffers
  if (stream_) cudaStreamSynchronize(stream_->get());
}

template <typename T>
std::shared_ptr<TransformEngine<T>> TransformEngine<T>::clone() const {
  // transform_internal.cpp:176-179: same parameters, new grid
  auto newGrid = std::make_shared<GridResources<T>>(*grid_);
  auto e = std::make_shared<TransformEngine<T>>(executionUnit_, std::move(newGrid), maps_, plan_);
  e->execMode_ = SPFFT_EXEC_SYNCHRONOUS;
  return e;
}

template <typename T>
size_t TransformEngine<T>::space_bytes() const {
  const size_t n = static_cast<size_t>(maps_->dimX) * static_cast<size_t>(maps_->dimY) *
                   static_cast<size_t>(maps_->local_planes());
  return n * (maps_->type == SPFFT_TRANS_R2C ? sizeof(T) : 2 * sizeof(T));
}

template <typename T>
T* TransformEngine<T>::device_space() const {
  // C2C: the x stage works in place on the plane buffer (like the reference,
  // execution_gpu.cpp:106-112); R2C: real rows live in array A (execution_gpu.cpp:97-103)
  return static_cast<T*>(maps_->type == SPFFT_TRANS_R2C ? grid_->array_a() : grid_->array_b());
}

template <typename T>
T* TransformEngine<T>::space_domain_data(SpfftProcessingUnitType location) {
  // transform_internal.cpp:212-232: a host transform has no device-side space domain
  if (location == SPFFT_PU_GPU && executionUnit_ != SPFFT_PU_GPU) throw InvalidParameterError();
  if (location == SPFFT_PU_GPU) return device_space();
  if (location == SPFFT_PU_HOST) return static_cast<T*>(grid_->host_space(space_bytes()));
  throw InvalidParameterError();
}

template <typename T>
void TransformEngine<T>::begin_call() {
  // sticky-error probe, execution_gpu.cpp:256-258,329-331
  if (cudaGetLastError() != cudaSuccess) throw GPUPrecedingError();
  // order after the default stream, execution_gpu.cpp:260-261
  check_gpu(cudaEventRecord(startEvent_->get(), nullptr));
  check_gpu(cudaStreamWaitEvent(stream_->get(), startEvent_->get(), 0));
  if (profiling_) record_stage(nullptr);  // start of a call
}

template <typename T>
void TransformEngine<T>::record_stage(const char* name) {
  if (!profiling_) return;
  if (profUsed_ == profEvents_.size()) profEvents_.emplace_back(new Event(true));
  check_gpu(cudaEventRecord(profEvents_[profUsed_]->get(), stream_->get()));
  profNames_.push_back(name);
  ++profUsed_;
}

template <typename T>
std::vector<StageTime> TransformEngine<T>::stage_times() {
  // Accumulates over every call since the previous query: average milliseconds per stage name
  // (in order of first appearance), then forgets the recorded events.
  std::vector<StageTime> out;
  std::vector<int> counts;
  if (profUsed_ < 2) {
    profUsed_ = 0;
    profNames_.clear();
    return out;
  }
  DeviceGuard guard(grid_->device_id());
  check_gpu(cudaEventSynchronize(profEvents_[profUsed_ - 1]->get()));
  for (size_t i = 1; i < profUsed_; ++i) {
    if (profNames_[i] == nullptr) continue;  // boundary between two calls
    float ms = 0.f;
    check_gpu(cudaEventElapsedTime(&ms, profEvents_[i - 1]->get(), profEvents_[i]->get()));
    size_t k = 0;
    while (k < out.size() && out[k].name != profNames_[i]) ++k;
    if (k == out.size()) {
      out.push_back(StageTime{profNames_[i], 0.f});
      counts.push_back(0);
    }
    out[k].ms += ms;
    ++counts[k];
  }
  for (size_t k = 0; k < out.size(); ++k) out[k].ms /= static_cast<float>(counts[k]);
  profUsed_ = 0;
  profNames_.clear();
  return out;
}

template <typename T>
void TransformEngine<T>::synchronize() {
  DeviceGuard guard(grid_->device_id());
  if (execMode_ == SPFFT_EXEC_SYNCHRONOUS) {
    check_gpu(cudaStreamSynchronize(stream_->get()));
  } else {
    // execution_gpu.cpp:403-410: the default stream waits for the transform
    check_gpu(cudaEventRecord(endEvent_->get(), stream_->get()));
    check_gpu(cudaStreamWaitEvent(nullptr, endEvent_->get(), 0));
  }
}

template <typename T>
int TransformEngine<T>::host_slabs() const {
  // SPFFT_B200_HOST_SLABS (default 4, 1 = one copy after / before the whole xy stage)
  static const int n = [] {
    const char* e = std::getenv("SPFFT_B200_HOST_SLABS");
    const int v = e ? std::atoi(e) : 4;
    return v < 1 ? 1 : (v > 16 ? 16 : v);
  }();
  const bool ok = !plan_->distributed && maps_->type == SPFFT_TRANS_C2C && maps_->local_planes() >= 8 * n;
  return ok ? n : 1;
}

template <typename T>
cudaStream_t TransformEngine<T>::copy_stream() {
  if (!copyStream_) copyStream_.reset(new Stream());
  return copyStream_->get();
}

template <typename T>
cudaEvent_t TransformEngine<T>::slab_event(int i) {
  while (static_cast<int>(slabEvents_.size()) <= i) slabEvents_.emplace_back(new Event());
  return slabEvents_[static_cast<size_t>(i)]->get();
}

template <typename T>
sb::XYArgs<T> TransformEngine<T>::make_xy_args(const TileMaps& geo, const T* spaceIn, T* spaceOut, bool forward,
                                               int parity, int planeBegin, int planeCount) {
  const IndexMaps& m = *maps_;
  sb::XYArgs<T> a{};
  a.y = make_y_stage_args(geo, forward, parity);
  a.y.planes = nullptr;
  a.x = make_x_args<T>(m, plan_->axes, plan_->ptrs, nullptr, spaceIn, spaceOut);
  if (planeCount >= 0) {
    // a slab of the local planes (host-pointer calls): stick rows / space planes from planeBegin on
    const size_t off = 2 * static_cast<size_t>(planeBegin) * static_cast<size_t>(m.dimX) * static_cast<size_t>(m.dimY);
    a.y.zRowOffset += planeBegin;
    a.y.numPlanes = planeCount;
    a.x.numPlanes = planeCount;
    if (spaceIn) a.x.spaceIn = spaceIn + off;
    if (spaceOut) a.x.spaceOut = spaceOut + off;
  }
  a.ring = plan_->xyRing;
  a.lag = plan_->xyLag;
  const size_t planeBytes = sizeof(sb::cx<T>) * static_cast<size_t>(m.dimX) * static_cast<size_t>(m.dimY);
  a.scratch = static_cast<sb::cx<T>*>(grid_->scratch(planeBytes * static_cast<size_t>(a.ring)));
  a.counters = grid_->counters(static_cast<size_t>(plan_->xyCounters));
  return a;
}

template <typename T>
sb::YArgs<T> TransformEngine<T>::make_y_stage_args(const TileMaps& geo, bool forward, int parity) {
  auto ya = make_y_args<T>(*maps_, geo, plan_->axes, plan_->ptrs, sticks(), planes());
  if (plan_->distributed && forward && peer_exchange()) {
    // forward over peer memory: every stick goes straight into its owner's stick buffer (copy `parity`)
    for (int r = 0; r < maps_->commSize; ++r) ya.peer[r] = static_cast<sb::cx<T>*>(grid_->peer_a(parity, r));
    ya.stickRank = plan_->stickRank;
    ya.fwdBase = plan_->fwdBase;
    ya.tileFwdBase = plan_->tileFwdBase;
    ya.xtOrder = plan_->fwdTileOrder;
    ya.xtRotate = plan_->exchange.fwdTileRotate;
  }
  if (plan_->distributed) {
    // sticks of all ranks, read from / written to the plane-side exchange buffer
    ya.sticks = static_cast<sb::cx<T>*>(grid_->array_q(forward ? 0 : parity));
    ya.srcBase = plan_->srcBase;
    ya.srcPitch = plan_->srcPitch;
    ya.tileBase = plan_->tileBase;
    ya.tilePitch = plan_->tilePitch;
    ya.zRowOffset = 0;
    ya.wireF32 = wire_f32() ? 1 : 0;
  }
  return ya;
}

template <typename T>
void TransformEngine<T>::enqueue_backward(const T* input, T* output) {
  DeviceGuard guard(grid_->device_id());
  begin_call();
  const IndexMaps& m = *maps_;
  const bool dist = plan_->distributed;
  if (static_cast<size_t>(m.dimX) * m.dimY * m.dimZ == 0) return;
  const bool haveSpace = space_bytes() > 0;
  if (!haveSpace && !dist) return;
  if (haveSpace && !output) throw InvalidParameterError();
  cudaStream_t s = stream_->get();
  const size_t ne = static_cast<size_t>(m.num_values());
  if (ne > 0 && !input) throw InvalidParameterError();

  // ---- z stage: decompress + stick symmetry + z-FFT (execution_gpu.cpp:327-369)
  const T* values = input;
  if (ne > 0 && !is_device_pointer(input)) {
    check_gpu(cudaMemcpyAsync(grid_->array_b(), input, ne * 2 * sizeof(T), cudaMemcpyHostToDevice, s));
    values = static_cast<const T*>(grid_->array_b());
    record_stage("h2d values");
  }
  // the stage wiring only reads the scalars of the tile maps
  TileMaps geo;
  geo.numStickTiles = plan_->numStickTiles;
  geo.pitch = plan_->pitch;
  geo.numXTiles = plan_->numXTiles;
  geo.symTile = plan_->symTile;
  geo.symLane = plan_->symLane;
  const bool peer = dist && peer_exchange();
  // Peers may still read the plane-side buffer the previous backward call filled: this call's z stages store into
  // the other copy (no barrier at the start of the call; the copy used two calls ago is free because every rank
  // passed the barrier of the previous call after its last read of it).
  const int parity = peer ? grid_->next_exchange_parity(0) : 0;
  if (plan_->numStickTiles > 0) {
    auto za = make_z_args<T>(m, geo, plan_->axes, plan_->ptrs, false, sticks(), values, nullptr,
                             false);
    if (peer) {
      // the z stage stores every row straight into the plane-side buffer of the rank that owns
      // the plane: compute and exchange are one kernel
      for (int r = 0; r < m.commSize; ++r) za.peer[r] = static_cast<sb::cx<T>*>(grid_->peer_q(parity, r));
      za.rowRank = plan_->rowRank;
      za.rowOff = plan_->rowOff;
    }
    za.wireF32 = wire_f32() ? 1 : 0;
    // (single precision: the backward warp kernel gathers two sticks per warp with 8-byte loads and is slower than
    // the round-1 kernel, 0.390 against 0.357 ms at 512^3 -- profiles/r02_summary.md; the forward one is faster)
    const bool warpZ = plan_->wfftZ && !peer && !za.wireF32 && sizeof(T) == 8;
    check_launch(warpZ ? Launch<T>::wz(0, za, s) : Launch<T>::z(0, za, s));
    record_stage(peer ? "z backward + exchange" : "z backward");
  }
  // ---- exchange: every rank sends, for every peer, the rows of that peer's slab (one contiguous
  // block of the plane-major stick buffer) -- replaces pack + MPI_Alltoallv + unpack
  // (transpose_mpi_compact_buffered_gpu.cpp:163-224)
  if (peer) {
    grid_->enqueue_peer_barrier(s);
    record_stage("barrier backward");
  } else if (dist) {
    const ExchangePlan& x = plan_->exchange;
    grid_->communicator()->all_to_all_v(sticks(), x.stickOffset.data(), x.stickCount.data(),
                                        grid_->array_q(), x.planeOffset.data(), x.planeCount.data(),
                                        static_cast<int>(wire_f32() ? sizeof(sb::cx<float>) : sizeof(sb::cx<T>)), s);
    record_stage("exchange backward");
  }
  if (!haveSpace) return;
  const bool outOnDevice = is_device_pointer(output);
  T* outDev = outOnDevice ? output : device_space();
  // (the warp-FFT kernels move rows with 16-byte bulk copies / vector accesses: a space pointer that is only
  // aligned to its scalar type takes the separate y and x kernels)
  bool fusedHere = plan_->fusedXY && !(plan_->wfftXY && (reinterpret_cast<size_t>(outDev) & 15) != 0) &&
                   !(dist && wire_f32());
  // Host output (reference: execution_gpu.cpp:391-397 copies after the whole stage): the xy stage runs slab by slab
  // and every finished slab leaves on a second stream while the next one is computed.
  const int slabs = outOnDevice ? 1 : host_slabs();
  const int P = m.local_planes();
  const size_t planeReals = space_bytes() / sizeof(T) / static_cast<size_t>(P > 0 ? P : 1);
  for (int c = 0; c < slabs; ++c) {
    const int p0 = static_cast<int>(static_cast<long long>(P) * c / slabs);
    const int p1 = static_cast<int>(static_cast<long long>(P) * (c + 1) / slabs);
    const bool whole = slabs == 1;
    if (fusedHere) {
      // ---- fused xy stage: y tiles and x tiles in one persistent kernel, hand-off through L2
      const int err = Launch<T>::wxy(0, whole ? make_xy_args(geo, nullptr, outDev, false, parity)
                                               : make_xy_args(geo, nullptr, outDev, false, parity, p0, p1 - p0), s);
      if (err == static_cast<int>(cudaErrorCooperativeLaunchTooLarge)) {
        // the device cannot hold the whole persistent grid right now (e.g. shared with another process):
        // the separate y and x kernels do the same work through the plane buffer
        cudaGetLastError();
        fusedHere = false;
      } else {
        check_launch(err);
        record_stage("xy backward");
      }
    }
    if (!fusedHere) {
      // ---- y stage: stick gather + plane symmetry + y-FFT (execution_gpu.cpp:371-390)
      auto ya = make_y_stage_args(geo, false, parity);
      auto xa = make_x_args<T>(m, plan_->axes, plan_->ptrs, planes(), nullptr, outDev);
      if (!whole) {
        const size_t planeElems = static_cast<size_t>(m.dimY) * static_cast<size_t>(m.dimXFreq);
        ya.zRowOffset += p0;
        ya.numPlanes = p1 - p0;
        ya.planes += planeElems * static_cast<size_t>(p0);
        xa.numPlanes = p1 - p0;
        xa.planes += planeElems * static_cast<size_t>(p0);
        xa.spaceOut = outDev + planeReals * static_cast<size_t>(p0);
      }
      check_launch(Launch<T>::y(0, ya, s));
      record_stage("y backward");
      // ---- x stage: x-FFT (C2C / C2R) into the space domain
      check_launch(Launch<T>::x(0, xa, s));
      record_stage("x backward");
    }
    if (!outOnDevice) {
      if (whole) {
        check_gpu(cudaMemcpyAsync(output, outDev, space_bytes(), cudaMemcpyDeviceToHost, s));
      } else {
        cudaStream_t cs = copy_stream();
        check_gpu(cudaEventRecord(slab_event(c), s));
        check_gpu(cudaStreamWaitEvent(cs, slab_event(c), 0));
        check_gpu(cudaMemcpyAsync(output + planeReals * static_cast<size_t>(p0), outDev + planeReals * static_cast<size_t>(p0),
                                  planeReals * static_cast<size_t>(p1 - p0) * sizeof(T), cudaMemcpyDeviceToHost, cs));
      }
    }
  }
  if (!outOnDevice) {
    if (slabs > 1) {
      // the transform's stream is the one callers synchronise with: it waits for the last slab's copy
      check_gpu(cudaEventRecord(slab_event(slabs), copy_stream()));
      check_gpu(cudaStreamWaitEvent(s, slab_event(slabs), 0));
    }
    record_stage("d2h space");
  }
}

template <typename T>
void TransformEngine<T>::enqueue_forward(const T* input, T* output, SpfftScalingType scaling) {
  DeviceGuard guard(grid_->device_id());
  begin_call();
  const IndexMaps& m = *maps_;
  const bool dist = plan_->distributed;
  if (static_cast<size_t>(m.dimX) * m.dimY * m.dimZ == 0) return;
  const bool haveSpace = space_bytes() > 0;
  if (!haveSpace && !dist) return;
  if (haveSpace && !input) throw InvalidParameterError();
  cudaStream_t s = stream_->get();
  const size_t ne = static_cast<size_t>(m.num_values());
  if (ne > 0 && !output) throw InvalidParameterError();

  TileMaps geo;
  geo.numStickTiles = plan_->numStickTiles;
  geo.pitch = plan_->pitch;
  geo.numXTiles = plan_->numXTiles;
  geo.symTile = plan_->symTile;
  geo.symLane = plan_->symLane;
  // sticks anywhere (all ranks)? otherwise nothing consumes the planes
  const bool anySticks = dist ? !plan_->exchange.stickSlot.empty() : (plan_->numStickTiles > 0 && ne > 0);

  const bool peer = dist && peer_exchange();
  // peers may still read the stick buffer the previous forward call filled: this call's y stages store into the
  // other copy (see enqueue_backward)
  const int parity = peer ? grid_->next_exchange_parity(1) : 0;
  // ---- x stage (execution_gpu.cpp:254-282)
  if (haveSpace) {
    const T* src = input;
    const bool inOnDevice = is_device_pointer(input);
    // Host input (reference: execution_gpu.cpp:263-276 copies before the whole stage): the space domain arrives slab
    // by slab on a second stream, the xy stage of a slab starts as soon as it is there.
    const int slabs = inOnDevice ? 1 : host_slabs();
    const int P = m.local_planes();
    const size_t planeReals = space_bytes() / sizeof(T) / static_cast<size_t>(P > 0 ? P : 1);
    if (!inOnDevice) {
      src = device_space();
      if (slabs == 1) {
        check_gpu(cudaMemcpyAsync(device_space(), input, space_bytes(), cudaMemcpyHostToDevice, s));
        record_stage("h2d space");
      } else {
        // the copies may overwrite the internal space buffer only after everything enqueued so far is done with it
        cudaStream_t cs = copy_stream();
        check_gpu(cudaEventRecord(slab_event(slabs), s));
        check_gpu(cudaStreamWaitEvent(cs, slab_event(slabs), 0));
        for (int c = 0; c < slabs; ++c) {
          const size_t r0 = planeReals * static_cast<size_t>(static_cast<long long>(P) * c / slabs);
          const size_t r1 = planeReals * static_cast<size_t>(static_cast<long long>(P) * (c + 1) / slabs);
          check_gpu(cudaMemcpyAsync(device_space() + r0, input + r0, (r1 - r0) * sizeof(T), cudaMemcpyHostToDevice, cs));
          check_gpu(cudaEventRecord(slab_event(c), cs));
        }
      }
    }
    bool fusedHere = plan_->fusedXY && !(plan_->wfftXY && (reinterpret_cast<size_t>(src) & 15) != 0) &&
                     !(dist && wire_f32());
    for (int c = 0; c < slabs; ++c) {
      const int p0 = static_cast<int>(static_cast<long long>(P) * c / slabs);
      const int p1 = static_cast<int>(static_cast<long long>(P) * (c + 1) / slabs);
      const bool whole = slabs == 1;
      if (!whole) check_gpu(cudaStreamWaitEvent(s, slab_event(c), 0));
      if (fusedHere && anySticks) {
        const int err = Launch<T>::wxy(1, whole ? make_xy_args(geo, src, nullptr, true, parity)
                                                 : make_xy_args(geo, src, nullptr, true, parity, p0, p1 - p0), s);
        if (err == static_cast<int>(cudaErrorCooperativeLaunchTooLarge)) {
          cudaGetLastError();  // see enqueue_backward
          fusedHere = false;
        } else {
          check_launch(err);
          record_stage(peer ? "xy forward + exchange" : "xy forward");
        }
      }
      if (!fusedHere) {
        auto xa = make_x_args<T>(m, plan_->axes, plan_->ptrs, planes(), src, nullptr);
        auto ya = make_y_stage_args(geo, true, parity);
        if (!whole) {
          const size_t planeElems = static_cast<size_t>(m.dimY) * static_cast<size_t>(m.dimXFreq);
          xa.numPlanes = p1 - p0;
          xa.planes += planeElems * static_cast<size_t>(p0);
          xa.spaceIn = src + planeReals * static_cast<size_t>(p0);
          ya.zRowOffset += p0;
          ya.numPlanes = p1 - p0;
          ya.planes += planeElems * static_cast<size_t>(p0);
        }
        check_launch(Launch<T>::x(1, xa, s));
        record_stage("x forward");
        if (anySticks) {
          // ---- y stage: y-FFT + scatter into the plane-major sticks / the exchange buffer
          check_launch(Launch<T>::y(1, ya, s));
          record_stage(peer ? "y forward + exchange" : "y forward");
        }
      }
    }
  }
  if (peer) {
    grid_->enqueue_peer_barrier(s);
    record_stage("barrier forward");
  } else if (dist) {
    const ExchangePlan& x = plan_->exchange;
    grid_->communicator()->all_to_all_v(grid_->array_q(), x.planeOffset.data(), x.planeCount.data(),
                                        sticks(), x.stickOffset.data(), x.stickCount.data(),
                                        static_cast<int>(wire_f32() ? sizeof(sb::cx<float>) : sizeof(sb::cx<T>)), s);
    record_stage("exchange forward");
  }
  if (plan_->numStickTiles == 0 || ne == 0) return;  // no local values to produce
  // ---- z stage: z-FFT + compress (+ scaling) (execution_gpu.cpp:291-324)
  const bool outOnDevice = is_device_pointer(output);
  T* outDev = outOnDevice ? output : static_cast<T*>(grid_->array_b());
  {
    auto za = make_z_args<T>(m, geo, plan_->axes, plan_->ptrs, true,
                             static_cast<sb::cx<T>*>(grid_->array_a(parity)), nullptr, outDev,
                             scaling == SPFFT_FULL_SCALING);
    za.wireF32 = wire_f32() ? 1 : 0;
    check_launch(plan_->wfftZ && !za.wireF32 ? Launch<T>::wz(1, za, s) : Launch<T>::z(1, za, s));
    record_stage("z forward");
  }
  if (!outOnDevice) {
    check_gpu(cudaMemcpyAsync(output, outDev, ne * 2 * sizeof(T), cudaMemcpyDeviceToHost, s));
    record_stage("d2h values");
  }
}

template <typename T>
bool TransformEngine<T>::run_batched(bool forward, int n, TransformEngine<T>* const* e, const T* const* in,
                                     T* const* out, const SpfftScalingType* scaling) {
  static const bool enabled = [] {
    const char* v = std::getenv("SPFFT_B200_BATCH");
    return !(v && std::atoi(v) == 0);
  }();
  if (!enabled || n < 2) return false;
  TransformEngine<T>& e0 = *e[0];
  if (!e0.plan_ || e0.plan_->distributed || e0.plan_->fusedXY) return false;
  const IndexMaps& m = *e0.maps_;
  const size_t ne = static_cast<size_t>(m.num_values());
  if (static_cast<size_t>(m.dimX) * m.dimY * m.dimZ == 0 || ne == 0 || e0.plan_->numStickTiles == 0) return false;
  const PlanPointers<T>& pp = e0.plan_->ptrs;
  if (!sb_band_kernel_available(m.dimX, pp.ftwX != nullptr) || !sb_band_kernel_available(m.dimY, pp.ftwY != nullptr) ||
      !sb_band_kernel_available(m.dimZ, pp.ftwZ != nullptr))
    return false;
  for (int i = 0; i < n; ++i) {
    const TransformEngine<T>& ei = *e[i];
    if (ei.plan_ != e0.plan_ || ei.execMode_ != e0.execMode_ || ei.profiling_ ||
        ei.grid_->device_id() != e0.grid_->device_id())
      return false;
    if (!in[i] || !out[i] || !is_device_pointer(in[i]) || !is_device_pointer(out[i])) return false;
    if (forward && scaling[i] != scaling[0]) return false;
  }
  DeviceGuard guard(e0.grid_->device_id());
  // All bands run on the first transform's stream. Ordering with the other transforms' streams
  // goes through the default stream like every call (begin_call / synchronize): each of them
  // waited for its previous work at the end of that call.
  e0.begin_call();
  cudaStream_t s = e0.stream_->get();
  TileMaps geo;
  geo.numStickTiles = e0.plan_->numStickTiles;
  geo.pitch = e0.plan_->pitch;
  geo.numXTiles = e0.plan_->numXTiles;
  geo.symTile = e0.plan_->symTile;
  geo.symLane = e0.plan_->symLane;
  for (int first = 0; first < n; first += sb::kMaxBands) {
    const int nb = std::min(sb::kMaxBands, n - first);
    sb::BandTable<T> bt{};
    for (int b = 0; b < nb; ++b) {
      TransformEngine<T>& eb = *e[first + b];
      bt.sticks[b] = eb.sticks();
      bt.planes[b] = eb.planes();
      if (forward) {
        bt.spaceIn[b] = in[first + b];
        bt.valuesOut[b] = reinterpret_cast<sb::cx<T>*>(out[first + b]);
      } else {
        bt.valuesIn[b] = reinterpret_cast<const sb::cx<T>*>(in[first + b]);
        bt.spaceOut[b] = out[first + b];
      }
    }
    auto za = make_z_args<T>(m, geo, e0.plan_->axes, pp, forward, e0.sticks(), nullptr, nullptr,
                             forward && scaling[0] == SPFFT_FULL_SCALING);
    auto ya = make_y_args<T>(m, geo, e0.plan_->axes, pp, e0.sticks(), e0.planes());
    auto xa = make_x_args<T>(m, e0.plan_->axes, pp, e0.planes(), nullptr, nullptr);
    if (!forward) {
      check_launch(Launch<T>::zb(0, za, bt, nb, s));
      check_launch(Launch<T>::yb(0, ya, bt, nb, s));
      check_launch(Launch<T>::xb(0, xa, bt, nb, s));
    } else {
      check_launch(Launch<T>::xb(1, xa, bt, nb, s));
      check_launch(Launch<T>::yb(1, ya, bt, nb, s));
      check_launch(Launch<T>::zb(1, za, bt, nb, s));
    }
  }
  e0.synchronize();
  return true;
}

template <typename T>
void TransformEngine<T>::backward(const T* input, T* output) {
  enqueue_backward(input, output);
  synchronize();
}

template <typename T>
void TransformEngine<T>::backward(const T* input, SpfftProcessingUnitType outputLocation) {
  backward(input, space_domain_data(outputLocation));
}

template <typename T>
void TransformEngine<T>::forward(const T* input, T* output, SpfftScalingType scaling) {
  enqueue_forward(input, output, scaling);
  synchronize();
}

template <typename T>
void TransformEngine<T>::forward(SpfftProcessingUnitType inputLocation, T* output,
                                 SpfftScalingType scaling) {
  forward(space_domain_data(inputLocation), output, scaling);
}

template class GridResources<double>;
template class GridResources<float>;
template class TransformEngine<double>;
template class TransformEngine<float>;

}  // namespace b200
}  // namespace spfft
