// transform_engine.hpp -- host side of a transform: work buffers (GridResources), the device
// resident plan shared by clones (DevicePlan) and the stage wiring (TransformEngine).
//
// Re-implements, for the GPU path only, what the reference spreads over
//   GridInternal<T>        src/spfft/grid_internal.cpp:47-262      (buffers, limits)
//   TransformInternal<T>   src/spfft/transform_internal.cpp:45-371 (validation, dispatch)
//   ExecutionGPU<T>        src/execution/execution_gpu.cpp:47-410  (stream/event ordering,
//                                                                   host/device pointer handling)
// The kernels are reached only through the C ABI in launch.h.
#pragma once
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "gpu_runtime.hpp"
#include "index_plan.hpp"
#include "nccl_dyn.hpp"
#include "peer_memory.hpp"
#include "spfft/types.h"
#include "stage_args.hpp"

namespace spfft {
namespace b200 {

template <typename T>
class GridResources {
public:
  GridResources(int maxDimX, int maxDimY, int maxDimZ, int maxNumLocalZSticks,
                SpfftProcessingUnitType processingUnit, int maxNumThreads);
  // Distributed grid over the GPUs of an NCCL communicator (reference: the MPI constructor,
  // grid_internal.cpp:104-206)
  GridResources(int maxDimX, int maxDimY, int maxDimZ, int maxNumLocalZSticks, int maxLocalZLength,
                SpfftProcessingUnitType processingUnit, int maxNumThreads,
                std::shared_ptr<Communicator> comm, SpfftExchangeType exchangeType);
  // same limits, new buffers (reference: GridInternal copy ctor, grid_internal.cpp:208-262)
  GridResources(const GridResources& other);
  GridResources& operator=(const GridResources&) = delete;

  int max_dim_x() const { return maxDimX_; }
  int max_dim_y() const { return maxDimY_; }
  int max_dim_z() const { return maxDimZ_; }
  int max_num_local_z_columns() const { return maxSticks_; }
  int max_num_local_xy_planes() const { return maxPlanes_; }
  SpfftProcessingUnitType processing_unit() const { return pu_; }
  int device_id() const { return deviceId_; }
  int num_threads() const { return numThreads_; }
  bool local() const { return !comm_ || comm_->size() == 1; }
  const std::shared_ptr<Communicator>& communicator() const { return comm_; }
  SpfftExchangeType exchange_type() const { return exchangeType_; }
  // plane-side exchange buffer of distributed transforms
  void* array_q() const { return q_.get(); }
  size_t bytes_q() const { return q_.bytes(); }

  // Device work arrays, capacities in bytes. A holds the plane-major stick buffer (and the real
  // space domain of R2C transforms), B the xy planes / the complex space domain / staged values.
  void* array_a() const { return a_.get(); }
  void* array_b() const { return b_.get(); }
  size_t bytes_a() const { return a_.bytes(); }
  size_t bytes_b() const { return b_.bytes(); }
  // Pinned host mirror of the space domain, allocated on first use.
  void* host_space(size_t bytes);
  // Scratch ring of the fused xy stage and its counters (grown on demand, shared by the
  // transforms of this grid like the other work buffers).
  void* scratch(size_t bytes);
  int* counters(size_t count);

  static size_t stick_capacity(int maxDimZ, int maxSticks);

  // Peer-memory exchange (peer_memory.hpp): true when arrays A and Q and the barrier flags of every
  // rank are mapped into this process. peer_a / peer_q: the mapped buffers in rank order.
  bool peer_exchange() const { return peerOk_; }
  void* peer_a(int rank) const { return peerA_.ptr(rank); }
  void* peer_q(int rank) const { return peerQ_.ptr(rank); }
  // Peer exchange: the buffers the OTHER ranks store into exist twice (A2 / Q2) and alternate from call to call,
  // so that a rank that is already one call ahead writes the copy nobody reads any more: the barrier at the start
  // of every call is not needed (2 instead of 4 barriers per backward + forward pair). parity 0 = A / Q.
  void* array_a(int parity) const { return parity ? a2_.get() : a_.get(); }
  void* array_q(int parity) const { return parity ? q2_.get() : q_.get(); }
  void* peer_a(int parity, int rank) const { return parity ? peerA2_.ptr(rank) : peerA_.ptr(rank); }
  void* peer_q(int parity, int rank) const { return parity ? peerQ2_.ptr(rank) : peerQ_.ptr(rank); }
  // parity of the next backward (0) / forward (1) call on this grid (collective call order = same on all ranks)
  int next_exchange_parity(int forward) { return peerOk_ ? (exchangeCalls_[forward]++ & 1) : 0; }
  // Enqueue a barrier over all ranks on `stream`: every rank's earlier work on its stream
  // (stores into peer memory included) is complete and visible before anything enqueued after the
  // barrier on any rank starts. All ranks must call it in the same sequence.
  void enqueue_peer_barrier(cudaStream_t stream);

private:
  void allocate();
  void map_peers();
  int maxDimX_, maxDimY_, maxDimZ_, maxSticks_, maxPlanes_;
  SpfftProcessingUnitType pu_;
  int deviceId_ = 0;
  int numThreads_;
  DeviceBuffer a_, b_, q_, a2_, q2_, scratch_, counters_;
  std::shared_ptr<Communicator> comm_;
  SpfftExchangeType exchangeType_ = SPFFT_EXCH_DEFAULT;
  PinnedBuffer host_;
  std::mutex hostMutex_;
  DeviceBuffer flags_;
  PeerWindow peerA_, peerQ_, peerA2_, peerQ2_, peerFlags_;
  int exchangeCalls_[2] = {0, 0};
  bool peerOk_ = false;
  int barrierEpoch_ = 0;
};

// Everything the kernels read that depends only on the index set: uploaded once, shared by clones.
template <typename T>
struct DevicePlan {
  AxisPlans axes;
  PlanPointers<T> ptrs;
  // tile geometry (host copies of the scalars of TileMaps)
  int numStickTiles = 0, pitch = 0, numXTiles = 0, symTile = -1, symLane = -1;
  // fused xy stage (wfft_xy.cu): scratch ring geometry; false = separate y and x kernels
  bool fusedXY = false;
  int xyRing = 0, xyLag = 0, xyCounters = 0;
  // warp-FFT kernels (wfft_xy.cu / wfft_z.cu): fused xy stage (implies fusedXY) / z stage
  bool wfftXY = false;
  bool wfftZ = false;
  // distributed transforms: the stick <-> slab exchange (host offsets/counts) and the y-stage
  // tables over all ranks' sticks
  bool distributed = false;
  ExchangePlan exchange;
  const int* srcBase = nullptr;
  const int* srcPitch = nullptr;
  const int* tileBase = nullptr;
  const int* tilePitch = nullptr;
  const unsigned short* distYInv = nullptr;
  // peer-memory form of the exchange (ExchangePlan::rowRank ...)
  const unsigned char* rowRank = nullptr;
  const long long* rowOff = nullptr;
  const unsigned char* stickRank = nullptr;
  const int* fwdBase = nullptr;
  const int* tileFwdBase = nullptr;
  const int* fwdTileOrder = nullptr;
  std::vector<DeviceBuffer> storage;
  size_t deviceBytes = 0;
};

struct StageTime {
  const char* name;
  float ms;
};

template <typename T>
class TransformEngine {
public:
  TransformEngine(SpfftProcessingUnitType executionUnit, std::shared_ptr<GridResources<T>> grid,
                  std::shared_ptr<IndexMaps> maps, std::shared_ptr<DevicePlan<T>> plan = nullptr);
  ~TransformEngine();

  std::shared_ptr<TransformEngine<T>> clone() const;

  // whole transforms (enqueue + synchronize according to the execution mode)
  void backward(const T* input, T* output);
  void backward(const T* input, SpfftProcessingUnitType outputLocation);
  void forward(const T* input, T* output, SpfftScalingType scaling);
  void forward(SpfftProcessingUnitType inputLocation, T* output, SpfftScalingType scaling);
  // enqueue only; used by multi_transform_* (reference: multi_transform_internal.hpp:63-176)
  void enqueue_backward(const T* input, T* output);
  void enqueue_forward(const T* input, T* output, SpfftScalingType scaling);
  void synchronize();
  // Batched multi-transform: all `n` transforms as ONE launch per stage (band_kernels.cu) when they
  // are clones of one local plan, use device pointers and the same execution mode / scaling.
  // Returns false (nothing enqueued) when the set does not qualify; the caller then runs the
  // transforms one by one on their own streams.
  static bool run_batched(bool forward, int n, TransformEngine<T>* const* engines, const T* const* in,
                          T* const* out, const SpfftScalingType* scaling);

  T* space_domain_data(SpfftProcessingUnitType location);

  const IndexMaps& maps() const { return *maps_; }
  const std::shared_ptr<GridResources<T>>& grid() const { return grid_; }
  SpfftProcessingUnitType processing_unit() const { return executionUnit_; }
  SpfftExecType execution_mode() const { return execMode_; }
  void set_execution_mode(SpfftExecType m) { execMode_ = m; }
  bool shared_grid(const TransformEngine<T>& other) const { return grid_ == other.grid_; }
  void* stream() const { return stream_->get(); }

  // distributed transform whose exchange is fused into the stage kernels (peer memory)
  bool uses_peer_exchange() const { return plan_ && plan_->distributed && peer_exchange(); }
  void set_profiling(bool on) { profiling_ = on; }
  std::vector<StageTime> stage_times();

private:
  void begin_call();
  sb::XYArgs<T> make_xy_args(const TileMaps& geo, const T* spaceIn, T* spaceOut, bool forward = false, int parity = 0,
                             int planeBegin = 0, int planeCount = -1);
  sb::YArgs<T> make_y_stage_args(const TileMaps& geo, bool forward, int parity = 0);
  void record_stage(const char* name);
  size_t space_bytes() const;
  T* device_space() const;
  sb::cx<T>* sticks() const { return static_cast<sb::cx<T>*>(grid_->array_a()); }
  sb::cx<T>* planes() const { return static_cast<sb::cx<T>*>(grid_->array_b()); }
  bool peer_exchange() const { return grid_->peer_exchange() && plan_->rowRank && plan_->stickRank; }
  // SPFFT_EXCH_*_FLOAT on a distributed double-precision transform: the exchanged buffers hold cx<float>
  bool wire_f32() const {
    return sizeof(T) == 8 && plan_ && plan_->distributed &&
           (grid_->exchange_type() == SPFFT_EXCH_COMPACT_BUFFERED_FLOAT ||
            grid_->exchange_type() == SPFFT_EXCH_BUFFERED_FLOAT);
  }

  SpfftProcessingUnitType executionUnit_;
  SpfftExecType execMode_ = SPFFT_EXEC_SYNCHRONOUS;
  std::shared_ptr<GridResources<T>> grid_;
  std::shared_ptr<IndexMaps> maps_;
  std::shared_ptr<DevicePlan<T>> plan_;
  std::unique_ptr<Stream> stream_;
  std::unique_ptr<Event> startEvent_, endEvent_;
  // host-pointer calls of local C2C transforms: the space domain moves in slabs of planes on a second stream while
  // the xy stage of the neighbouring slab runs (enqueue_backward / enqueue_forward)
  std::unique_ptr<Stream> copyStream_;
  std::vector<std::unique_ptr<Event>> slabEvents_;
  int host_slabs() const;
  cudaStream_t copy_stream();
  cudaEvent_t slab_event(int i);
  bool profiling_ = false;
  std::vector<std::unique_ptr<Event>> profEvents_;
  std::vector<const char*> profNames_;
  size_t profUsed_ = 0;
};

template <typename T>
std::shared_ptr<DevicePlan<T>> build_device_plan(const IndexMaps& maps, long long smemLimit);

// Parameters of a distributed transform (parameters.cpp:43-140): local index conversion, then the
// stick lists and plane counts of all ranks gathered over the communicator (collective call).
// Throws the exception type of a C error code (no-op for SPFFT_SUCCESS).
void throw_error_code(int code);
// Collective: all ranks pass their local error code; if any is not SPFFT_SUCCESS every rank throws it.
void agree_on_error(Communicator& comm, int localCode);

std::shared_ptr<IndexMaps> make_distributed_index_maps(Communicator& comm, SpfftTransformType type,
                                                       int dimX, int dimY, int dimZ, int localZLength,
                                                       int numLocalElements,
                                                       SpfftIndexFormatType indexFormat,
                                                       const int* indices);

// tile widths (log2 lanes) the stage kernels of a transform use; isFast* tell which axes run the
// register-FFT kernels
void choose_tile_lanes(const IndexMaps& m, int complexBytes, long long smemLimit, AxisPlans& ax,
                       bool& fastX, bool& fastY, bool& fastZ);

}  // namespace b200
}  // namespace spfft
