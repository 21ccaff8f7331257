// nccl_dyn.hpp -- NCCL reached through dlopen, so that the library loads (and the single-GPU path
// works) on systems without NCCL, and shares the copy a host application (e.g. PyTorch) already
// loaded. Replaces the reference's MPI plumbing (src/mpi_util/*, MPI_Alltoallv in
// src/transpose/transpose_mpi_compact_buffered_gpu.cpp:210-217, MPI_Allgather in
// src/parameters/parameters.cpp:89) for the GPUs of one NVSwitch box.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <memory>
#include <vector>

namespace spfft {
namespace b200 {

struct NcclUniqueId {
  char internal[128];
};

// One communicator = this rank's view of the group (ncclComm_t) on a fixed device.
class Communicator {
public:
  // collective over all ranks: every rank passes the id obtained by rank 0 from unique_id()
  Communicator(int numRanks, int rank, const NcclUniqueId& id);
  ~Communicator();
  Communicator(const Communicator&) = delete;
  Communicator& operator=(const Communicator&) = delete;

  static NcclUniqueId unique_id();

  int size() const { return size_; }
  int rank() const { return rank_; }
  int device_id() const { return device_; }

  // plan-time helpers (blocking): gather `count` ints from every rank
  std::vector<int> all_gather_ints(const int* local, int count);
  // same for `bytes` raw bytes per rank (IPC handles of the peer-memory exchange)
  std::vector<char> all_gather_bytes(const void* local, size_t bytes);

  // The stick<->slab exchange: for every peer r send `sendCount[r]` elements of `elemBytes` from
  // sendBuf + sendOffset[r] and receive recvCount[r] into recvBuf + recvOffset[r] (offsets and
  // counts in elements), as ONE grouped NCCL operation on `stream`; the block to itself is a
  // device copy on the same stream.
  void all_to_all_v(const void* sendBuf, const long long* sendOffset, const long long* sendCount,
                    void* recvBuf, const long long* recvOffset, const long long* recvCount,
                    int elemBytes, cudaStream_t stream);

private:
  void* comm_ = nullptr;
  int size_ = 1, rank_ = 0, device_ = 0;
};

}  // namespace b200
}  // namespace spfft
