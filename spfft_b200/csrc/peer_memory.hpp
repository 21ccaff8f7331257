// peer_memory.hpp -- device buffers of the other ranks of a communicator mapped into this process
// (CUDA IPC), so that kernels store straight into a peer GPU's memory over NVLink / NVSwitch.
//
// This is what replaces the reference's pack kernel + MPI_Alltoallv + unpack kernel
// (src/transpose/transpose_mpi_compact_buffered_gpu.cpp:163-287) on one NVSwitch box: the z stage
// (backward) and the y stage (forward) write their results directly where the next stage of the
// destination rank reads them, and a flag barrier through the same mapped memory orders the two.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <vector>

#include "nccl_dyn.hpp"

namespace spfft {
namespace b200 {

// One buffer per rank, all mapped. Construction is collective over the communicator and never
// throws for "peer access unavailable": ok() tells whether every rank mapped every peer.
class PeerWindow {
public:
  PeerWindow() = default;
  PeerWindow(const PeerWindow&) = delete;
  PeerWindow& operator=(const PeerWindow&) = delete;
  ~PeerWindow() { close(); }

  // `local` must be the base address of a cudaMalloc allocation (or nullptr for an empty buffer:
  // every rank must then pass nullptr, sizes agree by construction of the grid). Collective.
  void open(Communicator& comm, void* local);
  void close();

  bool mapped() const { return mapped_; }
  void* ptr(int rank) const { return ptrs_[static_cast<size_t>(rank)]; }
  const std::vector<void*>& ptrs() const { return ptrs_; }

private:
  std::vector<void*> ptrs_;
  int self_ = 0;
  bool mapped_ = false;
};

}  // namespace b200
}  // namespace spfft
