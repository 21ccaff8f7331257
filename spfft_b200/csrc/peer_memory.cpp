// peer_memory.cpp -- see peer_memory.hpp.
#include "peer_memory.hpp"

#include <cstring>

#include "gpu_runtime.hpp"

namespace spfft {
namespace b200 {

namespace {
struct Wire {
  cudaIpcMemHandle_t handle;
  int valid;  // 0: could not export, 1: handle valid, 2: the rank has no buffer (nothing to map)
};
}  // namespace

void PeerWindow::open(Communicator& comm, void* local) {
  close();
  const int P = comm.size();
  self_ = comm.rank();
  ptrs_.assign(static_cast<size_t>(P), nullptr);
  Wire mine;
  std::memset(&mine, 0, sizeof(mine));
  if (!local) {
    mine.valid = 2;
  } else if (cudaIpcGetMemHandle(&mine.handle, local) == cudaSuccess) {
    mine.valid = 1;
  } else {
    cudaGetLastError();
  }
  const std::vector<char> all = comm.all_gather_bytes(&mine, sizeof(Wire));
  int good = 1;
  for (int r = 0; r < P; ++r) {
    Wire w;
    std::memcpy(&w, all.data() + static_cast<size_t>(r) * sizeof(Wire), sizeof(Wire));
    if (r == self_) {
      ptrs_[r] = local;
      if (!mine.valid) good = 0;
      continue;
    }
    if (w.valid == 2) continue;  // empty buffer on that rank
    if (!w.valid) {
      good = 0;
      continue;
    }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, w.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      good = 0;
      continue;
    }
    ptrs_[r] = p;
  }
  // all ranks take the same decision
  const std::vector<int> votes = comm.all_gather_ints(&good, 1);
  mapped_ = true;
  for (int v : votes) mapped_ = mapped_ && v != 0;
  if (!mapped_) {
    for (int r = 0; r < P; ++r) {
      if (r != self_ && ptrs_[r]) cudaIpcCloseMemHandle(ptrs_[r]);
      if (r != self_) ptrs_[r] = nullptr;
    }
  }
}

void PeerWindow::close() {
  for (size_t r = 0; r < ptrs_.size(); ++r) {
    if (static_cast<int>(r) != self_ && ptrs_[r]) cudaIpcCloseMemHandle(ptrs_[r]);
  }
  ptrs_.clear();
  mapped_ = false;
}

}  // namespace b200
}  // namespace spfft
