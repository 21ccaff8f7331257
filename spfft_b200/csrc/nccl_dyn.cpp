// nccl_dyn.cpp -- see nccl_dyn.hpp.
#include "nccl_dyn.hpp"

#include <dlfcn.h>

#include <mutex>

#include "gpu_runtime.hpp"
#include "spfft/exceptions.hpp"

namespace spfft {
namespace b200 {

namespace {

// the subset of nccl.h we call (types reduced to what the ABI needs)
using ncclComm_t = void*;
enum { kNcclSuccess = 0 };
enum { kNcclInt8 = 0, kNcclInt32 = 2 };

struct Api {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  bool ok = false;
};

const Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return;
    auto sym = [&](const char* n) { return dlsym(h, n); };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd &&
           a.Send && a.Recv && a.AllGather;
  });
  if (!a.ok) throw MPISupportError();  // no NCCL on this system: distributed transforms unavailable
  return a;
}

inline void check_nccl(int r) {
  if (r != kNcclSuccess) throw MPIError();  // SPFFT_MPI_ERROR doubles as the NCCL error code
}

}  // namespace

NcclUniqueId Communicator::unique_id() {
  NcclUniqueId id;
  check_nccl(api().GetUniqueId(&id));
  return id;
}

Communicator::Communicator(int numRanks, int rank, const NcclUniqueId& id)
    : size_(numRanks), rank_(rank) {
  if (numRanks < 1 || rank < 0 || rank >= numRanks) throw InvalidParameterError();
  check_gpu(cudaGetDevice(&device_));
  ncclComm_t c = nullptr;
  check_nccl(api().CommInitRank(&c, numRanks, id, rank));
  comm_ = c;
}

Communicator::~Communicator() {
  if (comm_) api().CommDestroy(comm_);
}

std::vector<int> Communicator::all_gather_ints(const int* local, int count) {
  std::vector<int> out(static_cast<size_t>(count) * size_);
  if (count == 0) return out;
  DeviceGuard guard(device_);
  DeviceBuffer send(sizeof(int) * count), recv(sizeof(int) * count * size_);
  check_gpu(cudaMemcpy(send.get(), local, sizeof(int) * count, cudaMemcpyHostToDevice));
  check_nccl(api().AllGather(send.get(), recv.get(), static_cast<size_t>(count), kNcclInt32, comm_,
                             nullptr));
  check_gpu(cudaStreamSynchronize(nullptr));
  check_gpu(cudaMemcpy(out.data(), recv.get(), sizeof(int) * count * size_, cudaMemcpyDeviceToHost));
  return out;
}

std::vector<char> Communicator::all_gather_bytes(const void* local, size_t bytes) {
  std::vector<char> out(bytes * size_);
  if (bytes == 0) return out;
  DeviceGuard guard(device_);
  DeviceBuffer send(bytes), recv(bytes * size_);
  check_gpu(cudaMemcpy(send.get(), local, bytes, cudaMemcpyHostToDevice));
  check_nccl(api().AllGather(send.get(), recv.get(), bytes, kNcclInt8, comm_, nullptr));
  check_gpu(cudaStreamSynchronize(nullptr));
  check_gpu(cudaMemcpy(out.data(), recv.get(), bytes * size_, cudaMemcpyDeviceToHost));
  return out;
}

void Communicator::all_to_all_v(const void* sendBuf, const long long* sendOffset,
                                const long long* sendCount, void* recvBuf,
                                const long long* recvOffset, const long long* recvCount,
                                int elemBytes, cudaStream_t stream) {
  const char* sb = static_cast<const char*>(sendBuf);
  char* rb = static_cast<char*>(recvBuf);
  // own block: plain device copy, ordered on the same stream
  if (sendCount[rank_] > 0) {
    check_gpu(cudaMemcpyAsync(rb + recvOffset[rank_] * elemBytes, sb + sendOffset[rank_] * elemBytes,
                              static_cast<size_t>(sendCount[rank_]) * elemBytes,
                              cudaMemcpyDeviceToDevice, stream));
  }
  if (size_ == 1) return;
  const Api& a = api();
  check_nccl(a.GroupStart());
  // the group is ALWAYS closed, also when a send / recv fails in between (an open group would
  // swallow every later NCCL call of this thread)
  int firstError = kNcclSuccess;
  for (int r = 0; r < size_ && firstError == kNcclSuccess; ++r) {
    if (r == rank_) continue;
    if (sendCount[r] > 0)
      firstError = a.Send(sb + sendOffset[r] * elemBytes, static_cast<size_t>(sendCount[r]) * elemBytes,
                          kNcclInt8, r, comm_, stream);
    if (firstError == kNcclSuccess && recvCount[r] > 0)
      firstError = a.Recv(rb + recvOffset[r] * elemBytes, static_cast<size_t>(recvCount[r]) * elemBytes,
                          kNcclInt8, r, comm_, stream);
  }
  const int endError = a.GroupEnd();
  check_nccl(firstError);
  check_nccl(endError);
}

}  // namespace b200
}  // namespace spfft
