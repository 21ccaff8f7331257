// kernels.cu -- __global__ entry points of the stage kernels (sm_100a) and their launchers.
// The bodies live in stage_kernels.hpp / fft_tile.hpp.
#include <cuda_runtime.h>

#include <atomic>

#include "fast3_launch.cuh"
#include "fast_launch.cuh"
#include "launch.h"

namespace sb {

constexpr int kThreads = 256;

template <typename T, bool FWD>
__global__ void __launch_bounds__(kThreads) k_z_stage(const __grid_constant__ ZArgs<T> a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* smem = reinterpret_cast<cx<T>*>(smemRaw);
  if (FWD)
    z_forward_body<T>(a, (int)blockIdx.x, Ctx{kThreads}, smem);
  else
    z_backward_body<T>(a, (int)blockIdx.x, Ctx{kThreads}, smem);
}

template <typename T, bool FWD>
__global__ void __launch_bounds__(kThreads) k_y_stage(const __grid_constant__ YArgs<T> a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* smem = reinterpret_cast<cx<T>*>(smemRaw);
  if (FWD)
    y_forward_body<T>(a, (int)blockIdx.x, Ctx{kThreads}, smem);
  else
    y_backward_body<T>(a, (int)blockIdx.x, Ctx{kThreads}, smem);
}

template <typename T, bool FWD>
__global__ void __launch_bounds__(kThreads) k_x_stage(const __grid_constant__ XArgs<T> a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smemRaw[];
  cx<T>* smem = reinterpret_cast<cx<T>*>(smemRaw);
  if (FWD)
    x_forward_body<T>(a, (int)blockIdx.x, Ctx{kThreads}, smem);
  else
    x_backward_body<T>(a, (int)blockIdx.x, Ctx{kThreads}, smem);
}

// Barrier over the ranks of a distributed transform through peer-mapped flags (one int per
// source rank on every rank). Thread r publishes `epoch` in rank r's slot for this rank, then waits
// for rank r's epoch in the local slot. Launched after the kernel whose peer stores it publishes
// (stream order = kernel boundary, so those stores are complete), it replaces the
// cudaStreamSynchronize + MPI_Alltoallv pair of transpose_mpi_compact_buffered_gpu.cpp:186-217.
// Like MPI_Alltoallv the barrier simply WAITS for late peers (it spins with __nanosleep; a rank may
// enter a transform arbitrarily later than the others). Debug knob SPFFT_B200_BARRIER_TIMEOUT_S=<s>
// (> 0, opt-in, default off): trap after that many seconds instead of spinning on, to turn a lost
// rank into an error while debugging.
struct PeerBarrierArgs {
  int* flags[kMaxPeers];  // flags[r] = rank r's flag array (mapped), flags[me] local
  int numRanks, me, epoch;
  unsigned long long timeoutNs;  // 0 = wait for ever (default)
};

__global__ void k_peer_barrier(const __grid_constant__ PeerBarrierArgs a) {
  const int r = (int)threadIdx.x;
  if (r >= a.numRanks) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(a.flags[r] + a.me), "r"(a.epoch) : "memory");
  const int* mine = a.flags[a.me] + r;
  unsigned long long t0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (v - a.epoch >= 0) break;  // wrap-safe comparison
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (a.timeoutNs != 0 && t1 - t0 > a.timeoutNs) __trap();  // opt-in debug knob only
    __nanosleep(200);
  }
}

static std::atomic<long long> g_launches{0};

template <typename Kernel, typename Args>
static int launch(Kernel kernel, const Args& args, long long blocks, size_t smemBytes,
                  cudaStream_t stream) {
  if (blocks <= 0) return 0;
  if (blocks > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  if (smemBytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smemBytes);
    if (e != cudaSuccess) return (int)e;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return launch_stage_kernel(kernel, dim3((unsigned)blocks), kThreads, smemBytes, stream, args);
}

template <typename T>
static int launch_z(int forward, const ZArgs<T>& a, cudaStream_t s) {
  if (a.ftw) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return is_fast3_length(a.nz) ? launch_z_fast3<T>(forward, a, s) : launch_z_fast<T>(forward, a, s);
  }
  const size_t smem = 2 * ((size_t)a.nz << a.log2V) * sizeof(cx<T>);
  return forward ? launch(k_z_stage<T, true>, a, a.numTiles, smem, s)
                 : launch(k_z_stage<T, false>, a, a.numTiles, smem, s);
}
template <typename T>
static int launch_y(int forward, const YArgs<T>& a, cudaStream_t s) {
  if (a.ftw) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return is_fast3_length(a.ny) ? launch_y_fast3<T>(forward, a, s) : launch_y_fast<T>(forward, a, s);
  }
  const size_t smem = 2 * ((size_t)a.ny << a.log2V) * sizeof(cx<T>);
  const long long blocks = (long long)a.numXTiles * a.numPlanes;
  return forward ? launch(k_y_stage<T, true>, a, blocks, smem, s)
                 : launch(k_y_stage<T, false>, a, blocks, smem, s);
}
template <typename T>
static int launch_x(int forward, const XArgs<T>& a, cudaStream_t s) {
  if (a.ftw) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return is_fast3_length(a.nx) ? launch_x_fast3<T>(forward, a, s) : launch_x_fast<T>(forward, a, s);
  }
  const size_t smem = 2 * ((size_t)a.nx << a.log2V) * sizeof(cx<T>);
  const long long blocks = (long long)a.numRowTiles * a.numPlanes;
  return forward ? launch(k_x_stage<T, true>, a, blocks, smem, s)
                 : launch(k_x_stage<T, false>, a, blocks, smem, s);
}

}  // namespace sb

extern "C" {
int sb_launch_z_f64(int forward, const sb::ZArgs<double>* a, void* stream) {
  return sb::launch_z<double>(forward, *a, static_cast<cudaStream_t>(stream));
}
int sb_launch_z_f32(int forward, const sb::ZArgs<float>* a, void* stream) {
  return sb::launch_z<float>(forward, *a, static_cast<cudaStream_t>(stream));
}
int sb_launch_y_f64(int forward, const sb::YArgs<double>* a, void* stream) {
  return sb::launch_y<double>(forward, *a, static_cast<cudaStream_t>(stream));
}
int sb_launch_y_f32(int forward, const sb::YArgs<float>* a, void* stream) {
  return sb::launch_y<float>(forward, *a, static_cast<cudaStream_t>(stream));
}
int sb_launch_x_f64(int forward, const sb::XArgs<double>* a, void* stream) {
  return sb::launch_x<double>(forward, *a, static_cast<cudaStream_t>(stream));
}
int sb_launch_x_f32(int forward, const sb::XArgs<float>* a, void* stream) {
  return sb::launch_x<float>(forward, *a, static_cast<cudaStream_t>(stream));
}
int sb_launch_peer_barrier(int* const* flags, int numRanks, int me, int epoch, void* stream) {
  if (numRanks < 1 || numRanks > sb::kMaxPeers) return (int)cudaErrorInvalidValue;
  sb::PeerBarrierArgs a{};
  for (int r = 0; r < numRanks; ++r) a.flags[r] = flags[r];
  a.numRanks = numRanks;
  a.me = me;
  a.epoch = epoch;
  static const unsigned long long timeoutNs = [] {
    const char* e = getenv("SPFFT_B200_BARRIER_TIMEOUT_S");
    const double sec = e ? atof(e) : 0.0;
    return sec > 0.0 ? (unsigned long long)(sec * 1e9) : 0ull;
  }();
  a.timeoutNs = timeoutNs;
  sb::k_peer_barrier<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
  sb::g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}
void sb_note_launches(int n) { sb::g_launches.fetch_add(n, std::memory_order_relaxed); }
long long sb_launch_count(void) { return sb::g_launches.load(std::memory_order_relaxed); }
int sb_max_dynamic_smem(long long* bytes) {
  int dev = 0, v = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  *bytes = v;
  return (int)e;
}
}
