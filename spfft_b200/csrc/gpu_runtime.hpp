// gpu_runtime.hpp -- the small CUDA-only runtime layer of the host side: error mapping, RAII for
// device / pinned memory, streams, events, device guard, pointer classification.
//
// Takes the place of the reference's multi-backend wrappers in src/gpu_util/
// (gpu_runtime_api.hpp, gpu_stream_handle.hpp, gpu_event_handle.hpp, gpu_device_guard.hpp,
// gpu_pointer_translation.hpp, gpu_transfer.hpp) and of the containers in src/memory/.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <utility>

#include "spfft/exceptions.hpp"

namespace spfft {
namespace b200 {

// Error mapping of the reference: src/gpu_util/gpu_runtime_api.hpp:113-123
inline void check_gpu(cudaError_t e) {
  if (e == cudaSuccess) return;
  switch (e) {
    case cudaErrorMemoryAllocation: throw GPUAllocationError();
    case cudaErrorLaunchFailure:
    case cudaErrorLaunchOutOfResources:
    case cudaErrorLaunchTimeout: throw GPULaunchError();
    case cudaErrorNoDevice: throw GPUNoDeviceError();
    case cudaErrorInvalidValue: throw GPUInvalidValueError();
    case cudaErrorInvalidDevicePointer: throw GPUInvalidDevicePointerError();
    default: throw GPUError();
  }
}

// Holds a device for the lifetime of the object (the reference's GPUDeviceGuard,
// src/gpu_util/gpu_device_guard.hpp:38-63 -- used there as an unnamed temporary, which does not
// hold; here it is always a named local).
class DeviceGuard {
public:
  explicit DeviceGuard(int device) {
    check_gpu(cudaGetDevice(&previous_));
    if (previous_ != device) {
      check_gpu(cudaSetDevice(device));
      switched_ = true;
    }
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
  ~DeviceGuard() {
    if (switched_) cudaSetDevice(previous_);
  }

private:
  int previous_ = 0;
  bool switched_ = false;
};

class DeviceBuffer {
public:
  DeviceBuffer() = default;
  explicit DeviceBuffer(size_t bytes) { allocate(bytes); }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  DeviceBuffer(DeviceBuffer&& o) noexcept : ptr_(o.ptr_), bytes_(o.bytes_) {
    o.ptr_ = nullptr;
    o.bytes_ = 0;
  }
  DeviceBuffer& operator=(DeviceBuffer&& o) noexcept {
    std::swap(ptr_, o.ptr_);
    std::swap(bytes_, o.bytes_);
    return *this;
  }
  ~DeviceBuffer() { release(); }
  void allocate(size_t bytes) {
    release();
    if (bytes == 0) return;
    if (cudaMalloc(&ptr_, bytes) != cudaSuccess) {
      ptr_ = nullptr;
      cudaGetLastError();  // do not leave the allocation failure sticky
      throw GPUAllocationError();
    }
    bytes_ = bytes;
  }
  void release() {
    if (ptr_) cudaFree(ptr_);
    ptr_ = nullptr;
    bytes_ = 0;
  }
  void* get() const { return ptr_; }
  template <typename U>
  U* as() const {
    return static_cast<U*>(ptr_);
  }
  size_t bytes() const { return bytes_; }

private:
  void* ptr_ = nullptr;
  size_t bytes_ = 0;
};

// Page-locked host memory (the reference pins its HostArrays with cudaHostRegister,
// src/spfft/grid_internal.cpp:85-91).
class PinnedBuffer {
public:
  PinnedBuffer() = default;
  PinnedBuffer(const PinnedBuffer&) = delete;
  PinnedBuffer& operator=(const PinnedBuffer&) = delete;
  ~PinnedBuffer() { release(); }
  void allocate(size_t bytes) {
    release();
    if (bytes == 0) return;
    if (cudaMallocHost(&ptr_, bytes) != cudaSuccess) {
      ptr_ = nullptr;
      cudaGetLastError();
      throw HostAllocationError();
    }
    bytes_ = bytes;
  }
  void release() {
    if (ptr_) cudaFreeHost(ptr_);
    ptr_ = nullptr;
    bytes_ = 0;
  }
  void* get() const { return ptr_; }
  size_t bytes() const { return bytes_; }

private:
  void* ptr_ = nullptr;
  size_t bytes_ = 0;
};

class Stream {
public:
  Stream() { check_gpu(cudaStreamCreateWithFlags(&s_, cudaStreamNonBlocking)); }
  Stream(const Stream&) = delete;
  Stream& operator=(const Stream&) = delete;
  ~Stream() {
    if (s_) cudaStreamDestroy(s_);
  }
  cudaStream_t get() const { return s_; }

private:
  cudaStream_t s_ = nullptr;
};

class Event {
public:
  explicit Event(bool timing = false) {
    check_gpu(cudaEventCreateWithFlags(&e_, timing ? cudaEventDefault : cudaEventDisableTiming));
  }
  Event(const Event&) = delete;
  Event& operator=(const Event&) = delete;
  ~Event() {
    if (e_) cudaEventDestroy(e_);
  }
  cudaEvent_t get() const { return e_; }

private:
  cudaEvent_t e_ = nullptr;
};

// Is `ptr` device memory? Managed memory counts as host, like the reference
// (src/gpu_util/gpu_pointer_translation.hpp:42-71).
inline bool is_device_pointer(const void* ptr) {
  if (ptr == nullptr) return false;
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();  // unregistered host memory reports an error on old drivers
    return false;
  }
  return attr.type == cudaMemoryTypeDevice;
}

}  // namespace b200
}  // namespace spfft
