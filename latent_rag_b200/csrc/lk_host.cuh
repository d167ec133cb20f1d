// Host-side helpers shared by the translation units that implement the C ABI (lk_api.cu,
// lk_bert_api.cu): device selection, grow-only device workspaces.
#pragma once

#include "lk_common.cuh"

namespace lk {

// restores the caller's current device (torch tracks it) when an API call returns
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return LK_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      want = bytes;
      e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(workspace)", __FILE__, __LINE__);
    cap = want;
    return LK_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

inline int check_device(int device, int* sm_count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
  if (device < 0 || device >= n) {
    set_error("device %d out of range (%d visible)", device, n);
    return LK_ERR_INVALID;
  }
  int major = 0;
  LK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) {
    set_error("device %d has compute capability %d.x; liblatentknn is built for sm_100a only", device, major);
    return LK_ERR_UNSUPPORTED;
  }
  LK_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device));
  return LK_OK;
}

}  // namespace lk
