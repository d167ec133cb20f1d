// Host-side helpers shared by the translation units that implement the C ABI (lk_api.cu,
// lk_bert_api.cu): device selection, grow-only device workspaces.
#pragma once

#include <cstring>
#include <thread>

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

// Host memory -> device memory on a stream.  Page-locked sources go straight to cudaMemcpyAsync.  PAGEABLE sources
// (what a caller of the reference's API holds: plain CPU tensors) would be staged by the driver on the calling thread
// at ~11 GB/s; here they are cut into 16 MiB pieces that a few threads copy into one of two page-locked blocks while
// the previous piece is still on the link (the block's event tells when its DMA is done).  The source may be freed
// once the caller has synchronised the stream (pageable pieces have all been copied out by the time upload returns).
struct HostStager {
  static constexpr size_t kPiece = 16u << 20;
  unsigned char* pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  unsigned turn = 0;
  static bool pageable(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
      cudaGetLastError();
      return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
  }
  static void copy_threads(unsigned char* dst, const unsigned char* src, size_t n) {
    unsigned t = std::thread::hardware_concurrency();
    t = t >= 8 ? 4 : (t >= 4 ? 2 : 1);
    if (n < (2u << 20) || t == 1) {
      memcpy(dst, src, n);
      return;
    }
    std::thread th[4];
    const size_t per = (n / t + 63) & ~(size_t)63;
    unsigned started = 1;
    for (; started < t; ++started) {
      const size_t lo = started * per < n ? started * per : n, hi = (started + 1) * per < n ? (started + 1) * per : n;
      try {  // no exception may cross the C ABI: a thread that cannot start leaves its share to this one
        th[started] = std::thread([=] { if (hi > lo) memcpy(dst + lo, src + lo, hi - lo); });
      } catch (...) {
        break;
      }
    }
    memcpy(dst, src, per < n ? per : n);  // share 0, and the shares [started, t) that no thread took
    if (started < t && started * per < n) memcpy(dst + started * per, src + started * per, n - started * per);
    for (unsigned i = 1; i < started; ++i) th[i].join();
  }
  int upload(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes < (1u << 20) || !pageable(src)) {
      LK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
      return LK_OK;
    }
    for (int i = 0; i < 2; ++i) {
      if (!pin[i]) LK_CUDA(cudaHostAlloc((void**)&pin[i], kPiece, cudaHostAllocPortable));
      if (!ev[i]) LK_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    for (size_t off = 0; off < bytes; off += kPiece, ++turn) {
      const size_t n = bytes - off < kPiece ? bytes - off : kPiece;
      const int b = (int)(turn & 1u);
      LK_CUDA(cudaEventSynchronize(ev[b]));  // the DMA that last read this block (a fresh event is complete)
      copy_threads(pin[b], static_cast<const unsigned char*>(src) + off, n);
      LK_CUDA(cudaMemcpyAsync(static_cast<unsigned char*>(dst) + off, pin[b], n, cudaMemcpyHostToDevice, st));
      LK_CUDA(cudaEventRecord(ev[b], st));
    }
    return LK_OK;
  }
  void release() {
    for (int i = 0; i < 2; ++i) {
      if (ev[i]) {
        cudaEventSynchronize(ev[i]);
        cudaEventDestroy(ev[i]);
      }
      if (pin[i]) cudaFreeHost(pin[i]);
      pin[i] = nullptr;
      ev[i] = nullptr;
    }
  }
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
