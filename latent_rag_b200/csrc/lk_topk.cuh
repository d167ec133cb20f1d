// Top-k building blocks.
//
//  * WarpList: a best-first sorted list of k <= 128 (score, index) pairs in shared
//    memory, owned by one warp, with a warp-cooperative insert.  Used where the 32 lanes
//    of a warp look at 32 candidates of the SAME query (SIMT search, merges).
//  * RegTopK : a best-first sorted list held entirely in one thread's registers.  Used
//    where one thread owns one query (the tcgen05 epilogue: TMEM lane = query).
//
// Together they replace torch.topk at retrieval/bruteforce.py:82 (and the heap inside
// faiss IndexFlatIP.search, FAISSEmbeddingRetriever.py:322) without the [B, N] score
// matrix ever existing.
#pragma once

#include "lk_common.cuh"

namespace lk {

template <typename IdxT> struct IdxTraits;
template <> struct IdxTraits<int32_t> {
  static __device__ __forceinline__ int32_t sentinel() { return 0x7fffffff; }
};
template <> struct IdxTraits<int64_t> {
  static __device__ __forceinline__ int64_t sentinel() { return 0x7fffffffffffffffLL; }
};

template <typename IdxT>
__device__ __forceinline__ void warp_list_init(float* s, IdxT* ix, int k, int lane) {
  for (int j = lane; j < k; j += 32) {
    s[j] = -INFINITY;
    ix[j] = IdxTraits<IdxT>::sentinel();
  }
  __syncwarp();
}

// All 32 lanes call with the same (v, id).  No-op when the candidate does not make the list.
template <typename IdxT>
__device__ __forceinline__ void warp_list_insert(float* s, IdxT* ix, int k, float v, IdxT id, int lane) {
  int pos = 0;  // entries that stay ahead of the candidate (a prefix: the list is sorted)
  for (int base = 0; base < k; base += 32) {
    const int j = base + lane;
    const bool ahead = j < k && !better(v, (int64_t)id, s[j], (int64_t)ix[j]);
    pos += __popc(__ballot_sync(0xffffffffu, ahead));
  }
  if (pos >= k) return;
  float ts[kMaxK / 32];
  IdxT ti[kMaxK / 32];
#pragma unroll
  for (int c = 0; c < kMaxK / 32; ++c) {
    const int j = c * 32 + lane;
    if (j >= pos && j < k - 1) {
      ts[c] = s[j];
      ti[c] = ix[j];
    }
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < kMaxK / 32; ++c) {
    const int j = c * 32 + lane;
    if (j >= pos && j < k - 1) {
      s[j + 1] = ts[c];
      ix[j + 1] = ti[c];
    }
  }
  if (lane == 0) {
    s[pos] = v;
    ix[pos] = id;
  }
  __syncwarp();
}

// Each lane offers one candidate (valid or not); the ones that can still make the list
// are inserted one after the other.
template <typename IdxT>
__device__ __forceinline__ void warp_list_offer(float* s, IdxT* ix, int k, float v, IdxT id, bool valid,
                                                int lane) {
  const float thr = s[k - 1];
  unsigned mask = __ballot_sync(0xffffffffu, valid && v >= thr);  // NaN never passes
  while (mask) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    const float cv = __shfl_sync(0xffffffffu, v, src);
    const IdxT ci = __shfl_sync(0xffffffffu, id, src);
    warp_list_insert<IdxT>(s, ix, k, cv, ci, lane);
  }
}

// Sorted (best first) top-K in registers; rows must be offered in ascending index order
// so that a strict '>' keeps the lower index among equal scores.
template <int K>
struct RegTopK {
  float s[K];
  int32_t ix[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      s[j] = -INFINITY;
      ix[j] = 0x7fffffff;
    }
  }
  __device__ __forceinline__ float threshold() const { return s[K - 1]; }
  // precondition: v > threshold()
  __device__ __forceinline__ void insert(float v, int32_t id) {
    s[K - 1] = v;
    ix[K - 1] = id;
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
      const bool up = s[j] > s[j - 1];
      const float a = s[j], b = s[j - 1];
      const int32_t ia = ix[j], ib = ix[j - 1];
      s[j - 1] = up ? a : b;
      s[j] = up ? b : a;
      ix[j - 1] = up ? ia : ib;
      ix[j] = up ? ib : ia;
    }
  }
};

}  // namespace lk
