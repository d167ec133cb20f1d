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

// Selector interfaces used by the tcgen05 epilogue (one thread = one query):
//   begin(list_scores, list_idx, valid)  start a (query, partial-list) segment
//   threshold() / insert(score, id)      running top-k (insert requires score > threshold())
//   finish(scale)                        leave the list in list_scores / list_idx
template <int K>
struct RegSelector {
  RegTopK<K> top;
  float* out_s;
  int32_t* out_i;
  bool valid;
  __device__ __forceinline__ void begin(float* s, int32_t* ix, bool v, int /*k*/) {
    out_s = s; out_i = ix; valid = v;
    top.init();
  }
  __device__ __forceinline__ float threshold() const { return top.threshold(); }
  __device__ __forceinline__ void insert(float v, int32_t id) { top.insert(v, id); }
  __device__ __forceinline__ void finish(float scale, bool apply_scale, int /*k*/) {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      out_s[j] = apply_scale ? top.s[j] * scale : top.s[j];
      out_i[j] = top.ix[j];
    }
  }
};

// k up to 128: a binary heap (root = worst kept entry) living directly in the thread's
// partial-list slot in global memory (L2-resident; touched only on the rare insert).
struct HeapSelector {
  float* hs;
  int32_t* hi;
  float thr;
  bool valid;
  int k;
  __device__ __forceinline__ void begin(float* s, int32_t* ix, bool v, int kk) {
    hs = s; hi = ix; valid = v; k = kk;
    thr = v ? -INFINITY : INFINITY;  // lanes without a query never insert
    if (v)
      for (int j = 0; j < kk; ++j) {
        hs[j] = -INFINITY;
        hi[j] = 0x7fffffff;
      }
  }
  __device__ __forceinline__ float threshold() const { return thr; }
  // precondition: v > thr (rows arrive in ascending index order)
  __device__ __forceinline__ void insert(float v, int32_t id) {
    int i = 0;
    for (;;) {
      const int l = 2 * i + 1;
      if (l >= k) break;
      int c = l;
      float sc = hs[l];
      int32_t ic = hi[l];
      if (l + 1 < k) {  // descend towards the WORSE child: lower score, higher index on ties
        const float sr = hs[l + 1];
        const int32_t ir = hi[l + 1];
        if (sr < sc || (sr == sc && ir > ic)) {
          c = l + 1; sc = sr; ic = ir;
        }
      }
      if (sc < v || (sc == v && ic > id)) {  // child is worse than the new entry: move it up
        hs[i] = sc;
        hi[i] = ic;
        i = c;
      } else {
        break;
      }
    }
    hs[i] = v;
    hi[i] = id;
    thr = hs[0];
  }
  __device__ __forceinline__ void finish(float scale, bool apply_scale, int kk) {
    if (!valid || !apply_scale) return;
    for (int j = 0; j < kk; ++j) hs[j] *= scale;
  }
};

template <int KSEL> struct SelectorFor { using type = RegSelector<KSEL>; };
template <> struct SelectorFor<0> { using type = HeapSelector; };

}  // namespace lk
