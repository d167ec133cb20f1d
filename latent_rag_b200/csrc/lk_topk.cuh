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
#include "lk_ptx.cuh"

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
//   begin(list_scores, list_idx, valid, k, floor)  start a (query, partial-list) segment
//   threshold() / insert(score, id)      running top-k (insert requires score > threshold())
//   finish(scale)                        leave the list in list_scores / list_idx
// RegSelector keeps a sorted list in registers (k <= 32); BufSelector (kAppend) appends to a
// buffer in the list slot and compacts it cooperatively (k <= 128).
template <int K>
struct RegSelector {
  static constexpr bool kAppend = false;
  RegTopK<K> top;
  float* out_s;
  int32_t* out_i;
  bool valid;
  __device__ __forceinline__ void begin(float* s, int32_t* ix, bool v, int /*k*/, float /*floor*/) {
    out_s = s; out_i = ix; valid = v;
    top.init();
  }
  __device__ __forceinline__ float threshold() const { return top.threshold(); }
  __device__ __forceinline__ void insert(float v, int32_t id) { top.insert(v, id); }
  __device__ __forceinline__ void finish(float scale, bool apply_scale, int /*k*/, int /*lane*/) {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      out_s[j] = apply_scale ? top.s[j] * scale : top.s[j];
      out_i[j] = top.ix[j];
    }
  }
};

// k up to 128: an APPEND BUFFER of kBufCap entries living in the thread's partial-list slot
// in global memory (L2-resident: evict-last stores).  A candidate that beats the current
// threshold is appended with two fire-and-forget stores (no ordering kept, no dependent
// loads).  When a lane's buffer is nearly full the WARP shrinks it cooperatively: the 256
// entries sit in registers (8 per lane, coalesced loads, the next buffer's loads already in
// flight), a pivot t with k <= #{score >= t} <= k + kBufSlack is found by counting against 7
// candidates per pass (they split the known score range [threshold, running max] evenly in
// order-preserving key space, so every pass narrows the range 8x; three packed warp
// reductions per pass), then a stable rewrite keeps the entries at or above the pivot.
// Entries are always in ascending row order (rows arrive ascending, the rewrite is stable), so
// when the pivot is the exact k-th score the surplus entries equal to it are dropped from the
// back: (score desc, index asc) holds without storing an order.  Each compaction multiplies the
// rows seen by ~(1 + room / k), so a list needs only O(log(rows / k)) of them.
constexpr int kBufCap = 256;
constexpr int kBufSlack = 24;
constexpr int kBufPer = kBufCap / 32;

__device__ __forceinline__ uint32_t order_key(float f) {  // monotone float -> uint32
  const uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float order_key_inv(uint32_t k) {
  return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

// entries e = j * 32 + lane of one buffer -> registers (keys of missing entries are 0: below
// every real key)
__device__ __forceinline__ void buffer_load(const float* bs, const int32_t* bi, int n, int lane,
                                            uint32_t (&key)[kBufPer], int32_t (&id)[kBufPer]) {
#pragma unroll
  for (int j = 0; j < kBufPer; ++j) {
    const int e = j * 32 + lane;
    key[j] = 0u;
    id[j] = -1;
    if (e < n) {
      key[j] = order_key(__ldcg(bs + e));
      id[j] = __ldcg(bi + e);
    }
  }
}

// Whole warp, one buffer held in registers: every key is in [lo, hi].  Rewrites the survivors
// in place (stable), returns the pivot key; *kept = survivors.
__device__ __forceinline__ uint32_t buffer_shrink(float* bs, int32_t* bi, int n, int k, uint32_t lo, uint32_t hi,
                                                  int lane, uint64_t keep_policy, const uint32_t (&key)[kBufPer],
                                                  const int32_t (&id)[kBufPer], int* kept) {
  uint32_t t = lo;
  int c_t = n;  // #{key >= t}
#pragma unroll 1
  for (int pass = 0; pass < 16 && c_t > k + kBufSlack && hi > t; ++pass) {
    const uint32_t step = (hi - t) / 8u + 1u;
    int c[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) c[j] = 0;
#pragma unroll
    for (int i = 0; i < kBufPer; ++i)
#pragma unroll
      for (int j = 0; j < 7; ++j) c[j] += key[i] >= t + step * (uint32_t)(j + 1) ? 1 : 0;
    // per-lane counts are <= 8 and warp totals <= 256: three 10-bit fields per reduction
    const unsigned w0 = __reduce_add_sync(0xffffffffu, (unsigned)(c[0] | (c[1] << 10) | (c[2] << 20)));
    const unsigned w1 = __reduce_add_sync(0xffffffffu, (unsigned)(c[3] | (c[4] << 10) | (c[5] << 20)));
    const unsigned w2 = __reduce_add_sync(0xffffffffu, (unsigned)c[6]);
    const int tot[7] = {(int)(w0 & 1023u), (int)((w0 >> 10) & 1023u), (int)(w0 >> 20),
                        (int)(w1 & 1023u), (int)((w1 >> 10) & 1023u), (int)(w1 >> 20), (int)w2};
    uint32_t nt = t, nhi = t + step - 1u;
    int nc = c_t;
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (tot[j] >= k) {  // counts fall as j grows
        nt = t + step * (uint32_t)(j + 1);
        nc = tot[j];
        nhi = j < 6 ? t + step * (uint32_t)(j + 2) - 1u : hi;
      }
    t = nt;
    c_t = nc;
    hi = nhi < hi ? nhi : hi;
  }
  // ties: when more than the slack survive, t is the exact k-th key (hi == t) and only the
  // earliest k - #{key > t} of the entries equal to it are needed
  int need_eq = 1 << 30;
  if (c_t > k + kBufSlack) {
    int cgt = 0;
#pragma unroll
    for (int i = 0; i < kBufPer; ++i) cgt += key[i] > t ? 1 : 0;
    need_eq = k - (int)__reduce_add_sync(0xffffffffu, (unsigned)cgt);
    if (need_eq < 0) need_eq = 0;
  }
  const unsigned lt_mask = (1u << lane) - 1u;
  int base = 0, eq_seen = 0;
#pragma unroll
  for (int j = 0; j < kBufPer; ++j) {  // position order: e = j * 32 + lane
    const bool eq = key[j] == t;
    const unsigned eqm = __ballot_sync(0xffffffffu, eq);
    const bool keep = key[j] > t || (eq && eq_seen + __popc(eqm & lt_mask) < need_eq);
    eq_seen += __popc(eqm);
    const unsigned km = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = base + __popc(km & lt_mask);
      ptx::st_hint(bs + pos, order_key_inv(key[j]), keep_policy);
      ptx::st_hint(bi + pos, id[j], keep_policy);
    }
    base += __popc(km);
  }
  *kept = base;
  return t;
}

// Whole warp: shrink the buffers of (at most `budget`) lanes holding more than `limit` entries, one
// after the other, the next buffer's loads in flight while the current one is processed.
// Returns this lane's (new count, new threshold bits).
static __device__ __noinline__ uint2 buffer_compact_warp(float* bs, int32_t* bi, int cnt, float thr, float vmax, int k,
                                                  int limit, int lane, uint64_t keep_policy, int budget) {
  unsigned todo = __ballot_sync(0xffffffffu, cnt > limit);
  __syncwarp();  // the appends of every lane are visible to the warp
  auto ptr_s = [&](int src) {
    return reinterpret_cast<float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(bs), src));
  };
  auto ptr_i = [&](int src) {
    return reinterpret_cast<int32_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(bi), src));
  };
  uint32_t key[kBufPer], nkey[kBufPer];
  int32_t id[kBufPer], nid[kBufPer];
  int src = todo ? __ffs(todo) - 1 : -1;
  todo &= todo - 1;
  float* ps = nullptr;
  int32_t* pi = nullptr;
  int n = 0;
  if (src >= 0) {
    ps = ptr_s(src);
    pi = ptr_i(src);
    n = __shfl_sync(0xffffffffu, cnt, src);
    buffer_load(ps, pi, n, lane, key, id);
  }
  while (src >= 0) {
    int nsrc = -1, nn = 0;
    float* nps = nullptr;
    int32_t* npi = nullptr;
    if (todo && --budget > 0) {
      nsrc = __ffs(todo) - 1;
      todo &= todo - 1;
      nps = ptr_s(nsrc);
      npi = ptr_i(nsrc);
      nn = __shfl_sync(0xffffffffu, cnt, nsrc);
      buffer_load(nps, npi, nn, lane, nkey, nid);
    }
    const float t_old = __shfl_sync(0xffffffffu, thr, src);
    const uint32_t lo = t_old > -INFINITY ? order_key(t_old) : 0x007fffffu;
    const uint32_t hi = order_key(__shfl_sync(0xffffffffu, vmax, src));
    const int kk = __shfl_sync(0xffffffffu, k, src);
    int kept;
    const uint32_t t = buffer_shrink(ps, pi, n, kk, lo, hi, lane, keep_policy, key, id, &kept);
    if (lane == src) {
      cnt = kept;
      thr = order_key_inv(t);
    }
    src = nsrc;
    ps = nps;
    pi = npi;
    n = nn;
#pragma unroll
    for (int j = 0; j < kBufPer; ++j) {
      key[j] = nkey[j];
      id[j] = nid[j];
    }
  }
  __syncwarp();
  return make_uint2((unsigned)cnt, __float_as_uint(thr));
}

struct BufSelector {
  static constexpr bool kAppend = true;
  float* bs;
  int32_t* bi;
  float thr;    // entries must beat this; every buffered entry scores >= thr
  float vmax;   // upper bound of every buffered score
  int cnt;
  bool valid;
  int k;
  // floor: a score known to be BELOW the query's final k-th best (-inf when unknown)
  __device__ __forceinline__ void begin(float* s, int32_t* ix, bool v, int kk, float floor) {
    bs = s; bi = ix; valid = v; k = kk; cnt = 0;
    vmax = -INFINITY;
    thr = v ? floor : INFINITY;  // lanes without a query never append
  }
  __device__ __forceinline__ float threshold() const { return thr; }
  __device__ __forceinline__ void note_max(float m) { vmax = fmaxf(vmax, m); }
  // precondition: v > thr (and note_max has seen v).  The stores carry an evict-last hint: a
  // list is touched every few microseconds while the corpus streams through L2, and a partially
  // written sector that gets evicted in between costs a DRAM fill plus a write-back per append.
  template <bool HINT>
  __device__ __forceinline__ void append(float v, int32_t id, uint64_t keep_policy) {
    if (HINT) {
      ptx::st_hint(bs + cnt, v, keep_policy);
      ptx::st_hint(bi + cnt, id, keep_policy);
    } else {  // plain stores: the policy operand costs a register-to-uniform move per hinted store
      bs[cnt] = v;
      bi[cnt] = id;
    }
    ++cnt;
  }
  // Whole warp: shrink the buffers of (at most `budget`) lanes holding more than `limit` entries.
  // The work is in a non-inlined function (it is large and rare; the epilogue loop that calls it
  // from several places has to stay inside the instruction cache).
  __device__ __forceinline__ void compact(int limit, int lane, uint64_t keep_policy, int budget = 32) {
    if (!__any_sync(0xffffffffu, cnt > limit)) return;
    const uint2 res = buffer_compact_warp(bs, bi, cnt, thr, vmax, k, limit, lane, keep_policy, budget);
    cnt = (int)res.x;
    thr = __uint_as_float(res.y);
  }
  // Leaves the best entries seen (at least min(k, seen) of them, unordered, at most kBufCap) in
  // slots [0, cnt); the merge kernel does the final selection.
  __device__ __forceinline__ void finish(float scale, bool apply_scale, int /*kk*/, int /*lane*/) {
    if (valid && apply_scale)
      for (int j = 0; j < cnt; ++j) bs[j] = __ldcg(bs + j) * scale;
  }
};

template <int KSEL> struct SelectorFor { using type = RegSelector<KSEL>; };
template <> struct SelectorFor<0> { using type = BufSelector; };  // appends with an L2 evict-last hint
template <> struct SelectorFor<1> { using type = BufSelector; };  // plain appends (thousands of queries)

}  // namespace lk
