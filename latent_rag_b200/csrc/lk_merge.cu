// k-way merge of partial top-k lists: per-CTA partials inside one GPU (int32 local row
// ids + the shard's base row) and per-GPU candidates after the all-gather (int64 global
// row ids).  Net-new relative to the reference, which materialises [B, N] scores and
// calls torch.topk (retrieval/bruteforce.py:82); SURVEY.md section 8e.
#include <cstdlib>
#include <cstring>

#include "lk_topk.cuh"

namespace lk {

namespace {

constexpr int kMergeWarps = 8;

// Candidates are streamed 32 at a time through a WarpList (any order: the (score, index) total
// order makes the result independent of it).  WPQ = 1: one warp per query (large batches).
// WPQ = kMergeWarps: one CTA per query, every warp folds a slice of the candidates into its own
// list and warp 0 folds those lists (few queries with many partial lists, e.g. batch 1).
// List l of query q occupies [q][l][0, list_len) of arrays with `list_stride` entries per list.
template <typename IdxT, int WPQ>
__global__ void __launch_bounds__(kMergeWarps * 32) merge_kernel(const float* __restrict__ ps,
                                                                 const IdxT* __restrict__ pi,
                                                                 const int* __restrict__ pc,
                                                                 int64_t b, int n_lists, int list_len,
                                                                 int list_stride, int k, int64_t idx_base,
                                                                 float* __restrict__ out_s,
                                                                 int64_t* __restrict__ out_i) {
  __shared__ float ls[kMergeWarps][kMaxK];
  __shared__ IdxT li[kMergeWarps][kMaxK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = WPQ == 1 ? (int64_t)blockIdx.x * kMergeWarps + warp : (int64_t)blockIdx.x;
  if (q >= b) return;  // block-uniform when WPQ > 1
  float* s = ls[warp];
  IdxT* ix = li[warp];
  warp_list_init<IdxT>(s, ix, k, lane);
  const int n_cand = n_lists * list_len;
  const float* cs = ps + q * n_lists * list_stride;
  const IdxT* ci = pi + q * n_lists * list_stride;
  for (int base = (WPQ == 1 ? 0 : warp * 32); base < n_cand; base += 32 * WPQ) {
    const int c = base + lane;
    float v = 0.f;
    IdxT id = -1;
    if (c < n_cand) {
      const int l = c / list_len, e = c - l * list_len;
      if (pc == nullptr || e < pc[q * n_lists + l]) {  // slots past a list's fill count are not initialised
        v = cs[(int64_t)l * list_stride + e];
        id = ci[(int64_t)l * list_stride + e];
      }
    }
    const bool valid = c < n_cand && id >= 0 && id != IdxTraits<IdxT>::sentinel();
    warp_list_offer<IdxT>(s, ix, k, v, id, valid, lane);
  }
  if (WPQ > 1) {
    __syncthreads();
    if (warp != 0) return;
    for (int w = 1; w < WPQ; ++w)
      for (int base = 0; base < k; base += 32) {
        const int j = base + lane;
        const float v = j < k ? ls[w][j] : 0.f;
        const IdxT id = j < k ? li[w][j] : (IdxT)-1;
        warp_list_offer<IdxT>(s, ix, k, v, id, j < k && id != IdxTraits<IdxT>::sentinel(), lane);
      }
  }
  for (int j = lane; j < k; j += 32) {
    const bool filled = ix[j] != IdxTraits<IdxT>::sentinel();
    out_s[q * k + j] = s[j];
    out_i[q * k + j] = filled ? (int64_t)ix[j] + idx_base : (int64_t)-1;
  }
}

// ---------------------------------------------------------------------------------------
// Selection merge: one CTA per query, for few queries with MANY candidates (batch 1 over 296
// per-CTA lists, k = 100 append buffers ...).  Insertion is serial in the number of candidates
// that beat the running threshold; here the CTA instead
//   1. reads every filled slot once (per-list fill counts skip the empty tails) and stages the
//      order-preserving keys of the real candidates in shared memory,
//   2. finds a pivot t with k <= #{key >= t} <= kSelCap by counting, 3 key bits per pass over
//      the staged keys, starting at the highest bit in which the keys differ,
//   3. gathers those few entries (one more read of the lists) and ranks them all-pairs under
//      (score desc, index asc).
// If the keys do not fit in shared memory the passes re-read the lists from L2 instead.
// ---------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;
constexpr int kSelCap = 512;
constexpr int kSelMaxKeys = 48 * 1024;        // staged keys (192 KB of dynamic shared memory)
constexpr uint32_t kKeyNegInf = 0x007fffffu;  // order_key(-inf): real candidates score above it
constexpr uint32_t kKeyPosInf = 0xff800000u;  // order_key(+inf): keys above it are NaNs

template <typename IdxT>
__global__ void __launch_bounds__(kSelThreads) merge_select_kernel(const float* __restrict__ ps,
                                                                   const IdxT* __restrict__ pi,
                                                                   const int* __restrict__ pc, int64_t b,
                                                                   int n_lists, int list_len, int list_stride,
                                                                   int k, int64_t idx_base, int key_cap,
                                                                   float* __restrict__ out_s,
                                                                   int64_t* __restrict__ out_i) {
  extern __shared__ uint32_t s_keys[];
  __shared__ int s_cnt[8];
  __shared__ uint32_t s_min, s_max;
  __shared__ int s_n;
  __shared__ uint32_t c_key[kSelCap];
  __shared__ IdxT c_idx[kSelCap];
  __shared__ float ls[kMaxK];
  __shared__ IdxT li[kMaxK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int64_t q = blockIdx.x;
  const float* cs = ps + q * n_lists * list_stride;
  const IdxT* ci = pi + q * n_lists * list_stride;
  const int* cc = pc ? pc + q * n_lists : nullptr;
  // f(key, slot offset, is_candidate) for every filled slot, warp-converged (one list per warp turn)
  // (four 32-entry strides per turn: the 8 loads of a lane are in flight together -- with one stride per turn a
  // warp paid one L2 round trip per 32 candidates, ~40 in a row for 38 k candidates of a single query)
  auto for_each_slot = [&](auto&& f) {
    for (int l = warp; l < n_lists; l += kSelThreads / 32) {
      int n = list_len;
      if (cc) n = min(n, __ldcg(cc + l));
      const int64_t base = (int64_t)l * list_stride;
      for (int e0 = 0; e0 < n; e0 += 128) {
        float sv[4];
        IdxT iv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + 32 * u + lane;
          sv[u] = e < n ? __ldcg(cs + base + e) : 0.f;
          iv[u] = e < n ? __ldcg(ci + base + e) : (IdxT)-1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (e0 + 32 * u >= n) break;  // warp-uniform
          const int e = e0 + 32 * u + lane;
          uint32_t key = 0u;
          bool ok = false;
          if (e < n) {
            key = order_key(sv[u]);
            ok = key > kKeyNegInf && key <= kKeyPosInf && iv[u] >= 0 && iv[u] != IdxTraits<IdxT>::sentinel();
          }
          f(key, base + e, ok);
        }
      }
    }
  };
  if (tid == 0) {
    s_min = 0xffffffffu;
    s_max = 0u;
    s_n = 0;
  }
  if (tid < 8) s_cnt[tid] = 0;
  __syncthreads();

  // 1. stage the keys, find their range
  {
    uint32_t mn = 0xffffffffu, mx = 0u;
    for_each_slot([&](uint32_t key, int64_t, bool ok) {
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (m) {
        int pos = 0;
        if (lane == 0) pos = atomicAdd(&s_n, __popc(m));
        pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & lt_mask);
        if (ok) {
          if (pos < key_cap) s_keys[pos] = key;
          mn = min(mn, key);
          mx = max(mx, key);
        }
      }
    });
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) {
      atomicMin(&s_min, mn);
      atomicMax(&s_max, mx);
    }
  }
  __syncthreads();
  const uint32_t kmin = s_min, kmax = s_max;
  const int n_valid = s_n;
  const bool staged = n_valid <= key_cap;
  auto for_each_key = [&](auto&& f) {
    if (staged) {
#pragma unroll 4
      for (int i = tid; i < n_valid; i += kSelThreads) f(s_keys[i]);
    } else {
      for_each_slot([&](uint32_t key, int64_t, bool ok) {
        if (ok) f(key);
      });
    }
  };

  // 2. pivot search
  uint32_t t = 0u;
  int cge = n_valid;
  if (n_valid > kSelCap) {
    int bit = 31 - __clz(kmin ^ kmax);  // -1: every key is the same
    if (bit < 0) {
      t = kmin;
    } else {
      t = bit == 31 ? 0u : (kmin >> (bit + 1)) << (bit + 1);
      while (cge > kSelCap && bit >= 0) {
        const int sh = bit >= 2 ? bit - 2 : 0;
        const int nb = bit - sh + 1;  // bits decided this pass (3, or fewer at the bottom)
        int c[7] = {0, 0, 0, 0, 0, 0, 0};
        for_each_key([&](uint32_t key) {
#pragma unroll
          for (int j = 0; j < 7; ++j) c[j] += key >= (t | ((uint32_t)(j + 1) << sh)) ? 1 : 0;
        });
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const int v = __reduce_add_sync(0xffffffffu, c[j]);
          if (lane == 0 && v) atomicAdd(&s_cnt[j], v);
        }
        __syncthreads();
        int best = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j)
          if (j + 1 < (1 << nb) && s_cnt[j] >= k) best = j + 1;  // counts fall as j grows
        if (best) {
          t |= (uint32_t)best << sh;
          cge = s_cnt[best - 1];
        }
        __syncthreads();
        if (tid < 8) s_cnt[tid] = 0;
        __syncthreads();
        bit = sh - 1;
      }
    }
  }
  __syncthreads();  // everyone has read s_n (n_valid)

  if (cge <= kSelCap) {
    // 3. gather the survivors and rank them all-pairs
    if (tid == 0) s_n = 0;
    __syncthreads();
    for_each_slot([&](uint32_t key, int64_t o, bool ok) {
      const bool take = ok && key >= t;
      const unsigned m = __ballot_sync(0xffffffffu, take);
      if (m) {
        int pos = 0;
        if (lane == 0) pos = atomicAdd(&s_n, __popc(m));
        pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & lt_mask);
        if (take) {
          c_key[pos] = key;
          c_idx[pos] = __ldcg(ci + o);
        }
      }
    });
    __syncthreads();
    const int n = s_n;
    for (int i = tid; i < n; i += kSelThreads) {
      const uint32_t ki = c_key[i];
      const IdxT ii = c_idx[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const uint32_t kj = c_key[j];
        rank += (kj > ki || (kj == ki && c_idx[j] < ii)) ? 1 : 0;
      }
      if (rank < k) {
        out_s[q * k + rank] = order_key_inv(ki);
        out_i[q * k + rank] = (int64_t)ii + idx_base;
      }
    }
    for (int j = n + tid; j < k; j += kSelThreads) {  // fewer than k candidates exist
      out_s[q * k + j] = -INFINITY;
      out_i[q * k + j] = -1;
    }
    return;
  }
  // More than kSelCap candidates tie at the exact pivot (duplicate rows): everything above the
  // pivot plus the lowest-indexed ties, by insertion, one warp.
  if (tid >= 32) return;
  warp_list_init<IdxT>(ls, li, k, lane);
  for (int l = 0; l < n_lists; ++l) {
    int n = list_len;
    if (cc) n = min(n, __ldcg(cc + l));
    const int64_t base = (int64_t)l * list_stride;
    for (int e0 = 0; e0 < n; e0 += 32) {
      const int e = e0 + lane;
      uint32_t key = 0u;
      IdxT id = -1;
      if (e < n) {
        key = order_key(__ldcg(cs + base + e));
        id = __ldcg(ci + base + e);
      }
      const bool ok = key > kKeyNegInf && key <= kKeyPosInf && id >= 0 && id != IdxTraits<IdxT>::sentinel();
      warp_list_offer<IdxT>(ls, li, k, order_key_inv(key), id, ok && key >= t, lane);
    }
  }
  for (int j = lane; j < k; j += 32) {
    const bool filled = li[j] != IdxTraits<IdxT>::sentinel();
    out_s[q * k + j] = ls[j];
    out_i[q * k + j] = filled ? (int64_t)li[j] + idx_base : (int64_t)-1;
  }
}


// Merge of P SORTED lists of k entries per query (the second level of the two-level selection below; every list is
// sorted under the total order (score desc, id asc), unfilled slots (id < 0) at its end): by RANKING -- the final
// position of entry j of list r is j + the number of better entries in each other list, one binary search per other
// list, all threads of the CTA.  P * k <= 2048.
constexpr int kSortedMaxCand = 2048;
__global__ void __launch_bounds__(256) merge_sorted_kernel(const float* __restrict__ ps, const int64_t* __restrict__ pi,
                                                           int parts, int k, float* __restrict__ out_s,
                                                           int64_t* __restrict__ out_i) {
  __shared__ float cs[kSortedMaxCand];
  __shared__ int64_t ci[kSortedMaxCand];
  const int tid = threadIdx.x;
  const int64_t q = blockIdx.x;
  const int n = parts * k;
  for (int e = tid; e < n; e += 256) {
    float v = __ldcg(ps + q * n + e);
    int64_t id = __ldcg(pi + q * n + e);
    if (id < 0 || !(v == v)) {
      v = -INFINITY;
      id = INT64_MAX;
    }
    cs[e] = v;
    ci[e] = id;
  }
  for (int j = tid; j < k; j += 256) {
    out_s[q * k + j] = -INFINITY;
    out_i[q * k + j] = -1;
  }
  __syncthreads();
  for (int e = tid; e < n; e += 256) {
    const int r = e / k, j = e - r * k;
    const float v = cs[e];
    const int64_t id = ci[e];
    if (id == INT64_MAX) continue;
    int pos = j;
    for (int r2 = 0; r2 < parts && pos < k; ++r2) {
      if (r2 == r) continue;
      const float* l2s = cs + r2 * k;
      const int64_t* l2i = ci + r2 * k;
      int lo = 0, hi = k;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (better(l2s[mid], l2i[mid], v, id)) lo = mid + 1;
        else hi = mid;
      }
      pos += lo;
    }
    if (pos < k) {
      out_s[q * k + pos] = v;
      out_i[q * k + pos] = id;
    }
  }
}

template <typename IdxT>
int launch_merge(const float* ps, const IdxT* pi, const int* pc, int64_t b, int n_lists, int list_len,
                 int list_stride, int k, int64_t idx_base, float* out_s, int64_t* out_i, cudaStream_t st,
                 float* tmp_s = nullptr, int64_t* tmp_i = nullptr, int64_t tmp_entries = 0) {
  if (b <= 0) return LK_OK;
  if (k < 1 || k > kMaxK) {
    set_error("merge: k=%d outside 1..%d", k, kMaxK);
    return LK_ERR_INVALID;
  }
  if (list_len < 1 || list_stride < list_len || (int64_t)n_lists * list_stride > 0x7fffffff) {
    set_error("merge: bad list geometry (%d lists, len %d, stride %d)", n_lists, list_len, list_stride);
    return LK_ERR_INVALID;
  }
  const int64_t n_cand = (int64_t)n_lists * list_len;
  const char* force = getenv("LK_MERGE");  // bring-up: "insert" / "select"
  bool select = b <= 1024 && n_cand >= 2048;
  if (force) select = !strcmp(force, "select");
  // A few queries with tens of thousands of candidates each (one query over the 296 append buffers of a top-100
  // search): one CTA per query leaves the GPU empty and takes ~100 us.  Two levels instead: the lists of a query
  // are cut into P contiguous parts, P CTAs select each part's top-k (the same kernel on b * P pseudo-queries:
  // the list arrays are contiguous per query, so part p of query q is pseudo-query q * P + p), and a second,
  // small merge folds the P sorted lists.
  if (select && tmp_s != nullptr && tmp_i != nullptr && n_cand >= 16384 && b * 2 <= 148) {
    int parts = 0;
    for (int cand = 16; cand >= 2; cand >>= 1)
      if (n_lists % cand == 0 && b * cand <= 148 && (int64_t)(n_lists / cand) * list_len >= 2048 &&
          b * cand * (int64_t)k <= tmp_entries) {
        parts = cand;
        break;
      }
    if (parts > 1) {
      const int64_t n_part_cand = (int64_t)(n_lists / parts) * list_len;
      const int key_cap = (int)(n_part_cand < kSelMaxKeys ? n_part_cand : kSelMaxKeys);
      LK_CUDA(cudaFuncSetAttribute(merge_select_kernel<IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(kSelMaxKeys * sizeof(uint32_t))));
      merge_select_kernel<IdxT><<<(unsigned)(b * parts), kSelThreads, (size_t)key_cap * sizeof(uint32_t), st>>>(
          ps, pi, pc, b * parts, n_lists / parts, list_len, list_stride, k, idx_base, key_cap, tmp_s, tmp_i);
      LK_CHECK_LAUNCH("merge_select_kernel");
      merge_sorted_kernel<<<(unsigned)b, 256, 0, st>>>(tmp_s, tmp_i, parts, k, out_s, out_i);
      LK_CHECK_LAUNCH("merge_sorted_kernel");
      return LK_OK;
    }
  }
  if (select) {
    const int key_cap = (int)(n_cand < kSelMaxKeys ? n_cand : kSelMaxKeys);
    const size_t smem = (size_t)key_cap * sizeof(uint32_t);
    LK_CUDA(cudaFuncSetAttribute(merge_select_kernel<IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(kSelMaxKeys * sizeof(uint32_t))));
    merge_select_kernel<IdxT><<<(unsigned)b, kSelThreads, smem, st>>>(ps, pi, pc, b, n_lists, list_len, list_stride,
                                                                      k, idx_base, key_cap, out_s, out_i);
  } else if (b <= 2048 && n_cand >= 8 * 32) {
    merge_kernel<IdxT, kMergeWarps><<<(unsigned)b, kMergeWarps * 32, 0, st>>>(ps, pi, pc, b, n_lists, list_len,
                                                                               list_stride, k, idx_base, out_s, out_i);
  } else {
    const unsigned grid = (unsigned)((b + kMergeWarps - 1) / kMergeWarps);
    merge_kernel<IdxT, 1><<<grid, kMergeWarps * 32, 0, st>>>(ps, pi, pc, b, n_lists, list_len, list_stride, k,
                                                             idx_base, out_s, out_i);
  }
  LK_CHECK_LAUNCH("merge_kernel");
  return LK_OK;
}

}  // namespace

int launch_merge_i32(const float* ps, const int32_t* pi, const int* pc, int64_t b, int n_lists, int list_len,
                     int list_stride, int k, int64_t idx_base, float* out_s, int64_t* out_i, cudaStream_t st,
                     float* tmp_s, int64_t* tmp_i, int64_t tmp_entries) {
  return launch_merge<int32_t>(ps, pi, pc, b, n_lists, list_len, list_stride, k, idx_base, out_s, out_i, st, tmp_s,
                               tmp_i, tmp_entries);
}

int launch_merge_i64(const float* ps, const int64_t* pi, int64_t b, int n_lists, int list_len, int k,
                     float* out_s, int64_t* out_i, cudaStream_t st) {
  return launch_merge<int64_t>(ps, pi, nullptr, b, n_lists, list_len, list_len, k, 0, out_s, out_i, st);
}

}  // namespace lk
