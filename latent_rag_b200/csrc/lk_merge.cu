// k-way merge of partial top-k lists: per-CTA partials inside one GPU (int32 local row
// ids + the shard's base row) and per-GPU candidates after the all-gather (int64 global
// row ids).  Net-new relative to the reference, which materialises [B, N] scores and
// calls torch.topk (retrieval/bruteforce.py:82); SURVEY.md section 8e.
#include "lk_topk.cuh"

namespace lk {

namespace {

constexpr int kMergeWarps = 4;

// One warp per query.  Candidates are streamed 32 at a time through a WarpList.
template <typename IdxT>
__global__ void __launch_bounds__(kMergeWarps * 32) merge_kernel(const float* __restrict__ ps,
                                                                 const IdxT* __restrict__ pi,
                                                                 int64_t b, int n_cand, int k,
                                                                 int64_t idx_base,
                                                                 float* __restrict__ out_s,
                                                                 int64_t* __restrict__ out_i) {
  __shared__ float ls[kMergeWarps][kMaxK];
  __shared__ IdxT li[kMergeWarps][kMaxK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * kMergeWarps + warp;
  if (q >= b) return;
  float* s = ls[warp];
  IdxT* ix = li[warp];
  warp_list_init<IdxT>(s, ix, k, lane);
  const float* cs = ps + q * n_cand;
  const IdxT* ci = pi + q * n_cand;
  for (int base = 0; base < n_cand; base += 32) {
    const int c = base + lane;
    float v = 0.f;
    IdxT id = -1;
    if (c < n_cand) {
      v = cs[c];
      id = ci[c];
    }
    const bool valid = c < n_cand && id >= 0 && id != IdxTraits<IdxT>::sentinel();
    warp_list_offer<IdxT>(s, ix, k, v, id, valid, lane);
  }
  for (int j = lane; j < k; j += 32) {
    const bool filled = ix[j] != IdxTraits<IdxT>::sentinel();
    out_s[q * k + j] = s[j];
    out_i[q * k + j] = filled ? (int64_t)ix[j] + idx_base : (int64_t)-1;
  }
}

template <typename IdxT>
int launch_merge(const float* ps, const IdxT* pi, int64_t b, int n_lists, int list_len, int k,
                 int64_t idx_base, float* out_s, int64_t* out_i, cudaStream_t st) {
  if (b <= 0) return LK_OK;
  if (k < 1 || k > kMaxK) {
    set_error("merge: k=%d outside 1..%d", k, kMaxK);
    return LK_ERR_INVALID;
  }
  const int64_t n_cand = (int64_t)n_lists * list_len;
  if (n_cand > 0x7fffffff) {
    set_error("merge: too many candidates per query");
    return LK_ERR_INVALID;
  }
  const unsigned grid = (unsigned)((b + kMergeWarps - 1) / kMergeWarps);
  merge_kernel<IdxT><<<grid, kMergeWarps * 32, 0, st>>>(ps, pi, b, (int)n_cand, k, idx_base, out_s, out_i);
  LK_CHECK_LAUNCH("merge_kernel");
  return LK_OK;
}

}  // namespace

int launch_merge_i32(const float* ps, const int32_t* pi, int64_t b, int n_lists, int list_len, int k,
                     int64_t idx_base, float* out_s, int64_t* out_i, cudaStream_t st) {
  return launch_merge<int32_t>(ps, pi, b, n_lists, list_len, k, idx_base, out_s, out_i, st);
}

int launch_merge_i64(const float* ps, const int64_t* pi, int64_t b, int n_lists, int list_len, int k,
                     float* out_s, int64_t* out_i, cudaStream_t st) {
  return launch_merge<int64_t>(ps, pi, b, n_lists, list_len, k, 0, out_s, out_i, st);
}

}  // namespace lk
