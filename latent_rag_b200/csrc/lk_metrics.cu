// Retrieval metrics on the device: Recall@k, MRR(@k) and nDCG@k per query.
//
// Replaces the per-query Python of evaluation/retrieval_metrics.py:14-31 (called through
// evaluate_retrieval, :55-96, from main.py:321).  Semantics kept exactly:
//   recall  |set(retrieved[:k]) & set(relevant)| / len(relevant)      (0 when relevant is empty)
//   mrr     1 / rank of the first retrieved[:k] that is relevant      (k = all when not given)
//   ndcg    sum_i [retrieved[i] in relevant] * disc[i] / sum_{i < min(len(relevant), k)} disc[i]
// with disc[i] = 1 / log2(i + 2) supplied by the host in float64 (numpy's own values), and the
// sums taken left to right in float64 like Python's sum(), so the per-query values are the
// reference's bit for bit; the host takes numpy's mean / std over them as the reference does.
#include "lk_common.cuh"

namespace lk {

namespace {

constexpr int kMetricWarps = 4;
constexpr int kMaxRetrieved = 1024;

__global__ void __launch_bounds__(kMetricWarps * 32) retrieval_metrics_kernel(
    const int64_t* __restrict__ retrieved, int64_t q_total, int kr, const int64_t* __restrict__ rel_off,
    const int64_t* __restrict__ rel_ids, const int* __restrict__ kind, const int* __restrict__ mk, int n_metrics,
    const double* __restrict__ disc, double* __restrict__ out) {
  __shared__ unsigned char s_hit[kMetricWarps][kMaxRetrieved];    // retrieved[j] is relevant
  __shared__ unsigned char s_first[kMetricWarps][kMaxRetrieved];  // ... and is its first occurrence
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * kMetricWarps + warp;
  if (q >= q_total) return;
  const int64_t* r = retrieved + q * kr;
  const int64_t lo = rel_off[q], hi = rel_off[q + 1];
  const int64_t n_rel = hi - lo;  // len(relevant), duplicates included like the reference
  unsigned char* hit = s_hit[warp];
  unsigned char* first = s_first[warp];
  for (int j = lane; j < kr; j += 32) {
    const int64_t d = r[j];
    bool h = false;
    if (d >= 0)
      for (int64_t e = lo; e < hi && !h; ++e) h = rel_ids[e] == d;
    bool f = h;
    for (int e = 0; e < j && f; ++e) f = r[e] != d;
    hit[j] = h;
    first[j] = f;
  }
  __syncwarp();
  if (lane != 0) return;  // the sums below are sequential on purpose (left-to-right float64)
  for (int m = 0; m < n_metrics; ++m) {
    const int k_req = mk[m];  // the caller's cut-off: the ideal DCG keeps it even when fewer ids were retrieved
    int k = k_req;
    if (k <= 0 || k > kr) k = kr;
    double v = 0.0;
    if (kind[m] == 0) {  // recall
      int inter = 0;
      for (int j = 0; j < k; ++j) inter += first[j];
      v = n_rel > 0 ? (double)inter / (double)n_rel : 0.0;
    } else if (kind[m] == 1) {  // mrr
      for (int j = 0; j < k; ++j)
        if (hit[j]) {
          v = 1.0 / (double)(j + 1);
          break;
        }
    } else {  // ndcg
      double dcg = 0.0, idcg = 0.0;
      for (int j = 0; j < k; ++j)
        if (r[j] >= 0) dcg += hit[j] ? disc[j] : 0.0;
      // idcg runs over min(len(relevant), k) with the ORIGINAL k (retrieval_metrics.py:29), which may
      // exceed the retrieved length: disc holds max(n_retrieved, largest cut-off) entries
      const int64_t kk = k_req > 0 ? k_req : kr;
      const int64_t ni = n_rel < kk ? n_rel : kk;
      for (int64_t j = 0; j < ni; ++j) idcg += disc[j];
      v = idcg != 0.0 ? dcg / idcg : 0.0;
    }
    out[q * n_metrics + m] = v;
  }
}

// ---- all-pairs rank of the paired ("positive") document, evaluation/embedding_visualization.py:34-37:
//      sim = cosine_similarity(q[:, None], d[None, :]); rank_i = 1 + #{j : sim[i, j] > sim[i, i]}
//      (the reference materialises the [n, n, D] broadcast and argsorts twice).  fp32 like the
//      reference: rows normalised as x / max(|x|, 1e-8), one CTA per query, one warp per document.
constexpr int kRankThreads = 256;

__global__ void __launch_bounds__(kRankThreads) rank_positive_kernel(const float* __restrict__ q,
                                                                     const float* __restrict__ d, int64_t n, int dim,
                                                                     int64_t* __restrict__ out_rank) {
  extern __shared__ float s_q[];  // the normalised query
  __shared__ float s_red[kRankThreads / 32];
  __shared__ float s_self;
  __shared__ int s_count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t i = blockIdx.x;
  float ss = 0.f;
  for (int k = tid; k < dim; k += kRankThreads) {
    const float v = q[i * dim + k];
    s_q[k] = v;
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) s_red[warp] = ss;
  if (tid == 0) s_count = 0;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < kRankThreads / 32; ++w) tot += s_red[w];
  const float q_inv = 1.0f / fmaxf(sqrtf(tot), 1e-8f);
  auto cosine = [&](int64_t j) -> float {  // whole warp; every lane returns the value
    const float* dj = d + j * dim;
    float dot = 0.f, d2 = 0.f;
    for (int k = lane; k < dim; k += 32) {
      const float v = dj[k];
      dot = fmaf(s_q[k] * q_inv, v, dot);
      d2 = fmaf(v, v, d2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
      d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    return dot / fmaxf(sqrtf(d2), 1e-8f);
  };
  if (warp == 0) {
    const float self = cosine(i);
    if (lane == 0) s_self = self;
  }
  __syncthreads();
  const float self = s_self;
  int cnt = 0;
  for (int64_t j = warp; j < n; j += kRankThreads / 32)
    if (j != i) cnt += cosine(j) > self ? 1 : 0;
  if (lane == 0 && cnt) atomicAdd(&s_count, cnt);
  __syncthreads();
  if (tid == 0) out_rank[i] = (int64_t)s_count + 1;
}

}  // namespace
}  // namespace lk

using namespace lk;

extern "C" int lk_rank_positive(int device, const float* queries, const float* docs, int64_t n, int dim,
                                int64_t* out_rank, void* stream) {
  if (n < 0 || dim < 1 || dim > 8192 || (n > 0 && (!queries || !docs || !out_rank))) {
    set_error("lk_rank_positive: bad argument");
    return LK_ERR_INVALID;
  }
  if (n == 0) return LK_OK;
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
  if (device < 0 || device >= cnt) {
    set_error("device %d out of range (%d visible)", device, cnt);
    return LK_ERR_INVALID;
  }
  int prev = -1;
  cudaGetDevice(&prev);
  LK_CUDA(cudaSetDevice(device));
  rank_positive_kernel<<<(unsigned)n, kRankThreads, (size_t)dim * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      queries, docs, n, dim, out_rank);
  cudaError_t le = cudaGetLastError();
  if (prev >= 0) cudaSetDevice(prev);
  count_launch();
  if (le != cudaSuccess) return cuda_fail(le, "rank_positive_kernel", __FILE__, __LINE__);
  return LK_OK;
}

extern "C" int lk_retrieval_metrics(int device, const int64_t* retrieved, int64_t n_queries, int n_retrieved,
                                    const int64_t* rel_offsets, const int64_t* rel_ids, const int* metric_kind,
                                    const int* metric_k, int n_metrics, const double* discounts, double* out,
                                    void* stream) {
  if (n_queries < 0 || n_retrieved < 1 || n_retrieved > kMaxRetrieved || n_metrics < 1 ||
      (n_queries > 0 && (!retrieved || !rel_offsets || !metric_kind || !metric_k || !discounts || !out))) {
    set_error("lk_retrieval_metrics: bad argument (at most %d retrieved ids per query)", kMaxRetrieved);
    return LK_ERR_INVALID;
  }
  if (n_queries == 0) return LK_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
  if (device < 0 || device >= n) {
    set_error("device %d out of range (%d visible)", device, n);
    return LK_ERR_INVALID;
  }
  int prev = -1;
  cudaGetDevice(&prev);
  LK_CUDA(cudaSetDevice(device));
  const unsigned grid = (unsigned)((n_queries + kMetricWarps - 1) / kMetricWarps);
  retrieval_metrics_kernel<<<grid, kMetricWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      retrieved, n_queries, n_retrieved, rel_offsets, rel_ids, metric_kind, metric_k, n_metrics, discounts, out);
  cudaError_t le = cudaGetLastError();
  if (prev >= 0) cudaSetDevice(prev);
  count_launch();
  if (le != cudaSuccess) return cuda_fail(le, "retrieval_metrics_kernel", __FILE__, __LINE__);
  return LK_OK;
}
