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
    int k = mk[m];
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
      const int64_t ni = n_rel < k ? n_rel : k;
      for (int64_t j = 0; j < ni; ++j) idcg += disc[j];
      v = idcg != 0.0 ? dcg / idcg : 0.0;
    }
    out[q * n_metrics + m] = v;
  }
}

}  // namespace
}  // namespace lk

using namespace lk;

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
