// Document-level MaxSim aggregation of chunk candidates, on the device.
//
// Replaces the per-query Python loop of the reference's only production caller
// (main.py:270-282): candidates of a query arrive best first; every candidate row belongs to
// a document (chunk -> doc_id); a document scores the maximum over its chunks and documents are
// ranked by that score, ties in first-seen order (Python's sort is stable), truncated to top_k.
// With the candidates already sorted this is "the first occurrence of every doc id, in order".
#include "lk_common.cuh"

namespace lk {

namespace {

constexpr int kRerankWarps = 4;

// one warp per query; blockDim.x / 32 queries per CTA, cand_k doc ids per warp in dynamic shared memory
__global__ void __launch_bounds__(kRerankWarps * 32) maxsim_rerank_kernel(
    const float* __restrict__ scores, const int64_t* __restrict__ idx, int64_t b, int cand_k,
    const int64_t* __restrict__ row_doc, int64_t n_rows, int top_k, float* __restrict__ out_s,
    int64_t* __restrict__ out_doc) {
  extern __shared__ int64_t s_doc[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= b) return;
  int64_t* doc = s_doc + (size_t)warp * cand_k;
  for (int j = lane; j < cand_k; j += 32) {
    const int64_t r = idx[q * cand_k + j];
    doc[j] = (r >= 0 && r < n_rows) ? row_doc[r] : INT64_MIN;  // INT64_MIN = not a candidate
  }
  __syncwarp();
  int base = 0;  // documents kept so far
  for (int j0 = 0; j0 < cand_k; j0 += 32) {
    const int j = j0 + lane;
    bool keep = false;
    int64_t d = INT64_MIN;
    if (j < cand_k) {
      d = doc[j];
      keep = d != INT64_MIN;
      for (int e = 0; e < j && keep; ++e) keep = doc[e] != d;  // seen earlier = a lower-scoring chunk
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int rank = base + __popc(m & ((1u << lane) - 1u));
    if (keep && rank < top_k) {
      out_doc[q * top_k + rank] = d;
      out_s[q * top_k + rank] = scores[q * cand_k + j];
    }
    base += __popc(m);
  }
  for (int j = base + lane; j < top_k; j += 32) {  // fewer documents than top_k
    out_doc[q * top_k + j] = -1;
    out_s[q * top_k + j] = -INFINITY;
  }
}

}  // namespace
}  // namespace lk

using namespace lk;

extern "C" int lk_maxsim_rerank(int device, const float* cand_scores, const int64_t* cand_idx, int64_t b, int cand_k,
                                const int64_t* row_doc_ids, int64_t n_rows, int top_k, float* out_scores,
                                int64_t* out_doc_ids, void* stream) {
  if (b < 0 || cand_k < 1 || cand_k > kDeepMaxK || top_k < 1 || top_k > cand_k || n_rows < 0 ||
      (b > 0 && (!cand_scores || !cand_idx || !row_doc_ids || !out_scores || !out_doc_ids))) {
    set_error("lk_maxsim_rerank: bad argument");
    return LK_ERR_INVALID;
  }
  if (b == 0) return LK_OK;
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
  const int warps = cand_k <= 1024 ? kRerankWarps : 1;  // at most 32 KB of doc ids per CTA
  const unsigned grid = (unsigned)((b + warps - 1) / warps);
  maxsim_rerank_kernel<<<grid, warps * 32, (size_t)warps * cand_k * sizeof(int64_t), static_cast<cudaStream_t>(stream)>>>(
      cand_scores, cand_idx, b, cand_k, row_doc_ids, n_rows, top_k, out_scores, out_doc_ids);
  cudaError_t le = cudaGetLastError();
  if (prev >= 0) cudaSetDevice(prev);
  count_launch();
  if (le != cudaSuccess) return cuda_fail(le, "maxsim_rerank_kernel", __FILE__, __LINE__);
  return LK_OK;
}
