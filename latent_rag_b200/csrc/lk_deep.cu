// Top-k beyond one selector's reach (k > 128): the kernels behind the slab search of
// lk_api.cu.  The reference takes any k (k = min(k, N), retrieval/bruteforce.py:81-82) because
// it ranks a materialised [B, N] score matrix; the fused selectors keep at most 128 entries per
// list, so a deep search cuts the corpus into row slabs, takes the top 128 of every slab with
// the ordinary fused kernel, and folds those lists into the running [B, k] result here.
//
//   deep_merge_kernel      one CTA per query: (running result minus the ids that the incoming
//                          lists re-supply) + the incoming lists -> bitonic sort in shared memory
//                          under the library's (score desc, index asc) order -> best k back into
//                          the result.  Also notes every list's LAST score: a slab with more
//                          than `len` rows whose 128th best still reaches the final k-th score may
//                          hide further rows -- the caller splits it and searches the halves.
//   deep_saturated_kernel  that test, one flag per list.
#include "lk_topk.cuh"

namespace lk {

namespace {

constexpr int kDeepThreads = 1024;
constexpr uint32_t kKeyNegInf = 0x007fffffu;  // order_key(-inf): real candidates score above it
constexpr uint32_t kKeyPosInf = 0xff800000u;  // order_key(+inf): keys above it are NaNs

// a sorts before b
__device__ __forceinline__ bool deep_before(uint32_t ka, int64_t ia, uint32_t kb, int64_t ib) {
  return ka > kb || (ka == kb && ia < ib);
}

__global__ void __launch_bounds__(kDeepThreads) deep_merge_kernel(const float* __restrict__ cs,
                                                                  const int64_t* __restrict__ ci, DeepLists L,
                                                                  int64_t b, int k, int have_res, int n_sort,
                                                                  float* res_s, int64_t* res_i,
                                                                  float* __restrict__ list_last) {
  extern __shared__ __align__(16) unsigned char deep_smem[];
  int64_t* s_idx = reinterpret_cast<int64_t*>(deep_smem);      // [n_sort]
  uint32_t* s_key = reinterpret_cast<uint32_t*>(s_idx + n_sort);  // [n_sort]
  const int tid = threadIdx.x;
  const int64_t q = blockIdx.x;
  const int n_res = have_res ? k : 0;
  const int n_in = n_res + L.n_lists * L.len;

  // 1. load: empty slots and dropped entries become (key 0, id max): they sort behind everything
  for (int i = tid; i < n_sort; i += kDeepThreads) {
    uint32_t key = 0u;
    int64_t id = -1;
    if (i < n_res) {
      key = order_key(res_s[q * k + i]);
      id = res_i[q * k + i];
      for (int r = 0; r < L.n_ranges; ++r)
        if (id >= L.lo[r] && id < L.hi[r]) id = -1;  // comes back through list r
    } else if (i < n_in) {
      const int c = i - n_res, l = c / L.len, e = c - l * L.len;
      const int64_t off = (int64_t)l * L.list_stride + q * L.query_stride + e;
      key = order_key(__ldg(cs + off));
      id = __ldg(ci + off);
    }
    const bool ok = id >= 0 && key > kKeyNegInf && key <= kKeyPosInf;
    s_key[i] = ok ? key : 0u;
    s_idx[i] = ok ? id : INT64_MAX;
  }
  if (list_last != nullptr)
    for (int l = tid; l < L.n_lists; l += kDeepThreads) {
      const int64_t off = (int64_t)l * L.list_stride + q * L.query_stride + (L.len - 1);
      const float v = __ldg(cs + off);
      const uint32_t key = order_key(v);
      const bool full = __ldg(ci + off) >= 0 && key > kKeyNegInf && key <= kKeyPosInf;
      const bool more_rows = L.hi[l] - L.lo[l] > (int64_t)L.len;  // the slab holds rows the list does not
      list_last[(int64_t)l * b + q] = full && more_rows ? v : NAN;
    }

  // 2. bitonic sort, best first
  for (int size = 2; size <= n_sort; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = tid; t < (n_sort >> 1); t += kDeepThreads) {
        const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1)), j = i + stride;
        const uint32_t ka = s_key[i], kb = s_key[j];
        const int64_t ia = s_idx[i], ib = s_idx[j];
        const bool up = (i & size) == 0;  // this run is sorted best first
        if (up ? deep_before(kb, ib, ka, ia) : deep_before(ka, ia, kb, ib)) {
          s_key[i] = kb; s_idx[i] = ib;
          s_key[j] = ka; s_idx[j] = ia;
        }
      }
    }
  __syncthreads();

  // 3. the best k (every read of the old result happened in step 1)
  for (int j = tid; j < k; j += kDeepThreads) {
    const uint32_t key = j < n_sort ? s_key[j] : 0u;
    res_s[q * k + j] = key ? order_key_inv(key) : -INFINITY;
    res_i[q * k + j] = key ? s_idx[j] : (int64_t)-1;
  }
}

__global__ void __launch_bounds__(256) deep_saturated_kernel(const float* __restrict__ list_last, int64_t b,
                                                             const float* __restrict__ res_s, int k,
                                                             int* __restrict__ flags) {
  const int l = blockIdx.y;
  const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
  bool sat = false;
  if (q < b) {
    const float last = list_last[(int64_t)l * b + q];  // NaN: the list already holds every row of its slab
    sat = last >= res_s[q * k + (k - 1)];              // -inf while fewer than k candidates exist
  }
  if (__syncthreads_or(sat) && threadIdx.x == 0) flags[l] = 1;
}

}  // namespace

int deep_merge_capacity(int k, int have_res, int list_len) {
  const int n = (kDeepMaxCand - (have_res ? k : 0)) / list_len;
  return n < kDeepMaxLists ? n : kDeepMaxLists;
}

int launch_deep_merge(const float* cs, const int64_t* ci, const DeepLists& L, int64_t b, int k, int have_res,
                      float* res_s, int64_t* res_i, float* list_last, cudaStream_t st) {
  if (b <= 0) return LK_OK;
  const int64_t n_in = (int64_t)(have_res ? k : 0) + (int64_t)L.n_lists * L.len;
  if (k < 1 || k > kDeepMaxK || L.n_lists < 0 || L.len < 1 || n_in > kDeepMaxCand || L.n_ranges < 0 ||
      L.n_ranges > kDeepMaxLists || (list_last != nullptr && L.n_ranges != L.n_lists)) {
    set_error("deep merge: bad geometry (k=%d, %d lists of %d, %d ranges)", k, L.n_lists, L.len, L.n_ranges);
    return LK_ERR_INVALID;
  }
  int n_sort = 2;
  while (n_sort < n_in) n_sort <<= 1;
  const size_t smem = (size_t)n_sort * (sizeof(int64_t) + sizeof(uint32_t));
  LK_CUDA(cudaFuncSetAttribute(deep_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(kDeepMaxCand * (sizeof(int64_t) + sizeof(uint32_t)))));
  deep_merge_kernel<<<(unsigned)b, kDeepThreads, smem, st>>>(cs, ci, L, b, k, have_res, n_sort, res_s, res_i,
                                                             list_last);
  LK_CHECK_LAUNCH("deep_merge_kernel");
  return LK_OK;
}

int launch_deep_saturated(const float* list_last, int n_lists, int64_t b, const float* res_s, int k, int* flags,
                          cudaStream_t st) {
  if (b <= 0 || n_lists <= 0) return LK_OK;
  const dim3 grid((unsigned)((b + 255) / 256), (unsigned)n_lists);
  deep_saturated_kernel<<<grid, 256, 0, st>>>(list_last, b, res_s, k, flags);
  LK_CHECK_LAUNCH("deep_saturated_kernel");
  return LK_OK;
}

}  // namespace lk
