// The element-wise and attention kernels of the sentence encoder (all-MiniLM-L6-v2: the SBERT
// model behind retrieval/embedder.py:35-40 and main.py's corpus / query embedding; a 6-layer
// BERT, hidden 384, 12 heads of 32, FFN 1536, mean pooling + L2 normalisation).  The linear
// layers run on tcgen05 (lk_gemm_umma.cu); everything here is fp32 like the reference.
//
//   embed_ln_kernel    word + position + token-type embedding rows -> LayerNorm          (warp / token)
//   layernorm_kernel   LayerNorm over the hidden dimension, biased variance, in place     (warp / token)
//   attention_kernel   softmax(Q K^T / sqrt(32) + key mask) V for one (sentence, head): K and V
//                      of the head in shared memory, one warp per query row, lane = key for
//                      the scores and lane = head dimension for the context (head dim == 32)
//   pool_kernel        masked mean over the tokens (sum / clamp(count, 1e-9)) and L2
//                      normalisation (x / max(|x|, 1e-12)), one CTA per sentence
#include "lk_common.cuh"

namespace lk {

namespace {

constexpr int kMaxPerLane = 32;  // hidden <= 1024
constexpr int kHeadDim = 32;
constexpr int kAttnWarps = 4;
constexpr int kAttnQueries = 64;  // query rows per CTA

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// v[j] = element lane + 32 j of a row of `hidden` values -> normalised in place
__device__ __forceinline__ void row_layernorm(float (&v)[kMaxPerLane], int hidden, int lane, const float* g,
                                              const float* b, float eps, float* out) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j)
    if (lane + 32 * j < hidden) s += v[j];
  const float mean = warp_sum(s) / (float)hidden;
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j)
    if (lane + 32 * j < hidden) {
      const float d = v[j] - mean;
      ss = fmaf(d, d, ss);
    }
  const float rstd = 1.0f / sqrtf(warp_sum(ss) / (float)hidden + eps);
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j) {
    const int c = lane + 32 * j;
    if (c < hidden) out[c] = (v[j] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
  }
}

__global__ void __launch_bounds__(256) embed_ln_kernel(const int32_t* __restrict__ ids, int64_t n_tok, int s,
                                                       int vocab, int hidden, const float* __restrict__ word,
                                                       const float* __restrict__ pos, const float* __restrict__ type0,
                                                       const float* __restrict__ g, const float* __restrict__ b,
                                                       float eps, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n_tok) return;
  int id = ids[t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float* wr = word + (int64_t)id * hidden;
  const float* pr = pos + (int64_t)(t % s) * hidden;
  float v[kMaxPerLane];
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j) {
    const int c = lane + 32 * j;
    v[j] = c < hidden ? __ldg(wr + c) + __ldg(type0 + c) + __ldg(pr + c) : 0.f;
  }
  row_layernorm(v, hidden, lane, g, b, eps, out + t * hidden);
}

__global__ void __launch_bounds__(256) layernorm_kernel(float* __restrict__ x, int64_t n_tok, int hidden,
                                                        const float* __restrict__ g, const float* __restrict__ b,
                                                        float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n_tok) return;
  float* row = x + t * hidden;
  float v[kMaxPerLane];
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j) {
    const int c = lane + 32 * j;
    v[j] = c < hidden ? row[c] : 0.f;
  }
  row_layernorm(v, hidden, lane, g, b, eps, row);
}

// qkv [n_tok, 3 * hidden] (Q | K | V, head h at columns h * 32), mask [n_tok] (0 = padding key),
// ctx [n_tok, hidden].  grid (ceil(s / 64), heads, sentences).
__global__ void __launch_bounds__(kAttnWarps * 32) attention_kernel(const float* __restrict__ qkv,
                                                                    const int32_t* __restrict__ mask, int s,
                                                                    int hidden, float* __restrict__ ctx) {
  extern __shared__ float attn_smem[];
  float* ks = attn_smem;                       // [s][33]
  float* vs = ks + (size_t)s * (kHeadDim + 1);  // [s][32]
  float* ps = vs + (size_t)s * kHeadDim;        // [warps][s]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int64_t tok0 = (int64_t)blockIdx.z * s;
  const int ld = 3 * hidden;
  for (int i = threadIdx.x; i < s * kHeadDim; i += kAttnWarps * 32) {
    const int j = i >> 5, d = i & 31;
    const float* row = qkv + (tok0 + j) * ld + h * kHeadDim + d;
    ks[j * (kHeadDim + 1) + d] = __ldg(row + hidden);
    vs[j * kHeadDim + d] = __ldg(row + 2 * hidden);
  }
  __syncthreads();
  float* pw = ps + (size_t)warp * s;
  const float scale = rsqrtf((float)kHeadDim);
  const int q_hi = min(s, (int)(blockIdx.x + 1) * kAttnQueries);
  for (int qi = blockIdx.x * kAttnQueries + warp; qi < q_hi; qi += kAttnWarps) {
    const float qd = __ldg(qkv + (tok0 + qi) * ld + h * kHeadDim + lane);
    float mx = -INFINITY;
    for (int j0 = 0; j0 < s; j0 += 32) {
      const int j = j0 + lane, jj = j < s ? j : s - 1;  // every lane takes part in the shuffles
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < kHeadDim; ++d) acc = fmaf(__shfl_sync(0xffffffffu, qd, d), ks[jj * (kHeadDim + 1) + d], acc);
      if (j < s) {
        acc = mask[tok0 + j] != 0 ? acc * scale : -INFINITY;
        pw[j] = acc;
        mx = fmaxf(mx, acc);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < s; j += 32) {
      const float e = mx > -INFINITY ? expf(pw[j] - mx) : 0.f;
      pw[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float acc = 0.f;
    for (int j = 0; j < s; ++j) acc = fmaf(pw[j], vs[j * kHeadDim + lane], acc);
    ctx[(tok0 + qi) * hidden + h * kHeadDim + lane] = sum > 0.f ? acc / sum : 0.f;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) pool_kernel(const float* __restrict__ x, const int32_t* __restrict__ mask,
                                                   int s, int hidden, int normalize, float* __restrict__ out) {
  __shared__ float red[4];
  const int64_t tok0 = (int64_t)blockIdx.x * s;
  float acc[kMaxPerLane / 4];  // hidden <= 1024 over 128 threads
#pragma unroll
  for (int j = 0; j < kMaxPerLane / 4; ++j) acc[j] = 0.f;
  float cnt = 0.f;
  for (int t = 0; t < s; ++t) {
    if (mask[tok0 + t] == 0) continue;
    cnt += 1.f;
    const float* row = x + (tok0 + t) * hidden;
#pragma unroll
    for (int j = 0; j < kMaxPerLane / 4; ++j) {
      const int c = threadIdx.x + 128 * j;
      if (c < hidden) acc[j] += row[c];
    }
  }
  const float den = fmaxf(cnt, 1e-9f);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxPerLane / 4; ++j) {
    acc[j] /= den;
    if (threadIdx.x + 128 * j < hidden) ss = fmaf(acc[j], acc[j], ss);
  }
  float scale = 1.f;
  if (normalize) {
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    scale = 1.0f / fmaxf(sqrtf(red[0] + red[1] + red[2] + red[3]), 1e-12f);
  }
#pragma unroll
  for (int j = 0; j < kMaxPerLane / 4; ++j) {
    const int c = threadIdx.x + 128 * j;
    if (c < hidden) out[(int64_t)blockIdx.x * hidden + c] = acc[j] * scale;
  }
}

}  // namespace

int bert_shape_supported(int hidden, int heads, int ffn) {
  return hidden >= 64 && hidden <= 1024 && hidden % 128 == 0 && heads >= 1 && hidden == heads * kHeadDim &&
         ffn >= 128 && ffn % 128 == 0;
}

int launch_bert_embed_ln(const int32_t* ids, int64_t n_tok, int s, int vocab, int hidden, const float* word,
                         const float* pos, const float* type0, const float* g, const float* b, float eps, float* out,
                         cudaStream_t st) {
  if (n_tok <= 0) return LK_OK;
  embed_ln_kernel<<<(unsigned)((n_tok + 7) / 8), 256, 0, st>>>(ids, n_tok, s, vocab, hidden, word, pos, type0, g, b,
                                                               eps, out);
  LK_CHECK_LAUNCH("embed_ln_kernel");
  return LK_OK;
}

int launch_bert_layernorm(float* x, int64_t n_tok, int hidden, const float* g, const float* b, float eps,
                          cudaStream_t st) {
  if (n_tok <= 0) return LK_OK;
  layernorm_kernel<<<(unsigned)((n_tok + 7) / 8), 256, 0, st>>>(x, n_tok, hidden, g, b, eps);
  LK_CHECK_LAUNCH("layernorm_kernel");
  return LK_OK;
}

int launch_bert_attention(const float* qkv, const int32_t* mask, int64_t n_sent, int s, int hidden, int heads,
                          float* ctx, cudaStream_t st) {
  if (n_sent <= 0) return LK_OK;
  const size_t smem = ((size_t)s * (kHeadDim + 1) + (size_t)s * kHeadDim + (size_t)kAttnWarps * s) * sizeof(float);
  if (smem > 200 * 1024 || n_sent > 65535) {
    set_error("attention: %d tokens per sentence / %lld sentences per call are too many", s, (long long)n_sent);
    return LK_ERR_UNSUPPORTED;
  }
  LK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const dim3 grid((unsigned)((s + kAttnQueries - 1) / kAttnQueries), (unsigned)heads, (unsigned)n_sent);
  attention_kernel<<<grid, kAttnWarps * 32, smem, st>>>(qkv, mask, s, hidden, ctx);
  LK_CHECK_LAUNCH("attention_kernel");
  return LK_OK;
}

int launch_bert_pool(const float* x, const int32_t* mask, int64_t n_sent, int s, int hidden, int normalize, float* out,
                     cudaStream_t st) {
  if (n_sent <= 0) return LK_OK;
  pool_kernel<<<(unsigned)n_sent, 128, 0, st>>>(x, mask, s, hidden, normalize, out);
  LK_CHECK_LAUNCH("pool_kernel");
  return LK_OK;
}

}  // namespace lk
