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
#include <cstdlib>
#include <cstring>

#include "lk_common.cuh"
#include "lk_planes.cuh"

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

// A lane owns the 8-column chunks lane, lane + 32, ... of a row (hidden % 8 == 0): v[8 c + e] is
// column 8 (lane + 32 c) + e.  Normalises, writes the fp32 row and -- for the linear layer that
// reads it next -- the operand planes.
__device__ __forceinline__ void row_layernorm(float (&v)[kMaxPerLane], int hidden, int lane, const float* g,
                                              const float* b, float eps, float* out, unsigned char* planes,
                                              int64_t row, int n_planes) {
  const int n_chunks = hidden >> 3;
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxPerLane / 8; ++c)
    if (lane + 32 * c < n_chunks)
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[8 * c + e];
  const float mean = warp_sum(s) / (float)hidden;
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxPerLane / 8; ++c)
    if (lane + 32 * c < n_chunks)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[8 * c + e] - mean;
        ss = fmaf(d, d, ss);
      }
  const float rstd = 1.0f / sqrtf(warp_sum(ss) / (float)hidden + eps);
#pragma unroll
  for (int c = 0; c < kMaxPerLane / 8; ++c) {
    const int col = 8 * (lane + 32 * c);
    if (lane + 32 * c < n_chunks) {
      float y[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = (v[8 * c + e] - mean) * rstd * __ldg(g + col + e) + __ldg(b + col + e);
      *reinterpret_cast<float4*>(out + col) = make_float4(y[0], y[1], y[2], y[3]);
      *reinterpret_cast<float4*>(out + col + 4) = make_float4(y[4], y[5], y[6], y[7]);
      if (planes) store_planes8(planes, hidden >> 6, row, col, y, n_planes);
    }
  }
}

__global__ void __launch_bounds__(256) embed_ln_kernel(const int32_t* __restrict__ ids, int64_t n_tok, int s,
                                                       int vocab, int hidden, const float* __restrict__ word,
                                                       const float* __restrict__ pos, const float* __restrict__ type0,
                                                       const float* __restrict__ g, const float* __restrict__ b,
                                                       float eps, float* __restrict__ out,
                                                       unsigned char* __restrict__ planes, int n_planes) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n_tok) return;
  int id = ids[t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float* wr = word + (int64_t)id * hidden;
  const float* pr = pos + (int64_t)(t % s) * hidden;
  float v[kMaxPerLane];
#pragma unroll
  for (int c = 0; c < kMaxPerLane / 8; ++c) {
    const int col = 8 * (lane + 32 * c);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[8 * c + e] = col < hidden ? __ldg(wr + col + e) + __ldg(type0 + col + e) + __ldg(pr + col + e) : 0.f;
  }
  row_layernorm(v, hidden, lane, g, b, eps, out + t * hidden, planes, t, n_planes);
}

__global__ void __launch_bounds__(256) layernorm_kernel(float* __restrict__ x, int64_t n_tok, int hidden,
                                                        const float* __restrict__ g, const float* __restrict__ b,
                                                        float eps, unsigned char* __restrict__ planes, int n_planes) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n_tok) return;
  float* row = x + t * hidden;
  float v[kMaxPerLane];
#pragma unroll
  for (int c = 0; c < kMaxPerLane / 8; ++c) {
    const int col = 8 * (lane + 32 * c);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), d = a;
    if (col < hidden) {
      a = *reinterpret_cast<const float4*>(row + col);
      d = *reinterpret_cast<const float4*>(row + col + 4);
    }
    v[8 * c] = a.x; v[8 * c + 1] = a.y; v[8 * c + 2] = a.z; v[8 * c + 3] = a.w;
    v[8 * c + 4] = d.x; v[8 * c + 5] = d.y; v[8 * c + 6] = d.z; v[8 * c + 7] = d.w;
  }
  row_layernorm(v, hidden, lane, g, b, eps, row, planes, t, n_planes);
}

// qkv [n_tok, 3 * hidden] (Q | K | V, head h at columns h * 32), mask [n_tok] (0 = padding key),
// ctx: the operand planes of the [n_tok, hidden] context (the output projection reads nothing else).
// grid (ceil(s / 64), heads, sentences).
__global__ void __launch_bounds__(kAttnWarps * 32) attention_kernel(const float* __restrict__ qkv,
                                                                    const int32_t* __restrict__ mask, int s,
                                                                    int hidden, unsigned char* __restrict__ ctx,
                                                                    int n_planes) {
  extern __shared__ float attn_smem[];
  float* ks = attn_smem;                       // [s][33]
  float* vs = ks + (size_t)s * (kHeadDim + 1);  // [s][32]
  float* ps = vs + (size_t)s * kHeadDim;        // [warps][s]
  float* cs = ps + (size_t)kAttnWarps * s;      // [warps][32]
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
    cs[warp * kHeadDim + lane] = sum > 0.f ? acc / sum : 0.f;
    __syncwarp();
    if (lane < kHeadDim / 8) {  // one 16-byte chunk per plane and lane
      float y[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = cs[warp * kHeadDim + 8 * lane + e];
      store_planes8(ctx, hidden >> 6, tok0 + qi, h * kHeadDim + 8 * lane, y, n_planes);
    }
    __syncwarp();
  }
}

// Register-tiled attention for 64 <= s <= 256 tokens: one CTA per (128 query rows, head, sentence).
//   phase 1  scores^T[key][query] = K (Q / sqrt(32))^T: every thread owns an 8-query x 16-key tile,
//            operands transposed in shared memory so that a step over the head dimension is six
//            128-bit loads for 128 FMAs
//   phase 2  thread = query row: running max, exp, sum over the keys (masked keys: -inf)
//   phase 3  context = P V: every thread owns 8 queries x 4 head dimensions; V takes the place of
//            the Q / K operands, which are dead by then
// Queries 4g..4g+3 and 64+4g..64+4g+3 belong to thread group g, so the 16 lanes that differ in g
// read and write consecutive 16-byte pieces of a shared-memory row.
constexpr int kTileQ = 128;             // query rows per CTA
constexpr int kQRow = kTileQ + 4;       // row length of the transposed Q operand

__global__ void __launch_bounds__(128) attention_tiled_kernel(const float* __restrict__ qkv,
                                                              const int32_t* __restrict__ mask, int s, int s_pad,
                                                              int hidden, unsigned char* __restrict__ ctx,
                                                              int n_planes) {
  extern __shared__ __align__(16) float attn_smem[];
  const int k_row = s_pad + 4;
  float* st = attn_smem;                            // [s_pad][128]  scores^T, then exp(score - max)
  float* km = st + (size_t)s_pad * kTileQ;          // [s_pad]       0 / -inf per key
  float* inv = km + s_pad;                          // [128]         1 / sum per query
  float* qt = inv + kTileQ;                         // [32][132]     Q^T * scale
  float* kt = qt + kHeadDim * kQRow;                // [32][s_pad+4] K^T
  float* vs = qt;                                   // [s_pad][32]   V (after phase 1)
  const int tid = threadIdx.x;
  const int h = blockIdx.y;
  const int64_t tok0 = (int64_t)blockIdx.z * s;
  const int q0 = blockIdx.x * kTileQ;
  const int ld = 3 * hidden;
  const float scale = rsqrtf((float)kHeadDim);

  // operands: thread -> (row, 4 head dimensions), 128-byte rows read whole.  All the loads of a
  // batch are issued before the first transposed store, so their latencies overlap.
  float4 qv[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int i = tid + 128 * it, q = i >> 3, d4 = (i & 7) * 4;
    qv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + q < s) qv[it] = __ldg(reinterpret_cast<const float4*>(qkv + (tok0 + q0 + q) * ld + h * kHeadDim + d4));
  }
  for (int base = 0; base < s_pad * 8; base += 1024) {
    float4 kv[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int i = base + tid + 128 * it, j = i >> 3, d4 = (i & 7) * 4;
      kv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < s) kv[it] = __ldg(reinterpret_cast<const float4*>(qkv + (tok0 + j) * ld + hidden + h * kHeadDim + d4));
    }
    if (base == 0) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int i = tid + 128 * it, q = i >> 3, d4 = (i & 7) * 4;
        qt[(d4 + 0) * kQRow + q] = qv[it].x * scale;
        qt[(d4 + 1) * kQRow + q] = qv[it].y * scale;
        qt[(d4 + 2) * kQRow + q] = qv[it].z * scale;
        qt[(d4 + 3) * kQRow + q] = qv[it].w * scale;
      }
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int i = base + tid + 128 * it, j = i >> 3, d4 = (i & 7) * 4;
      if (j < s_pad) {
        kt[(d4 + 0) * k_row + j] = kv[it].x;
        kt[(d4 + 1) * k_row + j] = kv[it].y;
        kt[(d4 + 2) * k_row + j] = kv[it].z;
        kt[(d4 + 3) * k_row + j] = kv[it].w;
      }
    }
  }
  for (int j = tid; j < s_pad; j += 128) km[j] = (j < s && mask[tok0 + j] != 0) ? 0.f : -INFINITY;
  __syncthreads();

  // phase 1
  const int n_tiles = 16 * (s_pad >> 4);
  for (int tile = tid; tile < n_tiles; tile += 128) {
    const int qg = tile & 15, kg = tile >> 4;
    float acc[8][16];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < kHeadDim; ++d) {
      const float4 qa = *reinterpret_cast<const float4*>(qt + d * kQRow + 4 * qg);
      const float4 qb = *reinterpret_cast<const float4*>(qt + d * kQRow + 64 + 4 * qg);
      const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
      float kv[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(kt + d * k_row + 16 * kg + 4 * c);
        kv[4 * c] = t.x; kv[4 * c + 1] = t.y; kv[4 * c + 2] = t.z; kv[4 * c + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(qv[i], kv[j], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float* row = st + (size_t)(16 * kg + j) * kTileQ;
      *reinterpret_cast<float4*>(row + 4 * qg) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      *reinterpret_cast<float4*>(row + 64 + 4 * qg) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
    }
  }
  __syncthreads();

  // V over the dead operands (rows past s are zero), then phase 2: thread = query row
  for (int base = 0; base < s_pad * 8; base += 1024) {
    float4 vv[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int i = base + tid + 128 * it, j = i >> 3, d4 = (i & 7) * 4;
      vv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < s) vv[it] = __ldg(reinterpret_cast<const float4*>(qkv + (tok0 + j) * ld + 2 * hidden + h * kHeadDim + d4));
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int i = base + tid + 128 * it, j = i >> 3, d4 = (i & 7) * 4;
      if (j < s_pad) *reinterpret_cast<float4*>(vs + j * kHeadDim + d4) = vv[it];
    }
  }
  {  // four independent chains: the loads of a row's scores are 128 floats apart and latency-bound
    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 2
    for (int j = 0; j < s_pad; j += 4)
#pragma unroll
      for (int u = 0; u < 4; ++u) m4[u] = fmaxf(m4[u], st[(size_t)(j + u) * kTileQ + tid] + km[j + u]);
    const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    const float base = mx > -INFINITY ? mx : 0.f;  // every key masked: exp(-inf - 0) = 0, no NaN
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int j = 0; j < s_pad; j += 4)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float e = expf(st[(size_t)(j + u) * kTileQ + tid] + km[j + u] - base);
        st[(size_t)(j + u) * kTileQ + tid] = e;
        s4[u] += e;
      }
    const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    inv[tid] = sum > 0.f ? 1.0f / sum : 0.f;
  }
  __syncthreads();

  // phase 3
  {
    const int qg = tid & 15, dg = tid >> 4;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
#pragma unroll 4
    for (int j = 0; j < s_pad; ++j) {
      const float4 pa = *reinterpret_cast<const float4*>(st + (size_t)j * kTileQ + 4 * qg);
      const float4 pb = *reinterpret_cast<const float4*>(st + (size_t)j * kTileQ + 64 + 4 * qg);
      const float4 vv = *reinterpret_cast<const float4*>(vs + j * kHeadDim + 4 * dg);
      const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(pv[i], vv.x, acc[i][0]);
        acc[i][1] = fmaf(pv[i], vv.y, acc[i][1]);
        acc[i][2] = fmaf(pv[i], vv.z, acc[i][2]);
        acc[i][3] = fmaf(pv[i], vv.w, acc[i][3]);
      }
    }
    // lanes l and l ^ 16 hold dimensions 8m..8m+3 and 8m+4..8m+7 of the same queries: after the
    // exchange both have the 16-byte chunk; each writes it for four of the eight queries
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = (i < 4 ? 4 * qg + i : 64 + 4 * qg + i - 4);
      const float r = inv[q];
      float mine[4], other[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        mine[e] = acc[i][e] * r;
        other[e] = __shfl_xor_sync(0xffffffffu, mine[e], 16);
      }
      const bool odd = dg & 1;
      if (q0 + q < s && ((i >> 2) & 1) == (int)odd) {
        const float y[8] = {odd ? other[0] : mine[0], odd ? other[1] : mine[1], odd ? other[2] : mine[2],
                            odd ? other[3] : mine[3], odd ? mine[0] : other[0], odd ? mine[1] : other[1],
                            odd ? mine[2] : other[2], odd ? mine[3] : other[3]};
        store_planes8(ctx, hidden >> 6, tok0 + q0 + q, h * kHeadDim + 8 * (dg >> 1), y, n_planes);
      }
    }
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
                         unsigned char* planes, int n_planes, cudaStream_t st) {
  if (n_tok <= 0) return LK_OK;
  embed_ln_kernel<<<(unsigned)((n_tok + 7) / 8), 256, 0, st>>>(ids, n_tok, s, vocab, hidden, word, pos, type0, g, b,
                                                               eps, out, planes, n_planes);
  LK_CHECK_LAUNCH("embed_ln_kernel");
  return LK_OK;
}

int launch_bert_layernorm(float* x, int64_t n_tok, int hidden, const float* g, const float* b, float eps,
                          unsigned char* planes, int n_planes, cudaStream_t st) {
  if (n_tok <= 0) return LK_OK;
  layernorm_kernel<<<(unsigned)((n_tok + 7) / 8), 256, 0, st>>>(x, n_tok, hidden, g, b, eps, planes, n_planes);
  LK_CHECK_LAUNCH("layernorm_kernel");
  return LK_OK;
}

int launch_bert_attention(const float* qkv, const int32_t* mask, int64_t n_sent, int s, int hidden, int heads,
                          unsigned char* ctx_planes, int n_planes, cudaStream_t st) {
  if (n_sent <= 0) return LK_OK;
  if (n_sent > 65535) {
    set_error("attention: %lld sentences per call are too many", (long long)n_sent);
    return LK_ERR_UNSUPPORTED;
  }
  const char* force = getenv("LK_ATTN");  // bring-up: "warp" / "tiled"
  bool tiled = s >= 64 && s <= 256;
  if (force && s <= 256) tiled = !strcmp(force, "tiled");
  if (tiled) {
    const int s_pad = round_up(s, 16);
    const size_t operands = (size_t)kHeadDim * kQRow + (size_t)kHeadDim * (s_pad + 4), v = (size_t)s_pad * kHeadDim;
    const size_t smem = ((size_t)s_pad * kTileQ + s_pad + kTileQ + (operands > v ? operands : v)) * sizeof(float);
    LK_CUDA(cudaFuncSetAttribute(attention_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const dim3 grid((unsigned)((s + kTileQ - 1) / kTileQ), (unsigned)heads, (unsigned)n_sent);
    attention_tiled_kernel<<<grid, 128, smem, st>>>(qkv, mask, s, s_pad, hidden, ctx_planes, n_planes);
    LK_CHECK_LAUNCH("attention_tiled_kernel");
    return LK_OK;
  }
  const size_t smem = ((size_t)s * (kHeadDim + 1) + (size_t)s * kHeadDim + (size_t)kAttnWarps * (s + kHeadDim)) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("attention: %d tokens per sentence are too many", s);
    return LK_ERR_UNSUPPORTED;
  }
  LK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const dim3 grid((unsigned)((s + kAttnQueries - 1) / kAttnQueries), (unsigned)heads, (unsigned)n_sent);
  attention_kernel<<<grid, kAttnWarps * 32, smem, st>>>(qkv, mask, s, hidden, ctx_planes, n_planes);
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
