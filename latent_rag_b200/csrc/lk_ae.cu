// Fused autoencoder encoder forward: z = W1 relu(W0 x + b0) + b1 (+ L2 normalisation),
// hidden activations never leave shared memory.
//
// Replaces the two nn.Linear + ReLU calls of the reference encoders
// (models/denoising_autoencoder.py:19-23,33-34; models/contrastive_autoencoder.py:10-14,
// 23-25 incl. F.normalize; models/variational_autoencoder.py:11-16,27-28, mu half only as
// retrieval/embedder.py:44-45 keeps) -- fp32 FMA arithmetic like the reference.
#include "lk_common.cuh"

namespace lk {

namespace {

constexpr int kAeRows = 32;      // rows of X per CTA
constexpr int kAeThreads = 256;

// w0t: [d_in][d_hidden] (transposed nn.Linear weight), w1t: [d_hidden][d_latent]
__global__ void __launch_bounds__(kAeThreads) ae_encode_kernel(const float* __restrict__ x, int64_t m,
                                                               int d_in, int d_hidden, int d_latent,
                                                               const float* __restrict__ w0t,
                                                               const float* __restrict__ b0,
                                                               const float* __restrict__ w1t,
                                                               const float* __restrict__ b1, int l2norm,
                                                               float* __restrict__ z) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                         // [kAeRows][d_in]
  float* hs = xs + kAeRows * d_in;        // [kAeRows][d_hidden]
  float* zs = hs + kAeRows * d_hidden;    // [kAeRows][d_latent]
  const int t = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * kAeRows;
  const int nr = (int)min((int64_t)kAeRows, m - r0);

  for (int i = t; i < kAeRows * d_in; i += kAeThreads) {
    const int rr = i / d_in, cc = i - rr * d_in;
    xs[i] = rr < nr ? __ldg(x + (r0 + rr) * d_in + cc) : 0.f;
  }
  __syncthreads();

  // layer 0: thread = (row group of 16) x (4 columns strided by 128); columns in passes of 512
  {
    const int cg = t & 127, rg = t >> 7;
    for (int cb = 0; cb < d_hidden; cb += 512) {
      float acc[16][4];
#pragma unroll
      for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
      int col[4];
      bool okc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        col[j] = cb + cg + 128 * j;
        okc[j] = col[j] < d_hidden;
      }
      for (int k = 0; k < d_in; ++k) {
        float w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = okc[j] ? __ldg(w0t + (int64_t)k * d_hidden + col[j]) : 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float xv = xs[(rg * 16 + r) * d_in + k];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[r][j] = fmaf(xv, w[j], acc[r][j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!okc[j]) continue;
        const float bias = __ldg(b0 + col[j]);
#pragma unroll
        for (int r = 0; r < 16; ++r) hs[(rg * 16 + r) * d_hidden + col[j]] = fmaxf(acc[r][j] + bias, 0.f);
      }
    }
  }
  __syncthreads();

  // layer 1: thread = (row group of 8) x (one column of a 64-wide pass)
  {
    const int cl = t & 63, rg = t >> 6;
    for (int cb = 0; cb < d_latent; cb += 64) {
      const int col = cb + cl;
      const bool okc = col < d_latent;
      float acc[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = 0.f;
      for (int k = 0; k < d_hidden; ++k) {
        const float w = okc ? __ldg(w1t + (int64_t)k * d_latent + col) : 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = fmaf(hs[(rg * 8 + r) * d_hidden + k], w, acc[r]);
      }
      if (okc) {
        const float bias = __ldg(b1 + col);
#pragma unroll
        for (int r = 0; r < 8; ++r) zs[(rg * 8 + r) * d_latent + col] = acc[r] + bias;
      }
    }
  }
  __syncthreads();

  // optional F.normalize(z, dim=-1) (eps 1e-12), one warp per 4 rows, then the store
  const int warp = t >> 5, lane = t & 31;
  for (int rr = warp; rr < nr; rr += kAeThreads / 32) {
    float scale = 1.f;
    if (l2norm) {
      float ss = 0.f;
      for (int cidx = lane; cidx < d_latent; cidx += 32) {
        const float v = zs[rr * d_latent + cidx];
        ss = fmaf(v, v, ss);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      scale = fmaxf(sqrtf(ss), 1e-12f);
    }
    for (int cidx = lane; cidx < d_latent; cidx += 32) {
      const float v = zs[rr * d_latent + cidx];
      z[(r0 + rr) * d_latent + cidx] = l2norm ? v / scale : v;
    }
  }
}

}  // namespace

int launch_ae_encode(const float* x, int64_t m, int d_in, int d_hidden, int d_latent, const float* w0t,
                     const float* b0, const float* w1t, const float* b1, int l2norm, float* z,
                     cudaStream_t st) {
  if (m <= 0) return LK_OK;
  const size_t smem = (size_t)kAeRows * (d_in + d_hidden + d_latent) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("ae_encode: d_in + d_hidden + d_latent = %d too large for the fused kernel",
              d_in + d_hidden + d_latent);
    return LK_ERR_UNSUPPORTED;
  }
  LK_CUDA(cudaFuncSetAttribute(ae_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)((m + kAeRows - 1) / kAeRows);
  ae_encode_kernel<<<grid, kAeThreads, smem, st>>>(x, m, d_in, d_hidden, d_latent, w0t, b0, w1t, b1, l2norm, z);
  LK_CHECK_LAUNCH("ae_encode_kernel");
  return LK_OK;
}

}  // namespace lk
