// Index-build / query-prep kernels: row-major rows -> 128-row operand tiles (+ side
// values), and the Mahalanobis whitening product.
//
// Replaces the host-side preparation of the reference: F.normalize at
// retrieval/bruteforce.py:50,67 (retrieval/common.py:30-32), the per-call squared norms
// at retrieval/bruteforce.py:74-75, and faiss.normalize_L2 at retrieval/common.py:25.
#include "lk_common.cuh"

namespace lk {

namespace {

template <typename T> __device__ __forceinline__ float load_as_float(const T* p);
template <> __device__ __forceinline__ float load_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per row.  OutT = __nv_bfloat16 (8 per 16-byte chunk) or float (4 per chunk).
template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) tile_rows_kernel(const InT* __restrict__ rows, int64_t n, int dim,
                                                        int chunks, int64_t block_bytes,
                                                        unsigned char* __restrict__ tiles,
                                                        float* __restrict__ side, int64_t row0,
                                                        int side_mode, int prenorm) {
  constexpr int E = kChunkBytes / (int)sizeof(OutT);
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const InT* src = rows + r * (int64_t)dim;

  // pass 1: |row|^2 of the values as stored (or of the input when it is pre-normalised)
  float ss = 0.f;
  for (int c = lane; c < dim; c += 32) {
    float x = load_as_float<InT>(src + c);
    if (!prenorm && sizeof(OutT) == 2) x = __bfloat162float(__float2bfloat16_rn(x));
    ss = fmaf(x, x, ss);
  }
  ss = warp_sum(ss);
  const float nrm = fmaxf(sqrtf(ss), 1e-12f);

  // pass 2: scatter the row into its block, one 16-byte chunk per lane per step
  const int64_t g = row0 + r;
  const int rin = (int)(g % kBlockRows);
  unsigned char* blk = tiles + (g / kBlockRows) * block_bytes;
  for (int kc = lane; kc < chunks; kc += 32) {
    alignas(16) OutT v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int c = kc * E + e;
      float x = c < dim ? load_as_float<InT>(src + c) : 0.f;
      if (prenorm) x = x / nrm;
      if (sizeof(OutT) == 2) {
        reinterpret_cast<__nv_bfloat16*>(v)[e] = __float2bfloat16_rn(x);
      } else {
        reinterpret_cast<float*>(v)[e] = x;
      }
    }
    // K block kc>>3, logical chunk kc&7 -> swizzled position inside the slab
    *reinterpret_cast<uint4*>(blk + (int64_t)(kc >> 3) * kSlabBytes + slab_chunk_offset(rin, kc & 7)) =
        *reinterpret_cast<const uint4*>(v);
  }
  if (lane == 0) {
    float s;
    if (prenorm) s = side_mode == 0 ? 1.0f : ss / (nrm * nrm);
    else s = side_mode == 0 ? 1.0f / nrm : ss;
    side[g] = s;
  }
}

// out[r][j] = sum_i x[r][i] * L[i][j], fp64 accumulate.  CTA = 128 threads, 8 rows.
template <typename InT>
__global__ void __launch_bounds__(128) whiten_kernel(const InT* __restrict__ rows, int64_t n, int dim,
                                                     const double* __restrict__ L,
                                                     float* __restrict__ out) {
  extern __shared__ double xs[];  // [8][dim]
  const int64_t r0 = (int64_t)blockIdx.x * 8;
  const int nr = (int)min((int64_t)8, n - r0);
  for (int i = threadIdx.x; i < 8 * dim; i += blockDim.x) {
    const int rr = i / dim, c = i - rr * dim;
    xs[i] = rr < nr ? (double)load_as_float<InT>(rows + (r0 + rr) * dim + c) : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < dim; j += blockDim.x) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < dim; ++i) {
      const double l = L[(int64_t)i * dim + j];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) acc[rr] = fma(xs[rr * dim + i], l, acc[rr]);
    }
    for (int rr = 0; rr < nr; ++rr) out[(r0 + rr) * dim + j] = (float)acc[rr];
  }
}

}  // namespace

int launch_tile_rows(const void* rows, int rows_dtype, int64_t n, const TileGeom& g, void* tiles,
                     float* side, int64_t row0, int side_mode, int prenorm, cudaStream_t st) {
  if (n <= 0) return LK_OK;
  const int warps = 8;
  const unsigned grid = (unsigned)((n + warps - 1) / warps);
  unsigned char* t = static_cast<unsigned char*>(tiles);
#define LK_TILE(IN, OUT)                                                                      \
  tile_rows_kernel<IN, OUT><<<grid, warps * 32, 0, st>>>(static_cast<const IN*>(rows), n, g.dim, \
                                                         g.chunks(), g.block_bytes(), t, side,   \
                                                         row0, side_mode, prenorm)
  if (g.elem_bytes == 2) {
    if (rows_dtype == LK_F32) LK_TILE(float, __nv_bfloat16);
    else LK_TILE(__nv_bfloat16, __nv_bfloat16);
  } else {
    if (rows_dtype == LK_F32) LK_TILE(float, float);
    else LK_TILE(__nv_bfloat16, float);
  }
#undef LK_TILE
  LK_CHECK_LAUNCH("tile_rows_kernel");
  return LK_OK;
}

int launch_whiten(const void* rows, int rows_dtype, int64_t n, int dim, const double* L, float* out,
                  cudaStream_t st) {
  if (n <= 0) return LK_OK;
  const unsigned grid = (unsigned)((n + 7) / 8);
  const size_t smem = (size_t)8 * dim * sizeof(double);
  if (smem > 48 * 1024) {
    set_error("whitening supports dim <= 768 (got %d)", dim);
    return LK_ERR_UNSUPPORTED;
  }
  if (rows_dtype == LK_F32)
    whiten_kernel<float><<<grid, 128, smem, st>>>(static_cast<const float*>(rows), n, dim, L, out);
  else
    whiten_kernel<__nv_bfloat16>
        <<<grid, 128, smem, st>>>(static_cast<const __nv_bfloat16*>(rows), n, dim, L, out);
  LK_CHECK_LAUNCH("whiten_kernel");
  return LK_OK;
}

}  // namespace lk
