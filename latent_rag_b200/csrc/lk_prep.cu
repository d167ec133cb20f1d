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

// fp32 tiles -> split-bf16 planes of the same rows.  One warp per row; a lane converts 8 consecutive
// elements (two fp32 chunks -> one bf16 chunk per plane) per step.
__global__ void __launch_bounds__(256) planes_from_tiles_kernel(const unsigned char* __restrict__ t32, int kblocks32,
                                                                int64_t block_bytes32, int64_t blk0, int64_t n_rows,
                                                                int n_kb, int64_t block_bytes_p,
                                                                unsigned char* __restrict__ planes) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int64_t blk = blk0 + r / kBlockRows;
  const int rin = (int)(r % kBlockRows);
  const unsigned char* src = t32 + blk * block_bytes32;
  unsigned char* dst = planes + blk * block_bytes_p;
  for (int kc = lane; kc < n_kb * 8; kc += 32) {  // bf16 chunk kc = elements 8 kc .. 8 kc + 7
    float x[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c4 = kc * 2 + h;  // fp32 chunk (4 elements): K block c4 >> 3, logical chunk c4 & 7
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((c4 >> 3) < kblocks32)
        v = *reinterpret_cast<const float4*>(src + (int64_t)(c4 >> 3) * kSlabBytes + slab_chunk_offset(rin, c4 & 7));
      x[4 * h] = v.x; x[4 * h + 1] = v.y; x[4 * h + 2] = v.z; x[4 * h + 3] = v.w;
    }
    alignas(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      hi[e] = __float2bfloat16_rn(x[e]);
      lo[e] = __float2bfloat16_rn(x[e] - __bfloat162float(hi[e]));
    }
    const int64_t off = (int64_t)(kc >> 3) * kSlabBytes + slab_chunk_offset(rin, kc & 7);
    *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + (int64_t)n_kb * kSlabBytes + off) = *reinterpret_cast<const uint4*>(lo);
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

// The same product for a FEW rows (a query batch): the kernel above gives all dim columns of 8 rows to one CTA of
// 128 threads -- 9216 dependent fp64 FMAs per thread and the whole L (1.2 MB at dim 384) through one SM, ~150 us
// for a single query.  Here a CTA takes 8 rows x 32 columns (thread = column x row pair), so one query spreads
// over dim / 32 CTAs and finishes in a few microseconds.
template <typename InT>
__global__ void __launch_bounds__(128) whiten_few_kernel(const InT* __restrict__ rows, int64_t n, int dim,
                                                         const double* __restrict__ L, float* __restrict__ out) {
  extern __shared__ double xs[];  // [8][dim]
  const int64_t r0 = (int64_t)blockIdx.x * 8;
  const int nr = (int)min((int64_t)8, n - r0);
  for (int i = threadIdx.x; i < 8 * dim; i += blockDim.x) {
    const int rr = i / dim, c = i - rr * dim;
    xs[i] = rr < nr ? (double)load_as_float<InT>(rows + (r0 + rr) * dim + c) : 0.0;
  }
  __syncthreads();
  const int j = blockIdx.y * 32 + (threadIdx.x & 31);
  const int ra = (threadIdx.x >> 5) * 2;  // this thread's two rows
  if (j >= dim) return;
  double a0 = 0.0, a1 = 0.0;
  for (int i = 0; i < dim; ++i) {  // same summation order as whiten_kernel: bit-identical results
    const double l = L[(int64_t)i * dim + j];
    a0 = fma(xs[ra * dim + i], l, a0);
    a1 = fma(xs[(ra + 1) * dim + i], l, a1);
  }
  if (ra < nr) out[(r0 + ra) * dim + j] = (float)a0;
  if (ra + 1 < nr) out[(r0 + ra + 1) * dim + j] = (float)a1;
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

int launch_planes_from_tiles(const void* tiles32, const TileGeom& g32, int64_t blk0, int64_t n_blocks,
                             const TileGeom& gp, void* planes, cudaStream_t st) {
  if (n_blocks <= 0) return LK_OK;
  const int64_t n_rows = n_blocks * kBlockRows;
  const unsigned grid = (unsigned)((n_rows + 7) / 8);
  planes_from_tiles_kernel<<<grid, 256, 0, st>>>(static_cast<const unsigned char*>(tiles32), g32.kblocks,
                                                 g32.block_bytes(), blk0, n_rows, gp.kblocks / 2, gp.block_bytes(),
                                                 static_cast<unsigned char*>(planes));
  LK_CHECK_LAUNCH("planes_from_tiles_kernel");
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
  if (n <= 4096) {  // a query batch: spread the columns over CTAs as well
    const dim3 g2(grid, (unsigned)((dim + 31) / 32));
    if (rows_dtype == LK_F32)
      whiten_few_kernel<float><<<g2, 128, smem, st>>>(static_cast<const float*>(rows), n, dim, L, out);
    else
      whiten_few_kernel<__nv_bfloat16>
          <<<g2, 128, smem, st>>>(static_cast<const __nv_bfloat16*>(rows), n, dim, L, out);
    LK_CHECK_LAUNCH("whiten_few_kernel");
    return LK_OK;
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
