// Exact search on the CUDA cores: fp32 FMA dot products straight off the tiled corpus
// (bf16 or fp32 storage) + per-query warp-cooperative top-k.  This is the bit-faithful
// fp32 path (storage LK_F32 reproduces the reference's fp32 arithmetic up to summation
// order), the path for k > 32, and the cross-check for the tcgen05 kernel.
//
// Replaces retrieval/bruteforce.py:66-82: `q @ emb.T` (+ the euclidean expansion at
// :73-76) followed by torch.topk -- fused, so the [B, N] score matrix never exists.
#include "lk_topk.cuh"

namespace lk {

namespace {

template <typename ElemT> struct Chunk;
template <> struct Chunk<float> {
  static constexpr int E = 4;
  static __device__ __forceinline__ void unpack(const uint4& u, float* x) {
    x[0] = __uint_as_float(u.x); x[1] = __uint_as_float(u.y);
    x[2] = __uint_as_float(u.z); x[3] = __uint_as_float(u.w);
  }
};
template <> struct Chunk<__nv_bfloat16> {
  static constexpr int E = 8;
  static __device__ __forceinline__ void unpack(const uint4& u, float* x) {
    // bf16 -> fp32 is a 16-bit shift
    x[0] = __uint_as_float(u.x << 16); x[1] = __uint_as_float(u.x & 0xffff0000u);
    x[2] = __uint_as_float(u.y << 16); x[3] = __uint_as_float(u.y & 0xffff0000u);
    x[4] = __uint_as_float(u.z << 16); x[5] = __uint_as_float(u.z & 0xffff0000u);
    x[6] = __uint_as_float(u.w << 16); x[7] = __uint_as_float(u.w & 0xffff0000u);
  }
};

// the 8 chunks of a row's 128-byte line are fetched by 8 consecutive requests of the same
// thread, so let the line live in L1 between them
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// grid = (n_slices, ceil(B / QG)); 128 threads; thread t owns row t of each row block of
// the slice, warp g owns the top-k list of query g of the group.
template <typename ElemT, int QG>
__global__ void __launch_bounds__(kBlockRows) simt_search_kernel(SearchArgs a, int blocks_per_slice) {
  constexpr int E = Chunk<ElemT>::E;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int dim_pad = a.g.dim_pad;
  float* qs = reinterpret_cast<float*>(smem_raw);                 // [QG][dim_pad]
  float* sc = qs + QG * dim_pad;                                  // [QG][128]
  float* ls = sc + QG * kBlockRows;                               // [QG][k]
  int32_t* li = reinterpret_cast<int32_t*>(ls + QG * a.k);        // [QG][k]

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int slice = blockIdx.x;
  const int64_t q0 = (int64_t)blockIdx.y * QG;
  const int64_t nblk = (a.n_rows + kBlockRows - 1) / kBlockRows;
  const int64_t blk_lo = (int64_t)slice * blocks_per_slice;
  const int64_t blk_hi = min(nblk, blk_lo + blocks_per_slice);
  const int64_t block_bytes = a.g.block_bytes();

  // stage the group's queries as fp32 (gather out of the query tiles)
  const unsigned char* qt = static_cast<const unsigned char*>(a.q_tiles);
  for (int i = t; i < QG * dim_pad; i += kBlockRows) {
    const int g = i / dim_pad, c = i - g * dim_pad;
    const int64_t q = q0 + g;
    float v = 0.f;
    if (q < a.n_queries) {
      const int gc = c / E;  // global 16-byte chunk of the row
      const unsigned char* p = qt + (q / kBlockRows) * block_bytes + (int64_t)(gc >> 3) * kSlabBytes +
                               slab_chunk_offset((int)(q % kBlockRows), gc & 7) + (c % E) * sizeof(ElemT);
      if (sizeof(ElemT) == 2) v = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p));
      else v = *reinterpret_cast<const float*>(p);
    }
    qs[i] = v;
  }
  float qside[QG];
#pragma unroll
  for (int g = 0; g < QG; ++g) qside[g] = (q0 + g < a.n_queries) ? a.q_side[q0 + g] : 0.f;
  if (warp < QG) warp_list_init<int32_t>(ls + warp * a.k, li + warp * a.k, a.k, lane);
  __syncthreads();

  const unsigned char* tiles = static_cast<const unsigned char*>(a.tiles);
  const int kblocks = a.g.kblocks;
  const int sw = t & 7;  // this row's chunk swizzle
  for (int64_t blk = blk_lo; blk < blk_hi; ++blk) {
    const unsigned char* row = tiles + blk * block_bytes + t * kRowBytes;
    float acc[QG];
#pragma unroll
    for (int g = 0; g < QG; ++g) acc[g] = 0.f;
    for (int kb = 0; kb < kblocks; ++kb) {
      const uint4* src = reinterpret_cast<const uint4*>(row + (int64_t)kb * kSlabBytes);
      uint4 u[8];  // the row's whole 128-byte line of this K block: 8 independent loads;
                   // logical chunk c sits at physical position c ^ (row & 7)
#pragma unroll
      for (int c = 0; c < 8; ++c) u[c] = ld_stream(src + (c ^ sw));
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float x[E];
        Chunk<ElemT>::unpack(u[c], x);
#pragma unroll
        for (int g = 0; g < QG; ++g) {
          const float* qg = qs + g * dim_pad + (kb * 8 + c) * E;
#pragma unroll
          for (int e = 0; e < E; ++e) acc[g] = fmaf(x[e], qg[e], acc[g]);
        }
      }
    }
    // epilogue of the reference formulas; rows past the end carry side = NaN -> NaN score
    const float sd = a.side[blk * kBlockRows + t];
#pragma unroll
    for (int g = 0; g < QG; ++g) {
      float s;
      if (a.metric == LK_COSINE) s = acc[g] * sd * qside[g];          // bruteforce.py:66-69
      else s = fmaf(2.0f, acc[g], -(qside[g] + sd));                  // -(q2 + e2 - 2 q.e), :73-76
      sc[g * kBlockRows + t] = s;
    }
    __syncthreads();
    if (warp < QG && q0 + warp < a.n_queries) {
      float* s = ls + warp * a.k;
      int32_t* ix = li + warp * a.k;
#pragma unroll
      for (int c = 0; c < kBlockRows / 32; ++c) {
        const float v = sc[warp * kBlockRows + c * 32 + lane];
        const int32_t id = (int32_t)(blk * kBlockRows + c * 32 + lane);
        warp_list_offer<int32_t>(s, ix, a.k, v, id, true, lane);
      }
    }
    __syncthreads();
  }

  if (warp < QG && q0 + warp < a.n_queries) {
    const int64_t o = ((q0 + warp) * a.n_lists + slice) * a.ksel;
    for (int j = lane; j < a.k; j += 32) {
      a.part_scores[o + j] = ls[warp * a.k + j];
      a.part_idx[o + j] = li[warp * a.k + j];
    }
  }
}

int slices_for(const SearchArgs& a, int sm_count, int qg, int* blocks_per_slice) {
  const int64_t nblk = (a.n_rows + kBlockRows - 1) / kBlockRows;
  const int64_t groups = (a.n_queries + qg - 1) / qg;
  int64_t want = ((int64_t)sm_count * 8 + groups - 1) / groups;  // ~8 CTAs per SM overall
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  if (want > nblk) want = nblk;
  if (want < 1) want = 1;
  const int64_t bps = (nblk + want - 1) / want;
  *blocks_per_slice = (int)(bps < 1 ? 1 : bps);
  return (int)((nblk + *blocks_per_slice - 1) / (*blocks_per_slice) < 1
                   ? 1
                   : (nblk + *blocks_per_slice - 1) / (*blocks_per_slice));
}

inline int simt_qg(const SearchArgs& a) { return a.n_queries >= 2 ? kSimtQG : 1; }

}  // namespace

int simt_plan(const SearchArgs& a, int sm_count, int* n_lists, int* ksel) {
  int bps;
  *n_lists = slices_for(a, sm_count, simt_qg(a), &bps);
  *ksel = a.k;
  return LK_OK;
}

int launch_search_simt(const SearchArgs& a, int sm_count, cudaStream_t st) {
  if (a.n_queries <= 0 || a.n_rows <= 0) return LK_OK;
  const int qg = simt_qg(a);
  int bps;
  const int slices = slices_for(a, sm_count, qg, &bps);
  if (slices != a.n_lists) {
    set_error("simt search: plan mismatch (%d lists planned, %d slices)", a.n_lists, slices);
    return LK_ERR_INVALID;
  }
  const int64_t groups = (a.n_queries + qg - 1) / qg;
  if (groups > 65535) {
    set_error("simt search: batch of %lld queries is too large for one launch", (long long)a.n_queries);
    return LK_ERR_UNSUPPORTED;
  }
  const size_t smem = (size_t)qg * (a.g.dim_pad + kBlockRows + 2 * a.k) * 4;
  const dim3 grid((unsigned)slices, (unsigned)groups);
#define LK_SIMT(T, QG)                                                                              \
  do {                                                                                              \
    LK_CUDA(cudaFuncSetAttribute(simt_search_kernel<T, QG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)smem));                                                       \
    simt_search_kernel<T, QG><<<grid, kBlockRows, smem, st>>>(a, bps);                              \
  } while (0)
  if (a.g.elem_bytes == 2) {
    if (qg == 1) LK_SIMT(__nv_bfloat16, 1); else LK_SIMT(__nv_bfloat16, kSimtQG);
  } else {
    if (qg == 1) LK_SIMT(float, 1); else LK_SIMT(float, kSimtQG);
  }
#undef LK_SIMT
  LK_CHECK_LAUNCH("simt_search_kernel");
  return LK_OK;
}

}  // namespace lk
