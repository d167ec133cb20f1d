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
template <typename ElemT, int QG, bool FUSED>
__global__ void __launch_bounds__(kBlockRows) simt_search_kernel(SearchArgs a, int blocks_per_slice) {
  constexpr int E = Chunk<ElemT>::E;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int dim_pad = a.g.dim_pad;
  float* qs = reinterpret_cast<float*>(smem_raw);                 // [QG][dim_pad]
  float* sc = qs + QG * dim_pad;                                  // [QG][128]
  constexpr int kLists = FUSED && QG == 1 ? kBlockRows / 32 : QG;  // single-launch merge: one list per warp
  float* ls = sc + QG * kBlockRows;                               // [kLists][k]
  int32_t* li = reinterpret_cast<int32_t*>(ls + kLists * a.k);    // [kLists][k]

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int slice = blockIdx.x;
  const int64_t q0 = (int64_t)blockIdx.y * QG;
  const int64_t nblk = (a.n_rows + kBlockRows - 1) / kBlockRows;
  const int64_t blk_lo = (int64_t)slice * blocks_per_slice;
  const int64_t blk_hi = min(nblk, blk_lo + blocks_per_slice);
  const int64_t block_bytes = a.g.block_bytes();

  float qside[QG];
  if (FUSED) {
    // raw queries: round to the storage type like the tiling kernel does, and compute the side
    // value with ITS order of operations (one warp per query: lane-strided fmaf chain, xor-shuffle
    // sum), so the scores equal those of the tiled path bit for bit
    for (int i = t; i < QG * dim_pad; i += kBlockRows) {
      const int g = i / dim_pad, c = i - g * dim_pad;
      const int64_t q = q0 + g;
      float v = 0.f;
      if (q < a.n_queries && c < a.g.dim) {
        v = a.q_raw_dtype == LK_F32 ? static_cast<const float*>(a.q_raw)[q * a.g.dim + c]
                                    : __bfloat162float(static_cast<const __nv_bfloat16*>(a.q_raw)[q * a.g.dim + c]);
        if (sizeof(ElemT) == 2) v = __bfloat162float(__float2bfloat16_rn(v));
      }
      qs[i] = v;
    }
    __syncthreads();
    if (warp < QG) {
      float ss = 0.f;
      for (int c = lane; c < a.g.dim; c += 32) ss = fmaf(qs[warp * dim_pad + c], qs[warp * dim_pad + c], ss);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) sc[warp] = a.metric == LK_COSINE ? 1.0f / fmaxf(sqrtf(ss), 1e-12f) : ss;
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < QG; ++g) qside[g] = sc[g];
    __syncthreads();
  }
  // stage the group's queries as fp32 (gather out of the query tiles)
  const unsigned char* qt = static_cast<const unsigned char*>(a.q_tiles);
  for (int i = t; !FUSED && i < QG * dim_pad; i += kBlockRows) {
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
  if (!FUSED) {
#pragma unroll
    for (int g = 0; g < QG; ++g) qside[g] = (q0 + g < a.n_queries) ? a.q_side[q0 + g] : 0.f;
  }
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
    if (FUSED && blocks_per_slice == 1) {
      // one row block per slice (small corpora): no running list to maintain -- every thread ranks
      // its own score among the 128 (all-pairs, shared-memory broadcasts) and the best k land in
      // the list sorted; ~1 us instead of tens of serial list insertions
#pragma unroll
      for (int g = 0; g < QG; ++g) {
        const float v = sc[g * kBlockRows + t];
        int rank = kBlockRows;  // NaN (a row past the end): never selected
        if (v == v) {
          rank = 0;
          for (int u = 0; u < kBlockRows; ++u) {
            const float su = sc[g * kBlockRows + u];
            rank += (su > v || (su == v && u < t)) ? 1 : 0;
          }
        }
        if (rank < a.k) {
          ls[g * a.k + rank] = v;
          li[g * a.k + rank] = (int32_t)(blk * kBlockRows + t);
        }
      }
    } else if (warp < QG && q0 + warp < a.n_queries) {
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
  if (!FUSED) return;
  // The slice that finishes last merges all slices' lists of its query group and writes the final
  // result (threadFenceReduction pattern: lists, fence, ticket).
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (t == 0) {
    const int n_slices = (int)gridDim.x;
    const int done = atomicAdd(a.ticket + blockIdx.y, 1);
    s_last = done == n_slices - 1;
    if (s_last) a.ticket[blockIdx.y] = 0;  // ready for the next call
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // Final merge by the last slice.  Every list is sorted best first, so T = the largest k-th entry
  // over the slices bounds the global k-th best from below (that slice alone holds k rows >= T):
  // only candidates >= T can be in the result.  They are gathered (a few dozen, typically) and
  // ranked all-pairs under (score desc, index asc).  If more than kGather qualify (heavy ties),
  // the warps fall back to folding all candidates through sorted lists.
  constexpr int kGather = 512;
  float* gs = reinterpret_cast<float*>(li + kLists * a.k);   // [kGather] gathered scores
  int32_t* gi = reinterpret_cast<int32_t*>(gs + kGather);    // [kGather] gathered ids
  __shared__ float s_thr;
  __shared__ int s_thr_key;  // order-preserving key of the best bound, >> 1 so that it fits a signed atomicMax
  __shared__ int s_n;
  constexpr int kMaxSlices = 512;
  float* gi_end = reinterpret_cast<float*>(gi + kGather);
  const int n_slices = (int)gridDim.x;
  if (t == 0) s_thr_key = 0;
  __syncthreads();
  for (int g = 0; g < QG; ++g) {
    const int64_t q = q0 + g;
    if (q >= a.n_queries) break;  // block-uniform
    const float* ps = a.part_scores + q * a.n_lists * a.ksel;
    const int32_t* pi = a.part_idx + q * a.n_lists * a.ksel;
    // Lower bounds of the global k-th best: for a depth j, the ceil(k / j)-th largest of the
    // slices' j-th entries (that many slices hold j rows each at or above it).  Depths 1, 2, 4, 8
    // and k are tried -- one batch of loads -- and the largest bound wins; with 128-row slices the
    // k-th entries alone (depth k) would let through hundreds of candidates.
    constexpr int kDepths = 5;
    float* col = gi_end;  // [kDepths][kMaxSlices] staged entries
    const int depth[kDepths] = {1, 2, 4, 8, a.k};
    if (t == 0) {
      s_n = 0;
      s_thr = -INFINITY;
    }
    for (int sl = t; sl < n_slices; sl += kBlockRows)
#pragma unroll
      for (int d = 0; d < kDepths; ++d)
        col[d * kMaxSlices + sl] = depth[d] <= a.k ? __ldcg(ps + (int64_t)sl * a.ksel + (depth[d] - 1)) : -INFINITY;
    __syncthreads();
#pragma unroll
    for (int d = 0; d < kDepths; ++d) {
      const int need = (a.k + depth[d] - 1) / depth[d];  // slices that must reach the bound
      if (depth[d] > a.k || need > n_slices) continue;
      for (int sl = t; sl < n_slices; sl += kBlockRows) {
        const float v = col[d * kMaxSlices + sl];
        int rank = 0;
        for (int u = 0; u < n_slices; ++u) {
          const float w = col[d * kMaxSlices + u];
          rank += (w > v || (w == v && u < sl)) ? 1 : 0;
        }
        if (rank == need - 1) atomicMax(reinterpret_cast<int*>(&s_thr_key), (int)(order_key(v) >> 1));
      }
    }
    __syncthreads();
    float thr = s_thr_key > 0 ? order_key_inv(((uint32_t)s_thr_key << 1)) : -INFINITY;  // rounded DOWN: still a lower bound
    __syncthreads();
    if (t == 0) s_thr_key = 0;
    const int n_cand = n_slices * a.k;
    for (int c0 = t; c0 < n_cand; c0 += kBlockRows * 8) {  // eight independent loads in flight per thread
      float v[8];
      int64_t off[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * kBlockRows;
        v[u] = -INFINITY;
        off[u] = 0;
        if (c < n_cand) {
          const int sl = c / a.k, e = c - sl * a.k;
          off[u] = (int64_t)sl * a.ksel + e;
          v[u] = __ldcg(ps + off[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (c0 + u * kBlockRows < n_cand && v[u] >= thr) {
          const int32_t id = __ldcg(pi + off[u]);
          if (id >= 0 && id != IdxTraits<int32_t>::sentinel()) {
            const int pos = atomicAdd(&s_n, 1);
            if (pos < kGather) {
              gs[pos] = v[u];
              gi[pos] = id;
            }
          }
        }
      }
    }
    __syncthreads();
    const int n = s_n;
    if (n <= kGather) {
      for (int j = t; j < a.k; j += kBlockRows) {  // fewer than k candidates exist: padding
        a.out_scores[q * a.k + j] = -INFINITY;
        a.out_idx[q * a.k + j] = -1;
      }
      __syncthreads();
      for (int i = t; i < n; i += kBlockRows) {
        const float v = gs[i];
        const int32_t id = gi[i];
        int rank = 0;
        for (int u = 0; u < n; ++u) rank += (gs[u] > v || (gs[u] == v && gi[u] < id)) ? 1 : 0;
        if (rank < a.k) {
          a.out_scores[q * a.k + rank] = v;
          a.out_idx[q * a.k + rank] = (int64_t)id + a.idx_base;
        }
      }
    } else if (warp == 0) {  // heavy ties at the threshold: one warp, sorted-list insertion
      float* s = ls;
      int32_t* ix = li;
      warp_list_init<int32_t>(s, ix, a.k, lane);
      for (int base = 0; base < n_cand; base += 32) {
        const int c = base + lane;
        float v = 0.f;
        int32_t id = -1;
        if (c < n_cand) {
          const int sl = c / a.k, e = c - sl * a.k;
          v = __ldcg(ps + (int64_t)sl * a.ksel + e);
          id = __ldcg(pi + (int64_t)sl * a.ksel + e);
        }
        warp_list_offer<int32_t>(s, ix, a.k, v, id, c < n_cand && id >= 0 && id != IdxTraits<int32_t>::sentinel(), lane);
      }
      for (int j = lane; j < a.k; j += 32) {
        const bool filled = ix[j] != IdxTraits<int32_t>::sentinel();
        a.out_scores[q * a.k + j] = s[j];
        a.out_idx[q * a.k + j] = filled ? (int64_t)ix[j] + a.idx_base : (int64_t)-1;
      }
    }
    __syncthreads();
  }
}

int slices_for(const SearchArgs& a, int sm_count, int qg, int* blocks_per_slice) {
  const int64_t nblk = (a.n_rows + kBlockRows - 1) / kBlockRows;
  const int64_t groups = (a.n_queries + qg - 1) / qg;
  int64_t want = ((int64_t)sm_count * 8 + groups - 1) / groups;  // ~8 CTAs per SM overall
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  if (a.q_raw != nullptr && want > 512) want = 512;  // the single-launch merge stages one value per slice
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

// Measured against the general path (query tiling + tcgen05 search + merge, ~90 us per call): the
// single launch wins while rows x queries <= ~40k (B=1: 59 vs 87 us at 20k rows, 83 vs 99 at 35k;
// B=4: 69 vs 89 us at 8k rows, 118 vs 98 at 20k), whatever the dimension.
int simt_fused_supported(const TileGeom& g, int64_t n_rows, int64_t n_queries, int k) {
  return g.elem_bytes == 2 && n_queries >= 1 && n_queries <= kSimtQG && k >= 1 && k <= kMaxK &&
         n_rows * n_queries <= 40000 && n_rows * (int64_t)g.dim_pad * 2 <= (64ll << 20);
}

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
  const int n_lists_smem = a.q_raw != nullptr && qg == 1 ? kBlockRows / 32 : qg;
  const size_t smem = ((size_t)qg * (a.g.dim_pad + kBlockRows) + (size_t)n_lists_smem * 2 * a.k) * 4 +
                      (a.q_raw != nullptr ? 512 * 8 + 5 * 512 * 4 : 0);  // single-launch merge: gathered candidates, bounds
  const dim3 grid((unsigned)slices, (unsigned)groups);
#define LK_SIMT(T, QG, FU)                                                                          \
  do {                                                                                              \
    if (smem > 48 * 1024)                                                                           \
      LK_CUDA(cudaFuncSetAttribute(simt_search_kernel<T, QG, FU>,                                   \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    simt_search_kernel<T, QG, FU><<<grid, kBlockRows, smem, st>>>(a, bps);                          \
  } while (0)
  if (a.q_raw != nullptr) {  // single-launch mode
    if (a.g.elem_bytes != 2 || !a.ticket || !a.out_scores || !a.out_idx || groups != 1) {
      set_error("simt search: the single-launch mode needs bf16 storage and one query group");
      return LK_ERR_INVALID;
    }
    if (qg == 1) LK_SIMT(__nv_bfloat16, 1, true); else LK_SIMT(__nv_bfloat16, kSimtQG, true);
  } else if (a.g.elem_bytes == 2) {
    if (qg == 1) LK_SIMT(__nv_bfloat16, 1, false); else LK_SIMT(__nv_bfloat16, kSimtQG, false);
  } else {
    if (qg == 1) LK_SIMT(float, 1, false); else LK_SIMT(float, kSimtQG, false);
  }
#undef LK_SIMT
  LK_CHECK_LAUNCH("simt_search_kernel");
  return LK_OK;
}

}  // namespace lk
