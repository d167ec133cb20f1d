// Linear layer on tcgen05:  Y = act(X W^T + b) (+ R), fp32 in / fp32 out.
//
// The GEMMs of the sentence encoder (all-MiniLM-L6-v2 under retrieval/embedder.py:35-40, a BERT
// whose linear layers are torch.nn.Linear in fp32).  Same operand scheme as the autoencoder
// kernel (lk_ae_umma.cu): every operand is carried as two bf16 planes (x = hi + lo) and every
// product as three MMAs (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM), which keeps the
// result at fp32-level accuracy (~1e-5 of the row scale); one plane = plain bf16 operands.
//
// X arrives pre-split into planes in the SWIZZLE_128B slab format (split_rows_kernel,
// [m tile][plane][K block] x 16 KB), the weights likewise ([n tile][plane][K block], built once at
// load time), so a K step of one 128 x 128 output tile is four 16 KB cp.async.bulk copies into a
// 64 KB stage and 12 MMAs of 128 x 128 x 16.  Warp 0 produces, warp 1 issues, warps 4-11 read the
// accumulator (two of them in TMEM: the epilogue of tile t overlaps the MMAs of tile t+1), add the
// bias, apply GELU / the residual and store fp32 rows -- or, when the only reader is the next linear
// layer, that layer's operand planes directly.  Persistent CTAs, tiles n-fastest so the X
// planes of an m tile are re-read from L2.
#include "lk_common.cuh"
#include "lk_planes.cuh"
#include "lk_ptx.cuh"

namespace lk {

namespace {

constexpr int kThreads = 384;
constexpr int kFirstEpiWarp = 4;
constexpr int kEpiWarps = 8;
constexpr int kPlanes = 2;
constexpr int kStages = 3;
constexpr int kStageBytes = 2 * kPlanes * kSlabBytes;  // X planes + W planes of one K block: 64 KB
constexpr int kHeaderBytes = 256;
constexpr int kTmemCols = 256;  // two 128-column accumulators

enum GemmErr { kGemmProd = 301, kGemmMmaAcc = 302, kGemmMmaFull = 303, kGemmEpi = 304 };

struct GemmParams {
  const unsigned char* x_slabs;
  const unsigned char* w_slabs;
  const float* bias;      // [n] or null
  const float* residual;  // [m, n] or null
  float* y;               // [m, n] fp32, or null:
  unsigned char* y_planes;  // the result as operand planes (the next linear layer's X)
  int64_t m;
  int n, m_tiles, n_tiles, nkb;
  int act;                // 0 none, 1 GELU (erf)
  int np;                 // operand planes in use
  int* err_flag;
};

struct Ring {
  int idx = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++idx == n) {
      idx = 0;
      phase ^= 1u;
    }
  }
};

__global__ void __launch_bounds__(kThreads, 1) gemm_umma_kernel(const GemmParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t bar0 = ptx::smem_u32(smem);
  // barriers: full[3] empty[3] accfull[2] accempty[2]
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto accfull_bar = [&](int a) { return bar0 + 8u * (2 * kStages + a); };
  auto accempty_bar = [&](int a) { return bar0 + 8u * (2 * kStages + 2 + a); };
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + 200);
  unsigned char* data = smem + kHeaderBytes;
  data += (1024u - (ptx::smem_u32(data) & 1023u)) & 1023u;
  unsigned char* stage_sm = data;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(accfull_bar(a), 1);
      ptx::mbar_init(accempty_bar(a), kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr_s), kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  auto fail = [&](int code) {
    if (lane == 0) atomicCAS(p.err_flag, 0, code);
  };
  auto wait = [&](uint32_t bar, uint32_t parity) { return __all_sync(0xffffffffu, ptx::mbar_wait(bar, parity)); };

  const int64_t tile_stride = (int64_t)kPlanes * p.nkb * kSlabBytes;  // one m tile of X / one n tile of W
  const int64_t total = (int64_t)p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    Ring st;
    bool ok = true;
    for (int64_t w = blockIdx.x; w < total && ok; w += gridDim.x) {
      const unsigned char* xt = p.x_slabs + (w / p.n_tiles) * tile_stride;
      const unsigned char* wt = p.w_slabs + (w % p.n_tiles) * tile_stride;
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (!wait(empty_bar(st.idx), st.phase ^ 1u)) { fail(kGemmProd); ok = false; break; }
        if (ptx::elect_one()) {
          const uint32_t dst = ptx::smem_u32(stage_sm + st.idx * kStageBytes);
          ptx::mbar_arrive_expect_tx(full_bar(st.idx), (uint32_t)(2 * p.np * kSlabBytes));
#pragma unroll
          for (int pl = 0; pl < kPlanes; ++pl) {
            if (pl >= p.np) break;
            ptx::bulk_g2s(dst + pl * kSlabBytes, xt + ((int64_t)pl * p.nkb + kb) * kSlabBytes, kSlabBytes,
                          full_bar(st.idx));
            ptx::bulk_g2s(dst + (kPlanes + pl) * kSlabBytes, wt + ((int64_t)pl * p.nkb + kb) * kSlabBytes, kSlabBytes,
                          full_bar(st.idx));
          }
        }
        __syncwarp();
        st.advance(kStages);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ptx::idesc_bf16_f32(kBlockRows, kBlockRows);
    const uint64_t desc_hi = ptx::smem_desc(0, 16, 1024);
    const uint32_t st_lo = ptx::smem_u32(stage_sm) >> 4;
    auto desc = [&](uint32_t lo) { return desc_hi | (uint64_t)(lo & 0x3fffu); };
    Ring st;
    uint32_t work_no = 0;
    bool ok = true;
    for (int64_t w = blockIdx.x; w < total && ok; w += gridDim.x, ++work_no) {
      const uint32_t a = work_no & 1u;
      if (!wait(accempty_bar(a), ((work_no >> 1) & 1u) ^ 1u)) { fail(kGemmMmaAcc); break; }
      ptx::tc_fence_after();
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (!wait(full_bar(st.idx), st.phase)) { fail(kGemmMmaFull); ok = false; break; }
        ptx::tc_fence_after();
        const uint32_t s_lo = st_lo + (uint32_t)(st.idx * (kStageBytes >> 4));
        if (ptx::elect_one()) {
          const uint32_t slab16 = kSlabBytes >> 4;
#pragma unroll
          for (int t = 0; t < 3; ++t) {  // hi*hi, hi*lo, lo*hi (plain bf16: hi*hi only)
            if (t > 0 && p.np == 1) break;
            const uint32_t xa = s_lo + (t == 2 ? slab16 : 0u);
            const uint32_t wb = s_lo + 2u * slab16 + (t == 1 ? slab16 : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16(tmem_base + a * kBlockRows, desc(xa + 2u * k), desc(wb + 2u * k), idesc,
                             (kb | t | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(st.idx));
        }
        __syncwarp();
        st.advance(kStages);
      }
      if (!ok) break;
      if (ptx::elect_one()) ptx::umma_commit(accfull_bar(a));
      __syncwarp();
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================== epilogue =====================
    const int ew = warp - kFirstEpiWarp;
    const int quarter = warp & 3;
    const int ch = ew >> 2;               // which 64 of the tile's 128 columns
    const int row = quarter * 32 + lane;  // row of the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t work_no = 0;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x, ++work_no) {
      const uint32_t a = work_no & 1u;
      if (!wait(accfull_bar(a), (work_no >> 1) & 1u)) { fail(kGemmEpi); break; }
      ptx::tc_fence_after();
      const int64_t grow = (w / p.n_tiles) * kBlockRows + row;
      const int col0 = (int)(w % p.n_tiles) * kBlockRows + ch * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem_base + lane_addr + a * kBlockRows + ch * 64 + half * 32, r);
        const int col = col0 + half * 32;
        float4 bv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          bv[j] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        ptx::tmem_wait_ld();
        if (grow < p.m) {
          const float* bf = reinterpret_cast<const float*>(bv);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = __uint_as_float(r[j]) + bf[j];
            if (p.act == 1) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752f));
          }
          if (p.residual) {  // 256-bit accesses: a thread's piece of its row is a whole 32-byte sector
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float t[8];
              ptx::ldg256(p.residual + grow * p.n + col + 8 * j, t);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[8 * j + e] += t[e];
            }
          }
          if (p.y) {
            float* out = p.y + grow * p.n + col;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float t[8] = {v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3],
                                  v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]};
              ptx::stg256(out + 8 * j, t);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float y16[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) y16[e] = v[16 * c + e];
              store_planes16(p.y_planes, p.n >> 6, grow, col + 16 * c, y16, p.np);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(accempty_bar(a));
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

int gemm_umma_supported(int n, int k) { return n >= kBlockRows && n % kBlockRows == 0 && k >= 64 && k % 64 == 0; }

int launch_gemm_umma(const unsigned char* x_slabs, int64_t m, int k, const unsigned char* w_slabs, int n,
                     const float* bias, const float* residual, int act, int n_planes, float* y,
                     unsigned char* y_planes, int* err_flag, int sm_count, cudaStream_t st) {
  if (m <= 0) return LK_OK;
  if (!gemm_umma_supported(n, k)) {
    set_error("tcgen05 linear layer: unsupported shape (n=%d, k=%d)", n, k);
    return LK_ERR_UNSUPPORTED;
  }
  GemmParams p;
  p.x_slabs = x_slabs;
  p.w_slabs = w_slabs;
  p.bias = bias;
  p.residual = residual;
  p.y = y;
  p.y_planes = y_planes;
  p.m = m;
  p.n = n;
  p.m_tiles = (int)((m + kBlockRows - 1) / kBlockRows);
  p.n_tiles = n / kBlockRows;
  p.nkb = k / 64;
  p.act = act;
  p.np = n_planes == 1 ? 1 : 2;
  p.err_flag = err_flag;
  const size_t smem = kHeaderBytes + 1024 + (size_t)kStages * kStageBytes;
  LK_CUDA(cudaFuncSetAttribute(gemm_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t total = (int64_t)p.m_tiles * p.n_tiles;
  const int grid = (int)(total < sm_count ? total : sm_count);
  gemm_umma_kernel<<<grid, kThreads, smem, st>>>(p);
  LK_CHECK_LAUNCH("gemm_umma_kernel");
  return LK_OK;
}

}  // namespace lk
