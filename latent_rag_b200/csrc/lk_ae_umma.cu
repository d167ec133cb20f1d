// Fused autoencoder-encoder forward on tcgen05:  Z = relu(X W0^T + b0) W1^T + b1 (+ L2 norm).
//
// Replaces the encoder halves of the reference autoencoders (models/denoising_autoencoder.py:
// 19-23,33-34; models/contrastive_autoencoder.py:10-14,23-25; models/variational_autoencoder.py:
// 11-16,27-28, mu only as retrieval/embedder.py:44-45 keeps).  The reference computes in fp32;
// to stay at fp32-level accuracy on bf16 tensor cores every operand is carried as TWO bf16
// planes (x = hi + lo, 16 mantissa bits) and every product as three MMAs
// (hi*hi + hi*lo + lo*hi, fp32 accumulate): relative error ~1e-5 of the row scale.
//
// Per 128-row tile, per hidden chunk c of 128 units:
//   layer 0   acc0[c&1] (TMEM, 128 cols) = X(128 x d_in) . W0[c](128 x d_in)^T      warp 1
//   epilogue  h = relu(acc0 + b0) -> hi/lo planes -> swizzled smem (A operand of layer 1)
//   layer 1   acc1 (TMEM, d_latent cols) += H[c](128 x 128) . W1[:, c](d_latent x 128)^T
// so the hidden activations never leave the SM.  Layer 0 of chunk c+1 overlaps the epilogue of
// chunk c (two accumulators).  X (pre-split into planes by split_rows_kernel) and the weight
// slabs are all in the SWIZZLE_128B slab format of lk_common.cuh and arrive by cp.async.bulk.
#include <cstdlib>
#include <vector>

#include <cstring>

#include "lk_common.cuh"
#include "lk_ptx.cuh"

namespace lk {

namespace {

constexpr int kThreads = 384;
constexpr int kFirstEpiWarp = 4;
constexpr int kEpiWarps = 8;
constexpr int kPlanes = 2;
constexpr int kStages = 2;                                  // layer-0 smem stages
constexpr int kStageBytes = 2 * kPlanes * kSlabBytes;       // X planes + W0 planes of one K block: 64 KB
constexpr int kHBytes = kPlanes * 2 * kSlabBytes;           // hidden chunk: 2 planes x 2 K blocks: 64 KB
constexpr int kHeaderBytes = 256;
constexpr int kTmemCols = 512;
constexpr int kAcc1Col = 256;

enum AeErr { kAeProd = 201, kAeProdW1 = 202, kAeMmaFull = 203, kAeMmaAcc0 = 204, kAeMmaH = 205, kAeMmaW1 = 206,
             kAeMmaZ = 207, kAeEpiAcc0 = 208, kAeEpiH = 209, kAeEpiZ = 210 };

struct AeParams {
  const unsigned char* x_slabs;   // [tile][plane][kb0] slabs of 128 rows
  const unsigned char* w0_slabs;  // [chunk][plane][kb0]
  const unsigned char* w1_slabs;  // [plane][kb1] slabs of n1 rows (n1 * 128 bytes each)
  const float* b0;
  const float* b1;
  float* z;                       // [m, n1_true]
  int64_t m;
  int n_tiles, nkb0, n_chunks, n1, n1_true, l2norm;
  int np;                         // operand planes in use: 2 = split-bf16 (hi + lo, three MMAs per product), 1 = plain bf16
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

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kThreads, 1) ae_umma_kernel(const AeParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t bar0 = ptx::smem_u32(smem);
  // barriers: full[2] empty[2] w1full w1empty acc0full[2] acc0empty[2] hfull hempty zfull zempty
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 + s); };
  const uint32_t w1full_bar = bar0 + 8u * 4, w1empty_bar = bar0 + 8u * 5;
  auto a0full_bar = [&](int s) { return bar0 + 8u * (6 + s); };
  auto a0empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  const uint32_t hfull_bar = bar0 + 8u * 10, hempty_bar = bar0 + 8u * 11;
  const uint32_t zfull_bar = bar0 + 8u * 12, zempty_bar = bar0 + 8u * 13;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + 200);
  unsigned char* data = smem + kHeaderBytes;
  data += (1024u - (ptx::smem_u32(data) & 1023u)) & 1023u;
  unsigned char* stage_sm = data;                                  // kStages x 64 KB
  unsigned char* h_sm = stage_sm + kStages * kStageBytes;          // 64 KB
  unsigned char* w1_sm = h_sm + kHBytes;                           // planes x 2 K blocks x n1 rows
  const int w1_slab = p.n1 * kRowBytes;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
      ptx::mbar_init(a0full_bar(s), 1);
      ptx::mbar_init(a0empty_bar(s), kEpiWarps);
    }
    ptx::mbar_init(w1full_bar, 1);
    ptx::mbar_init(w1empty_bar, 1);
    ptx::mbar_init(hfull_bar, kEpiWarps);
    ptx::mbar_init(hempty_bar, 1);
    ptx::mbar_init(zfull_bar, 1);
    ptx::mbar_init(zempty_bar, kEpiWarps / 2);
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

  const int64_t slab_tile_stride = (int64_t)kPlanes * p.nkb0 * kSlabBytes;

  if (warp == 0) {
    // ===================== TMA producer =====================
    Ring st;
    uint32_t w1_uses = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x) {
      const unsigned char* xt = p.x_slabs + (int64_t)tile * slab_tile_stride;
      for (int c = 0; c < p.n_chunks && ok; ++c) {
        const unsigned char* wc = p.w0_slabs + (int64_t)c * slab_tile_stride;
        for (int kb = 0; kb < p.nkb0; ++kb) {
          if (!wait(empty_bar(st.idx), st.phase ^ 1u)) { fail(kAeProd); ok = false; break; }
          if (ptx::elect_one()) {
            const uint32_t dst = ptx::smem_u32(stage_sm + st.idx * kStageBytes);
            ptx::mbar_arrive_expect_tx(full_bar(st.idx), (uint32_t)(2 * p.np * kSlabBytes));
#pragma unroll
            for (int pl = 0; pl < kPlanes; ++pl) {
              if (pl >= p.np) break;
              ptx::bulk_g2s(dst + pl * kSlabBytes, xt + ((int64_t)pl * p.nkb0 + kb) * kSlabBytes, kSlabBytes,
                            full_bar(st.idx));
              ptx::bulk_g2s(dst + (kPlanes + pl) * kSlabBytes, wc + ((int64_t)pl * p.nkb0 + kb) * kSlabBytes,
                            kSlabBytes, full_bar(st.idx));
            }
          }
          __syncwarp();
          st.advance(kStages);
        }
        if (!ok) break;
        // W1 slabs of this hidden chunk (K blocks 2c, 2c+1 of layer 1), both planes
        if (!wait(w1empty_bar, (w1_uses & 1u) ^ 1u)) { fail(kAeProdW1); ok = false; break; }
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(w1full_bar, (uint32_t)(p.np * 2 * w1_slab));
          const int nkb1 = 2 * p.n_chunks;
#pragma unroll
          for (int pl = 0; pl < kPlanes; ++pl)
#pragma unroll
            for (int j = 0; j < 2; ++j)
              if (pl < p.np)
              ptx::bulk_g2s(ptx::smem_u32(w1_sm + (pl * 2 + j) * w1_slab),
                            p.w1_slabs + ((int64_t)pl * nkb1 + 2 * c + j) * w1_slab, (uint32_t)w1_slab, w1full_bar);
        }
        __syncwarp();
        ++w1_uses;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc0 = ptx::idesc_bf16_f32(kBlockRows, kBlockRows);
    const uint32_t idesc1 = ptx::idesc_bf16_f32(kBlockRows, p.n1);
    const uint64_t desc_hi = ptx::smem_desc(0, 16, 1024);
    const uint32_t st_lo = ptx::smem_u32(stage_sm) >> 4, h_lo = ptx::smem_u32(h_sm) >> 4,
                   w1_lo = ptx::smem_u32(w1_sm) >> 4;
    auto desc = [&](uint32_t lo) { return desc_hi | (uint64_t)(lo & 0x3fffu); };
    Ring st;
    uint32_t chunk_no = 0;   // global count of hidden chunks issued to layer 0
    uint32_t l1_no = 0;      // global count of hidden chunks issued to layer 1
    uint32_t tile_no = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x, ++tile_no) {
      for (int step = 0; step <= p.n_chunks && ok; ++step) {
        if (step < p.n_chunks) {  // ---- layer 0 of chunk `step`
          const uint32_t a = chunk_no & 1u;
          if (!wait(a0empty_bar(a), ((chunk_no >> 1) & 1u) ^ 1u)) { fail(kAeMmaAcc0); ok = false; break; }
          ptx::tc_fence_after();
          for (int kb = 0; kb < p.nkb0; ++kb) {
            if (!wait(full_bar(st.idx), st.phase)) { fail(kAeMmaFull); ok = false; break; }
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
                  ptx::umma_bf16(tmem_base + a * kBlockRows, desc(xa + 2u * k), desc(wb + 2u * k), idesc0,
                                 (kb | t | k) != 0 ? 1u : 0u);
              }
              ptx::umma_commit(empty_bar(st.idx));
            }
            __syncwarp();
            st.advance(kStages);
          }
          if (!ok) break;
          if (ptx::elect_one()) ptx::umma_commit(a0full_bar(a));
          __syncwarp();
          ++chunk_no;
        }
        if (step >= 1) {  // ---- layer 1 of chunk `step - 1`
          const int c = step - 1;
          if (!wait(hfull_bar, l1_no & 1u)) { fail(kAeMmaH); ok = false; break; }
          if (!wait(w1full_bar, l1_no & 1u)) { fail(kAeMmaW1); ok = false; break; }
          if (c == 0 && !wait(zempty_bar, (tile_no & 1u) ^ 1u)) { fail(kAeMmaZ); ok = false; break; }
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t hslab16 = kSlabBytes >> 4, wslab16 = (uint32_t)(p.n1 * kRowBytes) >> 4;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              if (t > 0 && p.np == 1) break;
              const uint32_t hp = t == 2 ? 1u : 0u, wp = t == 1 ? 1u : 0u;
#pragma unroll
              for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  ptx::umma_bf16(tmem_base + kAcc1Col, desc(h_lo + (hp * 2u + j) * hslab16 + 2u * k),
                                 desc(w1_lo + (wp * 2u + j) * wslab16 + 2u * k), idesc1,
                                 (c | t | j | k) != 0 ? 1u : 0u);
            }
            ptx::umma_commit(hempty_bar);
            ptx::umma_commit(w1empty_bar);
            if (c == p.n_chunks - 1) ptx::umma_commit(zfull_bar);
          }
          __syncwarp();
          ++l1_no;
        }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================== epilogue =====================
    const int ew = warp - kFirstEpiWarp;
    const int quarter = warp & 3;
    const int ch = ew >> 2;                  // which 64 of the chunk's 128 hidden units (= K block of H)
    const int row = quarter * 32 + lane;     // row of the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t chunk_no = 0, tile_no = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x, ++tile_no) {
      for (int c = 0; c < p.n_chunks; ++c, ++chunk_no) {
        const uint32_t a = chunk_no & 1u;
        if (!wait(a0full_bar(a), (chunk_no >> 1) & 1u)) { fail(kAeEpiAcc0); ok = false; break; }
        // the H buffer is free once layer 1 of the previous chunk has been read by the tensor core
        if (!wait(hempty_bar, (chunk_no & 1u) ^ 1u)) { fail(kAeEpiH); ok = false; break; }
        ptx::tc_fence_after();
        const float* bias = p.b0 + c * kBlockRows + ch * 64;
        unsigned char* hrow = h_sm + ch * kSlabBytes + row * kRowBytes;  // plane 0; plane 1 is + 2 slabs
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          ptx::tmem_ld32(tmem_base + lane_addr + a * kBlockRows + ch * 64 + half * 32, r);
          float4 bv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(bias + half * 32) + j);
          ptx::tmem_wait_ld();
          const float* bf = reinterpret_cast<const float*>(bv);
#pragma unroll
          for (int cj = 0; cj < 4; ++cj) {  // 4 chunks of 8 hidden units per 32 columns
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j0 = cj * 8 + 2 * e;
              const float h0 = fmaxf(__uint_as_float(r[j0]) + bf[j0], 0.f);
              const float h1 = fmaxf(__uint_as_float(r[j0 + 1]) + bf[j0 + 1], 0.f);
              const __nv_bfloat16 a0 = __float2bfloat16_rn(h0), a1 = __float2bfloat16_rn(h1);
              hi[e] = pack_bf16(h0, h1);
              lo[e] = pack_bf16(h0 - __bfloat162float(a0), h1 - __bfloat162float(a1));
            }
            const int chunk8 = half * 4 + cj;  // logical 16-byte chunk within the 128-byte row
            const int off = (chunk8 ^ (row & 7)) << 4;
            *reinterpret_cast<uint4*>(hrow + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (p.np == 2)
              *reinterpret_cast<uint4*>(hrow + 2 * kSlabBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        ptx::tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic stores -> tensor-core reads
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(a0empty_bar(a));
          ptx::mbar_arrive(hfull_bar);
        }
      }
      if (!ok) break;
      if (ch == 0) {  // ---- final: Z row = acc1 + b1 (+ L2 normalisation), one thread per row
        if (!wait(zfull_bar, tile_no & 1u)) { fail(kAeEpiZ); break; }
        ptx::tc_fence_after();
        float zr[64];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half * 32 < p.n1) {
            uint32_t r[32];
            ptx::tmem_ld32(tmem_base + lane_addr + kAcc1Col + half * 32, r);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = half * 32 + j;
              zr[col] = col < p.n1_true ? __uint_as_float(r[j]) + __ldg(p.b1 + col) : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) zr[half * 32 + j] = 0.f;
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(zempty_bar);
        float scale = 1.f;
        if (p.l2norm) {
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 64; ++j) ss = fmaf(zr[j], zr[j], ss);
          scale = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        }
        const int64_t grow = (int64_t)tile * kBlockRows + row;
        if (grow < p.m) {
          float* out = p.z + grow * p.n1_true;
          if (p.n1_true == 64) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              reinterpret_cast<float4*>(out)[j] = make_float4(zr[4 * j] * scale, zr[4 * j + 1] * scale,
                                                              zr[4 * j + 2] * scale, zr[4 * j + 3] * scale);
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (j < p.n1_true) out[j] = zr[j] * scale;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// fp32 rows -> hi/lo bf16 planes in slab format: [tile][plane][kb] x 16 KB; rows past `n` are zero.
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ rows, int64_t n, int64_t n_pad,
                                                         int dim, int nkb, int np,
                                                         unsigned char* __restrict__ slabs) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_pad) return;
  const int rin = (int)(r % kBlockRows);
  unsigned char* tile = slabs + (r / kBlockRows) * ((int64_t)kPlanes * nkb * kSlabBytes);
  for (int kc = lane; kc < nkb * 8; kc += 32) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c0 = kc * 8 + 2 * e;
      const float x0 = (r < n && c0 < dim) ? __ldg(rows + r * dim + c0) : 0.f;
      const float x1 = (r < n && c0 + 1 < dim) ? __ldg(rows + r * dim + c0 + 1) : 0.f;
      const __nv_bfloat16 a0 = __float2bfloat16_rn(x0), a1 = __float2bfloat16_rn(x1);
      hi[e] = pack_bf16(x0, x1);
      lo[e] = pack_bf16(x0 - __bfloat162float(a0), x1 - __bfloat162float(a1));
    }
    const int64_t off = (int64_t)(kc >> 3) * kSlabBytes + slab_chunk_offset(rin, kc & 7);
    *reinterpret_cast<uint4*>(tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (np == 2)
      *reinterpret_cast<uint4*>(tile + (int64_t)nkb * kSlabBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- host: bf16 round-to-nearest-even and weight slabs -------------------------------------
inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

}  // namespace

int ae_umma_supported(int d_in, int d_hidden, int d_latent) {
  return d_in % 64 == 0 && d_in >= 64 && d_hidden % 128 == 0 && d_hidden >= 128 && d_latent >= 1 && d_latent <= 64;
}

size_t ae_umma_x_slab_bytes(int64_t m, int d_in) {
  const int64_t tiles = (m + kBlockRows - 1) / kBlockRows;
  return (size_t)tiles * kPlanes * (d_in / 64) * kSlabBytes;
}

// W [rows_out, k_in] fp32 (nn.Linear layout) -> [rowblock][plane][kb] slabs of `slab_rows` rows
void ae_umma_weight_slabs(const float* w, int rows_out, int k_in, int slab_rows, std::vector<unsigned char>* out,
                          bool plane_major_over_kb_only) {
  const int nkb = k_in / 64;
  const int nrb = (rows_out + slab_rows - 1) / slab_rows;
  const size_t slab = (size_t)slab_rows * kRowBytes;
  out->assign((size_t)nrb * kPlanes * nkb * slab, 0);
  for (int o = 0; o < rows_out; ++o) {
    const int rb = o / slab_rows, rin = o % slab_rows;
    for (int i = 0; i < k_in; ++i) {
      const float v = w[(size_t)o * k_in + i];
      const uint16_t hi = f2bf(v), lo = f2bf(v - bf2f(hi));
      const int kb = i / 64, c = (i % 64) / 8, e = i % 8;
      const size_t within = (size_t)rin * kRowBytes + (size_t)((c ^ (rin & 7)) << 4) + (size_t)e * 2;
      for (int pl = 0; pl < kPlanes; ++pl) {
        const size_t base = plane_major_over_kb_only ? ((size_t)pl * nkb + kb) * slab
                                                     : (((size_t)rb * kPlanes + pl) * nkb + kb) * slab;
        const uint16_t bits = pl == 0 ? hi : lo;
        memcpy(out->data() + base + within, &bits, 2);
      }
    }
  }
}

int launch_ae_split_rows(const float* x, int64_t m, int d_in, int n_planes, unsigned char* slabs, cudaStream_t st) {
  const int64_t m_pad = round_up64(m, kBlockRows);
  const unsigned grid = (unsigned)((m_pad + 7) / 8);
  split_rows_kernel<<<grid, 256, 0, st>>>(x, m, m_pad, d_in, d_in / 64, n_planes, slabs);
  LK_CHECK_LAUNCH("split_rows_kernel");
  return LK_OK;
}

int launch_ae_umma(const unsigned char* x_slabs, int64_t m, int d_in, int d_hidden, int d_latent,
                   const unsigned char* w0_slabs, const unsigned char* w1_slabs, const float* b0, const float* b1,
                   int l2norm, int n_planes, float* z, int* err_flag, int sm_count, cudaStream_t st) {
  AeParams p;
  p.np = n_planes == 1 ? 1 : 2;
  p.x_slabs = x_slabs;
  p.w0_slabs = w0_slabs;
  p.w1_slabs = w1_slabs;
  p.b0 = b0;
  p.b1 = b1;
  p.z = z;
  p.m = m;
  p.n_tiles = (int)((m + kBlockRows - 1) / kBlockRows);
  p.nkb0 = d_in / 64;
  p.n_chunks = d_hidden / kBlockRows;
  p.n1 = round_up(d_latent, 16);
  p.n1_true = d_latent;
  p.l2norm = l2norm;
  p.err_flag = err_flag;
  const size_t smem = kHeaderBytes + 1024 + (size_t)kStages * kStageBytes + kHBytes + (size_t)kPlanes * 2 * p.n1 * kRowBytes;
  LK_CUDA(cudaFuncSetAttribute(ae_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.n_tiles < sm_count ? p.n_tiles : sm_count;
  ae_umma_kernel<<<grid, kThreads, smem, st>>>(p);
  LK_CHECK_LAUNCH("ae_umma_kernel");
  return LK_OK;
}

}  // namespace lk
