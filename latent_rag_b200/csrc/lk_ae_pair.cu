// Fused autoencoder-encoder forward on CTA PAIRS (tcgen05 cta_group::2), bf16 operands:
//     Z = relu(X W0^T + b0) W1^T + b1 (+ L2 norm),   X fp32 straight from HBM.
//
// Replaces the encoder halves of the reference autoencoders (models/denoising_autoencoder.py:19-23,
// 33-34; models/contrastive_autoencoder.py:10-14,23-25; models/variational_autoencoder.py:11-16,27-28,
// mu only as retrieval/embedder.py:44-45 keeps) at the stated precision of the bf16 search path:
// inputs, weights and hidden activations rounded to bf16 once, fp32 accumulation, fp32 biases.
//
// What the single-CTA kernel (lk_ae_umma.cu) paid for and this one does not:
//   * a separate pass that split the fp32 rows into operand planes (28 % of the time, 2 extra
//     trips through HBM): here 8 converter warps read the fp32 rows with 256-bit loads, round them
//     and write the swizzled bf16 A operand into shared memory themselves;
//   * X re-streamed once per hidden chunk (4x): the 128 x d_in bf16 tile stays RESIDENT in shared
//     memory for all hidden chunks (96 KB at d_in = 384);
//   * the whole W0 chunk per CTA: a CTA pair works on two row tiles at once (M = 256: the leader's
//     tile -> the leader's TMEM, the peer's -> the peer's) and each CTA loads only ITS 64 of the
//     chunk's 128 weight rows (8 KB per K block instead of 16) -- half the L2 -> SM operand traffic
//     per row, 8 stages in the same shared memory; W1 (its half: 32 KB) is resident as well.
// Per row tile and CTA: 192 KB of X from HBM, 192 KB of W0 from L2; tensor work 7.2 k cycles.
//
// Warp roles (640 threads per CTA, 1 CTA / SM, clusters of 2):
//   warp 0        TMA producer: W1 half once, then the ring of W0 half-slab stages
//   warp 1        leader: MMA issuer (one elected lane);  peer: relays "stage landed" / "X slab ready"
//   warp 2        TMEM allocator
//   warp 3        peer: relays "hidden chunk ready" (leader: idle)
//   warps 4-11    epilogue: acc0 -> +b0, ReLU, bf16 -> swizzled H operand in smem; final Z rows
//   warps 12-19   converters: fp32 X rows -> bf16 slabs
// Pipelines (mbarriers; the leader's collect both CTAs where the leader's MMA warp is the consumer):
//   W0 ring      full / pfull (relay) -> MMA -> empty (commit, multicast)
//   X slabs      xfull[kb] / pxfull[kb] (relay, cluster-scope release: generic-proxy stores of the
//                peer, read by the tensor core under the LEADER's instruction) -> MMA -> xempty[kb]
//                (commit after the last hidden chunk's K block kb: the converters then overwrite slab
//                kb with the next tile while the tensor core still works on the later K blocks)
//   acc0[2]      a0full (commit, multicast) -> epilogues -> a0empty (both CTAs' warps, on the leader)
//   H            hfull / phfull (relay) -> layer-1 MMA -> hempty (commit, multicast)
//   acc1         zfull (commit, multicast) -> Z epilogues -> zempty (on the leader)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lk_common.cuh"
#include "lk_ptx.cuh"

namespace lk {

namespace {

constexpr int kPThreads = 640;
constexpr int kPFirstEpi = 4, kPEpiWarps = 8;
constexpr int kPFirstCvt = 12, kPCvtWarps = 8;
constexpr int kPMaxStages = 8, kPMaxKb0 = 6;
constexpr int kHalfSlab = kSlabBytes / 2;  // this CTA's 64 rows of a (hidden chunk, K block) slab of W0
#ifndef LK_AE_STAGE_KB
#define LK_AE_STAGE_KB 2
#endif
constexpr int kStageKb = LK_AE_STAGE_KB;   // K blocks per W0 stage: one barrier wait and one commit per 8 MMAs (the
                                           // issuing thread, not the tensor pipe, was the bottleneck at one K block)
constexpr int kPStageBytes = kStageKb * kHalfSlab;
constexpr int kPHeader = 1024;
constexpr int kPTmemCols = 512, kPAcc1Col = 256;
constexpr int kPXCol = 320;            // TMEM columns 320..511: the bf16 row tile, the A operand of layer 0 (32 per K block)
constexpr int kXsRowBytes = 2 * 64 * kPMaxKb0 + 16;  // staging row: 768 bytes of bf16 + 16 of padding (conflict-free reads)
constexpr int kPSmemBudget = 227 * 1024;

// The barriers the leader's MMA warp waits on collect BOTH CTAs: the leader's own arrivals plus one from the
// peer's relay warp (which waits for the same event on its own barrier first), so the issuing thread makes
// one wait per event instead of two.
enum PairBar {
  B_FULL = 0, B_EMPTY = 8, B_XFULL = 16, B_XEMPTY = 22, B_A0FULL = 28, B_A0EMPTY = 30,
  B_HFULL = 32, B_HEMPTY = 33, B_ZFULL = 34, B_ZEMPTY = 35, B_W1FULL = 36, B_COUNT = 37
};

enum PairErr { kPeProd = 401, kPeRelay = 402, kPeRelayX = 403, kPeRelayH = 404, kPeMmaA0 = 405, kPeMmaX = 406,
               kPeMmaFull = 407, kPeMmaH = 408, kPeMmaW1 = 409, kPeMmaZ = 410, kPeEpiA0 = 411, kPeEpiH = 412,
               kPeEpiZ = 413, kPeCvt = 414 };

struct PairParams {
  const float* x;                 // [m, d_in] fp32 row-major, 32-byte aligned
  const unsigned char* w0_slabs;  // [chunk][plane 0..1][kb0] slabs of 128 rows (plane 0 = bf16(W0) is read)
  const unsigned char* w1_slabs;  // [plane][kb1] slabs of n1 rows
  const float* b0;
  const float* b1;
  float* z;                       // [m, n1_true]
  int64_t m;
  int d_in, n_pair_tiles, nkb0, n_chunks, nkb1, n1, n1_true, l2norm, n_stages, z_vec, prefetch;
  int* err_flag;
  long long* prof;  // bring-up (LK_AE_PROF): per cluster, cycles the MMA warp spent waiting per barrier kind + total
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

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// 256-bit streaming load of the fp32 rows: read once, so it must not push the bias vectors out of L1
__device__ __forceinline__ void ldg256_stream(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
// bring `bytes` (a multiple of 16) of global memory into L2 ahead of the loads that will read it
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T on a CTA pair: the A operand (this CTA's 128 rows, K elements packed two per
// 32-bit column) is read from tensor memory, so layer 0 spends its shared-memory bandwidth on the weights only
// (with A in shared memory a 256 x 128 x 16 step reads 8 KB per 64 cycles: the whole 128 B/clk port)
__device__ __forceinline__ void umma_bf16_2cta_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns, registers -> tensor memory (thread i writes lane base_lane + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cvt_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // the 8 converter warps

__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic stores -> tensor-core (async proxy) reads
}

__global__ void __launch_bounds__(kPThreads, 1) ae_pair_kernel(const PairParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t bar0 = ptx::smem_u32(smem);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + 512);
  unsigned char* data = smem + kPHeader;
  data += (1024u - (ptx::smem_u32(data) & 1023u)) & 1023u;
  const int w1_half = (p.n1 / 2) * kRowBytes;  // this CTA's rows of one W1 K-block slab
  unsigned char* h_sm = data;                                    // 2 slabs: one hidden chunk (128 units)
  unsigned char* w1_sm = h_sm + 2 * kSlabBytes;                  // nkb1 half-slabs
  unsigned char* stage_sm = w1_sm + p.nkb1 * w1_half;            // n_stages stages of kStageKb W0 half-slabs
  float* b0_sm = reinterpret_cast<float*>(stage_sm + p.n_stages * kPStageBytes);  // n_chunks * 128 hidden biases
  float* b1_sm = b0_sm + p.n_chunks * kBlockRows;                              // 64 latent biases (zero padded)
  unsigned char* xs_sm = reinterpret_cast<unsigned char*>(b1_sm + 64);         // staging: the NEXT row tile in bf16, 128 padded rows

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();  // 0 = leader
  const int n_clusters = gridDim.x / 2, cid = blockIdx.x / 2;

  if (threadIdx.x == 0) {
    const uint32_t relay = rank == 0 ? 1u : 0u;  // + the peer's relay on the leader's barriers
    for (int s = 0; s < kPMaxStages; ++s) {
      ptx::mbar_init(bar(B_FULL + s), 1 + relay);
      ptx::mbar_init(bar(B_EMPTY + s), 1);
    }
    for (int k = 0; k < kPMaxKb0; ++k) {
      ptx::mbar_init(bar(B_XFULL + k), 2 * kPCvtWarps);   // both CTAs' converter warps, on the leader
      ptx::mbar_init(bar(B_XEMPTY + k), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(bar(B_A0FULL + a), 1);
      ptx::mbar_init(bar(B_A0EMPTY + a), 2 * kPEpiWarps);  // the leader's collects both CTAs' epilogue warps
    }
    ptx::mbar_init(bar(B_HFULL), 2 * kPEpiWarps);          // both CTAs' epilogue warps, on the leader
    ptx::mbar_init(bar(B_HEMPTY), 1);
    ptx::mbar_init(bar(B_ZFULL), 1);
    ptx::mbar_init(bar(B_ZEMPTY), 2 * kPEpiWarps);  // every epilogue warp of both CTAs
    ptx::mbar_init(bar(B_W1FULL), 1 + relay);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc2(ptx::smem_u32(tmem_ptr_s), kPTmemCols);
    ptx::tmem_relinquish2();
  }
  // the bias vectors live in shared memory: the streamed rows and the polling of the barriers leave
  // nothing in L1 for long (the first version spent a quarter of the epilogue's time on bias loads)
  for (int i = threadIdx.x; i < p.n_chunks * kBlockRows; i += kPThreads) b0_sm[i] = p.b0[i];
  if (threadIdx.x < 64) b1_sm[threadIdx.x] = (int)threadIdx.x < p.n1_true ? p.b1[threadIdx.x] : 0.f;
  ptx::tc_fence_before();
  ptx::cluster_sync_all();  // the peer's barriers exist before anyone arrives on them (also publishes the biases)
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  auto fail = [&](int code) {
    if (lane == 0) atomicCAS(p.err_flag, 0, code);
  };
  auto wait = [&](uint32_t b, uint32_t parity) { return __all_sync(0xffffffffu, ptx::mbar_wait(b, parity)); };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, each its own halves) =====================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(bar(B_W1FULL), (uint32_t)(p.nkb1 * w1_half));
      for (int kb = 0; kb < p.nkb1; ++kb)
        ptx::bulk_g2s(ptx::smem_u32(w1_sm + kb * w1_half),
                      p.w1_slabs + (int64_t)kb * p.n1 * kRowBytes + rank * w1_half, (uint32_t)w1_half, bar(B_W1FULL));
    }
    __syncwarp();
    Ring st;
    bool ok = true;
    for (int pt = cid; pt < p.n_pair_tiles && ok; pt += n_clusters) {
      for (int c = 0; c < p.n_chunks && ok; ++c) {
        const unsigned char* wc = p.w0_slabs + (int64_t)c * 2 * p.nkb0 * kSlabBytes + rank * kHalfSlab;  // plane 0
        for (int kb = 0; kb < p.nkb0; kb += kStageKb) {
          const int nk = p.nkb0 - kb < kStageKb ? p.nkb0 - kb : kStageKb;
          if (!wait(bar(B_EMPTY + st.idx), st.phase ^ 1u)) { fail(kPeProd); ok = false; break; }
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(bar(B_FULL + st.idx), (uint32_t)(nk * kHalfSlab));
            for (int j = 0; j < nk; ++j)
              ptx::bulk_g2s(ptx::smem_u32(stage_sm + st.idx * kPStageBytes + j * kHalfSlab),
                            wc + (int64_t)(kb + j) * kSlabBytes, kHalfSlab, bar(B_FULL + st.idx));
          }
          __syncwarp();
          st.advance(p.n_stages);
        }
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ===================== peer relay: W0 stages and X slabs, in the order layer 0 consumes them =====
    // (CTA-scope arrivals on the leader's barriers: they only order work that has already completed here --
    // a landed bulk copy, shared-memory stores already fenced to the async proxy by their writers)
    Ring st;
    uint32_t t_local = 0;
    bool ok = true;
    for (int pt = cid; pt < p.n_pair_tiles && ok; pt += n_clusters, ++t_local) {
      for (int c = 0; c < p.n_chunks && ok; ++c) {
        for (int kb = 0; kb < p.nkb0 && ok; kb += kStageKb) {
          const int nk = p.nkb0 - kb < kStageKb ? p.nkb0 - kb : kStageKb;
          if (!wait(bar(B_FULL + st.idx), st.phase)) { fail(kPeRelay); ok = false; break; }
          if (ptx::elect_one()) ptx::mbar_arrive_remote(bar(B_FULL + st.idx), 0);
          __syncwarp();
          st.advance(p.n_stages);
        }
      }
    }
  } else if (warp == 3 && rank != 0) {
    // ===================== peer relay: W1 landed, hidden chunks ready =====================
    bool ok = wait(bar(B_W1FULL), 0u);
    if (!ok) fail(kPeRelayH);
    if (ok && ptx::elect_one()) ptx::mbar_arrive_remote(bar(B_W1FULL), 0);
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader) =====================
    const uint32_t idesc0 = ptx::idesc_bf16_f32(2 * kBlockRows, kBlockRows);  // M = 256 (pair), N = 128 hidden units
    const uint32_t idesc1 = ptx::idesc_bf16_f32(2 * kBlockRows, p.n1);        // N = latent columns
    const uint64_t desc_hi = ptx::smem_desc(0, 16, 1024);
    auto desc = [&](const void* sm) { return desc_hi | (uint64_t)((ptx::smem_u32(sm) >> 4) & 0x3fffu); };
    // descriptors of the operand slabs; a K step of 16 elements is +32 bytes = +2 in the start-address field
    // (no carry out of it: every operand lies below 256 KB)
    const uint64_t h_desc = desc(h_sm), w1_desc = desc(w1_sm), st_desc = desc(stage_sm);
    constexpr uint64_t kSlabStep = kSlabBytes >> 4, kHalfStep = kHalfSlab >> 4, kStageStep = kPStageBytes >> 4;
    const uint64_t w1_step = (uint64_t)(w1_half >> 4);
    Ring st;
    uint32_t chunk_no = 0, l1_no = 0, t_local = 0;
    bool ok = true, w1_ready = false;
    long long tw[6] = {0, 0, 0, 0, 0, 0};  // a0empty, xfull, full, hfull, zempty, w1full
    const long long t_begin = clock64();
    auto twait = [&](int kind, uint32_t b, uint32_t parity) -> bool {
      if (p.prof == nullptr) return wait(b, parity) != 0;
      const long long t0 = clock64();
      const bool r = wait(b, parity) != 0;
      tw[kind] += clock64() - t0;
      return r;
    };
    // layer 1 of a hidden chunk is issued AFTER layer 0 of the next chunk -- across row tiles too -- so that
    // the tensor pipe works on the next accumulator while the epilogue warps turn this one into H
    bool pend = false;       // a chunk whose layer 1 has not been issued yet
    int pend_c = 0;
    uint32_t pend_tile = 0;
    auto layer1 = [&](int c, uint32_t tile_no) -> bool {
      if (!w1_ready) {
        if (!twait(5, bar(B_W1FULL), 0u)) { fail(kPeMmaW1); return false; }
        w1_ready = true;
      }
      if (!twait(3, bar(B_HFULL), l1_no & 1u)) { fail(kPeMmaH); return false; }
      if (c == 0 && !twait(4, bar(B_ZEMPTY), (tile_no & 1u) ^ 1u)) { fail(kPeMmaZ); return false; }
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_2cta(tmem_base + kPAcc1Col, h_desc + j * kSlabStep + 2u * k,
                                w1_desc + (uint64_t)(2 * c + j) * w1_step + 2u * k, idesc1, (c | j | k) != 0 ? 1u : 0u);
        ptx::umma_commit_2cta(bar(B_HEMPTY), 3);
        if (c == p.n_chunks - 1) ptx::umma_commit_2cta(bar(B_ZFULL), 3);
      }
      __syncwarp();
      ++l1_no;
      return true;
    };
    for (int pt = cid; pt < p.n_pair_tiles && ok; pt += n_clusters, ++t_local) {
      for (int c = 0; c < p.n_chunks && ok; ++c) {
        // ---- layer 0 of hidden chunk c
        const uint32_t a = chunk_no & 1u;
        if (!twait(0, bar(B_A0EMPTY + a), ((chunk_no >> 1) & 1u) ^ 1u)) { fail(kPeMmaA0); ok = false; break; }
        for (int kb = 0; kb < p.nkb0 && ok; kb += kStageKb) {
          const int nk = p.nkb0 - kb < kStageKb ? p.nkb0 - kb : kStageKb;
          if (c == 0) {  // the row tiles of both CTAs
            for (int j = 0; j < nk; ++j)
              if (!twait(1, bar(B_XFULL + kb + j), t_local & 1u)) { fail(kPeMmaX); ok = false; }
            if (!ok) break;
          }
          if (!twait(2, bar(B_FULL + st.idx), st.phase)) { fail(kPeMmaFull); ok = false; break; }
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t xa = tmem_base + kPXCol + (uint32_t)kb * 32u;  // 64 bf16 of a K block = 32 columns
            const uint64_t wb = st_desc + (uint64_t)st.idx * kStageStep;
#pragma unroll
            for (int j = 0; j < kStageKb; ++j) {
              if (j < nk) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_2cta_ts(tmem_base + a * kBlockRows, xa + (uint32_t)(j * 32 + k * 8), wb + j * kHalfStep + 2u * k,
                                    idesc0, (kb | j | k) != 0 ? 1u : 0u);
              }
            }
            ptx::umma_commit_2cta(bar(B_EMPTY + st.idx), 3);
            if (c == p.n_chunks - 1)  // these slabs of X: last reader done
              for (int j = 0; j < nk; ++j) ptx::umma_commit_2cta(bar(B_XEMPTY + kb + j), 3);
          }
          __syncwarp();
          st.advance(p.n_stages);
        }
        if (!ok) break;
        if (ptx::elect_one()) ptx::umma_commit_2cta(bar(B_A0FULL + a), 3);
        __syncwarp();
        ++chunk_no;
        // ---- layer 1 of the chunk before
        if (pend && !layer1(pend_c, pend_tile)) { ok = false; break; }
        pend = true;
        pend_c = c;
        pend_tile = t_local;
      }
    }
    if (ok && pend) layer1(pend_c, pend_tile);
    if (p.prof != nullptr && lane == 0) {
      for (int i = 0; i < 6; ++i) p.prof[cid * 8 + i] = tw[i];
      p.prof[cid * 8 + 6] = clock64() - t_begin;
      p.prof[cid * 8 + 7] = t_local;
    }
  } else if (warp >= kPFirstCvt) {
    // ===================== converters: fp32 rows -> bf16 A operand in tensor memory =====================
    // Phase A (a whole tile ahead of its use): thread ct owns the 8-column piece `piece` of rows
    // (ct >> 3) + 32 j of every K block -- a warp reads 4 rows x 256 contiguous bytes per load instruction --
    // rounds it to bf16 and parks it in the shared-memory staging tile.  Phase B (as the tensor core releases the
    // K blocks of the tile before): thread = row (TMEM lane), two warps per lane quarter share a K block's 32
    // columns: 64 staged bytes -> tcgen05.st.  The bar.sync pairs order staging reuse among the 8 warps; they are
    // executed unconditionally (a timed-out wait only skips the work between them).
    const int ct = (int)threadIdx.x - kPFirstCvt * 32;
    const int piece = ct & 7, rbase = ct >> 3;
    const int cw = warp - kPFirstCvt, quarter = cw & 3, hsel = cw >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t x_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + kPXCol + (uint32_t)(hsel * 16);
    uint32_t t_local = 0;
    bool ok = true;
    auto stage_tile = [&](int pt) {  // phase A
      const int64_t row0 = ((int64_t)pt * 2 + rank) * kBlockRows;
#pragma unroll
      for (int kb = 0; kb < kPMaxKb0; ++kb) {
        if (kb < p.nkb0) {
          float v[4][8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int64_t grow = row0 + rbase + 32 * j;
            if (grow < p.m) {
              ldg256_stream(p.x + grow * p.d_in + kb * 64 + piece * 8, v[j]);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[j][e] = 0.f;
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(xs_sm + (rbase + 32 * j) * kXsRowBytes + kb * 128 + piece * 16) =
                make_uint4(pack2(v[j][0], v[j][1]), pack2(v[j][2], v[j][3]), pack2(v[j][4], v[j][5]), pack2(v[j][6], v[j][7]));
        }
      }
    };
    if (cid < p.n_pair_tiles) stage_tile(cid);
    cvt_bar_sync();
    for (int pt = cid; pt < p.n_pair_tiles; pt += n_clusters, ++t_local) {
      // phase B: staged tile -> tensor memory, K block by K block as the previous tile's readers finish
      for (int kb = 0; kb < p.nkb0 && ok; ++kb) {
        if (!wait(bar(B_XEMPTY + kb), (t_local & 1u) ^ 1u)) { fail(kPeCvt); ok = false; break; }
        ptx::tc_fence_after();
        uint32_t r[16];
        const unsigned char* src = xs_sm + row * kXsRowBytes + kb * 128 + hsel * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 q4 = *reinterpret_cast<const uint4*>(src + j * 16);
          r[4 * j] = q4.x; r[4 * j + 1] = q4.y; r[4 * j + 2] = q4.z; r[4 * j + 3] = q4.w;
        }
        tmem_st16(x_taddr + (uint32_t)kb * 32u, r);
        tmem_wait_st();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {  // straight onto the leader's barrier (no relay hop)
          if (rank != 0) ptx::mbar_arrive_remote(bar(B_XFULL + kb), 0);
          else ptx::mbar_arrive(bar(B_XFULL + kb));
        }
      }
      cvt_bar_sync();  // every warp has read its part of the staging tile
      if (pt + n_clusters < p.n_pair_tiles && ok) {
        {  // the tile after the next goes to L2 now (this warp's 16 rows of it)
          const int64_t nrow0 = ((int64_t)(pt + 2 * n_clusters) * 2 + rank) * kBlockRows + cw * (kBlockRows / kPCvtWarps);
          if (p.prefetch && pt + 2 * n_clusters < p.n_pair_tiles && nrow0 < p.m && lane == 0) {
            const int64_t rows = p.m - nrow0 < kBlockRows / kPCvtWarps ? p.m - nrow0 : kBlockRows / kPCvtWarps;
            prefetch_l2(p.x + nrow0 * p.d_in, (uint32_t)(rows * p.d_in * sizeof(float)));
          }
        }
        stage_tile(pt + n_clusters);
      }
      cvt_bar_sync();  // the staging tile is complete
    }
  } else if (warp >= kPFirstEpi) {
    // ===================== epilogue =====================
    const int ew = warp - kPFirstEpi;
    const int quarter = warp & 3;
    const int ch = ew >> 2;                  // which 64 of the chunk's 128 hidden units (= K block of H)
    const int row = quarter * 32 + lane;     // row of the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    unsigned char* hrow = h_sm + ch * kSlabBytes + row * kRowBytes;
    uint32_t g = 0, t_local = 0;
    bool ok = true;
    int z_pend_pt = -1;
    uint32_t z_pend_tile = 0;
    // Z row = acc1 + b1 (+ L2 normalisation): a thread owns one row and 32 of its 64 latent columns.  Run AFTER the first
    // hidden chunk of the next tile has been handed to layer 1: the tensor pipe is busy with that tile's second
    // chunk meanwhile, and layer 1 of the next tile only needs acc1 back after it
    auto z_epilogue = [&](int z_pt, uint32_t z_tile) -> bool {
      if (!wait(bar(B_ZFULL), z_tile & 1u)) { fail(kPeEpiZ); return false; }
      ptx::tc_fence_after();
      const int64_t grow = ((int64_t)z_pt * 2 + rank) * kBlockRows + row;
      // this warp's 32 of the 64 latent columns (ch picks the half), one TMEM read, values kept in registers
      float zv[32];
      const bool have = ch * 32 < p.n1;
      float ss = 0.f;
      if (p.l2norm && (1 - ch) * 32 < p.n1) {  // the other half of the row: only its squares, for the L2 norm
        uint32_t ro[32];
        ptx::tmem_ld32(tmem_base + lane_addr + kPAcc1Col + (1 - ch) * 32, ro);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = (1 - ch) * 32 + j;
          const float o = col < p.n1_true ? __uint_as_float(ro[j]) + b1_sm[col] : 0.f;
          ss = fmaf(o, o, ss);
        }
      }
      if (have) {  // this half: kept in registers until the store
        uint32_t r[32];
        ptx::tmem_ld32(tmem_base + lane_addr + kPAcc1Col + ch * 32, r);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          zv[j] = ch * 32 + j < p.n1_true ? __uint_as_float(r[j]) + b1_sm[ch * 32 + j] : 0.f;
          ss = fmaf(zv[j], zv[j], ss);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) zv[j] = 0.f;
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // acc1 goes back to the (leader's) MMA warp as soon as it has been read
        if (rank != 0) ptx::mbar_arrive_remote(bar(B_ZEMPTY), 0);
        else ptx::mbar_arrive(bar(B_ZEMPTY));
      }
      const float scale = p.l2norm ? 1.0f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
      if (have && grow < p.m) {
        float* out = p.z + grow * p.n1_true + ch * 32;
        if (p.z_vec) {
#pragma unroll
          for (int q8 = 0; q8 < 4; ++q8) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = zv[q8 * 8 + e] * scale;
            ptx::stg256(out + q8 * 8, o);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ch * 32 + j < p.n1_true) out[j] = zv[j] * scale;
        }
      }
      return true;
    };
    for (int pt = cid; pt < p.n_pair_tiles && ok; pt += n_clusters, ++t_local) {
      for (int c = 0; c < p.n_chunks; ++c, ++g) {
        const uint32_t a = g & 1u;
        if (!wait(bar(B_A0FULL + a), (g >> 1) & 1u)) { fail(kPeEpiA0); ok = false; break; }
        ptx::tc_fence_after();
        const float* bias = b0_sm + c * kBlockRows + ch * 64;
        uint32_t hpk[32];
        {
          uint32_t r0[32], r1[32];  // both halves in flight together
          ptx::tmem_ld32(tmem_base + lane_addr + a * kBlockRows + ch * 64, r0);
          ptx::tmem_ld32(tmem_base + lane_addr + a * kBlockRows + ch * 64 + 32, r1);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t* r = half == 0 ? r0 : r1;
#pragma unroll
            for (int q8 = 0; q8 < 4; ++q8) {  // 8 hidden units at a time: two bias vectors, four packed words
              const float4 ba = *reinterpret_cast<const float4*>(bias + half * 32 + q8 * 8);  // same address in every lane
              const float4 bb = *reinterpret_cast<const float4*>(bias + half * 32 + q8 * 8 + 4);
              const uint32_t* rr = r + q8 * 8;
              uint32_t* o = hpk + half * 16 + q8 * 4;
              o[0] = pack2(fmaxf(__uint_as_float(rr[0]) + ba.x, 0.f), fmaxf(__uint_as_float(rr[1]) + ba.y, 0.f));
              o[1] = pack2(fmaxf(__uint_as_float(rr[2]) + ba.z, 0.f), fmaxf(__uint_as_float(rr[3]) + ba.w, 0.f));
              o[2] = pack2(fmaxf(__uint_as_float(rr[4]) + bb.x, 0.f), fmaxf(__uint_as_float(rr[5]) + bb.y, 0.f));
              o[3] = pack2(fmaxf(__uint_as_float(rr[6]) + bb.z, 0.f), fmaxf(__uint_as_float(rr[7]) + bb.w, 0.f));
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {  // the accumulator goes back to the (leader's) MMA warp before H is written
          if (rank != 0) ptx::mbar_arrive_remote(bar(B_A0EMPTY + a), 0);
          else ptx::mbar_arrive(bar(B_A0EMPTY + a));
        }
        // the H buffer is free once layer 1 of the previous chunk has been read by the tensor cores
        if (!wait(bar(B_HEMPTY), (g & 1u) ^ 1u)) { fail(kPeEpiH); ok = false; break; }
#pragma unroll
        for (int cj = 0; cj < 8; ++cj)
          *reinterpret_cast<uint4*>(hrow + ((cj ^ (row & 7)) << 4)) =
              make_uint4(hpk[4 * cj], hpk[4 * cj + 1], hpk[4 * cj + 2], hpk[4 * cj + 3]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {  // straight onto the leader's barrier (no relay hop)
          if (rank != 0) ptx::mbar_arrive_remote(bar(B_HFULL), 0);
          else ptx::mbar_arrive(bar(B_HFULL));
        }
        if (c == 0 && z_pend_pt >= 0) {  // the previous tile's latent rows
          if (!z_epilogue(z_pend_pt, z_pend_tile)) { ok = false; break; }
          z_pend_pt = -1;
        }
      }
      if (!ok) break;
      z_pend_pt = pt;
      z_pend_tile = t_local;
    }
    if (ok && z_pend_pt >= 0) z_epilogue(z_pend_pt, z_pend_tile);
  }

  // ===================== teardown =====================
  ptx::tc_fence_before();
  ptx::cluster_sync_all();  // nobody touches the peer's barriers, shared memory or TMEM any more
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc2(tmem_base, kPTmemCols);
  }
}

inline int pair_stages(int d_in, int d_hidden, int n1) {
  (void)d_in;  // the row tile lives in tensor memory; its staging copy is sized for the widest supported row
  const int fixed = kPHeader + 1024 + 2 * kSlabBytes + (d_hidden / 64) * (n1 / 2) * kRowBytes +
                    (d_hidden + 64) * (int)sizeof(float) + kBlockRows * kXsRowBytes;
  int s = (kPSmemBudget - fixed) / kPStageBytes;
  return s > kPMaxStages ? kPMaxStages : s;
}

}  // namespace

// bf16-operand encoder on CTA pairs: d_in a multiple of 64 up to 384 (the bf16 row tile stays in shared
// memory), d_hidden a multiple of 128, d_latent <= 64, an even number of SMs, 3+ W0 stages
int ae_pair_supported(int d_in, int d_hidden, int d_latent, int sm_count) {
  if (d_in % 64 != 0 || d_in < 64 || d_in > 64 * kPMaxKb0 || d_hidden % 128 != 0 || d_hidden < 128 || d_latent < 1 ||
      d_latent > 64 || sm_count % 2 != 0)
    return 0;
  return pair_stages(d_in, d_hidden, round_up(d_latent, 16)) >= 2;
}

int launch_ae_pair(const float* x, int64_t m, int d_in, int d_hidden, int d_latent, const unsigned char* w0_slabs,
                   const unsigned char* w1_slabs, const float* b0, const float* b1, int l2norm, float* z, int* err_flag,
                   int sm_count, cudaStream_t st) {
  PairParams p;
  p.x = x;
  p.w0_slabs = w0_slabs;
  p.w1_slabs = w1_slabs;
  p.b0 = b0;
  p.b1 = b1;
  p.z = z;
  p.m = m;
  p.d_in = d_in;
  const int n_tiles = (int)((m + kBlockRows - 1) / kBlockRows);
  p.n_pair_tiles = (n_tiles + 1) / 2;
  p.nkb0 = d_in / 64;
  p.n_chunks = d_hidden / kBlockRows;
  p.nkb1 = d_hidden / 64;
  p.n1 = round_up(d_latent, 16);
  p.n1_true = d_latent;
  p.l2norm = l2norm;
  p.n_stages = pair_stages(d_in, d_hidden, p.n1);
  if (const char* e = getenv("LK_AE_STAGES")) {  // bring-up override
    const int v = atoi(e);
    if (v >= 2 && v <= p.n_stages) p.n_stages = v;
  }
  p.prefetch = 1;
  if (const char* e = getenv("LK_AE_PF")) p.prefetch = atoi(e) != 0;  // bring-up: L2 prefetch of the next row tile
  p.z_vec = (d_latent == 64 && (reinterpret_cast<uintptr_t>(z) & 31u) == 0) ? 1 : 0;
  p.err_flag = err_flag;
  const size_t smem = (size_t)kPHeader + 1024 + 2 * kSlabBytes + (size_t)p.nkb1 * (p.n1 / 2) * kRowBytes +
                      (size_t)p.n_stages * kPStageBytes + (size_t)(d_hidden + 64) * sizeof(float) +
                      (size_t)kBlockRows * kXsRowBytes;
  LK_CUDA(cudaFuncSetAttribute(ae_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int clusters = p.n_pair_tiles < sm_count / 2 ? p.n_pair_tiles : sm_count / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * 2));
  cfg.blockDim = dim3(kPThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  p.prof = nullptr;
  const bool prof = getenv("LK_AE_PROF") != nullptr;  // bring-up: where the MMA warp waits (synchronises, prints to stderr)
  if (prof) {
    LK_CUDA(cudaMalloc((void**)&p.prof, (size_t)clusters * 8 * sizeof(long long)));
    LK_CUDA(cudaMemsetAsync(p.prof, 0, (size_t)clusters * 8 * sizeof(long long), st));
  }
  LK_CUDA(cudaLaunchKernelEx(&cfg, ae_pair_kernel, p));
  LK_CHECK_LAUNCH("ae_pair_kernel");
  if (prof) {
    std::vector<long long> h((size_t)clusters * 8);
    LK_CUDA(cudaStreamSynchronize(st));
    LK_CUDA(cudaMemcpy(h.data(), p.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.prof);
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < clusters; ++c)
      for (int i = 0; i < 8; ++i) a[i] += (double)h[(size_t)c * 8 + i] / clusters;
    fprintf(stderr, "[ae_pair] m=%lld tiles/cluster=%.1f  MMA warp cycles: total %.0f  waits: a0empty %.0f  xfull %.0f  full %.0f  "
            "hfull %.0f  zempty %.0f  w1full %.0f  (per tile: total %.0f)\n", (long long)m, a[7], a[6], a[0], a[1], a[2], a[3],
            a[4], a[5], a[6] / (a[7] > 0 ? a[7] : 1));
  }
  return LK_OK;
}

}  // namespace lk
