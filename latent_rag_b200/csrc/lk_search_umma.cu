// Fused distance + top-k selection on the 5th-gen tensor cores (tcgen05 / TMEM / TMA).
//
// Replaces retrieval/bruteforce.py:66-82 of the reference (`q @ emb.T`, the euclidean
// expansion, `torch.topk`) and IndexFlatIP.search behind
// retrieval/FAISSEmbeddingRetriever.py:322 with ONE kernel in which the [B, N] score
// matrix only ever exists as 128x128 fp32 tiles in tensor memory.
//
// Shape of the computation
//   unit       = (query tile of 128 queries) x (PAIR of row blocks = 256 corpus rows)
//   D[q][row]  = sum_k Q[q][k] * E[row][k]       tcgen05.mma M=128 (queries -> TMEM lanes)
//                                                 N=256 (rows -> TMEM columns), K=16/instr
//   CTA c owns the contiguous unit range [c*U/G, (c+1)*U/G) of the query-tile-major unit
//   order, so a CTA changes query tile at most once when there are <= G query tiles and
//   CTAs whose ranges start at the same corpus offset share the row blocks through L2.
//
// Warp roles (384 threads, 1 CTA / SM)
//   warp 0      TMA producer: two 16 KB cp.async.bulk (one per row block) per K block of 64
//               into a ring of 32 KB shared-memory stages; also brings the query tile in
//   warp 1      MMA issuer (one elected lane): 4 tcgen05.mma (128x256x16) per K block,
//               accumulating a unit in one of 2 TMEM stages of 256 columns; tcgen05.commit
//               frees smem stages and publishes finished accumulators
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: TMEM lane = query, so each thread owns ONE query (for half of
//               the columns) and keeps its running top-k in registers: tcgen05.ld 32
//               columns, apply the metric (cosine norm / L2 expansion) with the per-row
//               side value, compare against the k-th best, insert on the rare hit.
//
// The operand slabs in HBM are already the tcgen05 K-major SWIZZLE_128B shared-memory
// image (lk_common.cuh), so the producer needs no tensor map: every copy is a contiguous
// 16 KB UBLKCP into a 1024-aligned stage.
#include <cstdlib>

#include "lk_ptx.cuh"
#include "lk_topk.cuh"

namespace lk {

namespace {

#ifndef LK_EPI_WARPS
#define LK_EPI_WARPS 8  // experiment: 16 = four epilogue warps per lane quarter, 64 columns each, one chunk per turn
#endif
constexpr int kThreads = 128 + 32 * LK_EPI_WARPS;
constexpr int kFirstEpiWarp = 4;
constexpr int kEpiWarps = LK_EPI_WARPS;
constexpr int kColSplit = kEpiWarps / 4;                 // epilogue warps per lane quarter
constexpr int kUnitBlocks = 2;                           // row blocks per unit
constexpr int kUnitCols = kUnitBlocks * kBlockRows;      // 256 accumulator columns per unit
constexpr int kColsPerWarp = kUnitCols / kColSplit;      // 128: one row block per epilogue warp
constexpr int kAccStages = 2;                            // 2 x 256 columns = all of TMEM
constexpr int kTmemCols = kAccStages * kUnitCols;        // 512
constexpr int kKBlockElems = kRowBytes / 2;              // 64 bf16 per row per K block
constexpr int kKBlockBytes = kSlabBytes;                 // 16384
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kBarBytes = 512;                           // barriers, TMEM pointer, abort flag
constexpr int kStashStride = 9;                          // floats per lane: 8 scores of a column group + 1 of padding
constexpr int kStashBytes = kEpiWarps * 32 * kStashStride * 4;
constexpr int kHeaderBytes = kBarBytes + kEpiWarps * 2 * kColsPerWarp * 4 + kStashBytes;  // + side rings + score stash
constexpr int kAlignSlack = 1024;                        // stages must be 1024-aligned (swizzle atom)
constexpr uint32_t kLbo = 16;                            // unused by swizzled K-major layouts
constexpr uint32_t kSbo = 8 * kRowBytes;                 // 1024: next 8-row group
constexpr uint32_t kUmmaKBytes = 32;                     // 16 bf16 = one MMA's K extent inside a row

enum UmmaErr {
  kErrProdEmpty = 101, kErrProdQEmpty = 102, kErrMmaFull = 103, kErrMmaTmemEmpty = 104,
  kErrMmaQFull = 105, kErrEpiTmemFull = 106, kErrMmaPeerFull = 107, kErrMmaPeerQFull = 108,
  kErrRelayFull = 109, kErrRelayQFull = 110
};

struct UmmaParams {
  const unsigned char* tiles;
  const float* side;
  const unsigned char* q_tiles;
  const float* q_side;
  int64_t n_queries;
  int64_t total_units;   // n_qt * nblk
  int nblk;              // row-block PAIRS in the corpus (units per query tile)
  int nkb;               // K blocks of 64 the MMA loop runs over per unit (split mode: 3 * split_n)
  int nkb_q;             // K-block slabs of a query tile in memory (= nkb; split mode: 2 * split_n)
  int nkb_res;           // the first nkb_res of them stay in shared memory for a whole segment (QRES: all of them);
                         // the others travel through the pipeline stages next to the corpus slabs of their K step
  int split_n;           // 0, or the K blocks per PLANE of split-bf16 operands (fp32 storage): rows and queries
                         // are stored as [hi plane | lo plane] and step kb multiplies q(hi,hi,lo) by e(hi,lo,hi)
  int n_stages;
  int n_lists;
  int64_t block_bytes;
  int ksel;              // entries per partial list slot (KSEL, or kBufCap for the buffer selector)
  int k;                 // requested k
  float* part_scores;    // [n_queries, n_lists, ksel]
  int32_t* part_idx;
  int* part_cnt;         // [n_queries, n_lists] entries left in each list slot (buffer selector), or null
  int* err_flag;
  const float* seed;     // [n_queries, k] best-first scores over a corpus sample, or null
  float* debug_tile;     // [128 queries][128 rows] raw dot products of unit 0, or null
  uint32_t lbo, sbo;
  int dbg;
  int rotate;            // corpus rotation per query-tile group (rot_cut)
};

__host__ __device__ inline int64_t unit_begin(int64_t c, int64_t total, int64_t grid) {
  return c * total / grid;
}
// the CTA whose range contains unit u
__host__ __device__ inline int64_t cta_of_unit(int64_t u, int64_t total, int64_t grid) {
  return ((u + 1) * grid - 1) / total;
}

// Corpus rotation.  Unit b of query-tile group qg works on corpus row-block pair
// (b - cut(qg)) mod nblk, where cut(qg) is the offset of the first scheduling-group boundary
// inside qg's unit range.  Every scheduling group that lies wholly inside one query-tile group
// then STARTS at a multiple of the per-group share of the corpus, so groups working for
// different query tiles stream the same row blocks at the same time and HBM sees the corpus
// about (1 + leftovers) times per batch instead of once per query-tile group and L2 sharing
// set.  A group boundary sits exactly at the wrap point, so the rows of every (group, query
// tile) segment are still visited in ascending order (the tie rule relies on it).
__host__ __device__ inline int64_t rot_cut(int64_t qg, int64_t nblk, int64_t total, int64_t grid) {
  const int64_t x = qg * nblk;
  const int64_t c = (x * grid + total - 1) / total;  // first group starting at or after x
  return c * total / grid - x;
}
__host__ __device__ inline int rot_block(int64_t qg, int b, int64_t nblk, int64_t total, int64_t grid, int rotate) {
  if (!rotate) return b;
  int cb = b - (int)rot_cut(qg, nblk, total, grid);
  return cb < 0 ? cb + (int)nblk : cb;
}

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

// CG = 1: one CTA per scheduling group (above).  CG = 2: a CTA PAIR per group
// (tcgen05 cta_group::2): the pair works on two query tiles at once (M = 256: the leader's tile
// -> the leader's TMEM, the peer's tile -> the peer's), each CTA loads only ITS half of the 256
// corpus rows of a unit (16 KB per K block instead of 32: half the L2 -> SM traffic and
// shared-memory operand reads per FLOP, twice the pipeline depth in the same shared memory).
// Only the leader issues MMAs; its commits are multicast to the barriers of both CTAs; the
// peer's warp 1 relays "my stage / query tile has landed" to the leader, and the peer's
// epilogue warps release the accumulator on the leader's barrier.
template <int KSEL, int METRIC, bool QRES, int CG>
__global__ void __launch_bounds__(kThreads, 1) umma_search_kernel(const UmmaParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // barrier slots: full[8] empty[8] peer_full[8] tfull[2] tempty[2] qfull qempty peer_qfull
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto pfull_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (3 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (3 * kMaxStages + kAccStages + s); };
  const uint32_t qfull_bar = bar0 + 8u * (3 * kMaxStages + 2 * kAccStages);
  const uint32_t qempty_bar = qfull_bar + 8u;
  const uint32_t pqfull_bar = qfull_bar + 16u;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + kBarBytes - 16);
  volatile int* abort_s = reinterpret_cast<volatile int*>(smem + kBarBytes - 12);
  float* side_ring = reinterpret_cast<float*>(smem + kBarBytes);
  unsigned char* q_sm = smem + kHeaderBytes;
  q_sm += (1024u - (ptx::smem_u32(q_sm) & 1023u)) & 1023u;  // swizzle atoms are 1024-byte aligned
  const int n_res = QRES ? p.nkb_q : p.nkb_res;  // query K blocks resident for a segment
  const bool res = QRES || n_res > 0;
  unsigned char* stage_sm = q_sm + n_res * kKBlockBytes;
  constexpr int kSlabsPerStage = kUnitBlocks / CG;  // corpus row blocks this CTA loads per K block
  constexpr int kStageBytes = (QRES ? kSlabsPerStage : kSlabsPerStage + 1) * kKBlockBytes;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int rank = CG == 2 ? (int)ptx::cluster_ctarank() : 0;  // 0 = leader
  const int64_t G = gridDim.x / CG, c = blockIdx.x / CG;        // scheduling groups, this CTA's group
  const int64_t u0 = unit_begin(c, p.total_units, G), u1 = unit_begin(c + 1, p.total_units, G);
  // split-bf16 operands (x = hi + lo, fp32-level products as hi*hi + hi*lo + lo*hi): K step kb of the
  // tripled loop reads corpus slab e_kb(kb) = (hi.., lo.., hi..) and query slab q_kb(kb) = (hi.., hi.., lo..)
  auto e_kb = [&](int kb) { return p.split_n && kb >= 2 * p.split_n ? kb - 2 * p.split_n : kb; };
  auto q_kb = [&](int kb) { return p.split_n && kb >= p.split_n ? kb - p.split_n : kb; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
      ptx::mbar_init(pfull_bar(s), 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);
      ptx::mbar_init(tempty_bar(s), kEpiWarps * CG);  // the leader's collects both CTAs' epilogue warps
    }
    ptx::mbar_init(qfull_bar, 1);
    ptx::mbar_init(qempty_bar, 1);
    ptx::mbar_init(pqfull_bar, 1);
    *abort_s = 0;
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) {
      ptx::tmem_alloc2(ptx::smem_u32(tmem_ptr_s), kTmemCols);
      ptx::tmem_relinquish2();
    } else {
      ptx::tmem_alloc(ptx::smem_u32(tmem_ptr_s), kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync_all();  // the peer's barriers exist before anyone arrives on them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  auto fail = [&](int code) {
    if (lane == 0) atomicCAS(p.err_flag, 0, code);
    *abort_s = 1;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop with warp-uniform values (so they live in uniform
    // registers); one elected lane issues the copies.
    Ring st;
    int seg = 0;
    int qt = (int)(u0 / p.nblk) * CG + rank, b = (int)(u0 % p.nblk);
    int cb = rot_block(u0 / p.nblk, b, p.nblk, p.total_units, G, p.rotate);  // corpus row-block pair of unit b
    bool ok = true;
    // one query tile = every row block is read exactly once: stream it evict-first so that the
    // query tile and the partial lists stay in L2; with several query tiles the row blocks are
    // shared through L2 by the CTAs working on the same corpus offset, so they keep the default
    const bool stream_once = p.total_units == p.nblk && !(p.dbg & 2);
    const uint64_t stream_policy = ptx::policy_evict_first();
    for (int64_t u = u0; u < u1 && ok; ++u) {
      const unsigned char* q_src = p.q_tiles + (int64_t)qt * p.block_bytes;
      if (res && (u == u0 || b == 0)) {
        if (seg > 0 && !__all_sync(0xffffffffu, ptx::mbar_wait(qempty_bar, (uint32_t)((seg - 1) & 1)))) {
          fail(kErrProdQEmpty);
          break;
        }
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(qfull_bar, (uint32_t)(n_res * kKBlockBytes));
          for (int kb = 0; kb < n_res; ++kb)
            ptx::bulk_g2s(ptx::smem_u32(q_sm + kb * kKBlockBytes), q_src + (int64_t)kb * kKBlockBytes,
                          kKBlockBytes, qfull_bar);
        }
        __syncwarp();
        ++seg;
      }
      const unsigned char* e_src = p.tiles + (int64_t)cb * kUnitBlocks * p.block_bytes;
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (!__all_sync(0xffffffffu, ptx::mbar_wait(empty_bar(st.idx), st.phase ^ 1u))) {
          fail(kErrProdEmpty);
          ok = false;
          break;
        }
        const bool q_streamed = !QRES && q_kb(kb) >= n_res;  // this K step's query slab rides in the stage
        if (ptx::elect_one()) {
          const uint32_t dst = ptx::smem_u32(stage_sm + st.idx * kStageBytes);
          ptx::mbar_arrive_expect_tx(full_bar(st.idx), (uint32_t)((kSlabsPerStage + (q_streamed ? 1 : 0)) * kKBlockBytes));
#pragma unroll
          for (int h = 0; h < kSlabsPerStage; ++h) {  // rows 0-127 and 128-255 of the N=256 operand (CG=2: this CTA's half)
            const unsigned char* src = e_src + (h + rank * kSlabsPerStage) * p.block_bytes + (int64_t)e_kb(kb) * kKBlockBytes;
            if (stream_once) ptx::bulk_g2s_hint(dst + h * kKBlockBytes, src, kKBlockBytes, full_bar(st.idx), stream_policy);
            else ptx::bulk_g2s(dst + h * kKBlockBytes, src, kKBlockBytes, full_bar(st.idx));
          }
          if (q_streamed)
            ptx::bulk_g2s(dst + kSlabsPerStage * kKBlockBytes, q_src + (int64_t)q_kb(kb) * kKBlockBytes, kKBlockBytes,
                          full_bar(st.idx));
        }
        __syncwarp();
        st.advance(p.n_stages);
      }
      if (++cb == p.nblk) cb = 0;
      if (++b == p.nblk) {
        b = 0;
        qt += CG;
        cb = rot_block(qt / CG, 0, p.nblk, p.total_units, G, p.rotate);
      }
    }
  } else if (warp == 1 && CG == 2 && rank != 0) {
    // ===================== peer relay =====================
    // Tells the leader's MMA warp that THIS CTA's copies have landed (a bulk copy can only
    // complete on a barrier of its own CTA).
    Ring st;
    int seg = 0;
    int b = (int)(u0 % p.nblk);
    bool ok = true;
    for (int64_t u = u0; u < u1 && ok; ++u) {
      if (res && (u == u0 || b == 0)) {
        if (!__all_sync(0xffffffffu, ptx::mbar_wait(qfull_bar, (uint32_t)(seg & 1)))) {
          fail(kErrRelayQFull);
          break;
        }
        if (ptx::elect_one()) ptx::mbar_arrive_remote(pqfull_bar, 0);
        __syncwarp();
        ++seg;
      }
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (!__all_sync(0xffffffffu, ptx::mbar_wait(full_bar(st.idx), st.phase))) {
          fail(kErrRelayFull);
          ok = false;
          break;
        }
        if (ptx::elect_one()) ptx::mbar_arrive_remote(pfull_bar(st.idx), 0);
        __syncwarp();
        st.advance(p.n_stages);
      }
      if (++b == p.nblk) b = 0;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform loop; one elected lane issues the 4 tcgen05.mma of a K block and the
    // commits.  Descriptors differ only in the 14-bit start-address field.
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(kBlockRows * CG, kUnitCols);
    const uint64_t desc_hi = ptx::smem_desc(0, p.lbo, p.sbo);
    const uint32_t q_base = ptx::smem_u32(q_sm) >> 4, st_base = ptx::smem_u32(stage_sm) >> 4;
    Ring st, acc;
    int seg = 0;
    int b = (int)(u0 % p.nblk);
    bool ok = true;
    for (int64_t u = u0; u < u1 && ok; ++u) {
      if (!__all_sync(0xffffffffu, ptx::mbar_wait(tempty_bar(acc.idx), acc.phase ^ 1u))) {
        fail(kErrMmaTmemEmpty);
        break;
      }
      if (res && (u == u0 || b == 0)) {
        if (!__all_sync(0xffffffffu, ptx::mbar_wait(qfull_bar, (uint32_t)(seg & 1)))) {
          fail(kErrMmaQFull);
          break;
        }
        if (CG == 2 && !__all_sync(0xffffffffu, ptx::mbar_wait(pqfull_bar, (uint32_t)(seg & 1)))) {
          fail(kErrMmaPeerQFull);
          break;
        }
        ++seg;
      }
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc.idx * kUnitCols);
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (!__all_sync(0xffffffffu, ptx::mbar_wait(full_bar(st.idx), st.phase))) {
          fail(kErrMmaFull);
          ok = false;
          break;
        }
        if (CG == 2 && !__all_sync(0xffffffffu, ptx::mbar_wait(pfull_bar(st.idx), st.phase))) {
          fail(kErrMmaPeerFull);
          ok = false;
          break;
        }
        ptx::tc_fence_after();
        const uint32_t e_lo = st_base + (uint32_t)(st.idx * (kStageBytes >> 4));
        const uint32_t q_lo = (QRES || q_kb(kb) < n_res) ? q_base + (uint32_t)(q_kb(kb) * (kKBlockBytes >> 4))
                                                         : e_lo + (uint32_t)(kSlabsPerStage * (kKBlockBytes >> 4));
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < kKBlockElems / 16; ++k) {  // +32 bytes per K step inside the swizzled row
            const uint64_t da = desc_hi | (uint64_t)((q_lo + 2u * k) & 0x3fffu);
            const uint64_t db = desc_hi | (uint64_t)((e_lo + 2u * k) & 0x3fffu);
            if (CG == 2) ptx::umma_bf16_2cta(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else ptx::umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // smem stage reusable (in both CTAs) once these MMAs retire
          if (CG == 2) ptx::umma_commit_2cta(empty_bar(st.idx), 3);
          else ptx::umma_commit(empty_bar(st.idx));
        }
        __syncwarp();
        st.advance(p.n_stages);
      }
      if (!ok) break;
      const bool seg_end = (u + 1 == u1) || (b + 1 == p.nblk);
      if (ptx::elect_one()) {
        if (CG == 2) {
          ptx::umma_commit_2cta(tfull_bar(acc.idx), 3);              // accumulators complete -> both epilogues
          if (res && seg_end) ptx::umma_commit_2cta(qempty_bar, 3);  // query tiles no longer read
        } else {
          ptx::umma_commit(tfull_bar(acc.idx));
          if (res && seg_end) ptx::umma_commit(qempty_bar);
        }
      }
      __syncwarp();
      acc.advance(kAccStages);
      if (++b == p.nblk) b = 0;
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================== epilogue: metric + running top-k =====================
    const int ew = warp - kFirstEpiWarp;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may read
    const int ch = ew >> 2;                  // which half of the 128 columns
    const int lane_q = quarter * 32 + lane;  // query within the tile
    float* my_side = side_ring + ew * (2 * kColsPerWarp);
    // this lane's 8 scores of a column group, for the lane-local insertions of crowded groups (below)
    float* my_stash = side_ring + kEpiWarps * 2 * kColsPerWarp + (ew * 32 + lane) * kStashStride;
    Ring acc;
    int slot = 0;
    int qt = (int)(u0 / p.nblk) * CG + rank, b = (int)(u0 % p.nblk);
    int cb = rot_block(u0 / p.nblk, b, p.nblk, p.total_units, G, p.rotate);  // corpus row-block pair of unit b
    typename SelectorFor<KSEL>::type top;
    // append buffers stay in L2 (evict-last) as long as all of them together leave most of it to
    // the corpus stream; at thousands of queries they would be 100 MB and push the row blocks the
    // scheduling groups share out of L2 (ncu, B=4096: hit rate 80 -> 49 %, HBM reads x3)
    // (KSEL 1 is the same selector with plain stores: no policy operand in the append blocks; the
    // compactions, which are rare, keep hinted stores with the normal policy)
    const uint64_t keep_policy = KSEL == 0 ? ptx::policy_evict_last() : ptx::policy_evict_normal();
#ifdef LK_EPI_PROF
    long long prof_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // tfull wait, tmem ld, fast math, select, unit total, hit turns, turns, units
    const long long prof_begin = clock64();
#endif
    float thr = INFINITY;
    float q_sd = 0.f;
    int* cnt_out = nullptr;
    bool seg_start = true;
    if (u0 < u1) {
      const float* sp = p.side + (int64_t)cb * kUnitCols + ch * kColsPerWarp;
#pragma unroll
      for (int j = 0; j < kColsPerWarp / 32; ++j) my_side[lane + 32 * j] = sp[lane + 32 * j];
      __syncwarp();
    }
    for (int64_t u = u0; u < u1; ++u) {
      if (seg_start) {
        const int64_t q = (int64_t)qt * kBlockRows + lane_q;
        const bool q_ok = q < p.n_queries;
        q_sd = q_ok ? p.q_side[q] : 0.f;
        const int64_t c_first = cta_of_unit((int64_t)(qt / CG) * p.nblk, p.total_units, G);
        const int list = (int)(c - c_first) * kColSplit + ch;
        const int64_t o = q_ok ? (q * p.n_lists + list) * p.ksel : 0;
        // seeded threshold: the k-th best score over a sample of the corpus (found by an earlier
        // launch over a prefix of the rows) bounds the final k-th best from below; taken a hair
        // lower so that the sample's own k-th row passes again, in this kernel's score space
        float floor = -INFINITY;
        if (p.seed != nullptr && q_ok) {
          float sd = p.seed[q * p.k + (p.k - 1)];
          if (sd > -INFINITY) {
            if (METRIC == LK_COSINE) sd = sd / q_sd;  // q_sd = 1/|q| > 0
            const float mag = METRIC == LK_COSINE ? fabsf(sd) : fabsf(sd) + q_sd;
            floor = sd - fmaxf(4e-6f * mag, 1e-37f);
          }
        }
        top.begin(p.part_scores + o, p.part_idx + o, q_ok, p.k, floor);
        cnt_out = (q_ok && p.part_cnt != nullptr) ? p.part_cnt + (q * p.n_lists + list) : nullptr;
        thr = top.threshold();
        seg_start = false;
      }
      // prefetch the next unit's side values while this unit's MMAs finish
      int nb = b + 1, nqt = qt, ncb = cb + 1 == p.nblk ? 0 : cb + 1;
      if (nb == p.nblk) {
        nb = 0;
        nqt += CG;
        ncb = rot_block(nqt / CG, 0, p.nblk, p.total_units, G, p.rotate);
      }
      float ns[kColsPerWarp / 32];
#pragma unroll
      for (int j = 0; j < kColsPerWarp / 32; ++j) ns[j] = 0.f;
      if (u + 1 < u1) {
        const float* sp = p.side + (int64_t)ncb * kUnitCols + ch * kColsPerWarp;
#pragma unroll
        for (int j = 0; j < kColsPerWarp / 32; ++j) ns[j] = sp[lane + 32 * j];
      }
#ifdef LK_EPI_PROF
      const long long tp0 = clock64();
#endif
      if (!__all_sync(0xffffffffu, ptx::mbar_wait(tfull_bar(acc.idx), acc.phase))) {
        fail(kErrEpiTmemFull);
        break;
      }
      ptx::tc_fence_after();
#ifdef LK_EPI_PROF
      const long long tp1 = clock64();
      prof_t[0] += tp1 - tp0;
#endif
      const float* sd = my_side + slot * kColsPerWarp;
      const int32_t row0 = cb * kUnitCols + ch * kColsPerWarp;
      // Fast path (branch-free, a few hundred instructions in total so it stays in the
      // instruction cache): score every column and keep the running max.  Only when some
      // lane's max beats its k-th best does the warp take the slow path, which holds the
      // ONE copy of the insertion network and re-reads the candidate columns from TMEM.
      auto score = [&](float dot, float e_sd) -> float {
        if (METRIC == LK_COSINE) return dot * e_sd;            // x 1/|e|; x 1/|q| at flush
        return fmaf(2.0f, dot, -(q_sd + e_sd));                // -(|q|^2 + |e|^2 - 2 q.e)
      };
      // running max per group of 8 columns (NaN never wins)
      auto group_max = [&](const uint32_t (&r)[32], const float* sdc, float (&mg)[4]) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          mg[g] = -INFINITY;
#pragma unroll
          for (int j = 8 * g; j < 8 * g + 8; ++j) mg[g] = fmaxf(mg[g], score(__uint_as_float(r[j]), sdc[j]));
        }
      };
      // everything after the fast path of one chunk of 32 columns
      auto select = [&](int chunk, uint32_t taddr, const uint32_t (&r)[32], const float* sdc, const float (&mg)[4]) {
        const float m = fmaxf(fmaxf(mg[0], mg[1]), fmaxf(mg[2], mg[3]));
        if constexpr (SelectorFor<KSEL>::type::kAppend) {
          // k > 10: every lane appends its own hits (predicated, no divergence, no TMEM re-read),
          // only in the 8-column groups where some lane has one; then the warp compacts the
          // buffers that are about to run out of room for a chunk.
          if (__any_sync(0xffffffffu, m > thr)) {
            top.note_max(m);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (__any_sync(0xffffffffu, mg[g] > thr)) {
#pragma unroll
                for (int j = 8 * g; j < 8 * g + 8; ++j) {  // pure predication: a vote + branch per column, or a
                  // uniform branch between hinted and plain stores, both measured ~25 % slower
                  const float sc = score(__uint_as_float(r[j]), sdc[j]);
                  if (sc > thr) top.template append<KSEL == 0>(sc, row0 + chunk * 32 + j, keep_policy);  // thr fixed between compactions
                }
              }
            }
            top.compact(kBufCap - 32, lane, keep_policy);  // room for one more chunk, or shrink now
            thr = top.threshold();
          }
        } else if (__any_sync(0xffffffffu, m > thr)) {
          // k <= 10: only the 8-column groups in which some lane has a hit are looked at again
          unsigned pend = 0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (!__any_sync(0xffffffffu, mg[g] > thr)) continue;
#pragma unroll
            for (int j = 8 * g; j < 8 * g + 8; ++j)
              pend |= score(__uint_as_float(r[j]), sdc[j]) > thr ? (1u << j) : 0u;
          }
          unsigned umask = __reduce_or_sync(0xffffffffu, pend);
          if (__popc(umask) >= 4) {
            // Crowded chunk (cold thresholds: the first rows of a segment, small corpora): serving the
            // candidate columns one per turn makes the WHOLE warp run the insertion network once per column
            // any lane hits.  Instead every lane parks the 8 scores of a group in shared memory and inserts
            // its own hits, in row order: the warp pays for the lane with the most hits, not for the number
            // of distinct columns.
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (!((umask >> (8 * g)) & 0xffu)) continue;  // warp-uniform
#pragma unroll
              for (int j = 0; j < 8; ++j) my_stash[j] = score(__uint_as_float(r[8 * g + j]), sdc[8 * g + j]);
              unsigned mine = (pend >> (8 * g)) & 0xffu;
              while (mine) {
                const int jj = __ffs(mine) - 1;
                mine &= mine - 1;
                const float sc = my_stash[jj];
                if (sc > thr) {  // rows in ascending order: strict '>' keeps the lower index
                  top.insert(sc, row0 + chunk * 32 + 8 * g + jj);
                  thr = top.threshold();
                }
              }
              __syncwarp();
            }
          } else {
            while (umask) {  // warp-uniform: one candidate column per turn
              const int j = __ffs(umask) - 1;
              umask &= umask - 1;
              const float dot = __uint_as_float(ptx::tmem_ld1(taddr + (uint32_t)j));
              ptx::tmem_wait_ld();
              const float sc = score(dot, sdc[j]);
              if (sc > thr) {  // rows arrive in ascending order: strict '>' keeps the lower index
                top.insert(sc, row0 + chunk * 32 + j);
                thr = top.threshold();
              }
            }
          }
        }
      };
      // Two chunks per turn: both tcgen05.ld are in flight together and the two independent
      // score / max chains interleave, which hides most of the latency two warps per scheduler
      // cannot; the selection steps then run in row order.
#if LK_EPI_WARPS > 8
#pragma unroll 1
      for (int chunk = 0; chunk < kColsPerWarp / 32; ++chunk) {  // one chunk per turn: 96 registers per thread
        uint32_t r0[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                               (uint32_t)(acc.idx * kUnitCols + ch * kColsPerWarp + chunk * 32);
        ptx::tmem_ld32(taddr, r0);
        ptx::tmem_wait_ld();
        const float* sdc = sd + chunk * 32;
        float mg0[4];
        group_max(r0, sdc, mg0);
        select(chunk, taddr, r0, sdc, mg0);
      }
#else
#pragma unroll 1
      for (int chunk = 0; chunk < kColsPerWarp / 32; chunk += 2) {
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                               (uint32_t)(acc.idx * kUnitCols + ch * kColsPerWarp + chunk * 32);
#ifdef LK_EPI_PROF
        const long long tq0 = clock64();
#endif
        ptx::tmem_ld32(taddr, r0);
        ptx::tmem_ld32(taddr + 32u, r1);
        ptx::tmem_wait_ld();
#ifdef LK_EPI_PROF
        const long long tq1 = clock64();
        prof_t[1] += tq1 - tq0;
#endif
        if (p.debug_tile != nullptr && u == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ch == 0) {
              p.debug_tile[lane_q * kBlockRows + chunk * 32 + j] = __uint_as_float(r0[j]);
              p.debug_tile[lane_q * kBlockRows + chunk * 32 + 32 + j] = __uint_as_float(r1[j]);
            }
        }
        const float* sdc = sd + chunk * 32;
        float mg0[4], mg1[4];
        group_max(r0, sdc, mg0);
        group_max(r1, sdc + 32, mg1);
#ifdef LK_EPI_PROF
        const bool any_hit = __any_sync(0xffffffffu, fmaxf(fmaxf(mg0[0], mg0[1]), fmaxf(mg0[2], mg0[3])) > thr ||
                                                         fmaxf(fmaxf(mg1[0], mg1[1]), fmaxf(mg1[2], mg1[3])) > thr);
        const long long tq2 = clock64();
        prof_t[2] += tq2 - tq1;
#endif
        select(chunk, taddr, r0, sdc, mg0);
        select(chunk + 1, taddr + 32u, r1, sdc + 32, mg1);
#ifdef LK_EPI_PROF
        prof_t[3] += clock64() - tq2;
        prof_t[5] += any_hit ? 1 : 0;
        prof_t[6] += 1;
#endif
      }
#ifdef LK_EPI_PROF
      prof_t[4] += clock64() - tp1;
      prof_t[7] += 1;
#endif
#endif
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // the accumulator stage goes back to the (leader's) MMA warp
        if (CG == 2 && rank != 0) ptx::mbar_arrive_remote(tempty_bar(acc.idx), 0);
        else ptx::mbar_arrive(tempty_bar(acc.idx));
      }
      acc.advance(kAccStages);
      if constexpr (SelectorFor<KSEL>::type::kAppend) {
        // routine compactions run here, after the accumulator went back to the MMA warp, a few
        // lanes per unit, so they overlap the next unit instead of stalling the pipeline
        top.compact(kBufCap - 64, lane, keep_policy, 4);
        thr = top.threshold();
      }
      slot ^= 1;
#pragma unroll
      for (int j = 0; j < kColsPerWarp / 32; ++j) my_side[slot * kColsPerWarp + lane + 32 * j] = ns[j];
      __syncwarp();

      const bool seg_end = (u + 1 == u1) || (b + 1 == p.nblk);
      if (seg_end) {
        top.finish(q_sd, METRIC == LK_COSINE, p.k, lane);  // cosine: x 1/|q| once per kept entry
        if constexpr (SelectorFor<KSEL>::type::kAppend)
          if (cnt_out != nullptr) *cnt_out = top.cnt;
        // The partial-list arrays are not cleared before the launch: the last group of a query
        // tile marks the list slots no group owns (n_lists is the maximum over query tiles).
        {
          const int64_t qg = qt / CG;
          const int64_t c_first = cta_of_unit(qg * p.nblk, p.total_units, G);
          const int64_t c_last = cta_of_unit((qg + 1) * p.nblk - 1, p.total_units, G);
          const int64_t q = (int64_t)qt * kBlockRows + lane_q;
          if (c == c_last && q < p.n_queries) {
            for (int l = (int)(c - c_first + 1) * kColSplit + ch; l < p.n_lists; l += kColSplit) {
              if constexpr (SelectorFor<KSEL>::type::kAppend) {
                p.part_cnt[q * p.n_lists + l] = 0;
              } else {
                int32_t* e = p.part_idx + (q * p.n_lists + l) * p.ksel;
                for (int j = 0; j < p.ksel; ++j) e[j] = -1;
              }
            }
          }
        }
        seg_start = true;
      }
      b = nb;
      qt = nqt;
      cb = ncb;
    }
#ifdef LK_EPI_PROF
    if (p.debug_tile != nullptr && ew == 0 && lane == 0) {
      long long* o = reinterpret_cast<long long*>(p.debug_tile) + (int64_t)blockIdx.x * 10;
      for (int i = 0; i < 8; ++i) o[i] = prof_t[i];
      o[8] = clock64() - prof_begin;
    }
#endif
  }

  // ===================== teardown =====================
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync_all();  // nobody touches the peer's barriers or TMEM any more
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if (CG == 2) ptx::tmem_dealloc2(tmem_base, kTmemCols);
    else ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

inline int n_kblocks(const TileGeom& g) { return g.kblocks; }  // slabs per row block in memory (split mode: both planes)
inline bool q_resident(const TileGeom& g) { return n_kblocks(g) <= 8; }
inline int stage_bytes_for(const TileGeom& g, int cg) {
  return (kUnitBlocks / cg + (q_resident(g) ? 0 : 1)) * kKBlockBytes;
}
// Query K blocks kept in shared memory for a whole segment.  All of them when the tile fits beside the pipeline
// (<= 8 K blocks: 512 bf16 dimensions).  Otherwise the tile is streamed with the corpus, which doubles the L2 -> SM
// traffic of a CTA pair's unit; at thousands of queries that traffic (9.8 TB/s over the 148 SMs, ncu, 768-d) and not
// the tensor pipe bounds the kernel, so CTA pairs keep the first 5 K blocks (80 KB) resident and run 4 stages of
// 32 KB instead of 6 (measured back to back on one box, tools/gpu_r2_qres.sh): 2M x 768 top-100 at 4096 queries
// 10.0 -> 9.36 ms (4 blocks: 9.5, 6 blocks = 3 stages: 9.9), fp32 planes 2M x 384 11.1 -> 10.75 ms; 512 queries lose
// 1.5 % (1.38 -> 1.40 ms).  Single CTAs (<= 128 queries, HBM-bound) keep the deeper pipeline: 10M x 768 at 64 queries
// 2.17 -> 2.21 ms with 4 resident blocks.
constexpr int kPartialResidentKb = 5;
inline int n_res_for(const TileGeom& g, int cg) {
  if (q_resident(g)) return n_kblocks(g);
  int v = cg == 2 ? kPartialResidentKb : 0;
  if (const char* e = getenv("LK_QRES_KB")) v = atoi(e);  // bring-up override
  if (v < 0) v = 0;
  const int stage = (kUnitBlocks / cg + 1) * kKBlockBytes;
  while (v > 0 && (kSmemBudget - kHeaderBytes - kAlignSlack - v * kKBlockBytes) / stage < 3) --v;  // >= 3 stages
  return v < n_kblocks(g) ? v : n_kblocks(g) - 1;
}
inline int n_stages_for(const TileGeom& g, int cg) {
  const int q_bytes = n_res_for(g, cg) * kKBlockBytes;
  int s = (kSmemBudget - kHeaderBytes - kAlignSlack - q_bytes) / stage_bytes_for(g, cg);
  return s > kMaxStages ? kMaxStages : s;
}
// sorted register list / append buffer.  (Measured: the append buffer is slower than the register
// list at k = 10 for every batch size, and 16 epilogue warps of 64 columns are slower than 8 of
// 128 - spills at 96 registers, twice the lists; profiles/r01_microbench.log.)
inline int ksel_for(int k) {
  if (const char* e = getenv("LK_KSEL_BUF"))  // bring-up override: the append-buffer selector for small k as well
    if (atoi(e)) return kBufCap;
  return k <= 10 ? 10 : kBufCap;
}

// CTA pairs (cta_group::2) once there are at least two query tiles to pair up; single CTAs for
// the bandwidth-bound case of one query tile.
inline int cta_group_for(int64_t n_queries, int sm_count) {
  if (const char* e = getenv("LK_CG")) {  // bring-up override
    const int v = atoi(e);
    if (v == 1 || (v == 2 && sm_count % 2 == 0)) return v;
  }
  return n_queries > kBlockRows && sm_count % 2 == 0 ? 2 : 1;
}

struct Schedule {
  int cg;            // CTAs per scheduling group
  int64_t nblk;      // units (256 corpus rows) per query-tile group
  int64_t nqg;       // query-tile groups (cg tiles each)
  int64_t total;     // units over all groups
  int64_t groups;    // scheduling groups launched
};
inline Schedule schedule_for(int64_t n_rows, int64_t n_queries, int sm_count) {
  Schedule s;
  s.cg = cta_group_for(n_queries, sm_count);
  s.nblk = (n_rows + kUnitCols - 1) / kUnitCols;
  const int64_t nqt = (n_queries + kBlockRows - 1) / kBlockRows;
  s.nqg = (nqt + s.cg - 1) / s.cg;
  s.total = s.nblk * s.nqg;
  const int64_t avail = sm_count / s.cg;
  s.groups = s.total < avail ? s.total : avail;
  return s;
}

}  // namespace

int umma_supported(const TileGeom& g, int k) {
  return g.elem_bytes == 2 && g.kblocks >= 1 && k >= 1 && k <= kMaxK && n_stages_for(g, 1) >= 2;
}

int umma_plan(const SearchArgs& a, int sm_count, int* n_lists, int* ksel) {
  const Schedule sc = schedule_for(a.n_rows, a.n_queries, sm_count);
  int64_t maxc = 1;
  for (int64_t qg = 0; qg < sc.nqg; ++qg) {
    const int64_t cf = cta_of_unit(qg * sc.nblk, sc.total, sc.groups),
                  cl = cta_of_unit((qg + 1) * sc.nblk - 1, sc.total, sc.groups);
    if (cl - cf + 1 > maxc) maxc = cl - cf + 1;
  }
  *n_lists = (int)maxc * kColSplit;
  *ksel = ksel_for(a.k);
  return LK_OK;
}

// Threshold seeding (k > 10, 8..1024 queries): every (group, column half) list starts cold, and
// with few query tiles each list sees only a small share of the rows, so most of what it
// appends is far from the global top-k.  Searching a short prefix of the corpus first (one unit
// = 256 rows per group: every list takes at most 128 entries and never compacts) gives every
// list a threshold that only k * rows / prefix_rows rows of the whole corpus beat.
int64_t umma_seed_rows(const SearchArgs& a, int sm_count) {
  if (ksel_for(a.k) != kBufCap || a.n_queries < 8 || a.n_queries > 1024) return 0;
  if (const char* e = getenv("LK_SEED"))  // bring-up override
    if (!atoi(e)) return 0;
  const Schedule sc = schedule_for(a.n_rows, a.n_queries, sm_count);
  const int64_t s_units = (sm_count / sc.cg + sc.nqg - 1) / sc.nqg;  // one unit per group
  if (sc.nblk < 16 * s_units) return 0;
  return s_units * kUnitCols;
}

int launch_search_umma(const SearchArgs& a, int sm_count, cudaStream_t st) {
  if (a.n_queries <= 0 || a.n_rows <= 0) return LK_OK;
  if (!umma_supported(a.g, a.k)) {
    set_error("tcgen05 search: unsupported shape (dim_pad=%d, k=%d)", a.g.dim_pad, a.k);
    return LK_ERR_UNSUPPORTED;
  }
  const Schedule sc = schedule_for(a.n_rows, a.n_queries, sm_count);
  UmmaParams p;
  p.tiles = static_cast<const unsigned char*>(a.tiles);
  p.side = a.side;
  p.q_tiles = static_cast<const unsigned char*>(a.q_tiles);
  p.q_side = a.q_side;
  p.n_queries = a.n_queries;
  p.total_units = sc.total;
  p.nblk = (int)sc.nblk;
  p.split_n = a.split_n;
  p.nkb_q = n_kblocks(a.g);
  p.nkb = a.split_n ? 3 * a.split_n : p.nkb_q;
  if (a.split_n && p.nkb_q != 2 * a.split_n) {
    set_error("tcgen05 search: split operands need 2 planes of %d K blocks, the geometry has %d", a.split_n, p.nkb_q);
    return LK_ERR_INVALID;
  }
  p.nkb_res = n_res_for(a.g, sc.cg);
  p.n_stages = n_stages_for(a.g, sc.cg);
  p.n_lists = a.n_lists;
  p.ksel = a.ksel;
  p.k = a.k;
  p.block_bytes = a.g.block_bytes();
  p.part_scores = a.part_scores;
  p.part_idx = a.part_idx;
  p.part_cnt = a.part_cnt;
  p.err_flag = a.err_flag;
  p.seed = a.seed;
  p.debug_tile = a.debug_tile;
  p.dbg = getenv("LK_DBG") ? atoi(getenv("LK_DBG")) : 0;
  p.rotate = (p.dbg & 4) ? 0 : 1;
  const bool keep_lists = (p.dbg & 16) || (int64_t)a.n_queries * a.n_lists * a.ksel * 8 <= (48ll << 20);
  p.lbo = kLbo;
  p.sbo = kSbo;
  if (const char* e = getenv("LK_UMMA_LBO")) p.lbo = (uint32_t)atoi(e);  // bring-up overrides
  if (const char* e = getenv("LK_UMMA_SBO")) p.sbo = (uint32_t)atoi(e);
  if (const char* e = getenv("LK_UMMA_STAGES")) {
    const int v = atoi(e);
    if (v >= 2 && v <= p.n_stages) p.n_stages = v;
  }
  const bool qres = q_resident(a.g);
  const size_t smem = (size_t)kHeaderBytes + kAlignSlack + (size_t)p.nkb_res * kKBlockBytes +
                      (size_t)p.n_stages * stage_bytes_for(a.g, sc.cg);
  const int ksel = ksel_for(a.k);
  if (ksel != a.ksel) {
    set_error("tcgen05 search: plan mismatch (ksel %d vs %d)", a.ksel, ksel);
    return LK_ERR_INVALID;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(sc.groups * sc.cg));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)sc.cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define LK_UMMA(KS, MET, QR, CGV)                                                                    \
  do {                                                                                               \
    LK_CUDA(cudaFuncSetAttribute(umma_search_kernel<KS, MET, QR, CGV>,                               \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    LK_CUDA(cudaLaunchKernelEx(&cfg, umma_search_kernel<KS, MET, QR, CGV>, p));                      \
  } while (0)
#define LK_UMMA_C(KS, MET, QR)                                             \
  do {                                                                     \
    if (sc.cg == 2) LK_UMMA(KS, MET, QR, 2);                               \
    else LK_UMMA(KS, MET, QR, 1);                                          \
  } while (0)
#define LK_UMMA_M(KS, QR)                                                  \
  do {                                                                     \
    if (a.metric == LK_COSINE) LK_UMMA_C(KS, LK_COSINE, QR);               \
    else LK_UMMA_C(KS, LK_EUCLIDEAN, QR);                                  \
  } while (0)
  if (ksel == 10) {
    if (qres) LK_UMMA_M(10, true); else LK_UMMA_M(10, false);
  } else if (keep_lists) {
    if (qres) LK_UMMA_M(0, true); else LK_UMMA_M(0, false);  // KSEL 0 = BufSelector, lists kept in L2
  } else {
    if (qres) LK_UMMA_M(1, true); else LK_UMMA_M(1, false);  // KSEL 1 = BufSelector, plain appends
  }
#undef LK_UMMA_M
#undef LK_UMMA_C
#undef LK_UMMA
  LK_CHECK_LAUNCH("umma_search_kernel");
  return LK_OK;
}

}  // namespace lk
