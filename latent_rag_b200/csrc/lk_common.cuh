// Internal declarations shared by the liblatentknn translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "latentknn.h"

namespace lk {

// ---------------------------------------------------------------------------------
// Data layout in HBM (DESIGN.md "Data layout")
//
// The corpus is stored as ROW BLOCKS of 128 rows, each split into K BLOCKS of 128 bytes
// per row (64 bf16 or 32 fp32).  One (row block, K block) SLAB is 128 rows x 128 B = 16 KB
// and is byte for byte the tcgen05 "K-major, SWIZZLE_128B" shared-memory operand image:
//
//     slab[r][p][16 bytes]   r = 0..127, p = physical 16-byte chunk = c ^ (r & 7)
//                            c = logical chunk 0..7 (8 bf16 / 4 fp32 each)
//
// i.e. rows are 128 B apart, 8-row groups 1024 B apart (SBO), and the 16-byte chunks of a
// row are XOR-swizzled with the row index so that tcgen05.mma reads operands at full
// shared-memory bandwidth.  A 16 KB cp.async.bulk brings a slab into a 1024-aligned stage
// with no tensor map.  Slabs are ordered [row block][K block].
// ---------------------------------------------------------------------------------
constexpr int kBlockRows = 128;       // rows per row block (= UMMA M / N granule)
constexpr int kChunkBytes = 16;       // one K chunk of one row
constexpr int kRowBytes = 128;        // bytes of one row inside a K block (swizzle span)
constexpr int kSlabBytes = kBlockRows * kRowBytes;  // 16384
constexpr int kMaxK = 128;            // largest top-k of one fused search (entries per selector list)
constexpr int kDeepMaxK = LK_MAX_K;   // largest top-k of a slab search (lk_deep.cu)
constexpr int kDeepMaxLists = 96;     // slab lists folded in by one deep merge launch
constexpr int kDeepMaxCand = 16384;   // entries one deep merge sorts in shared memory (192 KB)
constexpr int kSimtQG = 4;            // queries per CTA pass in the SIMT kernel

// byte offset of logical chunk c (0..7) of row r inside a slab
__host__ __device__ inline int slab_chunk_offset(int r, int c) {
  return r * kRowBytes + ((c ^ (r & 7)) << 4);
}

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// (score, index) total order used everywhere a choice between candidates is made:
// higher score first, lower index on equal scores.  Makes results independent of how
// the corpus is partitioned over CTAs / GPUs.
__device__ __forceinline__ bool better(float sa, int64_t ia, float sb, int64_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

// ---------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);

#define LK_CUDA(expr)                                                        \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) return ::lk::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define LK_CHECK_LAUNCH(name)                                                \
  do {                                                                       \
    ::lk::count_launch();                                                    \
    cudaError_t _e = cudaGetLastError();                                     \
    if (_e != cudaSuccess) return ::lk::cuda_fail(_e, name, __FILE__, __LINE__); \
  } while (0)

// ---------------------------------------------------------------------------------
// kernel launchers (defined in the .cu files, called from lk_api.cu)
// ---------------------------------------------------------------------------------
struct TileGeom {
  int dim;          // logical dimension
  int dim_pad;      // padded to whole K blocks: a multiple of 64 (bf16) / 32 (fp32) elements
  int elem_bytes;   // 2 (bf16) or 4 (fp32)
  int kblocks;      // dim_pad * elem_bytes / 128
  __host__ __device__ int chunks() const { return kblocks * 8; }
  __host__ __device__ int64_t block_bytes() const { return (int64_t)kblocks * kSlabBytes; }
};

// rows [n, dim] (fp32 or bf16 row-major, device) -> tiled storage starting at row
// `row0`, plus one fp32 side value per row.  side_mode: 0 = 1/max(|row|,1e-12) of the
// stored values, 1 = |row|^2 of the stored values.  prenorm: divide the row by
// max(|row|, 1e-12) before storing (fp32 cosine storage, like F.normalize).
int launch_tile_rows(const void* rows, int rows_dtype, int64_t n, const TileGeom& g, void* tiles,
                     float* side, int64_t row0, int side_mode, int prenorm, cudaStream_t st);

// fp32 tiles (geometry g32: 32 elements per K block) of row blocks [blk0, blk0 + n_blocks) -> split-bf16 planes
// (geometry gp: per row block [hi plane | lo plane], gp.kblocks / 2 K blocks of 64 each; x = hi + lo to 2^-17)
int launch_planes_from_tiles(const void* tiles32, const TileGeom& g32, int64_t blk0, int64_t n_blocks,
                             const TileGeom& gp, void* planes, cudaStream_t st);

// out[n, dim] fp32 = rows[n, dim] (fp32 or bf16) * L[dim, dim] (fp64, row-major), fp64 accumulate
int launch_whiten(const void* rows, int rows_dtype, int64_t n, int dim, const double* L, float* out,
                  cudaStream_t st);

struct SearchArgs {
  const void* tiles;       // corpus tiles
  const float* side;       // per-row side value (NaN for rows past the end)
  int64_t n_rows;
  TileGeom g;
  const void* q_tiles;     // query tiles (same layout, 128 queries per block)
  const float* q_side;
  int64_t n_queries;
  int metric;              // LK_COSINE or LK_EUCLIDEAN (mahalanobis is whitened L2)
  int split_n;             // tcgen05 kernel on fp32 storage: tiles / q_tiles are split-bf16 PLANES, [hi | lo] slabs
                           // of split_n K blocks each per row block (g.kblocks = 2 * split_n); 0 = plain bf16 tiles
  int k;                   // requested k
  int ksel;                // entries per partial list (>= k)
  int n_lists;             // partial lists per query (stride of the partial arrays)
  float* part_scores;      // [n_queries, n_lists, ksel]
  int32_t* part_idx;       // [n_queries, n_lists, ksel]
  int* part_cnt;           // [n_queries, n_lists] fill counts (append-buffer selector), or null
  int* err_flag;           // device int, set non-zero by a kernel that timed out
  const float* seed;       // tcgen05 kernel, k > 32: [n_queries, k] scores of a sample search, or null
  float* debug_tile;       // optional [128 x 128] dump of unit 0's raw accumulator (bring-up aid)
  // single-launch mode of the SIMT kernel (small batches on small corpora): raw queries in, final
  // results out -- query rounding / norms, search and the merge of the slices in one kernel
  const void* q_raw;       // [n_queries, dim] row-major device memory (q_tiles unused), or null
  int q_raw_dtype;         // LK_F32 / LK_BF16
  float* out_scores;       // [n_queries, k]
  int64_t* out_idx;        // [n_queries, k]   row position + idx_base
  int64_t idx_base;
  int* ticket;             // one zero-initialised counter per query group; the last slice to finish merges
};

int simt_plan(const SearchArgs& a, int sm_count, int* n_lists, int* ksel);
int launch_search_simt(const SearchArgs& a, int sm_count, cudaStream_t st);
// whether the single-launch mode serves this call (bf16 storage, <= 4 queries, rows x queries <= 40k)
int simt_fused_supported(const TileGeom& g, int64_t n_rows, int64_t n_queries, int k);
int umma_supported(const TileGeom& g, int k);
int umma_plan(const SearchArgs& a, int sm_count, int* n_lists, int* ksel);
// rows of the corpus prefix whose top-k seeds the selection thresholds (0 = do not seed)
int64_t umma_seed_rows(const SearchArgs& a, int sm_count);
int launch_search_umma(const SearchArgs& a, int sm_count, cudaStream_t st);

// merge [b, n_lists, list_len] candidates -> [b, k]; idx type int32 (+base) or int64
// pc: optional [b, n_lists] fill counts of the lists (entries past the count are not read)
// tmp_s / tmp_i (optional, tmp_entries each): scratch for the two-level merge of a few queries with very many candidates
int launch_merge_i32(const float* ps, const int32_t* pi, const int* pc, int64_t b, int n_lists, int list_len,
                     int list_stride, int k, int64_t idx_base, float* out_s, int64_t* out_i, cudaStream_t st,
                     float* tmp_s = nullptr, int64_t* tmp_i = nullptr, int64_t tmp_entries = 0);
int launch_merge_i64(const float* ps, const int64_t* pi, int64_t b, int n_lists, int list_len, int k,
                     float* out_s, int64_t* out_i, cudaStream_t st);

// Slab search for k > kMaxK (lk_deep.cu).  List l of query q starts at l * list_stride +
// q * query_stride of the candidate arrays (int64 ids, < 0 or NaN score = empty).  Entries of the
// running result whose id falls in [lo[r], hi[r]) are dropped before the merge: the incoming
// lists supply those rows again (a split slab is searched a second time, half by half).
struct DeepLists {
  int n_lists, len, n_ranges;
  int64_t list_stride, query_stride;
  int64_t lo[kDeepMaxLists], hi[kDeepMaxLists];
};
// lists of `list_len` entries one launch can take next to a running result of k entries
int deep_merge_capacity(int k, int have_res, int list_len);
// res (b x k, in place; read only if have_res) <- best k of res + lists.  list_last: optional
// [n_lists, b]; receives each list's last score when its slab (hi - lo rows) holds more rows
// than the list, NaN otherwise.
int launch_deep_merge(const float* cs, const int64_t* ci, const DeepLists& L, int64_t b, int k, int have_res,
                      float* res_s, int64_t* res_i, float* list_last, cudaStream_t st);
// flags[l] = 1 if for some query list l's last score reaches the query's k-th best
int launch_deep_saturated(const float* list_last, int n_lists, int64_t b, const float* res_s, int k, int* flags,
                          cudaStream_t st);

int launch_ae_encode(const float* x, int64_t m, int d_in, int d_hidden, int d_latent, const float* w0t,
                     const float* b0, const float* w1t, const float* b1, int l2norm, float* z,
                     cudaStream_t st);

// tensor-core (split-bf16) autoencoder path, lk_ae_umma.cu
int ae_umma_supported(int d_in, int d_hidden, int d_latent);
size_t ae_umma_x_slab_bytes(int64_t m, int d_in);
// n_planes: 2 = split-bf16 operands (x = hi + lo), 1 = plain bf16 (hi only)
int launch_ae_split_rows(const float* x, int64_t m, int d_in, int n_planes, unsigned char* slabs, cudaStream_t st);
int launch_ae_umma(const unsigned char* x_slabs, int64_t m, int d_in, int d_hidden, int d_latent,
                   const unsigned char* w0_slabs, const unsigned char* w1_slabs, const float* b0, const float* b1,
                   int l2norm, int n_planes, float* z, int* err_flag, int sm_count, cudaStream_t st);

// bf16-operand encoder on CTA pairs, fp32 rows converted in the kernel (lk_ae_pair.cu); weight slabs as above
int ae_pair_supported(int d_in, int d_hidden, int d_latent, int sm_count);
int launch_ae_pair(const float* x, int64_t m, int d_in, int d_hidden, int d_latent, const unsigned char* w0_slabs,
                   const unsigned char* w1_slabs, const float* b0, const float* b1, int l2norm, float* z, int* err_flag,
                   int sm_count, cudaStream_t st);

// linear layer Y = act(X W^T + b) (+ R) on tcgen05 (lk_gemm_umma.cu): X / W as split-bf16 slab planes
// (launch_ae_split_rows / ae_umma_weight_slabs with 128-row slabs), n % 128 == 0, k % 64 == 0
int gemm_umma_supported(int n, int k);
// y: fp32 [m, n], or null with y_planes: the result as operand planes of the next layer
int launch_gemm_umma(const unsigned char* x_slabs, int64_t m, int k, const unsigned char* w_slabs, int n,
                     const float* bias, const float* residual, int act, int n_planes, float* y,
                     unsigned char* y_planes, int* err_flag, int sm_count, cudaStream_t st);

// sentence-encoder kernels (lk_bert.cu); head dimension 32, hidden <= 1024
int bert_shape_supported(int hidden, int heads, int ffn);
int launch_bert_embed_ln(const int32_t* ids, int64_t n_tok, int s, int vocab, int hidden, const float* word,
                         const float* pos, const float* type0, const float* g, const float* b, float eps, float* out,
                         unsigned char* planes, int n_planes, cudaStream_t st);
// LayerNorm in place; planes (optional): the operand planes of the result for the next linear layer
int launch_bert_layernorm(float* x, int64_t n_tok, int hidden, const float* g, const float* b, float eps,
                          unsigned char* planes, int n_planes, cudaStream_t st);
// the context leaves as operand planes only (the output projection is its one reader)
int launch_bert_attention(const float* qkv, const int32_t* mask, int64_t n_sent, int s, int hidden, int heads,
                          unsigned char* ctx_planes, int n_planes, cudaStream_t st);
int launch_bert_pool(const float* x, const int32_t* mask, int64_t n_sent, int s, int hidden, int normalize, float* out,
                     cudaStream_t st);

}  // namespace lk
