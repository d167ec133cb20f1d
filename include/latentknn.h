/*
 * latentknn.h -- C ABI of liblatentknn.so, the B200 (sm_100a) exact nearest-neighbour
 * engine behind latent-rag's retriever classes.
 *
 * The reference (engares/latent-rag) is pure Python and has no FFI of its own: its
 * "operator interface" for this path is the duck-typed retriever class contract
 * (SURVEY.md section 8b).  Each entry point below names the reference code it replaces
 * (paths relative to the reference tree).  The Python mirror of those classes
 * (latent_rag_b200/retrieval/) is the only caller; it binds these symbols with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success and a negative lk_status on failure; the
 *     message is available from lk_last_error() (thread local).  Nothing throws.
 *   - pointers are plain host or device addresses; `mem` says which (lk_mem).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - one thread per handle at a time; handles are not re-entrant.
 *   - there is no CPU fallback anywhere: without a CUDA device every compute call
 *     fails with LK_ERR_CUDA.
 */
#ifndef LATENTKNN_H
#define LATENTKNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LK_ABI_VERSION 2

typedef enum lk_status {
  LK_OK = 0,
  LK_ERR_INVALID = -1,     /* bad argument                                  */
  LK_ERR_CUDA = -2,        /* CUDA runtime / launch failure                 */
  LK_ERR_OOM = -3,         /* device or host allocation failed              */
  LK_ERR_CAPACITY = -4,    /* lk_index_add past the capacity given at create */
  LK_ERR_UNSUPPORTED = -5  /* shape outside what the kernels implement      */
} lk_status;

/* similarity-name options of BruteForceRetriever (retrieval/bruteforce.py:14,49-54);
 * "mahalanobis" is the README feature (README.md:35) the reference never implemented. */
typedef enum lk_metric { LK_COSINE = 0, LK_EUCLIDEAN = 1, LK_MAHALANOBIS = 2 } lk_metric;
typedef enum lk_dtype { LK_F32 = 0, LK_BF16 = 1 } lk_dtype;
typedef enum lk_mem { LK_HOST = 0, LK_DEVICE = 1 } lk_mem;
/* which search kernel family to run (LK_KERNEL_AUTO picks by storage / batch / k) */
typedef enum lk_kernel { LK_KERNEL_AUTO = 0, LK_KERNEL_SIMT = 1, LK_KERNEL_UMMA = 2 } lk_kernel;
typedef enum lk_ae_kind { LK_AE_DAE = 0, LK_AE_CAE = 1, LK_AE_VAE_MU = 2 } lk_ae_kind;

typedef struct lk_index lk_index;
typedef struct lk_ae lk_ae;
typedef struct lk_comm lk_comm;
typedef struct lk_bert lk_bert;

#define LK_MAX_K 4096           /* largest top-k of lk_index_search / lk_merge_topk */
#define LK_MAX_K_FUSED 128      /* largest top-k one fused search selects in a single pass; also the
                                   limit of the peer exchange (lk_comm_*) */
#define LK_MAX_WORLD 16         /* ranks of one candidate exchange (one NVLink domain) */
#define LK_IPC_HANDLE_BYTES 64  /* sizeof(cudaIpcMemHandle_t) */

/* ---- library ----------------------------------------------------------------------- */
int lk_abi_version(void);
const char* lk_last_error(void);
/* number of visible CUDA devices (0 and LK_ERR_CUDA when there is no driver/device) */
int lk_device_count(int* out_count);
/* kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t lk_launch_count(void);

/* ---- index: replaces BruteForceRetriever.__init__ (retrieval/bruteforce.py:26-55) and
 *      FAISSEmbeddingRetriever.build's normalise+add (FAISSEmbeddingRetriever.py:206-257)
 *
 * storage LK_BF16: rows are rounded to bf16 and kept as 128-row tcgen05 operand tiles,
 *                  with one fp32 side value per row (1/|row| for cosine, |row|^2 for the
 *                  L2 metrics) applied in the kernel epilogue.
 * storage LK_F32 : rows are kept in fp32 (cosine: pre-normalised like the reference);
 *                  searched by the exact fp32 FMA kernel.
 * whiten         : for LK_MAHALANOBIS, the dim x dim row-major fp64 HOST matrix L with
 *                  precision = L L^T; rows and queries are multiplied by it (fp64
 *                  accumulate) and then searched with the L2 path.  NULL otherwise.
 */
int lk_index_create(lk_index** out, int device, int64_t capacity_rows, int dim, int metric,
                    int storage, const double* whiten);
/* append n_rows x dim row-major rows (dtype lk_dtype, host or device) */
int lk_index_add(lk_index* ix, const void* rows, int dtype, int mem, int64_t n_rows, void* stream);
int lk_index_size(const lk_index* ix, int64_t* out_rows, int* out_dim);
/* grow the capacity (device-to-device copy of the tiles); FAISSEmbeddingRetriever.build
 * appends to an existing index on every call (FAISSEmbeddingRetriever.py:252-257,294-296).
 * The copy is ordered on `stream` -- pass the stream the earlier lk_index_add calls ran on (or
 * synchronise first) -- and the call returns after it has completed. */
int lk_index_reserve(lk_index* ix, int64_t capacity_rows, void* stream);
int lk_index_destroy(lk_index* ix);

/* ---- persistence: replaces faiss.write_index / read_index
 *      (FAISSEmbeddingRetriever.py:65-69,300-304).  The payload is this library's own
 *      tiled image (not the upstream .faiss format): `tile_bytes` of operand tiles and
 *      `side_bytes` of fp32 side values covering the rows added so far. */
int lk_index_storage_bytes(const lk_index* ix, int64_t* out_tile_bytes, int64_t* out_side_bytes);
/* both copies are ordered on `stream` (behind the adds issued on it) and complete before return */
int lk_index_export(lk_index* ix, void* tiles_host, void* side_host, void* stream);
/* fill an empty index created with the same dim / metric / storage.  tile_bytes / side_bytes are
 * the lengths of the two host buffers and must be exactly what n_rows rows of this geometry take
 * (lk_index_storage_bytes of the exporting index): a truncated image is LK_ERR_INVALID, never read. */
int lk_index_import(lk_index* ix, const void* tiles_host, int64_t tile_bytes, const void* side_host,
                    int64_t side_bytes, int64_t n_rows, void* stream);

/* ---- search: replaces BruteForceRetriever.search (retrieval/bruteforce.py:58-83) and
 *      FAISSEmbeddingRetriever.search (FAISSEmbeddingRetriever.py:314-326).
 *
 * queries     b x dim row-major, q_dtype/q_mem as above
 * k           1..LK_MAX_K; the caller clamps to min(k, rows) like bruteforce.py:81.  Up to 128 the
 *             fused kernel selects in one pass over the corpus.  Above 128 the corpus is cut
 *             into row slabs, each slab's best 128 come from the same fused kernel and a
 *             sorting merge folds them into the result; a slab whose 128th best still reaches
 *             a query's k-th score is split and searched again, down to single 128-row blocks,
 *             so the result is the exact top-k under (score desc, row asc) for any data.
 * out_scores  b x k float32, best first, higher = better for every metric
 * out_idx     b x k int64 row positions + idx_base (idx_base = first global row of a shard)
 * out_mem     where the two outputs live; for LK_HOST the call returns after the copy
 * kernel      lk_kernel
 */
int lk_index_search(lk_index* ix, const void* queries, int q_dtype, int q_mem, int64_t b, int k,
                    float* out_scores, int64_t* out_idx, int out_mem, int64_t idx_base,
                    int kernel, void* stream);
/* With device outputs lk_index_search returns without synchronising; this waits for the device
 * and reports a search kernel whose bounded pipeline waits timed out since the last check
 * (LK_ERR_CUDA; with host outputs lk_index_search checks by itself). */
int lk_index_check(lk_index* ix);
/* device time (ms, CUDA events on `stream`) of the search kernel proper and of the whole
 * device side (query prep + search + merge) of the last lk_index_search on this handle;
 * used for StatsTracker (retrieval/common.py:37-65) and for bench.py's roofline. */
int lk_index_last_timing(lk_index* ix, float* out_search_kernel_ms, float* out_total_ms);
/* enable/disable that event timing (off by default: it synchronises the stream) */
int lk_index_set_timing(lk_index* ix, int enabled);

/* ---- k-way merge of per-shard candidates (net-new):
 * cand_* are b x n_lists x list_len, index < 0 or NaN score = padding.  Device or host.
 * k <= LK_MAX_K; for k > LK_MAX_K_FUSED at most 16384 candidates per query (n_lists x list_len). */
int lk_merge_topk(int device, const float* cand_scores, const int64_t* cand_idx, int64_t b,
                  int n_lists, int list_len, int k, float* out_scores, int64_t* out_idx,
                  int mem, void* stream);

/* ---- document-level MaxSim aggregation: replaces the per-query Python loop of the
 *      reference's caller (main.py:270-282).  cand_* are b x cand_k search results (best
 *      first, row ids; < 0 = padding), row_doc_ids maps a corpus row (chunk) to its document;
 *      out_* are b x top_k: the documents ranked by their best chunk, ties in first-seen
 *      order, padded with doc id -1 / score -inf.  All pointers are device memory. */
int lk_maxsim_rerank(int device, const float* cand_scores, const int64_t* cand_idx, int64_t b, int cand_k,
                     const int64_t* row_doc_ids, int64_t n_rows, int top_k, float* out_scores,
                     int64_t* out_doc_ids, void* stream);

/* ---- retrieval metrics per query: replaces recall_at_k / mrr / ndcg_at_k
 *      (evaluation/retrieval_metrics.py:14-31) under evaluate_retrieval (:55-96, main.py:321).
 *      retrieved: n_queries x n_retrieved ids (< 0 = padding of a shorter list); the relevant ids of
 *      query q are rel_ids[rel_offsets[q] .. rel_offsets[q+1]); metric m is metric_kind[m]
 *      (0 recall, 1 mrr, 2 ndcg) at cut-off metric_k[m] (<= 0: the whole list); discounts[i] =
 *      1 / log2(i + 2) in float64, max(n_retrieved, largest metric_k) entries (the ideal DCG keeps the
 *      caller's cut-off when fewer ids were retrieved, retrieval_metrics.py:29); out: n_queries x n_metrics float64.  All pointers are device
 *      memory.  The values equal the reference's bit for bit (same float64 sums, left to right). */
int lk_retrieval_metrics(int device, const int64_t* retrieved, int64_t n_queries, int n_retrieved,
                         const int64_t* rel_offsets, const int64_t* rel_ids, const int* metric_kind,
                         const int* metric_k, int n_metrics, const double* discounts, double* out,
                         void* stream);

/* ---- 1-based rank of every query's paired document among all documents by cosine similarity:
 *      replaces _rank_positive (evaluation/embedding_visualization.py:34-37), which broadcasts an
 *      [n, n, D] tensor and argsorts twice.  queries, docs: n x dim fp32 row-major, device memory;
 *      out_rank: n int64, device.  rank = 1 + #{j != i : cos(q_i, d_j) > cos(q_i, d_i)}. */
int lk_rank_positive(int device, const float* queries, const float* docs, int64_t n, int dim,
                     int64_t* out_rank, void* stream);

/* ---- candidate exchange between the GPUs of a row-sharded index, fused with that merge
 *      (net-new).  One process per GPU.  Every rank owns a symmetric
 *      buffer; the peers' buffers are mapped through CUDA IPC (the handles travel over the
 *      host-side process group).  lk_comm_exchange_merge is ONE kernel per rank and call:
 *      it stores this rank's b x k candidates (global ids) straight into every peer's buffer
 *      over NVLink, releases per-query flags, waits for the peers' flags and merges the
 *      world x k candidates of each query.  All ranks must call it in the same order.  Every rank's k
 *      candidates of a query must be sorted best first under (score desc, id asc), as lk_index_search
 *      returns them (short lists padded at the end with id -1): the merge ranks them by binary searches.  The wait for
 *      a peer is bounded (60 s; LK_XCHG_TIMEOUT_S overrides) only to turn a dead peer into an error:
 *      on a timeout the affected queries come back as id -1 / score -inf and lk_comm_check reports
 *      LK_ERR_CUDA -- callers that keep the outputs on the device must call lk_comm_check.
 *
 *      lk_comm_attach_local / lk_comm_begin / lk_comm_publish / lk_comm_collect split the
 *      same protocol into steps so that several ranks can be driven from ONE process (tests
 *      on a single GPU: begin + publish on every rank first, then collect on every rank). */
int lk_comm_create(lk_comm** out, int device, int rank, int world, int64_t max_b, int max_k);
/* out_handle: LK_IPC_HANDLE_BYTES bytes identifying this rank's buffer to other processes */
int lk_comm_ipc_handle(lk_comm* c, void* out_handle);
/* handles: world x LK_IPC_HANDLE_BYTES bytes, rank-major (the own entry is ignored) */
int lk_comm_open_peers(lk_comm* c, const void* handles);
int lk_comm_attach_local(lk_comm* c, int peer_rank, lk_comm* peer);
int lk_comm_exchange_merge(lk_comm* c, const float* local_scores, const int64_t* local_idx, int64_t b, int k,
                           float* out_scores, int64_t* out_idx, void* stream);
int lk_comm_begin(lk_comm* c);
int lk_comm_publish(lk_comm* c, const float* local_scores, const int64_t* local_idx, int64_t b, int k,
                    void* stream);
int lk_comm_collect(lk_comm* c, int64_t b, int k, float* out_scores, int64_t* out_idx, void* stream);
/* synchronises with the device: LK_ERR_CUDA if a wait for a peer timed out since the last check */
int lk_comm_check(lk_comm* c);
int lk_comm_destroy(lk_comm* c);

/* ---- autoencoder encoder forward: replaces DenoisingAutoencoder.encode
 *      (models/denoising_autoencoder.py:33-34), ContrastiveAutoencoder.encode
 *      (models/contrastive_autoencoder.py:23-25) and the mu half of
 *      VariationalAutoencoder.encode (models/variational_autoencoder.py:26-30,
 *      retrieval/embedder.py:44-45).  Weights are the nn.Linear tensors, fp32 HOST:
 *      w0 [d_hidden x d_in], b0 [d_hidden], w1 [d_latent x d_hidden], b1 [d_latent]. */
int lk_ae_create(lk_ae** out, int device, int kind, int d_in, int d_hidden, int d_latent,
                 const float* w0, const float* b0, const float* w1, const float* b1);
/* LK_KERNEL_UMMA: tensor-core kernel with split-bf16 operands (x = hi + lo, three MMAs per
 * product, fp32 accumulate; ~1e-5 of the row scale), available when d_in % 64 == 0,
 * d_hidden % 128 == 0 and d_latent <= 64; LK_KERNEL_SIMT: fp32 FMA kernel (any dims, the
 * reference's arithmetic); LK_KERNEL_AUTO (default): UMMA for calls of >= 256 rows when
 * available, SIMT otherwise. */
int lk_ae_set_kernel(lk_ae* ae, int kernel);
/* Operand precision of the tensor-core kernel.  LK_F32 (default): split-bf16 operands, fp32-level
 * results.  LK_BF16: inputs, weights and hidden activations rounded to bf16 once, one MMA per
 * product, fp32 accumulate (the same stated precision as the bf16 search): ~2x faster, and the
 * latents are stored in bf16 by the search index anyway.  Implies the tensor-core kernel. */
int lk_ae_set_precision(lk_ae* ae, int precision);
/* x: m x d_in fp32 row-major; z: m x d_latent fp32 row-major.  Device buffers: asynchronous on `stream`.
 * Host buffers (either side): the rows go through in chunks of 32768 whose host-to-device copy, kernel and
 * device-to-host copy overlap (two staging buffers each way, the copies on streams of the handle); page-locked
 * host memory makes the copies run at link speed (pageable rows are staged through page-locked blocks by a few
 * threads of the library, as for lk_index_add and host queries); the call returns when z is complete. */
int lk_ae_encode(lk_ae* ae, const float* x, int x_mem, int64_t m, float* z, int z_mem, void* stream);
int lk_ae_destroy(lk_ae* ae);

/* ---- sentence encoder forward: replaces SentenceTransformer("all-MiniLM-L6-v2").encode as the
 *      reference calls it (retrieval/embedder.py:35-40: convert_to_tensor, normalize_embeddings=True)
 *      from the token ids on: BERT embeddings + LayerNorm, n_layers x (self-attention, output
 *      projection + residual + LayerNorm, GELU feed-forward + residual + LayerNorm), masked mean
 *      pooling, L2 normalisation.  Tokenisation (WordPiece, third-party vocabulary) stays on the
 *      host.  Weights are the tensors of a transformers BertModel state_dict, fp32 HOST, nn.Linear
 *      layout [out, in].  The kernels implement head dimension 32 (hidden = 32 x heads), hidden and
 *      ffn multiples of 128, hidden <= 1024: all-MiniLM-L6-v2 is hidden 384, 12 heads, ffn 1536, 6
 *      layers, 512 positions, eps 1e-12. */
typedef struct lk_bert_layer_weights {
  const float *wq, *bq, *wk, *bk, *wv, *bv; /* attention.self.{query,key,value}.{weight,bias}          */
  const float *wo, *bo, *ln1_g, *ln1_b;     /* attention.output.dense, attention.output.LayerNorm      */
  const float *w1, *b1;                     /* intermediate.dense [ffn x hidden]                       */
  const float *w2, *b2, *ln2_g, *ln2_b;     /* output.dense [hidden x ffn], output.LayerNorm           */
} lk_bert_layer_weights;
typedef struct lk_bert_weights {
  const float* word_emb;                /* embeddings.word_embeddings.weight [vocab x hidden]          */
  const float* pos_emb;                 /* embeddings.position_embeddings.weight [max_pos x hidden]    */
  const float* type_emb;                /* embeddings.token_type_embeddings.weight: row 0 is used      */
  const float *emb_ln_g, *emb_ln_b;     /* embeddings.LayerNorm                                        */
  const lk_bert_layer_weights* layers;  /* n_layers entries: encoder.layer.<l>                         */
} lk_bert_weights;
int lk_bert_create(lk_bert** out, int device, int vocab, int max_pos, int hidden, int heads, int ffn,
                   int n_layers, float ln_eps, const lk_bert_weights* w);
/* Operand precision of the linear layers (tcgen05).  LK_F32 (default): split-bf16 operands, three
 * MMAs per product, fp32-level results.  LK_BF16: operands rounded to bf16, one MMA per product. */
int lk_bert_set_precision(lk_bert* m, int precision);
/* input_ids / attention_mask: n_sent x seq_len int32 (mask 0 = padding), host or device (ids_mem);
 * out: n_sent x hidden fp32, host or device (out_mem); normalize: L2-normalise the pooled rows.
 * With device outputs the call does not synchronise: lk_bert_check reports a timed-out pipeline. */
int lk_bert_encode(lk_bert* m, const int32_t* input_ids, const int32_t* attention_mask, int ids_mem,
                   int64_t n_sent, int seq_len, int normalize, float* out, int out_mem, void* stream);
int lk_bert_check(lk_bert* m);
int lk_bert_destroy(lk_bert* m);
/* One torch.nn.Linear forward through the encoder's tcgen05 path, for tests and bring-up:
 * y [m x n] = act(x [m x k] w^T + bias) + residual; w [n x k]; bias / residual may be NULL; act 0 =
 * none, 1 = GELU (erf); precision LK_F32 (split-bf16) or LK_BF16; all pointers fp32 HOST memory;
 * n % 128 == 0, k % 64 == 0. */
int lk_linear_forward(int device, const float* x, int64_t m, int k, const float* w, int n,
                      const float* bias, const float* residual, int act, int precision, float* y);

#ifdef __cplusplus
}
#endif
#endif /* LATENTKNN_H */
