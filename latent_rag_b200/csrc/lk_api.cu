// C ABI of liblatentknn.so (include/latentknn.h): handles, workspaces, staging of host
// buffers, kernel selection.  No exceptions cross this boundary and nothing here computes
// on the CPU: without a CUDA device every entry point fails.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "lk_common.cuh"
#include "lk_host.cuh"

namespace lk {
void ae_umma_weight_slabs(const float* w, int rows_out, int k_in, int slab_rows, std::vector<unsigned char>* out,
                          bool plane_major_over_kb_only);

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  const char* base = strrchr(file, '/');
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), base ? base + 1 : file, line,
            what);
  return e == cudaErrorMemoryAllocation ? LK_ERR_OOM : LK_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace {

TileGeom make_geom(int dim, int storage) {
  TileGeom g;
  g.dim = dim;
  g.elem_bytes = storage == LK_BF16 ? 2 : 4;
  g.dim_pad = round_up(dim, kRowBytes / g.elem_bytes);
  g.kblocks = g.dim_pad * g.elem_bytes / kRowBytes;
  return g;
}

// row blocks allocated for `rows` rows: whole PAIRS of blocks, because the tcgen05 kernel
// consumes two blocks per unit (the padding block is zeros with NaN side values)
inline int64_t alloc_blocks(int64_t rows) { return (rows + 2 * kBlockRows - 1) / (2 * kBlockRows) * 2; }

constexpr int64_t kStageRows = 1 << 16;  // rows staged per step when the input is on the host

}  // namespace
}  // namespace lk

using namespace lk;

struct lk_index {
  int device = 0, sm_count = 0;
  int dim = 0, metric = 0, kmetric = 0, storage = 0;
  int side_mode = 0, prenorm = 0;
  TileGeom g;
  int64_t capacity = 0, n_rows = 0;
  unsigned char* tiles = nullptr;
  float* side = nullptr;
  double* whiten = nullptr;
  int* err_flag = nullptr;
  int* ticket = nullptr;  // completion counters of the single-launch small-batch search
  unsigned char* pin = nullptr;  // pinned, device-visible host staging of that path: queries | scores | ids
  size_t pin_cap = 0;
  HostStager up;                 // host rows / queries -> device (pageable sources staged by threads)
  // fp32 storage on the tensor cores: split-bf16 planes of the rows (x = hi + lo), built lazily from the fp32
  // tiles by the first tcgen05 search and extended as rows are added; the fp32 tiles stay (exact FMA kernel)
  TileGeom gp;                    // geometry of the planes: kblocks = 2 * ceil(dim / 64)
  unsigned char* planes = nullptr;
  int64_t planes_blocks = 0;      // row blocks allocated
  int64_t planes_rows = 0;        // rows converted so far
  // what the search kernels of the current call read: the tiles as stored, or the planes
  const unsigned char* sv_tiles = nullptr;
  TileGeom sv_g;
  int sv_split = 0;
  Buf stage, white, q_tiles, q_side, q_planes, part_s, part_i, part_c, out_s, out_i, debug, mtmp_s, mtmp_i;
  Buf deep_s, deep_i, deep_last, deep_flags, deep_tiles, deep_side;  // slab search (k > 128)
  bool timing = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float last_search_ms = 0.f, last_total_ms = 0.f;
};

struct lk_ae {
  int device = 0, sm_count = 0;
  int kind = 0, d_in = 0, d_hidden = 0, d_latent = 0;
  int kernel = LK_KERNEL_AUTO;  // AUTO/UMMA: split-bf16 tensor-core kernel when the dims allow; SIMT: fp32 FMA
  int precision = LK_F32;       // tensor-core kernel: LK_F32 = split-bf16 operands (3 MMAs), LK_BF16 = plain bf16
  float *w0t = nullptr, *b0 = nullptr, *w1t = nullptr, *b1 = nullptr;
  unsigned char *w0_slabs = nullptr, *w1_slabs = nullptr;
  int* err_flag = nullptr;
  Buf xslabs;
  // host inputs / outputs: two staging buffers each way, copies on their own streams so that the upload of chunk
  // i + 1 and the download of chunk i - 1 run under the kernel of chunk i
  Buf xin[2], zout[2];
  HostStager up;
  cudaStream_t cs_in = nullptr, cs_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};

extern "C" {

int lk_abi_version(void) { return LK_ABI_VERSION; }
const char* lk_last_error(void) { return g_err; }
int64_t lk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int lk_device_count(int* out_count) {
  if (!out_count) return LK_ERR_INVALID;
  *out_count = 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
  *out_count = n;
  return LK_OK;
}

// ------------------------------------------------------------------------------------
// index
// ------------------------------------------------------------------------------------
int lk_index_destroy(lk_index* ix) {
  if (!ix) return LK_OK;
  DeviceGuard guard(ix->device);
  if (ix->tiles) cudaFree(ix->tiles);
  if (ix->side) cudaFree(ix->side);
  if (ix->whiten) cudaFree(ix->whiten);
  if (ix->err_flag) cudaFree(ix->err_flag);
  if (ix->ticket) cudaFree(ix->ticket);
  if (ix->pin) cudaFreeHost(ix->pin);
  ix->up.release();
  if (ix->planes) cudaFree(ix->planes);
  Buf* bufs[] = {&ix->stage, &ix->white, &ix->q_tiles, &ix->q_side, &ix->q_planes, &ix->mtmp_s, &ix->mtmp_i, &ix->part_s, &ix->part_i, &ix->part_c,
                 &ix->out_s, &ix->out_i, &ix->debug, &ix->deep_s, &ix->deep_i, &ix->deep_last, &ix->deep_flags,
                 &ix->deep_tiles, &ix->deep_side};
  for (Buf* b : bufs) b->release();
  for (cudaEvent_t e : ix->ev)
    if (e) cudaEventDestroy(e);
  delete ix;
  return LK_OK;
}

int lk_index_create(lk_index** out, int device, int64_t capacity_rows, int dim, int metric, int storage,
                    const double* whiten) {
  if (!out) return LK_ERR_INVALID;
  *out = nullptr;
  if (capacity_rows < 1 || capacity_rows > 0x7fffff00LL || dim < 1 || dim > 4096) {
    set_error("lk_index_create: capacity_rows=%lld dim=%d out of range", (long long)capacity_rows, dim);
    return LK_ERR_INVALID;
  }
  if (metric != LK_COSINE && metric != LK_EUCLIDEAN && metric != LK_MAHALANOBIS) {
    set_error("Unsupported metric: %d", metric);
    return LK_ERR_INVALID;
  }
  if (storage != LK_F32 && storage != LK_BF16) {
    set_error("lk_index_create: storage must be LK_F32 or LK_BF16");
    return LK_ERR_INVALID;
  }
  if ((metric == LK_MAHALANOBIS) != (whiten != nullptr)) {
    set_error("lk_index_create: the whitening matrix is required for mahalanobis and only for it");
    return LK_ERR_INVALID;
  }
  int sm_count = 0;
  int rc = check_device(device, &sm_count);
  if (rc != LK_OK) return rc;
  DeviceGuard guard(device);
  lk_index* ix = new (std::nothrow) lk_index();
  if (!ix) return LK_ERR_OOM;
  ix->device = device;
  ix->sm_count = sm_count;
  ix->dim = dim;
  ix->metric = metric;
  ix->kmetric = metric == LK_COSINE ? LK_COSINE : LK_EUCLIDEAN;
  ix->storage = storage;
  ix->g = make_geom(dim, storage);
  ix->gp = make_geom(dim, LK_BF16);
  ix->gp.kblocks *= 2;  // [hi plane | lo plane]
  ix->capacity = capacity_rows;
  ix->side_mode = ix->kmetric == LK_COSINE ? 0 : 1;
  ix->prenorm = (ix->kmetric == LK_COSINE && storage == LK_F32) ? 1 : 0;
  const int64_t nblk = alloc_blocks(capacity_rows);
  const size_t tile_bytes = (size_t)nblk * ix->g.block_bytes();
  const size_t side_bytes = (size_t)nblk * kBlockRows * sizeof(float);
#define LK_CREATE_CUDA(expr)                                              \
  do {                                                                    \
    cudaError_t _e = (expr);                                              \
    if (_e != cudaSuccess) {                                              \
      rc = cuda_fail(_e, #expr, __FILE__, __LINE__);                      \
      lk_index_destroy(ix);                                               \
      return rc;                                                          \
    }                                                                     \
  } while (0)
  LK_CREATE_CUDA(cudaMalloc((void**)&ix->tiles, tile_bytes));
  LK_CREATE_CUDA(cudaMalloc((void**)&ix->side, side_bytes));
  LK_CREATE_CUDA(cudaMalloc((void**)&ix->err_flag, sizeof(int)));
  LK_CREATE_CUDA(cudaMemset(ix->tiles, 0, tile_bytes));
  LK_CREATE_CUDA(cudaMemset(ix->side, 0xFF, side_bytes));  // NaN: rows that do not exist never rank
  LK_CREATE_CUDA(cudaMemset(ix->err_flag, 0, sizeof(int)));
  LK_CREATE_CUDA(cudaMalloc((void**)&ix->ticket, 16 * sizeof(int)));
  LK_CREATE_CUDA(cudaMemset(ix->ticket, 0, 16 * sizeof(int)));
  if (whiten) {
    const size_t wb = (size_t)dim * dim * sizeof(double);
    LK_CREATE_CUDA(cudaMalloc((void**)&ix->whiten, wb));
    LK_CREATE_CUDA(cudaMemcpy(ix->whiten, whiten, wb, cudaMemcpyHostToDevice));
  }
  for (int i = 0; i < 4; ++i) LK_CREATE_CUDA(cudaEventCreate(&ix->ev[i]));
#undef LK_CREATE_CUDA
  *out = ix;
  return LK_OK;
}

int lk_index_size(const lk_index* ix, int64_t* out_rows, int* out_dim) {
  if (!ix) return LK_ERR_INVALID;
  if (out_rows) *out_rows = ix->n_rows;
  if (out_dim) *out_dim = ix->dim;
  return LK_OK;
}

// rows (host or device, f32/bf16) -> tiles at [row0, row0+n); shared by add and by query prep
static int ingest_rows(lk_index* ix, const void* rows, int dtype, int mem, int64_t n, void* tiles, float* side,
                       int64_t row0, cudaStream_t st) {
  const size_t in_elem = dtype == LK_F32 ? 4 : 2;
  const bool direct = mem == LK_DEVICE && !ix->whiten;
  if (direct) return launch_tile_rows(rows, dtype, n, ix->g, tiles, side, row0, ix->side_mode, ix->prenorm, st);
  for (int64_t done = 0; done < n; done += kStageRows) {
    const int64_t cnt = n - done < kStageRows ? n - done : kStageRows;
    const unsigned char* src = static_cast<const unsigned char*>(rows) + (size_t)done * ix->dim * in_elem;
    const void* cur = src;
    if (mem == LK_HOST) {
      int rc = ix->stage.ensure((size_t)kStageRows * ix->dim * in_elem);
      if (rc != LK_OK) return rc;
      if ((rc = ix->up.upload(ix->stage.p, src, (size_t)cnt * ix->dim * in_elem, st)) != LK_OK) return rc;
      cur = ix->stage.p;
    }
    int cur_dtype = dtype;
    if (ix->whiten) {
      int rc = ix->white.ensure((size_t)kStageRows * ix->dim * sizeof(float));
      if (rc != LK_OK) return rc;
      rc = launch_whiten(cur, dtype, cnt, ix->dim, ix->whiten, ix->white.as<float>(), st);
      if (rc != LK_OK) return rc;
      cur = ix->white.p;
      cur_dtype = LK_F32;
    }
    int rc = launch_tile_rows(cur, cur_dtype, cnt, ix->g, tiles, side, row0 + done, ix->side_mode, ix->prenorm, st);
    if (rc != LK_OK) return rc;
    // the staging buffers are reused by the next step
    if (done + cnt < n && mem == LK_HOST) LK_CUDA(cudaStreamSynchronize(st));
  }
  return LK_OK;
}

int lk_index_add(lk_index* ix, const void* rows, int dtype, int mem, int64_t n_rows, void* stream) {
  if (!ix || (!rows && n_rows > 0) || n_rows < 0 || (dtype != LK_F32 && dtype != LK_BF16) ||
      (mem != LK_HOST && mem != LK_DEVICE)) {
    set_error("lk_index_add: bad argument");
    return LK_ERR_INVALID;
  }
  if (ix->n_rows + n_rows > ix->capacity) {
    set_error("lk_index_add: %lld + %lld rows exceed the capacity %lld", (long long)ix->n_rows,
              (long long)n_rows, (long long)ix->capacity);
    return LK_ERR_CAPACITY;
  }
  DeviceGuard guard(ix->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ingest_rows(ix, rows, dtype, mem, n_rows, ix->tiles, ix->side, ix->n_rows, st);
  if (rc != LK_OK) return rc;
  if (mem == LK_HOST) LK_CUDA(cudaStreamSynchronize(st));  // the caller may free `rows` on return
  ix->n_rows += n_rows;
  return LK_OK;
}

int lk_index_reserve(lk_index* ix, int64_t capacity_rows, void* stream) {
  if (!ix || capacity_rows > 0x7fffff00LL) {
    set_error("lk_index_reserve: bad argument");
    return LK_ERR_INVALID;
  }
  if (capacity_rows <= ix->capacity) return LK_OK;
  DeviceGuard guard(ix->device);
  // The copies run on the caller's stream, behind the tiling kernels of earlier lk_index_add calls
  // on it (torch's side streams do not synchronise with the legacy default stream, so a plain
  // cudaMemcpy could overtake them); the stream is drained before the old buffers are freed.
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t old_blk = alloc_blocks(ix->capacity);
  const int64_t new_blk = alloc_blocks(capacity_rows);
  const size_t bb = (size_t)ix->g.block_bytes();
  unsigned char* tiles = nullptr;
  float* side = nullptr;
  cudaError_t e = cudaMalloc((void**)&tiles, (size_t)new_blk * bb);
  if (e == cudaSuccess) e = cudaMalloc((void**)&side, (size_t)new_blk * kBlockRows * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpyAsync(tiles, ix->tiles, (size_t)old_blk * bb, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(tiles + (size_t)old_blk * bb, 0, (size_t)(new_blk - old_blk) * bb, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(side, ix->side, (size_t)old_blk * kBlockRows * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess)
    e = cudaMemsetAsync(side + (size_t)old_blk * kBlockRows, 0xFF,
                        (size_t)(new_blk - old_blk) * kBlockRows * sizeof(float), st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    if (tiles) cudaFree(tiles);
    if (side) cudaFree(side);
    return cuda_fail(e, "lk_index_reserve", __FILE__, __LINE__);
  }
  cudaFree(ix->tiles);
  cudaFree(ix->side);
  if (ix->planes) cudaFree(ix->planes);  // rebuilt from the new tiles by the next tensor-core search
  ix->planes = nullptr;
  ix->planes_blocks = ix->planes_rows = 0;
  ix->tiles = tiles;
  ix->side = side;
  ix->capacity = capacity_rows;
  return LK_OK;
}

int lk_index_storage_bytes(const lk_index* ix, int64_t* out_tile_bytes, int64_t* out_side_bytes) {
  if (!ix) return LK_ERR_INVALID;
  const int64_t nblk = (ix->n_rows + kBlockRows - 1) / kBlockRows;
  if (out_tile_bytes) *out_tile_bytes = nblk * ix->g.block_bytes();
  if (out_side_bytes) *out_side_bytes = nblk * kBlockRows * (int64_t)sizeof(float);
  return LK_OK;
}

int lk_index_export(lk_index* ix, void* tiles_host, void* side_host, void* stream) {
  if (!ix || !tiles_host || !side_host) return LK_ERR_INVALID;
  DeviceGuard guard(ix->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);  // behind the adds issued on this stream
  int64_t tb = 0, sb = 0;
  lk_index_storage_bytes(ix, &tb, &sb);
  LK_CUDA(cudaMemcpyAsync(tiles_host, ix->tiles, (size_t)tb, cudaMemcpyDeviceToHost, st));
  LK_CUDA(cudaMemcpyAsync(side_host, ix->side, (size_t)sb, cudaMemcpyDeviceToHost, st));
  LK_CUDA(cudaStreamSynchronize(st));
  return LK_OK;
}

int lk_index_import(lk_index* ix, const void* tiles_host, int64_t tile_bytes, const void* side_host,
                    int64_t side_bytes, int64_t n_rows, void* stream) {
  if (!ix || !tiles_host || !side_host || n_rows < 0) return LK_ERR_INVALID;
  if (ix->n_rows != 0) {
    set_error("lk_index_import: the index is not empty");
    return LK_ERR_INVALID;
  }
  if (n_rows > ix->capacity) {
    set_error("lk_index_import: %lld rows exceed the capacity %lld", (long long)n_rows, (long long)ix->capacity);
    return LK_ERR_CAPACITY;
  }
  const int64_t nblk = (n_rows + kBlockRows - 1) / kBlockRows;
  const int64_t want_tiles = nblk * ix->g.block_bytes(), want_side = nblk * kBlockRows * (int64_t)sizeof(float);
  if (tile_bytes != want_tiles || side_bytes != want_side) {  // a truncated or foreign image must not be read past its end
    set_error("lk_index_import: %lld rows of dim %d need %lld tile bytes and %lld side bytes, got %lld and %lld",
              (long long)n_rows, ix->dim, (long long)want_tiles, (long long)want_side, (long long)tile_bytes,
              (long long)side_bytes);
    return LK_ERR_INVALID;
  }
  DeviceGuard guard(ix->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LK_CUDA(cudaMemcpyAsync(ix->tiles, tiles_host, (size_t)want_tiles, cudaMemcpyHostToDevice, st));
  LK_CUDA(cudaMemcpyAsync(ix->side, side_host, (size_t)want_side, cudaMemcpyHostToDevice, st));
  LK_CUDA(cudaStreamSynchronize(st));  // the caller may free the host buffers on return
  ix->n_rows = n_rows;
  return LK_OK;
}

int lk_index_set_timing(lk_index* ix, int enabled) {
  if (!ix) return LK_ERR_INVALID;
  ix->timing = enabled != 0;
  return LK_OK;
}

int lk_index_last_timing(lk_index* ix, float* out_search_kernel_ms, float* out_total_ms) {
  if (!ix) return LK_ERR_INVALID;
  if (out_search_kernel_ms) *out_search_kernel_ms = ix->last_search_ms;
  if (out_total_ms) *out_total_ms = ix->last_total_ms;
  return LK_OK;
}

int lk_index_check(lk_index* ix) {
  if (!ix) return LK_ERR_INVALID;
  DeviceGuard guard(ix->device);
  int flag = 0;
  // every stream of the device: a plain cudaMemcpy only orders behind blocking streams, and torch's side streams are
  // cudaStreamNonBlocking (a search still running on one would be read as "no timeout")
  LK_CUDA(cudaDeviceSynchronize());
  LK_CUDA(cudaMemcpy(&flag, ix->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag != 0) {
    cudaMemset(ix->err_flag, 0, sizeof(int));
    set_error("search kernel pipeline timed out (barrier code %d); results are invalid", flag);
    return LK_ERR_CUDA;
  }
  return LK_OK;
}

// the geometry the tcgen05 kernel would read for this index: the bf16 tiles, or the split-bf16 planes of fp32 rows
static const TileGeom& umma_geom(const lk_index* ix) { return ix->storage == LK_BF16 ? ix->g : ix->gp; }

static int pick_kernel(const lk_index* ix, int requested, int k, int64_t b) {
  const char* env = getenv("LK_FORCE_KERNEL");
  if (env && !strcmp(env, "simt")) requested = LK_KERNEL_SIMT;
  if (env && !strcmp(env, "umma")) requested = LK_KERNEL_UMMA;
  if (requested == LK_KERNEL_AUTO) {
    if (!umma_supported(umma_geom(ix), k)) return LK_KERNEL_SIMT;
    // fp32 storage: the exact FMA kernel for tiny problems (no planes to build, bit-faithful fp32 products),
    // the tensor cores on split-bf16 planes (x = hi + lo, three MMAs per product) from ~1 M scores on
    if (ix->storage == LK_F32 && b * ix->n_rows < (1 << 20)) return LK_KERNEL_SIMT;
    return LK_KERNEL_UMMA;
  }
  return requested;
}

// fp32 storage, tensor-core search: make the planes cover every row added so far
static int ensure_planes(lk_index* ix, cudaStream_t st) {
  const int64_t need_blocks = alloc_blocks(ix->capacity);
  if (!ix->planes || ix->planes_blocks < need_blocks) {
    if (ix->planes) cudaFree(ix->planes);
    ix->planes = nullptr;
    ix->planes_rows = ix->planes_blocks = 0;
    LK_CUDA(cudaMalloc((void**)&ix->planes, (size_t)need_blocks * ix->gp.block_bytes()));
    ix->planes_blocks = need_blocks;
  }
  if (ix->planes_rows < ix->n_rows) {
    // whole row blocks, from the block the first new row lives in to the (pair-aligned) last: rows that do
    // not exist yet are zeros in the fp32 tiles and NaN-sided, so they convert to zeros and never rank
    const int64_t blk0 = ix->planes_rows / kBlockRows, blk1 = alloc_blocks(ix->n_rows);
    int rc = launch_planes_from_tiles(ix->tiles, ix->g, blk0, blk1 - blk0, ix->gp, ix->planes, st);
    if (rc != LK_OK) return rc;
    ix->planes_rows = ix->n_rows;
  }
  return LK_OK;
}


// ------------------------------------------------------------------------------------
// search over a run of rows (the whole corpus, or one slab of a deep search)
// ------------------------------------------------------------------------------------
static SearchArgs rows_args(const lk_index* ix, const unsigned char* tiles, const float* side, int64_t n_rows,
                            const unsigned char* q_tiles, const float* q_side, int64_t b, int k) {
  SearchArgs a = {};
  a.tiles = tiles;
  a.side = side;
  a.n_rows = n_rows;
  a.g = ix->sv_g;
  a.split_n = ix->sv_split;
  a.q_tiles = q_tiles;
  a.q_side = q_side;
  a.n_queries = b;
  a.metric = ix->kmetric;
  a.k = k;
  a.err_flag = ix->err_flag;
  return a;
}

// Plan, fused distance + selection, merge of the partial lists: [b, k] results into d_s / d_i
// (memory the device can write; ids = row position within the run + idx_base).  The run starts
// at a 128-row block boundary and covers whole 256-row units unless it ends with the corpus:
// the kernels score whole units and rely on the NaN side values past the last row.
// events: record ix->ev[1] / ev[2] around the search kernels.
static int search_rows(lk_index* ix, int which, SearchArgs a, int64_t idx_base, bool events, const char* dump_path,
                       float* d_s, int64_t* d_i, cudaStream_t st) {
  int rc;
  const int64_t b = a.n_queries;
  const int k = a.k;
  a.seed = nullptr;
  a.debug_tile = nullptr;
  if (dump_path) {  // bring-up aid: raw accumulator of unit 0 -> file
    if ((rc = ix->debug.ensure(kBlockRows * kBlockRows * sizeof(float))) != LK_OK) return rc;
    LK_CUDA(cudaMemsetAsync(ix->debug.p, 0xFF, kBlockRows * kBlockRows * sizeof(float), st));
    a.debug_tile = ix->debug.as<float>();
  }
  if (which == LK_KERNEL_UMMA) rc = umma_plan(a, ix->sm_count, &a.n_lists, &a.ksel);
  else rc = simt_plan(a, ix->sm_count, &a.n_lists, &a.ksel);
  if (rc != LK_OK) return rc;
  const int merge_len = a.ksel > kMaxK ? a.ksel : (k < a.ksel ? k : a.ksel);
  const int64_t seed_rows = which == LK_KERNEL_UMMA ? umma_seed_rows(a, ix->sm_count) : 0;
  SearchArgs a0 = a;  // the seeding search over a prefix of the run
  if (seed_rows > 0) {
    a0.n_rows = seed_rows;
    if ((rc = umma_plan(a0, ix->sm_count, &a0.n_lists, &a0.ksel)) != LK_OK) return rc;
  }
  const size_t n_part = (size_t)b * a.n_lists * a.ksel;
  const size_t n_part0 = seed_rows > 0 ? (size_t)b * a0.n_lists * a0.ksel : 0;
  const size_t n_part_max = n_part > n_part0 ? n_part : n_part0;
  if ((rc = ix->part_s.ensure(n_part_max * sizeof(float))) != LK_OK) return rc;
  if ((rc = ix->part_i.ensure(n_part_max * sizeof(int32_t))) != LK_OK) return rc;
  a.part_scores = a0.part_scores = ix->part_s.as<float>();
  a.part_idx = a0.part_idx = ix->part_i.as<int32_t>();
  a.part_cnt = a0.part_cnt = nullptr;
  const bool counted = which == LK_KERNEL_UMMA && a.ksel > kMaxK;  // append buffers report their fill
  const size_t n_cnt = (size_t)b * a.n_lists, n_cnt0 = seed_rows > 0 ? (size_t)b * a0.n_lists : 0;
  if (counted) {
    if ((rc = ix->part_c.ensure((n_cnt > n_cnt0 ? n_cnt : n_cnt0) * sizeof(int))) != LK_OK) return rc;
    a.part_cnt = a0.part_cnt = ix->part_c.as<int>();
  }

  // scratch of the two-level merge (a few queries, very many candidates each)
  constexpr int64_t kMergeTmpEntries = 148 * kMaxK;
  float* mt_s = nullptr;
  int64_t* mt_i = nullptr;
  if (b * 2 <= 148) {
    if ((rc = ix->mtmp_s.ensure(kMergeTmpEntries * sizeof(float))) != LK_OK) return rc;
    if ((rc = ix->mtmp_i.ensure(kMergeTmpEntries * sizeof(int64_t))) != LK_OK) return rc;
    mt_s = ix->mtmp_s.as<float>();
    mt_i = ix->mtmp_i.as<int64_t>();
  }

  // fused distance + selection
  if (events) LK_CUDA(cudaEventRecord(ix->ev[1], st));
  if (seed_rows > 0) {  // (the tcgen05 kernel marks the list slots it does not own itself)
    if ((rc = launch_search_umma(a0, ix->sm_count, st)) != LK_OK) return rc;
    rc = launch_merge_i32(a0.part_scores, a0.part_idx, a0.part_cnt, b, a0.n_lists, merge_len, a0.ksel, k, 0, d_s,
                          d_i, st, mt_s, mt_i, kMergeTmpEntries);
    if (rc != LK_OK) return rc;
    a.seed = d_s;  // read at the start of the main kernel's segments, overwritten by the final merge
  }
  if (which != LK_KERNEL_UMMA) {
    LK_CUDA(cudaMemsetAsync(a.part_scores, 0xFF, n_part * sizeof(float), st));   // NaN = empty slot
    LK_CUDA(cudaMemsetAsync(a.part_idx, 0xFF, n_part * sizeof(int32_t), st));    // -1
  } else if (const char* e = getenv("LK_DBG")) {
    if (atoi(e) & 8) {  // test aid: poison what the kernel must not rely on (huge scores, valid-looking ids / counts)
      LK_CUDA(cudaMemsetAsync(a.part_scores, 0x7F, n_part * sizeof(float), st));
      LK_CUDA(cudaMemsetAsync(a.part_idx, 0x00, n_part * sizeof(int32_t), st));
      if (a.part_cnt) LK_CUDA(cudaMemsetAsync(a.part_cnt, 0x01, n_cnt * sizeof(int), st));
    }
  }
  if (which == LK_KERNEL_UMMA) rc = launch_search_umma(a, ix->sm_count, st);
  else rc = launch_search_simt(a, ix->sm_count, st);
  if (rc != LK_OK) return rc;
  if (events) LK_CUDA(cudaEventRecord(ix->ev[2], st));

  if (dump_path) {
    std::vector<float> tile(kBlockRows * kBlockRows);
    LK_CUDA(cudaMemcpyAsync(tile.data(), ix->debug.p, tile.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    LK_CUDA(cudaStreamSynchronize(st));
    if (FILE* f = fopen(dump_path, "wb")) {
      fwrite(tile.data(), sizeof(float), tile.size(), f);
      fclose(f);
    }
  }

  // merge the per-CTA lists.  Sorted selectors leave their best k in the first k slots; the
  // append-buffer selector (ksel > kMaxK) leaves an unordered superset of its best k anywhere
  // in the slot
  return launch_merge_i32(a.part_scores, a.part_idx, a.part_cnt, b, a.n_lists, merge_len, a.ksel, k, idx_base, d_s,
                          d_i, st, mt_s, mt_i, kMergeTmpEntries);
}

// ------------------------------------------------------------------------------------
// k > 128: slab search (kernels in lk_deep.cu).  The reference ranks a materialised score
// matrix and so takes any k (retrieval/bruteforce.py:81-82).  Here the corpus is cut into slabs
// of whole 256-row units, about twice as many as lists of 128 would hold k; the fused kernel
// takes every slab's best 128 and a sorting merge folds them into the running result.  A slab
// with more than 128 rows whose 128th best still reaches some query's k-th score may hide
// better rows: it is split in two and searched again (the merge first drops what the result
// holds from those rows), down to single 128-row blocks, which are copied next to a padding
// block so that the unit-wide kernel sees nothing else.  Exact for any data; one pass over the
// corpus when the best rows are spread out, O(log) more over the crowded slabs when they are not.
// ------------------------------------------------------------------------------------
struct Slab {
  int64_t row0, rows;
  bool block;  // one 128-row block, searched through the scratch copy
};

static int deep_search(lk_index* ix, int which, const unsigned char* q_tiles, const float* q_side, int64_t b, int k,
                       int64_t idx_base, float* d_s, int64_t* d_i, cudaStream_t st) {
  constexpr int64_t kUnit = 2 * kBlockRows;
  int rc;
  const int64_t n = ix->n_rows, n_units = (n + kUnit - 1) / kUnit;
  const int64_t block_bytes = ix->sv_g.block_bytes();
  int64_t want = 2 * ((k + kMaxK - 1) / kMaxK);
  if (want > n_units) want = n_units;
  const int64_t upu = (n_units + want - 1) / want;  // units per slab
  std::vector<Slab> work, next;
  for (int64_t u = 0; u < n_units; u += upu) {
    const int64_t row0 = u * kUnit, rows = n - row0 < upu * kUnit ? n - row0 : upu * kUnit;
    work.push_back({row0, rows, false});
  }
  const int cap = deep_merge_capacity(k, 1, kMaxK);
  std::vector<int> flags;
  bool have_res = false;
  while (!work.empty()) {
    const int64_t n_round = (int64_t)work.size();
    const int64_t per = n_round < cap ? n_round : cap;  // slabs per merge launch
    if ((rc = ix->deep_s.ensure((size_t)per * b * kMaxK * sizeof(float))) != LK_OK) return rc;
    if ((rc = ix->deep_i.ensure((size_t)per * b * kMaxK * sizeof(int64_t))) != LK_OK) return rc;
    if ((rc = ix->deep_last.ensure((size_t)n_round * b * sizeof(float))) != LK_OK) return rc;
    if ((rc = ix->deep_flags.ensure((size_t)n_round * sizeof(int))) != LK_OK) return rc;
    LK_CUDA(cudaMemsetAsync(ix->deep_flags.p, 0, (size_t)n_round * sizeof(int), st));
    for (int64_t g0 = 0; g0 < n_round; g0 += per) {
      const int64_t g1 = g0 + per < n_round ? g0 + per : n_round;
      DeepLists L = {};
      L.n_lists = L.n_ranges = (int)(g1 - g0);
      L.len = kMaxK;
      L.list_stride = b * kMaxK;
      L.query_stride = kMaxK;
      for (int64_t g = g0; g < g1; ++g) {
        const Slab& s = work[(size_t)g];
        const unsigned char* tiles = ix->sv_tiles + (s.row0 / kBlockRows) * block_bytes;
        const float* side = ix->side + s.row0;
        if (s.block) {  // [the block][a padding block: zero rows, NaN side values]
          if (ix->deep_tiles.p == nullptr) {
            if ((rc = ix->deep_tiles.ensure((size_t)2 * block_bytes)) != LK_OK) return rc;
            if ((rc = ix->deep_side.ensure((size_t)kUnit * sizeof(float))) != LK_OK) return rc;
            LK_CUDA(cudaMemsetAsync(ix->deep_tiles.p, 0, (size_t)2 * block_bytes, st));
            LK_CUDA(cudaMemsetAsync(ix->deep_side.p, 0xFF, (size_t)kUnit * sizeof(float), st));
          }
          LK_CUDA(cudaMemcpyAsync(ix->deep_tiles.p, tiles, (size_t)block_bytes, cudaMemcpyDeviceToDevice, st));
          LK_CUDA(cudaMemcpyAsync(ix->deep_side.p, side, kBlockRows * sizeof(float), cudaMemcpyDeviceToDevice, st));
          tiles = ix->deep_tiles.as<unsigned char>();
          side = ix->deep_side.as<float>();
        }
        const SearchArgs a = rows_args(ix, tiles, side, s.rows, q_tiles, q_side, b, kMaxK);
        const size_t o = (size_t)(g - g0) * b * kMaxK;
        rc = search_rows(ix, which, a, idx_base + s.row0, false, nullptr, ix->deep_s.as<float>() + o,
                         ix->deep_i.as<int64_t>() + o, st);
        if (rc != LK_OK) return rc;
        L.lo[g - g0] = idx_base + s.row0;
        L.hi[g - g0] = idx_base + s.row0 + s.rows;
      }
      rc = launch_deep_merge(ix->deep_s.as<float>(), ix->deep_i.as<int64_t>(), L, b, k, have_res, d_s, d_i,
                             ix->deep_last.as<float>() + g0 * b, st);
      if (rc != LK_OK) return rc;
      have_res = true;
    }
    rc = launch_deep_saturated(ix->deep_last.as<float>(), (int)n_round, b, d_s, k, ix->deep_flags.as<int>(), st);
    if (rc != LK_OK) return rc;
    flags.resize((size_t)n_round);
    LK_CUDA(cudaMemcpyAsync(flags.data(), ix->deep_flags.p, (size_t)n_round * sizeof(int), cudaMemcpyDeviceToHost, st));
    LK_CUDA(cudaStreamSynchronize(st));
    next.clear();
    for (int64_t g = 0; g < n_round; ++g) {
      if (!flags[(size_t)g]) continue;
      const Slab& s = work[(size_t)g];
      const int64_t units = (s.rows + kUnit - 1) / kUnit;
      if (units >= 2) {
        const int64_t h = (units + 1) / 2 * kUnit;
        next.push_back({s.row0, h, false});
        next.push_back({s.row0 + h, s.rows - h, false});
      } else if (!s.block && s.rows > kBlockRows) {
        next.push_back({s.row0, kBlockRows, true});
        next.push_back({s.row0 + kBlockRows, s.rows - kBlockRows, true});
      }
    }
    work.swap(next);
  }
  return LK_OK;
}

int lk_index_search(lk_index* ix, const void* queries, int q_dtype, int q_mem, int64_t b, int k,
                    float* out_scores, int64_t* out_idx, int out_mem, int64_t idx_base, int kernel,
                    void* stream) {
  if (!ix || b < 0 || (b > 0 && (!queries || !out_scores || !out_idx)) ||
      (q_dtype != LK_F32 && q_dtype != LK_BF16) || (q_mem != LK_HOST && q_mem != LK_DEVICE) ||
      (out_mem != LK_HOST && out_mem != LK_DEVICE)) {
    set_error("lk_index_search: bad argument");
    return LK_ERR_INVALID;
  }
  if (k < 1 || k > kDeepMaxK) {
    set_error("lk_index_search: k=%d outside 1..%d", k, kDeepMaxK);
    return LK_ERR_INVALID;
  }
  if (ix->n_rows < 1) {
    set_error("lk_index_search: the index is empty");
    return LK_ERR_INVALID;
  }
  if (b == 0) return LK_OK;
  const bool deep = k > kMaxK;  // slab search: every fused search below asks for kMaxK
  const int which = pick_kernel(ix, kernel, deep ? kMaxK : k, b);
  if (which == LK_KERNEL_UMMA && !umma_supported(umma_geom(ix), deep ? kMaxK : k)) {
    set_error("lk_index_search: the tcgen05 kernel does not take this shape (dim=%d, k=%d, storage=%d)", ix->dim, k,
              ix->storage);
    return LK_ERR_UNSUPPORTED;
  }
  DeviceGuard guard(ix->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  // what the search kernels read: the tiles as stored, or -- fp32 storage on the tensor cores -- the planes
  const bool split = which == LK_KERNEL_UMMA && ix->storage == LK_F32;
  if (split) {
    if ((rc = ensure_planes(ix, st)) != LK_OK) return rc;
    ix->sv_tiles = ix->planes;
    ix->sv_g = ix->gp;
    ix->sv_split = ix->gp.kblocks / 2;
  } else {
    ix->sv_tiles = ix->tiles;
    ix->sv_g = ix->g;
    ix->sv_split = 0;
  }

  if (ix->timing) LK_CUDA(cudaEventRecord(ix->ev[0], st));

  // 0. A few queries over a small corpus (the reference's own caller issues ONE query per call,
  // main.py:270-271): everything in one launch -- query rounding and norms, search, merge of the
  // slices -- instead of tiling kernel + search + merge kernel.
  {
    const char* fe = getenv("LK_FUSED");
    const bool allow = kernel == LK_KERNEL_AUTO && !getenv("LK_FORCE_KERNEL") && !(fe && !atoi(fe));
    if (allow && !ix->whiten && ix->storage == LK_BF16 && simt_fused_supported(ix->g, ix->n_rows, b, k)) {
      SearchArgs a = {};
      a.tiles = ix->tiles;
      a.side = ix->side;
      a.n_rows = ix->n_rows;
      a.g = ix->g;
      a.n_queries = b;
      a.metric = ix->kmetric;
      a.k = k;
      a.idx_base = idx_base;
      a.ticket = ix->ticket;
      a.q_raw_dtype = q_dtype;
      // Host buffers go through a pinned staging block; the results are written into it by the
      // kernel itself (zero copy): one copy, a launch and one synchronisation instead of three copies.
      const size_t q_bytes = ((size_t)b * ix->dim * (q_dtype == LK_F32 ? 4 : 2) + 15) / 16 * 16;
      const size_t s_bytes = ((size_t)b * k * sizeof(float) + 15) / 16 * 16;
      const size_t i_bytes = (size_t)b * k * sizeof(int64_t);
      if (q_mem == LK_HOST || out_mem == LK_HOST) {
        const size_t need = q_bytes + s_bytes + i_bytes;
        if (need > ix->pin_cap) {
          if (ix->pin) cudaFreeHost(ix->pin);
          ix->pin = nullptr;
          ix->pin_cap = 0;
          LK_CUDA(cudaHostAlloc((void**)&ix->pin, need + 4096, cudaHostAllocMapped | cudaHostAllocPortable));
          ix->pin_cap = need + 4096;
        }
      }
      if (q_mem == LK_HOST) {  // every slice reads the queries: they go to device memory, through the pinned block
        const size_t qb = (size_t)b * ix->dim * (q_dtype == LK_F32 ? 4 : 2);
        if ((rc = ix->stage.ensure(qb)) != LK_OK) return rc;
        memcpy(ix->pin, queries, qb);
        LK_CUDA(cudaMemcpyAsync(ix->stage.p, ix->pin, qb, cudaMemcpyHostToDevice, st));
        a.q_raw = ix->stage.p;
      } else {
        a.q_raw = queries;
      }
      if ((rc = simt_plan(a, ix->sm_count, &a.n_lists, &a.ksel)) != LK_OK) return rc;
      const size_t n_part = (size_t)b * a.n_lists * a.ksel;  // every slice writes its whole lists: no clearing
      if ((rc = ix->part_s.ensure(n_part * sizeof(float))) != LK_OK) return rc;
      if ((rc = ix->part_i.ensure(n_part * sizeof(int32_t))) != LK_OK) return rc;
      a.part_scores = ix->part_s.as<float>();
      a.part_idx = ix->part_i.as<int32_t>();
      a.out_scores = out_scores;
      a.out_idx = out_idx;
      if (out_mem == LK_HOST) {
        a.out_scores = reinterpret_cast<float*>(ix->pin + q_bytes);
        a.out_idx = reinterpret_cast<int64_t*>(ix->pin + q_bytes + s_bytes);
      }
      if (ix->timing) LK_CUDA(cudaEventRecord(ix->ev[1], st));
      if ((rc = launch_search_simt(a, ix->sm_count, st)) != LK_OK) return rc;
      if (ix->timing) {
        LK_CUDA(cudaEventRecord(ix->ev[2], st));
        LK_CUDA(cudaEventRecord(ix->ev[3], st));
      }
      if (out_mem == LK_HOST) {
        LK_CUDA(cudaStreamSynchronize(st));
        memcpy(out_scores, a.out_scores, (size_t)b * k * sizeof(float));
        memcpy(out_idx, a.out_idx, (size_t)b * k * sizeof(int64_t));
      } else if (q_mem == LK_HOST) {
        LK_CUDA(cudaStreamSynchronize(st));  // the staged queries must outlive the kernel
      }
      if (ix->timing) {
        LK_CUDA(cudaEventSynchronize(ix->ev[3]));
        LK_CUDA(cudaEventElapsedTime(&ix->last_search_ms, ix->ev[1], ix->ev[2]));
        LK_CUDA(cudaEventElapsedTime(&ix->last_total_ms, ix->ev[0], ix->ev[3]));
      }
      return LK_OK;
    }
  }

  // 1. queries -> tiles (+ side values), same geometry and rounding as the corpus
  const int64_t b_pad = round_up64(b, 2 * kBlockRows);  // whole PAIRS of query tiles (CTA-pair kernel)
  const size_t qt_bytes = (size_t)(b_pad / kBlockRows) * ix->g.block_bytes();
  if ((rc = ix->q_tiles.ensure(qt_bytes)) != LK_OK) return rc;
  if ((rc = ix->q_side.ensure((size_t)b_pad * sizeof(float))) != LK_OK) return rc;
  // no clearing: padded query rows only feed accumulator lanes nobody reads, and the K padding of
  // the real rows is written by the tiling kernel
  if (const char* e = getenv("LK_DBG"))
    if (atoi(e) & 8) LK_CUDA(cudaMemsetAsync(ix->q_tiles.p, 0x7F, qt_bytes, st));  // test aid: huge garbage
  rc = ingest_rows(ix, queries, q_dtype, q_mem, b, ix->q_tiles.p, ix->q_side.as<float>(), 0, st);
  if (rc != LK_OK) return rc;
  const unsigned char* q_view = ix->q_tiles.as<unsigned char>();
  if (split) {  // the query tiles as planes, like the rows
    if ((rc = ix->q_planes.ensure((size_t)(b_pad / kBlockRows) * ix->gp.block_bytes())) != LK_OK) return rc;
    rc = launch_planes_from_tiles(ix->q_tiles.p, ix->g, 0, b_pad / kBlockRows, ix->gp, ix->q_planes.p, st);
    if (rc != LK_OK) return rc;
    q_view = ix->q_planes.as<unsigned char>();
  }

  // 2.-4. plan, fused distance + selection, merge of the partial lists
  const SearchArgs a = rows_args(ix, ix->sv_tiles, ix->side, ix->n_rows, q_view, ix->q_side.as<float>(), b, k);
  const char* dump_path = which == LK_KERNEL_UMMA && !deep ? getenv("LK_UMMA_DUMP") : nullptr;
  const int64_t seed_rows = which == LK_KERNEL_UMMA && !deep ? umma_seed_rows(a, ix->sm_count) : 0;
  // results land here: the caller's device buffers; for host outputs a device staging buffer, or --
  // small results that no later kernel reads back -- the pinned block, written by the merge kernel
  // itself (zero copy: two device-to-host copies less per call)
  float* d_s = out_scores;
  int64_t* d_i = out_idx;
  const size_t out_s_bytes = ((size_t)b * k * sizeof(float) + 15) / 16 * 16, out_i_bytes = (size_t)b * k * sizeof(int64_t);
  const bool zero_copy = out_mem == LK_HOST && !deep && seed_rows == 0 && out_s_bytes + out_i_bytes <= (64u << 10);
  if (zero_copy) {
    if (out_s_bytes + out_i_bytes > ix->pin_cap) {
      if (ix->pin) cudaFreeHost(ix->pin);
      ix->pin = nullptr;
      ix->pin_cap = 0;
      LK_CUDA(cudaHostAlloc((void**)&ix->pin, (64u << 10) + 4096, cudaHostAllocMapped | cudaHostAllocPortable));
      ix->pin_cap = (64u << 10) + 4096;
    }
    d_s = reinterpret_cast<float*>(ix->pin);
    d_i = reinterpret_cast<int64_t*>(ix->pin + out_s_bytes);
  } else if (out_mem == LK_HOST) {
    if ((rc = ix->out_s.ensure((size_t)b * k * sizeof(float))) != LK_OK) return rc;
    if ((rc = ix->out_i.ensure((size_t)b * k * sizeof(int64_t))) != LK_OK) return rc;
    d_s = ix->out_s.as<float>();
    d_i = ix->out_i.as<int64_t>();
  }
  if (!deep) {
    if ((rc = search_rows(ix, which, a, idx_base, ix->timing, dump_path, d_s, d_i, st)) != LK_OK) return rc;
  } else {
    // a chunk of queries at a time: every slab of a merge launch keeps 128 candidates per query
    constexpr int64_t kDeepChunk = 512;
    if (ix->timing) LK_CUDA(cudaEventRecord(ix->ev[1], st));
    for (int64_t q0 = 0; q0 < b; q0 += kDeepChunk) {
      const int64_t bc = b - q0 < kDeepChunk ? b - q0 : kDeepChunk;
      rc = deep_search(ix, which, q_view + (q0 / kBlockRows) * ix->sv_g.block_bytes(),
                       ix->q_side.as<float>() + q0, bc, k, idx_base, d_s + q0 * k, d_i + q0 * k, st);
      if (rc != LK_OK) return rc;
    }
    if (ix->timing) LK_CUDA(cudaEventRecord(ix->ev[2], st));
  }
  if (ix->timing) LK_CUDA(cudaEventRecord(ix->ev[3], st));

  // 5. results (and the kernels' error flag) back to the host
  if (out_mem == LK_HOST) {
    int flag = 0;
    if (!zero_copy) {
      LK_CUDA(cudaMemcpyAsync(out_scores, d_s, (size_t)b * k * sizeof(float), cudaMemcpyDeviceToHost, st));
      LK_CUDA(cudaMemcpyAsync(out_idx, d_i, (size_t)b * k * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    }
    LK_CUDA(cudaMemcpyAsync(&flag, ix->err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    LK_CUDA(cudaStreamSynchronize(st));
    if (zero_copy) {
      memcpy(out_scores, d_s, (size_t)b * k * sizeof(float));
      memcpy(out_idx, d_i, (size_t)b * k * sizeof(int64_t));
    }
    if (flag != 0) {
      cudaMemsetAsync(ix->err_flag, 0, sizeof(int), st);
      set_error("search kernel pipeline timed out (barrier code %d); results are invalid", flag);
      return LK_ERR_CUDA;
    }
  }
  if (ix->timing) {
    LK_CUDA(cudaEventSynchronize(ix->ev[3]));
    LK_CUDA(cudaEventElapsedTime(&ix->last_search_ms, ix->ev[1], ix->ev[2]));
    LK_CUDA(cudaEventElapsedTime(&ix->last_total_ms, ix->ev[0], ix->ev[3]));
    if (out_mem == LK_DEVICE) {
      int flag = 0;
      LK_CUDA(cudaMemcpy(&flag, ix->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
      if (flag != 0) {
        cudaMemset(ix->err_flag, 0, sizeof(int));
        set_error("search kernel pipeline timed out (barrier code %d); results are invalid", flag);
        return LK_ERR_CUDA;
      }
    }
  }
  return LK_OK;
}

// ------------------------------------------------------------------------------------
// merge of per-shard candidates
// ------------------------------------------------------------------------------------
int lk_merge_topk(int device, const float* cand_scores, const int64_t* cand_idx, int64_t b, int n_lists,
                  int list_len, int k, float* out_scores, int64_t* out_idx, int mem, void* stream) {
  if (b < 0 || n_lists < 1 || list_len < 1 || k < 1 || k > kDeepMaxK || (mem != LK_HOST && mem != LK_DEVICE) ||
      (b > 0 && (!cand_scores || !cand_idx || !out_scores || !out_idx))) {
    set_error("lk_merge_topk: bad argument");
    return LK_ERR_INVALID;
  }
  if (k > kMaxK && (int64_t)n_lists * list_len > kDeepMaxCand) {
    set_error("lk_merge_topk: k=%d > %d takes at most %d candidates per query (%d lists of %d given)", k, kMaxK,
              kDeepMaxCand, n_lists, list_len);
    return LK_ERR_UNSUPPORTED;
  }
  if (b == 0) return LK_OK;
  // up to kMaxK: insertion / selection merge; above: the sorting merge of the slab search
  auto merge = [&](const float* cs, const int64_t* ci, float* os, int64_t* oi, cudaStream_t st) {
    if (k <= kMaxK) return launch_merge_i64(cs, ci, b, n_lists, list_len, k, os, oi, st);
    DeepLists L = {};
    L.n_lists = n_lists;
    L.len = list_len;
    L.list_stride = list_len;
    L.query_stride = (int64_t)n_lists * list_len;
    return launch_deep_merge(cs, ci, L, b, k, 0, os, oi, nullptr, st);
  };
  int sm = 0;
  int rc = check_device(device, &sm);
  if (rc != LK_OK) return rc;
  DeviceGuard guard(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mem == LK_DEVICE) return merge(cand_scores, cand_idx, out_scores, out_idx, st);
  const size_t nc = (size_t)b * n_lists * list_len, no = (size_t)b * k;
  Buf cs, ci, os, oi;
  auto done = [&](int code) {
    cs.release(); ci.release(); os.release(); oi.release();
    return code;
  };
  if ((rc = cs.ensure(nc * 4)) != LK_OK || (rc = ci.ensure(nc * 8)) != LK_OK ||
      (rc = os.ensure(no * 4)) != LK_OK || (rc = oi.ensure(no * 8)) != LK_OK)
    return done(rc);
  cudaError_t e;
  if ((e = cudaMemcpyAsync(cs.p, cand_scores, nc * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
      (e = cudaMemcpyAsync(ci.p, cand_idx, nc * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess)
    return done(cuda_fail(e, "cudaMemcpyAsync(H2D)", __FILE__, __LINE__));
  rc = merge(cs.as<float>(), ci.as<int64_t>(), os.as<float>(), oi.as<int64_t>(), st);
  if (rc != LK_OK) return done(rc);
  if ((e = cudaMemcpyAsync(out_scores, os.p, no * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
      (e = cudaMemcpyAsync(out_idx, oi.p, no * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
      (e = cudaStreamSynchronize(st)) != cudaSuccess)
    return done(cuda_fail(e, "merge D2H", __FILE__, __LINE__));
  return done(LK_OK);
}

// ------------------------------------------------------------------------------------
// autoencoder encoder
// ------------------------------------------------------------------------------------
int lk_ae_destroy(lk_ae* ae) {
  if (!ae) return LK_OK;
  DeviceGuard guard(ae->device);
  float* ps[] = {ae->w0t, ae->b0, ae->w1t, ae->b1};
  for (float* p : ps)
    if (p) cudaFree(p);
  if (ae->w0_slabs) cudaFree(ae->w0_slabs);
  if (ae->w1_slabs) cudaFree(ae->w1_slabs);
  if (ae->err_flag) cudaFree(ae->err_flag);
  for (int i = 0; i < 2; ++i) {
    ae->xin[i].release();
    ae->zout[i].release();
    for (cudaEvent_t e : {ae->ev_in[i], ae->ev_comp[i], ae->ev_out[i]})
      if (e) cudaEventDestroy(e);
  }
  ae->up.release();
  if (ae->cs_in) cudaStreamDestroy(ae->cs_in);
  if (ae->cs_out) cudaStreamDestroy(ae->cs_out);
  ae->xslabs.release();
  delete ae;
  return LK_OK;
}

int lk_ae_create(lk_ae** out, int device, int kind, int d_in, int d_hidden, int d_latent, const float* w0,
                 const float* b0, const float* w1, const float* b1) {
  if (!out) return LK_ERR_INVALID;
  *out = nullptr;
  if (kind < LK_AE_DAE || kind > LK_AE_VAE_MU || d_in < 1 || d_hidden < 1 || d_latent < 1 || !w0 || !b0 ||
      !w1 || !b1) {
    set_error("lk_ae_create: bad argument");
    return LK_ERR_INVALID;
  }
  int sm = 0;
  int rc = check_device(device, &sm);
  if (rc != LK_OK) return rc;
  DeviceGuard guard(device);
  lk_ae* ae = new (std::nothrow) lk_ae();
  if (!ae) return LK_ERR_OOM;
  ae->device = device;
  ae->sm_count = sm;
  ae->kind = kind;
  ae->d_in = d_in;
  ae->d_hidden = d_hidden;
  ae->d_latent = d_latent;
  // nn.Linear keeps [out, in]; the kernel wants [in, out] so that threads of a warp read
  // consecutive output columns
  std::vector<float> w0t((size_t)d_in * d_hidden), w1t((size_t)d_hidden * d_latent);
  for (int o = 0; o < d_hidden; ++o)
    for (int i = 0; i < d_in; ++i) w0t[(size_t)i * d_hidden + o] = w0[(size_t)o * d_in + i];
  for (int o = 0; o < d_latent; ++o)
    for (int i = 0; i < d_hidden; ++i) w1t[(size_t)i * d_latent + o] = w1[(size_t)o * d_hidden + i];
  struct Up { float** dst; const float* src; size_t n; } ups[] = {
      {&ae->w0t, w0t.data(), w0t.size()}, {&ae->b0, b0, (size_t)d_hidden},
      {&ae->w1t, w1t.data(), w1t.size()}, {&ae->b1, b1, (size_t)d_latent}};
  for (auto& u : ups) {
    cudaError_t e = cudaMalloc((void**)u.dst, u.n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(*u.dst, u.src, u.n * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      rc = cuda_fail(e, "ae weight upload", __FILE__, __LINE__);
      lk_ae_destroy(ae);
      return rc;
    }
  }
  if (ae_umma_supported(d_in, d_hidden, d_latent)) {
    // split-bf16 weight planes in the tcgen05 slab format
    std::vector<unsigned char> s0, s1;
    ae_umma_weight_slabs(w0, d_hidden, d_in, kBlockRows, &s0, false);
    ae_umma_weight_slabs(w1, d_latent, d_hidden, round_up(d_latent, 16), &s1, true);
    cudaError_t e = cudaMalloc((void**)&ae->w0_slabs, s0.size());
    if (e == cudaSuccess) e = cudaMalloc((void**)&ae->w1_slabs, s1.size());
    if (e == cudaSuccess) e = cudaMalloc((void**)&ae->err_flag, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(ae->w0_slabs, s0.data(), s0.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(ae->w1_slabs, s1.data(), s1.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(ae->err_flag, 0, sizeof(int));
    if (e != cudaSuccess) {
      rc = cuda_fail(e, "ae weight slab upload", __FILE__, __LINE__);
      lk_ae_destroy(ae);
      return rc;
    }
  }
  *out = ae;
  return LK_OK;
}

int lk_ae_set_kernel(lk_ae* ae, int kernel) {
  if (!ae || kernel < LK_KERNEL_AUTO || kernel > LK_KERNEL_UMMA) return LK_ERR_INVALID;
  if (kernel == LK_KERNEL_UMMA && !ae->w0_slabs) {
    set_error("lk_ae_set_kernel: the tensor-core encoder needs d_in %% 64 == 0, d_hidden %% 128 == 0, d_latent <= 64");
    return LK_ERR_UNSUPPORTED;
  }
  ae->kernel = kernel;
  return LK_OK;
}

int lk_ae_set_precision(lk_ae* ae, int precision) {
  if (!ae || (precision != LK_F32 && precision != LK_BF16)) return LK_ERR_INVALID;
  if (precision == LK_BF16 && !ae->w0_slabs) {
    set_error("lk_ae_set_precision: bf16 operands need the tensor-core encoder (d_in %% 64 == 0, d_hidden %% 128 == 0, d_latent <= 64)");
    return LK_ERR_UNSUPPORTED;
  }
  ae->precision = precision;
  return LK_OK;
}

// one chunk of rows, device to device, on `st`
static int ae_encode_rows(lk_ae* ae, const float* xin, int64_t cnt, float* zdev, bool use_umma, cudaStream_t st) {
  const int l2 = ae->kind == LK_AE_CAE ? 1 : 0;
  const char* pe = getenv("LK_AE_PAIR");  // bring-up override: 0 = the single-CTA kernel for bf16 operands too
  const bool use_pair = use_umma && ae->precision == LK_BF16 && !(pe && !atoi(pe)) &&
                        ae_pair_supported(ae->d_in, ae->d_hidden, ae->d_latent, ae->sm_count) &&
                        (reinterpret_cast<uintptr_t>(xin) & 31u) == 0;
  if (use_pair)  // bf16 operands: CTA pairs, the fp32 rows are rounded inside the kernel (no split pass)
    return launch_ae_pair(xin, cnt, ae->d_in, ae->d_hidden, ae->d_latent, ae->w0_slabs, ae->w1_slabs, ae->b0, ae->b1, l2,
                          zdev, ae->err_flag, ae->sm_count, st);
  if (use_umma) {
    int rc = ae->xslabs.ensure(ae_umma_x_slab_bytes(cnt, ae->d_in));
    if (rc != LK_OK) return rc;
    const int planes = ae->precision == LK_BF16 ? 1 : 2;
    if ((rc = launch_ae_split_rows(xin, cnt, ae->d_in, planes, ae->xslabs.as<unsigned char>(), st)) != LK_OK) return rc;
    return launch_ae_umma(ae->xslabs.as<unsigned char>(), cnt, ae->d_in, ae->d_hidden, ae->d_latent, ae->w0_slabs,
                          ae->w1_slabs, ae->b0, ae->b1, l2, planes, zdev, ae->err_flag, ae->sm_count, st);
  }
  return launch_ae_encode(xin, cnt, ae->d_in, ae->d_hidden, ae->d_latent, ae->w0t, ae->b0, ae->w1t, ae->b1, l2, zdev, st);
}

static int ae_check_flag(lk_ae* ae) {
  if (!ae->err_flag) return LK_OK;
  int flag = 0;
  LK_CUDA(cudaMemcpy(&flag, ae->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag != 0) {
    cudaMemset(ae->err_flag, 0, sizeof(int));
    set_error("autoencoder kernel pipeline timed out (barrier code %d); results are invalid", flag);
    return LK_ERR_CUDA;
  }
  return LK_OK;
}

int lk_ae_encode(lk_ae* ae, const float* x, int x_mem, int64_t m, float* z, int z_mem, void* stream) {
  if (!ae || m < 0 || (m > 0 && (!x || !z)) || (x_mem != LK_HOST && x_mem != LK_DEVICE) ||
      (z_mem != LK_HOST && z_mem != LK_DEVICE)) {
    set_error("lk_ae_encode: bad argument");
    return LK_ERR_INVALID;
  }
  if (m == 0) return LK_OK;
  DeviceGuard guard(ae->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  // AUTO: tensor cores once there are two full row tiles of work, fp32 FMA for small batches
  const char* env = getenv("LK_AE_KERNEL");
  int which = ae->kernel;
  if (env && !strcmp(env, "simt")) which = LK_KERNEL_SIMT;
  if (env && !strcmp(env, "umma")) which = LK_KERNEL_UMMA;
  const bool use_umma = ae->w0_slabs && (which == LK_KERNEL_UMMA || ae->precision == LK_BF16 ||
                                         (which == LK_KERNEL_AUTO && m >= 2 * kBlockRows));
  const bool host_in = x_mem == LK_HOST, host_out = z_mem == LK_HOST;
  if (!host_in && !host_out) {  // device to device: chunks only bound the split-plane scratch
    const int64_t step = 1 << 18;
    for (int64_t done = 0; done < m; done += step) {
      const int64_t cnt = m - done < step ? m - done : step;
      if ((rc = ae_encode_rows(ae, x + (size_t)done * ae->d_in, cnt, z + (size_t)done * ae->d_latent, use_umma, st)) != LK_OK)
        return rc;
    }
    return LK_OK;
  }

  // Host rows and / or host latents: chunks of 32 Ki rows (48 MiB of fp32 input at 384 dims) through two staging
  // buffers each way.  H2D of chunk i + 1 (cs_in) and D2H of chunk i - 1 (cs_out) run under the kernel of chunk i
  // (the caller's stream); a buffer is reused two chunks later, behind the event of its previous consumer.
  const int64_t step = 1 << 15;
  if (!ae->cs_in) {
    LK_CUDA(cudaStreamCreateWithFlags(&ae->cs_in, cudaStreamNonBlocking));
    LK_CUDA(cudaStreamCreateWithFlags(&ae->cs_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      LK_CUDA(cudaEventCreateWithFlags(&ae->ev_in[i], cudaEventDisableTiming));
      LK_CUDA(cudaEventCreateWithFlags(&ae->ev_comp[i], cudaEventDisableTiming));
      LK_CUDA(cudaEventCreateWithFlags(&ae->ev_out[i], cudaEventDisableTiming));
    }
  }
  const int64_t n_chunks = (m + step - 1) / step;
  const int64_t rows_buf = m < step ? m : step;
  for (int i = 0; i < (n_chunks > 1 ? 2 : 1); ++i) {
    if (host_in && (rc = ae->xin[i].ensure((size_t)rows_buf * ae->d_in * 4)) != LK_OK) return rc;
    if (host_out && (rc = ae->zout[i].ensure((size_t)rows_buf * ae->d_latent * 4)) != LK_OK) return rc;
  }
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int buf = (int)(c & 1);
    const int64_t done = c * step, cnt = m - done < step ? m - done : step;
    const float* xin = x + (size_t)done * ae->d_in;
    float* zo = z + (size_t)done * ae->d_latent;
    if (host_in) {
      if (c >= 2) LK_CUDA(cudaStreamWaitEvent(ae->cs_in, ae->ev_comp[buf], 0));  // the kernel of chunk c - 2 read xin[buf]
      if ((rc = ae->up.upload(ae->xin[buf].p, xin, (size_t)cnt * ae->d_in * 4, ae->cs_in)) != LK_OK) return rc;
      LK_CUDA(cudaEventRecord(ae->ev_in[buf], ae->cs_in));
      LK_CUDA(cudaStreamWaitEvent(st, ae->ev_in[buf], 0));
      xin = ae->xin[buf].as<float>();
    }
    float* zdev = zo;
    if (host_out) {
      if (c >= 2) LK_CUDA(cudaStreamWaitEvent(st, ae->ev_out[buf], 0));  // zout[buf] of chunk c - 2 is on the host
      zdev = ae->zout[buf].as<float>();
    }
    if ((rc = ae_encode_rows(ae, xin, cnt, zdev, use_umma, st)) != LK_OK) {
      cudaStreamSynchronize(ae->cs_in);
      cudaStreamSynchronize(ae->cs_out);
      cudaStreamSynchronize(st);
      return rc;
    }
    LK_CUDA(cudaEventRecord(ae->ev_comp[buf], st));
    if (host_out) {
      LK_CUDA(cudaStreamWaitEvent(ae->cs_out, ae->ev_comp[buf], 0));
      LK_CUDA(cudaMemcpyAsync(zo, zdev, (size_t)cnt * ae->d_latent * 4, cudaMemcpyDeviceToHost, ae->cs_out));
      LK_CUDA(cudaEventRecord(ae->ev_out[buf], ae->cs_out));
    }
  }
  // the caller may free x and read z on return
  LK_CUDA(cudaStreamSynchronize(st));
  if (host_out) LK_CUDA(cudaStreamSynchronize(ae->cs_out));
  return ae_check_flag(ae);
}

}  // extern "C"
