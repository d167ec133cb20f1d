// Candidate exchange between the GPUs of a row-sharded index, fused with the final k-way
// merge, over NVLink peer memory.  Net-new relative to the reference, which is one process
// (SURVEY.md section 8e).
//
// Every rank owns a SYMMETRIC buffer (same layout on every GPU, peers mapped through CUDA
// IPC):   cand[2][world][max_b][max_k] scores + ids,   flag[2][world][max_b] epochs.
// One kernel per search call on every rank:
//   phase 1  publish: the CTA that owns query q stores this rank's k candidates of q into slot
//            [epoch & 1][rank][q] of EVERY rank's buffer (plain st.global on peer addresses:
//            k * 12 bytes per peer), then releases flag[..][rank][q] = epoch on every rank
//            (st.release.sys after a system-scope fence);
//   phase 2  merge: the same CTAs wait until all `world` flags of q carry this epoch
//            (ld.acquire.sys, bounded spin) and merge the world * k candidates, which by then
//            sit in LOCAL memory.
// The grid is sized to be fully resident and every CTA finishes publishing all its queries
// before it waits for anything, so no rank can wait on a CTA that has not been scheduled.
// Two slots alternate by epoch parity: a rank can only reach call e + 2 after every rank
// has published call e + 1, i.e. after every rank has finished reading call e.
// Against an all-gather this removes two collectives, their launch gaps and the staging
// copies from the batch-1 critical path; the payload is tiny (world * k * 12 bytes per
// query), so the exchange is latency- not bandwidth-bound.
#include <cstdlib>
#include <cstring>
#include <new>

#include "lk_topk.cuh"

namespace lk {

namespace {

constexpr int kXThreads = 256;
// A rank may legitimately arrive late (a cudaMalloc in lk_index_reserve, a first-call
// cudaFuncSetAttribute, a crowded shard, a host-side pause): the wait for a peer's flag is bounded
// only to turn a dead peer into an error instead of a hang.  LK_XCHG_TIMEOUT_S overrides the default.
constexpr double kDefaultTimeoutS = 60.0;
constexpr double kSpinClockHz = 2.0e9;

struct XView {  // one rank's symmetric buffer
  float* scores;     // [2][world][max_b][max_k]
  int64_t* idx;      // same shape
  unsigned* flag;    // [2][world][max_b]
};

struct XParams {
  XView peer[LK_MAX_WORLD];  // peer[r] = rank r's buffer as mapped into this process (peer[rank] = own)
  int rank, world;
  int64_t max_b;
  int max_k;
  unsigned epoch;
  const float* local_s;      // [b, k] this rank's candidates (global ids)
  const int64_t* local_i;
  int64_t b;
  int k;
  float* out_s;              // [b, k]
  int64_t* out_i;
  int* err_flag;
  int phases;                // bit 0: publish, bit 1: wait + merge
  long long spin_cycles;     // bound of the wait for one peer flag
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kXMaxCand = LK_MAX_WORLD * kMaxK;  // candidates of one query over all ranks

__global__ void __launch_bounds__(kXThreads) exchange_merge_kernel(const XParams p) {
  __shared__ float cs[kXMaxCand];      // the world sorted lists of one query, rank-major
  __shared__ int64_t ci[kXMaxCand];
  __shared__ int s_timeout;
  const int tid = threadIdx.x;
  const int slot = (int)(p.epoch & 1u);
  const int64_t slot_rows = (int64_t)p.world * p.max_b;  // rows of one slot
  if (p.phases & 1) {
    for (int64_t q = blockIdx.x; q < p.b; q += gridDim.x) {
      const int64_t row = ((int64_t)slot * slot_rows + (int64_t)p.rank * p.max_b + q);
      for (int e = tid; e < p.world * p.k; e += kXThreads) {
        const int r = e / p.k, j = e - r * p.k;
        p.peer[r].scores[row * p.max_k + j] = p.local_s[q * p.k + j];
        p.peer[r].idx[row * p.max_k + j] = p.local_i[q * p.k + j];
      }
      __syncthreads();
      if (tid < p.world) {
        __threadfence_system();  // the candidate stores of the whole CTA are visible system-wide first
        st_release_sys(p.peer[tid].flag + row, p.epoch);
      }
    }
  }
  if (!(p.phases & 2)) return;
  const XView me = p.peer[p.rank];
  for (int64_t q = blockIdx.x; q < p.b; q += gridDim.x) {
    if (tid == 0) s_timeout = 0;
    __syncthreads();
    if (tid < p.world) {
      const unsigned* f = me.flag + (int64_t)slot * slot_rows + (int64_t)tid * p.max_b + q;
      const long long t0 = clock64();
      while (ld_acquire_sys(f) != p.epoch) {
        if (clock64() - t0 > p.spin_cycles) {
          s_timeout = 1;
          break;
        }
      }
    }
    __syncthreads();
    if (s_timeout) {  // a peer never published: this and the remaining queries come back empty, never uninitialised
      if (tid == 0) atomicCAS(p.err_flag, 0, 301);
      for (int64_t qq = q; qq < p.b; qq += gridDim.x)
        for (int j = tid; j < p.k; j += kXThreads) {
          p.out_s[qq * p.k + j] = -INFINITY;
          p.out_i[qq * p.k + j] = -1;
        }
      return;
    }
    // Merge by RANKING, all threads: every rank's list arrives sorted under the total order (score desc, id asc),
    // so the final position of candidate j of list r is j + (for every other list) the number of its entries
    // that are better -- one binary search per other list.  world * k * (world - 1) * log2(k) shared-memory
    // probes (39 k for 8 ranks x 100) spread over the CTA, instead of up to world * k serial insertions into a
    // sorted list by one warp (which made the top-100 exchange of 4096 queries cost 1.3 ms on 8 GPUs).
    const int n = p.world * p.k;
    for (int e = tid; e < n; e += kXThreads) {
      const int r = e / p.k, j = e - r * p.k;
      const int64_t row = (int64_t)slot * slot_rows + (int64_t)r * p.max_b + q;
      float v = __ldcg(me.scores + row * p.max_k + j);  // written by a peer: not through L1
      int64_t id = __ldcg(me.idx + row * p.max_k + j);
      if (id < 0 || !(v == v)) {  // padding of a short list: worse than everything
        v = -INFINITY;
        id = INT64_MAX;
      }
      cs[e] = v;
      ci[e] = id;
    }
    for (int j = tid; j < p.k; j += kXThreads) {  // slots no candidate lands in (fewer than k rows in total)
      p.out_s[q * p.k + j] = -INFINITY;
      p.out_i[q * p.k + j] = -1;
    }
    __syncthreads();
    for (int e = tid; e < n; e += kXThreads) {
      const int r = e / p.k, j = e - r * p.k;
      const float v = cs[e];
      const int64_t id = ci[e];
      if (id == INT64_MAX) continue;
      int pos = j;
      for (int r2 = 0; r2 < p.world && pos < p.k; ++r2) {
        if (r2 == r) continue;
        const float* ls2 = cs + r2 * p.k;
        const int64_t* li2 = ci + r2 * p.k;
        int lo = 0, hi = p.k;  // first entry of list r2 that is NOT better than this candidate
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (better(ls2[mid], li2[mid], v, id)) lo = mid + 1;
          else hi = mid;
        }
        pos += lo;
      }
      if (pos < p.k) {
        p.out_s[q * p.k + pos] = v;
        p.out_i[q * p.k + pos] = id;
      }
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace lk

using namespace lk;

struct lk_comm {
  int device = 0, rank = 0, world = 1, sm_count = 0;
  int64_t max_b = 0;
  int max_k = 0;
  unsigned epoch = 0;
  void* base = nullptr;               // this rank's buffer (cudaMalloc)
  void* mapped[LK_MAX_WORLD] = {};    // every rank's buffer in this process' address space
  bool opened[LK_MAX_WORLD] = {};     // mapped through cudaIpcOpenMemHandle (to be closed)
  bool ready = false;
  int* err_flag = nullptr;
  size_t bytes = 0;
};

namespace {

size_t off_idx(const lk_comm* c) { return (size_t)2 * c->world * c->max_b * c->max_k * sizeof(float); }
size_t off_flag(const lk_comm* c) { return off_idx(c) + (size_t)2 * c->world * c->max_b * c->max_k * sizeof(int64_t); }
size_t total_bytes(const lk_comm* c) { return off_flag(c) + (size_t)2 * c->world * c->max_b * sizeof(unsigned); }

XView view_of(const lk_comm* c, void* base) {
  unsigned char* b = static_cast<unsigned char*>(base);
  XView v;
  v.scores = reinterpret_cast<float*>(b);
  v.idx = reinterpret_cast<int64_t*>(b + off_idx(c));
  v.flag = reinterpret_cast<unsigned*>(b + off_flag(c));
  return v;
}

struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int run_exchange(lk_comm* c, const float* local_s, const int64_t* local_i, int64_t b, int k, float* out_s,
                 int64_t* out_i, int phases, cudaStream_t st) {
  XParams p;
  for (int r = 0; r < c->world; ++r) p.peer[r] = view_of(c, c->mapped[r]);
  p.rank = c->rank;
  p.world = c->world;
  p.max_b = c->max_b;
  p.max_k = c->max_k;
  p.epoch = c->epoch;
  p.local_s = local_s;
  p.local_i = local_i;
  p.b = b;
  p.k = k;
  p.out_s = out_s;
  p.out_i = out_i;
  p.err_flag = c->err_flag;
  p.phases = phases;
  double timeout_s = kDefaultTimeoutS;
  if (const char* e = getenv("LK_XCHG_TIMEOUT_S")) {
    const double v = atof(e);
    if (v > 0.0) timeout_s = v;
  }
  p.spin_cycles = (long long)(timeout_s * kSpinClockHz);
  const int64_t resident = (int64_t)c->sm_count * 4;  // 256 threads, ~3 KB smem: at least 4 CTAs per SM fit
  const unsigned grid = (unsigned)(b < resident ? b : resident);
  exchange_merge_kernel<<<grid, kXThreads, 0, st>>>(p);
  LK_CHECK_LAUNCH("exchange_merge_kernel");
  return LK_OK;
}

int check_call(const lk_comm* c, const void* a, const void* b2, int64_t b, int k) {
  if (!c || !c->ready) {
    set_error("lk_comm: peers are not opened yet");
    return LK_ERR_INVALID;
  }
  if (b < 0 || b > c->max_b || k < 1 || k > c->max_k || (b > 0 && (!a || !b2))) {
    set_error("lk_comm: bad argument (b=%lld of %lld, k=%d of %d)", (long long)b, (long long)c->max_b, k, c->max_k);
    return LK_ERR_INVALID;
  }
  return LK_OK;
}

}  // namespace

extern "C" {

int lk_comm_create(lk_comm** out, int device, int rank, int world, int64_t max_b, int max_k) {
  if (!out || world < 1 || world > LK_MAX_WORLD || rank < 0 || rank >= world || max_b < 1 || max_k < 1 ||
      max_k > kMaxK) {
    set_error("lk_comm_create: bad argument");
    return LK_ERR_INVALID;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
  if (device < 0 || device >= n) {
    set_error("device %d out of range (%d visible)", device, n);
    return LK_ERR_INVALID;
  }
  lk_comm* c = new (std::nothrow) lk_comm();
  if (!c) return LK_ERR_OOM;
  c->device = device;
  c->rank = rank;
  c->world = world;
  c->max_b = max_b;
  c->max_k = max_k;
  DevGuard guard(device);
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  c->bytes = total_bytes(c);
  e = cudaMalloc(&c->base, c->bytes);
  if (e == cudaSuccess) e = cudaMemset(c->base, 0, c->bytes);  // epoch 0 = nothing published
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->err_flag, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(c->err_flag, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    int rc = cuda_fail(e, "lk_comm_create allocation", __FILE__, __LINE__);
    lk_comm_destroy(c);
    return rc;
  }
  c->mapped[rank] = c->base;
  c->ready = world == 1;
  *out = c;
  return LK_OK;
}

int lk_comm_ipc_handle(lk_comm* c, void* out_handle64) {
  if (!c || !out_handle64) return LK_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == LK_IPC_HANDLE_BYTES, "handle size");
  DevGuard guard(c->device);
  cudaIpcMemHandle_t h;
  LK_CUDA(cudaIpcGetMemHandle(&h, c->base));
  memcpy(out_handle64, &h, sizeof(h));
  return LK_OK;
}

int lk_comm_open_peers(lk_comm* c, const void* handles) {
  if (!c || !handles) return LK_ERR_INVALID;
  DevGuard guard(c->device);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank || c->mapped[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * LK_IPC_HANDLE_BYTES, sizeof(h));
    LK_CUDA(cudaIpcOpenMemHandle(&c->mapped[r], h, cudaIpcMemLazyEnablePeerAccess));
    c->opened[r] = true;
  }
  c->ready = true;
  return LK_OK;
}

int lk_comm_attach_local(lk_comm* c, int peer_rank, lk_comm* peer) {
  if (!c || !peer || peer_rank < 0 || peer_rank >= c->world || peer->world != c->world ||
      peer->max_b != c->max_b || peer->max_k != c->max_k || peer->rank != peer_rank) {
    set_error("lk_comm_attach_local: the communicators do not match");
    return LK_ERR_INVALID;
  }
  if (peer->device != c->device) {
    int can = 0;
    LK_CUDA(cudaDeviceCanAccessPeer(&can, c->device, peer->device));
    if (!can) {
      set_error("lk_comm_attach_local: device %d cannot access device %d", c->device, peer->device);
      return LK_ERR_UNSUPPORTED;
    }
    DevGuard guard(c->device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
    cudaGetLastError();
  }
  c->mapped[peer_rank] = peer->base;
  bool all = true;
  for (int r = 0; r < c->world; ++r) all = all && c->mapped[r] != nullptr;
  c->ready = all;
  return LK_OK;
}

int lk_comm_begin(lk_comm* c) {
  if (!c || !c->ready) {
    set_error("lk_comm: peers are not opened yet");
    return LK_ERR_INVALID;
  }
  ++c->epoch;
  if (c->epoch == 0) ++c->epoch;  // 0 means "nothing published"
  return LK_OK;
}

int lk_comm_publish(lk_comm* c, const float* local_scores, const int64_t* local_idx, int64_t b, int k,
                    void* stream) {
  int rc = check_call(c, local_scores, local_idx, b, k);
  if (rc != LK_OK || b == 0) return rc;
  DevGuard guard(c->device);
  return run_exchange(c, local_scores, local_idx, b, k, nullptr, nullptr, 1, static_cast<cudaStream_t>(stream));
}

int lk_comm_collect(lk_comm* c, int64_t b, int k, float* out_scores, int64_t* out_idx, void* stream) {
  int rc = check_call(c, out_scores, out_idx, b, k);
  if (rc != LK_OK || b == 0) return rc;
  DevGuard guard(c->device);
  return run_exchange(c, nullptr, nullptr, b, k, out_scores, out_idx, 2, static_cast<cudaStream_t>(stream));
}

int lk_comm_exchange_merge(lk_comm* c, const float* local_scores, const int64_t* local_idx, int64_t b, int k,
                           float* out_scores, int64_t* out_idx, void* stream) {
  int rc = check_call(c, local_scores, local_idx, b, k);
  if (rc != LK_OK) return rc;
  if (b > 0 && (!out_scores || !out_idx)) return LK_ERR_INVALID;
  if ((rc = lk_comm_begin(c)) != LK_OK || b == 0) return rc;
  DevGuard guard(c->device);
  return run_exchange(c, local_scores, local_idx, b, k, out_scores, out_idx, 3, static_cast<cudaStream_t>(stream));
}

int lk_comm_check(lk_comm* c) {
  if (!c) return LK_ERR_INVALID;
  DevGuard guard(c->device);
  int flag = 0;
  LK_CUDA(cudaDeviceSynchronize());  // non-blocking streams too (a plain cudaMemcpy would not wait for them)
  LK_CUDA(cudaMemcpy(&flag, c->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag != 0) {
    cudaMemset(c->err_flag, 0, sizeof(int));
    set_error("candidate exchange timed out waiting for a peer (code %d); results are invalid", flag);
    return LK_ERR_CUDA;
  }
  return LK_OK;
}

int lk_comm_destroy(lk_comm* c) {
  if (!c) return LK_OK;
  DevGuard guard(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r] && c->mapped[r]) cudaIpcCloseMemHandle(c->mapped[r]);
  if (c->base) cudaFree(c->base);
  if (c->err_flag) cudaFree(c->err_flag);
  delete c;
  return LK_OK;
}

}  // extern "C"
