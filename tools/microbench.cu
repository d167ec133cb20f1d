// Stand-alone sm_100a microbenchmarks used to size the search kernel's pipeline:
//   1. cp.async.bulk (UBLKCP) global->shared streaming bandwidth vs copy size / ring depth
//   2. tcgen05.mma issue rate (cycles per instruction) for the operand layouts we use
//   3. plain LDG.128 streaming bandwidth (reference point)
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I latent_rag_b200/csrc
//             -I include tools/microbench.cu -o tools/_bin/microbench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lk_ptx.cuh"

using namespace lk;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at line %d: %s\n", cudaGetErrorString(e), __LINE__, #x);  \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

// ---- 1. bulk copy streaming ---------------------------------------------------------
__global__ void __launch_bounds__(64) bulk_stream(const unsigned char* src, size_t bytes_per_cta, int copy_bytes,
                                                  int n_stages, int copies_per_stage, int* err) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
  unsigned char* data = sm + 1024;
  const uint32_t bar0 = ptx::smem_u32(bars);
  const int stage_bytes = copy_bytes * copies_per_stage;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      ptx::mbar_init(bar0 + 8 * s, 1);          // full
      ptx::mbar_init(bar0 + 8 * (32 + s), 1);   // empty
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const unsigned char* p = src + (size_t)blockIdx.x * bytes_per_cta;
  const size_t n_iter = bytes_per_cta / stage_bytes;
  if (threadIdx.x == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (size_t it = 0; it < n_iter; ++it) {
      if (!ptx::mbar_wait(bar0 + 8 * (32 + s), ph ^ 1)) { atomicExch(err, 1); break; }
      ptx::mbar_arrive_expect_tx(bar0 + 8 * s, stage_bytes);
      for (int c = 0; c < copies_per_stage; ++c)
        ptx::bulk_g2s(ptx::smem_u32(data + s * stage_bytes + c * copy_bytes),
                      p + it * stage_bytes + (size_t)c * copy_bytes, copy_bytes, bar0 + 8 * s);
      if (++s == n_stages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int s = 0;
    uint32_t ph = 0;
    for (size_t it = 0; it < n_iter; ++it) {
      if (!ptx::mbar_wait(bar0 + 8 * s, ph)) { atomicExch(err, 2); break; }
      ptx::mbar_arrive(bar0 + 8 * (32 + s));
      if (++s == n_stages) { s = 0; ph ^= 1; }
    }
  }
}

// ---- 3. LDG streaming -----------------------------------------------------------------
__global__ void __launch_bounds__(256) ldg_stream(const uint4* src, size_t n16, unsigned* out) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    uint4 a = __ldg(src + i), b = __ldg(src + i + stride), c = __ldg(src + i + 2 * stride), d = __ldg(src + i + 3 * stride);
    acc += a.x ^ b.y ^ c.z ^ d.w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// ---- 2. MMA rate ------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mma_rate(int n_cols, int layout, int iters, long long* cycles, int* err) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + 64);
  unsigned char* a_sm = sm + 1024;             // 16 KB
  unsigned char* b_sm = a_sm + 16384;          // 32 KB (up to N=256)
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a_sm)[i] = 0;
  const uint32_t bar = ptx::smem_u32(bars);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr), 512);
    ptx::tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::idesc_bf16_f32(128, n_cols);
    const uint32_t lbo = layout == 2 ? 16 : 2048, sbo = layout == 2 ? 1024 : 128;
    const uint32_t kstep = layout == 2 ? 32 : 4096;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + (uint32_t)((it & 1) * 256);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        ptx::umma_bf16(d, ptx::smem_desc(ptx::smem_u32(a_sm) + k * kstep, lbo, sbo, layout),
                       ptx::smem_desc(ptx::smem_u32(b_sm) + k * kstep, lbo, sbo, layout), idesc, k != 0);
    }
    ptx::umma_commit(bar);
    if (!ptx::mbar_wait(bar, 0)) atomicExch(err, 3);
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}


// ---- 2b. MMA rate on a CTA pair (cta_group::2, M = 256): N = 64 / 128 / 256, A operand in shared or tensor memory.
//          The leader issues `iters` groups of 4 MMAs back to back on static (zero) operands; one multicast commit.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) mma_rate_pair(int n_cols, int a_in_tmem, int iters,
                                                                                  long long* cycles, int* err) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + 64);
  unsigned char* a_sm = sm + 1024;             // 16 KB: this CTA's 128 rows of A
  unsigned char* b_sm = a_sm + 16384;          // 16 KB: this CTA's half of B (up to 128 of N = 256 rows)
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a_sm)[i] = 0;
  const uint32_t bar = ptx::smem_u32(bars);
  const uint32_t rank = ptx::cluster_ctarank();
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    ptx::tmem_alloc2(ptx::smem_u32(tmem_ptr), 512);
    ptx::tmem_relinquish2();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = ptx::idesc_bf16_f32(256, n_cols);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + (uint32_t)((it & 1) * 128);  // accumulators in columns 0..255, A (if in TMEM) in 320..
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t db = ptx::smem_desc(ptx::smem_u32(b_sm) + k * 32, 16, 1024, 2);
        if (a_in_tmem) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(tmem + 320u + 8u * k),
                       "l"(db), "r"(idesc), "r"((uint32_t)(k != 0))
                       : "memory");
        } else {
          ptx::umma_bf16_2cta(d, ptx::smem_desc(ptx::smem_u32(a_sm) + k * 32, 16, 1024, 2), db, idesc, k != 0);
        }
      }
    }
    ptx::umma_commit_2cta(bar, 3);
    if (!ptx::mbar_wait(bar, 0)) atomicExch(err, 4);
    if (blockIdx.x == 0) cycles[0] = clock64() - t0;
  } else if (threadIdx.x == 0) {
    if (!ptx::mbar_wait(bar, 0)) atomicExch(err, 5);
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc2(tmem, 512);
  }
}

// ---- 4. TMEM read rate: tcgen05.ld.32x32b.x32 from n_warps warps, `depth` loads in flight per warp
__global__ void __launch_bounds__(512) tmem_read_rate(int n_warps, int depth, int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_ptr), 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < n_warps) {
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t r0[32], r1[32];
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t col = (uint32_t)(((it * 2 + (warp >> 2)) * 32) & 511);
      ptx::tmem_ld32(tmem + lane_base + col, r0);
      if (depth == 2) ptx::tmem_ld32(tmem + lane_base + ((col + 256) & 511), r1);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc = fmaxf(acc, __uint_as_float(r0[j]));
      if (depth == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaxf(acc, __uint_as_float(r1[j]));
      }
    }
    t1 = clock64();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (acc == 12345.f) sink[0] = acc;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  const size_t total = (size_t)8 << 30;  // 8 GiB, far larger than L2
  unsigned char* buf;
  CK(cudaMalloc(&buf, total));
  CK(cudaMemset(buf, 1, total));
  int* err;
  CK(cudaMalloc(&err, 4));
  CK(cudaMemset(err, 0, 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));

  // LDG reference
  {
    unsigned* out;
    CK(cudaMalloc(&out, 4));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      ldg_stream<<<sms * 8, 256>>>(reinterpret_cast<const uint4*>(buf), total / 16, out);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("ldg_stream           : %.1f GB/s\n", total / (ms * 1e-3) / 1e9);
  }

  CK(cudaFuncSetAttribute(bulk_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  const int sizes[] = {2048, 4096, 8192, 16384, 32768};
  for (int cps : {1, 2, 4}) {
    for (int sz : sizes) {
      for (int st : {2, 4, 7, 12}) {
        const size_t stage = (size_t)sz * cps;
        if (stage * st + 1024 > 224 * 1024 || st > 30) continue;
        const size_t per_cta = total / sms / stage * stage;
        for (int rep = 0; rep < 2; ++rep) {
          CK(cudaEventRecord(e0));
          bulk_stream<<<sms, 64, 1024 + stage * st>>>(buf, per_cta, sz, st, cps, err);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        int herr = 0;
        CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
        printf("bulk copy %6d B x%d/stage, %2d stages (%3zu KB in flight): %7.1f GB/s  err=%d\n", sz, cps, st,
               stage * st / 1024, per_cta * sms / (ms * 1e-3) / 1e9, herr);
      }
    }
  }

  // MMA rate
  {
    long long* cyc;
    CK(cudaMalloc(&cyc, 8));
    CK(cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 49152));
    for (int layout : {0, 2}) {
      for (int n : {128, 256}) {
        const int iters = 2000;
        for (int rep = 0; rep < 2; ++rep) {
          mma_rate<<<sms, 128, 1024 + 49152>>>(n, layout, iters, cyc, err);
          CK(cudaDeviceSynchronize());
        }
        long long h = 0;
        int herr = 0;
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
        printf("tcgen05.mma M=128 N=%3d layout=%d: %.1f cycles per MMA (K=16), err=%d\n", n, layout,
               (double)h / (iters * 4), herr);
      }
    }
  }
  {
    long long* cyc;
    int* err2;
    CK(cudaMalloc(&cyc, 8));
    CK(cudaMalloc(&err2, 4));
    CK(cudaMemset(err2, 0, 4));
    CK(cudaFuncSetAttribute(mma_rate_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 32768));
    for (int a_in_tmem : {0, 1})
      for (int n : {64, 128, 256}) {
        const int iters = 2000;
        for (int rep = 0; rep < 2; ++rep) {
          mma_rate_pair<<<sms / 2 * 2, 128, 1024 + 32768>>>(n, a_in_tmem, iters, cyc, err2);
          CK(cudaDeviceSynchronize());
        }
        long long h = 0;
        int herr = 0;
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&herr, err2, 4, cudaMemcpyDeviceToHost));
        printf("tcgen05.mma cta_group::2 M=256 N=%3d A in %s: %.1f cycles per MMA (K=16), err=%d\n", n,
               a_in_tmem ? "TMEM" : "smem", (double)h / (iters * 4), herr);
      }
  }
  {
    long long* cyc;
    float* sink;
    CK(cudaMalloc(&cyc, 8));
    CK(cudaMalloc(&sink, 4));
    for (int depth : {1, 2})
      for (int nw : {1, 4, 8, 12, 16}) {
        const int iters = 4000;
        for (int rep = 0; rep < 2; ++rep) {
          tmem_read_rate<<<sms, 512>>>(nw, depth, iters, cyc, sink);
          CK(cudaDeviceSynchronize());
        }
        long long h = 0;
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double loads = (double)iters * depth;  // per warp; each moves 32 lanes x 32 cols x 4 B = 4 KB
        printf("tcgen05.ld.32x32b.x32 %2d warps, %d in flight: %.1f cycles per load per warp, %.1f B/cycle/SM\n", nw,
               depth, (double)h / loads, 4096.0 * loads * nw / (double)h);
      }
  }
  return 0;
}
