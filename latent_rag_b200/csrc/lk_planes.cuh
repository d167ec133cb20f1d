// Writers of the split-bf16 operand planes the tcgen05 linear layers read (lk_gemm_umma.cu,
// lk_ae_umma.cu): an fp32 matrix [rows, k] is stored per 128-row tile as [plane][K block] slabs of
// 16 KB in the SWIZZLE_128B layout of lk_common.cuh, plane 0 = bf16(x), plane 1 = bf16(x - plane 0).
// Kernels that PRODUCE an activation (LayerNorm, attention, a linear layer's epilogue) write the
// planes themselves, 8 consecutive columns (one 16-byte chunk per plane) at a time.
#pragma once

#include "lk_common.cuh"

namespace lk {

constexpr int kOperandPlanes = 2;

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// v[0..8) = columns col0 .. col0+7 (col0 % 8 == 0) of row `row` of a [*, 64 * nkb] matrix
__device__ __forceinline__ void store_planes8(unsigned char* planes, int nkb, int64_t row, int col0,
                                              const float (&v)[8], int n_planes) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float x0 = v[2 * e], x1 = v[2 * e + 1];
    hi[e] = pack_bf16x2(x0, x1);
    lo[e] = pack_bf16x2(x0 - __bfloat162float(__float2bfloat16_rn(x0)), x1 - __bfloat162float(__float2bfloat16_rn(x1)));
  }
  unsigned char* tile = planes + (row / kBlockRows) * ((int64_t)kOperandPlanes * nkb * kSlabBytes);
  const int64_t off = (int64_t)(col0 >> 6) * kSlabBytes + slab_chunk_offset((int)(row % kBlockRows), (col0 >> 3) & 7);
  *reinterpret_cast<uint4*>(tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (n_planes == 2)
    *reinterpret_cast<uint4*>(tile + (int64_t)nkb * kSlabBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace lk
