// Writers of the split-bf16 operand planes the tcgen05 linear layers read (lk_gemm_umma.cu,
// lk_ae_umma.cu): an fp32 matrix [rows, k] is stored per 128-row tile as [plane][K block] slabs of
// 16 KB in the SWIZZLE_128B layout of lk_common.cuh, plane 0 = bf16(x), plane 1 = bf16(x - plane 0).
// Kernels that PRODUCE an activation (LayerNorm, attention, a linear layer's epilogue) write the
// planes themselves, 8 consecutive columns (one 16-byte chunk per plane) at a time.
#pragma once

#include "lk_common.cuh"
#include "lk_ptx.cuh"

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

// v[0..16) = columns col0 .. col0+15 (col0 % 16 == 0): the two chunks sit next to each other in the
// swizzled row (positions p and p ^ 1), in an order that depends on the row's parity -- one 256-bit
// store per plane.
__device__ __forceinline__ void store_planes16(unsigned char* planes, int nkb, int64_t row, int col0,
                                               const float (&v)[16], int n_planes) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float x0 = v[2 * e], x1 = v[2 * e + 1];
    hi[e] = pack_bf16x2(x0, x1);
    lo[e] = pack_bf16x2(x0 - __bfloat162float(__float2bfloat16_rn(x0)), x1 - __bfloat162float(__float2bfloat16_rn(x1)));
  }
  const int rin = (int)(row % kBlockRows);
  const bool swap = rin & 1;  // chunk c (even) lands at position c ^ (rin & 7): the upper half of the pair for odd rows
  unsigned char* tile = planes + (row / kBlockRows) * ((int64_t)kOperandPlanes * nkb * kSlabBytes);
  const int pos = (((col0 >> 3) & 7) ^ (rin & 7)) & ~1;
  const int64_t off = (int64_t)(col0 >> 6) * kSlabBytes + rin * kRowBytes + (pos << 4);
  const uint32_t h[8] = {swap ? hi[4] : hi[0], swap ? hi[5] : hi[1], swap ? hi[6] : hi[2], swap ? hi[7] : hi[3],
                         swap ? hi[0] : hi[4], swap ? hi[1] : hi[5], swap ? hi[2] : hi[6], swap ? hi[3] : hi[7]};
  ptx::stg256(tile + off, h);
  if (n_planes == 2) {
    const uint32_t l[8] = {swap ? lo[4] : lo[0], swap ? lo[5] : lo[1], swap ? lo[6] : lo[2], swap ? lo[7] : lo[3],
                           swap ? lo[0] : lo[4], swap ? lo[1] : lo[5], swap ? lo[2] : lo[6], swap ? lo[3] : lo[7]};
    ptx::stg256(tile + (int64_t)nkb * kSlabBytes + off, l);
  }
}

}  // namespace lk
