// Device helpers shared by the fused FFN kernels (ffn.cu, ffn_pair.cu): cluster primitives, TMEM store, and row access
// to fp16 tiles stored as 64-column sub-tiles of 16 KB ([128 rows][64 cols], 128-byte swizzle).
#pragma once
#include "ptx.cuh"

namespace fseend {
namespace ffn_detail {

constexpr int kSlotBytes = 128 * 64 * 2;    // 16 KB

__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]: A is M x K fp16 packed two per 32-bit TMEM column (lane = row).
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_sync(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void load_vec32(const float* __restrict__ p, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
    v[4 * i + 0] = t.x;
    v[4 * i + 1] = t.y;
    v[4 * i + 2] = t.z;
    v[4 * i + 3] = t.w;
  }
}
// 32 fp16 columns [c*32, c*32+32) of row r in a tile made of 64-column sub-tiles of kSlotBytes each
__device__ __forceinline__ void tile_read32(const uint8_t* tile, int r, int c, float (&v)[32]) {
  const uint8_t* sub = tile + (c >> 1) * kSlotBytes;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u = *reinterpret_cast<const uint4*>(sub + sw128_offset(r, (c & 1) * 4 + q));
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      v[q * 8 + 2 * j] = f.x;
      v[q * 8 + 2 * j + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void tile_write32(uint8_t* tile, int r, int c, const float (&v)[32]) {
  uint8_t* sub = tile + (c >> 1) * kSlotBytes;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_half2(v[q * 8 + 0], v[q * 8 + 1]);
    u.y = pack_half2(v[q * 8 + 2], v[q * 8 + 3]);
    u.z = pack_half2(v[q * 8 + 4], v[q * 8 + 5]);
    u.w = pack_half2(v[q * 8 + 6], v[q * 8 + 7]);
    *reinterpret_cast<uint4*>(sub + sw128_offset(r, (c & 1) * 4 + q)) = u;
  }
}

}  // namespace ffn_detail
}  // namespace fseend
