// Fused FFN + residual + LayerNorm on a CTA PAIR (tcgen05 cta_group::2) — same contract as ffn.cu:
//     OUT = LN( X + relu(X W1^T + b1) W2^T + b2 )        X: [rows][256] fp16,  W1: [F][256],  W2: [256][F]
//
// Why a pair.  With one CTA per 128-row tile every MMA reads its whole B operand from that SM's shared memory and every
// weight byte is also written there by TMA: 2560 shared-memory wavefronts (128 B) per 128-wide hidden chunk against
// 2048 tensor-pipe cycles — the single-CTA kernel (ffn.cu) is shared-memory-bandwidth bound (measured: tensor pipe 64 %
// busy in its main loop, multicast of the weight loads does not help because the reads, not the L2 traffic, bind).
// Two CTAs of a cluster take adjacent row tiles (an M = 256 MMA, 128 rows per SM) and each holds HALF of every weight
// tile: B is read half from each SM, the weight stream per SM halves (1536 wavefronts per chunk).
//
// Per CTA:  X tile 64 KB (A operand of GEMM1 and residual; reused as the output staging tile), 10 x 16 KB weight ring,
// TMEM: Y [0,256) fp32 accumulator of GEMM2, H0/H1 [256,512) GEMM1 accumulators; the fp16 hidden chunk P(j) is written
// back over H(j) (packed two per column) and is the A operand of GEMM2 straight from TMEM.
//   W1 chunk j (128 hidden x 256 k): CTA r holds hidden rows [64 r, 64 r + 64)  -> 2 slots of two [64][64] k-sub-tiles
//   W2 chunk j (256 out x 128 hidden): CTA r holds out rows [128 r, +128)       -> 2 slots of one [128][64] k-sub-tile
// Roles (320 threads): warp 0 TMA producer (both CTAs; weight loads signal the LEADER's barrier), warp 1 tcgen05 issuer
// (leader CTA only; in the peer it relays "my X tile has landed"), warps 2-9 epilogue, two threads per row.
//
// Reference: torch.nn.TransformerEncoderLayer._ff_block + norm2 (FS:model:147) and
// TransformerEncoderFusionLayer._ff_block + norm22 (FS-EEND/nnet/modules/merge_tfm_encoder.py:373,397-399).
#include "once.h"
#include "ffn.cuh"
#include "ffn_tile.cuh"
#include "pair.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

using namespace ffn_detail;
using namespace pair;

constexpr int kRows = 128;
constexpr int kChunk = 128;                       // hidden units per chunk (across the pair)
constexpr int kXBytes = 4 * kSlotBytes;           // 64 KB
constexpr int kSlots = 10;
constexpr int kOffX = 0;
constexpr int kOffW = kXBytes;
constexpr int kOffLN = kOffW;                     // after the last MMA: b2 | gamma | beta (3 KB) + stats exchange (2 KB)
constexpr int kOffXchg = kOffW + 4096;
constexpr int kSmemBytes = kXBytes + kSlots * kSlotBytes + 1024;
constexpr uint32_t kTmemCols = 512;               // Y [0,256), H0 [256,384), H1 [384,512)

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
ffn_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO, const FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // identical offsets in both CTAs of the pair (same kernel image): remote barriers are addressed with mapa
  __shared__ __align__(8) uint64_t x_full, x_peer, w_full[kSlots], w_empty[kSlots], h_full[2], p_full[2], y_full;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float b1_smem[2][kChunk];

  // 1024-byte alignment as an OFFSET from the shared array: pointer arithmetic through uintptr_t would hide the shared
  // address space from the compiler and turn every access below into a generic load / store
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);

  const int m_tile = blockIdx.x;                // tiles (2c, 2c+1) form the pair's 256-row MMA tile
  const int seq = m_tile / p.tiles_per_seq;     // an odd tile count leaves one CTA past the end: loads zero-fill, stores clip
  const int t0 = (m_tile % p.tiles_per_seq) * kRows;
  const int n_chunks = p.F / kChunk;

  if (tid == 0) {
    mbar_init(&x_full, 1);
    mbar_init(&x_peer, 1);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&w_full[s], 1);      // the leader's arrive.expect_tx; bytes from both CTAs' loads
      mbar_init(&w_empty[s], 1);     // one multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&h_full[i], 1);
      mbar_init(&p_full[i], 512);    // every epilogue thread of both CTAs (leader's copy is the one waited on)
    }
    mbar_init(&y_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1) tmem_alloc_pair(&tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // peer barriers are initialised and both TMEM allocations done before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tmem_Y = tmem_base;

  // Slot sequence shared by the producers and the issuer:  W1(0) | W1(1) W2(0) | W1(2) W2(1) | ... | W2(n-1),
  // two slots per W1(j) (k-sub-tile pairs) and two per W2(j) (k-sub-tiles).
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected lane)
    const uint32_t wfull_leader = mapa_rank(smem_u32(&w_full[0]), 0);
    if (elect_one()) {
      mbar_arrive_expect_tx(&x_full, kXBytes);
      for (int ks = 0; ks < 4; ++ks) tma_load_3d(smem + kOffX + ks * kSlotBytes, &tmX, &x_full, ks * 64, t0, seq);
    }
    __syncwarp();
    uint32_t use = 0;
    auto load_w1 = [&](int j) {
      for (int q = 0; q < 2; ++q, ++use) {
        const int s = use % kSlots;
        mbar_wait(&w_empty[s], ((use / kSlots) & 1) ^ 1, 131);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&w_full[s], 2 * kSlotBytes);
          uint8_t* dst = smem + kOffW + s * kSlotBytes;
          const int row = j * kChunk + static_cast<int>(rank) * 64;
          tma_load_2d_pair(dst, &tmW1, wfull_leader + s * 8, (2 * q) * 64, row);
          tma_load_2d_pair(dst + kSlotBytes / 2, &tmW1, wfull_leader + s * 8, (2 * q + 1) * 64, row);
        }
        __syncwarp();
      }
    };
    auto load_w2 = [&](int j) {
      for (int ks2 = 0; ks2 < 2; ++ks2, ++use) {
        const int s = use % kSlots;
        mbar_wait(&w_empty[s], ((use / kSlots) & 1) ^ 1, 132);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&w_full[s], 2 * kSlotBytes);
          tma_load_2d_pair(smem + kOffW + s * kSlotBytes, &tmW2, wfull_leader + s * 8, j * kChunk + ks2 * 64,
                           static_cast<int>(rank) * 128);
        }
        __syncwarp();
      }
    };
    load_w1(0);
    for (int j = 0; j < n_chunks; ++j) {
      if (j + 1 < n_chunks) load_w1(j + 1);
      load_w2(j);
    }
  } else if (warp == 1) {
    if (!leader) {
      // the leader's MMAs read this CTA's X tile: tell it when the tile has landed
      mbar_wait(&x_full, 0, 133);
      if (elect_one()) mbar_arrive_cluster(mapa_rank(smem_u32(&x_peer), 0));
      __syncwarp();
    } else {
      // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane)
      constexpr uint32_t idesc_g1 = make_idesc_f16(256, 128, false);
      constexpr uint32_t idesc_g2 = make_idesc_f16(256, 256, false);
      uint32_t use = 0;
      auto gemm1 = [&](int j) {
        const int hb = j & 1;   // H(j) aliases P(j-2), whose GEMM2 precedes this GEMM1 in the in-order tensor pipe
        const uint32_t tmem_H = tmem_base + 256 + hb * 128;
        for (int q = 0; q < 2; ++q, ++use) {
          const int s = use % kSlots;
          mbar_wait(&w_full[s], (use / kSlots) & 1, 134);
          tc_fence_after();
          const uint32_t xa = smem_u32(smem + kOffX + 2 * q * kSlotBytes);
          const uint32_t wb = smem_u32(smem + kOffW + s * kSlotBytes);
          if (elect_one()) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint64_t adesc = smem_desc_sw128(xa + i * kSlotBytes);
              const uint64_t bdesc = smem_desc_sw128(wb + i * (kSlotBytes / 2));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma2_f16(tmem_H, adesc + 2 * kk, bdesc + 2 * kk, idesc_g1, (q > 0 || i > 0 || kk > 0) ? 1u : 0u);
            }
            umma2_commit(&w_empty[s]);
            if (q == 1) umma2_commit(&h_full[hb]);
          }
          __syncwarp();
        }
      };
      auto gemm2 = [&](int j) {
        const int pb = j & 1;
        mbar_wait_cluster(&p_full[pb], (j >> 1) & 1, 135);   // P(j) written by the epilogue warps of both CTAs
        tc_fence_after();
        for (int ks2 = 0; ks2 < 2; ++ks2, ++use) {
          const int s = use % kSlots;
          mbar_wait(&w_full[s], (use / kSlots) & 1, 136);
          tc_fence_after();
          const uint64_t bdesc = smem_desc_sw128(smem_u32(smem + kOffW + s * kSlotBytes));
          const uint32_t tmem_P = tmem_base + 256 + pb * 128 + ks2 * 32;   // 64 hidden = 32 packed columns
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma2_f16_ts(tmem_Y, tmem_P + 8 * kk, bdesc + 2 * kk, idesc_g2, (j > 0 || ks2 > 0 || kk > 0) ? 1u : 0u);
            umma2_commit(&w_empty[s]);
          }
          __syncwarp();
        }
      };
      mbar_wait(&x_full, 0, 137);
      mbar_wait_cluster(&x_peer, 0, 138);
      tc_fence_after();
      gemm1(0);
      for (int j = 0; j < n_chunks; ++j) {
        if (j + 1 < n_chunks) gemm1(j + 1);
        gemm2(j);
      }
      if (elect_one()) umma2_commit(&y_full);
      __syncwarp();
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps 2..9, two threads per row
    const int et = tid - 64;                      // 0..255
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int ch = (warp - 2) >> 2;               // column half owned by this thread
    const int r = quarter * 32 + lane;            // tile row
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t pfull_leader = mapa_rank(smem_u32(&p_full[0]), 0);
    // LayerNorm / bias vectors of the final epilogue: one element of each per thread, parked in registers until the
    // weight ring is free to hold them
    const float ln_b2 = __ldg(p.b2 + et), ln_g = __ldg(p.ln_g + et), ln_b = __ldg(p.ln_b + et);
    float b1_next = (et < kChunk) ? __ldg(p.b1 + et) : 0.f;
    for (int j = 0; j < n_chunks; ++j) {
      const int hb = j & 1;
      if (et < kChunk) {   // stage this chunk's 128 bias values (double-buffered), prefetch the next chunk's
        b1_smem[hb][et] = b1_next;
        if (j + 1 < n_chunks) b1_next = __ldg(p.b1 + (j + 1) * kChunk + et);
      }
      named_bar_sync(2, 256);
      mbar_wait(&h_full[hb], (j >> 1) & 1, 140);
      tc_fence_after();
      const uint32_t tH = tmem_base + 256 + hb * 128 + lane_base;
      uint32_t r0[32], r1[32];
      tmem_ld32(tH + ch * 64, r0);
      tmem_ld32(tH + ch * 64 + 32, r1);
      tmem_ld_wait();
      // P(j) (fp16 pairs) is written back over H(j): packed columns [32 ch, +32) overlay fp32 columns [32 ch, +32),
      // which for ch = 1 belong to the partner thread (same row, ch = 0) -> it must have finished reading them.
      if (ch == 0) named_bar_arrive(4 + quarter, 64);
      const float4* bs = reinterpret_cast<const float4*>(&b1_smem[hb][ch * 64]);
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = bs[i];     // same address in every lane: shared-memory broadcast
        pk[2 * i] = pack_half2(fmaxf(__uint_as_float(r0[4 * i + 0]) + t.x, 0.f),
                               fmaxf(__uint_as_float(r0[4 * i + 1]) + t.y, 0.f));
        pk[2 * i + 1] = pack_half2(fmaxf(__uint_as_float(r0[4 * i + 2]) + t.z, 0.f),
                                   fmaxf(__uint_as_float(r0[4 * i + 3]) + t.w, 0.f));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = bs[8 + i];
        pk[16 + 2 * i] = pack_half2(fmaxf(__uint_as_float(r1[4 * i + 0]) + t.x, 0.f),
                                    fmaxf(__uint_as_float(r1[4 * i + 1]) + t.y, 0.f));
        pk[16 + 2 * i + 1] = pack_half2(fmaxf(__uint_as_float(r1[4 * i + 2]) + t.z, 0.f),
                                        fmaxf(__uint_as_float(r1[4 * i + 3]) + t.w, 0.f));
      }
      if (ch == 1) named_bar_sync(4 + quarter, 64);
      tmem_st32(tH + ch * 32, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_cluster(pfull_leader + hb * 8);
    }

    // ---- final: Y + b2 + X -> LayerNorm -> fp16, written back over the X tile -> TMA store.  One pass: each thread
    // keeps its 128 columns in registers.
    mbar_wait(&y_full, 0, 142);   // every MMA of the pair has completed: X and the weight ring are no longer read
    mbar_wait(&x_full, 0, 143);   // (long since landed; makes the TMA-written X tile visible to this thread)
    tc_fence_after();
    float* ln_s = reinterpret_cast<float*>(smem + kOffLN);
    float2* xchg = reinterpret_cast<float2*>(smem + kOffXchg);
    ln_s[et] = ln_b2;
    ln_s[256 + et] = ln_g;
    ln_s[512 + et] = ln_b;
    named_bar_sync(1, 256);
    uint8_t* xtile = smem + kOffX;
    const uint32_t tY = tmem_Y + lane_base + ch * 128;
    float z[128];
    {
      uint32_t v0[32], v1[32], v2[32], v3[32];
      tmem_ld32(tY, v0);
      tmem_ld32(tY + 32, v1);
      tmem_ld32(tY + 64, v2);
      tmem_ld32(tY + 96, v3);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        z[i] = __uint_as_float(v0[i]);
        z[32 + i] = __uint_as_float(v1[i]);
        z[64 + i] = __uint_as_float(v2[i]);
        z[96 + i] = __uint_as_float(v3[i]);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float aux[32];
      tile_read32(xtile, r, ch * 4 + c, aux);
      const float4* b2s = reinterpret_cast<const float4*>(ln_s + ch * 128 + c * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = b2s[i];
        z[c * 32 + 4 * i + 0] += aux[4 * i + 0] + t.x;
        z[c * 32 + 4 * i + 1] += aux[4 * i + 1] + t.y;
        z[c * 32 + 4 * i + 2] += aux[4 * i + 2] + t.z;
        z[c * 32 + 4 * i + 3] += aux[4 * i + 3] + t.w;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) s += z[c * 32 + i];
    }
    const float mh = s * (1.f / 128.f);
    float m2h = 0.f;
#pragma unroll
    for (int i = 0; i < 128; ++i) {
      const float d = z[i] - mh;
      m2h = fmaf(d, d, m2h);
    }
    xchg[ch * 128 + r] = make_float2(mh, m2h);
    named_bar_sync(1, 256);
    const float2 o = xchg[(1 - ch) * 128 + r];
    const float mean = 0.5f * (mh + o.x);
    const float dm = mh - o.x;
    const float m2 = m2h + o.y + dm * dm * 64.f;      // Chan merge of two 128-element halves
    const float rstd = rsqrtf(m2 * (1.f / 256.f) + p.ln_eps);
    const bool zero_row = p.seq_len != nullptr && seq < p.n_seq && (t0 + r) >= p.seq_len[seq];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float outv[32];
      const float4* gs = reinterpret_cast<const float4*>(ln_s + 256 + ch * 128 + c * 32);
      const float4* bs = reinterpret_cast<const float4*>(ln_s + 512 + ch * 128 + c * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 g = gs[i], b = bs[i];
        outv[4 * i + 0] = zero_row ? 0.f : fmaf((z[c * 32 + 4 * i + 0] - mean) * rstd, g.x, b.x);
        outv[4 * i + 1] = zero_row ? 0.f : fmaf((z[c * 32 + 4 * i + 1] - mean) * rstd, g.y, b.y);
        outv[4 * i + 2] = zero_row ? 0.f : fmaf((z[c * 32 + 4 * i + 2] - mean) * rstd, g.z, b.z);
        outv[4 * i + 3] = zero_row ? 0.f : fmaf((z[c * 32 + 4 * i + 3] - mean) * rstd, g.w, b.w);
      }
      tile_write32(xtile, r, ch * 4 + c, outv);   // in place: this thread read exactly these elements above
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (et == 0) {
      for (int sub = 0; sub < 4; ++sub) tma_store_3d(&tmO, xtile + sub * kSlotBytes, sub * 64, t0, seq);
      tma_store_commit();
      tma_store_wait_read0();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA exits (or frees TMEM) while its peer may still signal it or the pair's MMAs run
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

}  // namespace

void launch_ffn_pair(const CUtensorMap& tmX, const CUtensorMap& tmW1_box64, const CUtensorMap& tmW2,
                     const CUtensorMap& tmO, const FfnParams& p, cudaStream_t stream) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(ffn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  }
  const int tiles = p.n_seq * p.tiles_per_seq;
  ffn_pair_kernel<<<(tiles + 1) / 2 * 2, 320, kSmemBytes, stream>>>(tmX, tmW1_box64, tmW2, tmO, p);
}

}  // namespace fseend
