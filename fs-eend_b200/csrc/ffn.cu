// Fused position-wise feed-forward + residual + LayerNorm (post-norm transformer FFN block):
//     OUT = LN( X + relu(X W1^T + b1) W2^T + b2 )        X: [rows][256] fp16,  W1: [F][256],  W2: [256][F]
// The F-wide intermediate never leaves the SM: per 128-row tile the hidden dimension is processed in
// chunks of 128; chunk j's GEMM1 accumulator (TMEM) is bias+ReLU'd into an fp16 smem tile that is the
// A operand of chunk j's GEMM2, which accumulates the 128x256 output in TMEM across all chunks.
//
// Reference: torch.nn.TransformerEncoderLayer._ff_block + norm2 (FS:model:147) and
// TransformerEncoderFusionLayer._ff_block + norm22 (FS-EEND/nnet/modules/merge_tfm_encoder.py:373,397-399).
//
// Roles (192 threads):  warp 0 lane 0 = TMA producer, warp 1 lane 0 = tcgen05 issuer, warps 2-5 = epilogue.
// Weights stream through a ring of 16 KB slots ([128 rows][64 k] fp16, 128B swizzle).  With kCluster = 2 the
// two CTAs of a cluster work on adjacent row tiles and share every weight slot: each CTA fetches half of the
// slots and TMA-multicasts them into both CTAs' rings, halving L2->SM weight traffic (the binding limit of this
// block: 2 MB of weights per 268 MFLOP tile).
#include "once.h"
#include "ffn.cuh"
#include "ffn_tile.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

using namespace ffn_detail;

constexpr int kRows = 128;
constexpr int kChunk = 128;                 // hidden units per chunk
using ffn_detail::kSlotBytes;              // 16 KB
constexpr int kXBytes = 4 * kSlotBytes;     // 64 KB: X tile as 4 k-sub-tiles
constexpr int kPBytes = 2 * kSlotBytes;     // 32 KB per P buffer (2 k-sub-tiles)
constexpr int kOffX = 0;
// SS variant: P (fp16 hidden chunk) double-buffered in smem, 6-slot weight ring.
// TS variant: P lives in TMEM (aliasing the GEMM1 accumulator it was computed from) and is the A operand of
//             GEMM2 straight from TMEM; the 64 KB this frees go to the weight ring (10 slots).
template <bool kTS> struct Lay {
  static constexpr int kSlots = kTS ? 10 : 6;
  static constexpr int kOffP = kOffX + kXBytes;                         // SS only: 2 buffers
  static constexpr int kOffW = kTS ? kOffX + kXBytes : kOffP + 2 * kPBytes;
  static constexpr int kOffStage = kTS ? kOffW : kOffP;                 // 64 KB output staging after the last MMA
};
constexpr int kMaxSlots = 10;
constexpr int kSmemBytes = kXBytes + 2 * kPBytes + 6 * kSlotBytes + 1024;   // = 64 + 160 KB + slack, both variants
constexpr uint32_t kTmemCols = 512;                  // Y [0,256), H0 [256,384), H1 [384,512)

template <int kCluster, bool kTS>
__global__ void __launch_bounds__(192, 1)
ffn_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
           const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO, const FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int kSlots = Lay<kTS>::kSlots;
  constexpr int kOffP = Lay<kTS>::kOffP;
  constexpr int kOffW = Lay<kTS>::kOffW;
  __shared__ __align__(8) uint64_t x_full, w_full[kMaxSlots], w_empty[kMaxSlots], h_full[2], h_empty[2], p_full[2],
      p_empty[2], y_full;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float b1_smem[2][kChunk];

  // 1024-byte alignment as an OFFSET from the shared array: pointer arithmetic through uintptr_t would hide the shared
  // address space from the compiler and turn every access below into a generic load / store
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = (kCluster > 1) ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (kCluster > 1) ? 0x3 : 0x1;

  const int m_tile = blockIdx.x;
  const int seq = m_tile / p.tiles_per_seq;
  const int t0 = (m_tile % p.tiles_per_seq) * kRows;
  const int n_chunks = p.F / kChunk;

  if (tid == 0) {
    mbar_init(&x_full, 1);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], kCluster);   // one tcgen05.commit arrival per CTA of the cluster
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&h_full[i], 1);
      mbar_init(&h_empty[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(&y_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();   // peers' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tmem_Y = tmem_base;

  // Slot sequence shared by producer and MMA issuer:  W1(0) | W1(1) W2(0) | W1(2) W2(1) | ... | W2(n-1)
  // every "group" is 4 slots; group g: g == 0 -> W1(0); g odd -> W1((g+1)/2) if it exists; g even -> W2(g/2 - 1)
  // Enumerated explicitly below to keep both roles in lock-step.

  // Role warps run their loops with all 32 lanes converged and elect one lane only around the TMA / tcgen05
  // instructions: descriptors, addresses and loop state then live in uniform registers.  (Issuing from a single
  // diverged lane made ptxas wrap every UTCHMMA in an ELECT loop with R2UR moves — ~13 dependent instructions per MMA,
  // which left the tensor pipe ~45 % busy.)
  if (warp == 0) {
    {
      // ------------------------------------------------------------------ TMA producer
      if (elect_one()) {
        mbar_arrive_expect_tx(&x_full, kXBytes);
        for (int ks = 0; ks < 4; ++ks) tma_load_3d(smem + kOffX + ks * kSlotBytes, &tmX, &x_full, ks * 64, t0, seq);
      }
      __syncwarp();
      uint32_t use = 0;   // global slot-use counter
      auto load_slot = [&](const CUtensorMap* tm, int c0, int c1) {
        const int s = use % kSlots;
        const uint32_t ph = (use / kSlots) & 1;
        mbar_wait(&w_empty[s], ph ^ 1, 31);
        if (elect_one()) {
          mbar_arrive_expect_tx(&w_full[s], kSlotBytes);
          if (kCluster == 1) {
            tma_load_2d(smem + kOffW + s * kSlotBytes, tm, &w_full[s], c0, c1);
          } else if ((use & 1u) == rank) {
            tma_load_2d_mc(smem + kOffW + s * kSlotBytes, tm, &w_full[s], c0, c1, kMask);
          }
        }
        __syncwarp();
        ++use;
      };
      auto load_w1 = [&](int j) {   // W1 rows [j*128, +128), k-sub-tile ks -> box (64 k, 128 rows)
        for (int ks = 0; ks < 4; ++ks) load_slot(&tmW1, ks * 64, j * kChunk);
      };
      auto load_w2 = [&](int j) {   // W2[:, j*128 .. +128): k-sub-tile ks2, N halves 0/1 -> box (64 k, 128 rows)
        for (int ks2 = 0; ks2 < 2; ++ks2)
          for (int nh = 0; nh < 2; ++nh) load_slot(&tmW2, j * kChunk + ks2 * 64, nh * 128);
      };
      load_w1(0);
      for (int j = 0; j < n_chunks; ++j) {
        if (j + 1 < n_chunks) load_w1(j + 1);
        load_w2(j);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      // ------------------------------------------------------------------ MMA issuer (warp-converged, see above)
      constexpr uint32_t idesc_g1 = make_idesc_f16(128, 128, false);
      constexpr uint32_t idesc_g2 = make_idesc_f16(128, 256, false);
      uint32_t use = 0;
      auto release_slot = [&](int s) {   // called by the elected lane
        if (kCluster == 1) umma_commit(&w_empty[s]);
        else umma_commit_mc(&w_empty[s], kMask);
      };
      auto gemm1 = [&](int j) {
        const int hb = j & 1;
        if (!kTS) {
          mbar_wait(&h_empty[hb], ((j >> 1) & 1) ^ 1, 32);   // epilogue has drained H(j-2)
          tc_fence_after();
        }   // TS: H(j) aliases P(j-2), whose GEMM2 precedes this GEMM1 in the in-order tensor pipe
        const uint32_t tmem_H = tmem_base + 256 + hb * 128;
        for (int ks = 0; ks < 4; ++ks) {
          const int s = use % kSlots;
          mbar_wait(&w_full[s], (use / kSlots) & 1, 33);
          tc_fence_after();
          const uint64_t adesc = smem_desc_sw128(smem_u32(smem + kOffX + ks * kSlotBytes));
          const uint64_t bdesc = smem_desc_sw128(smem_u32(smem + kOffW + s * kSlotBytes));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(tmem_H, adesc + 2 * kk, bdesc + 2 * kk, idesc_g1, (ks > 0 || kk > 0) ? 1u : 0u);
            release_slot(s);
            if (ks == 3) umma_commit(&h_full[hb]);
          }
          __syncwarp();
          ++use;
        }
      };
      auto gemm2 = [&](int j) {
        const int pb = j & 1;
        mbar_wait(&p_full[pb], (j >> 1) & 1, 34);            // P(j) written by the epilogue warps
        tc_fence_after();
        for (int ks2 = 0; ks2 < 2; ++ks2) {
          const int s0 = use % kSlots;                        // slots s0 (N rows 0-127) and s0+1 (128-255) are adjacent
          mbar_wait(&w_full[s0], (use / kSlots) & 1, 35);
          mbar_wait(&w_full[s0 + 1], ((use + 1) / kSlots) & 1, 36);
          tc_fence_after();
          const uint64_t bdesc = smem_desc_sw128(smem_u32(smem + kOffW + s0 * kSlotBytes));
          if (elect_one()) {
            if (kTS) {
              const uint32_t tmem_P = tmem_base + 256 + pb * 128 + ks2 * 32;   // 64 hidden = 32 packed columns
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_ts(tmem_Y, tmem_P + 8 * kk, bdesc + 2 * kk, idesc_g2, (j > 0 || ks2 > 0 || kk > 0) ? 1u : 0u);
            } else {
              const uint64_t adesc = smem_desc_sw128(smem_u32(smem + kOffP + pb * kPBytes + ks2 * kSlotBytes));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16(tmem_Y, adesc + 2 * kk, bdesc + 2 * kk, idesc_g2, (j > 0 || ks2 > 0 || kk > 0) ? 1u : 0u);
            }
            release_slot(s0);
            release_slot(s0 + 1);
            if (!kTS && ks2 == 1) umma_commit(&p_empty[pb]);   // TS: nobody waits on p_empty (the P columns are
                                                                 // protected by the in-order tensor pipe) - an arrive
                                                                 // that is never waited on is a synccheck finding
          }
          __syncwarp();
          use += 2;
        }
      };
      mbar_wait(&x_full, 0, 30);
      tc_fence_after();
      gemm1(0);
      for (int j = 0; j < n_chunks; ++j) {
        if (j + 1 < n_chunks) gemm1(j + 1);
        gemm2(j);
      }
      if (elect_one()) umma_commit(&y_full);
      __syncwarp();
    }
    __syncwarp();
  } else {
    // -------------------------------------------------------------------- epilogue warps (2..5)
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;            // tile row
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    float acc[32], aux[32];
    const int et = tid - 64;                       // 0..127 among the epilogue threads
    float b1_next = __ldg(p.b1 + et);              // bias of chunk 0, one element per thread
    for (int j = 0; j < n_chunks; ++j) {
      const int hb = j & 1;
      // stage this chunk's 128 bias values in smem (double-buffered), prefetch the next chunk's
      b1_smem[hb][et] = b1_next;
      if (j + 1 < n_chunks) b1_next = __ldg(p.b1 + (j + 1) * kChunk + et);
      named_bar_sync(2, 128);
      mbar_wait(&h_full[hb], (j >> 1) & 1, 40);
      tc_fence_after();
      if (!kTS && j >= 2) mbar_wait(&p_empty[hb], ((j >> 1) & 1) ^ 1, 41);   // GEMM2(j-2) has consumed this P buffer
      uint8_t* ptile = smem + kOffP + hb * kPBytes;
      const uint32_t tH = tmem_base + 256 + hb * 128 + lane_base;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        // two 32-column TMEM loads in flight per wait
        uint32_t r0[32], r1[32];
        tmem_ld32(tH + half * 64, r0);
        tmem_ld32(tH + half * 64 + 32, r1);
        tmem_ld_wait();
        const float4* bs = reinterpret_cast<const float4*>(&b1_smem[hb][half * 64]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = bs[i];     // same address in every lane: shared-memory broadcast
          acc[4 * i + 0] = fmaxf(__uint_as_float(r0[4 * i + 0]) + t.x, 0.f);
          acc[4 * i + 1] = fmaxf(__uint_as_float(r0[4 * i + 1]) + t.y, 0.f);
          acc[4 * i + 2] = fmaxf(__uint_as_float(r0[4 * i + 2]) + t.z, 0.f);
          acc[4 * i + 3] = fmaxf(__uint_as_float(r0[4 * i + 3]) + t.w, 0.f);
        }
        if (kTS) {
#pragma unroll
          for (int i = 0; i < 16; ++i) r0[i] = pack_half2(acc[2 * i], acc[2 * i + 1]);
        } else {
          tile_write32(ptile, r, half * 2, acc);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = bs[8 + i];
          acc[4 * i + 0] = fmaxf(__uint_as_float(r1[4 * i + 0]) + t.x, 0.f);
          acc[4 * i + 1] = fmaxf(__uint_as_float(r1[4 * i + 1]) + t.y, 0.f);
          acc[4 * i + 2] = fmaxf(__uint_as_float(r1[4 * i + 2]) + t.z, 0.f);
          acc[4 * i + 3] = fmaxf(__uint_as_float(r1[4 * i + 3]) + t.w, 0.f);
        }
        if (kTS) {
#pragma unroll
          for (int i = 0; i < 16; ++i) r0[16 + i] = pack_half2(acc[2 * i], acc[2 * i + 1]);
          // P(j) overwrites the H(j) columns this thread has already consumed: packed columns [half*32, +32)
          tmem_st32(tH + half * 32, r0);
        } else {
          tile_write32(ptile, r, half * 2 + 1, acc);
        }
      }
      if (kTS) tmem_st_wait();
      if (!kTS) fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[hb]);
      if (!kTS) mbar_arrive(&h_empty[hb]);
    }
    // ---- final: Y + b2 + X -> LayerNorm -> fp16 -> staging (P buffers, 64 KB) -> TMA store
    mbar_wait(&y_full, 0, 42);
    mbar_wait(&x_full, 0, 43);   // residual tile (long since landed; makes the TMA write visible to this thread)
    tc_fence_after();
    const uint8_t* xtile = smem + kOffX;
    uint8_t* staging = smem + Lay<kTS>::kOffStage;
    const uint32_t tY = tmem_Y + lane_base;
    float n = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      tmem_ld32_sync(tY + c * 32, acc);
      load_vec32(p.b2 + c * 32, aux);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] += aux[i];
      tile_read32(xtile, r, c, aux);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        acc[i] += aux[i];
        s += acc[i];
      }
      const float cm = s * (1.f / 32.f);
      float cm2 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float d = acc[i] - cm;
        cm2 = fmaf(d, d, cm2);
      }
      const float nn = n + 32.f;
      const float delta = cm - mean;
      m2 += cm2 + delta * delta * (n * 32.f / nn);
      mean += delta * (32.f / nn);
      n = nn;
    }
    const float rstd = rsqrtf(m2 * (1.f / 256.f) + p.ln_eps);
    const bool zero_row = p.seq_len != nullptr && seq < p.n_seq && (t0 + r) >= p.seq_len[seq];
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      tmem_ld32_sync(tY + c * 32, acc);
      load_vec32(p.b2 + c * 32, aux);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] += aux[i];
      tile_read32(xtile, r, c, aux);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = (acc[i] + aux[i] - mean) * rstd;
      load_vec32(p.ln_g + c * 32, aux);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] *= aux[i];
      load_vec32(p.ln_b + c * 32, aux);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = zero_row ? 0.f : acc[i] + aux[i];
      tile_write32(staging, r, c, acc);
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (tid == 64) {   // first epilogue thread
      for (int sub = 0; sub < 4; ++sub) tma_store_3d(&tmO, staging + sub * kSlotBytes, sub * 64, t0, seq);
      tma_store_commit();
      tma_store_wait_read0();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();   // no CTA exits while its peer may still multicast into / arrive on it
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int kCluster, bool kTS>
void launch_ffn_t(const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2, const CUtensorMap& tmO,
                  const FfnParams& p, cudaStream_t stream) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(ffn_kernel<kCluster, kTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  }
  const int tiles = p.n_seq * p.tiles_per_seq;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3((tiles + kCluster - 1) / kCluster * kCluster);   // odd count: the extra CTA sees zero rows, stores clip
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, ffn_kernel<kCluster, kTS>, tmX, tmW1, tmW2, tmO, p);
}

}  // namespace

// variant: 1 = SS, 2 = SS + 2-CTA multicast, 3 = TS (P in TMEM), 4 = TS + 2-CTA multicast
void launch_ffn(const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2, const CUtensorMap& tmO,
                const FfnParams& p, int variant, cudaStream_t stream) {
  switch (variant) {
    case 1: launch_ffn_t<1, false>(tmX, tmW1, tmW2, tmO, p, stream); break;
    case 2: launch_ffn_t<2, false>(tmX, tmW1, tmW2, tmO, p, stream); break;
    case 3: launch_ffn_t<1, true>(tmX, tmW1, tmW2, tmO, p, stream); break;
    default: launch_ffn_t<2, true>(tmX, tmW1, tmW2, tmO, p, stream); break;
  }
}

}  // namespace fseend
