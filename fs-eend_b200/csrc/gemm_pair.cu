// Weight-stationary row-tile GEMM on CTA pairs (tcgen05 cta_group::2) — same contract and epilogues as gemm.cu.
//
// The one-tile kernels re-load the 256 x K weight tile (128 KB at K = 256) for every 128-row output tile: the decoder's
// QKV projection moves 864 MB from L2 into the SMs to produce 295 MB (7.5 TB/s of L2->SM traffic — the bound).  Here a
// cluster of two CTAs owns ONE n-tile for the whole launch: each CTA keeps its half of that weight tile (128 of the 256
// output features x K, 64 KB) resident in shared memory and only the activations stream (one 16 KB k-block per stage
// and CTA).  Every MMA is M = 256 (128 rows from each CTA) x N = 256 x K = 16; L2->SM traffic drops 3x.
//   warp 0     : TMA producer (both CTAs): resident W half once, then A k-blocks through a 5-stage ring; completion is
//                signalled on the LEADER's barriers.
//   warp 1     : tcgen05 issuer (leader CTA only), two 256-column TMEM accumulators.
//   warps 2-9  : epilogue of this CTA's 128 rows, two threads per row (128 columns each).  Results are packed to fp16
//                in registers and the staging tile is only touched once the PREVIOUS item's TMA store has drained it,
//                so the store of item i overlaps the TMEM reads and math of item i+1.
// The residual of the LayerNorm GEMMs is added BY THE TENSOR CORE: the residual tile streams through the same ring as
// the activations and is multiplied by a 64 x 64 identity (4 KB per CTA) into the accumulator (four N = 64 MMA groups
// per item; fp16 x 1.0 accumulates exactly in fp32).  That keeps the residual on the deep TMA prefetch path instead of
// a strided per-thread global load or a second staging buffer that shared memory has no room for.
// Epilogues: EPI_BIAS (+ReLU/Swish) and EPI_LN (bias, optional residual, zero rows beyond seq_len).
#include "once.h"
#include "gemm.cuh"
#include "gemm_epilogue.cuh"
#include "pair.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

using namespace pair;

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kStages = 5;
constexpr int kKBlockBytes = BM * BK * 2;                 // 16 KB: [128 rows][64 k] fp16 (A stage, or one W-half k-block)
constexpr int kMaxKBlocks = 4;                            // resident W half: up to K = 256
constexpr int kOffW = 0;
constexpr int kOffA = kMaxKBlocks * kKBlockBytes;         // 64 KB
constexpr int kOffStage = kOffA + kStages * kKBlockBytes; // 144 KB
constexpr int kOffIdent = kOffStage + 4 * gemm_detail::kSubTileBytes;    // + 64 KB staging
constexpr int kIdentBytes = 32 * 128;                     // this CTA's 32 rows of the 64 x 64 identity, [32][64 k]
constexpr int kSmemBytes = kOffIdent + kIdentBytes;       // 212 KB
constexpr int kResBlocks = 4;                             // residual k-blocks per item (N = 256)
constexpr uint32_t kTmemCols = 512;

// kLn: EPI_LN (else EPI_BIAS); kRes: residual accumulated by identity MMAs.  Compile-time so that each variant keeps
// its working set in registers (a runtime-mode version spilled ~60 registers to local memory, and with the
// shared-memory carve-out this kernel needs, local memory lives in L2).
template <bool kLn, bool kRes>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                 const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t w_full, a_full[kStages], a_empty[kStages], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) gemm_detail::EpiParams epi_params;
  __shared__ __align__(16) float4 xchg[2 * 128];
  uint8_t* staging = smem + kOffStage;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);

  // work split: cluster c owns n-tile (c mod n_tiles) and every cnt-th pair of row tiles of that n-tile
  const int n_clusters = gridDim.x >> 1;
  const int cid = blockIdx.x >> 1;
  const int m_tiles = p.n_seq * p.tiles_per_seq;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int n_tile = cid % p.n_tiles;
  const int first = cid / p.n_tiles;
  const int cnt = (n_clusters - n_tile + p.n_tiles - 1) / p.n_tiles;
  const int n0 = n_tile * BN;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[fseend] gemm_pair: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    mbar_init(&w_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&a_full[i], 1);     // the leader's arrive.expect_tx; bytes from both CTAs' loads
      mbar_init(&a_empty[i], 1);    // one multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 512);   // every epilogue thread of both CTAs (the leader's copy is waited on)
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmBh);
    tma_prefetch_desc(&tmO);
  }
  if (kRes) {
    // rows [32 rank, +32) of the 64 x 64 identity as a K-major 128B-swizzled B tile (N = 64 across the pair)
    for (int i = tid; i < 32 * 8; i += blockDim.x) {
      const int row = i >> 3, chunk = i & 7;
      const int one = static_cast<int>(rank) * 32 + row - chunk * 8;   // position of the 1.0 inside this 8-half chunk
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (one >= 0 && one < 8) reinterpret_cast<uint16_t*>(&v)[one] = 0x3C00;   // fp16 1.0
      *reinterpret_cast<uint4*>(smem + kOffIdent + sw128_offset(row, chunk)) = v;
    }
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc_pair(&tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // peer barriers initialised and both TMEM allocations done before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected lane)
    const uint32_t a_full_leader = mapa_rank(smem_u32(&a_full[0]), 0);
    const uint32_t w_full_leader = mapa_rank(smem_u32(&w_full), 0);
    if (elect_one()) {
      if (leader) mbar_arrive_expect_tx(&w_full, 2 * p.k_blocks * kKBlockBytes);
      for (int kb = 0; kb < p.k_blocks; ++kb)
        tma_load_2d_pair(smem + kOffW + kb * kKBlockBytes, &tmBh, w_full_leader, kb * BK,
                         n0 + static_cast<int>(rank) * 128);
    }
    __syncwarp();
    uint32_t u = 0;
    for (int mp = first; mp < m_pairs; mp += cnt) {
      const int m_tile = 2 * mp + static_cast<int>(rank);   // past the end (odd tile count): loads zero-fill
      const int seq = m_tile / p.tiles_per_seq;
      const int t0 = (m_tile % p.tiles_per_seq) * BM;
      for (int kb = 0; kb < p.k_blocks; ++kb, ++u) {
        const int s = u % kStages;
        mbar_wait(&a_empty[s], ((u / kStages) & 1) ^ 1, 171);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&a_full[s], 2 * kKBlockBytes);
          tma_load_3d_pair(smem + kOffA + s * kKBlockBytes, &tmA, a_full_leader + s * 8, kb * BK,
                           t0 + p.tap_shift + p.a_row_offset, seq);
        }
        __syncwarp();
      }
      for (int rb = 0; kRes && rb < kResBlocks; ++rb, ++u) {   // residual tile: 4 more [128][64] blocks
        const int s = u % kStages;
        mbar_wait(&a_empty[s], ((u / kStages) & 1) ^ 1, 172);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&a_full[s], 2 * kKBlockBytes);
          tma_load_3d_pair(smem + kOffA + s * kKBlockBytes, &tmR, a_full_leader + s * 8, n0 + rb * 64, t0, seq);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane)
      constexpr uint32_t idesc = make_idesc_f16(2 * BM, BN, false);
      mbar_wait(&w_full, 0, 173);
      uint32_t u = 0, n = 0;
      for (int mp = first; mp < m_pairs; mp += cnt, ++n) {
        const int acc = n & 1;
        mbar_wait_cluster(&tmem_empty[acc], ((n >> 1) & 1) ^ 1, 174);
        tc_fence_after();
        const uint32_t tmem_D = tmem_base + acc * 256;
        for (int kb = 0; kb < p.k_blocks; ++kb, ++u) {
          const int s = u % kStages;
          mbar_wait(&a_full[s], (u / kStages) & 1, 175);
          tc_fence_after();
          const uint64_t adesc = smem_desc_sw128(smem_u32(smem + kOffA + s * kKBlockBytes));
          const uint64_t bdesc = smem_desc_sw128(smem_u32(smem + kOffW + kb * kKBlockBytes));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              umma2_f16(tmem_D, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            umma2_commit(&a_empty[s]);
            if (!kRes && kb == p.k_blocks - 1) umma2_commit(&tmem_full[acc]);
          }
          __syncwarp();
        }
        for (int rb = 0; kRes && rb < kResBlocks; ++rb, ++u) {
          // D[:, 64 rb .. +64) += R[:, 64 rb .. +64) * I64
          constexpr uint32_t idesc_r = make_idesc_f16(2 * BM, 64, false);
          const int s = u % kStages;
          mbar_wait(&a_full[s], (u / kStages) & 1, 178);
          tc_fence_after();
          const uint64_t adesc = smem_desc_sw128(smem_u32(smem + kOffA + s * kKBlockBytes));
          const uint64_t idesc_tile = smem_desc_sw128(smem_u32(smem + kOffIdent));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              umma2_f16(tmem_D + rb * 64, adesc + 2 * kk, idesc_tile + 2 * kk, idesc_r, 1u);
            umma2_commit(&a_empty[s]);
            if (rb == kResBlocks - 1) umma2_commit(&tmem_full[acc]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps 2..9
    using namespace gemm_detail;
    const int et = tid - 64;
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;          // warps 2-5: columns [0,128), warps 6-9: [128,256)
    const int r = quarter * 32 + lane;
    const bool store_thread = (et == 0);
    const uint32_t tmem_empty_leader = mapa_rank(smem_u32(&tmem_empty[0]), 0);
    auto sync = [] { named_bar_sync(1, 256); };
    load_epi_params(epi_params, p, n0, et, 256);   // this cluster's n-tile never changes
    sync();
    constexpr bool ln = kLn;
    uint32_t n = 0;
    for (int mp = first; mp < m_pairs; mp += cnt, ++n) {
      const int m_tile = 2 * mp + static_cast<int>(rank);
      const int seq = m_tile / p.tiles_per_seq;
      const int t0 = (m_tile % p.tiles_per_seq) * BM;
      const bool valid = m_tile < m_tiles;      // CTA-uniform
      const int acc = n & 1;
      uint32_t pk[ln ? 1 : 64];   // EPI_BIAS: this thread's 128 output columns as packed fp16 pairs
      mbar_wait(&tmem_full[acc], (n >> 1) & 1, 176);
      tc_fence_after();
      if (valid) {
        const uint32_t trow = tmem_base + acc * 256 + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
        float a[32], aux[32];
        if (!ln) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            tmem_ld32_sync(trow + c * 32, a);
            smem_vec32(epi_params.bias + half * 128 + c * 32, aux);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] += aux[i];
            if (p.relu == ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) a[i] = fmaxf(a[i], 0.f);
            } else if (p.relu == ACT_SWISH) {
#pragma unroll
              for (int i = 0; i < 32; ++i) a[i] = a[i] / (1.f + __expf(-a[i]));
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[c * 16 + i] = pack_half2(a[2 * i], a[2 * i + 1]);
          }
        } else {
          // pass 1: row statistics of z = acc + bias (the residual is already in the accumulator); Chan's merge of
          // 32-column chunks, then of the two half rows
          RowStats st{0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            tmem_ld32_sync(trow + c * 32, a);
            smem_vec32(epi_params.bias + half * 128 + c * 32, aux);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] += aux[i];
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) sum += a[i];
            const float cm = sum * (1.f / 32.f);
            float m2 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float d = a[i] - cm;
              m2 = fmaf(d, d, m2);
            }
            const float nn = st.n + 32.f;
            const float delta = cm - st.mean;
            st.m2 += m2 + delta * delta * (st.n * 32.f / nn);
            st.mean += delta * (32.f / nn);
            st.n = nn;
          }
          xchg[half * 128 + r] = make_float4(st.n, st.mean, st.m2, 0.f);
          if (store_thread) tma_store_wait_read0();   // the previous item's stores have drained the staging tile
          sync();
          const float4 o = xchg[(half ^ 1) * 128 + r];
          sync();
          const float dm = o.y - st.mean;
          const float mean = 0.5f * (st.mean + o.y);
          const float rstd = rsqrtf((st.m2 + o.z + dm * dm * 64.f) * (1.f / 256.f) + p.ln_eps);
          const bool zero_row = p.seq_len != nullptr && (t0 + r) >= p.seq_len[seq];
          // pass 2: normalise, affine, pack straight into the staging tile (free since the sync above)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            tmem_ld32_sync(trow + c * 32, a);
            smem_vec32(epi_params.bias + half * 128 + c * 32, aux);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] += aux[i];
            smem_vec32(epi_params.g + half * 128 + c * 32, aux);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = (a[i] - mean) * rstd * aux[i];
            smem_vec32(epi_params.b + half * 128 + c * 32, aux);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = zero_row ? 0.f : a[i] + aux[i];
            staging_write32(staging, r, half * 4 + c, a);
          }
          fence_proxy_async_smem();
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(tmem_empty_leader + acc * 8);   // this thread's TMEM reads of the accumulator are complete
      if (!ln) {
        if (store_thread) tma_store_wait_read0();          // the previous item's stores have drained the staging tile
        sync();
        if (valid) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int cg = half * 4 + c;                     // 32-column chunk of the 256-wide tile
            uint8_t* sub = staging + (cg >> 1) * kSubTileBytes;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(sub + sw128_offset(r, (cg & 1) * 4 + q)) = make_uint4(
                  pk[c * 16 + 4 * q], pk[c * 16 + 4 * q + 1], pk[c * 16 + 4 * q + 2], pk[c * 16 + 4 * q + 3]);
          }
          fence_proxy_async_smem();
        }
      } else if (!valid) {
        if (store_thread) tma_store_wait_read0();
      }
      sync();
      if (store_thread && valid) {
        for (int sub = 0; sub < 4; ++sub) tma_store_3d(&tmO, staging + sub * kSubTileBytes, n0 + sub * 64, t0, seq);
        tma_store_commit();
      }
    }
    if (store_thread) tma_store_wait_read0();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA exits (or frees TMEM) while its peer may still signal it or the pair's MMAs run
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

}  // namespace

bool gemm_pair_supported(const GemmParams& p) {
  if (p.tmB_half == nullptr || p.taps != 1 || p.k_blocks < 1 || p.k_blocks > kMaxKBlocks) return false;
  if (p.mode == EPI_BIAS) return true;
  return p.mode == EPI_LN && p.ln2_g == nullptr;
}

void launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmR, const CUtensorMap& tmO, const GemmParams& p,
                      cudaStream_t stream) {
  static int num_sms = 0;
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(gemm_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    cudaFuncSetAttribute(gemm_pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    cudaFuncSetAttribute(gemm_pair_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int m_pairs = (p.n_seq * p.tiles_per_seq + 1) / 2;
  const int items = m_pairs * p.n_tiles;                 // >= n_tiles: every n-tile gets at least one cluster
  const int clusters = items < num_sms / 2 ? items : num_sms / 2;
  const dim3 grid(2 * clusters), block(320);
  if (p.mode == EPI_BIAS) gemm_pair_kernel<false, false><<<grid, block, kSmemBytes, stream>>>(tmA, *p.tmB_half, tmR, tmO, p);
  else if (p.has_residual) gemm_pair_kernel<true, true><<<grid, block, kSmemBytes, stream>>>(tmA, *p.tmB_half, tmR, tmO, p);
  else gemm_pair_kernel<true, false><<<grid, block, kSmemBytes, stream>>>(tmA, *p.tmB_half, tmR, tmO, p);
}

}  // namespace fseend
