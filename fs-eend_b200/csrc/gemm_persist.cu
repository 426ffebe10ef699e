// Persistent row-tile GEMM (same contract and epilogues as gemm.cu, see gemm.cuh).
//
// The one-tile-per-CTA kernel runs load -> MMA -> epilogue -> store serially inside each CTA; with two CTAs per SM that
// keeps only ~1/3 of the HBM/L2 pipes busy on the short-K projections, which are memory-bound (2*K FLOP per 2-byte
// output element at K = 256).  Here one CTA per SM loops over its (m-tile, n-tile) items with three decoupled roles:
//   warp 0     : TMA producer — runs ahead across items through a 3-stage ring of (A 128x64, W 256x64) k-blocks
//                (144 KB in flight) and prefetches the residual tile of the next item into the staging buffer;
//   warp 1     : tcgen05 issuer — two 256-column TMEM accumulators, so the MMAs of item i+1 overlap the epilogue of i;
//   warps 2-9  : epilogue (gemm_epilogue.cuh), two threads per row (128 columns each; the row epilogues are
//                instruction/latency-bound, not memory-bound) -> 64 KB staging tile -> TMA store.
#include "once.h"
#include "gemm.cuh"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kStages = 3;
constexpr int kABytes = BM * BK * 2;          // 16 KB
constexpr int kBBytes = BN * BK * 2;          // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kOffStage = kStages * kStageBytes;                               // 144 KB
constexpr int kSmemBytes = kOffStage + 4 * gemm_detail::kSubTileBytes;         // + 64 KB staging = 208 KB
constexpr uint32_t kTmemCols = 512;

__global__ void __launch_bounds__(320, 1)
gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                    const __grid_constant__ CUtensorMap tmO2, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], tmem_full[2], tmem_empty[2], res_full,
      stg_free;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) gemm_detail::EpiParams epi_params;
  __shared__ __align__(16) float4 xchg[2 * 128];
  uint8_t* staging = smem + kOffStage;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n_items = p.n_seq * p.tiles_per_seq * p.n_tiles;
  const int total_it = p.taps * p.k_blocks;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[fseend] gemm: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    mbar_init(&res_full, 1);
    mbar_init(&stg_free, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // Role warps stay converged and elect one lane only around the TMA / tcgen05 instructions, so descriptors and loop
  // state live in uniform registers (a single diverged lane makes ptxas wrap each UTCHMMA in an ELECT loop + R2UR moves).
  if (warp == 0) {
    {
      // ------------------------------------------------------------------ TMA producer
      uint32_t u = 0, n = 0;
      for (int id = blockIdx.x; id < n_items; id += gridDim.x, ++n) {
        const int n_tile = id % p.n_tiles, m_tile = id / p.n_tiles;
        const int seq = m_tile / p.tiles_per_seq;
        const int t0 = (m_tile % p.tiles_per_seq) * BM;
        const int n0 = n_tile * BN;
        for (int it = 0; it < total_it; ++it, ++u) {
          const int s = u % kStages;
          mbar_wait(&empty_bar[s], ((u / kStages) & 1) ^ 1, 71);
          const int tap = it / p.k_blocks;
          const int kb = it - tap * p.k_blocks;
          uint8_t* sa = smem + s * kStageBytes;
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
            tma_load_3d(sa, &tmA, &full_bar[s], kb * BK, t0 + tap + p.tap_shift + p.a_row_offset, seq);
            tma_load_2d(sa + kABytes, &tmB, &full_bar[s], kb * BK, tap * (p.n_tiles * BN) + n0);
          }
          __syncwarp();
        }
        if (p.has_residual) {
          mbar_wait(&stg_free, (n & 1) ^ 1, 72);   // the previous item's stores have drained the staging tile
          if (elect_one()) {
            mbar_arrive_expect_tx(&res_full, 4 * gemm_detail::kSubTileBytes);
            for (int sub = 0; sub < 4; ++sub)
              tma_load_3d(staging + sub * gemm_detail::kSubTileBytes, &tmR, &res_full, n0 + sub * 64, t0, seq);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      // ------------------------------------------------------------------ MMA issuer (warp-converged)
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, false);
      uint32_t u = 0, n = 0;
      for (int id = blockIdx.x; id < n_items; id += gridDim.x, ++n) {
        const int acc = n & 1;
        mbar_wait(&tmem_empty[acc], ((n >> 1) & 1) ^ 1, 73);
        tc_fence_after();
        const uint32_t tmem_D = tmem_base + acc * 256;
        for (int it = 0; it < total_it; ++it, ++u) {
          const int s = u % kStages;
          mbar_wait(&full_bar[s], (u / kStages) & 1, 74);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * kStageBytes);
          const uint64_t adesc = smem_desc_sw128(sa);
          const uint64_t bdesc = smem_desc_sw128(sa + kABytes);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              umma_f16(tmem_D, adesc + 2 * kk, bdesc + 2 * kk, idesc, (it > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&empty_bar[s]);
            if (it == total_it - 1) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else {
    // -------------------------------------------------------------------- epilogue warps 2..9
    const int et = tid - 64;
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;          // warps 2-5: columns [0,128), warps 6-9: [128,256)
    const int r = quarter * 32 + lane;
    const bool store_thread = (et == 0);
    auto sync = [] { named_bar_sync(1, 256); };
    uint32_t n = 0;
    int cur_n0 = -1;
    for (int id = blockIdx.x; id < n_items; id += gridDim.x, ++n) {
      const int n_tile = id % p.n_tiles, m_tile = id / p.n_tiles;
      const int seq = m_tile / p.tiles_per_seq;
      const int t0 = (m_tile % p.tiles_per_seq) * BM;
      const int n0 = n_tile * BN;
      if (n0 != cur_n0) {   // per-column parameters of this n-tile (bias depends on n0; LN vectors loaded once)
        sync();
        gemm_detail::load_epi_params(epi_params, p, n0, et, 256);
        sync();
        cur_n0 = n0;
      }
      const int acc = n & 1;
      mbar_wait(&tmem_full[acc], (n >> 1) & 1, 75);
      if (p.has_residual) mbar_wait(&res_full, n & 1, 76);
      tc_fence_after();
      const uint32_t trow = tmem_base + acc * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
      gemm_detail::row_tile_epilogue<2>(p, epi_params, tmO, tmO2, trow, staging, r, store_thread, n0, n_tile, t0, seq,
                                        sync, half, xchg);
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);   // this thread's TMEM reads of the accumulator are complete
      sync();                          // the store thread has waited for its TMA stores to finish reading `staging`
      if (store_thread && p.has_residual) mbar_arrive(&stg_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

void launch_gemm_persist(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const CUtensorMap& tmO,
                         const CUtensorMap& tmO2, const GemmParams& p, cudaStream_t stream) {
  static int num_sms = 0;
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(gemm_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_items = p.n_seq * p.tiles_per_seq * p.n_tiles;
  const int grid = n_items < num_sms ? n_items : num_sms;
  gemm_persist_kernel<<<grid, 320, kSmemBytes, stream>>>(tmA, tmB, tmR, tmO, tmO2, p);
}

}  // namespace fseend
