// tcgen05 row-tile GEMM with fused epilogues (see gemm.cuh).
//
// One CTA (256 threads) computes one 128 x 256 output tile:
//   thread 0   : TMA producer  (A tile 128x64 + W tile 256x64 per k-block, 2-stage mbarrier ring)
//   thread 32  : MMA issuer    (tcgen05.mma kind::f16, M=128 N=256 K=16, accumulator = 256 TMEM columns)
//   all 8 warps: epilogue      (tcgen05.ld: two threads per output row, 128 columns each; row-local LN / L2 /
//                               bias / activation with the row statistics merged through smem; fp16 pack into a
//                               128B-swizzled staging tile, TMA store).  The row epilogues are instruction/latency
//                               bound, so they get all the warps the register file allows at 2 CTAs/SM.
// Two CTAs are resident per SM (2 x ~97 KB smem, 2 x 256 TMEM columns): one CTA's epilogue overlaps the
// other's main loop.
#include "once.h"
#include <stdlib.h>

#include "gemm.cuh"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kStages = 2;
constexpr int kABytes = BM * BK * 2;          // 16 KB
constexpr int kBBytes = BN * BK * 2;          // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
using gemm_detail::kSubTileBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 1024;  // + alignment slack
constexpr uint32_t kTmemCols = 256;

__global__ void __launch_bounds__(256, 2)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
            const __grid_constant__ CUtensorMap tmO2, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ __align__(8) uint64_t res_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) gemm_detail::EpiParams epi_params;
  __shared__ __align__(16) float4 xchg[2 * 128];

  // 1024-byte alignment as an OFFSET from the shared array: pointer arithmetic through uintptr_t would hide the shared
  // address space from the compiler and turn every access below into a generic load / store
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* staging = smem;  // aliases the pipeline stages; only touched after every MMA has completed

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  const int tile = blockIdx.x;
  const int n_tile = tile % p.n_tiles;
  const int m_tile = tile / p.n_tiles;
  const int seq = m_tile / p.tiles_per_seq;
  const int t0 = (m_tile % p.tiles_per_seq) * BM;
  const int n0 = n_tile * BN;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&res_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, kTmemCols);
  gemm_detail::load_epi_params(epi_params, p, n0, tid, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int total_it = p.taps * p.k_blocks;

  // Role warps stay converged; one elected lane issues the TMA / tcgen05 instructions (uniform-register operands).
  if (warp == 0) {
    {
      // ---------------- TMA producer
      const int a_row_offset = p.a_row_offset + (p.a_row_offset_dev ? *p.a_row_offset_dev : 0);
      for (int it = 0; it < total_it; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 1);
        const int tap = it / p.k_blocks;
        const int kb = it - tap * p.k_blocks;
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + kABytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
          tma_load_3d(sa, &tmA, &full_bar[s], kb * BK, t0 + tap + p.tap_shift + a_row_offset, seq);
          tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, tap * (p.n_tiles * BN) + n0);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      // ---------------- MMA issuer
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, false);
      for (int it = 0; it < total_it; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_bar[s], ph, 2);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * kStageBytes);
        const uint64_t adesc = smem_desc_sw128(sa);
        const uint64_t bdesc = smem_desc_sw128(sa + kABytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            // advance 16 fp16 (32 bytes) along K inside the 128-byte swizzle row: +2 in 16-byte units
            umma_f16(tmem_base, adesc + 2 * kk, bdesc + 2 * kk, idesc, (it > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees this smem stage when the MMAs above have read it
          if (it == total_it - 1) umma_commit(&tmem_full_bar);   // accumulator complete
        }
        __syncwarp();
      }
    }
    __syncwarp();
  }

  // ---------------- epilogue: all 128 threads, thread r <-> accumulator row r (TMEM lane r)
  mbar_wait(&tmem_full_bar, 0, 3);
  tc_fence_after();

  if (p.has_residual) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&res_bar, 4 * kSubTileBytes);
      for (int sub = 0; sub < 4; ++sub)
        tma_load_3d(staging + sub * kSubTileBytes, &tmR, &res_bar, sub * 64, t0, seq);
    }
    mbar_wait(&res_bar, 0, 4);
  }

  {
    const int quarter = warp & 3, half = warp >> 2;       // TMEM lane quarter; column half owned by this thread
    gemm_detail::row_tile_epilogue<2>(p, epi_params, tmO, tmO2, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16),
                                      staging, quarter * 32 + lane, tid == 0, n0, n_tile, t0, seq,
                                      [] { __syncthreads(); }, half, xchg);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

// gemm_persist.cu
void launch_gemm_persist(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const CUtensorMap& tmO,
                         const CUtensorMap& tmO2, const GemmParams& p, cudaStream_t stream);
// gemm_pair.cu
bool gemm_pair_supported(const GemmParams& p);
void launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmR, const CUtensorMap& tmO, const GemmParams& p,
                      cudaStream_t stream);

static int env_switch(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0] >= '0' && e[0] <= '9') ? e[0] - '0' : dflt;
}
// FSEEND_GEMM_PERSIST=0 selects the one-tile-per-CTA kernel below instead of the persistent one (gemm_persist.cu).
static bool use_persist() {
  static const int v = env_switch("FSEEND_GEMM_PERSIST", 1);
  return v == 1;
}
// FSEEND_GEMM_PAIR: 0 = never use the weight-stationary CTA-pair kernel (gemm_pair.cu), 1 = where it pays (default),
// 2 = wherever it is structurally possible (tests).
static int pair_mode() { return env_switch("FSEEND_GEMM_PAIR", 1); }   // read per launch: tests flip it

// Measured on B200 (B=64, T=500, S=6): the persistent kernel wins on the wide bias/activation projections
// (QKV 0.142 -> 0.113 ms) where many output tiles stream per SM; the row-epilogue GEMMs (residual + LayerNorm, L2,
// attractor broadcast) are epilogue-latency bound and run faster as two independent CTAs per SM.
static bool persist_pays(const GemmParams& p) {
  const int items = p.n_seq * p.tiles_per_seq * p.n_tiles;
  return (p.mode == EPI_BIAS || p.mode == EPI_GLU) && items >= 296;
}
// The pair kernel needs the weight n-tile to stay resident (K <= 256, one tap) and about one row-tile pair per cluster
// to amortise loading it (FSEEND_GEMM_PAIR_MIN overrides the item threshold).
static bool pair_pays(const GemmParams& p) {
  if (pair_mode() == 0 || !gemm_pair_supported(p)) return false;
  if (pair_mode() == 2) return true;
  const int items = (p.n_seq * p.tiles_per_seq + 1) / 2 * p.n_tiles;
  static const int min_items = [] {
    const char* e = getenv("FSEEND_GEMM_PAIR_MIN");
    return e ? atoi(e) : 74;   // one item per cluster: still ahead of the one-tile kernel (enc out-proj 23 -> 19 us)
  }();
  return items >= min_items;
}

void launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const CUtensorMap& tmO,
                  const CUtensorMap& tmO2, const GemmParams& p, cudaStream_t stream) {
  const bool plain_only = p.a_row_offset_dev != nullptr;   // only gemm_kernel reads the device-resident row offset
  if (!plain_only && pair_pays(p)) {
    launch_gemm_pair(tmA, tmR, tmO, p, stream);
    return;
  }
  if (!plain_only && use_persist() && persist_pays(p)) {
    launch_gemm_persist(tmA, tmB, tmR, tmO, tmO2, p, stream);
    return;
  }
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  }
  const int grid = p.n_seq * p.tiles_per_seq * p.n_tiles;
  gemm_kernel<<<grid, 256, kSmemBytes, stream>>>(tmA, tmB, tmR, tmO, tmO2, p);
}

void launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const CUtensorMap& tmO,
                 const GemmParams& p, cudaStream_t stream) {
  launch_gemm2(tmA, tmB, tmR, tmO, tmO, p, stream);
}

}  // namespace fseend
