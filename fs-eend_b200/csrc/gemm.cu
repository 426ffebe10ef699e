// tcgen05 row-tile GEMM with fused epilogues (see gemm.cuh).
//
// One CTA (128 threads) computes one 128 x 256 output tile:
//   thread 0   : TMA producer  (A tile 128x64 + W tile 256x64 per k-block, 2-stage mbarrier ring)
//   thread 32  : MMA issuer    (tcgen05.mma kind::f16, M=128 N=256 K=16, accumulator = 256 TMEM columns)
//   all 4 warps: epilogue      (tcgen05.ld: thread r owns output row r -> row-local LN / L2 / bias / ReLU,
//                               fp16 pack into a 128B-swizzled staging tile, TMA store)
// Two CTAs are resident per SM (2 x ~97 KB smem, 2 x 256 TMEM columns): one CTA's epilogue overlaps the
// other's main loop.
#include "gemm.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kStages = 2;
constexpr int kABytes = BM * BK * 2;          // 16 KB
constexpr int kBBytes = BN * BK * 2;          // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSubTileBytes = BM * 64 * 2;    // one 64-column staging sub-tile, 16 KB
constexpr int kSmemBytes = kStages * kStageBytes + 1024;  // + alignment slack
constexpr uint32_t kTmemCols = 256;

struct RowStats {
  float n, mean, m2;
};

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

// 32 consecutive fp16 columns [c*32, c*32+32) of row r in the staging tile (4 sub-tiles of 64 columns).
__device__ __forceinline__ void staging_read32(const uint8_t* staging, int r, int c, float (&v)[32]) {
  const uint8_t* sub = staging + (c >> 1) * kSubTileBytes;
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
__device__ __forceinline__ void staging_write32(uint8_t* staging, int r, int c, const float (&v)[32]) {
  uint8_t* sub = staging + (c >> 1) * kSubTileBytes;
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

__device__ __forceinline__ void tmem_ld32_sync(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(128, 2)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
            const __grid_constant__ CUtensorMap tmO2, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ __align__(8) uint64_t res_bar;
  __shared__ uint32_t tmem_base_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int total_it = p.taps * p.k_blocks;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer
      for (int it = 0; it < total_it; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 1);
        const int tap = it / p.k_blocks;
        const int kb = it - tap * p.k_blocks;
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + kABytes;
        mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
        tma_load_3d(sa, &tmA, &full_bar[s], kb * BK, t0 + tap + p.tap_shift + p.a_row_offset, seq);
        tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, tap * (p.n_tiles * BN) + n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
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
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          // advance 16 fp16 (32 bytes) along K inside the 128-byte swizzle row: +2 in 16-byte units
          umma_f16(tmem_base, adesc + 2 * kk, bdesc + 2 * kk, idesc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage when the MMAs above have read it
      }
      umma_commit(&tmem_full_bar);   // accumulator complete
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

  const int r = tid;
  const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  float acc[32], aux[32];

  if (p.mode == EPI_BIAS) {
    for (int c = 0; c < 8; ++c) {
      tmem_ld32_sync(trow + c * 32, acc);
      if (p.bias) {
        load_vec32(p.bias + n0 + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
      }
      if (p.relu == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = fmaxf(acc[i], 0.f);
      } else if (p.relu == ACT_SWISH) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = acc[i] / (1.f + __expf(-acc[i]));
      }
      staging_write32(staging, r, c, acc);
    }
  } else if (p.mode == EPI_GLU) {
    // columns [0,128) = value, [128,256) = gate of the same 128 output channels
    for (int c = 0; c < 4; ++c) {
      float gate[32];
      tmem_ld32_sync(trow + c * 32, acc);
      tmem_ld32_sync(trow + 128 + c * 32, gate);
      if (p.bias) {
        load_vec32(p.bias + n0 + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
        load_vec32(p.bias + n0 + 128 + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) gate[i] += aux[i];
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = acc[i] / (1.f + __expf(-gate[i]));
      staging_write32(staging, r, c, acc);
    }
  } else if (p.mode == EPI_LN || p.mode == EPI_L2 || p.mode == EPI_RESID) {
    const float alpha = (p.mode == EPI_RESID) ? p.alpha : 1.f;
    const bool do_ln = (p.mode == EPI_LN) || (p.mode == EPI_RESID && p.ln_g != nullptr);
    const bool need_stats = do_ln || p.mode == EPI_L2;
    // pass 1: row statistics (Chan's parallel merge of 32-column chunks: robust to large means)
    RowStats st{0.f, 0.f, 0.f};
    float sumsq = 0.f;
    for (int c = 0; need_stats && c < 8; ++c) {
      tmem_ld32_sync(trow + c * 32, acc);
      if (p.bias) {
        load_vec32(p.bias + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] *= alpha;
      if (p.has_residual) {
        staging_read32(staging, r, c, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
      }
      if (p.mode == EPI_L2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) sumsq = fmaf(acc[i], acc[i], sumsq);
      } else {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) s += acc[i];
        const float cm = s * (1.f / 32.f);
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = acc[i] - cm;
          m2 = fmaf(d, d, m2);
        }
        const float nn = st.n + 32.f;
        const float delta = cm - st.mean;
        st.m2 += m2 + delta * delta * (st.n * 32.f / nn);
        st.mean += delta * (32.f / nn);
        st.n = nn;
      }
    }
    float mean = 0.f, scale = 1.f;
    if (p.mode == EPI_L2) {
      scale = sumsq > 0.f ? rsqrtf(sumsq) : 0.f;
    } else if (do_ln) {
      mean = st.mean;
      scale = rsqrtf(st.m2 * (1.f / 256.f) + p.ln_eps);
    }
    const bool zero_row = p.seq_len != nullptr && (t0 + r) >= p.seq_len[seq];
    // pass 2: normalise, affine, pack
    for (int c = 0; c < 8; ++c) {
      tmem_ld32_sync(trow + c * 32, acc);
      if (p.bias) {
        load_vec32(p.bias + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] *= alpha;
      if (p.has_residual) {
        staging_read32(staging, r, c, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = (acc[i] - mean) * scale;
      if (do_ln) {
        load_vec32(p.ln_g + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] *= aux[i];
        load_vec32(p.ln_b + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
      }
      if (zero_row) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
      }
      staging_write32(staging, r, c, acc);
    }
  }

  if (p.mode == EPI_GLU) {
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      for (int sub = 0; sub < 2; ++sub)
        tma_store_3d(&tmO, staging + sub * kSubTileBytes, n_tile * 128 + sub * 64, t0, seq);
      tma_store_commit();
      tma_store_wait_read0();
    }
  } else if (p.mode != EPI_CONVERT) {
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      for (int sub = 0; sub < 4; ++sub)
        tma_store_3d(&tmO, staging + sub * kSubTileBytes, n0 + sub * 64, t0, seq);
      tma_store_commit();
      tma_store_wait_read0();
    }
    if (p.ln2_g != nullptr && (p.mode == EPI_LN || p.mode == EPI_RESID)) {
      // second output: LayerNorm of the (fp16) row just stored, with the next pre-norm sub-layer's affine
      __syncthreads();   // the TMA store above has finished reading the staging tile
      RowStats s2{0.f, 0.f, 0.f};
      for (int c = 0; c < 8; ++c) {
        staging_read32(staging, r, c, acc);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) s += acc[i];
        const float cm = s * (1.f / 32.f);
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = acc[i] - cm;
          m2 = fmaf(d, d, m2);
        }
        const float nn = s2.n + 32.f;
        const float delta = cm - s2.mean;
        s2.m2 += m2 + delta * delta * (s2.n * 32.f / nn);
        s2.mean += delta * (32.f / nn);
        s2.n = nn;
      }
      const float rstd2 = rsqrtf(s2.m2 * (1.f / 256.f) + p.ln_eps);
      for (int c = 0; c < 8; ++c) {
        staging_read32(staging, r, c, acc);
        load_vec32(p.ln2_g + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = (acc[i] - s2.mean) * rstd2 * aux[i];
        load_vec32(p.ln2_b + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
        staging_write32(staging, r, c, acc);
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        for (int sub = 0; sub < 4; ++sub)
          tma_store_3d(&tmO2, staging + sub * kSubTileBytes, n0 + sub * 64, t0, seq);
        tma_store_commit();
        tma_store_wait_read0();
      }
    }
  } else {
    // attractor init: S output rows per input row, out[row, s, :] = acc + pe_proj[s, :]
    const int row0 = seq * p.rows_per_seq + t0;
    for (int s = 0; s < p.S; ++s) {
      for (int c = 0; c < 8; ++c) {
        tmem_ld32_sync(trow + c * 32, acc);
        load_vec32(p.pe_proj + s * 256 + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
        staging_write32(staging, r, c, acc);
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        for (int sub = 0; sub < 4; ++sub) tma_store_3d(&tmO, staging + sub * kSubTileBytes, sub * 64, s, row0);
        tma_store_commit();
        tma_store_wait_read0();
      }
      __syncthreads();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

void launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const CUtensorMap& tmO,
                 const GemmParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    attr_set = true;
  }
  const int grid = p.n_seq * p.tiles_per_seq * p.n_tiles;
  gemm_kernel<<<grid, 128, kSmemBytes, stream>>>(tmA, tmB, tmR, tmO, tmO, p);
}

void launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const CUtensorMap& tmO,
                  const CUtensorMap& tmO2, const GemmParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    attr_set = true;
  }
  const int grid = p.n_seq * p.tiles_per_seq * p.n_tiles;
  gemm_kernel<<<grid, 128, kSmemBytes, stream>>>(tmA, tmB, tmR, tmO, tmO2, p);
}

}  // namespace fseend
