// Row-tile epilogues shared by the GEMM kernels: thread r owns accumulator row r (TMEM lane r); results are packed
// to fp16 into a 128B-swizzled 64 KB staging tile (4 sub-tiles of 64 columns) and written with TMA stores.
#pragma once
#include "gemm.cuh"
#include "ptx.cuh"

namespace fseend {
namespace gemm_detail {

constexpr int kSubTileBytes = 128 * 64 * 2;    // one 64-column staging sub-tile, 16 KB

// Per-column epilogue parameters of the current 256-column tile, staged in shared memory once so that the row
// epilogues read them with broadcast LDS instead of dependent global loads (with a ~200 KB smem carve-out the L1 is
// tiny and every such load costs an L2 round trip).
struct EpiParams {
  float bias[256], g[256], b[256], g2[256], b2[256];
};
__device__ __forceinline__ void load_epi_params(EpiParams& sp, const GemmParams& p, int n0, int tid, int nthreads) {
  for (int i = tid; i < 256; i += nthreads) {
    sp.bias[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
    sp.g[i] = p.ln_g ? __ldg(p.ln_g + i) : 1.f;
    sp.b[i] = p.ln_b ? __ldg(p.ln_b + i) : 0.f;
    sp.g2[i] = p.ln2_g ? __ldg(p.ln2_g + i) : 1.f;
    sp.b2[i] = p.ln2_b ? __ldg(p.ln2_b + i) : 0.f;
  }
}
__device__ __forceinline__ void smem_vec32(const float* __restrict__ v, float (&out)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = reinterpret_cast<const float4*>(v)[i];   // same address in every lane: broadcast
    out[4 * i + 0] = t.x;
    out[4 * i + 1] = t.y;
    out[4 * i + 2] = t.z;
    out[4 * i + 3] = t.w;
  }
}

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


// `staging` holds the residual tile on entry when p.has_residual.  `sync()` synchronises the 128 epilogue threads;
// `store_thread` issues the TMA stores.  trow = TMEM address of this thread's lane quarter, column 0 of the tile.
// kHalves = 2: two threads share a row (same TMEM lane quarter, different warps); thread `half` owns columns
// [half*128, half*128+128) and the row statistics are merged through `xchg` ([2][128] float4 in shared memory).
template <int kHalves, class Sync>
__device__ __forceinline__ void row_tile_epilogue(const GemmParams& p, const EpiParams& sp, const CUtensorMap& tmO,
                                                  const CUtensorMap& tmO2, uint32_t trow, uint8_t* staging, int r,
                                                  bool store_thread, int n0, int n_tile, int t0, int seq, Sync sync,
                                                  int half = 0, float4* xchg = nullptr) {
  float acc[32], aux[32];
  const int c0 = half * (8 / kHalves), c1 = c0 + 8 / kHalves;        // 32-column chunks owned by this thread
  const int g0 = half * (4 / kHalves), g1 = g0 + 4 / kHalves;        // GLU output chunks
  // merge this thread's (n, mean, m2 | sumsq) with its partner's
  auto merge_stats = [&](RowStats& st, float& sumsq) {
    if (kHalves == 2) {
      xchg[half * 128 + r] = make_float4(st.n, st.mean, st.m2, sumsq);
      sync();
      const float4 o = xchg[(half ^ 1) * 128 + r];
      sync();
      const float nn = st.n + o.x;
      if (nn > 0.f) {
        const float delta = o.y - st.mean;
        st.m2 += o.z + delta * delta * (st.n * o.x / nn);
        st.mean += delta * (o.x / nn);
        st.n = nn;
      }
      sumsq += o.w;
    }
  };

  if (p.mode == EPI_BIAS) {
    for (int c = c0; c < c1; ++c) {
      tmem_ld32_sync(trow + c * 32, acc);
      if (p.bias) {
        smem_vec32(sp.bias + c * 32, aux);
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
    for (int c = g0; c < g1; ++c) {
      float gate[32];
      tmem_ld32_sync(trow + c * 32, acc);
      tmem_ld32_sync(trow + 128 + c * 32, gate);
      if (p.bias) {
        smem_vec32(sp.bias + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
        smem_vec32(sp.bias + 128 + c * 32, aux);
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
    for (int c = c0; need_stats && c < c1; ++c) {
      tmem_ld32_sync(trow + c * 32, acc);
      if (p.bias) {
        smem_vec32(sp.bias + c * 32, aux);
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
    if (need_stats) merge_stats(st, sumsq);
    float mean = 0.f, scale = 1.f;
    if (p.mode == EPI_L2) {
      scale = sumsq > 0.f ? rsqrtf(sumsq) : 0.f;
    } else if (do_ln) {
      mean = st.mean;
      scale = rsqrtf(st.m2 * (1.f / 256.f) + p.ln_eps);
    }
    const bool zero_row = p.seq_len != nullptr && (t0 + r) >= p.seq_len[seq];
    // pass 2: normalise, affine, pack
    for (int c = c0; c < c1; ++c) {
      tmem_ld32_sync(trow + c * 32, acc);
      if (p.bias) {
        smem_vec32(sp.bias + c * 32, aux);
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
        smem_vec32(sp.g + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] *= aux[i];
        smem_vec32(sp.b + c * 32, aux);
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
    sync();
    if (store_thread) {
      for (int sub = 0; sub < 2; ++sub)
        tma_store_3d(&tmO, staging + sub * kSubTileBytes, n_tile * 128 + sub * 64, t0, seq);
      tma_store_commit();
      tma_store_wait_read0();
    }
  } else if (p.mode != EPI_CONVERT) {
    fence_proxy_async_smem();
    sync();
    if (store_thread) {
      for (int sub = 0; sub < 4; ++sub)
        tma_store_3d(&tmO, staging + sub * kSubTileBytes, n0 + sub * 64, t0, seq);
      tma_store_commit();
      tma_store_wait_read0();
    }
    if (p.ln2_g != nullptr && (p.mode == EPI_LN || p.mode == EPI_RESID)) {
      // second output: LayerNorm of the (fp16) row just stored, with the next pre-norm sub-layer's affine
      sync();   // the TMA store above has finished reading the staging tile
      RowStats s2{0.f, 0.f, 0.f};
      for (int c = c0; c < c1; ++c) {
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
      {
        float dummy = 0.f;
        merge_stats(s2, dummy);
      }
      const float rstd2 = rsqrtf(s2.m2 * (1.f / 256.f) + p.ln_eps);
      for (int c = c0; c < c1; ++c) {
        staging_read32(staging, r, c, acc);
        smem_vec32(sp.g2 + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = (acc[i] - s2.mean) * rstd2 * aux[i];
        smem_vec32(sp.b2 + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
        staging_write32(staging, r, c, acc);
      }
      fence_proxy_async_smem();
      sync();
      if (store_thread) {
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
      for (int c = c0; c < c1; ++c) {
        tmem_ld32_sync(trow + c * 32, acc);
        load_vec32(p.pe_proj + s * 256 + c * 32, aux);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += aux[i];
        staging_write32(staging, r, c, acc);
      }
      fence_proxy_async_smem();
      sync();
      if (store_thread) {
        for (int sub = 0; sub < 4; ++sub) tma_store_3d(&tmO, staging + sub * kSubTileBytes, sub * 64, s, row0);
        tma_store_commit();
        tma_store_wait_read0();
      }
      sync();
    }
  }

}

}  // namespace gemm_detail
}  // namespace fseend
