// LS-EEND chunkwise retention on tcgen05.
//
// Reference: MultiScaleRetention.chunk_recurrent_forward (LS-EEND/nnet/modules/retention.py:146-194) with
// RetNetRelPos (decay = log 1, :20; no rotation, :209-213), group norm + swish gate (:222-224).  Algebraically, for
// row j of a chunk (index inside the chunk) and c earlier chunks:
//     O_j      = sum_{i <= j, same chunk} (q_j.k_i) v_i  +  sqrt(C) * q_j R'_c         R'_c = (sum_{earlier} k^T v)/sqrt(C)
//     inner_j  = max(1, sum_{i <= j} |q_j.k_i| / sqrt(j+1)),   cross_c = max(1, max_d sum_e |R'_c[e][d]|)
//     ret_j    = O_j / (sqrt(j+1) * max(inner_j, cross_c))
//     out_j    = swish(g_j) * LayerNorm_64(ret_j; eps 1e-6, no affine)
// (k arrives pre-scaled by hd^-0.5, folded into the projection weights).
//
// Kernel structure = attn.cu without the softmax: one CTA per (sequence, chunk, head, 128-row tile); warp 4 lane 0
// drives TMA and the MMAs S = Q K^T, O1 += P V (P = masked raw scores as fp16) and O2 = Q R'; warps 0-3 own one row
// each: mask, |.|-sum, fp16 pack, and the final scale / group norm / gate.
#include "once.h"
#include "retention.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

constexpr int kTile = 128;
constexpr int kQBytes = kTile * 64 * 2;
constexpr int kKVBytes = kTile * 64 * 2;
constexpr int kPBytes = kTile * kTile * 2;
constexpr int kOffQ = 0;                        // Q, later the gate tile G
constexpr int kOffK = kOffQ + kQBytes;
constexpr int kOffV = kOffK + 2 * kKVBytes;
constexpr int kOffP = kOffV + 2 * kKVBytes;     // first 8 KB hold R' until the first P is written
constexpr int kOffBar = kOffP + kPBytes;
constexpr int kSmemBytes = kOffBar + 128;
constexpr uint32_t kTmemCols = 256;             // S [0,128), O1 [128,192), O2 [192,256)

__global__ void __launch_bounds__(160, 2)
retention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmO, const RetParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* kv_empty = bars + 5;  // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* p_ready = bars + 8;
  uint64_t* pv_full = bars + 9;
  uint64_t* g_full = bars + 10;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("[fseend] retention: dynamic smem base not 1024-aligned\n");
    __trap();
  }

  const int n_qt = (p.chunk + kTile - 1) / kTile;
  const int qt = n_qt - 1 - static_cast<int>(blockIdx.x);   // heaviest tiles first
  const int h = blockIdx.y;
  const int c = blockIdx.z % p.n_chunks;
  const int n = blockIdx.z / p.n_chunks;
  const int b = n / p.S, s = n % p.S;
  const int q0 = qt * kTile;          // row inside the chunk
  const int n_kv = qt + 1;
  const bool has_cross = c > 0;

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(pv_full, 1);
    mbar_init(g_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 4) tmem_alloc(tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t tmem_S = tmem_base, tmem_O1 = tmem_base + 128, tmem_O2 = tmem_base + 192;

  // The control warp stays converged and elects one lane around the TMA / tcgen05 instructions (uniform-register operands).
  if (warp == 4) {
    {
      auto load_kv = [&](int j) {
        const int st = j & 1;
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[st], kKVBytes);
          tma_load_5d(smem + kOffK + st * kKVBytes, &tmQ, &k_full[st], 256 + h * 64, s, j * kTile, c, b);
          mbar_arrive_expect_tx(&v_full[st], kKVBytes);
          tma_load_5d(smem + kOffV + st * kKVBytes, &tmQ, &v_full[st], 512 + h * 64, s, j * kTile, c, b);
        }
        __syncwarp();
      };
      constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, 64, true);
      const uint64_t qdesc = smem_desc_sw128(smem_u32(smem + kOffQ));
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1, 51);
        tc_fence_after();
        const uint64_t kdesc = smem_desc_sw128(smem_u32(smem + kOffK + st * kKVBytes));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_S, qdesc + 2 * kk, kdesc + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
          umma_commit(s_full);
        }
        __syncwarp();
      };

      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, kQBytes + (has_cross ? 64 * 64 * 2 : 0));
        tma_load_5d(smem + kOffQ, &tmQ, q_full, h * 64, s, q0, c, b);
        if (has_cross) tma_load_3d(smem + kOffP, &tmR, q_full, 0, 0, (n * p.H + h) * p.n_chunks + c);
      }
      __syncwarp();
      load_kv(0);
      if (n_kv > 1) load_kv(1);
      mbar_wait(q_full, 0, 50);
      tc_fence_after();
      if (has_cross) {
        // O2 = Q R'  (R' is [64 e][64 d] row-major = MN-major B operand, K = e)
        const uint64_t rdesc = smem_desc_sw128(smem_u32(smem + kOffP));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_O2, qdesc + 2 * kk, rdesc + 128 * kk, idesc_pv, kk > 0 ? 1u : 0u);
        }
        __syncwarp();
      }
      issue_qk(0);   // its commit (s_full) also covers the cross MMA: R' may be overwritten by P afterwards

      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(p_ready, j & 1, 52);
        if (j == n_kv - 1) {
          // every MMA reading Q has completed (the rows just consumed S(j)): reuse its buffer for the gate tile
          if (elect_one()) {
            mbar_arrive_expect_tx(g_full, kQBytes);
            tma_load_5d(smem + kOffQ, &tmQ, g_full, 768 + h * 64, s, q0, c, b);
          }
          __syncwarp();
        }
        mbar_wait(&v_full[st], (j >> 1) & 1, 53);
        tc_fence_after();
        const int valid_cols = min(kTile, q0 + kTile - j * kTile);   // diagonal tile: all 128; earlier tiles: 128
        const int n_k16 = (min(valid_cols, kTile) + 15) >> 4;
        const uint32_t p_addr = smem_u32(smem + kOffP);
        const uint64_t vdesc = smem_desc_sw128(smem_u32(smem + kOffV + st * kKVBytes));
        if (elect_one()) {
          for (int kk = 0; kk < n_k16; ++kk) {
            const uint64_t pdesc = smem_desc_sw128(p_addr + (kk >> 2) * (kTile * 128)) + 2 * (kk & 3);
            umma_f16(tmem_O1, pdesc, vdesc + 128 * kk, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(pv_full);
          umma_commit(&kv_empty[st]);
        }
        __syncwarp();
        if (j + 1 < n_kv) issue_qk(j + 1);
        if (j + 2 < n_kv) {
          mbar_wait(&kv_empty[st], (j >> 1) & 1, 54);
          load_kv(j + 2);
        }
      }
    }
    __syncwarp();
  } else {
    const int r = tid;
    const int jl = q0 + r;                                   // row index inside the chunk
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    uint8_t* ptile = smem + kOffP;
    float abs_sum = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int hi = jl - j * kTile;                         // visible tile columns: [0, hi]
      const bool full_tile = (j < qt);                       // strictly below the diagonal tile
      mbar_wait(s_full, j & 1, 60);
      tc_fence_after();
      uint32_t sv[128];
      {
        uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
        uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
        uint32_t(&a2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[64]);
        uint32_t(&a3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[96]);
        tmem_ld32(tmem_S + lane_base + 0, a0);
        tmem_ld32(tmem_S + lane_base + 32, a1);
        tmem_ld32(tmem_S + lane_base + 64, a2);
        tmem_ld32(tmem_S + lane_base + 96, a3);
        tmem_ld_wait();
      }
      if (!full_tile) {
#pragma unroll
        for (int i = 0; i < 128; ++i) sv[i] = (i <= hi) ? sv[i] : 0u;
      }
      if (j > 0) {
        mbar_wait(pv_full, (j - 1) & 1, 61);                 // PV(j-1) has consumed P(j-1)
        tc_fence_after();
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        uint4 u;
        uint32_t* e = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float x0 = __uint_as_float(sv[q * 8 + 2 * t]), x1 = __uint_as_float(sv[q * 8 + 2 * t + 1]);
          abs_sum += fabsf(x0) + fabsf(x1);
          e[t] = pack_half2(x0, x1);
        }
        *reinterpret_cast<uint4*>(ptile + (q >> 3) * (kTile * 128) + sw128_offset(r, q & 7)) = u;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // ---- epilogue
    mbar_wait(pv_full, (n_kv - 1) & 1, 62);
    mbar_wait(g_full, 0, 63);
    tc_fence_after();
    float o[64];
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t t1[32];
      tmem_ld32(tmem_O1 + lane_base + cc * 32, t1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[cc * 32 + i] = __uint_as_float(t1[i]);
    }
    float cross = 1.f;
    if (has_cross) {
      const float sqrt_c = sqrtf(static_cast<float>(p.chunk));
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t t2[32];
        tmem_ld32(tmem_O2 + lane_base + cc * 32, t2);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[cc * 32 + i] = fmaf(sqrt_c, __uint_as_float(t2[i]), o[cc * 32 + i]);
      }
      cross = p.cross_scale[(static_cast<size_t>(n) * p.H + h) * p.n_chunks + c];
    }
    const float rs = rsqrtf(static_cast<float>(jl + 1));
    const float inner = fmaxf(1.f, abs_sum * rs);
    const float scale = rs / fmaxf(inner, cross);
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      o[i] *= scale;
      mean += o[i];
    }
    mean *= (1.f / 64.f);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const float d = o[i] - mean;
      var = fmaf(d, d, var);
    }
    const float rstd = rsqrtf(var * (1.f / 64.f) + 1e-6f);
    const uint8_t* gtile = smem + kOffQ;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 gu = *reinterpret_cast<const uint4*>(gtile + sw128_offset(r, q));
      const __half2* gh = reinterpret_cast<const __half2*>(&gu);
      uint4 u;
      uint32_t* e = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 g = __half22float2(gh[t]);
        const float y0 = (o[q * 8 + 2 * t] - mean) * rstd * (g.x / (1.f + __expf(-g.x)));
        const float y1 = (o[q * 8 + 2 * t + 1] - mean) * rstd * (g.y / (1.f + __expf(-g.y)));
        e[t] = pack_half2(y0, y1);
      }
      *reinterpret_cast<uint4*>(ptile + sw128_offset(r, q)) = u;
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (tid == 0) {
      tma_store_5d(&tmO, ptile, h * 64, s, q0, c, b);     // rows beyond the chunk end are clipped
      tma_store_commit();
      tma_store_wait_read0();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// Chunk states on CUDA cores: one block per (sequence, head) walks its chunks in order, keeping the 64x64 running sum
// of k^T v (fp32) in shared memory; emits R'_c and cross_scale_c BEFORE adding chunk c.
// Register tiling: thread = (row subset rsub of 4, 8x8 block of the 64x64 state): per key/value row it reads 8 k and 8 v
// values (four 128-bit shared loads) for 64 FMAs, so the kernel is FMA-bound instead of load-bound (the first version
// did 17 shared loads per 16 FMAs: 1.0 ms per decoder launch at B=16, T=2000, S=10).  Rows are fetched with 16-byte
// loads (8 fp16) and converted once.
__global__ void __launch_bounds__(256)
ret_chunk_state_kernel(const __half* __restrict__ qkvg, RetParams p, __half* __restrict__ state,
                       float* __restrict__ cross_scale) {
  constexpr int R = 32;                         // rows per batch
  const int n = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int b = n / p.S, s = n % p.S;
  __shared__ __align__(16) float ks[R][64];
  __shared__ __align__(16) float vs[R][64];
  __shared__ __align__(16) float st[64][64];   // running sum over the chunks already added: st[e][d]
  __shared__ float red[2];
  const int rsub = tid >> 6, l64 = tid & 63;
  const int eb = (l64 >> 3) * 8, db = (l64 & 7) * 8;
  for (int i = tid; i < 4096; i += 256) (&st[0][0])[i] = 0.f;
  __syncthreads();
  const float inv_sqrt_c = rsqrtf(static_cast<float>(p.chunk));
  // loader role: thread -> (row lr of the batch, k or v, 16-byte piece)
  const int lr = tid >> 3, lpiece = tid & 7;
  for (int c = 0; c < p.n_chunks; ++c) {
    // ---- emit the state seen by chunk c
    __half* out = state + ((static_cast<size_t>(n) * p.H + h) * p.n_chunks + c) * 4096;
    for (int i = tid; i < 2048; i += 256) {     // two adjacent d per thread
      const int e = i >> 5, d2 = (i & 31) * 2;
      reinterpret_cast<__half2*>(out + e * 64 + d2)[0] =
          __floats2half2_rn(st[e][d2] * inv_sqrt_c, st[e][d2 + 1] * inv_sqrt_c);
    }
    if (tid < 64) {
      float cs = 0.f;
      for (int e = 0; e < 64; ++e) cs += fabsf(st[e][tid] * inv_sqrt_c);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) cs = fmaxf(cs, __shfl_xor_sync(0xffffffffu, cs, off));
      if ((tid & 31) == 0) red[tid >> 5] = cs;
    }
    __syncthreads();
    if (tid == 0)
      cross_scale[(static_cast<size_t>(n) * p.H + h) * p.n_chunks + c] = fmaxf(1.f, fmaxf(red[0], red[1]));
    if (c == p.n_chunks - 1) break;
    // ---- accumulate chunk c:  acc[e][d] += sum_i k_i[e] v_i[d] over this thread's rows
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    // software pipeline: the global loads of batch i0 + R are in flight while batch i0 is multiplied
    auto fetch = [&](int i0, uint4& kk, uint4& vv) {
      kk = make_uint4(0, 0, 0, 0);
      vv = make_uint4(0, 0, 0, 0);
      if (i0 + lr < p.chunk) {
        const int t = c * p.chunk + i0 + lr;
        const __half* row = qkvg + ((static_cast<size_t>(b) * p.T + t) * p.S + s) * 1024 + h * 64 + lpiece * 8;
        kk = *reinterpret_cast<const uint4*>(row + 256);
        vv = *reinterpret_cast<const uint4*>(row + 512);
      }
    };
    uint4 kk, vv;
    fetch(0, kk, vv);
    for (int i0 = 0; i0 < p.chunk; i0 += R) {
      {
        const __half2* kh = reinterpret_cast<const __half2*>(&kk);
        const __half2* vh = reinterpret_cast<const __half2*>(&vv);
        float4 k0, k1, v0, v1;
        float2 f;
        f = __half22float2(kh[0]); k0.x = f.x; k0.y = f.y;
        f = __half22float2(kh[1]); k0.z = f.x; k0.w = f.y;
        f = __half22float2(kh[2]); k1.x = f.x; k1.y = f.y;
        f = __half22float2(kh[3]); k1.z = f.x; k1.w = f.y;
        f = __half22float2(vh[0]); v0.x = f.x; v0.y = f.y;
        f = __half22float2(vh[1]); v0.z = f.x; v0.w = f.y;
        f = __half22float2(vh[2]); v1.x = f.x; v1.y = f.y;
        f = __half22float2(vh[3]); v1.z = f.x; v1.w = f.y;
        reinterpret_cast<float4*>(&ks[lr][lpiece * 8])[0] = k0;
        reinterpret_cast<float4*>(&ks[lr][lpiece * 8])[1] = k1;
        reinterpret_cast<float4*>(&vs[lr][lpiece * 8])[0] = v0;
        reinterpret_cast<float4*>(&vs[lr][lpiece * 8])[1] = v1;
      }
      if (i0 + R < p.chunk) fetch(i0 + R, kk, vv);
      __syncthreads();
#pragma unroll
      for (int q = 0; q < R / 4; ++q) {
        const int rr = q * 4 + rsub;
        const float4 ka = reinterpret_cast<const float4*>(&ks[rr][eb])[0];
        const float4 kb = reinterpret_cast<const float4*>(&ks[rr][eb])[1];
        const float4 va = reinterpret_cast<const float4*>(&vs[rr][db])[0];
        const float4 vb = reinterpret_cast<const float4*>(&vs[rr][db])[1];
        const float kv[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
        const float vv8[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(kv[i], vv8[j], acc[i][j]);
      }
      __syncthreads();
    }
    // ---- fold the four row subsets into the running state (fixed order: deterministic)
    for (int rs = 0; rs < 4; ++rs) {
      if (rsub == rs) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4* dst = reinterpret_cast<float4*>(&st[eb + i][db]);
          float4 x0 = dst[0], x1 = dst[1];
          x0.x += acc[i][0]; x0.y += acc[i][1]; x0.z += acc[i][2]; x0.w += acc[i][3];
          x1.x += acc[i][4]; x1.y += acc[i][5]; x1.z += acc[i][6]; x1.w += acc[i][7];
          dst[0] = x0;
          dst[1] = x1;
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace

void launch_ret_chunk_state(const __half* qkvg, const RetParams& p, __half* state, float* cross_scale,
                            cudaStream_t stream) {
  ret_chunk_state_kernel<<<dim3(p.B * p.S, p.H), 256, 0, stream>>>(qkvg, p, state, cross_scale);
}

void launch_retention(const CUtensorMap& tmQKVG, const CUtensorMap& tmState, const CUtensorMap& tmO, const RetParams& p,
                      cudaStream_t stream) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(retention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  }
  dim3 grid((p.chunk + kTile - 1) / kTile, p.H, p.B * p.S * p.n_chunks);
  retention_kernel<<<grid, 160, kSmemBytes, stream>>>(tmQKVG, tmState, tmO, p);
}

}  // namespace fseend
