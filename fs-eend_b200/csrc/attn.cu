// tcgen05 flash-attention forward, head_dim 64 (see attn.cuh).  Two masks share one kernel:
//   * ATTN_CAUSAL    : time-axis attention of the encoder / attractor decoder (key j visible iff j <= i + delay)
//   * ATTN_BLOCKDIAG : speaker-axis attention — the [frames*S] attractor rows are processed as 128-row tiles in
//                      which a row only sees the S rows of its own frame (block-diagonal mask, single KV tile).
//
// One CTA = one (sequence, head, 128-query tile).  160 threads:
//   warps 0-3 : softmax; thread r owns query row r (TMEM lane r).  One TMEM pass per KV tile (128 scores in
//               registers), exp2 in packed fp16 (ex2.approx.f16x2 — P is consumed as fp16 by the PV MMA anyway),
//               P written as fp16 into a 128B-swizzled smem tile (A operand of the PV MMA).
//               O accumulates in TMEM across KV tiles; when a row's running max grows, the warp rescales its
//               O rows in place (tcgen05.ld -> scale -> tcgen05.st), FlashAttention-4 style.
//   warp 4    : lane 0 issues TMA loads (Q once; K/V tiles through a 2-stage ring) and the tcgen05 MMAs
//               S = Q K^T (M128 N128 K64) and O += P V (M128 N64 K<=128, V consumed MN-major straight from its
//               row-major [kv][64] TMA tile).
// Causality skips KV tiles above the diagonal and trims the PV K-extent on the diagonal tile.
// Two CTAs per SM (112 KB smem, 256 TMEM columns each): one CTA's softmax overlaps the other's MMAs.
#include "attn.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

constexpr int kTile = 128;
constexpr int kQBytes = kTile * 64 * 2;  // 16 KB
constexpr int kKVBytes = kTile * 64 * 2;
constexpr int kPBytes = kTile * kTile * 2;  // 32 KB (two 64-column sub-tiles)
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQBytes;          // 2 stages
constexpr int kOffV = kOffK + 2 * kKVBytes;     // 2 stages
constexpr int kOffP = kOffV + 2 * kKVBytes;
constexpr int kOffBar = kOffP + kPBytes;        // mbarriers + tmem slot live at the tail of dynamic smem
constexpr int kSmemBytes = kOffBar + 128;
constexpr uint32_t kTmemCols = 256;             // S: [0,128), O: [128,192)

#ifndef FSEEND_ATTN_EXP_F32
#define FSEEND_ATTN_EXP_F32 0
#endif

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(160, 2)
attn_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* kv_empty = bars + 5;  // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* p_ready = bars + 8;
  uint64_t* pv_full = bars + 9;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  if (tid == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("[fseend] attn: dynamic smem base not 1024-aligned\n");
    __trap();
  }

  // ---- tile coordinates
  int q0, h, b, s, n_kv, kv_first, last_key = 0;
  h = blockIdx.y;
  if (p.mode == ATTN_CAUSAL) {
    const int n_qt = (p.T + kTile - 1) / kTile;
    const int qt = n_qt - 1 - static_cast<int>(blockIdx.x);  // heaviest (latest) query tiles first
    b = blockIdx.z / p.S;
    s = blockIdx.z % p.S;
    q0 = qt * kTile;
    last_key = min(q0 + kTile - 1 + p.mask_delay, p.T - 1);
    n_kv = last_key / kTile + 1;
    kv_first = 0;
  } else {
    b = 0;
    s = 0;
    q0 = blockIdx.x * p.tile_rows;   // T = total rows; tile_rows = (128 / S) * S
    n_kv = 1;
    kv_first = q0;
  }

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
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 4) tmem_alloc(tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 4) {
    if (lane == 0) {
      // ------------------------------------------------------------ control thread: TMA + MMA issue
      auto load_kv = [&](int j) {
        const int st = j & 1;
        const int row = kv_first + j * kTile;
        mbar_arrive_expect_tx(&k_full[st], kKVBytes);
        tma_load_4d(smem + kOffK + st * kKVBytes, &tmQKV, &k_full[st], 256 + h * 64, s, row, b);
        mbar_arrive_expect_tx(&v_full[st], kKVBytes);
        tma_load_4d(smem + kOffV + st * kKVBytes, &tmQKV, &v_full[st], 512 + h * 64, s, row, b);
      };
      constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, 64, true);
      const uint64_t qdesc = smem_desc_sw128(smem_u32(smem + kOffQ));
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1, 11);
        tc_fence_after();
        const uint64_t kdesc = smem_desc_sw128(smem_u32(smem + kOffK + st * kKVBytes));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_S, qdesc + 2 * kk, kdesc + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
        umma_commit(s_full);
      };

      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_4d(smem + kOffQ, &tmQKV, q_full, h * 64, s, q0, b);
      load_kv(0);
      if (n_kv > 1) load_kv(1);
      mbar_wait(q_full, 0, 10);
      issue_qk(0);

      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(p_ready, j & 1, 12);  // P(j) in smem, S(j) fully read, O rescaled
        mbar_wait(&v_full[st], (j >> 1) & 1, 13);
        tc_fence_after();
        const int valid_cols = (p.mode == ATTN_CAUSAL) ? min(kTile, last_key - j * kTile + 1) : kTile;
        const int n_k16 = (valid_cols + 15) >> 4;
        const uint32_t p_addr = smem_u32(smem + kOffP);
        const uint64_t vdesc = smem_desc_sw128(smem_u32(smem + kOffV + st * kKVBytes));
        for (int kk = 0; kk < n_k16; ++kk) {
          const uint64_t pdesc = smem_desc_sw128(p_addr + (kk >> 2) * (kTile * 128)) + 2 * (kk & 3);
          // V is MN-major: 16 kv rows = 16 * 128 B = 2048 B per K step -> +128 in 16-byte units
          umma_f16(tmem_O, pdesc, vdesc + 128 * kk, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(pv_full);
        umma_commit(&kv_empty[st]);
        if (j + 1 < n_kv) issue_qk(j + 1);
        if (j + 2 < n_kv) {
          mbar_wait(&kv_empty[st], (j >> 1) & 1, 14);
          load_kv(j + 2);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax warps
    const int r = tid;                 // query row inside the tile
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const float sl = p.scale * 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* ptile = smem + kOffP;

    for (int j = 0; j < n_kv; ++j) {
      // tile-relative visible column interval [lo, hi] of this row
      int lo, hi;
      bool full_tile;   // CTA-uniform: every row sees all 128 columns
      if (p.mode == ATTN_CAUSAL) {
        lo = 0;
        hi = min(q0 + r + p.mask_delay, p.T - 1) - j * kTile;
        full_tile = (j * kTile + kTile - 1) <= min(q0 + p.mask_delay, p.T - 1);
      } else {
        lo = (r / p.S) * p.S;
        hi = (r < p.tile_rows) ? lo + p.S - 1 : -1;
        if (q0 + hi >= p.T) hi = p.T - 1 - q0;
        full_tile = false;
      }

      mbar_wait(s_full, j & 1, 20);
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
        for (int i = 0; i < 128; ++i)
          sv[i] = (i >= lo && i <= hi) ? sv[i] : 0xff800000u;   // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
      const float m_new = fmaxf(m_run, mx);
      const float m_scaled = (m_new == -INFINITY) ? 0.f : m_new * sl;
      const float alpha = ex2(m_run * sl - m_scaled);   // m_run = -inf -> 0

      if (j > 0) {
        // PV(j-1) must have consumed P(j-1) before it is overwritten (and O must be complete before rescaling)
        mbar_wait(pv_full, (j - 1) & 1, 21);
        tc_fence_after();
      }
      float psum = 0.f;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        uint32_t e[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float x0 = fmaf(__uint_as_float(sv[q * 8 + 2 * t]), sl, -m_scaled);
          const float x1 = fmaf(__uint_as_float(sv[q * 8 + 2 * t + 1]), sl, -m_scaled);
#if FSEEND_ATTN_EXP_F32
          e[t] = pack_half2(ex2(x0), ex2(x1));
#else
          e[t] = ex2_h2(pack_half2(x0, x1));
#endif
        }
        const __half2 h01 = __hadd2(*reinterpret_cast<__half2*>(&e[0]), *reinterpret_cast<__half2*>(&e[1]));
        const __half2 h23 = __hadd2(*reinterpret_cast<__half2*>(&e[2]), *reinterpret_cast<__half2*>(&e[3]));
        const float2 f = __half22float2(__hadd2(h01, h23));
        psum += f.x + f.y;
        uint4 u = make_uint4(e[0], e[1], e[2], e[3]);
        *reinterpret_cast<uint4*>(ptile + (q >> 3) * (kTile * 128) + sw128_offset(r, q & 7)) = u;
      }
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        // some row of this warp raised its running max: rescale the warp's 32 O rows in TMEM
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          tmem_ld32(tmem_O + lane_base + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tmem_O + lane_base + c * 32, o);
        }
        tmem_st_wait();
      }
      l_run = l_run * alpha + psum;
      m_run = m_new;

      fence_proxy_async_smem();   // P visible to the tensor-core (async) proxy
      tc_fence_before();          // order our tcgen05.ld/st before the MMAs issued after the barrier
      mbar_arrive(p_ready);
    }
    // ---- epilogue: O / l -> fp16 -> staging (P buffer; every MMA that read it has completed) -> TMA store
    mbar_wait(pv_full, (n_kv - 1) & 1, 22);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_base + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_half2(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
        u.y = pack_half2(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
        u.z = pack_half2(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
        u.w = pack_half2(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(ptile + sw128_offset(r, c * 4 + q)) = u;
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (tid == 0) {
      tma_store_4d(&tmO, ptile, h * 64, s, q0, b);   // box rows = 128 (causal) or tile_rows (block-diagonal)
      tma_store_commit();
      tma_store_wait_read0();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

void launch_attn(const CUtensorMap& tmQKV, const CUtensorMap& tmO, const AttnParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    attr_set = true;
  }
  dim3 grid;
  if (p.mode == ATTN_CAUSAL) grid = dim3((p.T + kTile - 1) / kTile, p.H, p.B * p.S);
  else grid = dim3((p.T + p.tile_rows - 1) / p.tile_rows, p.H, 1);
  attn_kernel<<<grid, 160, kSmemBytes, stream>>>(tmQKV, tmO, p);
}

}  // namespace fseend
