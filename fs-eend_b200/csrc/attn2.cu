// Causal tcgen05 flash-attention, head_dim 64, with TWO query tiles per CTA (see attn.cuh; attn.cu keeps the
// one-tile kernel, which also serves the block-diagonal mode).
//
// Why: profiles/r01_attention_study.md — with T = 500 the one-tile kernel spends its time in the per-KV-tile handshake
// chain (MMA commit -> mbarrier -> TMEM read -> softmax -> arrive -> MMA issue), not in any execution unit.  Here every
// handshake carries four times the work and two independent chains run in one CTA:
//   * KV tiles are 128 keys wide (S tile 128 x 128), so a sequence of 500 frames is 1-4 steps per query tile;
//   * a work item is a PAIR of adjacent query tiles (A = the later one, B) of one (sequence, head): both read the same
//     K/V tiles from shared memory (half the L2 -> SM traffic), each has its own S / O accumulators in TMEM and its own
//     softmax warpgroup, and the tensor pipe alternates between them (ping-pong: QK^T / PV of one tile run while the
//     other tile's softmax does);
//   * P (fp16) is written back into the TMEM columns of the S tile it was computed from and consumed from there as
//     the A operand of the PV MMA (no P buffers in shared memory, no generic->async proxy fence); S is single-buffered,
//     the tensor pipe's in-order execution protects the S/P columns (QK^T(j+1) is issued behind PV(j));
//   * lazy rescaling (reference maximum moves only when outgrown by 2^8), packed-fp16 exp2, rows processed in 32-column
//     chunks (32 live S values per thread).
//
// PERSISTENT, one CTA per SM, 320 threads: warps 0-3 softmax of tile A, warps 4-7 softmax of tile B (thread r <-> query
// row r, TMEM lane r), warp 8 TMA producer, warp 9 tcgen05 issuer.
// TMEM (512 columns): S_A [0,128), S_B [128,256), O_A [256,320), O_B [320,384).
#include "once.h"
#include <stdlib.h>

#include "attn.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

constexpr int kTile = 128;                 // query rows per tile
constexpr int kKV = 64;                    // keys per KV tile
constexpr int kStages = 3;
constexpr int kQBytes = kTile * 64 * 2;    // 16 KB
constexpr int kKBytes = kKV * 64 * 2;      // 16 KB
constexpr int kStageBytes = 2 * kKBytes;   // K then V
constexpr int kOffQ = 0;                                    // [2 item buffers][A, B]
constexpr int kOffKV = kOffQ + 4 * kQBytes;                 // 64 KB
constexpr int kOffBar = kOffKV + kStages * kStageBytes;     // 160 KB
constexpr int kSmemBytes = kOffBar + 256;
constexpr uint32_t kTmemCols = 256;        // S_A [0,64), S_B [64,128), O_A [128,192), O_B [192,256)
constexpr int kThreads = 320;
constexpr float kTau = 8.f;                // lazy-rescale threshold (log2 units): P <= 2^8

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) {   // two exponentials per MUFU op
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ __half2 as_half2(uint32_t x) { return *reinterpret_cast<__half2*>(&x); }
// D[tmem] (+)= A[tmem] * B[smem desc]: A is M x K fp16 packed two per 32-bit TMEM column (lane = row)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {   // one full 32-byte sector
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z),
               "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// One work item: a pair of query tiles of one (sequence, head).  Tile A = q-tile index qa, tile B = qa - 1 (absent when
// qa == 0).  n_a / n_b: KV tiles each needs; last_a / last_b: last visible key of the tile's last row.
struct Item {
  int h, b, s, qa, n_a, n_b, last_a, last_b;
};
__device__ __forceinline__ Item decode_item(const AttnParams& p, int id) {
  // heavy pairs first (the latest query tiles of every (sequence, head)), lighter pairs after them
  const int n_qt = (p.T + kTile - 1) / kTile;
  const int n_sh = p.H * p.B * p.S;
  const int pair = id / n_sh;                 // 0 = (n_qt-1, n_qt-2), 1 = (n_qt-3, n_qt-4), ...
  const int rest = id - pair * n_sh;
  Item it;
  it.h = rest % p.H;
  const int z = rest / p.H;
  it.b = z / p.S;
  it.s = z % p.S;
  it.qa = n_qt - 1 - 2 * pair;
  it.last_a = min(it.qa * kTile + kTile - 1 + p.mask_delay, p.T - 1);
  it.n_a = it.last_a / kKV + 1;
  if (it.qa >= 1) {
    it.last_b = min((it.qa - 1) * kTile + kTile - 1 + p.mask_delay, p.T - 1);
    it.n_b = it.last_b / kKV + 1;
  } else {
    it.last_b = -1;
    it.n_b = 0;
  }
  return it;
}

__global__ void __launch_bounds__(kThreads, 2)
attn2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
             __half* __restrict__ out, const AttnParams p, const int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;        // [2]  both Q tiles of an item buffer have landed
  uint64_t* q_empty = bars + 2;       // [2]  every QK^T of the item in this buffer has completed
  uint64_t* kv_full = bars + 4;       // [kStages]
  uint64_t* kv_empty = bars + 7;      // [kStages]
  uint64_t* s_full = bars + 10;       // [2]  per query tile (A, B)
  uint64_t* p_ready = bars + 12;      // [2]
  uint64_t* o_full = bars + 14;       // [2]
  uint64_t* o_free = bars + 16;       // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 18);
  static_assert(kStages == 3, "barrier layout");

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("[fseend] attn2: dynamic smem base not 1024-aligned\n");
    __trap();
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 128);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 9) tmem_alloc(tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer (runs ahead across items)
    uint32_t g = 0, n = 0;
    for (int id = blockIdx.x; id < n_items; id += gridDim.x, ++n) {
      const Item it = decode_item(p, id);
      const int qb = n & 1;
      mbar_wait(&q_empty[qb], ((n >> 1) & 1) ^ 1, 10);
      if (elect_one()) {
        uint8_t* qa = smem + kOffQ + (2 * qb) * kQBytes;
        mbar_arrive_expect_tx(&q_full[qb], it.n_b > 0 ? 2 * kQBytes : kQBytes);
        tma_load_4d(qa, &tmQ, &q_full[qb], it.h * 64, it.s, it.qa * kTile, it.b);
        if (it.n_b > 0) tma_load_4d(qa + kQBytes, &tmQ, &q_full[qb], it.h * 64, it.s, (it.qa - 1) * kTile, it.b);
      }
      __syncwarp();
      for (int j = 0; j < it.n_a; ++j, ++g) {
        const int st = g % kStages;
        mbar_wait(&kv_empty[st], ((g / kStages) & 1) ^ 1, 11);
        uint8_t* dst = smem + kOffKV + st * kStageBytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[st], kStageBytes);
          tma_load_4d(dst, &tmKV, &kv_full[st], 256 + it.h * 64, it.s, j * kKV, it.b);
          tma_load_4d(dst + kKBytes, &tmKV, &kv_full[st], 512 + it.h * 64, it.s, j * kKV, it.b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer (warp-converged)
    constexpr uint32_t idesc_qk = make_idesc_f16(128, kKV, false);
    constexpr uint32_t idesc_pv = make_idesc_f16(128, 64, true);
    uint32_t g = 0, n = 0;
    uint32_t ta = 0, tb = 0;       // KV tiles processed so far by tile A / tile B (phase counters of s_full, p_ready)
    uint32_t ia = 0, ib = 0;       // items processed so far by A / B (phase counters of o_full, o_free)
    for (int id = blockIdx.x; id < n_items; id += gridDim.x, ++n) {
      const Item it = decode_item(p, id);
      const int qb = n & 1;
      const uint64_t qdesc_a = smem_desc_sw128(smem_u32(smem + kOffQ + (2 * qb) * kQBytes));
      const uint64_t qdesc_b = smem_desc_sw128(smem_u32(smem + kOffQ + (2 * qb + 1) * kQBytes));
      auto issue_qk = [&](int x, int j) {       // S_x = Q_x K(j)^T   (x = 0: A, 1: B)
        const int st = (g + j) % kStages;
        const uint64_t kdesc = smem_desc_sw128(smem_u32(smem + kOffKV + st * kStageBytes));
        const uint64_t qdesc = x == 0 ? qdesc_a : qdesc_b;
        const uint32_t tmem_S = tmem_base + x * kKV;
        const bool last = j == (x == 0 ? it.n_a : it.n_b) - 1;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_S, qdesc + 2 * kk, kdesc + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
          umma_commit(&s_full[x]);
          // the commit behind the item's last QK^T in program order (tile B's when both tiles need the same number of
          // KV tiles, else tile A's) frees both Q tiles
          if (last && (x == 0 ? it.n_b < it.n_a : it.n_b == it.n_a)) umma_commit(&q_empty[qb]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int x, int j) {       // O_x (+)= P_x V(j)
        const int st = (g + j) % kStages;
        const int last_key = x == 0 ? it.last_a : it.last_b;
        const int n_x = x == 0 ? it.n_a : it.n_b;
        const int valid_cols = min(kKV, last_key - j * kKV + 1);
        const int n_k16 = (valid_cols + 15) >> 4;
        const uint64_t vdesc = smem_desc_sw128(smem_u32(smem + kOffKV + st * kStageBytes + kKBytes));
        const uint32_t tmem_P = tmem_base + x * kKV, tmem_O = tmem_base + 2 * kKV + x * 64;
        if (elect_one()) {
          for (int kk = 0; kk < n_k16; ++kk)   // A: 16 keys = 8 packed TMEM columns; V MN-major: 16 rows = +128 units
            umma_f16_ts(tmem_O, tmem_P + 8 * kk, vdesc + 128 * kk, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
          if (j == n_x - 1) umma_commit(&o_full[x]);
        }
        __syncwarp();
      };
      mbar_wait(&q_full[qb], (n >> 1) & 1, 13);
      mbar_wait(&kv_full[g % kStages], (g / kStages) & 1, 12);
      tc_fence_after();
      issue_qk(0, 0);
      if (it.n_b > 0) issue_qk(1, 0);
      for (int j = 0; j < it.n_a; ++j) {
        const int st = (g + j) % kStages;
        // ---- tile A
        mbar_wait(&p_ready[0], ta & 1, 14);
        if (j == 0) mbar_wait(&o_free[0], (ia & 1) ^ 1, 15);
        tc_fence_after();
        issue_pv(0, j);
        ++ta;
        if (j + 1 < it.n_a) {
          mbar_wait(&kv_full[(g + j + 1) % kStages], ((g + j + 1) / kStages) & 1, 12);
          tc_fence_after();
          issue_qk(0, j + 1);     // behind PV_A(j) in the in-order tensor pipe: S_A / P_A columns are free
        }
        // ---- tile B
        if (j < it.n_b) {
          mbar_wait(&p_ready[1], tb & 1, 16);
          if (j == 0) mbar_wait(&o_free[1], (ib & 1) ^ 1, 17);
          tc_fence_after();
          issue_pv(1, j);
          ++tb;
          if (j + 1 < it.n_b) issue_qk(1, j + 1);
        }
        if (elect_one()) umma_commit(&kv_empty[st]);    // every MMA reading K/V tile j has been issued
        __syncwarp();
      }
      g += it.n_a;
      ++ia;
      if (it.n_b > 0) ++ib;
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups: x = 0 (tile A), 1 (tile B)
    const int x = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + (tid & 31);        // query row inside the tile
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tmem_S = tmem_base + x * kKV + lane_base;       // P (packed fp16) aliases the first kKV / 2 columns
    const uint32_t tmem_O = tmem_base + 2 * kKV + x * 64 + lane_base;
    const float sl = p.scale * 1.4426950408889634f;
    const size_t out_stride = static_cast<size_t>(p.S) * 256;      // between consecutive frames of one (b, s)
    uint32_t tx = 0, ix = 0;       // KV tiles / items processed by this warpgroup
    for (int id = blockIdx.x; id < n_items; id += gridDim.x) {
      const Item it = decode_item(p, id);
      const int n_x = x == 0 ? it.n_a : it.n_b;
      if (n_x == 0) continue;       // single-tile item: warpgroup B has nothing to do
      const int q0 = (it.qa - x) * kTile;
      const int hi_row = min(q0 + r + p.mask_delay, p.T - 1);                      // last visible key of this row
      const int hi_w1 = min(q0 + quarter * 32 + 31 + p.mask_delay, p.T - 1);       // ... of the warp's last row
      const int hi_w0 = min(q0 + quarter * 32 + p.mask_delay, p.T - 1);            // ... of the warp's first row
      float m_used = -INFINITY, l_run = 0.f;   // lazy reference maximum (scaled log2 units), row sum

      for (int j = 0; j < n_x; ++j, ++tx) {
        const int hi = hi_row - j * kKV, whi = hi_w1 - j * kKV, wlo = hi_w0 - j * kKV;
        mbar_wait(&s_full[x], tx & 1, 20);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < kKV / 32; ++cc) {
          // warp-uniform mode of this 32-column chunk: 0 = every row of the warp sees all of it, 1 = mixed (per-element
          // compare), 2 = no row sees any of it (no TMEM read, no exponentials, P = 0)
          const int cmode = (cc * 32 + 31 <= wlo) ? 0 : ((cc * 32 > whi) ? 2 : 1);
          uint32_t pk[16];
          if (cmode == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          } else {
            uint32_t sv[32];
            tmem_ld32(tmem_S + cc * 32, sv);
            tmem_ld_wait();
            if (cmode == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) sv[i] = (cc * 32 + i <= hi) ? sv[i] : 0xff800000u;   // -inf
            }
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent FMNMX3 chains
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              mx4[0] = fmax3(mx4[0], __uint_as_float(sv[i + 0]), __uint_as_float(sv[i + 1]));
              mx4[1] = fmax3(mx4[1], __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
              mx4[2] = fmax3(mx4[2], __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
              mx4[3] = fmax3(mx4[3], __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
            }
            const float m_chunk = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sl;   // -inf stays -inf
            float alpha = 1.f;
            if (m_chunk > m_used + kTau) {         // also the first visible columns of the row (m_used = -inf)
              alpha = ex2(m_used - m_chunk);       // m_used = -inf -> 0
              m_used = m_chunk;
            }
            const float m_sub = (m_used == -INFINITY) ? 0.f : m_used;
            // P = 2^(s*sl - m_used) as packed fp16 pairs; the row sum is taken from the rounded values the PV MMA
            // consumes: groups of 8 in fp16, then fp32
            float psum = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int c = q * 8 + 2 * t;
                pk[q * 4 + t] = ex2_h2(pack_half2(fmaf(__uint_as_float(sv[c]), sl, -m_sub),
                                                  fmaf(__uint_as_float(sv[c + 1]), sl, -m_sub)));
              }
              const float2 f = __half22float2(__hadd2(__hadd2(as_half2(pk[q * 4]), as_half2(pk[q * 4 + 1])),
                                                      __hadd2(as_half2(pk[q * 4 + 2]), as_half2(pk[q * 4 + 3]))));
              psum += f.x + f.y;
            }
            l_run = l_run * alpha + psum;
            // rare: some row of this warp moved its reference maximum -> rescale what was scaled with the old one:
            // the P chunks of this tile already written, and O (quiescent: PV(j-1) precedes QK^T(j) in the tensor
            // pipe and S(j) has been observed complete)
            if ((j > 0 || cc > 0) && __any_sync(0xffffffffu, alpha != 1.f)) {
              const __half2 a2 = __float2half2_rn(alpha);
              for (int pc = 0; pc < cc; ++pc) {
                uint32_t p0[16];
                tmem_ld16(tmem_S + pc * 16, p0);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const __half2 v = __hmul2(as_half2(p0[i]), a2);
                  p0[i] = *reinterpret_cast<const uint32_t*>(&v);
                }
                tmem_st16(tmem_S + pc * 16, p0);
              }
              if (j > 0) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  uint32_t o[32];
                  tmem_ld32(tmem_O + c * 32, o);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                  tmem_st32(tmem_O + c * 32, o);
                }
              }
            }
          }
          // written over S columns this thread has already consumed ([16 cc, 16 cc + 16) < 32 cc + 32)
          tmem_st16(tmem_S + cc * 16, pk);
        }
        tmem_st_wait();
        tc_fence_before();          // order our tcgen05.ld/st before the MMAs issued after the barrier
        mbar_arrive(&p_ready[x]);
      }

      // ---- tile epilogue: O / l -> fp16 -> global; each thread owns one 128-byte row segment (64 fp16 of head h),
      // written as four 32-byte (full sector) stores
      mbar_wait(&o_full[x], ix & 1, 23);
      ++ix;
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      const int rows = min(kTile, p.T - q0);
      __half* dst = out + ((static_cast<size_t>(it.b) * p.T + q0) * p.S + it.s) * 256 + it.h * 64 +
                    static_cast<size_t>(r) * out_stride;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(tmem_O + c * 32, o);
        tmem_ld_wait();
        if (c == 1) {
          tc_fence_before();
          mbar_arrive(&o_free[x]);      // O has been read out: the next item's first PV may overwrite it
        }
        if (r < rows) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint4 u, v;
            u.x = pack_half2(__uint_as_float(o[q * 16 + 0]) * inv, __uint_as_float(o[q * 16 + 1]) * inv);
            u.y = pack_half2(__uint_as_float(o[q * 16 + 2]) * inv, __uint_as_float(o[q * 16 + 3]) * inv);
            u.z = pack_half2(__uint_as_float(o[q * 16 + 4]) * inv, __uint_as_float(o[q * 16 + 5]) * inv);
            u.w = pack_half2(__uint_as_float(o[q * 16 + 6]) * inv, __uint_as_float(o[q * 16 + 7]) * inv);
            v.x = pack_half2(__uint_as_float(o[q * 16 + 8]) * inv, __uint_as_float(o[q * 16 + 9]) * inv);
            v.y = pack_half2(__uint_as_float(o[q * 16 + 10]) * inv, __uint_as_float(o[q * 16 + 11]) * inv);
            v.z = pack_half2(__uint_as_float(o[q * 16 + 12]) * inv, __uint_as_float(o[q * 16 + 13]) * inv);
            v.w = pack_half2(__uint_as_float(o[q * 16 + 14]) * inv, __uint_as_float(o[q * 16 + 15]) * inv);
            st_global_256(dst + c * 32 + q * 16, u, v);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

// tmQ: box (64,1,128,1) over (768, S, T, B); tmKV: the same tensor with box (64,1,kKV,1).
void launch_attn2(const CUtensorMap& tmQ, const CUtensorMap& tmKV, __half* out, const AttnParams& p,
                  cudaStream_t stream) {
  static int num_sms = 0;
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(attn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_qt = (p.T + kTile - 1) / kTile;
  const int n_items = ((n_qt + 1) / 2) * p.H * p.B * p.S;
  const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
  attn2_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tmQ, tmKV, out, p, n_items);
}

}  // namespace fseend
