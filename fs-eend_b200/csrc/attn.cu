// tcgen05 flash-attention forward, head_dim 64 (see attn.cuh).  Two masks share one kernel:
//   * ATTN_CAUSAL    : time-axis attention of the encoder / attractor decoder (key j visible iff j <= i + delay)
//   * ATTN_BLOCKDIAG : speaker-axis attention — the [frames*S] attractor rows are processed as 128-row tiles in
//                      which a row only sees the S rows of its own frame (block-diagonal mask).
//
// PERSISTENT kernel: 2 CTAs per SM, each looping over work items (sequence, head, 128-query tile), heaviest first.
// Profiling the one-item-per-CTA versions showed the same L2->SM sector rate (85 sectors/ns) for both masks and
// ~30 KB in flight per SM: load -> compute -> store ran serially per CTA, so the kernel was bound by memory-level
// parallelism, not by MUFU, issue slots or the tensor pipe.  Here a dedicated TMA warp runs ahead across item
// boundaries (Q double-buffered, 3-stage K/V ring), keeping loads of the next item in flight under the current one.
//
// Roles (192 threads):
//   warps 0-3 : softmax; thread r owns query row r (TMEM lane r).  One TMEM pass per 64-column KV tile, ex2.approx in
//               fp32, P packed to fp16 into a 128B-swizzled smem tile (A operand of the PV MMA).  O accumulates in
//               TMEM; when a row's running max grows the warp rescales its O rows in place (tcgen05.ld/st).
//   warp 4    : lane 0 = TMA producer (Q, K/V tiles).
//   warp 5    : lane 0 = tcgen05 issuer: S = Q K^T (M128 N64 K64) and O += P V (M128 N64 K64, V consumed MN-major
//               straight from its row-major [kv][64] tile).  S (TMEM) and P (smem) are double-buffered.
// Causality skips KV tiles above the diagonal; warps whose rows see nothing of a diagonal half-tile skip its
// exponentials.
//
// Softmax details: (1) LAZY rescaling — a row's reference maximum only moves when the running maximum outgrows it by
// more than 2^kTau, so P = 2^(s - m_ref) <= 2^kTau and the O rescale round trip through TMEM almost never happens after
// the first tile; (2) exp2 on packed fp16 pairs (one MUFU op per two probabilities; P is consumed as fp16 anyway);
// (3) the item epilogue (O / l -> fp16 -> global) is deferred until after the softmax of the next item's first tile
// and goes through a warp-private transpose in the free P buffer, so no CTA-level barrier and no wait for the last PV.
//
// Measured alternatives (B200, B=64 T=500 S=6, profiles/r01_attention_study.md): eight softmax warps with two threads
// per row exchanging maxima through spare TMEM columns; three small strictly serial CTAs per SM with P kept in TMEM as
// the A operand of the PV MMA.  Both were correct and no faster: with five 64-wide KV tiles per item the kernel is
// bound by the per-tile handshake chain (MMA commit -> mbarrier -> TMEM read -> ... -> arrive -> MMA issue), not by
// TMEM bandwidth, MUFU, DRAM or the tensor pipe.
#include "once.h"
#include <stdlib.h>

#include "attn.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

constexpr int kTile = 128;                 // query rows per item
constexpr int kKV = 64;                    // KV rows per tile
constexpr int kStages = 3;
constexpr int kQBytes = kTile * 64 * 2;    // 16 KB
constexpr int kKBytes = kKV * 64 * 2;      // 8 KB
constexpr int kStageBytes = 2 * kKBytes;   // K then V
constexpr int kPBytes = kTile * kKV * 2;   // 16 KB
constexpr int kOffQ = 0;                                   // 2 buffers
constexpr int kOffKV = kOffQ + 2 * kQBytes;
constexpr int kOffP = kOffKV + kStages * kStageBytes;      // 2 buffers
constexpr int kOffBar = kOffP + 2 * kPBytes;               // mbarriers + tmem slot at the tail of dynamic smem
constexpr int kSmemBytes = kOffBar + 256;                  // 112 KB + 256 B
constexpr uint32_t kTmemCols = 256;        // S0 [0,64), S1 [64,128), O [128,192)
constexpr float kTau = 8.f;                // lazy-rescale threshold (log2 units): P <= 2^8

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// two exponentials per MUFU op: packed fp16 in, packed fp16 out
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ __half2 as_half2(uint32_t x) { return *reinterpret_cast<__half2*>(&x); }

// One work item: coordinates and KV extent.
struct Item {
  int q0, h, b, s, n_kv, kv_first, last_key;
};
__device__ __forceinline__ Item decode_item(const AttnParams& p, int id) {
  Item it;
  if (p.mode == ATTN_CAUSAL) {
    // Item order (L2 locality): the query tiles of one (sequence, head) are adjacent, heaviest first, so the CTAs that
    // share its K/V run at the same time and K/V come from DRAM once (ordering all heavy tiles of the whole batch first
    // re-read K/V from DRAM for every query tile: 663 MB instead of 370 MB per launch at B=64, S=6).  The lightest tile
    // (qt = 0) of every (sequence, head) goes to a second phase at the end, so the tail of the persistent loop is made
    // of the cheapest items.  The grid size is chosen coprime to n_qt - 1 (launch_attn), so the static round-robin
    // hands every CTA an even mix of tile weights.
    const int n_qt = (p.T + kTile - 1) / kTile;
    const int n_sh = p.H * p.B * p.S;
    const int heavy = n_qt - 1;
    int qt, rest;
    if (p.order == 0) {
      qt = n_qt - 1 - id / n_sh;
      rest = id % n_sh;
    } else if (id < n_sh * heavy) {
      rest = id / heavy;
      qt = n_qt - 1 - (id - rest * heavy);
    } else {
      rest = id - n_sh * heavy;
      qt = 0;
    }
    it.h = rest % p.H;
    const int z = rest / p.H;
    it.b = z / p.S;
    it.s = z % p.S;
    it.q0 = qt * kTile;
    it.last_key = min(it.q0 + kTile - 1 + p.mask_delay, p.T - 1);
    it.n_kv = it.last_key / kKV + 1;
    it.kv_first = 0;
  } else {
    it.h = id % p.H;
    it.b = 0;
    it.s = 0;
    it.q0 = (id / p.H) * p.tile_rows;               // T = total rows; tile_rows = (128 / S) * S
    it.last_key = min(it.q0 + p.tile_rows, p.T) - 1 - it.q0;   // tile-relative
    it.n_kv = it.last_key / kKV + 1;
    it.kv_first = it.q0;
  }
  return it;
}

__global__ void __launch_bounds__(192, 2)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
            __half* __restrict__ out, const AttnParams p, const int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;      // [2]
  uint64_t* q_empty = bars + 2;     // [2]
  uint64_t* kv_full = bars + 4;     // [3]
  uint64_t* kv_empty = bars + 7;    // [3]
  uint64_t* s_full = bars + 10;     // [2]
  uint64_t* p_ready = bars + 12;    // [2]
  uint64_t* pv_done = bars + 14;    // [2]
  uint64_t* o_free = bars + 16;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("[fseend] attn: dynamic smem base not 1024-aligned\n");
    __trap();
  }

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(o_free, 128);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 5) tmem_alloc(tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t tmem_O = tmem_base + 128;

  // Role warps stay converged and elect one lane only around the TMA / tcgen05 instructions (operands then live in
  // uniform registers; a single diverged lane makes ptxas wrap each UTCHMMA in an ELECT loop with R2UR moves).
  if (warp == 4) {
    {
      // ------------------------------------------------------------ TMA producer (runs ahead across items)
      uint32_t g = 0;   // KV tiles issued so far (ring position)
      uint32_t n = 0;   // items issued so far
      for (int id = blockIdx.x; id < n_items; id += gridDim.x, ++n) {
        const Item it = decode_item(p, id);
        const int qb = n & 1;
        mbar_wait(&q_empty[qb], ((n >> 1) & 1) ^ 1, 10);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[qb], kQBytes);
          tma_load_4d(smem + kOffQ + qb * kQBytes, &tmQ, &q_full[qb], it.h * 64, it.s, it.q0, it.b);
        }
        __syncwarp();
        for (int j = 0; j < it.n_kv; ++j, ++g) {
          const int st = g % kStages;
          mbar_wait(&kv_empty[st], ((g / kStages) & 1) ^ 1, 11);
          uint8_t* dst = smem + kOffKV + st * kStageBytes;
          const int row = it.kv_first + j * kKV;
          if (elect_one()) {
            mbar_arrive_expect_tx(&kv_full[st], kStageBytes);
            tma_load_4d(dst, &tmKV, &kv_full[st], 256 + it.h * 64, it.s, row, it.b);
            tma_load_4d(dst + kKBytes, &tmKV, &kv_full[st], 512 + it.h * 64, it.s, row, it.b);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    {
      // ------------------------------------------------------------ MMA issuer (warp-converged)
      constexpr uint32_t idesc_qk = make_idesc_f16(128, kKV, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, 64, true);
      uint32_t g = 0, n = 0;
      for (int id = blockIdx.x; id < n_items; id += gridDim.x, ++n) {
        const Item it = decode_item(p, id);
        const int qb = n & 1;
        const uint64_t qdesc = smem_desc_sw128(smem_u32(smem + kOffQ + qb * kQBytes));
        auto issue_qk = [&](uint32_t gj, bool last_qk) {
          const int st = gj % kStages;
          mbar_wait(&kv_full[st], (gj / kStages) & 1, 12);
          tc_fence_after();
          const uint64_t kdesc = smem_desc_sw128(smem_u32(smem + kOffKV + st * kStageBytes));
          const uint32_t tmem_S = tmem_base + (gj & 1) * kKV;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_S, qdesc + 2 * kk, kdesc + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
            umma_commit(&s_full[gj & 1]);
            if (last_qk) umma_commit(&q_empty[qb]);   // every QK^T of this item has been issued
          }
          __syncwarp();
        };
        mbar_wait(&q_full[qb], (n >> 1) & 1, 13);
        issue_qk(g, it.n_kv == 1);
        if (it.n_kv > 1) issue_qk(g + 1, it.n_kv == 2);
        for (int j = 0; j < it.n_kv; ++j) {
          const uint32_t gj = g + j;
          const int st = gj % kStages, pb = gj & 1;
          mbar_wait(&p_ready[pb], (gj >> 1) & 1, 14);   // P(j) in smem, S(j) fully read, O rescaled if needed
          if (j == 0) mbar_wait(o_free, (n & 1) ^ 1, 15);   // the previous item's O has been read out
          tc_fence_after();
          const int valid_cols = min(kKV, it.last_key - j * kKV + 1);
          const int n_k16 = (valid_cols + 15) >> 4;
          const uint64_t pdesc = smem_desc_sw128(smem_u32(smem + kOffP + pb * kPBytes));
          const uint64_t vdesc = smem_desc_sw128(smem_u32(smem + kOffKV + st * kStageBytes + kKBytes));
          if (elect_one()) {
            for (int kk = 0; kk < n_k16; ++kk) {
              // V is MN-major: 16 kv rows = 16 * 128 B = 2048 B per K step -> +128 in 16-byte units
              umma_f16(tmem_O, pdesc + 2 * kk, vdesc + 128 * kk, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&pv_done[pb]);
            umma_commit(&kv_empty[st]);
          }
          __syncwarp();
          if (j + 2 < it.n_kv) issue_qk(gj + 2, j + 3 == it.n_kv);   // S(j) has been consumed: its TMEM buffer is free
        }
        g += it.n_kv;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax warps
    const int r = tid;                 // query row inside the tile
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const float sl = p.scale * 1.4426950408889634f;
    // Item epilogue, DEFERRED by one KV tile: the O rows of item n are read out, normalised and stored after the
    // softmax of the first tile of item n+1 (whose S = Q K^T was issued right behind the last PV of item n), so the
    // softmax warps never sit waiting for the last PV to retire.
    bool pend = false;
    float pend_inv = 0.f;
    __half* pend_base = nullptr;    // output address of row 0 of the pending tile (this head's 64 columns)
    int pend_rows = 0;              // rows of the tile that exist (sequence end / block-diagonal tile_rows)
    const size_t out_stride = p.mode == ATTN_CAUSAL ? static_cast<size_t>(p.S) * 256 : 256;   // between tile rows
    auto flush_pending = [&](uint32_t pend_gl) {   // pend_gl: last KV tile (running count) of the pending item
      mbar_wait(&pv_done[pend_gl & 1], (pend_gl >> 1) & 1, 23);   // every MMA of the pending item has completed
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld32(tmem_O + lane_base, o0);
      tmem_ld32(tmem_O + lane_base + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free);          // O has been read out: the next item's first PV may overwrite it
      // Normalise, pack to fp16 and transpose through this warp's own 4 KB slice of the free P buffer (the one the
      // pending item's last PV read; its next writer is this same warp, for the following KV tile, so __syncwarp
      // ordering is enough): thread r parks its 128-byte row, then 8 lanes store one row, so every global store
      // instruction covers four whole 128-byte lines (a 16-byte store per thread at a row stride costs 32 LSU passes).
      uint8_t* stg = smem + kOffP + (pend_gl & 1) * kPBytes;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_half2(__uint_as_float(o0[q * 8 + 0]) * pend_inv, __uint_as_float(o0[q * 8 + 1]) * pend_inv);
        u.y = pack_half2(__uint_as_float(o0[q * 8 + 2]) * pend_inv, __uint_as_float(o0[q * 8 + 3]) * pend_inv);
        u.z = pack_half2(__uint_as_float(o0[q * 8 + 4]) * pend_inv, __uint_as_float(o0[q * 8 + 5]) * pend_inv);
        u.w = pack_half2(__uint_as_float(o0[q * 8 + 6]) * pend_inv, __uint_as_float(o0[q * 8 + 7]) * pend_inv);
        *reinterpret_cast<uint4*>(stg + sw128_offset(r, q)) = u;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_half2(__uint_as_float(o1[q * 8 + 0]) * pend_inv, __uint_as_float(o1[q * 8 + 1]) * pend_inv);
        u.y = pack_half2(__uint_as_float(o1[q * 8 + 2]) * pend_inv, __uint_as_float(o1[q * 8 + 3]) * pend_inv);
        u.z = pack_half2(__uint_as_float(o1[q * 8 + 4]) * pend_inv, __uint_as_float(o1[q * 8 + 5]) * pend_inv);
        u.w = pack_half2(__uint_as_float(o1[q * 8 + 6]) * pend_inv, __uint_as_float(o1[q * 8 + 7]) * pend_inv);
        *reinterpret_cast<uint4*>(stg + sw128_offset(r, 4 + q)) = u;
      }
      __syncwarp();
      {
        const int lane = tid & 31, sub = lane >> 3, ch = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = warp * 32 + i * 4 + sub;           // row of the tile handled by this lane in pass i
          const uint4 u = *reinterpret_cast<const uint4*>(stg + sw128_offset(rr, ch));
          if (rr < pend_rows) *reinterpret_cast<uint4*>(pend_base + static_cast<size_t>(rr) * out_stride + ch * 8) = u;
        }
      }
      __syncwarp();
      pend = false;
    };
    uint32_t g = 0, n = 0;
    for (int id = blockIdx.x; id < n_items; id += gridDim.x, ++n) {
      const Item it = decode_item(p, id);
      const int q0 = it.q0;
      // reference maximum in scaled log2 units (lazy: moves only when outgrown by more than kTau) and row sum
      float m_ref = -INFINITY, l_run = 0.f;

      for (int j = 0; j < it.n_kv; ++j) {
        const uint32_t gj = g + j;
        const int pb = gj & 1;
        // visible columns of this KV tile for this row: [lo, hi] (tile-relative)
        int lo, hi, whi, wlo;   // whi / wlo: largest / smallest hi within the warp (warp-uniform)
        bool full_tile;         // CTA-uniform: every row sees all 64 columns
        if (p.mode == ATTN_CAUSAL) {
          lo = 0;
          hi = min(q0 + r + p.mask_delay, p.T - 1) - j * kKV;
          whi = min(q0 + warp * 32 + 31 + p.mask_delay, p.T - 1) - j * kKV;
          wlo = min(q0 + warp * 32 + p.mask_delay, p.T - 1) - j * kKV;
          full_tile = (j * kKV + kKV - 1) <= min(q0 + p.mask_delay, p.T - 1);
        } else {
          const int flo = (r / p.S) * p.S;                       // first row of this row's frame (128-tile relative)
          int fhi = (r < p.tile_rows) ? flo + p.S - 1 : -1;
          if (q0 + fhi >= p.T) fhi = p.T - 1 - q0;
          lo = flo - j * kKV;
          hi = fhi - j * kKV;
          whi = kKV;                                             // frames straddle tiles: no warp-level skipping
          wlo = -1;
          full_tile = false;
        }
        uint8_t* ptile = smem + kOffP + pb * kPBytes;
        const uint32_t tmem_S = tmem_base + pb * kKV + lane_base;

        mbar_wait(&s_full[pb], (gj >> 1) & 1, 20);
        tc_fence_after();
        float alpha = 1.f, psum = 0.f;
        if (gj >= 2) mbar_wait(&pv_done[pb], ((gj - 2) >> 1) & 1, 21);   // PV(g-2) has consumed this P buffer
        // Per 32-column half of the tile, warp-uniformly: 0 = every row of the warp sees all 32 columns,
        // 1 = mixed (per-element compare), 2 = no row sees any of them (skip the exponentials, P = 0).
        int hmode[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
          hmode[hh] = full_tile ? 0 : ((hh * 32 + 31 <= wlo) ? 0 : ((hh * 32 > whi) ? 2 : 1));
        if (hmode[0] != 2 || hmode[1] != 2) {
          uint32_t sv[64];
          {
            uint32_t(&a0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
            uint32_t(&a1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
            tmem_ld32(tmem_S, a0);
            tmem_ld32(tmem_S + 32, a1);
            tmem_ld_wait();
          }
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (hmode[hh] == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int c = hh * 32 + i;
                sv[c] = (c >= lo && c <= hi) ? sv[c] : 0xff800000u;   // -inf
              }
            }
            if (hmode[hh] != 2) {
              // four independent FMNMX3 chains (a single running max is a 64-deep dependent chain)
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                mx4[0] = fmax3(mx4[0], __uint_as_float(sv[hh * 32 + i + 0]), __uint_as_float(sv[hh * 32 + i + 1]));
                mx4[1] = fmax3(mx4[1], __uint_as_float(sv[hh * 32 + i + 2]), __uint_as_float(sv[hh * 32 + i + 3]));
                mx4[2] = fmax3(mx4[2], __uint_as_float(sv[hh * 32 + i + 4]), __uint_as_float(sv[hh * 32 + i + 5]));
                mx4[3] = fmax3(mx4[3], __uint_as_float(sv[hh * 32 + i + 6]), __uint_as_float(sv[hh * 32 + i + 7]));
              }
            }
          }
          const float m_tile = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sl;   // -inf stays -inf
          if (m_tile > m_ref + kTau) {           // also the first visible tile of the row (m_ref = -inf)
            alpha = ex2(m_ref - m_tile);         // m_ref = -inf -> 0
            m_ref = m_tile;
          }
          const float m_sub = (m_ref == -INFINITY) ? 0.f : m_ref;
          // P = 2^(s*sl - m_ref) as packed fp16 pairs; the row sum is taken from the rounded values the PV MMA
          // consumes: groups of 8 in fp16, then fp32
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (hmode[hh] == 2) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(ptile + sw128_offset(r, hh * 4 + q)) = make_uint4(0, 0, 0, 0);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint32_t e[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const int c = hh * 32 + q * 8 + 2 * t;
                  e[t] = ex2_h2(pack_half2(fmaf(__uint_as_float(sv[c]), sl, -m_sub),
                                           fmaf(__uint_as_float(sv[c + 1]), sl, -m_sub)));
                }
                *reinterpret_cast<uint4*>(ptile + sw128_offset(r, hh * 4 + q)) = make_uint4(e[0], e[1], e[2], e[3]);
                const float2 f = __half22float2(__hadd2(__hadd2(as_half2(e[0]), as_half2(e[1])),
                                                        __hadd2(as_half2(e[2]), as_half2(e[3]))));
                psum += f.x + f.y;
              }
            }
          }
        } else {
          // no row of this warp sees any column of this tile (upper part of a diagonal tile): P = 0
#pragma unroll
          for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(ptile + sw128_offset(r, q)) = make_uint4(0, 0, 0, 0);
        }
        if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
          // rare: some row of this warp moved its reference maximum: rescale the warp's 32 O rows in TMEM.  O must be
          // quiescent: PV(j-1) (the last MMA issued so far that writes O) has to be complete.
          mbar_wait(&pv_done[(gj - 1) & 1], ((gj - 1) >> 1) & 1, 22);
          tc_fence_after();
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

        fence_proxy_async_smem();   // P visible to the tensor-core (async) proxy
        tc_fence_before();          // order our tcgen05.ld/st before the MMAs issued after the barrier
        mbar_arrive(&p_ready[pb]);
        if (j == 0 && pend) flush_pending(g - 1);   // previous item's epilogue, overlapped with this item's QK^T / PV
      }
      g += it.n_kv;
      // ---- remember this item's epilogue (see flush_pending)
      pend = true;
      pend_inv = l_run > 0.f ? 1.f / l_run : 0.f;
      if (p.mode == ATTN_CAUSAL) {
        pend_base = out + ((static_cast<size_t>(it.b) * p.T + q0) * p.S + it.s) * 256 + it.h * 64;
        pend_rows = min(kTile, p.T - q0);
      } else {
        pend_base = out + static_cast<size_t>(q0) * 256 + it.h * 64;
        pend_rows = min(p.tile_rows, p.T - q0);
      }
    }
    if (pend) flush_pending(g - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

// attn2.cu: two query tiles per CTA, 128-key KV tiles (causal mode)
void launch_attn2(const CUtensorMap& tmQ, const CUtensorMap& tmKV, __half* out, const AttnParams& p,
                  cudaStream_t stream);
void launch_attn3(const CUtensorMap& tmQ, __half* out, const AttnParams& p, int step_keys, cudaStream_t stream);

void launch_attn(const CUtensorMap& tmQ, const CUtensorMap& tmKV, __half* out, const AttnParams& p,
                 cudaStream_t stream) {
  if (p.mode == ATTN_CAUSAL) {
    // FSEEND_ATTN: 1 = one query tile per CTA with role warps (this file), 2 = query-tile pairs with role warps
    // (attn2.cu), 3 = self-contained warpgroup per query tile with whole-step softmax (attn3.cu;
    // FSEEND_ATTN_STEP = 128 (default, 4 CTAs / SM) or 256 (2 CTAs / SM) keys per step).
    // Default (measured, profiles/r02_attention_study.md): attn3 where there are enough (sequence, head, tile) items to
    // balance 4 persistent CTAs per SM dynamically (the decoder: 6144 items), attn2 for small launches (the encoder).
    const char* e = getenv("FSEEND_ATTN");
    int variant = (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 0;
    if (variant == 0) {
      const long long items = 1ll * ((p.T + 127) / 128) * p.H * p.B * p.S;
      variant = items >= 4096 ? 3 : 2;
    }
    if (variant == 3) {
      const char* s = getenv("FSEEND_ATTN_STEP");
      launch_attn3(tmQ, out, p, (s && atoi(s) == 256) ? 256 : 128, stream);
      return;
    }
    if (variant == 2) {
      launch_attn2(tmQ, tmKV, out, p, stream);
      return;
    }
  }
  static int num_sms = 0;
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int n_items;
  if (p.mode == ATTN_CAUSAL) n_items = ((p.T + kTile - 1) / kTile) * p.H * p.B * p.S;
  else n_items = ((p.T + p.tile_rows - 1) / p.tile_rows) * p.H;
  int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
  if (const char* e = getenv("FSEEND_ATTN_GRID")) grid = atoi(e) < grid ? atoi(e) : grid;   // experiments
  AttnParams pp = p;
  if (const char* e = getenv("FSEEND_ATTN_ORDER")) pp.order = atoi(e);
  if (p.mode == ATTN_CAUSAL && pp.order == 1) {
    const int heavy = (p.T + kTile - 1) / kTile - 1;
    auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
    while (heavy > 1 && grid > 1 && gcd(grid, heavy) != 1) --grid;   // see decode_item
  }
  attn_kernel<<<grid, 192, kSmemBytes, stream>>>(tmQ, tmKV, out, pp, n_items);
}

}  // namespace fseend
