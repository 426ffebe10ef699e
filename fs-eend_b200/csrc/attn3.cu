// Causal tcgen05 attention, head_dim 64 — third structure (attn.cu: one query tile per CTA with producer / MMA role
// warps; attn2.cu: a pair of query tiles per CTA).
//
// What the round-1 profile of attn2 said (profiles/r01_final_attn_lines.txt): 63 % of the stall samples are long
// scoreboard, 44 % of ALL samples sit in mbar_try_wait — a tile step is a chain commit -> mbarrier -> softmax warps wake
// -> P written -> arrive -> MMA warp wakes -> issue, about 5700 clk per 128 x 64 tile for 256 clk of tensor work, and the
// bare chain (no softmax at all) already costs 1950 clk per tile.  This kernel removes handshakes instead of hiding them:
//   * a work item is ONE query tile (128 rows) of one (sequence, head), processed by ONE self-contained warpgroup: the
//     same 128 threads do the softmax and (an elected lane of warp 0) issue the TMA loads and the tcgen05.mma — no
//     producer warp, no MMA warp, no cross-warp mbarrier round trips; the only waits are on hardware completions (TMA
//     bytes landed, MMAs committed) plus one CTA barrier per MMA hand-over;
//   * keys are taken kStep (256 or 128) at a time: S = Q K^T for the whole step is one batch of MMAs, the row maximum of
//     the step is EXACT (two passes over the S columns in TMEM: no lazy-rescale bookkeeping), P goes back into the S
//     columns as packed fp16 and is consumed from TMEM by the PV MMAs; T = 500 is 6 steps per (sequence, head) instead
//     of 20 tile steps, i.e. 12 hardware waits instead of 40 software handshakes;
//   * O lives in REGISTERS across steps (the step's P V product is read out of TMEM and folded with the usual online
//     rescale), so a chain needs only kStep TMEM columns (S, P and the step's O alias) and two CTAs per SM hold two
//     independent chains in 512 columns: one chain's softmax overlaps the other's MMAs;
//   * persistent CTAs pull items from an atomic counter in (sequence, head)-major order, heaviest tile first: the four
//     query tiles of a (sequence, head) run at about the same time on different SMs and share its K/V through L2, and the
//     next item's Q / K are prefetched while the current item finishes.
// Reference arithmetic: nn.MultiheadAttention with the additive causal mask, FS-EEND/nnet/model/onl_tfm_..._l2norm.py:
// 147-155 and nnet/modules/merge_tfm_encoder.py:379-385 (softmax in fp32, P rounded to fp16 for the PV product, row sums
// taken from the rounded P).
#include <stdlib.h>

#include "once.h"
#include "attn.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

constexpr int kTile = 128;                 // query rows per item
constexpr int kQBytes = kTile * 64 * 2;    // 16 KB
constexpr int kThreads = 128;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ __half2 as_h2(uint32_t x) { return *reinterpret_cast<__half2*>(&x); }
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
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

struct Item3 {
  int h, b, s, q0, n_keys;     // n_keys: keys [0, n_keys) are visible to at least one row of the tile
};
__device__ __forceinline__ Item3 decode3(const AttnParams& p, int id) {
  const int n_qt = (p.T + kTile - 1) / kTile;
  const int grp = id / n_qt;                       // (sequence, head), heads fastest
  const int qt = n_qt - 1 - (id - grp * n_qt);     // heaviest query tile of the group first
  Item3 it;
  it.h = grp % p.H;
  const int z = grp / p.H;
  it.s = z % p.S;
  it.b = z / p.S;
  it.q0 = qt * kTile;
  it.n_keys = min(it.q0 + kTile - 1 + p.mask_delay, p.T - 1) + 1;
  return it;
}

template <int kStep>
__global__ void __launch_bounds__(kThreads, kStep == 256 ? 2 : 4)
attn3_kernel(const __grid_constant__ CUtensorMap tmQ, __half* __restrict__ out, const AttnParams p, const int n_items,
             int* __restrict__ counter) {
  constexpr int kKBytes = kStep * 64 * 2;            // K (or V) tile of one step
  constexpr int kOffK = kQBytes, kOffV = kOffK + kKBytes, kOffBar = kOffV + kKBytes;
  constexpr uint32_t kTmemCols = kStep;              // S fp32 [0, kStep); P fp16 [0, kStep/2); step O [kStep/2, kStep/2+64)
  constexpr uint32_t kColO = kStep / 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* s_full = bars + 3;
  uint64_t* o_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  int* next_slot = reinterpret_cast<int*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int wq = warp;                                       // TMEM lane quarter
  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[fseend] attn3: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_full, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    next_slot[0] = atomicAdd(counter, 1);
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
  const uint32_t tS = tmem_base + lane_base;

  constexpr uint32_t idesc_pv = make_idesc_f16(128, 64, true);
  const uint64_t qdesc = smem_desc_sw128(smem_u32(smem));
  const uint64_t kdesc = smem_desc_sw128(smem_u32(smem + kOffK));
  const uint64_t vdesc = smem_desc_sw128(smem_u32(smem + kOffV));
  const float sl = p.scale * 1.4426950408889634f;
  const size_t out_stride = static_cast<size_t>(p.S) * 256;
  const int r = wq * 32 + (tid & 31);                  // query row inside the tile = TMEM lane
  constexpr int kOC = 64;                              // O columns folded / stored by this thread

  // loads of `rows` keys starting at key k0 (two 128-row boxes for a 256-key step); issued by the elected leader lane
  auto load_kv = [&](const Item3& it, int k0, int rows, int col, uint8_t* dst, uint64_t* bar) {
    mbar_arrive_expect_tx(bar, static_cast<uint32_t>(((rows + 127) / 128) * kQBytes));
    for (int h0 = 0; h0 < rows; h0 += 128)
      tma_load_4d(dst + (h0 / 128) * kQBytes, &tmQ, bar, col + it.h * 64, it.s, k0 + h0, it.b);
  };
  auto step_rows = [&](const Item3& it, int st) { return min(kStep, ((it.n_keys - st * kStep) + 127) / 128 * 128); };

  uint32_t ph_q = 0, ph_k = 0, ph_v = 0, ph_s = 0, ph_o = 0;    // phase bits of the five barriers
  int cur = next_slot[0];
  bool prefetched = false;                                       // Q / K(step 0) / V(step 0) of `cur` already in flight
  __syncthreads();

  while (cur < n_items) {
    const Item3 it = decode3(p, cur);
    const int n_steps = (it.n_keys + kStep - 1) / kStep;
    if (warp == 0) {
      if (elect_one()) {
        if (!prefetched) {
          mbar_arrive_expect_tx(q_full, kQBytes);
          tma_load_4d(smem, &tmQ, q_full, it.h * 64, it.s, it.q0, it.b);
          load_kv(it, 0, step_rows(it, 0), 256, smem + kOffK, k_full);
          load_kv(it, 0, step_rows(it, 0), 512, smem + kOffV, v_full);
        }
      }
      __syncwarp();
    }
    const int hi_row = min(it.q0 + r + p.mask_delay, p.T - 1);                    // last visible key of this row
    const int hi_w1 = min(it.q0 + wq * 32 + 31 + p.mask_delay, p.T - 1);          // ... of the warp's last row
    const int hi_w0 = min(it.q0 + wq * 32 + p.mask_delay, p.T - 1);               // ... of the warp's first row
    float m_run = -INFINITY, l_run = 0.f;
    float o_run[kOC];
#pragma unroll
    for (int i = 0; i < kOC; ++i) o_run[i] = 0.f;
    int nxt = n_items;

    for (int st = 0; st < n_steps; ++st) {
      const int k0 = st * kStep;
      const int rows = step_rows(it, st);                 // 128 or 256 key rows loaded / multiplied
      const bool last = st == n_steps - 1;
      // ---- S = Q K^T for the whole step
      if (warp == 0) {
        if (st == 0) {
          mbar_wait(q_full, ph_q, 30);
        }
        mbar_wait(k_full, ph_k, 31);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idesc_qk = make_idesc_f16(128, rows, false);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base, qdesc + 2 * kk, kdesc + 2 * kk, idesc_qk, kk > 0 ? 1u : 0u);
          umma_commit(s_full);
          // the index of the NEXT item: fetched here so that the atomic's round trip hides behind the MMAs just issued
          if (st == 0) next_slot[0] = atomicAdd(counter, 1);
        }
        __syncwarp();
      }
      if (st == 0) ph_q ^= 1;
      ph_k ^= 1;
      mbar_wait(s_full, ph_s, 32);
      ph_s ^= 1;
      tc_fence_after();
      // K (and after the item's last QK^T also Q) is free: fetch the next step's K, or the next item's Q and K
      if (warp == 0) {
        if (last) nxt = next_slot[0];
        if (elect_one()) {
          if (!last) {
            load_kv(it, k0 + kStep, step_rows(it, st + 1), 256, smem + kOffK, k_full);
          } else if (nxt < n_items) {
            const Item3 ni = decode3(p, nxt);
            mbar_arrive_expect_tx(q_full, kQBytes);
            tma_load_4d(smem, &tmQ, q_full, ni.h * 64, ni.s, ni.q0, ni.b);
            load_kv(ni, 0, step_rows(ni, 0), 256, smem + kOffK, k_full);
          }
        }
        __syncwarp();
      }
      // ---- softmax of this step: pass 1 = exact row maximum, pass 2 = P (packed fp16, back into the S columns)
      const int hi = hi_row - k0, whi = hi_w1 - k0, wlo = hi_w0 - k0;
      const int n_chunks = rows / 32;
      const int c_lo = 0, c_hi = n_chunks;
      float mx = -INFINITY;
      for (int cc = c_lo; cc < c_hi; ++cc) {
        if (cc * 32 > whi) break;                                  // no row of this warp sees this or any later chunk
        uint32_t sv[32];
        tmem_ld32(tS + cc * 32, sv);
        tmem_ld_wait();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (cc * 32 + 31 <= wlo) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            m4[0] = fmax3(m4[0], __uint_as_float(sv[i + 0]), __uint_as_float(sv[i + 1]));
            m4[1] = fmax3(m4[1], __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
            m4[2] = fmax3(m4[2], __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
            m4[3] = fmax3(m4[3], __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = (cc * 32 + i <= hi) ? __uint_as_float(sv[i]) : -INFINITY;
            m4[i & 3] = fmaxf(m4[i & 3], v);
          }
        }
        mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      }
      const float m_new = fmaxf(m_run, mx * sl);                   // key 0 is visible to every row: finite from step 0 on
      const float alpha = ex2f(m_run - m_new);                     // first step: 2^-inf = 0
      m_run = m_new;
      float psum = 0.f;
      for (int cc = c_lo; cc < c_hi; ++cc) {
        uint32_t pk[16];
        if (cc * 32 > whi) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
        } else {
          uint32_t sv[32];
          tmem_ld32(tS + cc * 32, sv);
          tmem_ld_wait();
          if (cc * 32 + 31 > wlo) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sv[i] = (cc * 32 + i <= hi) ? sv[i] : 0xff800000u;   // -inf -> P = 0
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              // fp32 ex2 then one packed conversion: ex2.approx.f16x2 is TWO MUFU.EX2.F16 plus PRMT unpack / repack on
              // this part (ncu: 6.9 M MUFU + 6.7 M PRMT warp-instructions per launch), i.e. no MUFU saving at all
              const int c = q * 8 + 2 * t;
              pk[q * 4 + t] = pack_half2(ex2f(fmaf(__uint_as_float(sv[c]), sl, -m_new)),
                                         ex2f(fmaf(__uint_as_float(sv[c + 1]), sl, -m_new)));
            }
            // row sum from the rounded values the PV MMA consumes: groups of 8 in fp16, then fp32
            const float2 f = __half22float2(__hadd2(__hadd2(as_h2(pk[q * 4]), as_h2(pk[q * 4 + 1])),
                                                    __hadd2(as_h2(pk[q * 4 + 2]), as_h2(pk[q * 4 + 3]))));
            psum += f.x + f.y;
          }
        }
        tmem_st16(tS + cc * 16, pk);          // over S columns this thread has already consumed (16 cc + 16 <= 32 cc + 32)
      }
      l_run = l_run * alpha + psum;
      tmem_st_wait();
      tc_fence_before();
      __syncthreads();                         // every row's P is in TMEM
      // ---- step O = P V, fresh accumulator
      if (warp == 0) {
        tc_fence_after();
        mbar_wait(v_full, ph_v, 33);
        tc_fence_after();
        if (elect_one()) {
          const int valid = min(rows, min(it.q0 + kTile - 1 + p.mask_delay, p.T - 1) - k0 + 1);   // keys the tile's last row sees
          const int n_k16 = (valid + 15) >> 4;
          for (int kk = 0; kk < n_k16; ++kk)     // A: 16 keys = 8 packed TMEM columns; V MN-major: 16 rows = +128 units
            umma_ts(tmem_base + kColO, tmem_base + 8 * kk, vdesc + 128 * kk, idesc_pv, kk > 0 ? 1u : 0u);
          umma_commit(o_full);
        }
        __syncwarp();
      }
      ph_v ^= 1;
      mbar_wait(o_full, ph_o, 34);
      ph_o ^= 1;
      tc_fence_after();
      // V is free: next step's V, or the next item's first V
      if (warp == 0) {
        if (elect_one()) {
          if (!last) {
            load_kv(it, k0 + kStep, step_rows(it, st + 1), 512, smem + kOffV, v_full);
          } else if (nxt < n_items) {
            const Item3 ni = decode3(p, nxt);
            load_kv(ni, 0, step_rows(ni, 0), 512, smem + kOffV, v_full);
          }
        }
        __syncwarp();
      }
#pragma unroll
      for (int c = 0; c < kOC / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(tS + kColO + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_run[c * 32 + i] = fmaf(o_run[c * 32 + i], alpha, __uint_as_float(o[i]));
      }
      tc_fence_before();
      __syncthreads();                         // the step's O has been read out: the next QK^T may overwrite the columns
    }

    // ---- item epilogue: O / l -> fp16 -> global; each thread owns kOC columns of its row (full 32-byte sectors)
    {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      const int rows_valid = min(kTile, p.T - it.q0);
      if (r < rows_valid) {
        __half* dst = out + ((static_cast<size_t>(it.b) * p.T + it.q0) * p.S + it.s) * 256 + it.h * 64 +
                      static_cast<size_t>(r) * out_stride;
#pragma unroll
        for (int q = 0; q < kOC / 16; ++q) {
          uint4 u, v;
          u.x = pack_half2(o_run[q * 16 + 0] * inv, o_run[q * 16 + 1] * inv);
          u.y = pack_half2(o_run[q * 16 + 2] * inv, o_run[q * 16 + 3] * inv);
          u.z = pack_half2(o_run[q * 16 + 4] * inv, o_run[q * 16 + 5] * inv);
          u.w = pack_half2(o_run[q * 16 + 6] * inv, o_run[q * 16 + 7] * inv);
          v.x = pack_half2(o_run[q * 16 + 8] * inv, o_run[q * 16 + 9] * inv);
          v.y = pack_half2(o_run[q * 16 + 10] * inv, o_run[q * 16 + 11] * inv);
          v.z = pack_half2(o_run[q * 16 + 12] * inv, o_run[q * 16 + 13] * inv);
          v.w = pack_half2(o_run[q * 16 + 14] * inv, o_run[q * 16 + 15] * inv);
          st_global_256(dst + q * 16, u, v);
        }
      }
    }
    // next item (its index was published before this item's last barrier; every thread reads it after that barrier)
    cur = next_slot[0];
    prefetched = cur < n_items;
    __syncthreads();                           // nobody still reads next_slot when warp 0 overwrites it
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

constexpr int kCounterRing = 256;

}  // namespace

// tmQ: box (64,1,128,1) over (768, S, T, B).  step_keys: 128 (4 CTAs / SM, default) or 256 (2 CTAs / SM).
void launch_attn3(const CUtensorMap& tmQ, __half* out, const AttnParams& p, int step_keys, cudaStream_t stream) {
  static int num_sms = 0;
  static int* counters_dev[64] = {};
  static int next_counter = 0;
  static PerDeviceOnce once;
  constexpr int smem256 = kQBytes + 2 * 256 * 128 + 128, smem128 = kQBytes + 2 * 128 * 128 + 128;
  if (once.first()) {
    cudaFuncSetAttribute(attn3_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem256);
    cudaFuncSetAttribute(attn3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem128);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaMalloc(&counters_dev[once.device], kCounterRing * sizeof(int));
  }
  int* counters = counters_dev[once.device];
  // one work counter per launch from a ring (launches on different streams may overlap; 256 launches in flight never do)
  int* counter = counters + next_counter;
  next_counter = (next_counter + 1) % kCounterRing;
  cudaMemsetAsync(counter, 0, sizeof(int), stream);
  const int n_qt = (p.T + kTile - 1) / kTile;
  const int n_items = n_qt * p.H * p.B * p.S;
  if (step_keys == 256) {
    const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
    attn3_kernel<256><<<grid, kThreads, smem256, stream>>>(tmQ, out, p, n_items, counter);
  } else {
    const int grid = n_items < 4 * num_sms ? n_items : 4 * num_sms;
    attn3_kernel<128><<<grid, kThreads, smem128, stream>>>(tmQ, out, p, n_items, counter);
  }
}

}  // namespace fseend
