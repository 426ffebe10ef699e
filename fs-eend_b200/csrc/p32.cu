// Parity-precision kernels (see p32.cuh): split-precision tcgen05 GEMM on fp32 activations, fp32 retention core,
// fp32 normalisation / activation kernels.  Reference arithmetic restated per kernel below (paths relative to
// /root/reference/LS-EEND/nnet).
#include "once.h"
#include "p32.cuh"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "ptx.cuh"

namespace fseend {

namespace {

// =====================================================================================================================
// Split-precision GEMM.  One CTA (256 threads) = one 128 x 128 output tile.
//   all 8 warps : A producer — fp32 rows from global (coalesced float4), split x = hi + lo (fp16), both halves written
//                 into 128B-swizzled K-major smem tiles (the layout TMA would have produced)
//   thread 0    : TMA for the two weight tiles (W_hi, W_lo: split once at model creation)
//   warp 1      : MMA issue — per 64-wide k block twelve tcgen05.mma into a FRESH TMEM accumulator: the eight small
//                 products first (A_lo W_hi, A_hi W_lo), then A_hi W_hi (lo.lo dropped: 2^-22 relative)
//   all 8 warps : fold the finished k-block accumulator into fp32 REGISTERS (tcgen05.ld + add), epilogue from registers
// Why registers: the tensor core truncates (round-toward-zero) when it adds into the fp32 TMEM accumulator — measured
// on B200, the error of a plain TMEM-resident accumulation grows linearly with the number of MMA steps (3e-5 absolute at
// K = 1024, 6.5e-5 at K = 4864 on O(1) outputs: ~17 bits).  With one fresh accumulator per k-block only the four
// A_hi W_hi steps truncate at full magnitude, and the sum across k-blocks is round-to-nearest on the CUDA cores.
// Two stages of (A_hi, A_lo, W_hi, W_lo) = 128 KB and two 128-column accumulators: the split of k-block i+1 and the
// fold of k-block i-1 overlap the MMAs of k-block i.
constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 2;
constexpr int kTileBytes = 128 * BK * 2;          // 16 KB
constexpr int kStageBytes = 4 * kTileBytes;       // A_hi | A_lo | W_hi | W_lo
constexpr int kGemmSmem = kStages * kStageBytes + 1024;
constexpr uint32_t kTmemCols = 256;            // two ping-pong 128-column accumulators

// x = hi + lo with both halves fp16.  Packed conversions only: cvt.rn.f16x2.f32 (F2FP, full rate) and HADD2.F32 —
// the scalar F2F.F16.F32 form runs on the quarter-rate conversion pipe and was 30 % of the first version's stall samples.
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split8(const float4 a, const float4 b, uint4& hi, uint4& lo) {
  split2(a.x, a.y, hi.x, lo.x);
  split2(a.z, a.w, hi.y, lo.y);
  split2(b.x, b.y, hi.z, lo.z);
  split2(b.z, b.w, hi.w, lo.w);
}

template <int kAct>
__device__ __forceinline__ float act_apply(float v) {
  if (kAct == P32_RELU) return fmaxf(v, 0.f);
  // swish: v * 1 / (1 + e^-v) with a correctly rounded reciprocal (no IEEE-division slow path)
  if (kAct == P32_SWISH) return v * __frcp_rn(1.f + expf(-v));
  return v;
}

template <int kAct, bool kTrain = false>
__global__ void __launch_bounds__(256, 1)
p32_gemm_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                const P32GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t done_bar[kStages];   // MMAs of a k-block complete: smem stage free + accumulator ready
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[BN];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = p.N / BN;
  const int n_tile = blockIdx.x % n_tiles;
  const int m_tile = blockIdx.x / n_tiles;
  const int tiles_per_seq = (p.rows_per_seq + BM - 1) / BM;
  const int seq = m_tile / tiles_per_seq;
  const int t0 = (m_tile % tiles_per_seq) * BM;
  const int n0 = n_tile * BN;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&done_bar[s], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmWhi);
    tma_prefetch_desc(&tmWlo);
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, kTmemCols);
  if (tid < BN) bias_s[tid] = p.bias ? __ldg(p.bias + n0 + tid) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int total_it = p.taps * p.k_blocks;
  const int a_off = p.a_row_offset + (p.a_row_offset_dev ? *p.a_row_offset_dev : 0);
  const float* a_seq = p.A + static_cast<size_t>(seq) * p.a_seq_rows * p.lda;
  constexpr uint32_t idesc = make_idesc_f16(BM, BN, false);

  // this thread's slice of the output tile: row (warp & 3) * 32 + lane, columns (warp >> 2) * 64 .. + 64
  const int cbase = (warp >> 2) * 64;
  const uint32_t lane_addr = (static_cast<uint32_t>((warp & 3) * 32) << 16) + cbase;
  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
  auto fold = [&](int buf) {      // acc += finished k-block accumulator `buf`
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t raw[32];
      tmem_ld32(tmem_base + buf * 128 + lane_addr + half * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[half * 32 + j] += __uint_as_float(raw[j]);
    }
  };

  // A tile of k-block `it`: 4 chunks of 8 floats per thread, fetched one iteration AHEAD into registers so that the
  // global-load latency hides behind the MMAs of the current k-block (the first version loaded and consumed in the same
  // iteration: ~1000 clk of exposed L2 latency per k-block)
  float4 pre[8];
  const float ascale = (kTrain && p.a_scale_dev) ? __ldg(p.a_scale_dev) : 1.f;
  auto fetch = [&](int it) {
    const int tap = it / p.k_blocks;
    const int kb = it - tap * p.k_blocks;
    if (kTrain && p.a_transposed) {
      // lanes walk the tile's rows m (consecutive floats of one source row): coalesced scalar loads
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int q = tid + 256 * i;
        const int m = q & 127, c = q >> 7;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
        if (t0 + m < p.rows_per_seq) {
          const int kbase = seq * (p.k_blocks * BK) + kb * BK + c * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j)      // raw loads only: anything computed here would serialise them on their latency
            if (kbase + j < p.a_k_rows) v[j] = __ldg(p.A + static_cast<size_t>(kbase + j) * p.lda + t0 + m);
        }
        pre[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
        pre[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = tid + 256 * i;
      const int r = q >> 3, c = q & 7;                 // tile row, 16-byte chunk (8 halfs) inside the 64-wide k block
      const int t = t0 + r;
      const int arow = t + tap + p.tap_shift + a_off;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (t < p.rows_per_seq && arow >= 0 && arow < p.a_seq_rows) {
        const size_t o = static_cast<size_t>(arow) * p.lda + kb * BK + c * 8;
        const float4* src = reinterpret_cast<const float4*>(a_seq + o);
        v0 = __ldg(src);
        v1 = __ldg(src + 1);
      }
      pre[2 * i] = v0;
      pre[2 * i + 1] = v1;
    }
  };
  fetch(0);

  for (int it = 0; it < total_it; ++it) {
    const int s = it & 1;
    const uint32_t ph = (it >> 1) & 1;
    if (it >= kStages) {
      mbar_wait(&done_bar[s], ph ^ 1, 71);   // the MMAs of iteration it-2 are complete: fold their accumulator,
      tc_fence_after();                      // then stage s and accumulator s are free
      fold(s);
      tc_fence_before();
    }
    const int tap = it / p.k_blocks;
    const int kb = it - tap * p.k_blocks;
    uint8_t* sAhi = smem + s * kStageBytes;
    uint8_t* sAlo = sAhi + kTileBytes;
    uint8_t* sWhi = sAlo + kTileBytes;
    uint8_t* sWlo = sWhi + kTileBytes;
    if (tid == 0) {
      mbar_arrive_expect_tx(&full_bar[s], 2 * kTileBytes);
      const int wrow = tap * p.N + n0 + (kTrain ? seq * p.w_seq_stride : 0);
      tma_load_2d(sWhi, &tmWhi, &full_bar[s], kb * BK, wrow);
      tma_load_2d(sWlo, &tmWlo, &full_bar[s], kb * BK, wrow);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = tid + 256 * i;
      const bool tr = kTrain && p.a_transposed;
      const int r = tr ? (q & 127) : (q >> 3), c = tr ? (q >> 7) : (q & 7);
      uint4 hi, lo;
      if (kTrain) {
        float4& a = pre[2 * i];
        float4& b = pre[2 * i + 1];
        a.x *= ascale; a.y *= ascale; a.z *= ascale; a.w *= ascale;
        b.x *= ascale; b.y *= ascale; b.z *= ascale; b.w *= ascale;
      }
      split8(pre[2 * i], pre[2 * i + 1], hi, lo);
      const uint32_t off = sw128_offset(r, c);
      *reinterpret_cast<uint4*>(sAhi + off) = hi;
      *reinterpret_cast<uint4*>(sAlo + off) = lo;
    }
    if (it + 1 < total_it) fetch(it + 1);
    fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    __syncthreads();               // (also orders every thread's fold of accumulator s before its re-use below)
    if (warp == 1) {
      mbar_wait(&full_bar[s], ph, 72);
      tc_fence_after();
      const uint64_t dAhi = smem_desc_sw128(smem_u32(sAhi)), dAlo = smem_desc_sw128(smem_u32(sAlo));
      const uint64_t dWhi = smem_desc_sw128(smem_u32(sWhi)), dWlo = smem_desc_sw128(smem_u32(sWlo));
      const uint32_t d = tmem_base + s * 128;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          umma_f16(d, dAlo + 2 * kk, dWhi + 2 * kk, idesc, kk > 0 ? 1u : 0u);
          umma_f16(d, dAhi + 2 * kk, dWlo + 2 * kk, idesc, 1u);
        }
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) umma_f16(d, dAhi + 2 * kk, dWhi + 2 * kk, idesc, 1u);
        umma_commit(&done_bar[s]);
      }
      __syncwarp();
    }
  }
  // drain the last (up to) two k-blocks in issue order
  for (int j = (total_it >= 2 ? total_it - 2 : 0); j < total_it; ++j) {
    mbar_wait(&done_bar[j & 1], (j >> 1) & 1, 73);
    tc_fence_after();
    fold(j & 1);
  }

  // ---------------- epilogue: registers -> padded fp32 tile in shared memory (the pipeline stages are idle now) ->
  // whole rows per warp, so that the residual loads and the output stores are 512-byte coalesced (the first version
  // stored 64 floats per THREAD: every store instruction touched 32 different rows with half-filled sectors)
  {
    constexpr int kLdT = BN + 4;                       // 132 floats: conflict-free float4 row writes and reads
    float* tile = reinterpret_cast<float*>(smem);
    __syncthreads();                                   // every warp has folded the last accumulators: stages are free
    {
      const int r = (warp & 3) * 32 + lane;
      float* trow = tile + r * kLdT + cbase;
#pragma unroll
      for (int j = 0; j < 64; j += 4) *reinterpret_cast<float4*>(trow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
    __syncthreads();
    const float4 bb = *reinterpret_cast<const float4*>(&bias_s[4 * lane]);
    const float wsc = p.w_inv_scale * ((kTrain && p.out_scale_dev) ? __ldg(p.out_scale_dev) : 1.f), alpha = p.alpha;
    const int col = n0 + 4 * lane;
    for (int r = warp; r < BM; r += 8) {
      const int t = t0 + r;
      if (t >= p.rows_per_seq) break;                  // rows are handed out in increasing order per warp
      const size_t orow = static_cast<size_t>(seq) * p.rows_per_seq + t;
      const float4 a = *reinterpret_cast<const float4*>(tile + r * kLdT + 4 * lane);
      float4 v;
      v.x = alpha * act_apply<kAct>(fmaf(a.x, wsc, bb.x));
      v.y = alpha * act_apply<kAct>(fmaf(a.y, wsc, bb.y));
      v.z = alpha * act_apply<kAct>(fmaf(a.z, wsc, bb.z));
      v.w = alpha * act_apply<kAct>(fmaf(a.w, wsc, bb.w));
      if (p.residual) {
        const float4 rr = *reinterpret_cast<const float4*>(p.residual + orow * p.ldr + col);
        v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
      }
      if (kTrain && p.out_mask) {
        const float4 mk = __ldg(reinterpret_cast<const float4*>(p.out_mask + orow * p.ldo + col));
        v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f; v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
      }
      *reinterpret_cast<float4*>(p.out + orow * p.ldo + col) = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// Persistent, warp-specialised form of the split-precision GEMM (the batch path's kernel; the one-tile kernel above stays
// for small grids).  One CTA (16 warps) per SM walks output tiles; inside it
//   warps 0-6  : A producers — fp32 rows from global, fetched TWO k-blocks ahead into registers (ncu on the first
//                persistent version, one k-block ahead: 2700 clk per k-block = the loaded-L2/HBM latency, tensor pipe 28 %
//                active; bandwidth x latency asks for ~100 KB in flight per SM, the register budget allows 64 KB), hi / lo split with packed
//                conversions, swizzled smem stores, arrive on the stage's `full` barrier; thread 0 also issues the TMA
//                for W_hi / W_lo
//   warp 7     : MMA issuer — 12 tcgen05.mma per k-block into one of FOUR 128-column TMEM accumulators (fresh per k-block:
//                the tensor core truncates on accumulate, see above), commits `empty[stage]` and `acc_full[buf]`
//   warps 8-15 : fold each finished k-block accumulator into fp32 registers (64 columns per thread, round-to-nearest) and
//                release it; after a tile's last k-block write the row through a padded smem tile and store whole rows
// so the operand split, the register fold and the epilogue of tile i all overlap the MMAs (of tile i+1).  16 warps keep
// the register budget at 128 per thread (17 warps are allocated as 20: 96 registers and spills in both hot roles).
constexpr int kPStages = 2, kPAcc = 4;
constexpr int kPLdT = BN + 4;
constexpr int kPersistSmem = kPStages * kStageBytes + BM * kPLdT * 4 + 1024;
constexpr int kPThreads = 16 * 32;
constexpr int kPProd = 7 * 32;                       // producer threads

template <int kAct, bool kTrain = false>
__global__ void __launch_bounds__(kPThreads, 1)
p32_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                        const P32GemmParams p, const int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kPStages];
  __shared__ __align__(8) uint64_t empty_bar[kPStages];
  __shared__ __align__(8) uint64_t acc_full[kPAcc];
  __shared__ __align__(8) uint64_t acc_empty[kPAcc];
  __shared__ uint32_t tmem_base_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* tile_s = reinterpret_cast<float*>(smem + kPStages * kStageBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kPStages; ++s) {
      mbar_init(&full_bar[s], kPProd);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < kPAcc; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 256);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmWhi);
    tma_prefetch_desc(&tmWlo);
  }
  if (warp == 7) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int n_tiles_n = p.N / BN;
  const int tiles_per_seq = (p.rows_per_seq + BM - 1) / BM;
  const int total_it = p.taps * p.k_blocks;
  const int a_off = p.a_row_offset + (p.a_row_offset_dev ? *p.a_row_offset_dev : 0);
  const int n_local = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const uint32_t G = static_cast<uint32_t>(n_local) * total_it;        // k-blocks this CTA processes, in order

  if (warp < 7) {
    // ------------------------------------------------------------ A producers (+ W TMA by thread 0)
    // k-block gi of this CTA = (tile blockIdx.x + (gi / total_it) * gridDim.x, it = gi % total_it)
    const float ascale = (kTrain && p.a_scale_dev) ? __ldg(p.a_scale_dev) : 1.f;
    auto fetch = [&](float4 (&buf)[10], uint32_t gi) {
      if (gi >= G) return;
      const int tile = blockIdx.x + (gi / total_it) * gridDim.x, it = gi % total_it;
      const int m_tile = tile / n_tiles_n;
      const int seq = m_tile / tiles_per_seq, t0 = (m_tile % tiles_per_seq) * BM;
      const int tap = it / p.k_blocks, kb = it - tap * p.k_blocks;
      if (kTrain && p.a_transposed) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int q = tid + kPProd * i;
          const int m = q & 127, c = q >> 7;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;
          if (q < 1024 && t0 + m < p.rows_per_seq) {
            const int kbase = seq * (p.k_blocks * BK) + kb * BK + c * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j)    // raw loads only: anything computed here would serialise them on their latency
              if (kbase + j < p.a_k_rows) v[j] = __ldg(p.A + static_cast<size_t>(kbase + j) * p.lda + t0 + m);
          }
          buf[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
          buf[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        return;
      }
      const float* a_seq = p.A + static_cast<size_t>(seq) * p.a_seq_rows * p.lda;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int q = tid + kPProd * i;                     // 1024 chunks of 8 floats over 224 threads
        const int r = q >> 3, c = q & 7;
        const int t = t0 + r;
        const int arow = t + tap + p.tap_shift + a_off;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (q < 1024 && t < p.rows_per_seq && arow >= 0 && arow < p.a_seq_rows) {
          const size_t o = static_cast<size_t>(arow) * p.lda + kb * BK + c * 8;
          const float4* src = reinterpret_cast<const float4*>(a_seq + o);
          v0 = __ldg(src);
          v1 = __ldg(src + 1);
        }
        buf[2 * i] = v0;
        buf[2 * i + 1] = v1;
      }
    };
    auto step = [&](float4 (&buf)[10], uint32_t gi) {
      const int s = gi % kPStages;
      mbar_wait(&empty_bar[s], ((gi / kPStages) & 1) ^ 1, 81);
      uint8_t* sAhi = smem + s * kStageBytes;
      uint8_t* sAlo = sAhi + kTileBytes;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int q = tid + kPProd * i;
        if (q < 1024) {
          const bool tr = kTrain && p.a_transposed;
          const int r = tr ? (q & 127) : (q >> 3), c = tr ? (q >> 7) : (q & 7);
          uint4 hi, lo;
          if (kTrain) {
            float4& a = buf[2 * i];
            float4& b = buf[2 * i + 1];
            a.x *= ascale; a.y *= ascale; a.z *= ascale; a.w *= ascale;
            b.x *= ascale; b.y *= ascale; b.z *= ascale; b.w *= ascale;
          }
          split8(buf[2 * i], buf[2 * i + 1], hi, lo);
          const uint32_t off = sw128_offset(r, c);
          *reinterpret_cast<uint4*>(sAhi + off) = hi;
          *reinterpret_cast<uint4*>(sAlo + off) = lo;
        }
      }
      fetch(buf, gi + 2);                                   // refill this register set two k-blocks ahead
      fence_proxy_async_smem();
      if (tid == 0) {
        const int tile = blockIdx.x + (gi / total_it) * gridDim.x, it = gi % total_it;
        const int n0 = (tile % n_tiles_n) * BN;
        const int tap = it / p.k_blocks, kb = it - tap * p.k_blocks;
        const int wrow = tap * p.N + n0 + (kTrain ? ((tile / n_tiles_n) / tiles_per_seq) * p.w_seq_stride : 0);
        mbar_arrive_expect_tx(&full_bar[s], 2 * kTileBytes);
        tma_load_2d(sAlo + kTileBytes, &tmWhi, &full_bar[s], kb * BK, wrow);
        tma_load_2d(sAlo + 2 * kTileBytes, &tmWlo, &full_bar[s], kb * BK, wrow);
      } else {
        mbar_arrive(&full_bar[s]);
      }
    };
    float4 b0[10], b1[10];
    fetch(b0, 0);
    fetch(b1, 1);
    for (uint32_t gi = 0; gi < G; gi += 2) {
      step(b0, gi);
      if (gi + 1 < G) step(b1, gi + 1);
    }
  } else if (warp == 7) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(BM, BN, false);
    for (uint32_t g = 0; g < G; ++g) {
      const int s = g % kPStages, buf = g % kPAcc;
      mbar_wait(&full_bar[s], (g / kPStages) & 1, 82);
      mbar_wait(&acc_empty[buf], ((g / kPAcc) & 1) ^ 1, 83);
      tc_fence_after();
      const uint32_t sb = smem_u32(smem + s * kStageBytes);
      const uint64_t dAhi = smem_desc_sw128(sb), dAlo = smem_desc_sw128(sb + kTileBytes);
      const uint64_t dWhi = smem_desc_sw128(sb + 2 * kTileBytes), dWlo = smem_desc_sw128(sb + 3 * kTileBytes);
      const uint32_t d = tmem_base + buf * 128;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          umma_f16(d, dAlo + 2 * kk, dWhi + 2 * kk, idesc, kk > 0 ? 1u : 0u);
          umma_f16(d, dAhi + 2 * kk, dWlo + 2 * kk, idesc, 1u);
        }
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) umma_f16(d, dAhi + 2 * kk, dWhi + 2 * kk, idesc, 1u);
        umma_commit(&empty_bar[s]);
        umma_commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ fold + epilogue warps 8-15: thread <-> (row, column half)
    const int wq = warp & 3, half = (warp - 8) >> 2;
    const int r = wq * 32 + lane;
    const uint32_t t_addr = (static_cast<uint32_t>(wq * 32) << 16) + half * 64;
    const int ew = warp - 8;                               // 0..7: rows ew, ew + 8, ... in the store phase
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m_tile = tile / n_tiles_n;
      const int seq = m_tile / tiles_per_seq, t0 = (m_tile % tiles_per_seq) * BM;
      const int n0 = (tile % n_tiles_n) * BN;
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
      for (int it = 0; it < total_it; ++it, ++g) {
        const int buf = g % kPAcc;
        mbar_wait(&acc_full[buf], (g / kPAcc) & 1, 84);
        tc_fence_after();
        uint32_t ra[32], rb[32];
        tmem_ld32(tmem_base + buf * 128 + t_addr, ra);
        tmem_ld32(tmem_base + buf * 128 + t_addr + 32, rb);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);                     // the accumulator is in registers: release it before the adds
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          acc[j] += __uint_as_float(ra[j]);
          acc[32 + j] += __uint_as_float(rb[j]);
        }
      }
      // row -> padded smem tile -> whole rows out (512-byte coalesced residual loads and stores)
      named_bar_sync(1, 256);                             // the previous tile's rows have been read out of tile_s
      {
        float* trow = tile_s + r * kPLdT + half * 64;
#pragma unroll
        for (int j = 0; j < 64; j += 4) *reinterpret_cast<float4*>(trow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      }
      named_bar_sync(1, 256);
      const int col = n0 + 4 * lane;
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      const float wsc = p.w_inv_scale * ((kTrain && p.out_scale_dev) ? __ldg(p.out_scale_dev) : 1.f), alpha = p.alpha;
      for (int rr = ew; rr < BM; rr += 8) {
        const int t = t0 + rr;
        if (t >= p.rows_per_seq) break;
        const size_t orow = static_cast<size_t>(seq) * p.rows_per_seq + t;
        const float4 a = *reinterpret_cast<const float4*>(tile_s + rr * kPLdT + 4 * lane);
        float4 v;
        v.x = alpha * act_apply<kAct>(fmaf(a.x, wsc, bb.x));
        v.y = alpha * act_apply<kAct>(fmaf(a.y, wsc, bb.y));
        v.z = alpha * act_apply<kAct>(fmaf(a.z, wsc, bb.z));
        v.w = alpha * act_apply<kAct>(fmaf(a.w, wsc, bb.w));
        if (p.residual) {
          const float4 rs = *reinterpret_cast<const float4*>(p.residual + orow * p.ldr + col);
          v.x += rs.x; v.y += rs.y; v.z += rs.z; v.w += rs.w;
        }
        if (kTrain && p.out_mask) {
          const float4 mk = __ldg(reinterpret_cast<const float4*>(p.out_mask + orow * p.ldo + col));
          v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f; v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
        }
        *reinterpret_cast<float4*>(p.out + orow * p.ldo + col) = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 7) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// TMA-fed form: A arrives as pre-split fp16 hi / lo planes [rows][K] (row-major, one 2-D tensor map each) — no register
// producers, so the per-k-block cost is the 12 MMAs plus the fold.  Used by the training products (train_ops.cu), whose
// operands are split by one elementwise pass that also applies the power-of-two gradient scale.  ncu on the register-fed
// persistent kernel (profiles/r02c_train_gemm_ncu.txt): tensor pipe 14-22 % active, the producer warps need ~2000+ clk per
// k-block against 768 clk of MMAs.
//   warp 0 lane 0 : TMA for A_hi, A_lo, W_hi, W_lo of every k-block (3-stage ring, 192 KB)
//   warp 1        : MMA issuer, four 128-column TMEM accumulators (fresh per k-block)
//   warps 4-11    : fold + epilogue (the tile goes out through a 64-row staging tile, two passes)
// One sequence, one tap.  kMask: out = mask > 0 ? v : 0 (ReLU backward in the epilogue); out_scale_dev as in the train variant.
constexpr int kTStages = 3;
constexpr int kTHalfRows = 64;
constexpr int kPlanesSmem = kTStages * kStageBytes + kTHalfRows * kPLdT * 4 + 1024;
constexpr int kTThreads = 12 * 32;

template <int kAct>
__global__ void __launch_bounds__(kTThreads, 1)
p32_gemm_planes_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                       const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                       const P32GemmParams p, const int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kTStages];
  __shared__ __align__(8) uint64_t empty_bar[kTStages];
  __shared__ __align__(8) uint64_t acc_full[kPAcc];
  __shared__ __align__(8) uint64_t acc_empty[kPAcc];
  __shared__ uint32_t tmem_base_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* tile_s = reinterpret_cast<float*>(smem + kTStages * kStageBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kTStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < kPAcc; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 256);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmAhi);
    tma_prefetch_desc(&tmAlo);
    tma_prefetch_desc(&tmWhi);
    tma_prefetch_desc(&tmWlo);
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int n_tiles_n = p.N / BN;
  const int KB = p.k_blocks;
  const int n_local = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const uint32_t G = static_cast<uint32_t>(n_local > 0 ? n_local : 0) * KB;

  if (warp == 0) {
    if (lane == 0) {
      for (uint32_t gi = 0; gi < G; ++gi) {
        const int s = gi % kTStages;
        mbar_wait(&empty_bar[s], ((gi / kTStages) & 1) ^ 1, 91);
        const int tile = blockIdx.x + (gi / KB) * gridDim.x, kb = gi % KB;
        const int t0 = (tile / n_tiles_n) * BM, n0 = (tile % n_tiles_n) * BN;
        uint8_t* st = smem + s * kStageBytes;
        mbar_arrive_expect_tx(&full_bar[s], 4 * kTileBytes);
        tma_load_2d(st, &tmAhi, &full_bar[s], kb * BK, t0);                 // rows beyond the tensor are zero-filled
        tma_load_2d(st + kTileBytes, &tmAlo, &full_bar[s], kb * BK, t0);
        tma_load_2d(st + 2 * kTileBytes, &tmWhi, &full_bar[s], kb * BK, n0);
        tma_load_2d(st + 3 * kTileBytes, &tmWlo, &full_bar[s], kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_f16(BM, BN, false);
    for (uint32_t g = 0; g < G; ++g) {
      const int s = g % kTStages, buf = g % kPAcc;
      mbar_wait(&full_bar[s], (g / kTStages) & 1, 92);
      mbar_wait(&acc_empty[buf], ((g / kPAcc) & 1) ^ 1, 93);
      tc_fence_after();
      const uint32_t sb = smem_u32(smem + s * kStageBytes);
      const uint64_t dAhi = smem_desc_sw128(sb), dAlo = smem_desc_sw128(sb + kTileBytes);
      const uint64_t dWhi = smem_desc_sw128(sb + 2 * kTileBytes), dWlo = smem_desc_sw128(sb + 3 * kTileBytes);
      const uint32_t d = tmem_base + buf * 128;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          umma_f16(d, dAlo + 2 * kk, dWhi + 2 * kk, idesc, kk > 0 ? 1u : 0u);
          umma_f16(d, dAhi + 2 * kk, dWlo + 2 * kk, idesc, 1u);
        }
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) umma_f16(d, dAhi + 2 * kk, dWhi + 2 * kk, idesc, 1u);
        umma_commit(&empty_bar[s]);
        umma_commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int wq = warp & 3, half = (warp - 4) >> 2;
    const int r = wq * 32 + lane;
    const uint32_t t_addr = (static_cast<uint32_t>(wq * 32) << 16) + half * 64;
    const int ew = warp - 4;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int t0 = (tile / n_tiles_n) * BM, n0 = (tile % n_tiles_n) * BN;
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int buf = g % kPAcc;
        mbar_wait(&acc_full[buf], (g / kPAcc) & 1, 94);
        tc_fence_after();
        uint32_t ra[32], rb[32];
        tmem_ld32(tmem_base + buf * 128 + t_addr, ra);
        tmem_ld32(tmem_base + buf * 128 + t_addr + 32, rb);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          acc[j] += __uint_as_float(ra[j]);
          acc[32 + j] += __uint_as_float(rb[j]);
        }
      }
      const int col = n0 + 4 * lane;
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      const float wsc = p.w_inv_scale * (p.out_scale_dev ? __ldg(p.out_scale_dev) : 1.f), alpha = p.alpha;
#pragma unroll 1
      for (int hp = 0; hp < 2; ++hp) {                  // rows hp * 64 .. + 64 through the staging tile
        named_bar_sync(1, 256);                         // the previous pass' rows have been read out
        if ((wq >> 1) == hp) {
          float* trow = tile_s + (r - hp * kTHalfRows) * kPLdT + half * 64;
#pragma unroll
          for (int j = 0; j < 64; j += 4) *reinterpret_cast<float4*>(trow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
        named_bar_sync(1, 256);
        for (int rr = ew; rr < kTHalfRows; rr += 8) {
          const int t = t0 + hp * kTHalfRows + rr;
          if (t >= p.rows_per_seq) break;
          const size_t orow = static_cast<size_t>(t);
          const float4 a = *reinterpret_cast<const float4*>(tile_s + rr * kPLdT + 4 * lane);
          float4 v;
          v.x = alpha * act_apply<kAct>(fmaf(a.x, wsc, bb.x));
          v.y = alpha * act_apply<kAct>(fmaf(a.y, wsc, bb.y));
          v.z = alpha * act_apply<kAct>(fmaf(a.z, wsc, bb.z));
          v.w = alpha * act_apply<kAct>(fmaf(a.w, wsc, bb.w));
          if (p.residual) {
            const float4 rs = *reinterpret_cast<const float4*>(p.residual + orow * p.ldr + col);
            v.x += rs.x; v.y += rs.y; v.z += rs.z; v.w += rs.w;
          }
          if (p.out_mask) {
            const float4 mk = __ldg(reinterpret_cast<const float4*>(p.out_mask + orow * p.ldo + col));
            v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f; v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
          }
          *reinterpret_cast<float4*>(p.out + orow * p.ldo + col) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// A-stationary persistent form for K <= 256, one tap, N >= 256 (QKV(G) projections, FFN up-projections, speaker-attention
// in-projection: most launches of the parity path).  ncu on the persistent kernel above (r02_final_p32persist.txt): 844 M
// warp-instructions for one decoder FFN up-projection, i.e. ~900 issue cycles per k-block against 768 clk of MMAs — the A
// tile is re-fetched and re-split for every one of the N / 128 column tiles.  Here a work item is a 128-row tile: its
// hi / lo planes for the whole K (up to 128 KB) are built ONCE and stay in shared memory while the CTA walks all column
// tiles, streaming only the weights (TMA, 3-stage ring).  Per-k-block barriers let the next row tile's planes be built
// as soon as the last column tile has consumed the corresponding k-block.
//   warps 0-5 : A producers;  warp 6 : W loader (TMA);  warp 7 : MMA issuer;  warps 8-15 : fold + epilogue (registers ->
//   global, 64 consecutive floats per thread).
constexpr int kSWStages = 3;
constexpr int kSMaxKB = 4;
constexpr int kSWStageBytes = 2 * kTileBytes;                                   // W_hi | W_lo
constexpr int kAstatSmem = kSMaxKB * 2 * kTileBytes + kSWStages * kSWStageBytes + 1024;   // 128 KB + 96 KB
constexpr int kSProd = 6 * 32;

template <int kAct>
__global__ void __launch_bounds__(kPThreads, 1)
p32_gemm_astat_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                      const P32GemmParams p, const int n_m_tiles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kSMaxKB];
  __shared__ __align__(8) uint64_t a_empty[kSMaxKB];
  __shared__ __align__(8) uint64_t w_full[kSWStages];
  __shared__ __align__(8) uint64_t w_empty[kSWStages];
  __shared__ __align__(8) uint64_t acc_full[kPAcc];
  __shared__ __align__(8) uint64_t acc_empty[kPAcc];
  __shared__ uint32_t tmem_base_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smemW = smem + kSMaxKB * 2 * kTileBytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int k = 0; k < kSMaxKB; ++k) {
      mbar_init(&a_full[k], kSProd);
      mbar_init(&a_empty[k], 1);
    }
    for (int s = 0; s < kSWStages; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int b = 0; b < kPAcc; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 256);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmWhi);
    tma_prefetch_desc(&tmWlo);
  }
  if (warp == 7) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int n_tiles_n = p.N / BN;
  const int tiles_per_seq = (p.rows_per_seq + BM - 1) / BM;
  const int KB = p.k_blocks;

  if (warp < 6) {
    // ------------------------------------------------------------ A producers: planes of one row tile, once
    float4 pre[12];
    auto fetch = [&](int m_tile, int kb) {
      const int seq = m_tile / tiles_per_seq, t0 = (m_tile % tiles_per_seq) * BM;
      const float* a_seq = p.A + static_cast<size_t>(seq) * p.a_seq_rows * p.lda;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int q = tid + kSProd * i;                     // 1024 chunks of 8 floats over 192 threads
        const int r = q >> 3, c = q & 7;
        const int t = t0 + r;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (q < 1024 && t < p.rows_per_seq) {
          const float4* src = reinterpret_cast<const float4*>(a_seq + static_cast<size_t>(t) * p.lda + kb * BK + c * 8);
          v0 = __ldg(src);
          v1 = __ldg(src + 1);
        }
        pre[2 * i] = v0;
        pre[2 * i + 1] = v1;
      }
    };
    int m_tile = blockIdx.x;
    if (m_tile < n_m_tiles) fetch(m_tile, 0);
    for (uint32_t j = 0; m_tile < n_m_tiles; m_tile += gridDim.x, ++j) {
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&a_empty[kb], (j & 1) ^ 1, 91);           // the previous row tile's last column tile has consumed it
        uint8_t* sAhi = smem + kb * 2 * kTileBytes;
        uint8_t* sAlo = sAhi + kTileBytes;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int q = tid + kSProd * i;
          if (q < 1024) {
            const int r = q >> 3, c = q & 7;
            uint4 hi, lo;
            split8(pre[2 * i], pre[2 * i + 1], hi, lo);
            const uint32_t off = sw128_offset(r, c);
            *reinterpret_cast<uint4*>(sAhi + off) = hi;
            *reinterpret_cast<uint4*>(sAlo + off) = lo;
          }
        }
        if (kb + 1 < KB) fetch(m_tile, kb + 1);
        else if (m_tile + static_cast<int>(gridDim.x) < n_m_tiles) fetch(m_tile + gridDim.x, 0);
        fence_proxy_async_smem();
        mbar_arrive(&a_full[kb]);
      }
    }
  } else if (warp == 6) {
    // ------------------------------------------------------------ W loader
    uint32_t g = 0;
    for (int m_tile = blockIdx.x; m_tile < n_m_tiles; m_tile += gridDim.x) {
      for (int nt = 0; nt < n_tiles_n; ++nt) {
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int s = g % kSWStages;
          mbar_wait(&w_empty[s], ((g / kSWStages) & 1) ^ 1, 92);
          if (elect_one()) {
            uint8_t* sW = smemW + s * kSWStageBytes;
            mbar_arrive_expect_tx(&w_full[s], kSWStageBytes);
            tma_load_2d(sW, &tmWhi, &w_full[s], kb * BK, nt * BN);
            tma_load_2d(sW + kTileBytes, &tmWlo, &w_full[s], kb * BK, nt * BN);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 7) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(BM, BN, false);
    uint32_t g = 0, j = 0;
    for (int m_tile = blockIdx.x; m_tile < n_m_tiles; m_tile += gridDim.x, ++j) {
      for (int nt = 0; nt < n_tiles_n; ++nt) {
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int s = g % kSWStages, buf = g % kPAcc;
          mbar_wait(&w_full[s], (g / kSWStages) & 1, 93);
          if (nt == 0) mbar_wait(&a_full[kb], j & 1, 94);
          mbar_wait(&acc_empty[buf], ((g / kPAcc) & 1) ^ 1, 95);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + kb * 2 * kTileBytes);
          const uint32_t sw = smem_u32(smemW + s * kSWStageBytes);
          const uint64_t dAhi = smem_desc_sw128(sa), dAlo = smem_desc_sw128(sa + kTileBytes);
          const uint64_t dWhi = smem_desc_sw128(sw), dWlo = smem_desc_sw128(sw + kTileBytes);
          const uint32_t d = tmem_base + buf * 128;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              umma_f16(d, dAlo + 2 * kk, dWhi + 2 * kk, idesc, kk > 0 ? 1u : 0u);
              umma_f16(d, dAhi + 2 * kk, dWlo + 2 * kk, idesc, 1u);
            }
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) umma_f16(d, dAhi + 2 * kk, dWhi + 2 * kk, idesc, 1u);
            umma_commit(&w_empty[s]);
            umma_commit(&acc_full[buf]);
            if (nt == n_tiles_n - 1) umma_commit(&a_empty[kb]);   // this k-block's planes may be rebuilt for the next row tile
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------ fold + epilogue warps 8-15
    const int wq = warp & 3, half = (warp - 8) >> 2;
    const int r = wq * 32 + lane;
    const uint32_t t_addr = (static_cast<uint32_t>(wq * 32) << 16) + half * 64;
    uint32_t g = 0;
    for (int m_tile = blockIdx.x; m_tile < n_m_tiles; m_tile += gridDim.x) {
      const int seq = m_tile / tiles_per_seq, t0 = (m_tile % tiles_per_seq) * BM;
      const int t = t0 + r;
      const size_t orow = static_cast<size_t>(seq) * p.rows_per_seq + t;
      for (int nt = 0; nt < n_tiles_n; ++nt) {
        float acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.f;
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int buf = g % kPAcc;
          mbar_wait(&acc_full[buf], (g / kPAcc) & 1, 96);
          tc_fence_after();
          uint32_t ra[32], rb[32];
          tmem_ld32(tmem_base + buf * 128 + t_addr, ra);
          tmem_ld32(tmem_base + buf * 128 + t_addr + 32, rb);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&acc_empty[buf]);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            acc[i] += __uint_as_float(ra[i]);
            acc[32 + i] += __uint_as_float(rb[i]);
          }
        }
        if (t < p.rows_per_seq) {
          const int col0 = nt * BN + half * 64;
          float* op = p.out + orow * p.ldo + col0;
          const float* rp = p.residual ? p.residual + orow * p.ldr + col0 : nullptr;
          const float wsc = p.w_inv_scale, alpha = p.alpha;
#pragma unroll
          for (int i = 0; i < 64; i += 4) {
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
            float4 v;
            v.x = alpha * act_apply<kAct>(fmaf(acc[i + 0], wsc, bb.x));
            v.y = alpha * act_apply<kAct>(fmaf(acc[i + 1], wsc, bb.y));
            v.z = alpha * act_apply<kAct>(fmaf(acc[i + 2], wsc, bb.z));
            v.w = alpha * act_apply<kAct>(fmaf(acc[i + 3], wsc, bb.w));
            if (rp) {
              const float4 rs = *reinterpret_cast<const float4*>(rp + i);
              v.x += rs.x; v.y += rs.y; v.z += rs.z; v.w += rs.w;
            }
            *reinterpret_cast<float4*>(op + i) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 7) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// Small-row form of the same product (see p32.cuh): at most 16 output rows.  CTA = 8 warps x 2 output columns; K is
// walked in blocks of 512 whose activations (R x 512 fp32) are staged in shared memory once per CTA.
constexpr int kRvMaxRows = 16, kRvKB = 512, kRvCols = 2, kRvWarps = 8;
__device__ __forceinline__ void warp_ln8(float (&v)[8], const float* g, const float* b, int lane, float eps);
__device__ __forceinline__ void rowvec_row_epilogue(const P32GemmParams& p, int R);

template <int kAct>
__global__ void __launch_bounds__(kRvWarps * 32)
p32_rowvec_kernel(const __half* __restrict__ whi, const __half* __restrict__ wlo, const P32GemmParams p) {
  __shared__ __align__(16) float a_s[kRvMaxRows][kRvKB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.n_seq * p.rows_per_seq;
  const int K = p.k_blocks * 64, Ktot = p.taps * K;
  const int n0 = (blockIdx.x * kRvWarps + warp) * kRvCols;
  const int a_off = p.a_row_offset + (p.a_row_offset_dev ? *p.a_row_offset_dev : 0);
  float acc[kRvCols][kRvMaxRows];
#pragma unroll
  for (int c = 0; c < kRvCols; ++c)
#pragma unroll
    for (int r = 0; r < kRvMaxRows; ++r) acc[c][r] = 0.f;

  for (int kb0 = 0; kb0 < Ktot; kb0 += kRvKB) {
    __syncthreads();
    // stage A[r][kb0 .. kb0 + 512) (taps: output row t of a sequence reads A row t + tap + tap_shift + a_off, zero outside)
    for (int i = tid; i < R * (kRvKB / 4); i += kRvWarps * 32) {
      const int r = i / (kRvKB / 4), k = kb0 + (i - r * (kRvKB / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < Ktot) {
        const int tap = k / K, c = k - tap * K;
        const int seq = r / p.rows_per_seq, t = r - seq * p.rows_per_seq;
        const int arow = t + tap + p.tap_shift + a_off;
        if (arow >= 0 && arow < p.a_seq_rows)
          v = __ldg(reinterpret_cast<const float4*>(p.A + (static_cast<size_t>(seq) * p.a_seq_rows + arow) * p.lda + c));
      }
      *reinterpret_cast<float4*>(&a_s[r][k - kb0]) = v;
    }
    __syncthreads();
    if (n0 < p.N) {
#pragma unroll
      for (int j = 0; j < kRvKB / 256; ++j) {
        const int kl = (lane + 32 * j) * 8, k = kb0 + kl;
        if (k < Ktot) {
          const int tap = k / K, c = k - tap * K;
          float w[kRvCols][8];
#pragma unroll
          for (int cc = 0; cc < kRvCols; ++cc) {
            const size_t off = (static_cast<size_t>(tap) * p.N + n0 + cc) * K + c;
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(whi + off));
            const __half2* hh = reinterpret_cast<const __half2*>(&h);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = __half22float2(hh[q]);
              w[cc][2 * q] = f.x;
              w[cc][2 * q + 1] = f.y;
            }
            if (wlo) {
              const uint4 l = __ldg(reinterpret_cast<const uint4*>(wlo + off));
              const __half2* ll = reinterpret_cast<const __half2*>(&l);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 f = __half22float2(ll[q]);
                w[cc][2 * q] += f.x;          // hi + lo: exact in fp32 (two non-overlapping 11-bit mantissas)
                w[cc][2 * q + 1] += f.y;
              }
            }
          }
#pragma unroll
          for (int r = 0; r < kRvMaxRows; ++r) {
            if (r < R) {
              const float4 a0 = *reinterpret_cast<const float4*>(&a_s[r][kl]);
              const float4 a1 = *reinterpret_cast<const float4*>(&a_s[r][kl + 4]);
#pragma unroll
              for (int cc = 0; cc < kRvCols; ++cc) {
                float s = acc[cc][r];
                s = fmaf(w[cc][0], a0.x, s); s = fmaf(w[cc][1], a0.y, s); s = fmaf(w[cc][2], a0.z, s); s = fmaf(w[cc][3], a0.w, s);
                s = fmaf(w[cc][4], a1.x, s); s = fmaf(w[cc][5], a1.y, s); s = fmaf(w[cc][6], a1.z, s); s = fmaf(w[cc][7], a1.w, s);
                acc[cc][r] = s;
              }
            }
          }
        }
      }
    }
  }
  if (n0 < p.N) {
#pragma unroll
  for (int cc = 0; cc < kRvCols; ++cc)
#pragma unroll
    for (int r = 0; r < kRvMaxRows; ++r) {
      if (r < R) {
        float v = acc[cc][r];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        acc[cc][r] = v;
      }
    }
  // lane l < R * kRvCols finishes output (row l / kRvCols, column n0 + l % kRvCols)
#pragma unroll
  for (int cc = 0; cc < kRvCols; ++cc)
#pragma unroll
    for (int r = 0; r < kRvMaxRows; ++r) {
      if (r < R && lane == ((r * kRvCols + cc) & 31)) {
        const int n = n0 + cc;
        float x = fmaf(acc[cc][r], p.w_inv_scale, p.bias ? __ldg(p.bias + n) : 0.f);
        x = p.alpha * act_apply<kAct>(x);
        if (p.residual) x += p.residual[static_cast<size_t>(r) * p.ldr + n];
        p.out[static_cast<size_t>(r) * p.ldo + n] = x;
      }
    }
  }
  if (p.row_epi_counter) rowvec_row_epilogue(p, R);
}

// Fused row epilogue of the row-vector product (see P32GemmParams::row_epi_counter): runs in the last CTA to finish.
__device__ __forceinline__ void rowvec_row_epilogue(const P32GemmParams& p, int R) {
  __shared__ bool is_last;
  __threadfence();                                   // this CTA's output columns are visible device-wide
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(p.row_epi_counter, 1u);
    is_last = prev == gridDim.x - 1;
    if (is_last) *p.row_epi_counter = 0u;            // ready for the next launch on this stream
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += kRvWarps) {
    const float4* xp = reinterpret_cast<const float4*>(p.out + static_cast<size_t>(r) * p.ldo);
    const float4 a = __ldcg(xp + lane), b = __ldcg(xp + 32 + lane);      // L2 (other CTAs wrote these): bypass L1
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (p.l2norm) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(v[i], v[i], ss);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
      const float inv = 1.f / sqrtf(ss);
      float4* op = reinterpret_cast<float4*>(p.out + static_cast<size_t>(r) * p.ldo);
      op[lane] = make_float4(v[0] * inv, v[1] * inv, v[2] * inv, v[3] * inv);
      op[32 + lane] = make_float4(v[4] * inv, v[5] * inv, v[6] * inv, v[7] * inv);
      continue;
    }
    if (p.ln_g1) warp_ln8(v, p.ln_g1, p.ln_b1, lane, p.ln_eps);
    if (p.ln_out1) {
      float4* op = reinterpret_cast<float4*>(p.ln_out1 + static_cast<size_t>(r) * 256);
      op[lane] = make_float4(v[0], v[1], v[2], v[3]);
      op[32 + lane] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (p.ln_out2) {
      warp_ln8(v, p.ln_g2, p.ln_b2, lane, p.ln_eps);
      float4* op = reinterpret_cast<float4*>(p.ln_out2 + static_cast<size_t>(r) * 256);
      op[lane] = make_float4(v[0], v[1], v[2], v[3]);
      op[32 + lane] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

// Streaming attention step in fp32 (see elementwise.cu: step_attn_kernel for the fp16-cache form).
__global__ void __launch_bounds__(128)
p32_step_attn_kernel(const float* __restrict__ qkv, float* __restrict__ kcache, float* __restrict__ vcache, int cap,
                     int pos_arg, const int* __restrict__ pos_dev, float scale, float* __restrict__ out) {
  const int n = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int pos = pos_dev ? *pos_dev : pos_arg;
  __shared__ float q_s[64];
  __shared__ float red_m[128], red_l[128];
  __shared__ float red_o[4][64];
  const float* row = qkv + static_cast<size_t>(n) * 768;
  float* kc = kcache + (static_cast<size_t>(n) * cap) * 256 + h * 64;
  float* vc = vcache + (static_cast<size_t>(n) * cap) * 256 + h * 64;
  if (tid < 64) {
    q_s[tid] = row[h * 64 + tid] * scale * 1.4426950408889634f;
    kc[static_cast<size_t>(pos) * 256 + tid] = row[256 + h * 64 + tid];
    vc[static_cast<size_t>(pos) * 256 + tid] = row[512 + h * 64 + tid];
  }
  __syncthreads();
  float m = -INFINITY, l = 0.f, o[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] = 0.f;
  for (int j = tid; j <= pos; j += 128) {
    const float4* kp = reinterpret_cast<const float4*>(kc + static_cast<size_t>(j) * 256);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 f = kp[i];
      s = fmaf(q_s[4 * i], f.x, s); s = fmaf(q_s[4 * i + 1], f.y, s);
      s = fmaf(q_s[4 * i + 2], f.z, s); s = fmaf(q_s[4 * i + 3], f.w, s);
    }
    const float m_new = fmaxf(m, s);
    const float a = exp2f(m - m_new), pj = exp2f(s - m_new);
    l = l * a + pj;
    const float4* vp = reinterpret_cast<const float4*>(vc + static_cast<size_t>(j) * 256);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 f = vp[i];
      o[4 * i] = fmaf(pj, f.x, o[4 * i] * a); o[4 * i + 1] = fmaf(pj, f.y, o[4 * i + 1] * a);
      o[4 * i + 2] = fmaf(pj, f.z, o[4 * i + 2] * a); o[4 * i + 3] = fmaf(pj, f.w, o[4 * i + 3] * a);
    }
    m = m_new;
  }
  red_m[tid] = m;
  __syncthreads();
  float gm = -INFINITY;
  for (int i = 0; i < 128; ++i) gm = fmaxf(gm, red_m[i]);
  const float w = (m == -INFINITY) ? 0.f : exp2f(m - gm);
  red_l[tid] = l * w;
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    float v = o[i] * w;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((tid & 31) == 0) red_o[tid >> 5][i] = v;
  }
  __syncthreads();
  if (tid < 64) {
    float lt = 0.f;
    for (int i = 0; i < 128; ++i) lt += red_l[i];
    const float v = red_o[0][tid] + red_o[1][tid] + red_o[2][tid] + red_o[3][tid];
    out[static_cast<size_t>(n) * 256 + h * 64 + tid] = v / lt;
  }
}

// =====================================================================================================================
// LayerNorm over rows of 256 fp32 (biased variance, two-pass in registers): one warp per row, lane owns columns
// [4 lane, 4 lane + 4) and [128 + 4 lane, ...).  y1 = g1 ? LN(x) : x -> out1;  out2 = LN(y1; g2, b2).
__device__ __forceinline__ void warp_ln8(float (&v)[8], const float* g, const float* b, int lane, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s * (1.f / 256.f);
  float m2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float d = v[i] - mean;
    m2 = fmaf(d, d, m2);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, off);
  const float rstd = 1.f / sqrtf(m2 * (1.f / 256.f) + eps);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + lane), g1 = __ldg(reinterpret_cast<const float4*>(g + 128) + lane);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + lane), b1 = __ldg(reinterpret_cast<const float4*>(b + 128) + lane);
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * gg[i] + bb[i];
}

__global__ void __launch_bounds__(256)
p32_layernorm_kernel(const float* __restrict__ x, int rows, const float* __restrict__ g1, const float* __restrict__ b1,
                     float* __restrict__ out1, const float* __restrict__ g2, const float* __restrict__ b2,
                     float* __restrict__ out2, float eps, const int* __restrict__ seq_len, int rows_per_seq) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xp = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * 256);
  const float4 a = xp[lane], b = xp[32 + lane];
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  bool zero = false;
  if (seq_len) {
    const int sq = row / rows_per_seq;
    zero = (row - sq * rows_per_seq) >= seq_len[sq];
  }
  if (g1) warp_ln8(v, g1, b1, lane, eps);
  if (zero) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
  if (out1) {
    float4* op = reinterpret_cast<float4*>(out1 + static_cast<size_t>(row) * 256);
    op[lane] = make_float4(v[0], v[1], v[2], v[3]);
    op[32 + lane] = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (out2) {
    warp_ln8(v, g2, b2, lane, eps);
    if (zero) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    float4* op = reinterpret_cast<float4*>(out2 + static_cast<size_t>(row) * 256);
    op[lane] = make_float4(v[0], v[1], v[2], v[3]);
    op[32 + lane] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// optional per-channel affine (BatchNorm folded, FS:model:165-166: rows t >= len take pad_value BEFORE the affine)
__global__ void p32_pad_input_kernel(const float* __restrict__ x, const int* __restrict__ cu, int Tmax, int Din,
                                     int Kpad, float* __restrict__ out, const float* __restrict__ sc,
                                     const float* __restrict__ sh, float pad_value) {
  const int row = blockIdx.x;
  const int b = row / Tmax, t = row - b * Tmax;
  const int start = cu[b], len = cu[b + 1] - start;
  const float* src = (t < len) ? x + static_cast<size_t>(start + t) * Din : nullptr;
  float* dst = out + static_cast<size_t>(row) * Kpad;
  for (int i = threadIdx.x; i < Kpad; i += blockDim.x) {
    float v = 0.f;
    if (i < Din) {
      v = src ? __ldg(src + i) : pad_value;
      if (sc) v = fmaf(v, sc[i], sh[i]);
    }
    dst[i] = v;
  }
}

// GLU over channel halves (conformer/activation.py:40-42): out = a * sigmoid(b)
__global__ void __launch_bounds__(256)
p32_glu_kernel(const float* __restrict__ h, size_t n4, float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // float4 index into [rows][64 float4]
  if (i >= n4) return;
  const size_t row = i >> 6, c = i & 63;
  const float4 a = reinterpret_cast<const float4*>(h)[row * 128 + c];
  const float4 g = reinterpret_cast<const float4*>(h)[row * 128 + 64 + c];
  float4 o;
  o.x = a.x / (1.f + expf(-g.x));
  o.y = a.y / (1.f + expf(-g.y));
  o.z = a.z / (1.f + expf(-g.z));
  o.w = a.w / (1.f + expf(-g.w));
  reinterpret_cast<float4*>(out)[i] = o;
}

// Causal depthwise Conv1d (left zero padding K-1, conformer/convolution.py:144,65-68) -> BatchNorm1d (eval, folded)
// -> swish.  Thread = (channel, 8 output frames); one-step form reads and slides the (K-1)-frame cache.
__global__ void __launch_bounds__(256)
p32_dwconv_bn_swish_kernel(const float* __restrict__ u, const float* __restrict__ w, const float* __restrict__ sc,
                           const float* __restrict__ sh, int T, int K, float* __restrict__ hist,
                           float* __restrict__ out) {
  const int n = blockIdx.y, c = threadIdx.x;
  const int t0 = blockIdx.x * 8;
  float wk[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) wk[k] = k < K ? w[c * K + k] : 0.f;
  const float* up = u + static_cast<size_t>(n) * T * 256 + c;
  float* hp = hist ? hist + static_cast<size_t>(n) * (K - 1) * 256 + c : nullptr;
  const float s0 = sc[c], h0 = sh[c];
  for (int t = t0; t < min(t0 + 8, T); ++t) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k < K) {
        const int tt = t - (K - 1) + k;
        float v = 0.f;
        if (tt >= 0) v = up[static_cast<size_t>(tt) * 256];
        else if (hp) v = hp[static_cast<size_t>(tt + K - 1) * 256];
        acc = fmaf(wk[k], v, acc);
      }
    }
    acc = fmaf(acc, s0, h0);
    out[(static_cast<size_t>(n) * T + t) * 256 + c] = acc / (1.f + expf(-acc));
  }
  if (hp && T == 1) {   // slide the cache: drop the oldest frame, append u[0]
    float prev[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) prev[k] = (k < K - 1) ? hp[static_cast<size_t>(k) * 256] : 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k < K - 2) hp[static_cast<size_t>(k) * 256] = prev[k + 1];
    }
    if (K >= 2) hp[static_cast<size_t>(K - 2) * 256] = up[0];
  }
}

__global__ void __launch_bounds__(256)
p32_l2norm_kernel(float* __restrict__ x, int rows) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4* xp = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * 256);
  float4 a = xp[lane], b = xp[32 + lane];
  float ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const float inv = 1.f / sqrtf(ss);      // no eps, as the reference (x / norm): a zero row gives NaN there too
  a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
  b.x *= inv; b.y *= inv; b.z *= inv; b.w *= inv;
  xp[lane] = a;
  xp[32 + lane] = b;
}

__global__ void __launch_bounds__(256)
p32_convert_kernel(const float* __restrict__ y, const float* __restrict__ pe_proj, size_t n4, int S,
                   float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // float4 index into [rows][S][64]
  if (i >= n4) return;
  const size_t c = i & 63, rs = i >> 6;
  const size_t row = rs / S, s = rs - row * S;
  const float4 a = reinterpret_cast<const float4*>(y)[row * 64 + c];
  const float4 b = __ldg(reinterpret_cast<const float4*>(pe_proj) + s * 64 + c);
  reinterpret_cast<float4*>(out)[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// Speaker-axis attention (merge_retnet_layer.py:244-249 -> nn.MultiheadAttention, no mask): thread = (frame, slot, head).
// The K / V rows of the block's frames are staged in shared memory with coalesced 16-byte loads (per head 64 + 4 floats:
// the four heads of a row land in different bank groups); the first version read them straight from global memory and
// ran at 1.4 TB/s (ncu r02_ls_p32_kernels.txt: 0.91 ms per decoder layer at B=16, T=2000, S=10).
constexpr int kMaxS = 16;
constexpr int kSpkHead = 68;                      // floats per (row, head) in shared memory
constexpr int kSpkRow = 2 * 4 * kSpkHead;         // K then V

__global__ void __launch_bounds__(128)
p32_spk_attn_kernel(const float* __restrict__ qkv, float* __restrict__ out, int n_frames, int S, float scale,
                    int frames_per_block) {
  extern __shared__ __align__(16) float kv_s[];
  const int f0 = blockIdx.x * frames_per_block;
  const int nf = min(frames_per_block, n_frames - f0);
  const int rows = nf * S;
  for (int i = threadIdx.x; i < rows * 128; i += blockDim.x) {
    const int rr = i >> 7, c4 = i & 127;                       // 128 float4 per row: K (64) then V (64)
    const float4 v = __ldg(reinterpret_cast<const float4*>(qkv + (static_cast<size_t>(f0) * S + rr) * 768 + 256) + c4);
    const int part = c4 >> 6, h = (c4 >> 4) & 3, d4 = c4 & 15;
    *reinterpret_cast<float4*>(kv_s + rr * kSpkRow + (part * 4 + h) * kSpkHead + 4 * d4) = v;
  }
  __syncthreads();
  const int tpf = 4 * S;
  const int f = threadIdx.x / tpf;
  if (f >= nf) return;
  const int rem = threadIdx.x - f * tpf;
  const int a = rem >> 2, h = rem & 3;                         // heads fastest: 4 threads read 1 KB of one q row
  const size_t row = (static_cast<size_t>(f0) + f) * S + a;
  float q[64];
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + row * 768 + h * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 t = __ldg(qp + i);
      q[4 * i] = t.x * scale; q[4 * i + 1] = t.y * scale; q[4 * i + 2] = t.z * scale; q[4 * i + 3] = t.w * scale;
    }
  }
  float sc[kMaxS];
  float mx = -INFINITY;
#pragma unroll
  for (int b = 0; b < kMaxS; ++b) {
    sc[b] = -INFINITY;
    if (b < S) {
      const float4* kp = reinterpret_cast<const float4*>(kv_s + (f * S + b) * kSpkRow + h * kSpkHead);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 t = kp[i];
        d = fmaf(q[4 * i], t.x, d); d = fmaf(q[4 * i + 1], t.y, d);
        d = fmaf(q[4 * i + 2], t.z, d); d = fmaf(q[4 * i + 3], t.w, d);
      }
      sc[b] = d;
      mx = fmaxf(mx, d);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int b = 0; b < kMaxS; ++b) {
    if (b < S) {
      sc[b] = expf(sc[b] - mx);
      sum += sc[b];
    }
  }
  const float inv = 1.f / sum;
  float o[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] = 0.f;
#pragma unroll
  for (int b = 0; b < kMaxS; ++b) {
    if (b < S) {
      const float pw = sc[b] * inv;
      const float4* vp = reinterpret_cast<const float4*>(kv_s + (f * S + b) * kSpkRow + (4 + h) * kSpkHead);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 t = vp[i];
        o[4 * i] = fmaf(pw, t.x, o[4 * i]); o[4 * i + 1] = fmaf(pw, t.y, o[4 * i + 1]);
        o[4 * i + 2] = fmaf(pw, t.z, o[4 * i + 2]); o[4 * i + 3] = fmaf(pw, t.w, o[4 * i + 3]);
      }
    }
  }
  float4* op = reinterpret_cast<float4*>(out + row * 256 + h * 64);
#pragma unroll
  for (int i = 0; i < 16; ++i) op[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
}

// Logits head (LS:model:136-143): att_n = att / ||att||, y = emb . att_n.  One warp per frame.
__global__ void __launch_bounds__(256)
p32_head_kernel(const float* __restrict__ emb, const float* __restrict__ att, int n_frames, int S,
                float* __restrict__ logits, float* __restrict__ emb_out, float* __restrict__ att_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_frames) return;
  const float4* ep = reinterpret_cast<const float4*>(emb + static_cast<size_t>(warp) * 256);
  const float4 e0 = ep[lane], e1 = ep[32 + lane];
  if (emb_out) {
    float4* eo = reinterpret_cast<float4*>(emb_out + static_cast<size_t>(warp) * 256);
    eo[lane] = e0;
    eo[32 + lane] = e1;
  }
  for (int s = 0; s < S; ++s) {
    const size_t row = static_cast<size_t>(warp) * S + s;
    const float4* ap = reinterpret_cast<const float4*>(att + row * 256);
    const float4 a0 = ap[lane], a1 = ap[32 + lane];
    float ss = a0.x * a0.x + a0.y * a0.y + a0.z * a0.z + a0.w * a0.w + a1.x * a1.x + a1.y * a1.y + a1.z * a1.z + a1.w * a1.w;
    float dot = a0.x * e0.x + a0.y * e0.y + a0.z * e0.z + a0.w * e0.w + a1.x * e1.x + a1.y * e1.y + a1.z * e1.z + a1.w * e1.w;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, off);
      dot += __shfl_xor_sync(0xffffffffu, dot, off);
    }
    const float inv = 1.f / sqrtf(ss);
    if (lane == 0) logits[row] = dot * inv;
    if (att_out) {
      float4* ao = reinterpret_cast<float4*>(att_out + row * 256);
      ao[lane] = make_float4(a0.x * inv, a0.y * inv, a0.z * inv, a0.w * inv);
      ao[32 + lane] = make_float4(a1.x * inv, a1.y * inv, a1.z * inv, a1.w * inv);
    }
  }
}

// =====================================================================================================================
// Retention, pass 1: per (sequence n = (b, s), head h) the exclusive prefix KV_c = sum_{t in chunks < c} k_t^T v_t
// (modules/retention.py:167-180 without the 1/sqrt(C) weights, re-applied where used) and
// cross_scale_c = max(1, max_d sum_e |KV_c[e][d]| / sqrt(C)).  256 threads = 16 x 16, each owns a 4 x 4 block of KV.
__global__ void __launch_bounds__(256)
p32_ret_chunk_state_kernel(const float* __restrict__ qkvg, int S, int T, int chunk, int n_chunks,
                           float* __restrict__ state, float* __restrict__ cross_scale) {
  const int n = blockIdx.x >> 2, h = blockIdx.x & 3;
  const int b = n / S, s = n - b * S;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  __shared__ __align__(16) float ks[32][64];
  __shared__ __align__(16) float vs[32][64];
  __shared__ float colpart[16][64];     // per-ty partial column |.|-sums (fixed-order reduction: bit-reproducible)
  __shared__ float colsum[64];
  float kv[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) kv[i][j] = 0.f;
  const size_t row_stride = static_cast<size_t>(S) * 1024;
  const float* base = qkvg + (static_cast<size_t>(b) * T * S + s) * 1024 + h * 64;
  const float inv_sqrt_c = 1.f / sqrtf(static_cast<float>(chunk));
  for (int c = 0; c < n_chunks; ++c) {
    // state entering chunk c
    float* sp = state + ((static_cast<size_t>(n) * 4 + h) * n_chunks + c) * 4096;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(sp + (4 * ty + i) * 64 + 4 * tx) = make_float4(kv[i][0], kv[i][1], kv[i][2], kv[i][3]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      colpart[ty][4 * tx + j] = fabsf(kv[0][j]) + fabsf(kv[1][j]) + fabsf(kv[2][j]) + fabsf(kv[3][j]);
    __syncthreads();
    if (tid < 64) {
      float a = 0.f;
#pragma unroll
      for (int y = 0; y < 16; ++y) a += colpart[y][tid];
      colsum[tid] = a;
    }
    __syncthreads();
    if (tid == 0) {
      float mx = 0.f;
      for (int d = 0; d < 64; ++d) mx = fmaxf(mx, colsum[d]);
      cross_scale[(static_cast<size_t>(n) * 4 + h) * n_chunks + c] = fmaxf(1.f, mx * inv_sqrt_c);
    }
    if (c == n_chunks - 1) break;
    for (int r0 = 0; r0 < chunk; r0 += 32) {
      __syncthreads();
      // 32 rows x (64 k + 64 v) floats = 1024 float4: 4 per thread
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int q = tid + 256 * i;
        const int rr = q >> 5, part = (q >> 4) & 1, c4 = q & 15;
        const int t = c * chunk + r0 + rr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + rr < chunk) v = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(t) * row_stride + 256 + part * 256) + c4);
        *reinterpret_cast<float4*>(part ? &vs[rr][4 * c4] : &ks[rr][4 * c4]) = v;
      }
      __syncthreads();
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) {
        const float4 kk = *reinterpret_cast<const float4*>(&ks[rr][4 * ty]);
        const float4 vv = *reinterpret_cast<const float4*>(&vs[rr][4 * tx]);
        const float ke[4] = {kk.x, kk.y, kk.z, kk.w};
        const float vd[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) kv[i][j] = fmaf(ke[i], vd[j], kv[i][j]);
      }
    }
  }
}

// Retention, pass 2 (modules/retention.py:146-194,222-224 restated in retention.cu's header):
//   O_j = sum_{i<=j, same chunk} (q_j.k_i) v_i + q_j KV_c;   inner_j = max(1, sum_{i<=j} |q_j.k_i| / sqrt(j+1))
//   ret_j = O_j / (sqrt(j+1) max(inner_j, cross_c));   out_j = swish(g_j) * LayerNorm_64(ret_j; eps 1e-6)
// CTA = 64 query rows of one (sequence, head, chunk); 256 threads = 16 (ty: rows 4ty..) x 16 (tx); S tile columns are
// owned strided (tx + 16 jj) so that the row-major K tile is read without bank conflicts, O columns contiguous (4 tx..).
constexpr int kLD = 68;
constexpr int kRetSmem = 4 * 64 * kLD * 4;

__global__ void __launch_bounds__(256, 2)
p32_retention_kernel(const float* __restrict__ qkvg, const float* __restrict__ state,
                     const float* __restrict__ cross_scale, int S, int T, int chunk, int n_chunks,
                     float* __restrict__ out) {
  extern __shared__ __align__(16) float rsm[];
  float* Qs = rsm;
  float* Ks = Qs + 64 * kLD;
  float* Vs = Ks + 64 * kLD;
  float* Ss = Vs + 64 * kLD;
  const int n_qt = (chunk + 63) / 64;
  const int qt = n_qt - 1 - static_cast<int>(blockIdx.x);     // heaviest tiles first
  const int h = blockIdx.y;
  const int c = blockIdx.z % n_chunks, n = blockIdx.z / n_chunks;
  const int b = n / S, s = n - b * S;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const size_t row_stride = static_cast<size_t>(S) * 1024;
  const float* base = qkvg + ((static_cast<size_t>(b) * T + static_cast<size_t>(c) * chunk) * S + s) * 1024 + h * 64;
  const int q0 = qt * 64;

  auto load_tile = [&](float* dst, int col_off, int r0) {      // rows r0.. of the chunk, 64 floats at col_off
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = tid + 256 * i;
      const int rr = q >> 4, c4 = q & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + rr < chunk) v = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(r0 + rr) * row_stride + col_off) + c4);
      *reinterpret_cast<float4*>(dst + rr * kLD + 4 * c4) = v;
    }
  };
  // O[4ty+i][4tx+j] += sum_k L[4ty+i][k] * R[k][4tx+j]   (L, R: 64 x 64 tiles with leading dimension kLD)
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  auto mac_lr = [&](const float* L, const float* R) {
#pragma unroll 4
    for (int k4 = 0; k4 < 64; k4 += 4) {
      float l[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(L + (4 * ty + i) * kLD + k4);
        l[i][0] = t.x; l[i][1] = t.y; l[i][2] = t.z; l[i][3] = t.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 t = *reinterpret_cast<const float4*>(R + (k4 + k) * kLD + 4 * tx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o[i][0] = fmaf(l[i][k], t.x, o[i][0]);
          o[i][1] = fmaf(l[i][k], t.y, o[i][1]);
          o[i][2] = fmaf(l[i][k], t.z, o[i][2]);
          o[i][3] = fmaf(l[i][k], t.w, o[i][3]);
        }
      }
    }
  };

  load_tile(Qs, 0, q0);
  float asum[4] = {0.f, 0.f, 0.f, 0.f};
  if (c > 0) {   // cross-chunk term: O += Q KV_c
    const float* sp = state + ((static_cast<size_t>(n) * 4 + h) * n_chunks + c) * 4096;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = tid + 256 * i;
      const int rr = q >> 4, c4 = q & 15;
      *reinterpret_cast<float4*>(Vs + rr * kLD + 4 * c4) = __ldg(reinterpret_cast<const float4*>(sp + rr * 64) + c4);
    }
    __syncthreads();
    mac_lr(Qs, Vs);
  }
  for (int kt = 0; kt <= qt; ++kt) {
    __syncthreads();                       // previous tile's readers of Ks / Vs / Ss are done (and Qs is loaded)
    load_tile(Ks, 256, kt * 64);
    load_tile(Vs, 512, kt * 64);
    __syncthreads();
    // S[4ty+i][tx+16jj] = q_row . k_col
    float sacc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[i][j] = 0.f;
#pragma unroll 4
    for (int d4 = 0; d4 < 64; d4 += 4) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(Qs + (4 * ty + i) * kLD + d4);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * kLD + d4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sacc[i][j] = fmaf(qv[i].x, kv[j].x, sacc[i][j]);
          sacc[i][j] = fmaf(qv[i].y, kv[j].y, sacc[i][j]);
          sacc[i][j] = fmaf(qv[i].z, kv[j].z, sacc[i][j]);
          sacc[i][j] = fmaf(qv[i].w, kv[j].w, sacc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int jr = q0 + 4 * ty + i;                 // query index inside the chunk
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ic = kt * 64 + tx + 16 * j;         // key index inside the chunk
        const float v = (ic <= jr) ? sacc[i][j] : 0.f;
        asum[i] += fabsf(v);
        Ss[(4 * ty + i) * kLD + tx + 16 * j] = v;
      }
    }
    __syncthreads();
    mac_lr(Ss, Vs);
  }
  // row-wise |.| sums across the 16 tx lanes of a row
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) asum[i] += __shfl_xor_sync(0xffffffffu, asum[i], off);
  }
  const float cross = (c > 0) ? cross_scale[(static_cast<size_t>(n) * 4 + h) * n_chunks + c] : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int jr = q0 + 4 * ty + i;
    const float sq = sqrtf(static_cast<float>(jr + 1));
    const float inner = fmaxf(1.f, asum[i] / sq);
    const float den = sq * fmaxf(inner, cross);
    float r[4] = {o[i][0] / den, o[i][1] / den, o[i][2] / den, o[i][3] / den};
    float sm = r[0] + r[1] + r[2] + r[3];
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, off);
    const float mean = sm * (1.f / 64.f);
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) m2 = fmaf(r[j] - mean, r[j] - mean, m2);
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, off);
    const float rstd = 1.f / sqrtf(m2 * (1.f / 64.f) + 1e-6f);
    if (jr < chunk) {
      const size_t grow = (static_cast<size_t>(b) * T + static_cast<size_t>(c) * chunk + jr) * S + s;
      const float4 g = __ldg(reinterpret_cast<const float4*>(qkvg + grow * 1024 + 768 + h * 64) + tx);
      const float gg[4] = {g.x, g.y, g.z, g.w};
      float4 ov;
      ov.x = (r[0] - mean) * rstd * (gg[0] / (1.f + expf(-gg[0])));
      ov.y = (r[1] - mean) * rstd * (gg[1] / (1.f + expf(-gg[1])));
      ov.z = (r[2] - mean) * rstd * (gg[2] / (1.f + expf(-gg[2])));
      ov.w = (r[3] - mean) * rstd * (gg[3] / (1.f + expf(-gg[3])));
      *reinterpret_cast<float4*>(out + grow * 256 + h * 64 + 4 * tx) = ov;
    }
  }
}

// Recurrent retention step (modules/retention.py:126-144), fp32 in / fp32 state / fp32 out; see elementwise.cu.
__global__ void __launch_bounds__(64)
p32_ret_step_kernel(const float* __restrict__ qkvg, float* __restrict__ state, int t_arg, const int* __restrict__ t_dev,
                    float* __restrict__ out) {
  const int n = blockIdx.x, h = blockIdx.y, d = threadIdx.x;
  const int t = t_dev ? *t_dev : t_arg;
  __shared__ float q_s[64], k_s[64], red[2];
  const float* row = qkvg + static_cast<size_t>(n) * 1024 + h * 64;
  q_s[d] = row[d];
  k_s[d] = row[256 + d];
  const float v = row[512 + d];
  const float g = row[768 + d];
  __syncthreads();
  float* st = state + (static_cast<size_t>(n) * 4 + h) * 4096;
  const float a = sqrtf(static_cast<float>(t)) / sqrtf(static_cast<float>(t + 1));
  const float bsc = 1.f / sqrtf(static_cast<float>(t + 1));
  float o = 0.f;
#pragma unroll 8
  for (int e = 0; e < 64; ++e) {
    float kv = st[e * 64 + d];
    kv = (t > 0 ? kv * a : 0.f) + k_s[e] * v * bsc;
    st[e * 64 + d] = kv;
    o = fmaf(q_s[e], kv, o);
  }
  float sm = o;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, off);
  if ((d & 31) == 0) red[d >> 5] = sm;
  __syncthreads();
  const float mean = (red[0] + red[1]) * (1.f / 64.f);
  __syncthreads();
  float dv = (o - mean) * (o - mean);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) dv += __shfl_xor_sync(0xffffffffu, dv, off);
  if ((d & 31) == 0) red[d >> 5] = dv;
  __syncthreads();
  const float rstd = 1.f / sqrtf((red[0] + red[1]) * (1.f / 64.f) + 1e-6f);
  out[static_cast<size_t>(n) * 256 + h * 64 + d] = (o - mean) * rstd * (g / (1.f + expf(-g)));
}

__global__ void p32_hist_append_kernel(const float* __restrict__ src, float* __restrict__ hist, int cap, int pos_arg,
                                       const int* __restrict__ pos_dev) {
  const int n = blockIdx.x;
  const int pos = pos_dev ? *pos_dev : pos_arg;
  hist[(static_cast<size_t>(n) * cap + pos) * 256 + threadIdx.x] = src ? src[static_cast<size_t>(n) * 256 + threadIdx.x] : 0.f;
}

}  // namespace

// =====================================================================================================================
void launch_p32_gemm(const CUtensorMap& tmWhi, const CUtensorMap& tmWlo, const P32GemmParams& p, cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(p32_gemm_kernel<P32_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    cudaFuncSetAttribute(p32_gemm_kernel<P32_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    cudaFuncSetAttribute(p32_gemm_kernel<P32_SWISH>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
  }
  const int tiles_per_seq = (p.rows_per_seq + BM - 1) / BM;
  const int grid = p.n_seq * tiles_per_seq * (p.N / BN);
  const bool train = p.a_transposed || p.w_seq_stride || p.a_scale_dev || p.out_scale_dev || p.out_mask;
  // persistent warp-specialised kernel once there is more than one tile per SM (FSEEND_P32_GEMM=0: always one-tile)
  static int num_sms = 0;
  int use_persist = 1;
  static PerDeviceOnce once_p;
  if (once_p.first()) {
    cudaFuncSetAttribute(p32_gemm_persist_kernel<P32_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistSmem);
    cudaFuncSetAttribute(p32_gemm_persist_kernel<P32_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistSmem);
    cudaFuncSetAttribute(p32_gemm_persist_kernel<P32_SWISH>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistSmem);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(p32_gemm_kernel<P32_NONE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    cudaFuncSetAttribute(p32_gemm_persist_kernel<P32_NONE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistSmem);
    cudaFuncSetAttribute(p32_gemm_astat_kernel<P32_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAstatSmem);
    cudaFuncSetAttribute(p32_gemm_astat_kernel<P32_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAstatSmem);
    cudaFuncSetAttribute(p32_gemm_astat_kernel<P32_SWISH>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAstatSmem);
  }
  {
    // read per launch (tests flip it): 0 = one-tile kernel only, 1 (default) = + persistent kernel, 2 = + A-stationary
    // kernel where it applies.  The A-stationary form is correct but measured NO faster (26.1 vs 26.0 ms per parity forward):
    // its profile (profiles/r02_p32astat.txt) moves the bound to the fold / epilogue warps (accurate expf + reciprocal of
    // the swish epilogue: 20 % of the samples; tail imbalance at the final barrier: 21 %), i.e. re-splitting A per column
    // tile was not what limits the persistent kernel.
    use_persist = 1;
    if (const char* e = getenv("FSEEND_P32_GEMM")) use_persist = (e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  if (train) {        // backward products (train_ops.cu): transposed / scaled / masked A, per-sequence W; no activation
    if (grid > num_sms) p32_gemm_persist_kernel<P32_NONE, true><<<num_sms, kPThreads, kPersistSmem, st>>>(tmWhi, tmWlo, p, grid);
    else p32_gemm_kernel<P32_NONE, true><<<grid, 256, kGemmSmem, st>>>(tmWhi, tmWlo, p);
    return;
  }
  if (use_persist >= 2 && p.taps == 1 && p.k_blocks <= kSMaxKB && p.N >= 2 * BN && p.n_seq * tiles_per_seq > num_sms &&
      p.a_row_offset == 0 && p.a_row_offset_dev == nullptr && p.tap_shift == 0 && p.a_seq_rows == p.rows_per_seq) {
    const int n_m = p.n_seq * tiles_per_seq;
    if (p.act == P32_RELU) p32_gemm_astat_kernel<P32_RELU><<<num_sms, kPThreads, kAstatSmem, st>>>(tmWhi, tmWlo, p, n_m);
    else if (p.act == P32_SWISH) p32_gemm_astat_kernel<P32_SWISH><<<num_sms, kPThreads, kAstatSmem, st>>>(tmWhi, tmWlo, p, n_m);
    else p32_gemm_astat_kernel<P32_NONE><<<num_sms, kPThreads, kAstatSmem, st>>>(tmWhi, tmWlo, p, n_m);
    return;
  }
  if (use_persist && grid > num_sms) {
    if (p.act == P32_RELU) p32_gemm_persist_kernel<P32_RELU><<<num_sms, kPThreads, kPersistSmem, st>>>(tmWhi, tmWlo, p, grid);
    else if (p.act == P32_SWISH) p32_gemm_persist_kernel<P32_SWISH><<<num_sms, kPThreads, kPersistSmem, st>>>(tmWhi, tmWlo, p, grid);
    else p32_gemm_persist_kernel<P32_NONE><<<num_sms, kPThreads, kPersistSmem, st>>>(tmWhi, tmWlo, p, grid);
    return;
  }
  if (p.act == P32_RELU) p32_gemm_kernel<P32_RELU><<<grid, 256, kGemmSmem, st>>>(tmWhi, tmWlo, p);
  else if (p.act == P32_SWISH) p32_gemm_kernel<P32_SWISH><<<grid, 256, kGemmSmem, st>>>(tmWhi, tmWlo, p);
  else p32_gemm_kernel<P32_NONE><<<grid, 256, kGemmSmem, st>>>(tmWhi, tmWlo, p);
}


void launch_p32_gemm_planes(const CUtensorMap& tmAhi, const CUtensorMap& tmAlo, const CUtensorMap& tmWhi,
                            const CUtensorMap& tmWlo, const P32GemmParams& p, cudaStream_t st) {
  static PerDeviceOnce once;
  static int num_sms = 148;
  if (once.first()) {
    cudaFuncSetAttribute(p32_gemm_planes_kernel<P32_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPlanesSmem);
    cudaFuncSetAttribute(p32_gemm_planes_kernel<P32_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPlanesSmem);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_tiles = ((p.rows_per_seq + BM - 1) / BM) * (p.N / BN);
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  if (p.act == P32_RELU) p32_gemm_planes_kernel<P32_RELU><<<grid, kTThreads, kPlanesSmem, st>>>(tmAhi, tmAlo, tmWhi, tmWlo, p, n_tiles);
  else p32_gemm_planes_kernel<P32_NONE><<<grid, kTThreads, kPlanesSmem, st>>>(tmAhi, tmAlo, tmWhi, tmWlo, p, n_tiles);
}

bool launch_p32_rowvec(const __half* whi, const __half* wlo, const P32GemmParams& p, cudaStream_t st) {
  const int R = p.n_seq * p.rows_per_seq;
  if (R < 1 || R > kRvMaxRows || (p.k_blocks * 64) % 8 || p.lda % 4) return false;
  if (p.row_epi_counter && (p.N != 256 || p.ldo != 256)) return false;
  const int grid = (p.N + kRvWarps * kRvCols - 1) / (kRvWarps * kRvCols);
  if (p.act == P32_RELU) p32_rowvec_kernel<P32_RELU><<<grid, kRvWarps * 32, 0, st>>>(whi, wlo, p);
  else if (p.act == P32_SWISH) p32_rowvec_kernel<P32_SWISH><<<grid, kRvWarps * 32, 0, st>>>(whi, wlo, p);
  else p32_rowvec_kernel<P32_NONE><<<grid, kRvWarps * 32, 0, st>>>(whi, wlo, p);
  return true;
}

void launch_p32_step_attn(const float* qkv, float* kcache, float* vcache, int n_seq, int cap, int pos, float scale,
                          float* out, cudaStream_t st, const int* pos_dev) {
  p32_step_attn_kernel<<<dim3(n_seq, 4), 128, 0, st>>>(qkv, kcache, vcache, cap, pos, pos_dev, scale, out);
}

void launch_p32_layernorm(const float* x, int rows, const float* g1, const float* b1, float* out1, const float* g2,
                          const float* b2, float* out2, float eps, const int* seq_len, int rows_per_seq,
                          cudaStream_t st) {
  if (rows <= 0) return;
  p32_layernorm_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, rows, g1, b1, out1, g2, b2, out2, eps, seq_len,
                                                      rows_per_seq > 0 ? rows_per_seq : rows);
}

void launch_p32_pad_input(const float* x, const int* cu, int B, int Tmax, int Din, int Kpad, float* out, cudaStream_t st,
                          const float* sc, const float* sh, float pad_value) {
  p32_pad_input_kernel<<<B * Tmax, 128, 0, st>>>(x, cu, Tmax, Din, Kpad, out, sc, sh, pad_value);
}

void launch_p32_glu(const float* h, int rows, float* out, cudaStream_t st) {
  const size_t n4 = static_cast<size_t>(rows) * 64;
  if (n4 == 0) return;
  p32_glu_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(h, n4, out);
}

int launch_p32_dwconv_bn_swish(const float* u, const float* w, const float* sc, const float* sh, int n_seq, int T, int K,
                               float* hist, float* out, cudaStream_t st) {
  if (K < 1 || K > 32) return -1;
  p32_dwconv_bn_swish_kernel<<<dim3((T + 7) / 8, n_seq), 256, 0, st>>>(u, w, sc, sh, T, K, hist, out);
  return 0;
}

void launch_p32_l2norm(float* x, int rows, cudaStream_t st) {
  if (rows <= 0) return;
  p32_l2norm_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, rows);
}

void launch_p32_convert(const float* y, const float* pe_proj, int rows, int S, float* out, cudaStream_t st) {
  const size_t n4 = static_cast<size_t>(rows) * S * 64;
  if (n4 == 0) return;
  p32_convert_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(y, pe_proj, n4, S, out);
}

int launch_p32_spk_attn(const float* qkv, float* out, int n_frames, int S, float scale, cudaStream_t st) {
  if (S < 1 || S > kMaxS) return -1;
  if (n_frames == 0) return 0;
  const int fpb = 128 / (4 * S);                               // whole frames per 128-thread block
  const int smem = fpb * S * kSpkRow * static_cast<int>(sizeof(float));      // <= 2 * 16 * 2176 B = 68 KB
  static PerDeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(p32_spk_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
  p32_spk_attn_kernel<<<(n_frames + fpb - 1) / fpb, 128, smem, st>>>(qkv, out, n_frames, S, scale, fpb);
  return 0;
}

void launch_p32_head(const float* emb, const float* att, int n_frames, int S, float* logits, float* emb_out,
                     float* att_out, cudaStream_t st) {
  if (n_frames <= 0) return;
  p32_head_kernel<<<(n_frames + 7) / 8, 256, 0, st>>>(emb, att, n_frames, S, logits, emb_out, att_out);
}

void launch_p32_ret_chunk_state(const float* qkvg, int B, int S, int T, int chunk, float* state, float* cross_scale,
                                cudaStream_t st) {
  p32_ret_chunk_state_kernel<<<B * S * 4, 256, 0, st>>>(qkvg, S, T, chunk, T / chunk, state, cross_scale);
}

void launch_p32_retention(const float* qkvg, const float* state, const float* cross_scale, int B, int S, int T,
                          int chunk, float* out, cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(p32_retention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRetSmem);
  }
  const int nc = T / chunk;
  dim3 grid((chunk + 63) / 64, 4, B * S * nc);
  p32_retention_kernel<<<grid, 256, kRetSmem, st>>>(qkvg, state, cross_scale, S, T, chunk, nc, out);
}

void launch_p32_ret_step(const float* qkvg, float* state, int n_seq, int t, float* out, cudaStream_t st,
                         const int* t_dev) {
  p32_ret_step_kernel<<<dim3(n_seq, 4), 64, 0, st>>>(qkvg, state, t, t_dev, out);
}

void launch_p32_hist_append(const float* src, float* hist, int n_seq, int cap, int pos, cudaStream_t st,
                            const int* pos_dev) {
  p32_hist_append_kernel<<<n_seq, 256, 0, st>>>(src, hist, cap, pos, pos_dev);
}

}  // namespace fseend
