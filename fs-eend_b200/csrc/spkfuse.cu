// Speaker-axis self-attention with its QKV projection fused in (reference: the live path of
// TransformerEncoderFusionLayer, FS-EEND/nnet/modules/merge_tfm_encoder.py:366-372 / _sa_block2 — nn.MultiheadAttention
// over the S attractor slots of one frame; LS-EEND/nnet/modules/merge_retnet_layer.py:244-249).
//
// The unfused path wrote the [frames*S][768] fp16 QKV activations to HBM (295 MB at B=64, T=500, S=6) and read them
// back in a block-diagonal tcgen05 attention whose 128x128 score tiles were 95 % masked: 0.075 + 0.114 ms per layer.
// Here a work item is (row tile, head): the tile's q, k, v of that head are projected on the tensor cores
// ([128 rows] x [192 = 64 q | 64 k | 64 v] accumulators in TMEM, K = 256) and the attention is finished on CUDA cores —
// S keys per row, all of them rows of the same tile (tiles hold whole frames: tile_rows = (128 / S) * S), exchanged
// through shared memory in fp32.  HBM traffic: X read (4 heads of a tile run back to back: L2 hits), output written.
//
// PERSISTENT, one CTA per SM, 320 threads:
//   warps 0-3 / 4-7 : two epilogue groups working on alternate items, each with its own TMEM accumulator (2 x 192
//               columns) and its own k / v buffers; thread r of a group <-> tile row r (TMEM lane r).  The groups run
//               out of phase, so the TMEM reads, shared-memory traffic and FMA chains of one overlap the other's.
//   warp 8    : TMA producer (X k-block + the head's q / k / v weight rows), 2-stage ring, runs ahead across items.
//   warp 9    : tcgen05 issuer.  An accumulator is released as soon as q / k / v have left TMEM.
// History (B=64, T=500, S=6, in-model): one item per CTA, 2 CTAs/SM: 0.206 ms (every latency exposed once per item);
// persistent with 8 warps in lock step on one item (two threads per row): 0.219 ms with the row reads emitted as
// generic loads (pointer re-alignment arithmetic hid the shared address space), 0.167 ms as LDS.
#include "once.h"
#include "ptx.cuh"
#include "spkfuse.cuh"

namespace fseend {

namespace {

constexpr int BM = 128, BN = 192, BK = 64;
constexpr int kStages = 3;
constexpr int kKBlocks = 256 / BK;
constexpr int kABytes = BM * BK * 2;                 // 16 KB
constexpr int kBBytes = BN * BK * 2;                 // 24 KB: 64 q rows | 64 k rows | 64 v rows of this head
constexpr int kStageBytes = kABytes + kBBytes;       // 40 KB
// k / v rows in shared memory: fp16 (the precision the unfused path had), 128 B per row + 16 B pad.  The shared-memory
// data pipe is what bounds this kernel (ncu: 78 % of peak wavefronts with fp32 rows): the S query rows of a frame all
// read the same key row, but a 128-bit LDS still costs one wavefront per quarter warp, so bytes per row are what counts.
// Row stride 144 B = 9 bank groups: the rows a quarter warp writes (consecutive) or reads (<= 3 frames, S rows apart)
// fall in distinct bank groups.
constexpr int kRowB = 144;
constexpr int kBufBytes = BM * kRowB;                       // 18 432
constexpr int kOffRing = 0;
constexpr int kOffKV = kOffRing + kStages * kStageBytes;    // [2 groups][k | v]
constexpr int kOffBias = kOffKV + 4 * kBufBytes;            // in_proj_bias [768]
constexpr int kOffBar = kOffBias + 768 * 4;                 // mbarriers + TMEM base slot
constexpr int kDynBytes = kOffBar + 128;                    // ~ 155 KB
constexpr uint32_t kTmemCols = 512;                         // accumulators at columns [0,192) and [256,448)
constexpr int kThreads = 320;
static_assert(kDynBytes <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads, 1)
spkfuse_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const SpkFuseParams p,
               const int n_items) {
  // No static __shared__ and no pointer re-alignment arithmetic: the dynamic buffer is the only shared allocation, so
  // it starts 1024-aligned and the compiler keeps every access below in the shared address space (LDS / STS; with a
  // uintptr_t round-up the k / v row reads of the epilogue were emitted as generic loads and stalled in L1TEX).
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* full_bar = bars;                 // [kStages]
  uint64_t* empty_bar = bars + kStages;      // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;       // [2]
  uint64_t* acc_free = bars + 2 * kStages + 2;   // [2]
  uint32_t* tmem_base_slot_p = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  static_assert((2 * kStages + 4) * 8 + 4 <= 128, "barrier area");
#define tmem_base_slot (*tmem_base_slot_p)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* bias_s = reinterpret_cast<float*>(smem + kOffBias);
  if (tid == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("[fseend] spkfuse: dynamic smem base not 1024-aligned\n");
    __trap();
  }

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&acc_full[g], 1);
      mbar_init(&acc_free[g], 128);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 9) tmem_alloc(&tmem_base_slot, kTmemCols);
  for (int i = tid; i < 768; i += kThreads) bias_s[i] = __ldg(p.bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 8) {
    // ---------------- TMA producer
    uint32_t it = 0;   // k-blocks issued so far (ring position)
    for (int id = blockIdx.x; id < n_items; id += gridDim.x) {
      const int head = id & 3, t0 = (id >> 2) * p.tile_rows;
      for (int kb = 0; kb < kKBlocks; ++kb, ++it) {
        const int s = it % kStages;
        mbar_wait(&empty_bar[s], ((it / kStages) & 1) ^ 1, 50);
        uint8_t* sa = smem + kOffRing + s * kStageBytes;
        uint8_t* sb = sa + kABytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
          tma_load_3d(sa, &tmX, &full_bar[s], kb * BK, t0, 0);
#pragma unroll
          for (int part = 0; part < 3; ++part)
            tma_load_2d(sb + part * (64 * BK * 2), &tmW, &full_bar[s], kb * BK, part * 256 + head * 64);
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer: acc[128][192] = X[128][256] * W_head[192][256]^T
    constexpr uint32_t idesc = make_idesc_f16(BM, BN, false);
    uint32_t it = 0, n = 0;
    for (int id = blockIdx.x; id < n_items; id += gridDim.x, ++n) {
      const int g = n & 1;                               // accumulator / epilogue group of this item
      mbar_wait(&acc_free[g], ((n >> 1) & 1) ^ 1, 53);   // its previous q / k / v have been read out of TMEM
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + g * 256;
      for (int kb = 0; kb < kKBlocks; ++kb, ++it) {
        const int s = it % kStages;
        mbar_wait(&full_bar[s], (it / kStages) & 1, 51);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + kOffRing + s * kStageBytes);
        const uint64_t adesc = smem_desc_sw128(sa);
        const uint64_t bdesc = smem_desc_sw128(sa + kABytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)
            umma_f16(tmem_acc, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (kb == kKBlocks - 1) umma_commit(&acc_full[g]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue groups: group g = warps 4g .. 4g+3 takes items n = g, g+2, ...; thread r <-> tile row r
    const int g = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t taddr = tmem_base + g * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    const int S = p.S;
    const float qs = p.scale * 1.4426950408889634f;          // scores in log2 units
    uint8_t* kbuf = smem + kOffKV + (2 * g) * kBufBytes;
    uint8_t* vbuf = smem + kOffKV + (2 * g + 1) * kBufBytes;
    const bool active = r < p.tile_rows;                     // rows >= tile_rows belong to the next tile
    const int frame = r / S;
    const int my_row_off = r * kRowB;                        // this thread's k / v row (bytes)
    const int frame_off = active ? (frame * S) * kRowB : 0;  // row 0 of this row's frame
    const int group_bar = 1 + g;

    uint32_t n = g;
    for (int id = blockIdx.x + g * gridDim.x; id < n_items; id += 2 * gridDim.x, n += 2) {
      const int head = id & 3, t0 = (id >> 2) * p.tile_rows;
      mbar_wait(&acc_full[g], (n >> 1) & 1, 52);
      tc_fence_after();
      float q[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t a[32];
        tmem_ld32(taddr + c * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = reinterpret_cast<const float4*>(bias_s + head * 64 + c * 32)[i];
          q[c * 32 + 4 * i + 0] = (__uint_as_float(a[4 * i + 0]) + b4.x) * qs;
          q[c * 32 + 4 * i + 1] = (__uint_as_float(a[4 * i + 1]) + b4.y) * qs;
          q[c * 32 + 4 * i + 2] = (__uint_as_float(a[4 * i + 2]) + b4.z) * qs;
          q[c * 32 + 4 * i + 3] = (__uint_as_float(a[4 * i + 3]) + b4.w) * qs;
        }
      }
      named_bar_sync(group_bar, 128);    // the group is done with its previous item's k / v rows
#pragma unroll
      for (int part = 0; part < 2; ++part) {      // k then v -> fp16 rows in shared memory
        uint8_t* dst = (part == 0 ? kbuf : vbuf) + my_row_off;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4* bs = reinterpret_cast<const float4*>(bias_s + (part + 1) * 256 + head * 64 + c * 32);
          uint32_t a[32];
          tmem_ld32(taddr + 64 + part * 64 + c * 32, a);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b0 = bs[2 * i], b1 = bs[2 * i + 1];
            uint4 u;
            u.x = pack_half2(__uint_as_float(a[8 * i + 0]) + b0.x, __uint_as_float(a[8 * i + 1]) + b0.y);
            u.y = pack_half2(__uint_as_float(a[8 * i + 2]) + b0.z, __uint_as_float(a[8 * i + 3]) + b0.w);
            u.z = pack_half2(__uint_as_float(a[8 * i + 4]) + b1.x, __uint_as_float(a[8 * i + 5]) + b1.y);
            u.w = pack_half2(__uint_as_float(a[8 * i + 6]) + b1.z, __uint_as_float(a[8 * i + 7]) + b1.w);
            reinterpret_cast<uint4*>(dst)[c * 4 + i] = u;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_free[g]);         // this accumulator may be overwritten by the MMAs of item n + 2
      named_bar_sync(group_bar, 128);    // every row's k / v is in shared memory (frames straddle warps)

      float sc[16];
      float m = -INFINITY;
      {
        const uint8_t* kf = kbuf + frame_off;
        for (int i = 0; i < S; ++i) {
          const uint4* k0 = reinterpret_cast<const uint4*>(kf + i * kRowB);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 t = k0[c];
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
            const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&t.z));
            const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&t.w));
            a0 = fmaf(q[8 * c + 0], f0.x, a0);
            a1 = fmaf(q[8 * c + 1], f0.y, a1);
            a2 = fmaf(q[8 * c + 2], f1.x, a2);
            a3 = fmaf(q[8 * c + 3], f1.y, a3);
            a0 = fmaf(q[8 * c + 4], f2.x, a0);
            a1 = fmaf(q[8 * c + 5], f2.y, a1);
            a2 = fmaf(q[8 * c + 6], f3.x, a2);
            a3 = fmaf(q[8 * c + 7], f3.y, a3);
          }
          sc[i] = (a0 + a1) + (a2 + a3);
          m = fmaxf(m, sc[i]);
        }
      }
      float acc[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) acc[c] = 0.f;
      float l = 0.f;
      {
        const uint8_t* vf = vbuf + frame_off;
        for (int i = 0; i < S; ++i) {
          const float p0 = ex2f(sc[i] - m);
          l += p0;
          const uint4* v0 = reinterpret_cast<const uint4*>(vf + i * kRowB);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 t = v0[c];
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
            const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&t.z));
            const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&t.w));
            acc[8 * c + 0] = fmaf(p0, f0.x, acc[8 * c + 0]);
            acc[8 * c + 1] = fmaf(p0, f0.y, acc[8 * c + 1]);
            acc[8 * c + 2] = fmaf(p0, f1.x, acc[8 * c + 2]);
            acc[8 * c + 3] = fmaf(p0, f1.y, acc[8 * c + 3]);
            acc[8 * c + 4] = fmaf(p0, f2.x, acc[8 * c + 4]);
            acc[8 * c + 5] = fmaf(p0, f2.y, acc[8 * c + 5]);
            acc[8 * c + 6] = fmaf(p0, f3.x, acc[8 * c + 6]);
            acc[8 * c + 7] = fmaf(p0, f3.y, acc[8 * c + 7]);
          }
        }
      }
      const float inv = 1.f / l;
      // this row's 64 outputs of head `head`: 128 contiguous bytes, written as four full 32-byte sectors
      if (active && t0 + r < p.rows) {
        __half* orow = p.out + static_cast<size_t>(t0 + r) * 256 + head * 64;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          uint4 u, v;
          u.x = pack_half2(acc[qd * 16 + 0] * inv, acc[qd * 16 + 1] * inv);
          u.y = pack_half2(acc[qd * 16 + 2] * inv, acc[qd * 16 + 3] * inv);
          u.z = pack_half2(acc[qd * 16 + 4] * inv, acc[qd * 16 + 5] * inv);
          u.w = pack_half2(acc[qd * 16 + 6] * inv, acc[qd * 16 + 7] * inv);
          v.x = pack_half2(acc[qd * 16 + 8] * inv, acc[qd * 16 + 9] * inv);
          v.y = pack_half2(acc[qd * 16 + 10] * inv, acc[qd * 16 + 11] * inv);
          v.z = pack_half2(acc[qd * 16 + 12] * inv, acc[qd * 16 + 13] * inv);
          v.w = pack_half2(acc[qd * 16 + 14] * inv, acc[qd * 16 + 15] * inv);
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + qd * 16), "r"(u.x),
                       "r"(u.y), "r"(u.z), "r"(u.w), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                       : "memory");
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, kTmemCols);
#undef tmem_base_slot
}

}  // namespace

void launch_spkfuse(const CUtensorMap& tmX, const CUtensorMap& tmW, const SpkFuseParams& p, cudaStream_t stream) {
  static int num_sms = 0;
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(spkfuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynBytes);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_items = ((p.rows + p.tile_rows - 1) / p.tile_rows) * 4;
  const int grid = n_items < num_sms ? n_items : num_sms;
  spkfuse_kernel<<<grid, kThreads, kDynBytes, stream>>>(tmX, tmW, p, n_items);
}

}  // namespace fseend
