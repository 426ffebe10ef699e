// tcgen05 causal flash-attention forward, head_dim 64 (see attn.cuh).
//
// One CTA = one (sequence, head, 128-query tile).  160 threads:
//   warps 0-3 : softmax; thread r owns query row r (TMEM lane r).  Online softmax in fp32; P is written
//               as fp16 into a 128B-swizzled smem tile (A operand of the PV MMA); O accumulates in registers.
//   warp 4    : lane 0 issues TMA loads (Q once; K/V tiles through a 2-stage ring) and the tcgen05 MMAs
//               S = Q K^T (M128 N128 K64) and PV = P V (M128 N64 K<=128, V consumed MN-major straight from
//               its row-major [kv][64] TMA tile).
// Causality skips whole KV tiles above the diagonal and trims the PV K-extent on the diagonal tile.
// Two CTAs per SM (112 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
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
constexpr int kSmemBytes = kOffP + kPBytes + 1024;
constexpr uint32_t kTmemCols = 256;  // S: [0,128), PV: [128,192)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_ld32_sync(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tile_write32(uint8_t* tile, int r, int c, const float (&v)[32]) {
  uint8_t* sub = tile + (c >> 1) * (kTile * 128);
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

__global__ void __launch_bounds__(160, 2)
attn_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full[2], v_full[2], kv_empty[2], s_full, p_ready, pv_full;
  __shared__ uint32_t tmem_base_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  const int n_qt = (p.T + kTile - 1) / kTile;
  const int qt = n_qt - 1 - static_cast<int>(blockIdx.x);  // heaviest (latest) query tiles first
  const int h = blockIdx.y;
  const int b = blockIdx.z / p.S;
  const int s = blockIdx.z % p.S;
  const int q0 = qt * kTile;
  const int last_key = min(q0 + kTile - 1 + p.mask_delay, p.T - 1);
  const int n_kv = last_key / kTile + 1;

  if (tid == 0) {
    mbar_init(&q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(&s_full, 1);
    mbar_init(&p_ready, 128);
    mbar_init(&pv_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 4) tmem_alloc(&tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_PV = tmem_base + 128;

  if (warp == 4) {
    if (lane == 0) {
      // ------------------------------------------------------------ control thread: TMA + MMA issue
      auto load_kv = [&](int j) {
        const int st = j & 1;
        mbar_arrive_expect_tx(&k_full[st], kKVBytes);
        tma_load_4d(smem + kOffK + st * kKVBytes, &tmQKV, &k_full[st], 256 + h * 64, s, j * kTile, b);
        mbar_arrive_expect_tx(&v_full[st], kKVBytes);
        tma_load_4d(smem + kOffV + st * kKVBytes, &tmQKV, &v_full[st], 512 + h * 64, s, j * kTile, b);
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
        umma_commit(&s_full);
      };

      mbar_arrive_expect_tx(&q_full, kQBytes);
      tma_load_4d(smem + kOffQ, &tmQKV, &q_full, h * 64, s, q0, b);
      load_kv(0);
      if (n_kv > 1) load_kv(1);
      mbar_wait(&q_full, 0, 10);
      issue_qk(0);

      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&p_ready, j & 1, 12);  // P(j) in smem, S(j) fully read, PV(j-1) fully read
        mbar_wait(&v_full[st], (j >> 1) & 1, 13);
        tc_fence_after();
        const int valid_cols = min(kTile, last_key - j * kTile + 1);
        const int n_k16 = (valid_cols + 15) >> 4;
        const uint32_t p_addr = smem_u32(smem + kOffP);
        const uint64_t vdesc = smem_desc_sw128(smem_u32(smem + kOffV + st * kKVBytes));
        for (int kk = 0; kk < n_k16; ++kk) {
          const uint64_t pdesc = smem_desc_sw128(p_addr + (kk >> 2) * (kTile * 128)) + 2 * (kk & 3);
          // V is MN-major: 16 kv rows = 16 * 128 B = 2048 B per K step -> +128 in 16-byte units
          umma_f16(tmem_PV, pdesc, vdesc + 128 * kk, idesc_pv, kk > 0 ? 1u : 0u);
        }
        umma_commit(&pv_full);
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
    const int qi = q0 + r;             // global query index
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const float sl = p.scale * 1.4426950408889634f;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* ptile = smem + kOffP;
    float v[32];

    for (int j = 0; j < n_kv; ++j) {
      // columns c of this KV tile visible to this row: c <= limit
      const int limit = min(qi + p.mask_delay, p.T - 1) - j * kTile;
      // warp-uniform bound so that tcgen05.ld stays convergent: the last lane has the largest limit
      const int wlimit = min(q0 + warp * 32 + 31 + p.mask_delay, p.T - 1) - j * kTile;
      const int n_chunks = wlimit < 0 ? 0 : min(4, (wlimit >> 5) + 1);

      mbar_wait(&s_full, j & 1, 20);
      tc_fence_after();

      float mx = -INFINITY;
      for (int c = 0; c < n_chunks; ++c) {
        tmem_ld32_sync(tmem_S + lane_base + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c * 32 + i <= limit) ? v[i] : -INFINITY);
      }
      const float m_new = fmaxf(m_run, mx);
      // m_new is finite from tile 0 on (column 0 is visible to every row); guard anyway
      const float m_scaled = (m_new == -INFINITY) ? 0.f : m_new * sl;
      const float alpha = ex2(m_run * sl - m_scaled);   // m_run = -inf -> 0
      float psum = 0.f;

      if (j > 0) {
        // PV(j-1) must have consumed P(j-1) before it is overwritten; fold its result into O
        mbar_wait(&pv_full, (j - 1) & 1, 21);
        tc_fence_after();
      }
      for (int c = 0; c < 4; ++c) {
        if (c < n_chunks) {
          tmem_ld32_sync(tmem_S + lane_base + c * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = ex2(fmaf(v[i], sl, -m_scaled));
            v[i] = (c * 32 + i <= limit) ? e : 0.f;
            psum += v[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        tile_write32(ptile, r, c, v);
      }
      if (j > 0) {
        // O = O * alpha_{j-1} + PV(j-1): alpha_prev was applied lazily -> apply here
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld32_sync(tmem_PV + lane_base + c * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] += v[i];
        }
      }
      // rescale the accumulated O (which now includes PV(j-1), computed against m_run) to m_new
#pragma unroll
      for (int i = 0; i < 64; ++i) o[i] *= alpha;
      l_run = l_run * alpha + psum;
      m_run = m_new;

      fence_proxy_async_smem();   // P visible to the tensor-core (async) proxy
      tc_fence_before();          // order our tcgen05.ld of S / PV before the next MMAs
      mbar_arrive(&p_ready);
    }
    // last PV
    mbar_wait(&pv_full, (n_kv - 1) & 1, 22);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      tmem_ld32_sync(tmem_PV + lane_base + c * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c * 32 + i] += v[i];
    }
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    // stage O (fp16) in the P buffer (all MMAs reading it have completed) and TMA-store it
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = o[c * 32 + i] * inv;
      tile_write32(ptile, r, c, v);   // c in {0,1} -> sub-tile 0 (64 columns)
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (tid == 0) {
      tma_store_4d(&tmO, ptile, h * 64, s, q0, b);
      tma_store_commit();
      tma_store_wait_read0();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

void launch_causal_attn(const CUtensorMap& tmQKV, const CUtensorMap& tmO, const AttnParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    attr_set = true;
  }
  dim3 grid((p.T + kTile - 1) / kTile, p.H, p.B * p.S);
  attn_kernel<<<grid, 160, kSmemBytes, stream>>>(tmQKV, tmO, p);
}

}  // namespace fseend
