// Embedding-consistency loss:
//   FS-EEND (reference FS model file :46-57):  mean over B*T*T of ( E E^T/(|e||e|^T + 1e-6) - L L^T/(|l||l|^T + 1e-6) )^2
//   LS-EEND (reference LS model file :92-113): the same sum restricted to rows/cols < ilen[b], divided by sum(ilen^2).
// The reference materialises four (B,T,T) fp32 maps; here one CTA owns a 128 x 128 tile of one sequence's (T,T) map:
// the cosine Gram tile is one tcgen05 accumulator (M128 N128, K = 256 in 16 MMAs), the label Gram tile (K = S <= 16) is
// evaluated on CUDA cores in the epilogue, and only the squared-difference partial sum leaves the SM.  Both maps are
// symmetric, so only tiles on or below the diagonal are computed and off-diagonal tiles count twice.
//
// The embeddings arrive as the fp32 rows the model returns; they are rounded to fp16 while being staged into the
// 128B-swizzled K-major operand layout (plain loads: the op has no TMA descriptor to build and no alignment demands
// beyond 16 bytes).  Row norms are computed from the same fp16 values the MMA consumes.
// Partial sums go to a workspace and are reduced in a fixed order in fp64 by embloss_finalize_kernel (deterministic).
#include "once.h"
#include "embloss.cuh"
#include "ptx.cuh"

namespace fseend {

namespace {

constexpr int kD = 256;                      // embedding width
constexpr int kTile = 128;
constexpr int kKBlocks = kD / 64;
constexpr int kOperandBytes = kTile * kD * 2;            // 64 KB: 4 k-blocks of [128 rows][64 halfs]
constexpr int kMaxS = 16;
constexpr int kSmemBytes = 2 * kOperandBytes + 1024;     // row tile + column tile + alignment slack
constexpr uint32_t kTmemCols = 128;

// Stage 128 fp32 rows (zero beyond `lim`) as fp16 into the swizzled operand tile and return this thread's row norm.
// Warp w handles rows w, w+4, ...: one fully coalesced 1 KB row per iteration (lane l: floats [4l,4l+4) and
// [128+4l, 128+4l+4)); the norm of a row is a warp reduction, written to norms[row].
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, int row0, int lim, uint8_t* tile,
                                           float* norms, int warp, int lane) {
  for (int r = warp; r < kTile; r += 4) {
    const int row = row0 + r;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (row < lim) {
      const float4* p = reinterpret_cast<const float4*>(src + static_cast<size_t>(row) * kD);
      a = __ldg(p + lane);
      b = __ldg(p + 32 + lane);
    }
    const __half2 a0 = __floats2half2_rn(a.x, a.y), a1 = __floats2half2_rn(a.z, a.w);
    const __half2 b0 = __floats2half2_rn(b.x, b.y), b1 = __floats2half2_rn(b.z, b.w);
    // column c = 4*lane (+128): k-block c/64, 16-byte chunk (c%64)/8, 8-byte half of the chunk (lane&1)
    {
      const int c = 4 * lane;
      uint8_t* dst = tile + (c >> 6) * (kTile * 128) + sw128_offset(r, (c & 63) >> 3) + (lane & 1) * 8;
      *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&a0),
                                                  *reinterpret_cast<const uint32_t*>(&a1));
    }
    {
      const int c = 128 + 4 * lane;
      uint8_t* dst = tile + (c >> 6) * (kTile * 128) + sw128_offset(r, (c & 63) >> 3) + (lane & 1) * 8;
      *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&b0),
                                                  *reinterpret_cast<const uint32_t*>(&b1));
    }
    const float2 fa0 = __half22float2(a0), fa1 = __half22float2(a1), fb0 = __half22float2(b0), fb1 = __half22float2(b1);
    float ss = fa0.x * fa0.x + fa0.y * fa0.y + fa1.x * fa1.x + fa1.y * fa1.y + fb0.x * fb0.x + fb0.y * fb0.y +
               fb1.x * fb1.x + fb1.y * fb1.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) norms[r] = sqrtf(ss);
  }
}

// labels of 128 rows -> smem [128][kMaxS] (zero padded) and their norms
__device__ __forceinline__ void stage_labels(const float* __restrict__ lab, int row0, int lim, int S, float* dst,
                                             float* norms, int tid) {
  const int row = row0 + tid;
  float ss = 0.f;
#pragma unroll
  for (int s = 0; s < kMaxS; ++s) {
    float v = 0.f;
    if (s < S && row < lim) v = __ldg(lab + static_cast<size_t>(row) * S + s);
    dst[tid * kMaxS + s] = v;
    ss += v * v;
  }
  norms[tid] = sqrtf(ss);
}

__global__ void __launch_bounds__(128)
embloss_kernel(const EmbLossParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float en_i[kTile], en_j[kTile], ln_i[kTile], ln_j[kTile];
  __shared__ __align__(16) float lab_i[kTile * kMaxS];
  __shared__ __align__(16) float lab_j[kTile * kMaxS];
  __shared__ float warp_sum[4];

  // 1024-byte alignment as an OFFSET from the shared array: pointer arithmetic through uintptr_t would hide the shared
  // address space from the compiler and turn every access below into a generic load / store
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* tile_i = smem;
  uint8_t* tile_j = smem + kOperandBytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // blockIdx.x -> (sequence, ti >= tj)
  const int b = blockIdx.x / p.n_pairs;
  int pr = blockIdx.x % p.n_pairs, ti = 0;
  while (pr >= ti + 1) { pr -= ti + 1; ++ti; }
  const int tj = pr;
  const int lim = p.seq_len ? min(p.seq_len[b], p.T) : p.T;   // rows/cols >= lim do not contribute
  const bool diag = ti == tj;
  const bool active = tj * kTile < lim && ti * kTile < lim;

  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_slot, kTmemCols);

  const float* emb_b = p.emb + static_cast<size_t>(b) * p.T * kD;
  const float* lab_b = p.labels + static_cast<size_t>(b) * p.T * p.S;
  if (active) {
    stage_rows(emb_b, ti * kTile, lim, tile_i, en_i, warp, lane);
    stage_labels(lab_b, ti * kTile, lim, p.S, lab_i, ln_i, tid);
    if (!diag) {
      stage_rows(emb_b, tj * kTile, lim, tile_j, en_j, warp, lane);
      stage_labels(lab_b, tj * kTile, lim, p.S, lab_j, ln_j, tid);
    }
  }
  fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  float acc_sum = 0.f;
  if (active) {
    if (warp == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kTile, kTile, false);
      const uint32_t sa = smem_u32(tile_i), sb = smem_u32(diag ? tile_i : tile_j);
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < kKBlocks; ++kb) {
          const uint64_t adesc = smem_desc_sw128(sa + kb * (kTile * 128));
          const uint64_t bdesc = smem_desc_sw128(sb + kb * (kTile * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb | kk) ? 1u : 0u);
        }
        umma_commit(&mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(&mma_bar, 0, 40);
    tc_fence_after();

    const float* nj = diag ? en_i : en_j;
    const float* lj = diag ? lab_i : lab_j;
    const float* lnj = diag ? ln_i : ln_j;
    const int row = ti * kTile + tid;
    const float ni = en_i[tid], li_n = ln_i[tid];
    float li[kMaxS];
#pragma unroll
    for (int s = 0; s < kMaxS; ++s) li[s] = lab_i[tid * kMaxS + s];
    const int s4 = (p.S + 3) >> 2;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < kTile; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + c0, v);
      tmem_ld_wait();
      if (row < lim) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int c = c0 + i;
          const float g = __uint_as_float(v[i]) / (ni * nj[c] + 1e-6f);
          float l = 0.f;
          const float4* lr = reinterpret_cast<const float4*>(lj + c * kMaxS);
#pragma unroll
          for (int q = 0; q < kMaxS / 4; ++q) {
            if (q < s4) {   // warp-uniform
              const float4 t = lr[q];
              l = fmaf(li[4 * q + 0], t.x, l);
              l = fmaf(li[4 * q + 1], t.y, l);
              l = fmaf(li[4 * q + 2], t.z, l);
              l = fmaf(li[4 * q + 3], t.w, l);
            }
          }
          l = l / (li_n * lnj[c] + 1e-6f);
          const float d = (tj * kTile + c < lim) ? g - l : 0.f;
          acc_sum = fmaf(d, d, acc_sum);
        }
      }
    }
    if (!diag) acc_sum *= 2.f;   // the mirrored tile (tj, ti)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc_sum += __shfl_xor_sync(0xffffffffu, acc_sum, o);
  if (lane == 0) warp_sum[warp] = acc_sum;
  tc_fence_before();
  __syncthreads();
  if (tid == 0) p.partials[blockIdx.x] = (warp_sum[0] + warp_sum[1]) + (warp_sum[2] + warp_sum[3]);
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

__global__ void __launch_bounds__(256)
embloss_finalize_kernel(const float* __restrict__ partials, int n, double inv_divisor, float* __restrict__ loss) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += static_cast<double>(partials[i]);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = static_cast<float>(sh[0] * inv_divisor);
}

}  // namespace

int embloss_num_partials(int B, int T) {
  const int nt = (T + kTile - 1) / kTile;
  return B * (nt * (nt + 1) / 2);
}

void launch_embloss(EmbLossParams p, double divisor, float* loss, cudaStream_t stream) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(embloss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  }
  const int nt = (p.T + kTile - 1) / kTile;
  p.n_pairs = nt * (nt + 1) / 2;
  const int grid = p.B * p.n_pairs;
  embloss_kernel<<<grid, 128, kSmemBytes, stream>>>(p);
  embloss_finalize_kernel<<<1, 256, 0, stream>>>(p.partials, grid, 1.0 / divisor, loss);
}

}  // namespace fseend
