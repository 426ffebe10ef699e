// Tensor-core form of the training attention (train_attn.cu): the same three kernels (forward with log-sum-exp, dQ, dK/dV)
// with every 64 x 64 x 64 tile product on mma.sync.m16n8k16 (fp16 operands, fp32 accumulate) in split precision —
// x = hi + lo, x.y ~= lo.hi + hi.lo + hi.hi, three MMAs — so that results stay fp32-grade (~22 operand mantissa bits;
// accumulation chains are 12 MMAs long into a fresh accumulator that is then added in fp32 registers, the same rule as
// p32.cuh).  Warp-level MMA rather than tcgen05: the backward needs five different products per tile pair with operands
// that are produced in registers (P, dS); on this first version they stay in registers between products (the C fragment
// of one product is the A fragment of the next).  The CUDA-core kernels remain as FSEEND_TRAIN_ATTN=0.
//
// Shared-memory tiles are fp16 [64][72] (row stride 36 words: conflict-free fragment loads); "t" tiles hold the transpose.
#pragma once

namespace tc {

constexpr int LDH = 72;
constexpr int TILE_H = 64 * LDH;          // halfs per tile

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += (ah + al) (bh + bl) without the lo.lo term, small terms first
__device__ __forceinline__ void mma_split(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                          uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma16816(c, al, bh0, bh1);
  mma16816(c, ah, bl0, bl1);
  mma16816(c, ah, bh0, bh1);
}
__device__ __forceinline__ uint32_t ld32(const __half* tile, int row, int col) {
  return *reinterpret_cast<const uint32_t*>(tile + row * LDH + col);
}
// A fragment (16 x 16) of rows r0.., columns k0.. of a row-major tile
__device__ __forceinline__ void load_a(const __half* tile, int r0, int k0, int g, int t, uint32_t (&a)[4]) {
  a[0] = ld32(tile, r0 + g, k0 + 2 * t);
  a[1] = ld32(tile, r0 + g + 8, k0 + 2 * t);
  a[2] = ld32(tile, r0 + g, k0 + 2 * t + 8);
  a[3] = ld32(tile, r0 + g + 8, k0 + 2 * t + 8);
}
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - f.x, y - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// acc[nt] (16 x 64 per warp, nt = 8 column tiles) += A(rows r0.. of a_hi/a_lo, 64 deep) . B^T, B row-major [n][k]
__device__ __forceinline__ void warp_mm_nt(const __half* a_hi, const __half* a_lo, int r0, const __half* b_hi,
                                           const __half* b_lo, int g, int t, float (&acc)[8][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t ah[4], al[4];
    load_a(a_hi, r0, ks * 16, g, t, ah);
    load_a(a_lo, r0, ks * 16, g, t, al);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const uint32_t bh0 = ld32(b_hi, nt * 8 + g, ks * 16 + 2 * t), bh1 = ld32(b_hi, nt * 8 + g, ks * 16 + 2 * t + 8);
      const uint32_t bl0 = ld32(b_lo, nt * 8 + g, ks * 16 + 2 * t), bl1 = ld32(b_lo, nt * 8 + g, ks * 16 + 2 * t + 8);
      mma_split(acc[nt], ah, al, bh0, bh1, bl0, bl1);
    }
  }
}
// acc[nt] += P . B^T where P (16 x 64) lives in registers as a C-fragment set p[8][4] (fp32): the C fragments of column
// tiles 2 ks, 2 ks + 1 are exactly the A fragment of k-step ks.  B row-major [n][k] (k = P's column index).
__device__ __forceinline__ void warp_mm_reg_nt(const float (&p)[8][4], const __half* b_hi, const __half* b_lo, int g, int t,
                                               float (&acc)[8][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t ah[4], al[4];
    split_pair(p[2 * ks][0], p[2 * ks][1], ah[0], al[0]);
    split_pair(p[2 * ks][2], p[2 * ks][3], ah[1], al[1]);
    split_pair(p[2 * ks + 1][0], p[2 * ks + 1][1], ah[2], al[2]);
    split_pair(p[2 * ks + 1][2], p[2 * ks + 1][3], ah[3], al[3]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const uint32_t bh0 = ld32(b_hi, nt * 8 + g, ks * 16 + 2 * t), bh1 = ld32(b_hi, nt * 8 + g, ks * 16 + 2 * t + 8);
      const uint32_t bl0 = ld32(b_lo, nt * 8 + g, ks * 16 + 2 * t), bl1 = ld32(b_lo, nt * 8 + g, ks * 16 + 2 * t + 8);
      mma_split(acc[nt], ah, al, bh0, bh1, bl0, bl1);
    }
  }
}

// 64 x 64 fp32 tile (rows row0.. of src, zero beyond T) * scale -> hi / lo fp16 tiles; 128 threads
__device__ __forceinline__ void load_split(__half* hi, __half* lo, const float* src, int row0, int T, int ld, float scale) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = threadIdx.x + 128 * it, r = idx >> 4, c = (idx & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < T) v = *reinterpret_cast<const float4*>(src + static_cast<size_t>(row0 + r) * ld + c);
    uint2 h, l;
    split_pair(v.x * scale, v.y * scale, h.x, l.x);
    split_pair(v.z * scale, v.w * scale, h.y, l.y);
    *reinterpret_cast<uint2*>(hi + r * LDH + c) = h;
    *reinterpret_cast<uint2*>(lo + r * LDH + c) = l;
  }
}
// the same tile, transposed: out[c][r] = src[row0 + r][c] * scale.  A warp handles all 32 row pairs of one 4-column group
// (stores are conflict-free 32-bit words (rows 2 rp, 2 rp + 1); the 16-byte global reads of a warp hit 64 different rows,
// the other column groups of those rows are served by L1)
__device__ __forceinline__ void load_split_t(__half* hi, __half* lo, const float* src, int row0, int T, int ld, float scale) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = threadIdx.x + 128 * it, rp = idx & 31, c = (idx >> 5) * 4;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (row0 + 2 * rp < T) v0 = *reinterpret_cast<const float4*>(src + static_cast<size_t>(row0 + 2 * rp) * ld + c);
    if (row0 + 2 * rp + 1 < T) v1 = *reinterpret_cast<const float4*>(src + static_cast<size_t>(row0 + 2 * rp + 1) * ld + c);
    const float a[4] = {v0.x, v0.y, v0.z, v0.w}, b[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t h, l;
      split_pair(a[i] * scale, b[i] * scale, h, l);
      *reinterpret_cast<uint32_t*>(hi + (c + i) * LDH + 2 * rp) = h;
      *reinterpret_cast<uint32_t*>(lo + (c + i) * LDH + 2 * rp) = l;
    }
  }
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ---------------------------------------------------------------------------------------------------------------------
// forward: grid (q tiles, heads, sequences), 128 threads; warp w owns query rows i0 + 16 w .. + 16
constexpr int kFwdSmem = 6 * TILE_H * 2;
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ lse, int T, int S_in, int delay,
                float scale, const Dropout drop) {
  extern __shared__ __align__(16) __half smh[];
  __half *Qh = smh, *Ql = Qh + TILE_H, *Kh = Ql + TILE_H, *Kl = Kh + TILE_H, *Vth = Kl + TILE_H, *Vtl = Vth + TILE_H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int i0 = blockIdx.x * 64, h = blockIdx.y, n = blockIdx.z, r0 = warp * 16;
  // sequence n of an interleaved batch [n_seq / S_in][T][S_in][.]: frame t lives at row r0g + t * S_in (S_in = 1: plain [n][T])
  const size_t r0g = static_cast<size_t>(n / S_in) * T * S_in + (n % S_in);
  const float* base = qkv + r0g * 768 + h * 64;
  load_split(Qh, Ql, base, i0, T, 768 * S_in, scale);          // the 64^-1/2 scale is a power of two: folded into q exactly
  float o[8][4] = {}, m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  const int row[2] = {i0 + r0 + g, i0 + r0 + g + 8};
  const int jlast = min(T - 1, i0 + 63 + delay);
  for (int j0 = 0; j0 <= jlast; j0 += 64) {
    __syncthreads();
    load_split(Kh, Kl, base + 256, j0, T, 768 * S_in, 1.f);
    load_split_t(Vth, Vtl, base + 512, j0, T, 768 * S_in, 1.f);
    __syncthreads();
    float s[8][4] = {};
    warp_mm_nt(Qh, Ql, r0, Kh, Kl, g, t, s);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = j0 + nt * 8 + 2 * t + (e & 1), rw = row[e >> 1];
        if (!(col < T && col <= rw + delay)) s[nt][e] = -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const float mnew = fmaxf(m[a], quad_max(mx[a]));      // finite from the first key tile on (key 0 is always visible)
      corr[a] = expf(m[a] - mnew);
      m[a] = mnew;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int a = e >> 1;
        const float p = expf(s[nt][e] - m[a]);
        rs[a] += p;                                          // the normaliser sums the undropped probabilities
        const uint64_t rbase = ((static_cast<uint64_t>(n) * kHeads + h) * T + row[a]) * T;
        s[nt][e] = p * keep_scale(drop, rbase, j0 + nt * 8 + 2 * t + (e & 1));
        o[nt][e] *= corr[a];
      }
#pragma unroll
    for (int a = 0; a < 2; ++a) l[a] = l[a] * corr[a] + quad_sum(rs[a]);
    float of[8][4] = {};
    warp_mm_reg_nt(s, Vth, Vtl, g, t, of);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[nt][e] += of[nt][e];
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    if (row[a] >= T) continue;
    const float inv = 1.f / l[a];
    float* op = out + (r0g + static_cast<size_t>(row[a]) * S_in) * 256 + h * 64 + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<float2*>(op + nt * 8) = make_float2(o[nt][2 * a] * inv, o[nt][2 * a + 1] * inv);
    if (t == 0) lse[(static_cast<size_t>(n) * kHeads + h) * T + row[a]] = m[a] + logf(l[a]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dQ: grid (q tiles, heads, sequences).  Also writes dsum = rowsum(dO * O) for the dK/dV kernel.
constexpr int kDqSmem = 10 * TILE_H * 2 + 64 * 4;
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ dout,
                   const float* __restrict__ lse, float* __restrict__ dqkv, float* __restrict__ dsum, int T, int S_in,
                   int delay, float scale, const Dropout drop, const float* __restrict__ gsc) {
  extern __shared__ __align__(16) __half smh[];
  __half *Qh = smh, *Ql = Qh + TILE_H, *dOh = Ql + TILE_H, *dOl = dOh + TILE_H, *Kh = dOl + TILE_H, *Kl = Kh + TILE_H;
  __half *Kth = Kl + TILE_H, *Ktl = Kth + TILE_H, *Vh = Ktl + TILE_H, *Vl = Vh + TILE_H;
  float* Ds = reinterpret_cast<float*>(Vl + TILE_H);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int i0 = blockIdx.x * 64, h = blockIdx.y, n = blockIdx.z, r0 = warp * 16;
  // sequence n of an interleaved batch [n_seq / S_in][T][S_in][.]: frame t lives at row r0g + t * S_in (S_in = 1: plain [n][T])
  const size_t r0g = static_cast<size_t>(n / S_in) * T * S_in + (n % S_in);
  const float* base = qkv + r0g * 768 + h * 64;
  const float* dob = dout + r0g * 256 + h * 64;
  const size_t stat = (static_cast<size_t>(n) * kHeads + h) * T;
  // gradients are ~1e-6 (fp16-subnormal): dO is multiplied by a power of two (max |dO| -> [16, 32)) before the split,
  // every quantity derived from it (D, dP, dS) carries the factor, the result is multiplied by its inverse
  const float gscale = __ldg(gsc), ginv = __ldg(gsc + 1);
  load_split(Qh, Ql, base, i0, T, 768 * S_in, scale);
  load_split(dOh, dOl, dob, i0, T, 256 * S_in, gscale);
  for (int rr = 0; rr < 16; ++rr) {                      // D = rowsum(dO * O), one row per iteration, lanes over d
    const int rw = i0 + r0 + rr;
    float d = 0.f;
    if (rw < T) {
      const float* op = out + (r0g + static_cast<size_t>(rw) * S_in) * 256 + h * 64;
      const float* dp = dob + static_cast<size_t>(rw) * 256 * S_in;
      d = op[lane] * dp[lane] + op[lane + 32] * dp[lane + 32];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) {
      Ds[r0 + rr] = d * gscale;
      if (rw < T) dsum[stat + rw] = d;
    }
  }
  __syncthreads();
  const int row[2] = {i0 + r0 + g, i0 + r0 + g + 8};
  float D[2], L[2], dq[8][4] = {};
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    D[a] = Ds[r0 + g + 8 * a];
    L[a] = row[a] < T ? lse[stat + row[a]] : 0.f;
  }
  const int jlast = min(T - 1, i0 + 63 + delay);
  for (int j0 = 0; j0 <= jlast; j0 += 64) {
    __syncthreads();
    load_split(Kh, Kl, base + 256, j0, T, 768 * S_in, 1.f);
    load_split_t(Kth, Ktl, base + 256, j0, T, 768 * S_in, 1.f);
    load_split(Vh, Vl, base + 512, j0, T, 768 * S_in, 1.f);
    __syncthreads();
    float s[8][4] = {}, dp[8][4] = {};
    warp_mm_nt(Qh, Ql, r0, Kh, Kl, g, t, s);
    warp_mm_nt(dOh, dOl, r0, Vh, Vl, g, t, dp);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int a = e >> 1, col = j0 + nt * 8 + 2 * t + (e & 1);
        const bool ok = row[a] < T && col < T && col <= row[a] + delay;
        const float p = ok ? expf(s[nt][e] - L[a]) : 0.f;
        const uint64_t rbase = (stat + row[a]) * T;
        s[nt][e] = p * (dp[nt][e] * keep_scale(drop, rbase, col) - D[a]) * scale;       // dS
      }
    float f[8][4] = {};
    warp_mm_reg_nt(s, Kth, Ktl, g, t, f);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[nt][e] += f[nt][e];
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    if (row[a] >= T) continue;
    float* op = dqkv + (r0g + static_cast<size_t>(row[a]) * S_in) * 768 + h * 64 + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<float2*>(op + nt * 8) = make_float2(dq[nt][2 * a] * ginv, dq[nt][2 * a + 1] * ginv);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dK, dV: grid (key tiles, heads, sequences); warp w owns key rows j0 + 16 w .. + 16; the score tile is computed
// transposed (S^T = K Q^T) so that P^T and dS^T come out as the A operands of the two accumulating products.
constexpr int kDkvSmem = 12 * TILE_H * 2 + 2 * 64 * 4;
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, const float* __restrict__ lse,
                    const float* __restrict__ dsum, float* __restrict__ dqkv, int T, int S_in, int delay, float scale,
                    const Dropout drop, const float* __restrict__ gsc) {
  extern __shared__ __align__(16) __half smh[];
  __half *Kh = smh, *Kl = Kh + TILE_H, *Vh = Kl + TILE_H, *Vl = Vh + TILE_H, *Qh = Vl + TILE_H, *Ql = Qh + TILE_H;
  __half *Qth = Ql + TILE_H, *Qtl = Qth + TILE_H, *dOh = Qtl + TILE_H, *dOl = dOh + TILE_H, *dOth = dOl + TILE_H, *dOtl = dOth + TILE_H;
  float* Ls = reinterpret_cast<float*>(dOtl + TILE_H);
  float* Ds = Ls + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int j0 = blockIdx.x * 64, h = blockIdx.y, n = blockIdx.z, r0 = warp * 16;
  // sequence n of an interleaved batch [n_seq / S_in][T][S_in][.]: frame t lives at row r0g + t * S_in (S_in = 1: plain [n][T])
  const size_t r0g = static_cast<size_t>(n / S_in) * T * S_in + (n % S_in);
  const float* base = qkv + r0g * 768 + h * 64;
  const float* dob = dout + r0g * 256 + h * 64;
  const size_t stat = (static_cast<size_t>(n) * kHeads + h) * T;
  const float gscale = __ldg(gsc), ginv = __ldg(gsc + 1);
  load_split(Kh, Kl, base + 256, j0, T, 768 * S_in, 1.f);
  load_split(Vh, Vl, base + 512, j0, T, 768 * S_in, 1.f);
  const int key[2] = {j0 + r0 + g, j0 + r0 + g + 8};
  float dk[8][4] = {}, dv[8][4] = {};
  for (int i0 = max(0, j0 - delay) / 64 * 64; i0 < T; i0 += 64) {
    __syncthreads();
    load_split(Qh, Ql, base, i0, T, 768 * S_in, scale);
    load_split_t(Qth, Qtl, base, i0, T, 768 * S_in, scale);
    load_split(dOh, dOl, dob, i0, T, 256 * S_in, gscale);
    load_split_t(dOth, dOtl, dob, i0, T, 256 * S_in, gscale);
    if (threadIdx.x < 64) {
      const int q = i0 + threadIdx.x;
      Ls[threadIdx.x] = q < T ? lse[stat + q] : 0.f;
      Ds[threadIdx.x] = q < T ? dsum[stat + q] * gscale : 0.f;
    }
    __syncthreads();
    float s[8][4] = {}, dp[8][4] = {};
    warp_mm_nt(Kh, Kl, r0, Qh, Ql, g, t, s);             // S^T[key][q]
    warp_mm_nt(Vh, Vl, r0, dOh, dOl, g, t, dp);          // dP^T[key][q] = V dO^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int a = e >> 1, ql = nt * 8 + 2 * t + (e & 1), q = i0 + ql;
        const bool ok = q < T && key[a] < T && key[a] <= q + delay;
        const float p = ok ? expf(s[nt][e] - Ls[ql]) : 0.f;
        const float ks = keep_scale(drop, (stat + q) * T, key[a]);
        s[nt][e] = p * ks;                                // P^T with dropout: dV's operand
        dp[nt][e] = p * (dp[nt][e] * ks - Ds[ql]);        // dS^T without the 64^-1/2 (q is pre-scaled): dK's operand
      }
    float f[8][4] = {};
    warp_mm_reg_nt(s, dOth, dOtl, g, t, f);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dv[nt][e] += f[nt][e];
        f[nt][e] = 0.f;
      }
    warp_mm_reg_nt(dp, Qth, Qtl, g, t, f);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dk[nt][e] += f[nt][e];
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    if (key[a] >= T) continue;
    float* op = dqkv + (r0g + static_cast<size_t>(key[a]) * S_in) * 768 + h * 64 + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<float2*>(op + 256 + nt * 8) = make_float2(dk[nt][2 * a] * ginv, dk[nt][2 * a + 1] * ginv);
      *reinterpret_cast<float2*>(op + 512 + nt * 8) = make_float2(dv[nt][2 * a] * ginv, dv[nt][2 * a + 1] * ginv);
    }
  }
}

}  // namespace tc
