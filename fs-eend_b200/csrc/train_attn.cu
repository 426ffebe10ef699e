// Causal multi-head self-attention, forward (with log-sum-exp) and backward, in fp32 on CUDA cores — the training
// counterpart of attn3.cu for the parity arithmetic (started).  Reference semantics:
// torch.nn.MultiheadAttention inside nn.TransformerEncoderLayer with the additive float mask of FS:model:152-155
// (key j visible to query i iff j <= i + mask_delay), 4 heads x 64, q scaled by 64^-1/2, fp32 softmax.
//
// qkv  fp32 [n_seq][T][768]  (q | k | v, head h at columns h*64 of each third)     out fp32 [n_seq][T][256]
// lse  fp32 [n_seq][4][T]    log-sum-exp of the scaled, masked scores of every query row
//
// 64 x 64 tiles in shared memory (row stride 68 floats), 256 threads, every thread owns a 4 x 4 register block with
// rows ty + 16 i and columns tx + 16 j (interleaved: shared-memory reads are conflict-free and row reductions stay
// inside a half-warp).  Backward is the standard recomputation form in two kernels (no atomics, fixed summation order):
//   dq kernel  per query tile:  D = rowsum(dO * O);  for key tiles: P = exp(S - lse); dS = P (dO V^T - D) scale; dQ += dS K
//   dkv kernel per key tile:    for query tiles:     dV += P^T dO;  dK += dS^T Q
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <stdint.h>

#include <algorithm>
#include <stdexcept>
#include <string>

#include "../../include/fseend_b200.h"
#include "once.h"
#include "p32.cuh"
#include "tmap.h"

namespace fseend {
namespace {

constexpr int kLD = 68;
constexpr int kTile = 64 * kLD;      // floats per tile
constexpr int kHeads = 4;

__device__ __forceinline__ void load_tile(float* dst, const float* src, int row0, int T, int ld) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = threadIdx.x + 256 * k, r = idx >> 4, c = (idx & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < T) v = *reinterpret_cast<const float4*>(src + static_cast<size_t>(row0 + r) * ld + c);
    *reinterpret_cast<float4*>(dst + r * kLD + c) = v;
  }
}

// acc[i][j] += sum_k A[ty + 16 i][k] * B[tx + 16 j][k]
__device__ __forceinline__ void mm_nt(const float* A, const float* B, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll 4
  for (int k = 0; k < 64; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty + 16 * i) * kLD + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(B + (tx + 16 * j) * kLD + k);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
      }
  }
}
// acc[i][j] += sum_k A[ty + 16 i][k] * B[k][tx + 16 j]
__device__ __forceinline__ void mm_nn(const float* A, const float* B, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll 4
  for (int k = 0; k < 64; k += 4) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty + 16 * i) * kLD + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = B[(k + kk) * kLD + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av, b[j], acc[i][j]);
      }
    }
  }
}
// acc[i][j] += sum_k A[k][ty + 16 i] * B[k][tx + 16 j]
__device__ __forceinline__ void mm_tn(const float* A, const float* B, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll 8
  for (int k = 0; k < 64; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = A[k * kLD + ty + 16 * i];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = B[k * kLD + tx + 16 * j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}
// Attention-probability dropout (nn.MultiheadAttention's `dropout`, applied to the softmax output): a counter-based
// hash of (seed, sequence, head, query, key) decides every element, so forward and backward regenerate the same mask
// with no stored state.  keep_scale() returns 0 (dropped) or 1 / (1 - p).  tests/test_train_ops_gpu.py carries the same
// hash in torch integer arithmetic.
struct Dropout {
  uint32_t thresh;      // drop iff hash < thresh;  thresh = p * 2^32 (0: no dropout)
  uint32_t seed_lo, seed_hi;
  float inv_keep;       // 1 / (1 - p)
};
__device__ __forceinline__ uint32_t drop_hash(uint32_t seed_lo, uint32_t seed_hi, uint64_t idx) {
  uint32_t x = static_cast<uint32_t>(idx) ^ seed_lo;
  x *= 0x9E3779B1u;
  x ^= x >> 15;
  x += static_cast<uint32_t>(idx >> 32) * 0x85EBCA77u + seed_hi;
  x *= 0xC2B2AE3Du;
  x ^= x >> 13;
  x *= 0x27D4EB2Fu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float keep_scale(const Dropout& d, uint64_t row_base, int col) {
  if (d.thresh == 0u) return 1.f;
  return drop_hash(d.seed_lo, d.seed_hi, row_base + static_cast<uint64_t>(col)) < d.thresh ? 0.f : d.inv_keep;
}

__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
train_attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ lse, int T, int S_in, int delay,
                      float scale, const Dropout drop) {
  extern __shared__ __align__(16) float sm[];
  float *Qs = sm, *Ks = sm + kTile, *Vs = sm + 2 * kTile, *Ps = sm + 3 * kTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.x * 64, h = blockIdx.y, n = blockIdx.z;
  const size_t r0g = static_cast<size_t>(n / S_in) * T * S_in + (n % S_in);      // see train_attn_tc.cuh
  const float* base = qkv + r0g * 768 + h * 64;
  load_tile(Qs, base, i0, T, 768 * S_in);
  float o[4][4] = {}, m[4], l[4] = {};
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = -INFINITY;
  const int jlast = min(T - 1, i0 + 63 + delay);
  for (int j0 = 0; j0 <= jlast; j0 += 64) {
    __syncthreads();
    load_tile(Ks, base + 256, j0, T, 768 * S_in);
    load_tile(Vs, base + 512, j0, T, 768 * S_in);
    __syncthreads();
    float s[4][4] = {};
    mm_nt(Qs, Ks, ty, tx, s);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = i0 + ty + 16 * i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = j0 + tx + 16 * j;
        s[i][j] = (col < T && col <= row + delay) ? s[i][j] * scale : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
      mx = half_warp_max(mx);
      const float mnew = fmaxf(m[i], mx);          // finite from the first key tile on (key 0 is visible to every row)
      const float corr = expf(m[i] - mnew);
      const uint64_t rbase = ((static_cast<uint64_t>(n) * kHeads + h) * T + row) * T;
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = expf(s[i][j] - mnew);
        rs += p;                                    // the softmax normaliser sums the undropped probabilities
        Ps[(ty + 16 * i) * kLD + tx + 16 * j] = p * keep_scale(drop, rbase, j0 + tx + 16 * j);
        o[i][j] *= corr;
      }
      l[i] = l[i] * corr + half_warp_sum(rs);
      m[i] = mnew;
    }
    __syncthreads();
    mm_nn(Ps, Vs, ty, tx, o);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i0 + ty + 16 * i;
    if (row >= T) continue;
    const float inv = 1.f / l[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(r0g + static_cast<size_t>(row) * S_in) * 256 + h * 64 + tx + 16 * j] = o[i][j] * inv;
    if (tx == 0) lse[(static_cast<size_t>(n) * kHeads + h) * T + row] = m[i] + logf(l[i]);
  }
}

__global__ void __launch_bounds__(256)
train_attn_bwd_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ dout,
                         const float* __restrict__ lse, float* __restrict__ dqkv, float* __restrict__ dsum, int T, int S_in,
                         int delay, float scale, const Dropout drop) {
  extern __shared__ __align__(16) float sm[];
  float *Qs = sm, *dOs = sm + kTile, *Ks = sm + 2 * kTile, *Vs = sm + 3 * kTile, *Ps = sm + 4 * kTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.x * 64, h = blockIdx.y, n = blockIdx.z;
  const size_t r0g = static_cast<size_t>(n / S_in) * T * S_in + (n % S_in);      // see train_attn_tc.cuh
  const float* base = qkv + r0g * 768 + h * 64;
  load_tile(Qs, base, i0, T, 768 * S_in);
  load_tile(dOs, dout + r0g * 256 + h * 64, i0, T, 256 * S_in);
  __syncthreads();
  float D[4], L[4], dq[4][4] = {};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i0 + ty + 16 * i;
    float d = 0.f;
    if (row < T) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        d = fmaf(out[(r0g + static_cast<size_t>(row) * S_in) * 256 + h * 64 + tx + 16 * j], dOs[(ty + 16 * i) * kLD + tx + 16 * j], d);
    }
    D[i] = half_warp_sum(d);
    L[i] = row < T ? lse[(static_cast<size_t>(n) * kHeads + h) * T + row] : 0.f;
    if (tx == 0 && row < T) dsum[(static_cast<size_t>(n) * kHeads + h) * T + row] = D[i];
  }
  const int jlast = min(T - 1, i0 + 63 + delay);
  for (int j0 = 0; j0 <= jlast; j0 += 64) {
    __syncthreads();
    load_tile(Ks, base + 256, j0, T, 768 * S_in);
    load_tile(Vs, base + 512, j0, T, 768 * S_in);
    __syncthreads();
    float s[4][4] = {}, dp[4][4] = {};
    mm_nt(Qs, Ks, ty, tx, s);
    mm_nt(dOs, Vs, ty, tx, dp);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = i0 + ty + 16 * i;
      const uint64_t rbase = ((static_cast<uint64_t>(n) * kHeads + h) * T + row) * T;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = j0 + tx + 16 * j;
        const bool ok = row < T && col < T && col <= row + delay;
        const float p = ok ? expf(s[i][j] * scale - L[i]) : 0.f;
        Ps[(ty + 16 * i) * kLD + tx + 16 * j] = p * (dp[i][j] * keep_scale(drop, rbase, col) - D[i]) * scale;
      }
    }
    __syncthreads();
    mm_nn(Ps, Ks, ty, tx, dq);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i0 + ty + 16 * i;
    if (row >= T) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) dqkv[(r0g + static_cast<size_t>(row) * S_in) * 768 + h * 64 + tx + 16 * j] = dq[i][j];
  }
}

__global__ void __launch_bounds__(256)
train_attn_bwd_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, const float* __restrict__ lse,
                          const float* __restrict__ dsum, float* __restrict__ dqkv, int T, int S_in, int delay, float scale,
                          const Dropout drop) {
  extern __shared__ __align__(16) float sm[];
  float *Ks = sm, *Vs = sm + kTile, *Qs = sm + 2 * kTile, *dOs = sm + 3 * kTile, *Ps = sm + 4 * kTile, *dSs = sm + 5 * kTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int j0 = blockIdx.x * 64, h = blockIdx.y, n = blockIdx.z;
  const size_t r0g = static_cast<size_t>(n / S_in) * T * S_in + (n % S_in);      // see train_attn_tc.cuh
  const float* base = qkv + r0g * 768 + h * 64;
  load_tile(Ks, base + 256, j0, T, 768 * S_in);
  load_tile(Vs, base + 512, j0, T, 768 * S_in);
  float dk[4][4] = {}, dv[4][4] = {};
  const size_t stat = (static_cast<size_t>(n) * kHeads + h) * T;
  for (int i0 = max(0, j0 - delay) / 64 * 64; i0 < T; i0 += 64) {
    __syncthreads();
    load_tile(Qs, base, i0, T, 768 * S_in);
    load_tile(dOs, dout + r0g * 256 + h * 64, i0, T, 256 * S_in);
    __syncthreads();
    float s[4][4] = {}, dp[4][4] = {};
    mm_nt(Qs, Ks, ty, tx, s);
    mm_nt(dOs, Vs, ty, tx, dp);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = i0 + ty + 16 * i;
      const float L = row < T ? lse[stat + row] : 0.f, D = row < T ? dsum[stat + row] : 0.f;
      const uint64_t rbase = ((static_cast<uint64_t>(n) * kHeads + h) * T + row) * T;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = j0 + tx + 16 * j;
        const bool ok = row < T && col < T && col <= row + delay;
        const float p = ok ? expf(s[i][j] * scale - L) : 0.f;
        const float ks = keep_scale(drop, rbase, col);
        Ps[(ty + 16 * i) * kLD + tx + 16 * j] = p * ks;
        dSs[(ty + 16 * i) * kLD + tx + 16 * j] = p * (dp[i][j] * ks - D) * scale;
      }
    }
    __syncthreads();
    mm_tn(Ps, dOs, ty, tx, dv);
    mm_tn(dSs, Qs, ty, tx, dk);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int key = j0 + ty + 16 * i;
    if (key >= T) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t o = (r0g + static_cast<size_t>(key) * S_in) * 768 + h * 64 + tx + 16 * j;
      dqkv[o + 256] = dk[i][j];
      dqkv[o + 512] = dv[i][j];
    }
  }
}

#include "train_attn_tc.cuh"

// ---------------------------------------------------------------------------------------------------------------------
// Speaker-axis attention backward (FS:fusion:390 self_attn2: S x S attention per frame, no mask).  One warp per
// (frame, head): q / k / v / dO rows [S][64] staged in shared memory (row stride 65), the S x S probabilities and score
// gradients recomputed there; lanes own (query, key) pairs for the scores and head dims (lane, lane + 32) for the products.
constexpr int kSpkMaxS = 16;
constexpr int kSpkLd = 65;
constexpr int kSpkWarpFloats = 4 * kSpkMaxS * kSpkLd + 2 * kSpkMaxS * kSpkMaxS;      // q k v dO + P dS at S = 16
inline int spk_bwd_smem(int S) { return 4 * (4 * S * kSpkLd + 2 * S * S) * 4; }

__global__ void __launch_bounds__(128)
train_spk_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, float* __restrict__ dqkv,
                          int n_frames, int S, float scale, const Dropout drop) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + warp;                  // (frame, head), heads fastest
  if (item >= n_frames * kHeads) return;
  const int f = item >> 2, h = item & 3;
  float* q = sm + warp * (4 * S * kSpkLd + 2 * S * S);       // sized by the actual S (host: spk_bwd_smem)
  float *k = q + S * kSpkLd, *v = k + S * kSpkLd, *dO = v + S * kSpkLd;
  float *P = dO + S * kSpkLd, *dS = P + S * S;
  const float* base = qkv + static_cast<size_t>(f) * S * 768 + h * 64;
  const float* dbase = dout + static_cast<size_t>(f) * S * 256 + h * 64;
  for (int i = 0; i < S; ++i) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int d = lane + 32 * hh;
      q[i * kSpkLd + d] = base[static_cast<size_t>(i) * 768 + d];
      k[i * kSpkLd + d] = base[static_cast<size_t>(i) * 768 + 256 + d];
      v[i * kSpkLd + d] = base[static_cast<size_t>(i) * 768 + 512 + d];
      dO[i * kSpkLd + d] = dbase[static_cast<size_t>(i) * 256 + d];
    }
  }
  __syncwarp();
  for (int idx = lane; idx < S * S; idx += 32) {
    const int i = idx / S, j = idx - i * S;
    float sc = 0.f, dp = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      sc = fmaf(q[i * kSpkLd + d], k[j * kSpkLd + d], sc);
      dp = fmaf(dO[i * kSpkLd + d], v[j * kSpkLd + d], dp);
    }
    P[i * S + j] = sc * scale;
    dS[i * S + j] = dp;
  }
  __syncwarp();
  if (lane < S) {                                          // one query row per lane
    const int i = lane;
    float mx = -INFINITY;
    for (int j = 0; j < S; ++j) mx = fmaxf(mx, P[i * S + j]);
    float sum = 0.f;
    for (int j = 0; j < S; ++j) {
      const float e = expf(P[i * S + j] - mx);
      P[i * S + j] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    const uint64_t rbase = ((static_cast<uint64_t>(f) * kHeads + h) * S + i) * S;
    float D = 0.f;
    for (int j = 0; j < S; ++j) {
      const float pj = P[i * S + j] * inv;
      P[i * S + j] = pj;
      dS[i * S + j] *= keep_scale(drop, rbase, j);        // gradient w.r.t. the undropped probability
      D = fmaf(pj, dS[i * S + j], D);
    }
    for (int j = 0; j < S; ++j) {
      dS[i * S + j] = P[i * S + j] * (dS[i * S + j] - D) * scale;
      P[i * S + j] *= keep_scale(drop, rbase, j);         // dV uses the dropped probabilities
    }
  }
  __syncwarp();
  float* obase = dqkv + static_cast<size_t>(f) * S * 768 + h * 64;
  for (int a = 0; a < S; ++a) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int d = lane + 32 * hh;
      float dq = 0.f, dk = 0.f, dv = 0.f;
      for (int b = 0; b < S; ++b) {
        dq = fmaf(dS[a * S + b], k[b * kSpkLd + d], dq);     // a = query
        dk = fmaf(dS[b * S + a], q[b * kSpkLd + d], dk);     // a = key
        dv = fmaf(P[b * S + a], dO[b * kSpkLd + d], dv);
      }
      obase[static_cast<size_t>(a) * 768 + d] = dq;
      obase[static_cast<size_t>(a) * 768 + 256 + d] = dk;
      obase[static_cast<size_t>(a) * 768 + 512 + d] = dv;
    }
  }
}

// Speaker-axis attention forward WITH dropout (the inference kernel p32_spk_attn_kernel serves p = 0): same staging as
// the backward kernel, one warp per (frame, head).
__global__ void __launch_bounds__(128)
train_spk_attn_fwd_drop_kernel(const float* __restrict__ qkv, float* __restrict__ out, int n_frames, int S, float scale,
                               const Dropout drop) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + warp;
  if (item >= n_frames * kHeads) return;
  const int f = item >> 2, h = item & 3;
  float* q = sm + warp * (3 * S * kSpkLd + S * S);
  float *k = q + S * kSpkLd, *v = k + S * kSpkLd, *P = v + S * kSpkLd;
  const float* base = qkv + static_cast<size_t>(f) * S * 768 + h * 64;
  for (int i = 0; i < S; ++i) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int d = lane + 32 * hh;
      q[i * kSpkLd + d] = base[static_cast<size_t>(i) * 768 + d];
      k[i * kSpkLd + d] = base[static_cast<size_t>(i) * 768 + 256 + d];
      v[i * kSpkLd + d] = base[static_cast<size_t>(i) * 768 + 512 + d];
    }
  }
  __syncwarp();
  for (int idx = lane; idx < S * S; idx += 32) {
    const int i = idx / S, j = idx - i * S;
    float sc = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) sc = fmaf(q[i * kSpkLd + d], k[j * kSpkLd + d], sc);
    P[i * S + j] = sc * scale;
  }
  __syncwarp();
  if (lane < S) {
    const int i = lane;
    const uint64_t rbase = ((static_cast<uint64_t>(f) * kHeads + h) * S + i) * S;
    float mx = -INFINITY;
    for (int j = 0; j < S; ++j) mx = fmaxf(mx, P[i * S + j]);
    float sum = 0.f;
    for (int j = 0; j < S; ++j) {
      const float e = expf(P[i * S + j] - mx);
      P[i * S + j] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    for (int j = 0; j < S; ++j) P[i * S + j] *= inv * keep_scale(drop, rbase, j);
  }
  __syncwarp();
  float* obase = out + static_cast<size_t>(f) * S * 256 + h * 64;
  for (int a = 0; a < S; ++a) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int d = lane + 32 * hh;
      float o = 0.f;
      for (int b = 0; b < S; ++b) o = fmaf(P[a * S + b], v[b * kSpkLd + d], o);
      obase[static_cast<size_t>(a) * 256 + d] = o;
    }
  }
}

template <class F>
int aguard(F&& f) {
  try {
    f();
    return FSEEND_OK;
  } catch (const std::invalid_argument& e) {
    set_last_error(e.what());
    return FSEEND_ERR_INVALID;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return FSEEND_ERR_CUDA;
  }
}
void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
PerDeviceOnce g_attr_once;
void set_attrs() {
  if (g_attr_once.first()) {
    cudaFuncSetAttribute(train_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kTile * 4);
    cudaFuncSetAttribute(train_attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * kTile * 4);
    cudaFuncSetAttribute(train_attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * kTile * 4);
    cudaFuncSetAttribute(tc::attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kFwdSmem);
    cudaFuncSetAttribute(tc::attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kDqSmem);
    cudaFuncSetAttribute(tc::attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kDkvSmem);
    cudaFuncSetAttribute(train_spk_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kSpkWarpFloats * 4);
    cudaFuncSetAttribute(train_spk_attn_fwd_drop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kSpkWarpFloats * 4);
  }
}
// max |x| -> power-of-two scale (the gradient-scaling rule of train_ops.cu, target [16, 32): dP = dO V^T must stay inside
// the fp16 range after the multiplication)
__global__ void __launch_bounds__(256) attn_absmax_kernel(const float* __restrict__ x, size_t n4, unsigned int* __restrict__ slot) {
  float m = 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(slot, __float_as_uint(m));
}
__global__ void attn_make_scale_kernel(const unsigned int* __restrict__ slot, float* __restrict__ sc) {
  const float m = __uint_as_float(*slot);
  float s = 1.f;
  if (m > 0.f && m < 3e38f) s = exp2f(static_cast<float>(max(-100, min(100, 4 - ilogbf(m)))));
  sc[0] = s;
  sc[1] = 1.f / s;
}

// FSEEND_TRAIN_ATTN=0 selects the CUDA-core kernels (read per call: the tests run both)
bool use_tensor_cores() {
  const char* e = getenv("FSEEND_TRAIN_ATTN");
  return !(e && e[0] == '0');
}
Dropout make_dropout(float p, unsigned long long seed) {
  if (!(p >= 0.f && p < 1.f)) throw std::invalid_argument("dropout probability must be in [0, 1)");
  Dropout d;
  d.thresh = p > 0.f ? static_cast<uint32_t>(std::min(4294967295.0, static_cast<double>(p) * 4294967296.0)) : 0u;
  d.seed_lo = static_cast<uint32_t>(seed);
  d.seed_hi = static_cast<uint32_t>(seed >> 32);
  d.inv_keep = 1.f / (1.f - p);
  return d;
}

}  // namespace
}  // namespace fseend

using namespace fseend;

extern "C" {

int fseend_train_attn_fwd(const float* qkv, int n_seq, int T, int seq_inner, int mask_delay, float dropout_p,
                          unsigned long long seed, float* out, float* lse, void* stream) {
  return aguard([&] {
    if (!qkv || !out || !lse || n_seq < 1 || T < 1 || mask_delay < 0 || n_seq > 65535 || seq_inner < 1 || n_seq % seq_inner)
      throw std::invalid_argument("train_attn_fwd: bad arguments");
    set_attrs();
    const dim3 grid((T + 63) / 64, kHeads, n_seq);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (use_tensor_cores())
      tc::attn_fwd_kernel<<<grid, 128, tc::kFwdSmem, st>>>(qkv, out, lse, T, seq_inner, mask_delay, 0.125f, make_dropout(dropout_p, seed));
    else
      train_attn_fwd_kernel<<<grid, 256, 4 * kTile * 4, st>>>(qkv, out, lse, T, seq_inner, mask_delay, 0.125f, make_dropout(dropout_p, seed));
    check_launch("train_attn_fwd");
  });
}

// dsum: scratch fp32 [n_seq][4][T] + 16 floats (row sums of dO * O; the gradient-scale slot behind them)
int fseend_train_attn_bwd(const float* qkv, const float* out, const float* dout, const float* lse, int n_seq, int T,
                          int seq_inner, int mask_delay, float dropout_p, unsigned long long seed, float* dqkv, float* dsum,
                          void* stream) {
  return aguard([&] {
    if (!qkv || !out || !dout || !lse || !dqkv || !dsum || n_seq < 1 || T < 1 || mask_delay < 0 || n_seq > 65535 ||
        seq_inner < 1 || n_seq % seq_inner)
      throw std::invalid_argument("train_attn_bwd: bad arguments");
    set_attrs();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid((T + 63) / 64, kHeads, n_seq);
    const Dropout drop = make_dropout(dropout_p, seed);
    if (use_tensor_cores()) {
      unsigned int* slot = reinterpret_cast<unsigned int*>(dsum + static_cast<size_t>(n_seq) * kHeads * T);
      float* gsc = reinterpret_cast<float*>(slot) + 4;
      if (cudaMemsetAsync(slot, 0, 4, st) != cudaSuccess) throw std::runtime_error("train_attn_bwd: memset failed");
      attn_absmax_kernel<<<592, 256, 0, st>>>(dout, static_cast<size_t>(n_seq) * T * 64, slot);
      attn_make_scale_kernel<<<1, 1, 0, st>>>(slot, gsc);
      tc::attn_bwd_dq_kernel<<<grid, 128, tc::kDqSmem, st>>>(qkv, out, dout, lse, dqkv, dsum, T, seq_inner, mask_delay, 0.125f, drop, gsc);
      tc::attn_bwd_dkv_kernel<<<grid, 128, tc::kDkvSmem, st>>>(qkv, dout, lse, dsum, dqkv, T, seq_inner, mask_delay, 0.125f, drop, gsc);
    } else {
      train_attn_bwd_dq_kernel<<<grid, 256, 5 * kTile * 4, st>>>(qkv, out, dout, lse, dqkv, dsum, T, seq_inner, mask_delay, 0.125f, drop);
      train_attn_bwd_dkv_kernel<<<grid, 256, 6 * kTile * 4, st>>>(qkv, dout, lse, dsum, dqkv, T, seq_inner, mask_delay, 0.125f, drop);
    }
    check_launch("train_attn_bwd");
  });
}

// Speaker-axis attention (no mask) on projected qkv fp32 [n_frames][S][768] -> out fp32 [n_frames][S][256], S <= 16.
int fseend_train_spk_attn_fwd(const float* qkv, int n_frames, int S, float dropout_p, unsigned long long seed, float* out,
                              void* stream) {
  return aguard([&] {
    if (!qkv || !out || n_frames < 1 || S < 1 || S > kSpkMaxS) throw std::invalid_argument("train_spk_attn_fwd: bad arguments");
    if (dropout_p > 0.f) {
      set_attrs();
      const int items = n_frames * kHeads;
      train_spk_attn_fwd_drop_kernel<<<(items + 3) / 4, 128, 4 * (3 * S * kSpkLd + S * S) * 4, static_cast<cudaStream_t>(stream)>>>(
          qkv, out, n_frames, S, 0.125f, make_dropout(dropout_p, seed));
      check_launch("train_spk_attn_fwd");
      return;
    }
    if (launch_p32_spk_attn(qkv, out, n_frames, S, 0.125f, static_cast<cudaStream_t>(stream)) != 0)
      throw std::invalid_argument("train_spk_attn_fwd: S out of range");
    check_launch("train_spk_attn_fwd");
  });
}
int fseend_train_spk_attn_bwd(const float* qkv, const float* dout, int n_frames, int S, float dropout_p,
                              unsigned long long seed, float* dqkv, void* stream) {
  return aguard([&] {
    if (!qkv || !dout || !dqkv || n_frames < 1 || S < 1 || S > kSpkMaxS) throw std::invalid_argument("train_spk_attn_bwd: bad arguments");
    set_attrs();
    const int items = n_frames * kHeads;
    train_spk_attn_bwd_kernel<<<(items + 3) / 4, 128, spk_bwd_smem(S), static_cast<cudaStream_t>(stream)>>>(
        qkv, dout, dqkv, n_frames, S, 0.125f, make_dropout(dropout_p, seed));
    check_launch("train_spk_attn_bwd");
  });
}

}  // extern "C"
