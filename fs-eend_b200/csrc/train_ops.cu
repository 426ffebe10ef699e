// Training-step building blocks (STARTED — not a training step yet): forward + backward of the ops
// that carry ~90 % of the step's FLOPs, in the parity ("p32") arithmetic: fp32 tensors in HBM, every matrix product on
// tcgen05 with split-precision operands (p32.cuh), reductions in fixed order.  Behind fseend_train_* C-ABI exports and
// torch.autograd.Functions (fseend_b200/autograd.py); pinned against torch autograd in tests/test_train_ops_gpu.py.
//
//   Linear  y = act(x W^T + b)           reference call sites: every nn.Linear of FS:model / FS:fusion
//     dX = dY W            p32 GEMM, A = dY [rows][N], planes = W^T [K][N]
//     dW = dY^T X          split-K p32 GEMM over the rows: the LARGER of dY / X is the A operand, read through its transpose
//                          in place (no copy); the smaller one is transposed once into fp16 hi / lo planes laid out per
//                          K-slice; the slices' partial products are summed in fixed order by a reduce kernel
//     db = column sums of dY (two-stage, fixed order)
//   Gradients are ~1e-6 in magnitude (the loss is a mean over B*T frames) — fp16-subnormal — so every backward product
//   first takes max|dY| (one pass) and multiplies dY by a power of two that brings the maximum to [512, 1024); the
//   epilogue multiplies by the inverse.  The ReLU mask (y > 0) is applied while dY is read, never materialised.
//   LayerNorm (biased variance, eps inside the sqrt): dx per row, dgamma / dbeta as two-stage column sums
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <stdexcept>
#include <string>

#include "../../include/fseend_b200.h"
#include "p32.cuh"
#include "tmap.h"

namespace fseend {
namespace {

constexpr float kWScale = 64.f;           // weights are multiplied by 2^6 before the fp16 split (lo stays a normal number)

#define TCHECK(expr)                                                                              \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      throw std::runtime_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));        \
  } while (0)

template <class F>
int tguard(F&& f) {
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

__device__ __forceinline__ void split1(float x, __half& h, __half& l) {
  h = __float2half_rn(x);
  l = __float2half_rn(x - __half2float(h));
}

// src fp32 [R][C] -> hi / lo fp16 planes [R][Cp] (columns >= C zero), values multiplied by scale * (*scale_dev)
__global__ void split_planes_kernel(const float* __restrict__ src, int R, int C, int Cp, float scale,
                                    const float* __restrict__ scale_dev, __half* __restrict__ hi, __half* __restrict__ lo) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(R) * Cp) return;
  const int r = static_cast<int>(i / Cp), c = static_cast<int>(i % Cp);
  const float sc = scale * (scale_dev ? __ldg(scale_dev) : 1.f);
  __half h, l;
  split1(c < C ? src[static_cast<size_t>(r) * C + c] * sc : 0.f, h, l);
  hi[i] = h;
  lo[i] = l;
}
// the same for C == Cp, C % 4 == 0: four elements per thread
__global__ void __launch_bounds__(256)
split_planes4_kernel(const float* __restrict__ src, size_t n4, float scale, const float* __restrict__ scale_dev,
                     __half* __restrict__ hi, __half* __restrict__ lo) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float sc = scale * (scale_dev ? __ldg(scale_dev) : 1.f);
  const float4 v = reinterpret_cast<const float4*>(src)[i];
  __half h[4], l[4];
  split1(v.x * sc, h[0], l[0]);
  split1(v.y * sc, h[1], l[1]);
  split1(v.z * sc, h[2], l[2]);
  split1(v.w * sc, h[3], l[3]);
  reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<const uint2*>(h);
  reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<const uint2*>(l);
}

// src fp32 [R][C] -> fp16 hi / lo planes of its transpose, laid out per K-slice: plane[(r / ks) * Cp + c][r % ks] for
// r < n_slices * ks, c < Cp (zero where r >= R or c >= C); values are multiplied by scale_host * (*scale_dev) and read as
// zero where mask[r][c] <= 0.  32 x 32 smem tiles; ks % 32 == 0 so a tile never straddles two slices.
__global__ void __launch_bounds__(256)
transpose_split_kernel(const float* __restrict__ src, const float* __restrict__ mask, int R, int C, int ks, int Cp,
                       float scale_host, const float* __restrict__ scale_dev, __half* __restrict__ hi,
                       __half* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  const float scale = scale_host * (scale_dev ? __ldg(scale_dev) : 1.f);
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      const size_t o = static_cast<size_t>(r) * C + c;
      v = src[o];
      if (mask && mask[o] <= 0.f) v = 0.f;
    }
    tile[j][tx] = v * scale;
  }
  __syncthreads();
  const int slice = r0 / ks, rin0 = r0 - slice * ks;
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;                                        // output row = source column
    if (c < Cp) {
      const size_t o = (static_cast<size_t>(slice) * Cp + c) * ks + rin0 + tx;
      __half h, l;
      split1(tile[tx][j], h, l);
      hi[o] = h;
      lo[o] = l;
    }
  }
}

// max |x| over n floats into *slot (float bits of a non-negative value order like unsigned integers); slot zeroed before
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, size_t n, unsigned int* __restrict__ slot) {
  float m = 0.f;
  const size_t n4 = n / 4;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = fmaxf(m, fabsf(x[n4 * 4 + threadIdx.x]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, wm[w]);
    if (m > 0.f) atomicMax(slot, __float_as_uint(m));            // NaN / inf propagate as "huge": scale falls back to 1
  }
}
// sc[0] = 2^e with max * 2^e in [512, 1024), sc[1] = 2^-e  (1, 1 for an all-zero or non-finite tensor)
__global__ void make_scale_kernel(const unsigned int* __restrict__ slot, float* __restrict__ sc) {
  const float m = __uint_as_float(*slot);
  float s = 1.f;
  if (m > 0.f && m < 3e38f) {
    int e = 9 - ilogbf(m);
    e = max(-100, min(100, e));
    s = exp2f(static_cast<float>(e));
  }
  sc[0] = s;
  sc[1] = 1.f / s;
}

// dw[n][k] = sum over slices of partial[s][n][k]   (k < K; partial rows are n_out floats)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int n_slices, int M, int n_out, int K, float* __restrict__ dw) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(M) * K) return;
  const int n = static_cast<int>(i / K), k = static_cast<int>(i % K);
  float s = 0.f;
  for (int sl = 0; sl < n_slices; ++sl) s += partial[(static_cast<size_t>(sl) * M + n) * n_out + k];
  dw[i] = s;
}
// the swapped product: partial[s][k][n] (M = K rows of N floats) -> dw[n][k]
__global__ void __launch_bounds__(256)
splitk_reduce_t_kernel(const float* __restrict__ partial, int n_slices, int K, int N, float* __restrict__ dw) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int k = k0 + j, n = n0 + tx;
    float s = 0.f;
    if (k < K && n < N)
      for (int sl = 0; sl < n_slices; ++sl) s += partial[(static_cast<size_t>(sl) * K + k) * N + n];
    tile[j][tx] = s;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int n = n0 + j, k = k0 + tx;
    if (n < N && k < K) dw[static_cast<size_t>(n) * K + k] = tile[tx][j];
  }
}

// partial[blk][c] = sum over the block's rows of f(row, c); then out[c] = sum over blocks (fixed order)
constexpr int kRowsPerBlk = 256;
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mask, int R, int C,
                      float* __restrict__ partial) {
  // a [R][C]; b optional [R][C]: sums a * b (LayerNorm dgamma) or a alone; mask optional: rows' elements with mask <= 0 skipped
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= C) return;
  const int r0 = blockIdx.x * kRowsPerBlk, r1 = min(R, r0 + kRowsPerBlk);
  float s4[4] = {0.f, 0.f, 0.f, 0.f};        // four independent chains (rows r, r+1, r+2, r+3), combined in fixed order
  auto term = [&](int r) {
    float v = a[static_cast<size_t>(r) * C + c];
    if (mask && mask[static_cast<size_t>(r) * C + c] <= 0.f) v = 0.f;
    return b ? v * b[static_cast<size_t>(r) * C + c] : v;
  };
  int r = r0;
  for (; r + 4 <= r1; r += 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) s4[j] += term(r + j);
  }
  for (; r < r1; ++r) s4[0] += term(r);
  partial[static_cast<size_t>(blockIdx.x) * C + c] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
}
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, int n_blk, int C, float* __restrict__ out) {
  // 32 columns per block, 8 interleaved groups of partial rows per column, combined in fixed order (deterministic)
  __shared__ double sm[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (c < C)
    for (int k = grp; k < n_blk; k += 8) s += static_cast<double>(partial[static_cast<size_t>(k) * C + c]);
  sm[grp][lane] = s;
  __syncthreads();
  if (grp == 0 && c < C) {
    for (int j = 1; j < 8; ++j) s += sm[j][lane];
    out[c] = static_cast<float>(s);
  }
}

// out = dy where y > 0 else 0   (ReLU backward on the saved output; only the stand-alone Linear(+ReLU) backward needs it —
// the FFN pair folds the mask into the epilogue of the product that creates dY, see relu_input below)
__global__ void relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, size_t n4, float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 a = reinterpret_cast<const float4*>(y)[i], d = reinterpret_cast<const float4*>(dy)[i];
  reinterpret_cast<float4*>(out)[i] = make_float4(a.x > 0.f ? d.x : 0.f, a.y > 0.f ? d.y : 0.f, a.z > 0.f ? d.z : 0.f, a.w > 0.f ? d.w : 0.f);
}

// LayerNorm backward, one warp per row of 256: dx = rstd * (dy g - mean(dy g) - xhat mean(dy g xhat)); also writes
// xhat (for dgamma = colsum(dy * xhat)).
__global__ void __launch_bounds__(256)
ln_bwd_row_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ dy, int rows,
                  float eps, float* __restrict__ dx, float* __restrict__ xhat_out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xp = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * 256);
  const float4* dp = reinterpret_cast<const float4*>(dy + static_cast<size_t>(row) * 256);
  const float4 a = xp[lane], b = xp[32 + lane], da = dp[lane], db = dp[32 + lane];
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + lane), g1 = __ldg(reinterpret_cast<const float4*>(g) + 32 + lane);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  const float d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / 256.f);
  float m2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) m2 = fmaf(v[i] - mean, v[i] - mean, m2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
  const float rstd = 1.f / sqrtf(m2 * (1.f / 256.f) + eps);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = (v[i] - mean) * rstd;                 // xhat
    s1 += d[i] * gg[i];
    s2 = fmaf(d[i] * gg[i], v[i], s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  s1 *= (1.f / 256.f);
  s2 *= (1.f / 256.f);
  float o8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o8[i] = rstd * (d[i] * gg[i] - s1 - v[i] * s2);
  float4* op = reinterpret_cast<float4*>(dx + static_cast<size_t>(row) * 256);
  op[lane] = make_float4(o8[0], o8[1], o8[2], o8[3]);
  op[32 + lane] = make_float4(o8[4], o8[5], o8[6], o8[7]);
  float4* hp = reinterpret_cast<float4*>(xhat_out + static_cast<size_t>(row) * 256);
  hp[lane] = make_float4(v[0], v[1], v[2], v[3]);
  hp[32 + lane] = make_float4(v[4], v[5], v[6], v[7]);
}

// y = LayerNorm(x + r) (r optional; the sum is stored when sum_out != nullptr — the backward needs it), warp per row of 256
__global__ void __launch_bounds__(256)
ln_fwd_row_kernel(const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ g,
                  const float* __restrict__ b, int rows, float eps, float* __restrict__ sum_out, float* __restrict__ y) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const size_t off = static_cast<size_t>(row) * 64;      // in float4
  float4 a = reinterpret_cast<const float4*>(x)[off + lane], c = reinterpret_cast<const float4*>(x)[off + 32 + lane];
  if (r) {
    const float4 ra = reinterpret_cast<const float4*>(r)[off + lane], rc = reinterpret_cast<const float4*>(r)[off + 32 + lane];
    a.x += ra.x; a.y += ra.y; a.z += ra.z; a.w += ra.w;
    c.x += rc.x; c.y += rc.y; c.z += rc.z; c.w += rc.w;
  }
  if (sum_out) {
    reinterpret_cast<float4*>(sum_out)[off + lane] = a;
    reinterpret_cast<float4*>(sum_out)[off + 32 + lane] = c;
  }
  float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / 256.f);
  float m2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) m2 = fmaf(v[i] - mean, v[i] - mean, m2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
  const float rstd = 1.f / sqrtf(m2 * (1.f / 256.f) + eps);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + lane), g1 = __ldg(reinterpret_cast<const float4*>(g) + 32 + lane);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + lane), b1 = __ldg(reinterpret_cast<const float4*>(b) + 32 + lane);
  reinterpret_cast<float4*>(y)[off + lane] = make_float4((v[0] - mean) * rstd * g0.x + b0.x, (v[1] - mean) * rstd * g0.y + b0.y,
                                                         (v[2] - mean) * rstd * g0.z + b0.z, (v[3] - mean) * rstd * g0.w + b0.w);
  reinterpret_cast<float4*>(y)[off + 32 + lane] = make_float4((v[4] - mean) * rstd * g1.x + b1.x, (v[5] - mean) * rstd * g1.y + b1.y,
                                                              (v[6] - mean) * rstd * g1.z + b1.z, (v[7] - mean) * rstd * g1.w + b1.w);
}

// ---------------------------------------------------------------------------------------------------------------------
// L2 normalisation of rows of 256 (FS:model:41,43: emb / ||emb||, attractors / ||attractors||; no eps, as the reference)
// forward: y = x / ||x||, inv[row] = 1 / ||x||;  backward: dx = (dy - y (y . dy)) * inv
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ x, int rows, float* __restrict__ y, float* __restrict__ inv) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xp = reinterpret_cast<const float4*>(x) + static_cast<size_t>(row) * 64;
  const float4 a = xp[lane], b = xp[32 + lane];
  float s = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float r = 1.f / sqrtf(s);
  float4* yp = reinterpret_cast<float4*>(y) + static_cast<size_t>(row) * 64;
  yp[lane] = make_float4(a.x * r, a.y * r, a.z * r, a.w * r);
  yp[32 + lane] = make_float4(b.x * r, b.y * r, b.z * r, b.w * r);
  if (lane == 0) inv[row] = r;
}
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ inv, const float* __restrict__ dy, int rows,
                  float* __restrict__ dx) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* yp = reinterpret_cast<const float4*>(y) + static_cast<size_t>(row) * 64;
  const float4* dp = reinterpret_cast<const float4*>(dy) + static_cast<size_t>(row) * 64;
  const float4 a = yp[lane], b = yp[32 + lane], da = dp[lane], db = dp[32 + lane];
  float s = a.x * da.x + a.y * da.y + a.z * da.z + a.w * da.w + b.x * db.x + b.y * db.y + b.z * db.z + b.w * db.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float r = inv[row];
  float4* op = reinterpret_cast<float4*>(dx) + static_cast<size_t>(row) * 64;
  op[lane] = make_float4((da.x - a.x * s) * r, (da.y - a.y * s) * r, (da.z - a.z * s) * r, (da.w - a.w * s) * r);
  op[32 + lane] = make_float4((db.x - b.x * s) * r, (db.y - b.y * s) * r, (db.z - b.z * s) * r, (db.w - b.w * s) * r);
}

// Dot-product head (FS:model:60): y[f][s] = emb[f][:] . att[f][s][:], one warp per frame f.
__global__ void __launch_bounds__(256)
head_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ att, int frames, int S, float* __restrict__ y) {
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (f >= frames) return;
  const float4* ep = reinterpret_cast<const float4*>(emb) + static_cast<size_t>(f) * 64;
  const float4 a = ep[lane], b = ep[32 + lane];
  for (int s = 0; s < S; ++s) {
    const float4* ap = reinterpret_cast<const float4*>(att) + (static_cast<size_t>(f) * S + s) * 64;
    const float4 c = ap[lane], d = ap[32 + lane];
    float v = a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w + b.x * d.x + b.y * d.y + b.z * d.z + b.w * d.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) y[static_cast<size_t>(f) * S + s] = v;
  }
}
// demb[f][:] = sum_s dy[f][s] att[f][s][:];  datt[f][s][:] = dy[f][s] emb[f][:]
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ emb, const float* __restrict__ att, const float* __restrict__ dy, int frames, int S,
                float* __restrict__ demb, float* __restrict__ datt) {
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (f >= frames) return;
  const float4* ep = reinterpret_cast<const float4*>(emb) + static_cast<size_t>(f) * 64;
  const float4 a = ep[lane], b = ep[32 + lane];
  float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
  for (int s = 0; s < S; ++s) {
    const float g = dy[static_cast<size_t>(f) * S + s];
    const float4* ap = reinterpret_cast<const float4*>(att) + (static_cast<size_t>(f) * S + s) * 64;
    const float4 c = ap[lane], d = ap[32 + lane];
    ga.x = fmaf(g, c.x, ga.x); ga.y = fmaf(g, c.y, ga.y); ga.z = fmaf(g, c.z, ga.z); ga.w = fmaf(g, c.w, ga.w);
    gb.x = fmaf(g, d.x, gb.x); gb.y = fmaf(g, d.y, gb.y); gb.z = fmaf(g, d.z, gb.z); gb.w = fmaf(g, d.w, gb.w);
    float4* op = reinterpret_cast<float4*>(datt) + (static_cast<size_t>(f) * S + s) * 64;
    op[lane] = make_float4(g * a.x, g * a.y, g * a.z, g * a.w);
    op[32 + lane] = make_float4(g * b.x, g * b.y, g * b.z, g * b.w);
  }
  float4* dp = reinterpret_cast<float4*>(demb) + static_cast<size_t>(f) * 64;
  dp[lane] = ga;
  dp[32 + lane] = gb;
}

// BatchNorm1d over the rows of x [rows][C] in training mode (FS:model:166 on the -1-padded batch): batch mean and
// biased variance per feature from two-stage fixed-order column sums (second stage in double), then
//   y = (x - mean) * rstd * gamma + beta          dx = gamma * rstd * (dy - mean(dy) - xhat * mean(dy * xhat))
// stats[0..C) = mean, stats[C..2C) = biased variance (the caller updates the running statistics from them).
__global__ void __launch_bounds__(256)
bn_colstat_partial_kernel(const float* __restrict__ x, int R, int C, float* __restrict__ p1, float* __restrict__ p2) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= C) return;
  const int r0 = blockIdx.x * kRowsPerBlk, r1 = min(R, r0 + kRowsPerBlk);
  float s = 0.f, q = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float v = x[static_cast<size_t>(r) * C + c];
    s += v;
    q = fmaf(v, v, q);
  }
  p1[static_cast<size_t>(blockIdx.x) * C + c] = s;
  p2[static_cast<size_t>(blockIdx.x) * C + c] = q;
}
__global__ void __launch_bounds__(256)
bn_stat_final_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int n_blk, int R, int C, float* __restrict__ stats) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < n_blk; ++k) {
    s += static_cast<double>(p1[static_cast<size_t>(k) * C + c]);
    q += static_cast<double>(p2[static_cast<size_t>(k) * C + c]);
  }
  const double mean = s / R;
  stats[c] = static_cast<float>(mean);
  stats[C + c] = static_cast<float>(fmax(q / R - mean * mean, 0.0));
}
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ g,
                const float* __restrict__ b, size_t n, int C, float eps, float* __restrict__ y) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = static_cast<int>(i % C);
  y[i] = (x[i] - stats[c]) * (1.f / sqrtf(stats[C + c] + eps)) * g[c] + b[c];
}
// sums[0..C) = colsum(dy), sums[C..2C) = colsum(dy * xhat) (xhat recomputed from x and the saved statistics)
__global__ void __launch_bounds__(256)
bn_bwd_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ stats, int R, int C,
                      float eps, float* __restrict__ p1, float* __restrict__ p2) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= C) return;
  const float mean = stats[c], rstd = 1.f / sqrtf(stats[C + c] + eps);
  const int r0 = blockIdx.x * kRowsPerBlk, r1 = min(R, r0 + kRowsPerBlk);
  float s = 0.f, q = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float d = dy[static_cast<size_t>(r) * C + c];
    s += d;
    q = fmaf(d, (x[static_cast<size_t>(r) * C + c] - mean) * rstd, q);
  }
  p1[static_cast<size_t>(blockIdx.x) * C + c] = s;
  p2[static_cast<size_t>(blockIdx.x) * C + c] = q;
}
__global__ void __launch_bounds__(256)
bn_bwd_final_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int n_blk, int C, float* __restrict__ db,
                    float* __restrict__ dg) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < n_blk; ++k) {
    s += static_cast<double>(p1[static_cast<size_t>(k) * C + c]);
    q += static_cast<double>(p2[static_cast<size_t>(k) * C + c]);
  }
  db[c] = static_cast<float>(s);
  dg[c] = static_cast<float>(q);
}
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ stats,
                    const float* __restrict__ g, const float* __restrict__ db, const float* __restrict__ dg, size_t n, int R,
                    int C, float eps, float* __restrict__ dx) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = static_cast<int>(i % C);
  const float rstd = 1.f / sqrtf(stats[C + c] + eps);
  const float xhat = (x[i] - stats[c]) * rstd;
  const float invR = 1.f / static_cast<float>(R);
  dx[i] = g[c] * rstd * (dy[i] - db[c] * invR - xhat * dg[c] * invR);
}

inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }
inline int pad64(int x) { return (x + 63) / 64 * 64; }
inline int pad128(int x) { return (x + 127) / 128 * 128; }

// Weight-gradient product dW[N][K] = dY^T X, split over the rows.  The operand with the larger feature dimension is A
// (M = that dimension, read through its transpose); the other one becomes the per-slice planes (n_out columns).
struct WgradPlan {
  bool swap;        // false: A = dY (M = N), planes = X^T (n_out = pad128(K));  true: A = X (M = K), planes = dY^T (n_out = N)
  int M, n_out, n_slices, ks;
};
inline WgradPlan wgrad_plan(int rows, int K, int N) {
  WgradPlan pl;
  pl.swap = K > N;
  pl.M = pl.swap ? K : N;
  pl.n_out = pl.swap ? N : pad128(K);
  const int tiles = ((pl.M + 127) / 128) * (pl.n_out / 128);
  // number of K-slices: at least two tiles per SM, then the count whose total tile number fills whole waves of the
  // persistent kernel best (ncu on the first version, 10 slices x 32 tiles = 2.16 waves: SMs idle 29 % of the launch),
  // while a slice keeps >= 1024 rows (16 k-blocks) so that the per-tile epilogue stays amortised
  constexpr int kSms = 148;
  const int s_max = std::max(1, std::min(64, rows / 1024));
  const int s_min = std::min(s_max, std::max(1, (2 * kSms + tiles - 1) / tiles));
  int sl = s_min;
  double best = 0.0;
  for (int c = s_min; c <= s_max; ++c) {
    const int total = tiles * c;
    const double eff = static_cast<double>(total) / (static_cast<double>((total + kSms - 1) / kSms) * kSms);
    if (eff > best + 1e-9) {
      best = eff;
      sl = c;
    }
  }
  pl.ks = pad64((rows + sl - 1) / sl);
  pl.n_slices = (rows + pl.ks - 1) / pl.ks;
  return pl;
}

// fp32 [R][C] -> fp16 hi / lo planes [R][Cp]
void split_rows(const float* src, int R, int C, int Cp, float scale, const float* scale_dev, __half* hi, __half* lo,
                cudaStream_t st) {
  if (C == Cp && C % 4 == 0) {
    const size_t n4 = static_cast<size_t>(R) * C / 4;
    split_planes4_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(src, n4, scale, scale_dev, hi, lo);
  } else {
    const size_t n = static_cast<size_t>(R) * Cp;
    split_planes_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(src, R, C, Cp, scale, scale_dev, hi, lo);
  }
}

CUtensorMap plane_map(const void* base, int rows, int K) {
  uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(rows)};
  uint64_t str[1] = {static_cast<uint64_t>(K)};
  uint32_t box[2] = {64, 128};
  return make_tmap_f16(base, 2, dims, str, box);
}

}  // namespace
}  // namespace fseend

using namespace fseend;

extern "C" {

// Workspace (bytes) of fseend_train_linear_fwd / _bwd for a [rows][K] x [N][K]^T layer.
size_t fseend_train_linear_workspace_bytes(int rows, int K, int N) {
  const size_t Kp = pad64(K), Np = pad128(N), Kp128 = pad128(K);
  const size_t rows_p = pad128(rows);            // A planes are declared to TMA with whole 128-row boxes
  const size_t fwd = 2 * align256(Np * Kp * 2) + 2 * align256(rows_p * Kp * 2);
  const WgradPlan pl = wgrad_plan(rows, K, N);
  const size_t bwd = 256                                                        // max slot + scale pair
                     + 2 * align256(Kp128 * Np * 2)                             // W^T planes [K][N]
                     + 2 * align256(rows_p * Np * 2)                            // dY planes (scaled) for the dgrad
                     + align256(static_cast<size_t>(rows) * Kp128 * 4)          // dX padded
                     + 2 * align256(static_cast<size_t>(pl.n_slices) * pl.n_out * pl.ks * 2)   // small operand's planes
                     + align256(static_cast<size_t>(pl.n_slices) * pl.M * pl.n_out * 4)        // split-K partial products
                     + align256((static_cast<size_t>(rows) / kRowsPerBlk + 1) * N * 4)         // column-sum partials
                     + align256(static_cast<size_t>(rows) * N * 4);                            // ReLU-masked dY (act == 1 only)
  return (fwd > bwd ? fwd : bwd) + 4096;
}

// y[rows][N] = act(x[rows][K] w[N][K]^T + b); all fp32 on the device; act 0 none / 1 ReLU.  N % 128 == 0; any K (padded to 64 inside).
int fseend_train_linear_fwd(const float* x, int rows, int K, const float* w, int N, const float* bias, int act, float* y,
                            void* workspace, size_t ws_bytes, void* stream) {
  return tguard([&] {
    if (!x || !w || !y || !workspace || rows < 1 || K < 1 || N < 1) throw std::invalid_argument("linear_fwd: bad arguments");
    if (N % 128) throw std::invalid_argument("linear_fwd: N must be a multiple of 128");
    if (ws_bytes < fseend_train_linear_workspace_bytes(rows, K, N)) throw std::invalid_argument("linear_fwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Kp = pad64(K), Np = pad128(N);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    __half* hi = reinterpret_cast<__half*>(ws);
    __half* lo = reinterpret_cast<__half*>(ws + align256(static_cast<size_t>(Np) * Kp * 2));
    __half* xhi = reinterpret_cast<__half*>(ws + 2 * align256(static_cast<size_t>(Np) * Kp * 2));
    __half* xlo = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(xhi) + align256(static_cast<size_t>(pad128(rows)) * Kp * 2));
    split_rows(w, N, K, Kp, kWScale, nullptr, hi, lo, st);
    split_rows(x, rows, K, Kp, 1.f, nullptr, xhi, xlo, st);      // activations as planes: both operands arrive by TMA
    P32GemmParams p{};
    p.rows_per_seq = rows;
    p.a_seq_rows = rows;
    p.n_seq = 1;
    p.k_blocks = Kp / 64;
    p.taps = 1;
    p.N = Np;
    p.bias = bias;
    p.act = act;
    p.alpha = 1.f;
    p.w_inv_scale = 1.f / kWScale;
    p.out = y;
    p.ldo = N;
    // rows of the planes beyond `rows` (up to the next multiple of 128) are never written: they only feed output rows that
    // are not stored
    launch_p32_gemm_planes(plane_map(xhi, pad128(rows), Kp), plane_map(xlo, pad128(rows), Kp), plane_map(hi, Np, Kp), plane_map(lo, Np, Kp), p, st);
    TCHECK(cudaGetLastError());
  });
}

// dx[rows][K] (nullable), dw[N][K], db[N] (nullable) from dy[rows][N]; act 1: dy is first masked with y > 0 (y = saved output).
// relu_input != 0: x is itself the output of a ReLU and dx is written as zero where x <= 0 (the FFN pair: the ReLU's
// backward rides in the epilogue of the down-projection's dgrad; the up-projection's backward then runs with act 0).
int fseend_train_linear_bwd(const float* x, const float* w, const float* y, const float* dy, int rows, int K, int N, int act,
                            int relu_input, float* dx, float* dw, float* db, void* workspace, size_t ws_bytes, void* stream) {
  return tguard([&] {
    if (!x || !w || !dy || !dw || !workspace || rows < 1) throw std::invalid_argument("linear_bwd: bad arguments");
    if (N % 128) throw std::invalid_argument("linear_bwd: N must be a multiple of 128");
    if (act == 1 && !y) throw std::invalid_argument("linear_bwd: ReLU backward needs the saved output");
    if (relu_input && K % 128) throw std::invalid_argument("linear_bwd: relu_input needs K % 128 == 0");
    if (ws_bytes < fseend_train_linear_workspace_bytes(rows, K, N)) throw std::invalid_argument("linear_bwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Kp = pad128(K);
    const WgradPlan pl = wgrad_plan(rows, K, N);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    auto take = [&](size_t bytes) {
      uint8_t* p = ws;
      ws += align256(bytes);
      return p;
    };
    unsigned int* slot = reinterpret_cast<unsigned int*>(take(256));
    float* sc = reinterpret_cast<float*>(slot) + 4;                 // sc[0] scale, sc[1] inverse
    __half* wt_hi = reinterpret_cast<__half*>(take(static_cast<size_t>(Kp) * N * 2));
    __half* wt_lo = reinterpret_cast<__half*>(take(static_cast<size_t>(Kp) * N * 2));
    float* dxp = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * Kp * 4));
    __half* dy_hi = reinterpret_cast<__half*>(take(static_cast<size_t>(pad128(rows)) * N * 2));
    __half* dy_lo = reinterpret_cast<__half*>(take(static_cast<size_t>(pad128(rows)) * N * 2));
    const size_t plane_elems = static_cast<size_t>(pl.n_slices) * pl.n_out * pl.ks;
    __half* sp_hi = reinterpret_cast<__half*>(take(plane_elems * 2));
    __half* sp_lo = reinterpret_cast<__half*>(take(plane_elems * 2));
    float* partial = reinterpret_cast<float*>(take(static_cast<size_t>(pl.n_slices) * pl.M * pl.n_out * 4));
    float* part = reinterpret_cast<float*>(take((static_cast<size_t>(rows) / kRowsPerBlk + 1) * N * 4));
    if (act == 1) {
      float* dym = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * N * 4));
      const size_t n4 = static_cast<size_t>(rows) * N / 4;
      relu_bwd_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(y, dy, n4, dym);
      dy = dym;
    }

    // gradient scale: a power of two from max |dY|
    TCHECK(cudaMemsetAsync(slot, 0, 4, st));
    absmax_kernel<<<592, 256, 0, st>>>(dy, static_cast<size_t>(rows) * N, slot);
    make_scale_kernel<<<1, 1, 0, st>>>(slot, sc);

    if (db) {
      const int nb = (rows + kRowsPerBlk - 1) / kRowsPerBlk;
      colsum_partial_kernel<<<dim3(nb, (N + 255) / 256), 256, 0, st>>>(dy, nullptr, nullptr, rows, N, part);
      colsum_final_kernel<<<(N + 31) / 32, 256, 0, st>>>(part, nb, N, db);
    }
    if (dx) {
      // W^T planes [Kp][N] (rows >= K zero): one slice of N "rows"
      transpose_split_kernel<<<dim3(N / 32, Kp / 32), 256, 0, st>>>(w, nullptr, N, K, N, Kp, kWScale, nullptr, wt_hi, wt_lo);
      float* out = (Kp == K) ? dx : dxp;
      split_rows(dy, rows, N, N, 1.f, sc, dy_hi, dy_lo, st);          // scaled gradient planes
      P32GemmParams p{};
      p.rows_per_seq = rows;
      p.a_seq_rows = rows;
      p.n_seq = 1;
      p.k_blocks = N / 64;
      p.taps = 1;
      p.N = Kp;
      p.alpha = 1.f;
      p.w_inv_scale = 1.f / kWScale;
      p.out = out;
      p.ldo = Kp;
      p.out_scale_dev = sc + 1;
      p.out_mask = relu_input ? x : nullptr;            // Kp == K here (checked above): same layout as out
      launch_p32_gemm_planes(plane_map(dy_hi, pad128(rows), N), plane_map(dy_lo, pad128(rows), N), plane_map(wt_hi, Kp, N), plane_map(wt_lo, Kp, N), p, st);
      if (Kp != K)
        TCHECK(cudaMemcpy2DAsync(dx, static_cast<size_t>(K) * 4, dxp, static_cast<size_t>(Kp) * 4, static_cast<size_t>(K) * 4, rows,
                                 cudaMemcpyDeviceToDevice, st));
    }
    {
      // dW: A = the larger operand, read transposed in place; planes = the smaller one, transposed per K-slice
      const float* a_src = pl.swap ? x : dy;
      const int a_cols = pl.swap ? K : N;                      // = M
      const float* s_src = pl.swap ? dy : x;
      const int s_cols = pl.swap ? N : K;
      transpose_split_kernel<<<dim3(pl.n_slices * pl.ks / 32, pl.n_out / 32), 256, 0, st>>>(
          s_src, nullptr, rows, s_cols, pl.ks, pl.n_out, 1.f, pl.swap ? sc : nullptr, sp_hi, sp_lo);
      P32GemmParams p{};
      p.A = a_src;
      p.lda = a_cols;
      p.a_seq_rows = pl.M;
      p.rows_per_seq = pl.M;
      p.n_seq = pl.n_slices;
      p.k_blocks = pl.ks / 64;
      p.taps = 1;
      p.N = pl.n_out;
      p.alpha = 1.f;
      p.w_inv_scale = 1.f;
      p.out = partial;
      p.ldo = pl.n_out;
      p.a_transposed = 1;
      p.a_k_rows = rows;
      p.w_seq_stride = pl.n_out;
      p.a_scale_dev = pl.swap ? nullptr : sc;
      p.out_scale_dev = sc + 1;
      const int plane_rows = pl.n_slices * pl.n_out;
      launch_p32_gemm(plane_map(sp_hi, plane_rows, pl.ks), plane_map(sp_lo, plane_rows, pl.ks), p, st);
      if (pl.swap)
        splitk_reduce_t_kernel<<<dim3((K + 31) / 32, (N + 31) / 32), 256, 0, st>>>(partial, pl.n_slices, K, N, dw);
      else
        splitk_reduce_kernel<<<static_cast<unsigned>((static_cast<size_t>(N) * K + 255) / 256), 256, 0, st>>>(
            partial, pl.n_slices, N, pl.n_out, K, dw);
    }
    TCHECK(cudaGetLastError());
  });
}

// y = LayerNorm(x + r; g, b) over rows of 256 (r nullable).  sum_out (nullable) receives x + r, the tensor the backward
// differentiates through.  Reference: the post-norm residual blocks of nn.TransformerEncoderLayer (FS:model:147).
int fseend_train_add_layernorm_fwd(const float* x, const float* r, const float* g, const float* b, int rows, float eps,
                                   float* sum_out, float* y, void* stream) {
  return tguard([&] {
    if (!x || !g || !b || !y || rows < 1) throw std::invalid_argument("add_layernorm_fwd: bad arguments");
    ln_fwd_row_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, r, g, b, rows, eps, sum_out, y);
    TCHECK(cudaGetLastError());
  });
}

size_t fseend_train_layernorm_workspace_bytes(int rows) {
  return align256(static_cast<size_t>(rows) * 256 * 4) + align256((static_cast<size_t>(rows) / kRowsPerBlk + 1) * 256 * 4) + 1024;
}

// LayerNorm(256) backward: x, dy [rows][256], g [256] -> dx [rows][256], dg [256], db [256].
int fseend_train_layernorm_bwd(const float* x, const float* g, const float* dy, int rows, float eps, float* dx, float* dg,
                               float* db, void* workspace, size_t ws_bytes, void* stream) {
  return tguard([&] {
    if (!x || !g || !dy || !dx || !dg || !db || !workspace || rows < 1) throw std::invalid_argument("layernorm_bwd: bad arguments");
    if (ws_bytes < fseend_train_layernorm_workspace_bytes(rows)) throw std::invalid_argument("layernorm_bwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* xhat = static_cast<float*>(workspace);
    float* part = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + align256(static_cast<size_t>(rows) * 256 * 4));
    ln_bwd_row_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, g, dy, rows, eps, dx, xhat);
    const int nb = (rows + kRowsPerBlk - 1) / kRowsPerBlk;
    colsum_partial_kernel<<<dim3(nb, 1), 256, 0, st>>>(dy, xhat, nullptr, rows, 256, part);
    colsum_final_kernel<<<8, 256, 0, st>>>(part, nb, 256, dg);
    colsum_partial_kernel<<<dim3(nb, 1), 256, 0, st>>>(dy, nullptr, nullptr, rows, 256, part);
    colsum_final_kernel<<<8, 256, 0, st>>>(part, nb, 256, db);
    TCHECK(cudaGetLastError());
  });
}

// ---- L2 normalisation, head, BatchNorm (training)
int fseend_train_l2norm_fwd(const float* x, int rows, float* y, float* inv_norm, void* stream) {
  return tguard([&] {
    if (!x || !y || !inv_norm || rows < 1) throw std::invalid_argument("l2norm_fwd: bad arguments");
    l2norm_fwd_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, y, inv_norm);
    TCHECK(cudaGetLastError());
  });
}
int fseend_train_l2norm_bwd(const float* y, const float* inv_norm, const float* dy, int rows, float* dx, void* stream) {
  return tguard([&] {
    if (!y || !inv_norm || !dy || !dx || rows < 1) throw std::invalid_argument("l2norm_bwd: bad arguments");
    l2norm_bwd_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, inv_norm, dy, rows, dx);
    TCHECK(cudaGetLastError());
  });
}
int fseend_train_head_fwd(const float* emb, const float* att, int frames, int S, float* y, void* stream) {
  return tguard([&] {
    if (!emb || !att || !y || frames < 1 || S < 1) throw std::invalid_argument("head_fwd: bad arguments");
    head_fwd_kernel<<<(frames + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(emb, att, frames, S, y);
    TCHECK(cudaGetLastError());
  });
}
int fseend_train_head_bwd(const float* emb, const float* att, const float* dy, int frames, int S, float* demb, float* datt,
                          void* stream) {
  return tguard([&] {
    if (!emb || !att || !dy || !demb || !datt || frames < 1 || S < 1) throw std::invalid_argument("head_bwd: bad arguments");
    head_bwd_kernel<<<(frames + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(emb, att, dy, frames, S, demb, datt);
    TCHECK(cudaGetLastError());
  });
}
size_t fseend_train_batchnorm_workspace_bytes(int rows, int C) {
  return 2 * align256((static_cast<size_t>(rows) / kRowsPerBlk + 1) * C * 4) + 1024;
}
// y = BatchNorm(x) with batch statistics; stats [2 * C] receives mean | biased variance
int fseend_train_batchnorm_fwd(const float* x, const float* g, const float* b, int rows, int C, float eps, float* y,
                               float* stats, void* workspace, size_t ws_bytes, void* stream) {
  return tguard([&] {
    if (!x || !g || !b || !y || !stats || !workspace || rows < 1 || C < 1) throw std::invalid_argument("batchnorm_fwd: bad arguments");
    if (ws_bytes < fseend_train_batchnorm_workspace_bytes(rows, C)) throw std::invalid_argument("batchnorm_fwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nb = (rows + kRowsPerBlk - 1) / kRowsPerBlk;
    float* p1 = static_cast<float*>(workspace);
    float* p2 = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + align256(static_cast<size_t>(nb) * C * 4));
    bn_colstat_partial_kernel<<<dim3(nb, (C + 255) / 256), 256, 0, st>>>(x, rows, C, p1, p2);
    bn_stat_final_kernel<<<(C + 255) / 256, 256, 0, st>>>(p1, p2, nb, rows, C, stats);
    const size_t n = static_cast<size_t>(rows) * C;
    bn_apply_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(x, stats, g, b, n, C, eps, y);
    TCHECK(cudaGetLastError());
  });
}
int fseend_train_batchnorm_bwd(const float* x, const float* g, const float* stats, const float* dy, int rows, int C, float eps,
                               float* dx, float* dg, float* db, void* workspace, size_t ws_bytes, void* stream) {
  return tguard([&] {
    if (!x || !g || !stats || !dy || !dg || !db || !workspace || rows < 1 || C < 1)      // dx may be null (input features)
      throw std::invalid_argument("batchnorm_bwd: bad arguments");
    if (ws_bytes < fseend_train_batchnorm_workspace_bytes(rows, C)) throw std::invalid_argument("batchnorm_bwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nb = (rows + kRowsPerBlk - 1) / kRowsPerBlk;
    float* p1 = static_cast<float*>(workspace);
    float* p2 = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + align256(static_cast<size_t>(nb) * C * 4));
    bn_bwd_partial_kernel<<<dim3(nb, (C + 255) / 256), 256, 0, st>>>(x, dy, stats, rows, C, eps, p1, p2);
    bn_bwd_final_kernel<<<(C + 255) / 256, 256, 0, st>>>(p1, p2, nb, C, db, dg);
    const size_t n = static_cast<size_t>(rows) * C;
    if (dx) bn_bwd_apply_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(x, dy, stats, g, db, dg, n, rows, C, eps, dx);
    TCHECK(cudaGetLastError());
  });
}

}  // extern "C"
