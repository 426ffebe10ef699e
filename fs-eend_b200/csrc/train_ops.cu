// Training-step building blocks (STARTED — not a training step yet): forward + backward of the ops
// that carry ~90 % of the step's FLOPs, in the parity ("p32") arithmetic: fp32 tensors in HBM, every matrix product on
// tcgen05 with split-precision operands (p32.cuh), reductions in fixed order.  Behind fseend_train_* C-ABI exports and
// torch.autograd.Functions (fseend_b200/autograd.py); pinned against torch autograd in tests/test_train_ops_gpu.py.
//
//   Linear  y = act(x W^T + b)           reference call sites: every nn.Linear of FS:model / FS:fusion
//     dX = dY W            p32 GEMM, A = dY [rows][N], planes = W^T [K][N]
//     dW = dY^T X          p32 GEMM, A = dY^T [N][rows], planes = X^T [K][rows]   (both operands transposed on the device)
//     db = column sums of dY (two-stage, fixed order)
//   LayerNorm (biased variance, eps inside the sqrt): dx per row, dgamma / dbeta as two-stage column sums
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

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

// src fp32 [R][C] -> hi / lo fp16 planes [R][Cp] (columns >= C zero), values multiplied by `scale`
__global__ void split_planes_kernel(const float* __restrict__ src, int R, int C, int Cp, float scale,
                                    __half* __restrict__ hi, __half* __restrict__ lo) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(R) * Cp) return;
  const int r = static_cast<int>(i / Cp), c = static_cast<int>(i % Cp);
  __half h, l;
  split1(c < C ? src[static_cast<size_t>(r) * C + c] * scale : 0.f, h, l);
  hi[i] = h;
  lo[i] = l;
}

// src fp32 [R][C] -> transposed [Cp][Rp] (zero padded): fp32 (dst32) and / or split planes (hi, lo), 32 x 32 smem tiles
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, int R, int C, int Rp, int Cp, float scale, float* __restrict__ dst32,
                 __half* __restrict__ hi, __half* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < R && c < C) ? src[static_cast<size_t>(r) * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;                           // output row = source column
    if (c < Cp && r < Rp) {
      const float v = tile[tx][j] * scale;
      const size_t o = static_cast<size_t>(c) * Rp + r;
      if (dst32) dst32[o] = v;
      if (hi) {
        __half h, l;
        split1(v, h, l);
        hi[o] = h;
        lo[o] = l;
      }
    }
  }
}

// partial[blk][c] = sum over the block's rows of f(row, c); then out[c] = sum over blocks (fixed order)
constexpr int kRowsPerBlk = 256;
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int R, int C, float* __restrict__ partial) {
  // a [R][C]; b optional [R][C]: sums a * b (LayerNorm dgamma) or a alone
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= C) return;
  const int r0 = blockIdx.x * kRowsPerBlk, r1 = min(R, r0 + kRowsPerBlk);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float v = a[static_cast<size_t>(r) * C + c];
    s += b ? v * b[static_cast<size_t>(r) * C + c] : v;
  }
  partial[static_cast<size_t>(blockIdx.x) * C + c] = s;
}
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, int n_blk, int C, float* __restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int k = 0; k < n_blk; ++k) s += static_cast<double>(partial[static_cast<size_t>(k) * C + c]);
  out[c] = static_cast<float>(s);
}

// dy *= (y > 0)   (ReLU backward on the saved output)
__global__ void relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, size_t n, float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = y[i] > 0.f ? dy[i] : 0.f;
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

inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }
inline int pad64(int x) { return (x + 63) / 64 * 64; }
inline int pad128(int x) { return (x + 127) / 128 * 128; }

CUtensorMap plane_map(const void* base, int rows, int K) {
  uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(rows)};
  uint64_t str[1] = {static_cast<uint64_t>(K)};
  uint32_t box[2] = {64, 128};
  return make_tmap_f16(base, 2, dims, str, box);
}

void run_gemm(const float* A, int lda, int rows, int kdim, const __half* hi, const __half* lo, int n_out, const float* bias,
              int act, float inv_scale, float* out, int ldo, cudaStream_t st) {
  // out[rows][n_out] = act(inv_scale * A[rows][kdim] planes[n_out][kdim]^T + bias)
  P32GemmParams p{};
  p.A = A;
  p.lda = lda;
  p.a_seq_rows = rows;
  p.rows_per_seq = rows;
  p.n_seq = 1;
  p.k_blocks = kdim / 64;
  p.taps = 1;
  p.N = n_out;
  p.bias = bias;
  p.act = act;
  p.alpha = 1.f;
  p.w_inv_scale = inv_scale;
  p.out = out;
  p.ldo = ldo;
  const CUtensorMap th = plane_map(hi, n_out, kdim), tl = plane_map(lo, n_out, kdim);
  launch_p32_gemm(th, tl, p, st);
}

}  // namespace
}  // namespace fseend

using namespace fseend;

extern "C" {

// Workspace (bytes) of fseend_train_linear_fwd / _bwd for a [rows][K] x [N][K]^T layer.
size_t fseend_train_linear_workspace_bytes(int rows, int K, int N) {
  const size_t Kp = pad64(K), Np = pad128(N), Kp128 = pad128(K), Rp = pad64(rows);
  size_t fwd = 2 * align256(Np * Kp * 2) + align256(static_cast<size_t>(rows) * Kp * 4) + align256(static_cast<size_t>(rows) * Np * 4);
  size_t bwd = 2 * align256(Kp128 * pad64(N) * 2)                  // W^T planes [K][N]
               + align256(static_cast<size_t>(rows) * pad64(N) * 4)   // dY (relu-masked / column padded)
               + align256(static_cast<size_t>(rows) * Kp128 * 4)      // dX padded
               + align256(Np * Rp * 4)                                // dY^T fp32
               + 2 * align256(Kp128 * Rp * 2)                         // X^T planes
               + align256(Np * Kp128 * 4)                             // dW padded
               + align256((static_cast<size_t>(rows) / kRowsPerBlk + 1) * N * 4);   // column-sum partials
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
    float* xp = reinterpret_cast<float*>(ws + 2 * align256(static_cast<size_t>(Np) * Kp * 2));
    const size_t nw = static_cast<size_t>(Np) * Kp;
    TCHECK(cudaMemsetAsync(hi, 0, nw * 2, st));
    TCHECK(cudaMemsetAsync(lo, 0, nw * 2, st));
    split_planes_kernel<<<static_cast<unsigned>((static_cast<size_t>(N) * Kp + 255) / 256), 256, 0, st>>>(w, N, K, Kp, kWScale, hi, lo);
    const float* xa = x;
    if (Kp != K) {      // pad the activation columns (the GEMM reads K in blocks of 64 with 16-byte loads)
      TCHECK(cudaMemsetAsync(xp, 0, static_cast<size_t>(rows) * Kp * 4, st));
      TCHECK(cudaMemcpy2DAsync(xp, static_cast<size_t>(Kp) * 4, x, static_cast<size_t>(K) * 4, static_cast<size_t>(K) * 4, rows,
                               cudaMemcpyDeviceToDevice, st));
      xa = xp;
    }
    run_gemm(xa, Kp, rows, Kp, hi, lo, Np, bias, act, 1.f / kWScale, y, N, st);
    TCHECK(cudaGetLastError());
  });
}

// dx[rows][K] (nullable), dw[N][K], db[N] (nullable) from dy[rows][N]; act 1: dy is first masked with y > 0 (y = saved output).
int fseend_train_linear_bwd(const float* x, const float* w, const float* y, const float* dy, int rows, int K, int N, int act,
                            float* dx, float* dw, float* db, void* workspace, size_t ws_bytes, void* stream) {
  return tguard([&] {
    if (!x || !w || !dy || !dw || !workspace || rows < 1) throw std::invalid_argument("linear_bwd: bad arguments");
    if (N % 128) throw std::invalid_argument("linear_bwd: N must be a multiple of 128");
    if (act == 1 && !y) throw std::invalid_argument("linear_bwd: ReLU backward needs the saved output");
    if (ws_bytes < fseend_train_linear_workspace_bytes(rows, K, N)) throw std::invalid_argument("linear_bwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Kp = pad128(K), Rp = pad64(rows);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    auto take = [&](size_t bytes) {
      uint8_t* p = ws;
      ws += align256(bytes);
      return p;
    };
    __half* wt_hi = reinterpret_cast<__half*>(take(static_cast<size_t>(Kp) * N * 2));
    __half* wt_lo = reinterpret_cast<__half*>(take(static_cast<size_t>(Kp) * N * 2));
    float* dym = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * N * 4));
    float* dxp = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * Kp * 4));
    float* dyt = reinterpret_cast<float*>(take(static_cast<size_t>(N) * Rp * 4));
    __half* xt_hi = reinterpret_cast<__half*>(take(static_cast<size_t>(Kp) * Rp * 2));
    __half* xt_lo = reinterpret_cast<__half*>(take(static_cast<size_t>(Kp) * Rp * 2));
    float* dwp = reinterpret_cast<float*>(take(static_cast<size_t>(N) * Kp * 4));
    float* part = reinterpret_cast<float*>(take((static_cast<size_t>(rows) / kRowsPerBlk + 1) * N * 4));
    const float* g = dy;
    if (act == 1) {
      const size_t n = static_cast<size_t>(rows) * N;
      relu_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(y, dy, n, dym);
      g = dym;
    }
    if (db) {
      const int nb = (rows + kRowsPerBlk - 1) / kRowsPerBlk;
      colsum_partial_kernel<<<dim3(nb, (N + 255) / 256), 256, 0, st>>>(g, nullptr, rows, N, part);
      colsum_final_kernel<<<(N + 255) / 256, 256, 0, st>>>(part, nb, N, db);
    }
    if (dx) {
      // W^T planes [Kp][N] (rows >= K zero): transpose of w [N][K]
      transpose_kernel<<<dim3((N + 31) / 32, (Kp + 31) / 32), 256, 0, st>>>(w, N, K, N, Kp, kWScale, nullptr, wt_hi, wt_lo);
      float* out = (Kp == K) ? dx : dxp;
      run_gemm(g, N, rows, N, wt_hi, wt_lo, Kp, nullptr, 0, 1.f / kWScale, out, Kp, st);
      if (Kp != K)
        TCHECK(cudaMemcpy2DAsync(dx, static_cast<size_t>(K) * 4, dxp, static_cast<size_t>(Kp) * 4, static_cast<size_t>(K) * 4, rows,
                                 cudaMemcpyDeviceToDevice, st));
    }
    // dW[N][K] = sum_r dY[r][n] X[r][k]: A = dY^T [N][Rp] fp32, planes = X^T [Kp][Rp]
    transpose_kernel<<<dim3(Rp / 32, (N + 31) / 32), 256, 0, st>>>(g, rows, N, Rp, N, 1.f, dyt, nullptr, nullptr);    // the grid covers the zero padding too
    transpose_kernel<<<dim3(Rp / 32, (Kp + 31) / 32), 256, 0, st>>>(x, rows, K, Rp, Kp, 1.f, nullptr, xt_hi, xt_lo);
    float* out = (Kp == K) ? dw : dwp;
    run_gemm(dyt, Rp, N, Rp, xt_hi, xt_lo, Kp, nullptr, 0, 1.f, out, Kp, st);
    if (Kp != K)
      TCHECK(cudaMemcpy2DAsync(dw, static_cast<size_t>(K) * 4, dwp, static_cast<size_t>(Kp) * 4, static_cast<size_t>(K) * 4, N,
                               cudaMemcpyDeviceToDevice, st));
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
    colsum_partial_kernel<<<dim3(nb, 1), 256, 0, st>>>(dy, xhat, rows, 256, part);
    colsum_final_kernel<<<1, 256, 0, st>>>(part, nb, 256, dg);
    colsum_partial_kernel<<<dim3(nb, 1), 256, 0, st>>>(dy, nullptr, rows, 256, part);
    colsum_final_kernel<<<1, 256, 0, st>>>(part, nb, 256, db);
    TCHECK(cudaGetLastError());
  });
}

}  // extern "C"
