// Bandwidth-bound kernels of the FS-EEND hot path: input BatchNorm+cast, speaker-axis attention, logits head.
#include "elementwise.cuh"

#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>

namespace fseend {

namespace {

// ---------------------------------------------------------------------------------------------
// x (packed fp32 rows, cu_seqlens) -> BN(eval) -> fp16 [B][Tmax][Kpad]; rows t >= len take x = -1
// (the reference pads with -1 BEFORE BatchNorm, FS:model:165-166); columns >= Din are zero.
__global__ void prep_input_kernel(const float* __restrict__ x, const int* __restrict__ cu, int Tmax, int Din,
                                  int Kpad, const float* __restrict__ sc, const float* __restrict__ sh,
                                  float pad_value, __half* __restrict__ out) {
  const int row = blockIdx.x;  // b * Tmax + t
  const int b = row / Tmax, t = row - b * Tmax;
  const int start = cu[b], len = cu[b + 1] - start;
  const float* src = (t < len) ? x + static_cast<size_t>(start + t) * Din : nullptr;
  __half2* dst = reinterpret_cast<__half2*>(out + static_cast<size_t>(row) * Kpad);
  for (int i = threadIdx.x; i < Kpad / 2; i += blockDim.x) {
    const int c0 = 2 * i, c1 = 2 * i + 1;
    float v0 = 0.f, v1 = 0.f;
    if (c0 < Din) v0 = fmaf(src ? __ldg(src + c0) : pad_value, sc ? sc[c0] : 1.f, sh ? sh[c0] : 0.f);
    if (c1 < Din) v1 = fmaf(src ? __ldg(src + c1) : pad_value, sc ? sc[c1] : 1.f, sh ? sh[c1] : 0.f);
    dst[i] = __floats2half2_rn(v0, v1);
  }
}

// ---------------------------------------------------------------------------------------------
// Speaker-axis attention (FS:fusion:390, no mask): for every frame, S x S attention per head over the
// S attractor rows.  qkv: [frames][S][768] fp16, out: [frames][S][256] fp16.
// One thread = (frame, head, query slot); K/V of the block's frames are staged in shared memory.
constexpr int kMaxS = 16;

__global__ void __launch_bounds__(128)
spk_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int n_frames, int S, float scale,
                int frames_per_block) {
  extern __shared__ uint4 kv_smem[];  // [F][S][64 uint4] : K (32 uint4) then V (32 uint4) per row
  const int f0 = blockIdx.x * frames_per_block;
  const int nf = min(frames_per_block, n_frames - f0);
  const int rows = nf * S;
  for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) {
    const int rr = i >> 6, cc = i & 63;
    kv_smem[i] = __ldg(reinterpret_cast<const uint4*>(qkv + (static_cast<size_t>(f0) * S + rr) * 768 + 256) + cc);
  }
  __syncthreads();
  const int tpf = 4 * S;
  const int f = threadIdx.x / tpf;
  if (f >= nf) return;
  const int rem = threadIdx.x - f * tpf;
  const int a = rem >> 2, h = rem & 3;   // heads fastest: 4 consecutive threads read 512 contiguous bytes of q

  float q[64];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(qkv + (static_cast<size_t>(f0 + f) * S + a) * 768 + h * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 u = __ldg(qp + i);
      const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 t = __half22float2(hh[j]);
        q[i * 8 + 2 * j] = t.x * scale;
        q[i * 8 + 2 * j + 1] = t.y * scale;
      }
    }
  }
  float sc[kMaxS];
  float mx = -INFINITY;
#pragma unroll
  for (int bb = 0; bb < kMaxS; ++bb) {
    sc[bb] = -INFINITY;
    if (bb < S) {
      const uint4* kp = kv_smem + (f * S + bb) * 64 + h * 8;
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 u = kp[i];
        const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 t = __half22float2(hh[j]);
          d = fmaf(q[i * 8 + 2 * j], t.x, d);
          d = fmaf(q[i * 8 + 2 * j + 1], t.y, d);
        }
      }
      sc[bb] = d;
      mx = fmaxf(mx, d);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int bb = 0; bb < kMaxS; ++bb) {
    if (bb < S) {
      sc[bb] = __expf(sc[bb] - mx);
      sum += sc[bb];
    }
  }
  const float inv = 1.f / sum;
  float o[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] = 0.f;
#pragma unroll
  for (int bb = 0; bb < kMaxS; ++bb) {
    if (bb < S) {
      // the reference multiplies fp32 probabilities with V; keep P in fp32 (no tensor core here)
      const float pw = sc[bb] * inv;
      const uint4* vp = kv_smem + (f * S + bb) * 64 + 32 + h * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 u = vp[i];
        const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 t = __half22float2(hh[j]);
          o[i * 8 + 2 * j] = fmaf(pw, t.x, o[i * 8 + 2 * j]);
          o[i * 8 + 2 * j + 1] = fmaf(pw, t.y, o[i * 8 + 2 * j + 1]);
        }
      }
    }
  }
  uint4* op = reinterpret_cast<uint4*>(out + (static_cast<size_t>(f0 + f) * S + a) * 256 + h * 64);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 u;
    __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) hh[j] = __floats2half2_rn(o[i * 8 + 2 * j], o[i * 8 + 2 * j + 1]);
    op[i] = u;
  }
}

// ---------------------------------------------------------------------------------------------
// Logits head (FS:model:43,60): att_n = att / ||att||, y[frame][s] = emb[frame] . att_n[frame][s].
// One warp per frame; each lane owns 8 of the 256 channels.
__global__ void __launch_bounds__(256)
head_kernel(const __half* __restrict__ emb, const __half* __restrict__ att, int n_frames, int S,
            float* __restrict__ logits, float* __restrict__ emb_f32, float* __restrict__ att_f32) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_frames) return;
  float e[8];
  {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(emb + static_cast<size_t>(warp) * 256) + lane);
    const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 t = __half22float2(hh[j]);
      e[2 * j] = t.x;
      e[2 * j + 1] = t.y;
    }
    if (emb_f32) {
      float4* ep = reinterpret_cast<float4*>(emb_f32 + static_cast<size_t>(warp) * 256 + lane * 8);
      ep[0] = make_float4(e[0], e[1], e[2], e[3]);
      ep[1] = make_float4(e[4], e[5], e[6], e[7]);
    }
  }
  for (int s = 0; s < S; ++s) {
    const size_t row = static_cast<size_t>(warp) * S + s;
    uint4 u = __ldg(reinterpret_cast<const uint4*>(att + row * 256) + lane);
    const __half2* hh = reinterpret_cast<const __half2*>(&u);
    float a[8];
    float ss = 0.f, dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 t = __half22float2(hh[j]);
      a[2 * j] = t.x;
      a[2 * j + 1] = t.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ss = fmaf(a[j], a[j], ss);
      dot = fmaf(a[j], e[j], dot);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, off);
      dot += __shfl_xor_sync(0xffffffffu, dot, off);
    }
    const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
    if (lane == 0) logits[row] = dot * inv;
    if (att_f32) {
      float4* ap = reinterpret_cast<float4*>(att_f32 + row * 256 + lane * 8);
      ap[0] = make_float4(a[0] * inv, a[1] * inv, a[2] * inv, a[3] * inv);
      ap[1] = make_float4(a[4] * inv, a[5] * inv, a[6] * inv, a[7] * inv);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Streaming (frame-by-frame) attention step: one new query row per sequence attends over the projected K/V cache
// [n_seq][cap][256] (the reference re-projects its cached layer inputs every step, FS:stream_mod:28-35 — same
// arithmetic).  The block first appends this frame's K/V at `pos`, then attends over keys 0..pos.
// grid (n_seq, 4 heads), 128 threads: keys are strided over threads, partial softmax states merged through smem.
__global__ void __launch_bounds__(128)
step_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ kcache, __half* __restrict__ vcache, int cap,
                 int pos_arg, const int* __restrict__ pos_dev, float scale, __half* __restrict__ out) {
  const int n = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int pos = pos_dev ? *pos_dev : pos_arg;   // device-resident frame counter: the launch is CUDA-graph replayable
  __shared__ float q_s[64];
  __shared__ float red_m[128], red_l[128];
  __shared__ float red_o[4][64];
  const __half* row = qkv + static_cast<size_t>(n) * 768;
  __half* kc = kcache + (static_cast<size_t>(n) * cap) * 256 + h * 64;
  __half* vc = vcache + (static_cast<size_t>(n) * cap) * 256 + h * 64;
  if (tid < 64) {
    q_s[tid] = __half2float(row[h * 64 + tid]) * scale * 1.4426950408889634f;
    kc[static_cast<size_t>(pos) * 256 + tid] = row[256 + h * 64 + tid];
    vc[static_cast<size_t>(pos) * 256 + tid] = row[512 + h * 64 + tid];
  }
  __syncthreads();
  float m = -INFINITY, l = 0.f, o[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] = 0.f;
  for (int j = tid; j <= pos; j += 128) {
    const uint4* kp = reinterpret_cast<const uint4*>(kc + static_cast<size_t>(j) * 256);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 u = kp[i];
      const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = __half22float2(hh[t]);
        s = fmaf(q_s[i * 8 + 2 * t], f.x, s);
        s = fmaf(q_s[i * 8 + 2 * t + 1], f.y, s);
      }
    }
    const float m_new = fmaxf(m, s);
    const float a = exp2f(m - m_new), pj = exp2f(s - m_new);
    l = l * a + pj;
    const uint4* vp = reinterpret_cast<const uint4*>(vc + static_cast<size_t>(j) * 256);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 u = vp[i];
      const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = __half22float2(hh[t]);
        o[i * 8 + 2 * t] = fmaf(pj, f.x, o[i * 8 + 2 * t] * a);
        o[i * 8 + 2 * t + 1] = fmaf(pj, f.y, o[i * 8 + 2 * t + 1] * a);
      }
    }
    m = m_new;
  }
  // merge the 128 partial (m, l, o) states
  red_m[tid] = m;
  __syncthreads();
  float gm = -INFINITY;
  for (int i = 0; i < 128; ++i) gm = fmaxf(gm, red_m[i]);   // smem broadcast reads
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
    out[static_cast<size_t>(n) * 256 + h * 64 + tid] = __float2half_rn(v / lt);
  }
}

// rows [n_seq] of width 256 copied (or zero-filled when src == nullptr) into hist[n][pos]
__global__ void hist_append_kernel(const __half* __restrict__ src, __half* __restrict__ hist, int cap, int pos_arg,
                                   const int* __restrict__ pos_dev) {
  const int n = blockIdx.x;
  const int pos = pos_dev ? *pos_dev : pos_arg;
  const uint32_t* s = src ? reinterpret_cast<const uint32_t*>(src + static_cast<size_t>(n) * 256) : nullptr;
  uint32_t* d = reinterpret_cast<uint32_t*>(hist + (static_cast<size_t>(n) * cap + pos) * 256);
  d[threadIdx.x] = s ? s[threadIdx.x] : 0u;
}

// ---------------------------------------------------------------------------------------------
// Conformer conv module middle (LS:conf/convolution.py:144-146): causal depthwise Conv1d (kernel K <= 32, left
// zero padding K-1, no bias) -> BatchNorm1d (eval, folded to scale/shift) -> swish.
// u, out: [n_seq][T][256] fp16.  Block = (64 frames, one sequence); thread = 2 adjacent channels.
// hist (optional): [n_seq][K-1][256] fp16 one-step cache: prepended instead of zeros, and updated when T == 1.
__global__ void __launch_bounds__(128)
dwconv_bn_swish_kernel(const __half* __restrict__ u, const float* __restrict__ w, const float* __restrict__ sc,
                       const float* __restrict__ sh, int T, int K, __half* __restrict__ hist,
                       __half* __restrict__ out) {
  const int n = blockIdx.y, t0 = blockIdx.x * 64, c = threadIdx.x * 2;
  float w0[32], w1[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    w0[k] = k < K ? w[c * K + k] : 0.f;
    w1[k] = k < K ? w[(c + 1) * K + k] : 0.f;
  }
  const float s0 = sc[c], s1 = sc[c + 1], h0 = sh[c], h1 = sh[c + 1];
  const __half2* up = reinterpret_cast<const __half2*>(u + static_cast<size_t>(n) * T * 256 + c);
  __half2* hp = hist ? reinterpret_cast<__half2*>(hist + static_cast<size_t>(n) * (K - 1) * 256 + c) : nullptr;
  float2 win[32];   // win[k] = input at t - (K-1) + k
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int t = t0 - (K - 1) + k;
    float2 v = make_float2(0.f, 0.f);
    if (k < K - 1) {
      if (t >= 0) v = __half22float2(up[static_cast<size_t>(t) * 128]);
      else if (hp) v = __half22float2(hp[static_cast<size_t>(t + K - 1) * 128]);
    }
    win[k] = v;
  }
  const int t_end = min(t0 + 64, T);
  for (int t = t0; t < t_end; ++t) {
    const float2 cur = __half22float2(up[static_cast<size_t>(t) * 128]);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k < K - 1) {
        a0 = fmaf(w0[k], win[k].x, a0);
        a1 = fmaf(w1[k], win[k].y, a1);
      }
    }
    a0 = fmaf(w0[K - 1], cur.x, a0);
    a1 = fmaf(w1[K - 1], cur.y, a1);
#pragma unroll
    for (int k = 0; k < 31; ++k) win[k] = (k + 1 < K - 1) ? win[k + 1] : win[k];
    if (K >= 2) {
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k == K - 2) win[k] = cur;
    }
    a0 = fmaf(a0, s0, h0);
    a1 = fmaf(a1, s1, h1);
    a0 = a0 / (1.f + __expf(-a0));
    a1 = a1 / (1.f + __expf(-a1));
    reinterpret_cast<__half2*>(out + (static_cast<size_t>(n) * T + t) * 256 + c)[0] = __floats2half2_rn(a0, a1);
  }
  if (hp && T == 1) {   // one-step: slide the cache
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < K - 1) hp[static_cast<size_t>(k) * 128] = __floats2half2_rn(win[k].x, win[k].y);
  }
}

// Batch variant for K = 16 (the published LS-EEND configuration): block = 32 frames x 256 channels of one sequence,
// 256 threads; thread = (channel pair, 16-frame half).  The 31 input frames a thread needs are read once (coalesced
// half2 loads: consecutive threads = consecutive channel pairs) and kept in registers with static indices; 16 x 16
// FMAs per channel.  HBM-bound: read + write of the [n_seq][T][256] fp16 activations (the first version walked 64
// frames serially per thread with a shifting register window: 189 us at B=16, T=2000 against ~10 us of traffic).
__global__ void __launch_bounds__(256)
dwconv16_bn_swish_kernel(const __half* __restrict__ u, const float* __restrict__ w, const float* __restrict__ sc,
                         const float* __restrict__ sh, int T, __half* __restrict__ out) {
  constexpr int K = 16, F = 16;                       // taps; output frames per thread
  const int n = blockIdx.y, cp = threadIdx.x & 127, c = cp * 2;
  const int t0 = blockIdx.x * 32 + (threadIdx.x >> 7) * F;
  if (t0 >= T) return;
  float w0[K], w1[K];
#pragma unroll
  for (int k = 0; k < K; k += 4) {
    const float4 a = *reinterpret_cast<const float4*>(w + c * K + k);
    const float4 b = *reinterpret_cast<const float4*>(w + (c + 1) * K + k);
    w0[k] = a.x; w0[k + 1] = a.y; w0[k + 2] = a.z; w0[k + 3] = a.w;
    w1[k] = b.x; w1[k + 1] = b.y; w1[k + 2] = b.z; w1[k + 3] = b.w;
  }
  const float s0 = sc[c], s1 = sc[c + 1], h0 = sh[c], h1 = sh[c + 1];
  const __half2* up = reinterpret_cast<const __half2*>(u + static_cast<size_t>(n) * T * 256 + c);
  float2 x[F + K - 1];                                // x[i] = input frame t0 - (K-1) + i (zero before the start)
#pragma unroll
  for (int i = 0; i < F + K - 1; ++i) {
    const int t = t0 - (K - 1) + i;
    x[i] = (t >= 0 && t < T) ? __half22float2(up[static_cast<size_t>(t) * 128]) : make_float2(0.f, 0.f);
  }
  __half2* op = reinterpret_cast<__half2*>(out + static_cast<size_t>(n) * T * 256 + c);
#pragma unroll
  for (int j = 0; j < F; ++j) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      a0 = fmaf(w0[k], x[j + k].x, a0);
      a1 = fmaf(w1[k], x[j + k].y, a1);
    }
    a0 = fmaf(a0, s0, h0);
    a1 = fmaf(a1, s1, h1);
    a0 = a0 / (1.f + __expf(-a0));
    a1 = a1 / (1.f + __expf(-a1));
    if (t0 + j < T) op[static_cast<size_t>(t0 + j) * 128] = __floats2half2_rn(a0, a1);
  }
}

// ---------------------------------------------------------------------------------------------
// Post-processing of the speaker-activity posteriors (reference: train/utils/make_rttm.py:10-15 and metrics.py:58-60):
// decision = pred > threshold, then a median filter of odd width `median` along time with zero padding
// (scipy.signal.medfilt(pred, (median, 1))).  On 0/1 input the median is a majority vote: 1 iff more than median/2 of
// the window are 1.  pred: [T][C] fp32, out: [T][C] uint8.  One thread per (t, c); the window re-reads hit L1/L2
// (HBM-bound: T*C*4 bytes in, T*C bytes out).  Bit-exact with the reference: the only arithmetic is the comparison.
__global__ void __launch_bounds__(256)
decide_median_kernel(const float* __restrict__ pred, int T, int C, float threshold, int median,
                     unsigned char* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(T) * C) return;
  const int t = static_cast<int>(i / C), c = static_cast<int>(i % C);
  if (median <= 1) {
    out[i] = pred[i] > threshold ? 1 : 0;
    return;
  }
  const int half = median / 2;
  int ones = 0;
  for (int k = -half; k <= half; ++k) {
    const int tt = t + k;
    if (tt >= 0 && tt < T) ones += pred[static_cast<long long>(tt) * C + c] > threshold ? 1 : 0;
  }
  out[i] = ones > half ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Feature front-end tail (reference datasets/feature.py:103-133, called at :259-261 / :348-352): frame splicing with
// +-context zero-padded neighbours followed by frame subsampling.  y [T][F] fp32 -> out [ceil(T / sub)][(2 ctx + 1) F]:
// out[j][k * F + f] = y[j * sub - ctx + k][f] (0 outside [0, T)).  Pure data movement (HBM-bound, bit-exact); doing it
// on the device lets the host ship the 23-dim log-mel frames instead of the 345-dim spliced ones.
__global__ void __launch_bounds__(256)
splice_subsample_kernel(const float* __restrict__ y, int T, int F, int ctx, int sub, int T_out,
                        float* __restrict__ out) {
  const int W = (2 * ctx + 1) * F;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(T_out) * W) return;
  const int j = static_cast<int>(i / W), c = static_cast<int>(i % W);
  const int k = c / F, f = c - k * F;
  const int t = j * sub - ctx + k;
  out[i] = (t >= 0 && t < T) ? y[static_cast<long long>(t) * F + f] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// Recurrent retention step (LS:ret:126-144, decay = 1): per (sequence, head) the fp32 state kv (64 x 64) becomes
// kv * sqrt(t)/sqrt(t+1) + k^T v / sqrt(t+1)  (t = frames already seen), out = q kv -> group norm (eps 1e-6) ->
// * swish(g).  qkvg: [n_seq][1024] fp16 (q | k*hd^-.5 | v | g); state: [n_seq][4][64][64] fp32; out: [n_seq][256].
__global__ void __launch_bounds__(64)
ret_step_kernel(const __half* __restrict__ qkvg, float* __restrict__ state, int t_arg, const int* __restrict__ t_dev,
                __half* __restrict__ out) {
  const int n = blockIdx.x, h = blockIdx.y, d = threadIdx.x;   // thread d owns column d of the state
  const int t = t_dev ? *t_dev : t_arg;                        // device-resident frame counter (CUDA-graph replay)
  __shared__ float q_s[64], k_s[64], red[2];
  const __half* row = qkvg + static_cast<size_t>(n) * 1024 + h * 64;
  q_s[d] = __half2float(row[d]);
  k_s[d] = __half2float(row[256 + d]);
  const float v = __half2float(row[512 + d]);
  const float g = __half2float(row[768 + d]);
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
  // group norm over the 64 columns of this head
  float s = o;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((d & 31) == 0) red[d >> 5] = s;
  __syncthreads();
  const float mean = (red[0] + red[1]) * (1.f / 64.f);
  __syncthreads();
  float dv = (o - mean) * (o - mean);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) dv += __shfl_xor_sync(0xffffffffu, dv, off);
  if ((d & 31) == 0) red[d >> 5] = dv;
  __syncthreads();
  const float rstd = rsqrtf((red[0] + red[1]) * (1.f / 64.f) + 1e-6f);
  out[static_cast<size_t>(n) * 256 + h * 64 + d] = __float2half_rn((o - mean) * rstd * (g / (1.f + __expf(-g))));
}

}  // namespace

void launch_prep_input(const float* x, const int* cu_seqlens, int B, int Tmax, int Din, int Kpad, const float* sc,
                       const float* sh, __half* out, cudaStream_t stream, float pad_value) {
  prep_input_kernel<<<B * Tmax, 192, 0, stream>>>(x, cu_seqlens, Tmax, Din, Kpad, sc, sh, pad_value, out);
}

int launch_spk_attn(const __half* qkv, __half* out, int n_frames, int S, float scale, cudaStream_t stream) {
  if (S < 1 || S > kMaxS) return -1;
  const int fpb = 128 / (4 * S);
  const int smem = fpb * S * 64 * 16;
  const int grid = (n_frames + fpb - 1) / fpb;
  spk_attn_kernel<<<grid, 128, smem, stream>>>(qkv, out, n_frames, S, scale, fpb);
  return 0;
}

void launch_head(const __half* emb, const __half* att, int n_frames, int S, float* logits, float* emb_f32,
                 float* att_f32, cudaStream_t stream) {
  const int warps_per_block = 8;
  const int grid = (n_frames + warps_per_block - 1) / warps_per_block;
  head_kernel<<<grid, warps_per_block * 32, 0, stream>>>(emb, att, n_frames, S, logits, emb_f32, att_f32);
}

void launch_step_attn(const __half* qkv, __half* kcache, __half* vcache, int n_seq, int cap, int pos, float scale,
                      __half* out, cudaStream_t stream, const int* pos_dev) {
  step_attn_kernel<<<dim3(n_seq, 4), 128, 0, stream>>>(qkv, kcache, vcache, cap, pos, pos_dev, scale, out);
}

// counters[i] += inc[i] (i < 4): advances the device-resident frame counters at the end of a streaming step
__global__ void advance_counters_kernel(int* counters, int4 inc) {
  if (threadIdx.x == 0) {
    counters[0] += inc.x;
    counters[1] += inc.y;
    counters[2] += inc.z;
    counters[3] += inc.w;
  }
}
void launch_advance_counters(int* counters, int i0, int i1, int i2, int i3, cudaStream_t stream) {
  advance_counters_kernel<<<1, 32, 0, stream>>>(counters, make_int4(i0, i1, i2, i3));
}

void launch_splice_subsample(const float* y, int T, int F, int ctx, int sub, float* out, cudaStream_t stream) {
  const int T_out = (T + sub - 1) / sub;
  const long long n = static_cast<long long>(T_out) * (2 * ctx + 1) * F;
  if (n == 0) return;
  splice_subsample_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(y, T, F, ctx, sub, T_out, out);
}

void launch_decide_median(const float* pred, int T, int C, float threshold, int median, unsigned char* out,
                          cudaStream_t stream) {
  const long long n = static_cast<long long>(T) * C;
  if (n == 0) return;
  decide_median_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(pred, T, C, threshold, median, out);
}

void launch_hist_append(const __half* src, __half* hist, int n_seq, int cap, int pos, cudaStream_t stream,
                        const int* pos_dev) {
  hist_append_kernel<<<n_seq, 128, 0, stream>>>(src, hist, cap, pos, pos_dev);
}

int launch_dwconv_bn_swish(const __half* u, const float* w, const float* sc, const float* sh, int n_seq, int T, int K,
                           __half* hist, __half* out, cudaStream_t stream) {
  if (K < 1 || K > 32) return -1;
  if (K == 16 && hist == nullptr && T >= 32) {
    dwconv16_bn_swish_kernel<<<dim3((T + 31) / 32, n_seq), 256, 0, stream>>>(u, w, sc, sh, T, out);
    return 0;
  }
  dwconv_bn_swish_kernel<<<dim3((T + 63) / 64, n_seq), 128, 0, stream>>>(u, w, sc, sh, T, K, hist, out);
  return 0;
}

void launch_ret_step(const __half* qkvg, float* state, int n_seq, int t, __half* out, cudaStream_t stream,
                     const int* t_dev) {
  ret_step_kernel<<<dim3(n_seq, 4), 64, 0, stream>>>(qkvg, state, t, t_dev, out);
}

}  // namespace fseend
