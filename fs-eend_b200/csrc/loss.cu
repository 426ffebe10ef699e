// Training-step label pipeline and diarization loss on the device.
// Reference: FS-EEND/train/oln_tfm_enc_dec.py:51-76 (speaker columns re-ordered by first appearance, silence and
// "no speaker" columns added) and train/utils/loss.py:119-125 (standard_loss: per-recording BCE-with-logits with a label
// delay, weighted by the recording length, divided by the total number of frames).  The reference walks Python lists
// and synchronises on np.sum; here the batch is two small kernels each, with fixed-order fp64 final reductions.
#include "loss.cuh"

#include <math.h>

namespace fseend {

namespace {

constexpr int kMaxSpk = 14;   // max_spk + 2 <= 16 attractor slots

// One block per recording: first active frame of every speaker column (1-based, +inf if never active), then the
// stable rank of the columns by that frame -> perm[b][k] = column that becomes speaker k.
__global__ void __launch_bounds__(256)
label_order_kernel(const float* __restrict__ labels, int T, int C, int* __restrict__ perm) {
  const int b = blockIdx.x, tid = threadIdx.x;
  __shared__ int first[kMaxSpk];
  if (tid < kMaxSpk) first[tid] = 0x7fffffff;
  __syncthreads();
  const float* lab = labels + static_cast<size_t>(b) * T * C;
  int mine[kMaxSpk];
#pragma unroll
  for (int c = 0; c < kMaxSpk; ++c) mine[c] = 0x7fffffff;
  for (int t = tid; t < T; t += 256) {
#pragma unroll
    for (int c = 0; c < kMaxSpk; ++c)
      if (c < C && lab[static_cast<size_t>(t) * C + c] != 0.f) mine[c] = min(mine[c], t + 1);
  }
#pragma unroll
  for (int c = 0; c < kMaxSpk; ++c)
    if (c < C && mine[c] != 0x7fffffff) atomicMin(&first[c], mine[c]);
  __syncthreads();
  if (tid < C) {
    // stable rank: columns with an earlier first frame come first, ties keep their original order
    int rank = 0;
    for (int c = 0; c < C; ++c) rank += (first[c] < first[tid]) || (first[c] == first[tid] && c < tid);
    perm[b * C + rank] = tid;
  }
}

// out[b][t][0] = 1 - max_c labels; out[b][t][1 + k] = labels[b][t][perm[b][k]]; out[b][t][C + 1] = 0
__global__ void __launch_bounds__(256)
label_build_kernel(const float* __restrict__ labels, const int* __restrict__ perm, int B, int T, int C,
                   float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * T) return;
  const int b = static_cast<int>(i / T);
  const float* src = labels + i * C;
  float* dst = out + i * (C + 2);
  float mx = 0.f;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, src[c]);
  dst[0] = 1.f - mx;
  for (int k = 0; k < C; ++k) dst[1 + k] = src[perm[b * C + k]];
  dst[C + 1] = 0.f;
}

// partial[b][chunk] = sum over the chunk's frames t in [delay, len_b) and classes c < cls_b of
// BCEWithLogits(y[b][t][c], tgt[b][t - delay][c])
__global__ void __launch_bounds__(256)
bce_partial_kernel(const float* __restrict__ logits, int ldy, const float* __restrict__ target, int ldt, int T,
                   const int* __restrict__ lens, const int* __restrict__ n_cls, int delay, int frames_per_chunk,
                   float* __restrict__ partial) {
  const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int len = min(lens[b], T), C = n_cls[b];
  const int t_lo = max(delay, chunk * frames_per_chunk), t_hi = min(len, (chunk + 1) * frames_per_chunk);
  const float* y = logits + static_cast<size_t>(b) * T * ldy;
  const float* tg = target + static_cast<size_t>(b) * T * ldt;
  float acc = 0.f;
  const int n = (t_hi - t_lo) * C;
  for (int i = tid; i < n; i += 256) {
    const int t = t_lo + i / C, c = i % C;
    const float x = y[static_cast<size_t>(t) * ldy + c];
    const float z = tg[static_cast<size_t>(t - delay) * ldt + c];
    acc += fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[b * gridDim.x + chunk] = s;
  }
}

// loss = sum_b (sum_chunk partial[b][chunk]) / n_cls[b]  /  (sum_b len_b - delay * B)
__global__ void bce_finalize_kernel(const float* __restrict__ partial, int B, int n_chunks, const int* __restrict__ lens,
                                    const int* __restrict__ n_cls, int T, int delay, float* __restrict__ loss) {
  if (threadIdx.x != 0) return;
  double tot = 0.0;
  long long frames = 0;
  for (int b = 0; b < B; ++b) {
    double s = 0.0;
    for (int k = 0; k < n_chunks; ++k) s += static_cast<double>(partial[b * n_chunks + k]);
    tot += s / static_cast<double>(n_cls[b]);
    frames += min(lens[b], T);
  }
  frames -= static_cast<long long>(delay) * B;
  *loss = static_cast<float>(tot / static_cast<double>(frames));
}

// PIT pair costs (reference train/utils/loss.py:69-96 pit_loss, :257-327 batch_pit_n_speaker_loss, :329-403 its
// label-delay form): cost[b][i][j] = sum over frames t in [0, len_b - delay) of BCEWithLogits(y[b][t + delay][i],
// lab[b][t][j]) (+ pad_frames_b * BCEWithLogits(-1, -1) when `pad_term`: the reference pads predictions AND labels of
// shorter recordings with -1 and sums over the padded frames too).  Element-wise fp32 as torch computes it, fixed-order
// fp64 accumulation (bit-reproducible).  One block per (b, i * C + j).
__global__ void __launch_bounds__(256)
pit_cost_kernel(const float* __restrict__ logits, const float* __restrict__ labels, int T, int C,
                const int* __restrict__ lens, int delay, int pad_term, double* __restrict__ cost) {
  const int b = blockIdx.y, i = blockIdx.x / C, j = blockIdx.x % C, tid = threadIdx.x;
  const int len = min(lens[b], T);
  const int n = max(len - delay, 0);
  const float* y = logits + static_cast<size_t>(b) * T * C;
  const float* z = labels + static_cast<size_t>(b) * T * C;
  double acc = 0.0;
  for (int t = tid; t < n; t += 256) {
    const float x = y[static_cast<size_t>(t + delay) * C + i];
    const float tg = z[static_cast<size_t>(t) * C + j];
    acc += static_cast<double>(fmaxf(x, 0.f) - x * tg + log1pf(expf(-fabsf(x))));
  }
  __shared__ double red[256];
  red[tid] = acc;
  __syncthreads();
  for (int sft = 128; sft > 0; sft >>= 1) {
    if (tid < sft) red[tid] += red[tid + sft];
    __syncthreads();
  }
  if (tid == 0) {
    double v = red[0];
    if (pad_term) v += static_cast<double>(T - len) * static_cast<double>(0.f - (-1.f) * (-1.f) + log1pf(expf(-1.f)));
    cost[(static_cast<size_t>(b) * C + i) * C + j] = v;
  }
}

}  // namespace

int launch_pit_costs(const float* logits, const float* labels, int B, int T, int C, const int* lens, int delay,
                     int pad_term, double* cost, cudaStream_t stream) {
  if (C < 1 || C > 16 || B < 1 || T < 1 || delay < 0) return -1;
  pit_cost_kernel<<<dim3(C * C, B), 256, 0, stream>>>(logits, labels, T, C, lens, delay, pad_term, cost);
  return 0;
}

int launch_label_prepare(const float* labels, int B, int T, int C, int* perm, float* out, cudaStream_t stream) {
  if (C < 1 || C > kMaxSpk) return -1;
  label_order_kernel<<<B, 256, 0, stream>>>(labels, T, C, perm);
  const long long n = static_cast<long long>(B) * T;
  label_build_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(labels, perm, B, T, C, out);
  return 0;
}

int bce_loss_chunks(int T) { return (T + 255) / 256; }

void launch_bce_loss(const float* logits, int ldy, const float* target, int ldt, int B, int T, const int* lens,
                     const int* n_cls, int delay, float* partial, float* loss, cudaStream_t stream) {
  const int n_chunks = bce_loss_chunks(T);
  bce_partial_kernel<<<dim3(n_chunks, B), 256, 0, stream>>>(logits, ldy, target, ldt, T, lens, n_cls, delay, 256, partial);
  bce_finalize_kernel<<<1, 32, 0, stream>>>(partial, B, n_chunks, lens, n_cls, T, delay, loss);
}

}  // namespace fseend
