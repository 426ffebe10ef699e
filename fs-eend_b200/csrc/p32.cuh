// Parity-precision ("p32") kernels: fp32 activations in HBM end to end, every matrix product on tcgen05 with
// split-precision operands (x = hi + lo, both fp16; x.w ~= hi.hi + lo.hi + hi.lo, three MMAs, fp32 accumulate in TMEM:
// ~22 operand mantissa bits), retention core / normalisations / activations in fp32 on CUDA cores.
//
// Why it exists: LS-EEND's per-head group norm (eps 1e-6, LS-EEND/nnet/modules/retention.py:222-226) amplifies operand
// rounding 10^2-10^3 x at isolated frames, so the fp16-operand pipeline meets the 1e-3 logit bound only in the bulk
// (DESIGN.md §1).  This path trades ~3x tensor work and 2x activation bytes for fp32-grade results; the fp16 pipeline
// stays as the throughput mode.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fseend {

enum P32Act : int { P32_NONE = 0, P32_RELU = 1, P32_SWISH = 2 };

// OUT[row][n] = residual[row][n] + alpha * act(w_inv_scale * sum_k A[arow][k] W[tap][n][k] + bias[n])
// Rows are organised as n_seq sequences; with taps > 1 (Conv1d as shifted GEMMs) output row t of a sequence reads
// A rows t + tap + tap_shift + a_row_offset of the same sequence, zero outside [0, a_seq_rows).
struct P32GemmParams {
  const float* A;
  int lda;             // floats, multiple of 4; K = k_blocks * 64 columns are read
  int a_seq_rows;      // A rows per sequence
  int rows_per_seq;    // output rows per sequence
  int n_seq;
  int k_blocks, taps, tap_shift, a_row_offset;
  const int* a_row_offset_dev;   // optional device int added to a_row_offset (graph-replayable streaming conv)
  int N;               // output columns, multiple of 128 (W rows per tap)
  const float* bias;   // [N] or nullptr
  int act;             // P32Act
  float alpha;
  float w_inv_scale;   // the weights were multiplied by a power of two before the fp16 split
  const float* residual;
  int ldr;
  float* out;
  int ldo;
  // Row-vector kernel only (N == 256): fused row epilogue, executed by the LAST CTA of the launch to finish (threadfence +
  // atomic ticket; the counter is reset for the next launch): y1 = ln_g1 ? LN(out row; g1, b1) : out row -> ln_out1
  // (optional); ln_out2 = LN(y1; g2, b2) (optional); or, with l2norm != 0, out row /= ||out row||_2 in place.
  // Saves one dependent launch per LayerNorm in the per-frame streaming steps.
  unsigned int* row_epi_counter = nullptr;
  const float* ln_g1 = nullptr;
  const float* ln_b1 = nullptr;
  float* ln_out1 = nullptr;
  const float* ln_g2 = nullptr;
  const float* ln_b2 = nullptr;
  float* ln_out2 = nullptr;
  float ln_eps = 1e-5f;
  int l2norm = 0;
  // Training (backward) products only — any of these selects the kTrain instantiation (act must be P32_NONE):
  //  a_transposed: A is read through its transpose, A[m][k] = src[(seq * k_blocks * 64 + k) * lda + m] for k rows below
  //                a_k_rows (zero beyond): the weight-gradient product reads dY / X as they lie, no transposed copy;
  //                sequences are then split-K slices, their partial products land in out[seq][rows_per_seq][N]
  //  w_seq_stride: W rows per sequence (the TMA row coordinate is tap * N + n0 + seq * w_seq_stride)
  //  a_scale_dev : device float, A is multiplied by it before the fp16 split (a power of two: gradients are ~1e-6 and
  //                would be fp16-subnormal); out_scale_dev: device float multiplied into the epilogue (its inverse)
  //  out_mask    : same layout / ldo as out; outputs whose mask value is <= 0 are written as zero (ReLU backward fused
  //                into the product that creates the gradient of the ReLU's output)
  int a_transposed = 0;
  int a_k_rows = 0;
  int w_seq_stride = 0;
  const float* a_scale_dev = nullptr;
  const float* out_scale_dev = nullptr;
  const float* out_mask = nullptr;
};
// tmWhi / tmWlo: 2-D (K, taps*N) fp16, box (64, 128), 128B swizzle
void launch_p32_gemm(const CUtensorMap& tmWhi, const CUtensorMap& tmWlo, const P32GemmParams& p, cudaStream_t st);

// TMA-fed form for pre-split A: tmAhi / tmAlo: 2-D (K, rows) fp16, box (64, 128), 128B swizzle; one sequence, one tap,
// act none / ReLU; honours bias, alpha, w_inv_scale, residual, out_scale_dev, out_mask.
void launch_p32_gemm_planes(const CUtensorMap& tmAhi, const CUtensorMap& tmAlo, const CUtensorMap& tmWhi,
                            const CUtensorMap& tmWlo, const P32GemmParams& p, cudaStream_t st);

// The same product for at most 16 output rows (the per-frame streaming steps: 1 .. B * S rows): with so few rows the
// GEMM is a weight-streaming matrix-vector product — L2-bandwidth bound, tensor cores idle either way — so it runs on
// CUDA cores: every warp owns two output columns, lanes split K, the weights (fp16 hi [+ lo] planes, [taps*N][K]
// row-major) are read once with 16-byte loads and reconstructed as fp32 (hi + lo is exact), activations stay fp32.
// wlo == nullptr: single-plane fp16 weights (the FS-EEND streaming path).  Returns false if rows > 16.
bool launch_p32_rowvec(const __half* whi, const __half* wlo, const P32GemmParams& p, cudaStream_t st);

// y1 = g1 ? LN(x; g1, b1) : x  -> out1 (optional);  out2 = LN(y1; g2, b2) (optional).  Rows of 256 fp32.
// seq_len (optional, with rows_per_seq): rows t >= seq_len[b] are written as zeros (both outputs).
void launch_p32_layernorm(const float* x, int rows, const float* g1, const float* b1, float* out1, const float* g2,
                          const float* b2, float* out2, float eps, const int* seq_len, int rows_per_seq,
                          cudaStream_t st);
// packed fp32 rows (cu_seqlens) -> [B][Tmax][Kpad] fp32, rows t >= len and columns >= Din are zero
void launch_p32_pad_input(const float* x, const int* cu, int B, int Tmax, int Din, int Kpad, float* out, cudaStream_t st,
                          const float* sc = nullptr, const float* sh = nullptr, float pad_value = 0.f);
// out[r][c] = h[r][c] * sigmoid(h[r][256 + c]),  h: [rows][512]
void launch_p32_glu(const float* h, int rows, float* out, cudaStream_t st);
// causal depthwise conv (K <= 32) + folded BatchNorm + swish on fp32 [n_seq][T][256]; hist: optional [n_seq][K-1][256]
// one-step cache (used and slid when T == 1)
int launch_p32_dwconv_bn_swish(const float* u, const float* w, const float* sc, const float* sh, int n_seq, int T, int K,
                               float* hist, float* out, cudaStream_t st);
// x[r][:] /= ||x[r]||_2 in place (rows of 256)
void launch_p32_l2norm(float* x, int rows, cudaStream_t st);
// out[row][s][:] = y[row][:] + pe_proj[s][:]
void launch_p32_convert(const float* y, const float* pe_proj, int rows, int S, float* out, cudaStream_t st);
// speaker-axis attention on projected fp32 qkv [frames][S][768] -> [frames][S][256]
int launch_p32_spk_attn(const float* qkv, float* out, int n_frames, int S, float scale, cudaStream_t st);
// L2-normalise attractors, logits = emb . att_n; optional fp32 copies of emb / normalised attractors
void launch_p32_head(const float* emb, const float* att, int n_frames, int S, float* logits, float* emb_out,
                     float* att_out, cudaStream_t st);

// Retention (see retention.cu for the algebra).  qkvg: fp32 [B][T][S][1024] (q | k*hd^-.5 | v | g).
// Pass 1: exclusive prefix state KV_c = sum over earlier chunks of k^T v, fp32 [n][h][c][64][64], and
// cross_scale[n][h][c] = max(1, max_d sum_e |KV_c[e][d]| / sqrt(chunk)).  Pass 2: out fp32 [B][T][S][256].
void launch_p32_ret_chunk_state(const float* qkvg, int B, int S, int T, int chunk, float* state, float* cross_scale,
                                cudaStream_t st);
void launch_p32_retention(const float* qkvg, const float* state, const float* cross_scale, int B, int S, int T,
                          int chunk, float* out, cudaStream_t st);
// recurrent step: qkvg fp32 [n_seq][1024], state fp32 [n_seq][4][64][64] (updated), out fp32 [n_seq][256]
void launch_p32_ret_step(const float* qkvg, float* state, int n_seq, int t, float* out, cudaStream_t st,
                         const int* t_dev);
// Streaming attention step on fp32 (FS-EEND frame loop, FS:stream_mod:28-35): appends this frame's K / V rows of
// qkv [n_seq][768] to the fp32 caches [n_seq][cap][256] at `pos`, then attends over keys 0..pos.  out fp32 [n_seq][256].
void launch_p32_step_attn(const float* qkv, float* kcache, float* vcache, int n_seq, int cap, int pos, float scale,
                          float* out, cudaStream_t st, const int* pos_dev);
// rows [n_seq] of 256 fp32 copied (zero-filled when src == nullptr) into hist[n][pos]
void launch_p32_hist_append(const float* src, float* hist, int n_seq, int cap, int pos, cudaStream_t st,
                            const int* pos_dev);

}  // namespace fseend
