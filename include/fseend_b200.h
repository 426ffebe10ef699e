/* fseend_b200 — C ABI of the B200-native FS-EEND hot path (encoder + attractor decoder forward).
 *
 * The reference (Audio-WestlakeU/FS-EEND) has no FFI: its hot path is the Python method
 *   OnlineTransformerDADiarization.test(src, ilens, max_nspks)
 *   (FS-EEND/nnet/model/onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm.py:67-84)
 * and .forward() (same file :32-65).  This header is the boundary a maintainer would bind with ctypes
 * (see INTEGRATION.md); `fs-eend_b200/nnet/` is the Python mirror of the reference's nnet/ API on top of it.
 *
 * Conventions: plain pointers and sizes only; all "dev" pointers are CUDA device pointers owned by the
 * caller; every function returns 0 on success or a negative error code and records a message retrievable
 * with fseend_last_error(); no hidden host synchronisation in the *_dev entry points; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  The library requires an sm_100a device and fails
 * loudly otherwise — there is no CPU fallback.
 */
#ifndef FSEEND_B200_H_
#define FSEEND_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSEEND_VERSION 100

enum {
  FSEEND_OK = 0,
  FSEEND_ERR_INVALID = -1,      /* bad argument / unsupported shape */
  FSEEND_ERR_CUDA = -2,         /* CUDA runtime / driver error      */
  FSEEND_ERR_MISSING = -3,      /* a required state_dict tensor is missing or has the wrong size */
  FSEEND_ERR_NO_DEVICE = -4     /* no sm_100 device                 */
};

/* Constructor kwargs of OnlineTransformerDADiarization (reference model file :11). */
typedef struct fseend_fs_config {
  int in_size;              /* 345 = (2*7+1)*23 spliced log-mel                          */
  int n_units;              /* 256 (only supported width: one LayerNorm row per MMA tile) */
  int n_heads;              /* n_units / n_heads must be 64                               */
  int enc_n_layers;
  int dec_n_layers;
  int enc_dim_feedforward;  /* 2048: nn.TransformerEncoderLayer default (reference :147)  */
  int dec_dim_feedforward;
  int conv_kernel;          /* 2*conv_delay+1 = 19                                        */
  int conv_padding;         /* 9 (hard-coded in the reference, :30)                       */
  int mask_delay;
  int has_mask;             /* 0: encoder attends to all frames (decoder is always causal, reference :116) */
  float bn_eps;             /* 1e-5 */
  float ln_eps;             /* 1e-5 */
} fseend_fs_config;

typedef struct fseend_fs_model fseend_fs_model;

int fseend_version(void);
const char* fseend_last_error(void);
/* 1 if the current CUDA device is compute capability 10.x, else 0. */
int fseend_device_ok(void);

/* Build a model from a reference state_dict: `names[i]` are the reference's own keys
 * ("enc.bn.weight", "enc.transformer_encoder.layers.0.self_attn.in_proj_weight", ..., "cnn.weight",
 * "dec.convert.weight", "dec.attractor_decoder.layers.1.norm22.bias"), `data[i]` HOST fp32 pointers,
 * `numel[i]` element counts.  Weights are converted once (fp16 operands, BatchNorm folded to a
 * per-channel affine, conv taps split, attractor-init weight split) and uploaded. */
int fseend_fs_create(const fseend_fs_config* cfg, int n_tensors, const char* const* names,
                     const float* const* data, const long long* numel, fseend_fs_model** out);
void fseend_fs_destroy(fseend_fs_model* m);

/* test(): x_packed = concatenation of the B feature matrices (sum(ilens) x in_size, fp32, device),
 * ilens on the HOST.  Outputs are padded to Tmax = max(ilens):
 *   logits [B][Tmax][max_nspks] fp32   (rows t >= ilens[b] are unspecified, as in the reference's padded tensors)
 *   emb    [B][Tmax][n_units]   fp32   or NULL
 *   att    [B][Tmax][max_nspks][n_units] fp32 (L2-normalised attractors) or NULL */
int fseend_fs_forward(fseend_fs_model* m, const float* x_packed_dev, const int* ilens_host, int B, int max_nspks,
                      float* logits_dev, float* emb_dev, float* att_dev, void* stream);

/* Same with HOST buffers (pinned or pageable): H2D copy, forward, D2H copy, stream synchronise. */
int fseend_fs_forward_host(fseend_fs_model* m, const float* x_packed_host, const int* ilens_host, int B,
                           int max_nspks, float* logits_host, float* emb_host, float* att_host);
/* Pipelined host-buffer form: enqueue one forward (H2D on a copy stream, kernels + D2H on a compute stream) and return a
 * ticket.  Two calls may be in flight, so the copy of call i+1 overlaps the kernels of call i; fseend_fs_host_wait
 * blocks until the ticket's logits are in logits_host.  x_packed_host and logits_host must stay valid until then and
 * should be pinned (cudaHostAlloc) for the copies to be asynchronous.  Logits are [B][max ilen][max_nspks]. */
int fseend_fs_forward_host_async(fseend_fs_model* m, const float* x_packed_host, const int* ilens_host, int B,
                                 int max_nspks, float* logits_host, long long* ticket);
int fseend_fs_host_wait(fseend_fs_model* m, long long ticket);

/* Tuning switches.  "ffn": 0 = two GEMM launches (hidden activations through HBM); fused FFN kernel:
 * 1 = hidden chunk via smem, 2 = 1 + 2-CTA clusters sharing weight tiles by TMA multicast, 3 = hidden chunk kept in
 * TMEM (A operand from TMEM), 4 = 3 + multicast, 5 = 3 on a CTA pair (tcgen05 cta_group::2, M = 256 across two SMs,
 * each SM holding half of every weight tile; default: measured fastest).
 * "spk": speaker-axis attention: 0 = QKV GEMM + CUDA-core attention, 1 = QKV GEMM + tcgen05 block-diagonal attention,
 * 2 = projection and attention fused in one kernel (default).
 * "host_chunks": fseend_fs_forward_host splits the batch into this many chunks of whole sequences and overlaps the
 * host->device copy of chunk i+1 with the kernels of chunk i (0 = automatic, 1 = no overlap, <= 8). */
int fseend_fs_set_option(fseend_fs_model* m, const char* key, int value);

/* Per-kernel timing of the next forward calls (CUDA events around every launch; off by default).
 * get_profile returns the number of distinct kernels; names[i] (<= 31 chars), total ms and launch count
 * accumulated since profiling was switched on. */
int fseend_fs_set_profiling(fseend_fs_model* m, int on);
int fseend_fs_get_profile(fseend_fs_model* m, int max_entries, char (*names)[32], float* total_ms, int* launches);
/* Number of kernel launches one forward issues for (B, Tmax, max_nspks). */
int fseend_fs_launches_per_forward(const fseend_fs_model* m);
/* Bytes of device workspace currently held by the model's (B, Tmax, S) plan. */
size_t fseend_fs_workspace_bytes(const fseend_fs_model* m);

/* ---- frame-by-frame streaming (reference: StreamingTransformerEDADiarization.test,
 * FS-EEND/nnet/model/streaming_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm.py:31-60) --------------------
 * A stream holds the device-resident state of B parallel recordings: projected K/V caches per layer, the encoder
 * history feeding the 19-tap look-ahead conv, and the frame counter.  The model must outlive its streams. */
typedef struct fseend_fs_stream fseend_fs_stream;
int fseend_fs_stream_create(fseend_fs_model* m, int B, int max_nspks, fseend_fs_stream** out);
void fseend_fs_stream_destroy(fseend_fs_stream* s);
/* Push one frame: x_t_dev fp32 [B][in_size], or NULL for a flush step (the reference's dummy_conv_input=True, used
 * conv_delay times at the end).  *produced = 1 and logits_dev [B][max_nspks] is written once the conv window is
 * full (from the (conv_delay+1)-th call on), else *produced = 0 (the reference returns None). */
int fseend_fs_stream_step(fseend_fs_stream* s, const float* x_t_dev, float* logits_dev, int* produced, void* stream);
int fseend_fs_stream_frames(const fseend_fs_stream* s);

/* ---- LS-EEND (Conformer-retention encoder + retention attractor decoder) -----------------------------------------
 * Reference: OnlineConformerRetentionDADiarization (LS-EEND/nnet/model/onl_conformer_retention_enc_1dcnn_tfm_
 * retention_enc_linear_non_autoreg_pos_enc_l2norm_emb_loss_mask.py:14-147), constructor kwargs :15-32. */
typedef struct fseend_ls_config {
  int in_size;                        /* 345 */
  int n_units;                        /* 256 */
  int n_heads;                        /* 4 (head_dim 64) */
  int enc_n_layers;
  int dec_n_layers;
  int feed_forward_expansion_factor;  /* 4 (conf yaml :35); only 4 is supported */
  int dec_dim_feedforward;            /* 2048 */
  int conv_kernel_size;               /* 16: causal depthwise conv of the Conformer conv module */
  int recurrent_chunk_size;           /* 500: retention chunk; inputs are zero-padded to a multiple of it */
  int conv_delay;                     /* 9: look-ahead Conv1d kernel = 2*conv_delay+1 */
} fseend_ls_config;
typedef struct fseend_ls_model fseend_ls_model;

/* names[i]: the reference's state_dict keys ("enc.encoder.layers.0.sequential.1.module.self_attn.q_proj.weight", ...). */
int fseend_ls_create(const fseend_ls_config* cfg, int n_tensors, const char* const* names, const float* const* data,
                     const long long* numel, fseend_ls_model** out);
void fseend_ls_destroy(fseend_ls_model* m);
/* Tp = max(ilens) rounded up to a multiple of recurrent_chunk_size: the padded length of every output below. */
int fseend_ls_padded_len(const fseend_ls_model* m, int max_ilen);
/* test(): x_packed [sum(ilens)][in_size] fp32 device; logits [B][Tp][max_nspks], emb [B][Tp][256] or NULL,
 * att [B][Tp][max_nspks][256] or NULL (fp32, device). */
int fseend_ls_forward(fseend_ls_model* m, const float* x_packed_dev, const int* ilens_host, int B, int max_nspks,
                      float* logits_dev, float* emb_dev, float* att_dev, void* stream);
int fseend_ls_launches_per_forward(const fseend_ls_model* m);
/* Host-buffer form of fseend_ls_forward (packed features and outputs in host memory; H2D + forward + D2H inside the
 * call, synchronous on return): what bench.py's end-to-end leg times for LS-EEND.  Outputs are padded to
 * Tp = fseend_ls_padded_len(max ilen). */
int fseend_ls_forward_host(fseend_ls_model* m, const float* x_packed_host, const int* ilens_host, int B, int max_nspks,
                           float* logits_host, float* emb_host, float* att_host);
/* Options.  "precision": 1 (default) = parity mode — fp32 activations, split-precision (hi + lo fp16, three
 * tcgen05.mma per product) GEMMs, fp32 retention core: logits within 1e-3 max-abs of the reference on every frame
 * (LS-EEND/nnet/modules/retention.py:222-226: the eps = 1e-6 per-head group norm amplifies operand rounding);
 * 0 = throughput mode — fp16 operands / fp16 activations (median error 2-3e-4, isolated frames up to 7e-2).
 * The environment variable FSEEND_LS_PRECISION=fp16|fp32 sets the default at creation.  Streams inherit the model's
 * precision when they are created. */
int fseend_ls_set_option(fseend_ls_model* m, const char* key, int value);
int fseend_ls_get_option(const fseend_ls_model* m, const char* key);

/* One-step (recurrent) LS-EEND.  State per stream: retention states + conv caches (O(1) per frame) and the encoder
 * history of the look-ahead conv.  fseend_ls_stream_step is the fused loop body of streaming_predict
 * (LS-EEND/streaming_infer_dia.py:52-97; x_t NULL = flush step); enc_step / dec_step are the split entry points behind
 * model.enc.forward_one_step (conformer/encoder.py:223-228) and model.dec.forward_one_step (model file :235-243),
 * t = 0 resets the corresponding states. */
typedef struct fseend_ls_stream fseend_ls_stream;
int fseend_ls_stream_create(fseend_ls_model* m, int B, int max_nspks, fseend_ls_stream** out);
void fseend_ls_stream_destroy(fseend_ls_stream* s);
int fseend_ls_stream_reset(fseend_ls_stream* s);
int fseend_ls_stream_step(fseend_ls_stream* s, const float* x_t_dev, float* logits_dev, int* produced, void* stream);
int fseend_ls_stream_enc_step(fseend_ls_stream* s, const float* x_t_dev, int t, float* emb_dev, void* stream);
int fseend_ls_stream_dec_step(fseend_ls_stream* s, const float* emb_dev, int t, float* att_dev, void* stream);

/* ---- single-kernel entry points (used by the parity tests; all pointers are device pointers) ---------- */

/* OUT = epilogue(A * W^T): A fp16 [n_seq][rows_per_seq][K], W fp16 [taps*N][K], fp32 accumulate.
 * mode 0: +bias (+ReLU), N multiple of 256 | 1: LayerNorm(+bias +residual) | 2: L2-normalise(+bias)
 * | 3: attractor init, out[row][s][:] = acc + pe_proj[s][:].   taps>1: row-shifted accumulation (Conv1d). */
int fseend_op_gemm(const void* a_f16, int rows_per_seq, int n_seq, int K, const void* w_f16, int N, int taps,
                   int tap_shift, int mode, int relu, const float* bias, const void* residual_f16,
                   const float* ln_g, const float* ln_b, float ln_eps, const float* pe_proj, int S,
                   const int* seq_len_dev, void* out_f16, void* stream);
/* OUT = LayerNorm(X + relu(X W1^T + b1) W2^T + b2): X fp16 [n_seq][rows_per_seq][256], W1 fp16 [F][256],
 * W2 fp16 [256][F]; the F-wide hidden activations stay on chip.  cluster = kernel variant 1..5 (see "ffn" option). */
int fseend_op_ffn(const void* x_f16, int rows_per_seq, int n_seq, const void* w1_f16, const float* b1,
                  const void* w2_f16, const float* b2, int F, const float* ln_g, const float* ln_b, float ln_eps,
                  const int* seq_len_dev, int cluster, void* out_f16, void* stream);
/* qkv fp16 [B][T][S][768] -> out fp16 [B][T][S][256]; key j visible to query i iff j <= i + mask_delay. */
int fseend_op_causal_attn(const void* qkv_f16, int B, int T, int S, int H, int mask_delay, float scale,
                          void* out_f16, void* stream);
/* Speaker-axis attention, qkv fp16 [n_frames][S][768] -> out fp16 [n_frames][S][256]: CUDA-core kernel and the
 * tcgen05 block-diagonal variant (the one the model uses). */
int fseend_op_spk_attn(const void* qkv_f16, int n_frames, int S, float scale, void* out_f16, void* stream);
int fseend_op_spk_attn_tc(const void* qkv_f16, int n_frames, int S, float scale, void* out_f16, void* stream);
/* The model's speaker-axis attention: QKV projection (in_proj weight fp16 [768][256], bias fp32 [768]) fused with the
 * attention over the S slots of each frame; x fp16 [n_frames][S][256] -> out fp16 [n_frames][S][256] (heads concatenated,
 * before the out-projection).  Reference: _sa_block2 of TransformerEncoderFusionLayer (merge_tfm_encoder.py:366-372). */
int fseend_op_spk_qkv_attn(const void* x_f16, const void* w_f16, const float* bias, int n_frames, int S, float scale,
                           void* out_f16, void* stream);
int fseend_op_head(const void* emb_f16, const void* att_f16, int n_frames, int S, float* logits, float* emb_f32,
                   float* att_f32, void* stream);
int fseend_op_prep_input(const float* x_packed, const int* cu_seqlens_dev, int B, int Tmax, int Din, int Kpad,
                         const float* scale, const float* shift, void* out_f16, void* stream);

/* Embedding-consistency loss (reference FS model file :46-57; LS model file :92-113):
 *   *loss_dev = sum_{b,i,j} ( <e_i,e_j>/(|e_i||e_j| + 1e-6) - <l_i,l_j>/(|l_i||l_j| + 1e-6) )^2 / divisor
 * emb fp32 [B][T][256] (16-byte aligned), labels fp32 [B][T][S] zero padded, S <= 16.  seq_len_dev NULL: all T rows of
 * every sequence count (FS-EEND: divisor = B*T*T); otherwise only rows/cols < seq_len[b] (LS-EEND: divisor = sum len^2).
 * workspace: fseend_op_embloss_workspace_bytes(B, T) bytes of device memory.  Deterministic (fixed-order reduction). */
size_t fseend_op_embloss_workspace_bytes(int B, int T);
int fseend_op_embloss(const float* emb_f32, const float* labels, const int* seq_len_dev, int B, int T, int S,
                      double divisor, float* workspace, float* loss_dev, void* stream);

/* Training / validation step label pipeline (reference train/oln_tfm_enc_dec.py:51-76): speaker columns of the padded
 * 0/1 activity labels [B][T][n_spk] re-ordered by first appearance (stable for ties; never-active speakers last), a
 * silence column in front (1 - max) and a "no speaker" column of zeros behind: labels_out [B][T][n_spk + 2];
 * perm [B][n_spk] = source column of speaker k.  n_spk <= 14.  All pointers device pointers. */
int fseend_op_label_prepare(const float* labels, int B, int T, int n_spk, int* perm, float* labels_out, void* stream);
/* standard_loss (reference train/utils/loss.py:119-125): sum_b [ sum_{t=delay}^{len_b-1} sum_{c<n_cls_b}
 * BCEWithLogits(logits[b][t][c], target[b][t-delay][c]) / n_cls_b ] / (sum_b len_b - delay * B).
 * logits [B][T][ld_logits], target [B][T][ld_target] fp32, lens / n_cls device int [B];
 * workspace: fseend_op_bce_loss_workspace_bytes(B, T).  Deterministic (fixed-order reduction). */
size_t fseend_op_bce_loss_workspace_bytes(int B, int T);
int fseend_op_bce_loss(const float* logits, int ld_logits, const float* target, int ld_target, int B, int T,
                       const int* lens_dev, const int* n_cls_dev, int label_delay, float* workspace, float* loss_dev,
                       void* stream);
/* Permutation-invariant-training pair costs (reference train/utils/loss.py:69-96 pit_loss / :98-116 batch_pit_loss,
 * :257-327 batch_pit_n_speaker_loss, :329-403 its label-delay form): cost[b][i][j] = sum over frames t < len_b - delay
 * of BCEWithLogits(logits[b][t + delay][i], labels[b][t][j]); pad_term != 0 adds (T - len_b) * BCE(-1, -1), the
 * contribution of the reference's -1-padded frames.  logits, labels: fp32 [B][T][C] on the device, lens: int [B] on the
 * device, cost: fp64 [B][C][C] on the device (fixed-order reduction).  The permutation search over the C x C costs is the
 * host-side part (fseend_b200.loss.batch_pit_loss / batch_pit_n_speaker_loss). */
int fseend_op_pit_costs(const float* logits, const float* labels, int B, int T, int C, const int* lens_dev,
                        int label_delay, int pad_term, double* cost_dev, void* stream);

/* Feature front-end tail (reference datasets/feature.py: splice :111-133 then subsample :103-108, as called at
 * :259-261 / :348-352): feat fp32 [T][F] (e.g. 23-dim log-mel) -> out fp32 [ceil(T / subsampling)][(2 context_size + 1) F],
 * out[j][k F + f] = feat[j subsampling - context_size + k][f], zero outside the signal.  Device pointers; bit-exact. */
int fseend_op_splice_subsample(const float* feat, int T, int F, int context_size, int subsampling, float* out,
                               void* stream);

/* Post-processing in front of the RTTM writer (reference train/utils/make_rttm.py:10-15, metrics.py:58-60):
 * decisions[t][c] = medfilt(pred > threshold, (median, 1))[t][c] — threshold, then a zero-padded median filter of odd
 * width along time (median <= 1: no filter).  pred fp32 [T][C] (sigmoid posteriors), decisions uint8 [T][C], both on the
 * device.  Bit-exact with the reference (the only arithmetic is the comparison). */
int fseend_op_decide_median(const float* pred, int T, int C, float threshold, int median, unsigned char* decisions,
                            void* stream);

/* GEMM with the LS-EEND epilogues: mode 0 (+bias, act 0 none / 1 ReLU / 2 swish), 4 (GLU: N/2 outputs), 1 (LayerNorm),
 * 5 (y = residual + alpha*(acc+bias); out = ln_g ? LN(y) : y); out2 (optional) = LayerNorm(out; ln2_g, ln2_b). */
int fseend_op_gemm_ex(const void* a_f16, int rows_per_seq, int n_seq, int K, const void* w_f16, int N, int mode, int act,
                      const float* bias, const void* residual_f16, float alpha, const float* ln_g, const float* ln_b,
                      const float* ln2_g, const float* ln2_b, float ln_eps, const int* seq_len_dev, void* out_f16,
                      void* out2_f16, void* stream);
/* Chunkwise retention + group norm + swish gate: qkvg fp16 [B][T][S][1024] (q | k*hd^-.5 | v | g) -> out fp16
 * [B][T][S][256].  state fp16 [B*S*4*(T/chunk)][64][64] and cross_scale fp32 [B*S*4*(T/chunk)] are workspaces. */
int fseend_op_retention(const void* qkvg_f16, int B, int S, int T, int chunk, void* state_f16, float* cross_scale,
                        void* out_f16, void* stream);
/* Causal depthwise conv (weight fp32 [256][K]) -> per-channel affine -> swish; u/out fp16 [n_seq][T][256];
 * hist: optional one-step cache fp16 [n_seq][K-1][256] (updated when T == 1). */
int fseend_op_dwconv_bn_swish(const void* u_f16, const float* w, const float* scale, const float* shift, int n_seq,
                              int T, int K, void* hist_f16, void* out_f16, void* stream);
/* Recurrent retention step for 0-based frame index t: state fp32 [n_seq][4][64][64] updated in place. */
int fseend_op_ret_step(const void* qkvg_f16, float* state, int n_seq, int t, void* out_f16, void* stream);

/* Parity-precision kernels (csrc/p32.cu).
 * fseend_p32_linear_*: a linear layer with split-precision weights (w fp32 [N][K] in HOST memory, split into fp16
 *   hi + lo once at creation; K % 64 == 0, N % 128 == 0).  apply: out[rows][N] = residual + alpha * act(a[rows][K] w^T
 *   + bias), fp32 device buffers, asynchronous on `stream`; act: 0 none, 1 ReLU, 2 swish.  Replaces nn.Linear /
 *   nn.Conv1d-as-GEMM call sites where the fp16 pipeline is not accurate enough (StreamingConv1d of the LS-EEND model
 *   file :151-186 in parity mode).
 * fseend_op_p32_retention: qkvg fp32 [B][T][S][1024] -> out fp32 [B][T][S][256] = swish(g) * GroupNorm(retention),
 *   LS-EEND/nnet/modules/retention.py:146-194,222-224 in fp32. */
typedef struct fseend_p32_linear fseend_p32_linear;
int fseend_p32_linear_create(const float* w_host, int N, int K, fseend_p32_linear** out);
void fseend_p32_linear_destroy(fseend_p32_linear* h);
int fseend_p32_linear_apply(const fseend_p32_linear* h, const float* a_dev, int rows, const float* bias_dev, int act,
                            float alpha, const float* residual_dev, float* out_dev, void* stream);
int fseend_op_p32_retention(const float* qkvg_dev, int B, int S, int T, int chunk, float* out_dev, void* stream);

/* Training building blocks (csrc/train_ops.cu, csrc/train_attn.cu) — SURVEY.md §8f "N1", STARTED: forward + backward of
 * the operators that carry the FS-EEND encoder layer (nn.TransformerEncoderLayer, post-norm, ReLU; constructed at
 * FS-EEND/nnet/model/onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm.py:147, looped at
 * nnet/modules/transformer_encoder_fusion.py:129-131), in the parity arithmetic (fp32 tensors, split-precision tcgen05
 * products, fixed-order reductions).  All pointers are DEVICE fp32; calls are asynchronous on `stream`.  The weights
 * are re-split on the device on every call (they change every optimizer step); |w| must stay below 1000.
 * The torch.autograd.Functions over these live in fseend_b200/autograd.py; gradients are pinned against torch autograd
 * in tests/test_train_ops_gpu.py.  There is no optimizer / DDP loop here. */
size_t fseend_train_linear_workspace_bytes(int rows, int K, int N);
/* y[rows][N] = act(x[rows][K] w[N][K]^T + bias); act 0 none, 1 ReLU; N % 128 == 0, any K. */
int fseend_train_linear_fwd(const float* x, int rows, int K, const float* w, int N, const float* bias, int act, float* y,
                            void* workspace, size_t ws_bytes, void* stream);
/* dx[rows][K] (nullable) = dy' w;  dw[N][K] = dy'^T x;  db[N] (nullable) = column sums of dy';  dy' = dy masked by
 * y > 0 when act == 1 (y = the saved forward output, else may be null).  relu_input != 0: x is the output of a ReLU
 * and dx is zeroed where x <= 0 (that ReLU's backward, fused; K % 128 == 0). */
int fseend_train_linear_bwd(const float* x, const float* w, const float* y, const float* dy, int rows, int K, int N,
                            int act, int relu_input, float* dx, float* dw, float* db, void* workspace, size_t ws_bytes,
                            void* stream);
/* y = LayerNorm(x + r; g, b), rows of 256, biased variance, eps inside the sqrt; r and sum_out (= x + r) nullable. */
int fseend_train_add_layernorm_fwd(const float* x, const float* r, const float* g, const float* b, int rows, float eps,
                                   float* sum_out, float* y, void* stream);
size_t fseend_train_layernorm_workspace_bytes(int rows);
/* LayerNorm backward: x = the normalised tensor's input (x + r of the forward). */
int fseend_train_layernorm_bwd(const float* x, const float* g, const float* dy, int rows, float eps, float* dx, float* dg,
                               float* db, void* workspace, size_t ws_bytes, void* stream);
/* Causal 4-head self-attention on projected qkv fp32 [n_seq][T][768] -> out fp32 [n_seq][T][256]; key j is visible to
 * query i iff j <= i + mask_delay (FS model file :152-155); lse fp32 [n_seq][4][T] is saved for the backward.
 * dropout_p / seed: attention-probability dropout (nn.MultiheadAttention's `dropout`, on the softmax output, kept
 * elements scaled by 1/(1-p)); the mask is a counter-based hash of (seed, sequence, head, query, key), so the backward
 * regenerates it from the same (dropout_p, seed).  dropout_p = 0 disables it.
 * seq_inner: 1 for the plain layout above; S for an interleaved batch [n_seq / S][T][S][768] -> [n_seq / S][T][S][256] (the
 * attractor decoder's (B, T, S, D) tensor, whose time attention runs per (b, s) — FS-EEND/nnet/modules/merge_tfm_encoder.py
 * :358-365 transposes instead): sequence n = b * S + s is read and written in place with row stride S, no layout copy. */
int fseend_train_attn_fwd(const float* qkv, int n_seq, int T, int seq_inner, int mask_delay, float dropout_p,
                          unsigned long long seed, float* out, float* lse, void* stream);
/* dqkv fp32 [n_seq][T][768] from dout [n_seq][T][256]; dsum: scratch fp32 [n_seq * 4 * T + 16]. */
int fseend_train_attn_bwd(const float* qkv, const float* out, const float* dout, const float* lse, int n_seq, int T,
                          int seq_inner, int mask_delay, float dropout_p, unsigned long long seed, float* dqkv, float* dsum,
                          void* stream);
/* Speaker-axis attention (FS-EEND/nnet/modules/transformer_encoder_fusion.py:390 self_attn2: S x S per frame, no mask)
 * on projected qkv fp32 [n_frames][S][768] -> out fp32 [n_frames][S][256]; S <= 16.  The backward recomputes the
 * probabilities: dqkv fp32 [n_frames][S][768] from dout [n_frames][S][256].  Dropout as above. */
int fseend_train_spk_attn_fwd(const float* qkv, int n_frames, int S, float dropout_p, unsigned long long seed, float* out,
                              void* stream);
int fseend_train_spk_attn_bwd(const float* qkv, const float* dout, int n_frames, int S, float dropout_p,
                              unsigned long long seed, float* dqkv, void* stream);

/* Row kernels of the training graph (csrc/train_ops.cu), fp32:
 *  l2norm: y = x / ||x||_2 over rows of 256 (FS model file :41,:43, no eps), inv_norm[rows] saved for the backward
 *          dx = (dy - y (y . dy)) * inv_norm;
 *  head:   y[f][s] = emb[f][:] . att[f][s][:] (:60) and its two gradients;
 *  batchnorm: nn.BatchNorm1d in training mode over x [rows][C] (:166; the -1 padding rows are part of the batch, as in
 *          the reference): stats[2C] = batch mean | biased variance (the caller updates running_mean / running_var),
 *          backward dx, dgamma, dbeta. */
int fseend_train_l2norm_fwd(const float* x, int rows, float* y, float* inv_norm, void* stream);
int fseend_train_l2norm_bwd(const float* y, const float* inv_norm, const float* dy, int rows, float* dx, void* stream);
int fseend_train_head_fwd(const float* emb, const float* att, int frames, int S, float* y, void* stream);
int fseend_train_head_bwd(const float* emb, const float* att, const float* dy, int frames, int S, float* demb, float* datt,
                          void* stream);
size_t fseend_train_batchnorm_workspace_bytes(int rows, int C);
int fseend_train_batchnorm_fwd(const float* x, const float* g, const float* b, int rows, int C, float eps, float* y,
                               float* stats, void* workspace, size_t ws_bytes, void* stream);
int fseend_train_batchnorm_bwd(const float* x, const float* g, const float* stats, const float* dy, int rows, int C,
                               float eps, float* dx, float* dg, float* db, void* workspace, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSEEND_B200_H_ */
