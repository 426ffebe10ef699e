// Bandwidth-bound kernels of the FS-EEND hot path (see elementwise.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fseend {

// x: packed fp32 rows [sum(len)][Din]; cu_seqlens: device int[B+1]; out: fp16 [B][Tmax][Kpad].
// out = x * sc + sh (BatchNorm eval folded into per-channel scale/shift; nullptr = identity), rows t >= len use
// x = pad_value (FS-EEND pads with -1 before BatchNorm, LS-EEND with 0).
void launch_prep_input(const float* x, const int* cu_seqlens, int B, int Tmax, int Din, int Kpad, const float* sc,
                       const float* sh, __half* out, cudaStream_t stream, float pad_value = -1.f);

// qkv: [n_frames][S][768] fp16 -> out [n_frames][S][256] fp16; S <= 16.  Returns -1 on unsupported S.
int launch_spk_attn(const __half* qkv, __half* out, int n_frames, int S, float scale, cudaStream_t stream);

// emb: [n_frames][256], att: [n_frames][S][256] (un-normalised) -> logits [n_frames][S] fp32;
// optional fp32 copies of emb and of the L2-normalised attractors.
void launch_head(const __half* emb, const __half* att, int n_frames, int S, float* logits, float* emb_f32,
                 float* att_f32, cudaStream_t stream);

// Streaming step: append this frame's K/V (from qkv [n_seq][768]) to the caches [n_seq][cap][256] at `pos`, then
// attend the single query row over keys 0..pos.  out: [n_seq][256] fp16.
// pos_dev (optional): device int holding `pos` (then `pos` is ignored): the launch can be replayed from a CUDA graph.
void launch_step_attn(const __half* qkv, __half* kcache, __half* vcache, int n_seq, int cap, int pos, float scale,
                      __half* out, cudaStream_t stream, const int* pos_dev = nullptr);
// counters[i] += inc_i: device-resident frame counters of a streaming step
void launch_advance_counters(int* counters, int i0, int i1, int i2, int i3, cudaStream_t stream);
// hist[n][pos][:] = src[n][:] (or zeros when src == nullptr); hist: [n_seq][cap][256] fp16.
void launch_hist_append(const __half* src, __half* hist, int n_seq, int cap, int pos, cudaStream_t stream,
                        const int* pos_dev = nullptr);

// Conformer conv-module middle: causal depthwise conv (K <= 32 taps, weight [256][K]) -> BN(eval) affine -> swish.
// u/out: [n_seq][T][256] fp16; hist: optional [n_seq][K-1][256] one-step cache (updated when T == 1).
int launch_dwconv_bn_swish(const __half* u, const float* w, const float* sc, const float* sh, int n_seq, int T, int K,
                           __half* hist, __half* out, cudaStream_t stream);
// out[j][k * F + f] = y[j * sub - ctx + k][f] (zero outside the signal); out: [ceil(T / sub)][(2 ctx + 1) * F] fp32.
void launch_splice_subsample(const float* y, int T, int F, int ctx, int sub, float* out, cudaStream_t stream);
// decisions[t][c] = median filter (odd width, zero padded, along t) of (pred[t][c] > threshold); pred [T][C] fp32.
void launch_decide_median(const float* pred, int T, int C, float threshold, int median, unsigned char* out,
                          cudaStream_t stream);
// Recurrent retention step for frame index t (0-based): state [n_seq][4][64][64] fp32 updated in place.
// t_dev (optional): device int holding t (then `t` is ignored).
void launch_ret_step(const __half* qkvg, float* state, int n_seq, int t, __half* out, cudaStream_t stream,
                     const int* t_dev = nullptr);

}  // namespace fseend
