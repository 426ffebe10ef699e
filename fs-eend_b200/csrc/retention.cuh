// Chunkwise multi-scale retention of LS-EEND (decay = 1, no rotation): causal *linear* attention with the
// reference's chunk normalisation, per-head group norm and swish gate (see retention.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fseend {

struct RetParams {
  int B, S, T, H;     // sequences n = (b, s); T padded to a multiple of `chunk`; H heads of 64
  int chunk;          // recurrent_chunk_size (500)
  int n_chunks;       // T / chunk
  const float* cross_scale;   // [B*S][H][n_chunks] fp32: max(1, max_d sum_e |R'_c[e][d]|)
};

// Pass 1 (CUDA cores): per (sequence, head): running state R'_c = (sum over earlier chunks of K^T V) / sqrt(chunk)
// as fp16 [n][h][c][64 e][64 d], and cross_scale[n][h][c].  qkvg: [B][T][S][1024] fp16 (q | k*hd^-.5 | v | g).
void launch_ret_chunk_state(const __half* qkvg, const RetParams& p, __half* state, float* cross_scale,
                            cudaStream_t stream);

// Pass 2 (tcgen05): tmQKVG 5-D (1024, S, chunk, n_chunks, B) box (64,1,128,1,1); tmState 3-D (64, 64, n_states)
// box (64,64,1); tmO 5-D (256, S, chunk, n_chunks, B) box (64,1,128,1,1).  Output = swish(g) * GroupNorm(ret).
void launch_retention(const CUtensorMap& tmQKVG, const CUtensorMap& tmState, const CUtensorMap& tmO, const RetParams& p,
                      cudaStream_t stream);

}  // namespace fseend
