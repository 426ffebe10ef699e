// Causal (look-ahead `mask_delay`) multi-head self-attention along the time axis, head_dim 64.
// Serves both the encoder (S = 1) and the attractor decoder's time attention (S speaker slots):
// the packed QKV activations are [B][T][S][3*256] fp16 and sequence n = (b, s) is a strided row set.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace fseend {

struct AttnParams {
  int B, S, T, H;    // H heads of 64
  int mask_delay;    // key j visible to query i iff j <= i + mask_delay (and j < T)
  float scale;       // applied to q.k (hd^-0.5)
};

// tmQKV: 4-D (768, S, T, B) box (64,1,128,1);  tmO: 4-D (256, S, T, B) box (64,1,128,1)
void launch_causal_attn(const CUtensorMap& tmQKV, const CUtensorMap& tmO, const AttnParams& p, cudaStream_t stream);

}  // namespace fseend
