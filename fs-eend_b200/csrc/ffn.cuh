// Fused FFN + residual + LayerNorm block (see ffn.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace fseend {

struct FfnParams {
  int rows_per_seq;
  int n_seq;
  int tiles_per_seq;   // ceil(rows_per_seq / 128)
  int F;               // hidden width, multiple of 128
  float ln_eps;
  const float* b1;     // [F]
  const float* b2;     // [256]
  const float* ln_g;   // [256]
  const float* ln_b;   // [256]
  const int* seq_len;  // optional [n_seq]: rows t >= seq_len[b] are written as zeros
};

// tmX / tmO: 3-D (256, rows_per_seq, n_seq) box (64,128,1);  tmW1: 2-D (256, F) box (64,128);
// tmW2: 2-D (F, 256) box (64,128).  variant: 1 = SS, 2 = SS + 2-CTA weight multicast, 3 = TS (hidden chunk in TMEM),
// 4 = TS + multicast.
void launch_ffn(const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2, const CUtensorMap& tmO,
                const FfnParams& p, int variant, cudaStream_t stream);

// CTA-pair variant (tcgen05 cta_group::2, see ffn_pair.cu): same tmX / tmO / tmW2; tmW1 with box (64, 64).
void launch_ffn_pair(const CUtensorMap& tmX, const CUtensorMap& tmW1_box64, const CUtensorMap& tmW2,
                     const CUtensorMap& tmO, const FfnParams& p, cudaStream_t stream);

}  // namespace fseend
