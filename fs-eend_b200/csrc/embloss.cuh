// Embedding-consistency loss kernel (see embloss.cu).
#pragma once
#include <cuda_runtime.h>

namespace fseend {

struct EmbLossParams {
  const float* emb;      // [B][T][256] fp32 (L2-normalised embeddings as returned by the forward)
  const float* labels;   // [B][T][S] fp32, zero padded (rows >= ilen and speakers >= n_spk[b])
  const int* seq_len;    // optional [B]: only rows/cols < seq_len[b] contribute (LS-EEND masking); nullptr = all T
  float* partials;       // workspace, embloss_num_partials(B, T) floats
  int B, T, S;
  int n_pairs;           // filled by launch_embloss
};

int embloss_num_partials(int B, int T);
// *loss = sum of squared differences / divisor
void launch_embloss(EmbLossParams p, double divisor, float* loss, cudaStream_t stream);

}  // namespace fseend
