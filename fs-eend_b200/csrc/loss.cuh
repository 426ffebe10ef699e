// Training-step label pipeline and diarization loss kernels (see loss.cu).
#pragma once
#include <cuda_runtime.h>

namespace fseend {

// labels [B][T][C] fp32 0/1 (zero padded) -> perm [B][C] (column that becomes speaker k, by first appearance, stable)
// and out [B][T][C + 2] = silence | re-ordered speakers | zeros.  Returns -1 if C is out of range (1..14).
int launch_label_prepare(const float* labels, int B, int T, int C, int* perm, float* out, cudaStream_t stream);

int bce_loss_chunks(int T);   // partial sums per recording
// logits [B][T][ldy], target [B][T][ldt] fp32; lens [B], n_cls [B] device ints; partial: B * bce_loss_chunks(T) floats.
void launch_bce_loss(const float* logits, int ldy, const float* target, int ldt, int B, int T, const int* lens,
                     const int* n_cls, int delay, float* partial, float* loss, cudaStream_t stream);

// PIT pair costs: logits, labels [B][T][C] fp32 (zero padded), lens [B] device ints -> cost [B][C][C] fp64 with
// cost[b][i][j] = sum_t BCEWithLogits(logits[b][t + delay][i], labels[b][t][j]), t < len_b - delay; pad_term adds the
// reference's (T - len_b) * BCE(-1, -1) of its -1-padded batches.  Returns -1 on bad arguments (C in 1..16).
int launch_pit_costs(const float* logits, const float* labels, int B, int T, int C, const int* lens, int delay,
                     int pad_term, double* cost, cudaStream_t stream);

}  // namespace fseend
