// Multi-head self-attention, head_dim 64, on tcgen05 (see attn.cu).
//  * ATTN_CAUSAL: attention along the time axis with look-ahead `mask_delay`.  Serves the encoder (S = 1) and the
//    attractor decoder's time attention (S speaker slots): the packed QKV activations are [B][T][S][3*256] fp16
//    and sequence n = (b, s) is a strided row set.
//  * ATTN_BLOCKDIAG: attention along the speaker axis: rows [frames*S][3*256]; each row attends to the S rows of
//    its own frame.  T = total rows, tile_rows = (128 / S) * S rows per CTA.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fseend {

enum AttnMode : int { ATTN_CAUSAL = 0, ATTN_BLOCKDIAG = 1 };

struct AttnParams {
  int B, S, T, H;    // H heads of 64
  int mask_delay;    // causal: key j visible to query i iff j <= i + mask_delay (and j < T)
  float scale;       // applied to q.k (hd^-0.5)
  int mode;          // AttnMode
  int tile_rows;     // block-diagonal: rows per CTA
  int order = 0;     // causal item order: 0 = all heavy query tiles of the batch first (default: measured faster),
                     // 1 = (sequence, head)-major (K/V read from DRAM once; FSEEND_ATTN_ORDER=1)
};

// causal:     tmQ 4-D (768, S, T, B) box (64,1,128,1), tmKV same tensor with box (64,1,64,1);
//             out fp16 [B][T][S][256]
// block-diag: tmQ 4-D (768, 1, rows, 1) box (64,1,128,1), tmKV box (64,1,64,1); out fp16 [rows][256]
void launch_attn(const CUtensorMap& tmQ, const CUtensorMap& tmKV, __half* out, const AttnParams& p,
                 cudaStream_t stream);

}  // namespace fseend
