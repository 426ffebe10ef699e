// Fused speaker-axis attention: QKV projection + softmax(q k^T) v over the S slots of each frame (see spkfuse.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fseend {

struct SpkFuseParams {
  int rows;            // n_frames * S attractor rows
  int S;               // slots per frame (<= 16)
  int tile_rows;       // (128 / S) * S: whole frames per 128-row tile
  float scale;         // head_dim^-0.5
  const float* bias;   // in_proj_bias [768] = q | k | v
  __half* out;         // [rows][256] attention output (heads concatenated), input of the out-projection
};

// tmX: activations [rows][256] fp16 as the 3-D row map (256, rows, 1) box (64,128,1);
// tmW: in_proj_weight [768][256] fp16 as a 2-D map box (64 k, 64 rows).
void launch_spkfuse(const CUtensorMap& tmX, const CUtensorMap& tmW, const SpkFuseParams& p, cudaStream_t stream);

}  // namespace fseend
