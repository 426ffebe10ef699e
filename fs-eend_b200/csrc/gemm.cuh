// Row-tile GEMM for the FS-EEND hot path:  OUT = epilogue( A[rows][K] * W[N][K]^T )  with fp16 operands
// (A and W both K-major, i.e. torch.nn.Linear's weight layout is consumed as is), fp32 accumulation in TMEM,
// fused epilogues (bias / ReLU / residual + LayerNorm / L2-norm / attractor-init broadcast).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fseend {

enum GemmEpilogue : int {
  EPI_BIAS = 0,     // out = acc + bias            (optionally ReLU);   N = n_tiles * 256
  EPI_LN = 1,       // out = LN(acc + bias [+ residual]) * g + b;      N = 256
  EPI_L2 = 2,       // out = (acc + bias) / ||acc + bias||_2;          N = 256
  EPI_CONVERT = 3,  // out[row, s, :] = acc + pe_proj[s, :], s < S;    N = 256  (attractor init)
  EPI_GLU = 4,      // out[:, c] = u[c] * sigmoid(u[128 + c]), u = acc + bias; each 256-wide tile -> 128 outputs
                    //   (weight rows pre-arranged per tile as [128 value rows | 128 gate rows])
  EPI_RESID = 5,    // y = residual + alpha * (acc + bias); out = ln_g ? LN(y) : y;  optional second output
                    //   out2 = LN2(out) (the next pre-norm sub-layer's input), N = 256
};
enum GemmAct : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SWISH = 2 };

struct GemmParams {
  int rows_per_seq;   // rows in one sequence (flat GEMM: all rows, n_seq = 1)
  int n_seq;
  int tiles_per_seq;  // ceil(rows_per_seq / 128)
  int n_tiles;        // N / 256
  int k_blocks;       // K / 64 (per tap)
  int taps;           // 1 (Linear) or kernel width (Conv1d as shifted GEMMs)
  int tap_shift;      // row offset of tap 0 (Conv1d: -padding)
  int a_row_offset;   // added to every A row coordinate (streaming conv: output row 0 = window centre)
  int mode;           // GemmEpilogue
  int relu;           // EPI_BIAS activation: GemmAct
  int has_residual;
  int S;              // EPI_CONVERT: attractor slots
  float ln_eps;
  float alpha;        // EPI_RESID: scale of the GEMM branch (Conformer half-step residual = 0.5)
  const float* bias;     // [N] or nullptr
  const float* ln_g;     // [256]
  const float* ln_b;     // [256]
  const float* ln2_g;    // EPI_RESID / EPI_LN: optional second LayerNorm -> tmO2
  const float* ln2_b;
  const float* pe_proj;  // [S][256]
  const int* seq_len;    // optional [n_seq]: EPI_LN rows t >= seq_len[b] are written as zeros
  const int* a_row_offset_dev = nullptr;   // optional: device int added to a_row_offset (graph-replayable streaming conv)
  const CUtensorMap* tmB_half = nullptr;   // optional: the weight with box (64,128) — enables the weight-stationary
                                           // CTA-pair kernel (gemm_pair.cu) for EPI_BIAS / EPI_LN, one tap, K <= 256
};

// tmA: 3-D (K, rows_per_seq, n_seq) box (64,128,1);  tmB: 2-D (K, taps*N) box (64,256);
// tmR: residual, same geometry as the output;  tmO: 3-D (N, rows_per_seq, n_seq) box (64,128,1)
//      (EPI_CONVERT: 3-D (256, S, rows) box (64,1,128)).
void launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const CUtensorMap& tmO,
                 const GemmParams& p, cudaStream_t stream);
// with a second output descriptor (same geometry as tmO) for the ln2 output
void launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const CUtensorMap& tmO,
                  const CUtensorMap& tmO2, const GemmParams& p, cudaStream_t stream);

}  // namespace fseend
