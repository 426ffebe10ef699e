// FS-EEND model object behind the C ABI (include/fseend_b200.h): weight conversion, workspace/TMA-descriptor
// plan per (B, Tmax, S), and the launch sequence of the forward pass.
//
// Reference path restated here (file:line relative to /root/reference/FS-EEND/nnet):
//   model/onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm.py:67-84   test()
//   ...:162-188  encoder  | :38-41 conv + L2 | :112-118 decoder | modules/merge_tfm_encoder.py:356-376 fusion layer
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fseend_b200.h"
#include "attn.cuh"
#include "elementwise.cuh"
#include "embloss.cuh"
#include "ffn.cuh"
#include "gemm.cuh"
#include "loss.cuh"
#include "p32.cuh"
#include "spkfuse.cuh"
#include "tmap.h"

namespace fseend {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }

#define CUDA_CHECK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      throw std::runtime_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));        \
  } while (0)

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  void alloc(size_t n) {
    free();
    CUDA_CHECK(cudaMalloc(&p, n));
    bytes = n;
  }
  void free() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  ~DevBuf() { free(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// fp16 weight matrix [rows][K] on the device + its TMA descriptor (box 64 x 256 rows)
struct WMat {
  DevBuf buf;
  int rows = 0, K = 0;
  CUtensorMap tm;     // box (64 k, 256 rows): B operand of the 128x256 GEMM tile
  CUtensorMap tm128;  // box (64 k, 128 rows): 16 KB weight slots of the fused FFN
  CUtensorMap tm64;   // box (64 k, 64 rows): half-tile of the CTA-pair FFN (each CTA holds half of W1's chunk rows)
  void upload(const std::vector<__half>& h, int rows_, int K_) {
    rows = rows_;
    K = K_;
    buf.alloc(h.size() * sizeof(__half));
    CUDA_CHECK(cudaMemcpy(buf.p, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice));
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(rows)};
    uint64_t str[1] = {static_cast<uint64_t>(K)};
    uint32_t box[2] = {64, 256};
    tm = make_tmap_f16(buf.p, 2, dims, str, box);
    uint32_t box128[2] = {64, 128};
    tm128 = make_tmap_f16(buf.p, 2, dims, str, box128);
    uint32_t box64[2] = {64, 64};
    tm64 = make_tmap_f16(buf.p, 2, dims, str, box64);
  }
};

struct FVec {
  DevBuf buf;
  void upload(const std::vector<float>& h) {
    buf.alloc(h.size() * sizeof(float));
    CUDA_CHECK(cudaMemcpy(buf.p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  const float* f() const { return static_cast<const float*>(buf.p); }
};

struct EncLayer {
  WMat wqkv, wo, w1, w2;
  FVec bqkv, bo, b1, b2, g1, be1, g2, be2;
};
struct DecLayer {
  WMat wqkv1, wo1, wqkv2, wo2, w1, w2;
  FVec bqkv1, bo1, bqkv2, bo2, b1, b2, g11, be11, g21, be21, g22, be22;
};

struct ProfEntry {
  std::string name;
  cudaEvent_t a, b;
};

}  // namespace
}  // namespace fseend

using namespace fseend;

struct fseend_fs_model {
  fseend_fs_config cfg;
  int Kin = 0;  // in_size padded to a multiple of 64
  FVec bn_scale, bn_shift;
  WMat w_in;
  FVec b_in, g_in, be_in;
  std::vector<std::unique_ptr<EncLayer>> enc;
  WMat w_conv;
  FVec b_conv;
  WMat w_cvt;
  FVec pe_proj;  // [kMaxSlots][256]
  std::vector<std::unique_ptr<DecLayer>> dec;

  // ---- plan (workspace + descriptors) for the current (B, T, S)
  int pB = 0, pT = 0, pS = 0;
  size_t ws_bytes = 0;
  DevBuf x16, hA, hB, qkv_e, ao_e, f_e, emb16, aX, aY, aZ, qkv_d, ao_d, f_d, cu_dev, len_dev, x_stage, logits_stage;
  // pinned staging of (cu_seqlens [B+1] | lens [B]): kHostSlots rotating slots, each guarded by an event, so that
  // back-to-back asynchronous forwards never overwrite a slot whose copy is still pending
  static constexpr int kHostSlots = 8;
  int* cu_host = nullptr;
  cudaEvent_t cu_ev[kHostSlots] = {};
  int cu_slot = 0;
  // host-buffer forward (fseend_fs_forward_host): copy stream + compute stream + one event per chunk
  static constexpr int kMaxHostChunks = 8;
  int host_chunks = 0;     // 0 = automatic
  cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
  cudaEvent_t chunk_ev[kMaxHostChunks] = {};
  // pipelined host-buffer forward (fseend_fs_forward_host_async): kHostDepth calls in flight, each with its own device
  // staging of the features / logits; the H2D copy of call i+1 overlaps the kernels of call i
  static constexpr int kHostDepth = 2;
  DevBuf hx[kHostDepth], hl[kHostDepth];
  cudaEvent_t h2d_ev[kHostDepth] = {}, done_ev[kHostDepth] = {};
  long long host_calls = 0;            // tickets handed out so far

  // descriptors
  CUtensorMap tm_x16, tm_hA, tm_hB, tm_qkv_e_out, tm_qkv_e_attn, tm_ao_e_attn, tm_ao_e, tm_f_e_out, tm_f_e_in;
  CUtensorMap tm_hconv_in, tm_hB_seq, tm_emb_out, tm_emb_in, tm_cvt_out;
  CUtensorMap tm_aX, tm_aY, tm_aZ, tm_qkv_d_out, tm_qkv_d_attn, tm_ao_d_attn, tm_qkv_d_spk, tm_ao_d_spk, tm_ao_d, tm_qkv_e_kv, tm_qkv_d_kv, tm_qkv_d_spk_kv, tm_f_d_out, tm_f_d_in;

  int spk_mode = 2;  // 0: CUDA-core speaker attention, 1: tcgen05 block-diagonal attention (both after a QKV GEMM),
                     // 2: QKV projection + attention fused in one kernel (spkfuse.cu)
  int ffn_mode = 5;  // 0: two GEMM launches (hidden layer through HBM); fused: 1 = SS, 2 = SS + 2-CTA weight multicast,
                     // 3 = TS (hidden chunk stays in TMEM), 4 = TS + multicast, 5 = TS on a CTA pair (cta_group::2)
  bool profiling = false;
  std::vector<ProfEntry> prof_pending;
  std::map<std::string, std::pair<double, int>> prof_acc;
  int launches_last = 0;

  ~fseend_fs_model() {
    if (cu_host) cudaFreeHost(cu_host);
    for (auto& e : cu_ev)
      if (e) cudaEventDestroy(e);
    for (auto& e : chunk_ev)
      if (e) cudaEventDestroy(e);
    for (auto& e : h2d_ev)
      if (e) cudaEventDestroy(e);
    for (auto& e : done_ev)
      if (e) cudaEventDestroy(e);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (compute_stream) cudaStreamDestroy(compute_stream);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Frame-by-frame streaming state (reference: StreamingTransformerEDADiarization.test, FS:stream_model:31-60 and
// FS:stream_mod:82-269).  Device-resident: projected K/V caches per layer (the reference caches layer inputs and
// re-projects them every step — identical arithmetic, O(t·d²) less work), the encoder-output history that feeds
// the look-ahead conv, and one-tile workspaces.
struct fseend_fs_stream {
  fseend_fs_model* m = nullptr;
  int B = 0, S = 0;
  int cap = 0;        // frames of capacity in every cache
  int t = 0;          // frames pushed so far (real + flush)
  std::vector<std::unique_ptr<DevBuf>> enc_k, enc_v, dec_k, dec_v;
  DevBuf hist, x16, h0, h1, qkv, ao, a0, a1, a2, emb, cu;
  CUtensorMap tm_x16, tm_h0, tm_h1, tm_qkv_e, tm_ao_e, tm_hist, tm_emb_out, tm_emb_in, tm_cvt_out, tm_a0, tm_a1, tm_a2,
      tm_qkv_d, tm_ao_d, tm_qkv_spk, tm_qkv_spk_kv, tm_ao_spk;
  // CUDA-graph replay of the steady-state step: every per-frame quantity a kernel needs (encoder frame index,
  // decoder frame index) lives in `ctr` on the device and is advanced by the last node of the step, the frame is
  // staged through x_in and the logits through y_out, so one instantiated graph serves every frame.
  DevBuf ctr, x_in, y_out;                 // int[4]: {encoder frame index, decoder frame index, -, -}
  cudaGraphExec_t graph_step = nullptr;    // frame in, logits out (steady state)
  cudaGraphExec_t graph_flush = nullptr;   // flush step (zero conv input), logits out
  int graph_cap = 0;                       // cache capacity the graphs were captured for
  int eager_decodes = 0;                   // decoder steps run eagerly so far (first ones: lazy kernel attributes)
  bool use_graph = true;
  cudaStream_t cap_stream = nullptr;       // private capture stream
  // Small-row path (B * S <= 16 rows per step, the usual B = 1 recording): every product is a weight-streaming
  // matrix-vector job, so the step runs on the CUDA-core row-vector kernel of p32.cu with the model's fp16 weights and
  // fp32 activations / caches instead of 128-row tcgen05 tiles (FSEEND_STREAM_RV=0 keeps the tile kernels).
  bool rv = false;
  std::vector<std::unique_ptr<DevBuf>> enc_k32, enc_v32, dec_k32, dec_v32;
  DevBuf hist32, rX, rh0, rh1, rqkv, rao, rf, ra0, ra1, ra2, rT1, remb, rY, epi_ctr;
  ~fseend_fs_stream() {
    if (graph_step) cudaGraphExecDestroy(graph_step);
    if (graph_flush) cudaGraphExecDestroy(graph_flush);
    if (cap_stream) cudaStreamDestroy(cap_stream);
  }
};

namespace fseend {
namespace {

constexpr int kMaxSlots = 16;

struct TensorTable {
  std::map<std::string, std::pair<const float*, long long>> t;
  const float* get(const std::string& name, long long numel) const {
    auto it = t.find(name);
    if (it == t.end()) throw std::invalid_argument("missing state_dict tensor: " + name);
    if (it->second.second != numel)
      throw std::invalid_argument("state_dict tensor " + name + " has " + std::to_string(it->second.second) +
                                  " elements, expected " + std::to_string(numel));
    return it->second.first;
  }
  // positional-encoding buffer may be longer than needed
  const float* get_atleast(const std::string& name, long long numel) const {
    auto it = t.find(name);
    if (it == t.end()) throw std::invalid_argument("missing state_dict tensor: " + name);
    if (it->second.second < numel) throw std::invalid_argument("state_dict tensor " + name + " too small");
    return it->second.first;
  }
};

std::vector<__half> to_half(const float* w, size_t n) {
  std::vector<__half> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = __float2half_rn(w[i]);
  return h;
}
std::vector<float> to_vec(const float* w, size_t n) { return std::vector<float>(w, w + n); }

void load_linear(const TensorTable& tt, const std::string& name, int out_f, int in_f, WMat& w, FVec& b) {
  w.upload(to_half(tt.get(name + ".weight", 1ll * out_f * in_f), 1ull * out_f * in_f), out_f, in_f);
  b.upload(to_vec(tt.get(name + ".bias", out_f), out_f));
}
void load_ln(const TensorTable& tt, const std::string& name, int n, FVec& g, FVec& b) {
  g.upload(to_vec(tt.get(name + ".weight", n), n));
  b.upload(to_vec(tt.get(name + ".bias", n), n));
}
void load_attn(const TensorTable& tt, const std::string& name, int D, WMat& wqkv, FVec& bqkv, WMat& wo, FVec& bo) {
  wqkv.upload(to_half(tt.get(name + ".in_proj_weight", 3ll * D * D), 3ull * D * D), 3 * D, D);
  bqkv.upload(to_vec(tt.get(name + ".in_proj_bias", 3 * D), 3 * D));
  load_linear(tt, name + ".out_proj", D, D, wo, bo);
}

void build_model(fseend_fs_model* m, const TensorTable& tt) {
  const fseend_fs_config& c = m->cfg;
  const int D = c.n_units, Din = c.in_size;
  m->Kin = (Din + 63) / 64 * 64;
  // BatchNorm1d (eval) -> per-channel affine applied in fp32 before the fp16 cast (FS:model:166)
  {
    const float* w = tt.get("enc.bn.weight", Din);
    const float* b = tt.get("enc.bn.bias", Din);
    const float* mu = tt.get("enc.bn.running_mean", Din);
    const float* var = tt.get("enc.bn.running_var", Din);
    std::vector<float> sc(Din), sh(Din);
    for (int i = 0; i < Din; ++i) {
      sc[i] = w[i] / sqrtf(var[i] + c.bn_eps);
      sh[i] = b[i] - mu[i] * sc[i];
    }
    m->bn_scale.upload(sc);
    m->bn_shift.upload(sh);
  }
  {  // input projection, K zero-padded to Kin (FS:model:173-174)
    const float* w = tt.get("enc.encoder.weight", 1ll * D * Din);
    std::vector<__half> h(1ull * D * m->Kin, __float2half_rn(0.f));
    for (int o = 0; o < D; ++o)
      for (int i = 0; i < Din; ++i) h[1ull * o * m->Kin + i] = __float2half_rn(w[1ull * o * Din + i]);
    m->w_in.upload(h, D, m->Kin);
    m->b_in.upload(to_vec(tt.get("enc.encoder.bias", D), D));
    load_ln(tt, "enc.encoder_norm", D, m->g_in, m->be_in);
  }
  for (int l = 0; l < c.enc_n_layers; ++l) {
    auto L = std::make_unique<EncLayer>();
    const std::string p = "enc.transformer_encoder.layers." + std::to_string(l);
    load_attn(tt, p + ".self_attn", D, L->wqkv, L->bqkv, L->wo, L->bo);
    load_linear(tt, p + ".linear1", c.enc_dim_feedforward, D, L->w1, L->b1);
    load_linear(tt, p + ".linear2", D, c.enc_dim_feedforward, L->w2, L->b2);
    load_ln(tt, p + ".norm1", D, L->g1, L->be1);
    load_ln(tt, p + ".norm2", D, L->g2, L->be2);
    m->enc.push_back(std::move(L));
  }
  {  // Conv1d weight (Dout, Din, K) -> K tap matrices [tap][Dout][Din] (FS:model:30,40)
    const int Kc = c.conv_kernel;
    const float* w = tt.get("cnn.weight", 1ll * D * D * Kc);
    std::vector<__half> h(1ull * Kc * D * D);
    for (int k = 0; k < Kc; ++k)
      for (int o = 0; o < D; ++o)
        for (int i = 0; i < D; ++i)
          h[(1ull * k * D + o) * D + i] = __float2half_rn(w[(1ull * o * D + i) * Kc + k]);
    m->w_conv.upload(h, Kc * D, D);
    m->b_conv.upload(to_vec(tt.get("cnn.bias", D), D));
  }
  {  // attractor init: convert(cat[emb, pe_s]) = Wc[:, :D] emb + (Wc[:, D:] pe_s + bc)  (FS:model:113-114)
    const float* w = tt.get("dec.convert.weight", 2ll * D * D);
    const float* b = tt.get("dec.convert.bias", D);
    const float* pe = tt.get_atleast("dec.pos_enc.pe", 1ll * kMaxSlots * D);
    std::vector<__half> h(1ull * D * D);
    for (int o = 0; o < D; ++o)
      for (int i = 0; i < D; ++i) h[1ull * o * D + i] = __float2half_rn(w[1ull * o * 2 * D + i]);
    m->w_cvt.upload(h, D, D);
    std::vector<float> pp(1ull * kMaxSlots * D);
    for (int s = 0; s < kMaxSlots; ++s)
      for (int o = 0; o < D; ++o) {
        double acc = b[o];
        for (int i = 0; i < D; ++i) acc += static_cast<double>(w[1ull * o * 2 * D + D + i]) * pe[1ull * s * D + i];
        pp[1ull * s * D + o] = static_cast<float>(acc);
      }
    m->pe_proj.upload(pp);
  }
  for (int l = 0; l < c.dec_n_layers; ++l) {
    auto L = std::make_unique<DecLayer>();
    const std::string p = "dec.attractor_decoder.layers." + std::to_string(l);
    load_attn(tt, p + ".self_attn1", D, L->wqkv1, L->bqkv1, L->wo1, L->bo1);
    load_attn(tt, p + ".self_attn2", D, L->wqkv2, L->bqkv2, L->wo2, L->bo2);
    load_linear(tt, p + ".linear1", c.dec_dim_feedforward, D, L->w1, L->b1);
    load_linear(tt, p + ".linear2", D, c.dec_dim_feedforward, L->w2, L->b2);
    load_ln(tt, p + ".norm11", D, L->g11, L->be11);
    load_ln(tt, p + ".norm21", D, L->g21, L->be21);
    load_ln(tt, p + ".norm22", D, L->g22, L->be22);
    m->dec.push_back(std::move(L));
  }
}

CUtensorMap rows_map(const DevBuf& buf, uint64_t cols, uint64_t rows_per_seq, uint64_t n_seq) {
  return make_tmap_rows3d(buf.p, cols, cols, rows_per_seq, n_seq, 128);
}
CUtensorMap attn_map(const DevBuf& buf, uint64_t cols, int S, int T, int B, uint32_t box_rows = 128) {
  uint64_t dims[4] = {cols, static_cast<uint64_t>(S), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
  uint64_t str[3] = {cols, cols * S, cols * S * T};
  uint32_t box[4] = {64, 1, box_rows, 1};
  return make_tmap_f16(buf.p, 4, dims, str, box);
}

void make_plan(fseend_fs_model* m, int B, int T, int S) {
  if (m->pB == B && m->pT == T && m->pS == S) return;
  m->pB = m->pT = m->pS = 0;   // the key is only valid once every buffer and descriptor below exists (an allocation may throw)
  const fseend_fs_config& c = m->cfg;
  const int D = c.n_units;
  const size_t Me = 1ull * B * T, Md = Me * S;
  size_t total = 0;
  auto A = [&](DevBuf& b, size_t bytes) {
    b.alloc(bytes);
    total += bytes;
  };
  A(m->x16, Me * m->Kin * 2);
  A(m->hA, Me * D * 2);
  A(m->hB, Me * D * 2);
  A(m->qkv_e, Me * 3 * D * 2);
  A(m->ao_e, Me * D * 2);
  A(m->f_e, Me * c.enc_dim_feedforward * 2);
  A(m->emb16, Me * D * 2);
  A(m->aX, Md * D * 2);
  A(m->aY, Md * D * 2);
  A(m->aZ, Md * D * 2);
  A(m->qkv_d, Md * 3 * D * 2);
  A(m->ao_d, Md * D * 2);
  A(m->f_d, Md * c.dec_dim_feedforward * 2);
  A(m->cu_dev, (B + 1) * sizeof(int));
  A(m->len_dev, B * sizeof(int));
  m->x_stage.free();
  m->logits_stage.free();
  if (m->cu_host) {
    CUDA_CHECK(cudaDeviceSynchronize());   // pending copies out of the old staging buffer
    cudaFreeHost(m->cu_host);
    m->cu_host = nullptr;
  }
  CUDA_CHECK(cudaMallocHost(&m->cu_host, fseend_fs_model::kHostSlots * (2 * B + 1) * sizeof(int)));
  m->ws_bytes = total;

  m->tm_x16 = rows_map(m->x16, m->Kin, Me, 1);
  m->tm_hA = rows_map(m->hA, D, Me, 1);
  m->tm_hB = rows_map(m->hB, D, Me, 1);
  m->tm_qkv_e_out = rows_map(m->qkv_e, 3 * D, Me, 1);
  m->tm_qkv_e_attn = attn_map(m->qkv_e, 3 * D, 1, T, B);
  m->tm_qkv_e_kv = attn_map(m->qkv_e, 3 * D, 1, T, B, 64);
  m->tm_ao_e_attn = attn_map(m->ao_e, D, 1, T, B);
  m->tm_ao_e = rows_map(m->ao_e, D, Me, 1);
  m->tm_f_e_out = rows_map(m->f_e, c.enc_dim_feedforward, Me, 1);
  m->tm_f_e_in = m->tm_f_e_out;
  // conv: per-sequence tiles so that shifted taps zero-fill across sequence ends
  m->tm_hconv_in = rows_map(m->hA, D, T, B);   // encoder output always ends in hA (see forward)
  m->tm_hB_seq = rows_map(m->hB, D, T, B);
  m->tm_emb_out = rows_map(m->emb16, D, T, B);
  m->tm_emb_in = rows_map(m->emb16, D, Me, 1);
  {
    uint64_t dims[3] = {static_cast<uint64_t>(D), static_cast<uint64_t>(S), Me};
    uint64_t str[2] = {static_cast<uint64_t>(D), static_cast<uint64_t>(D) * S};
    uint32_t box[3] = {64, 1, 128};
    m->tm_cvt_out = make_tmap_f16(m->aX.p, 3, dims, str, box);
  }
  m->tm_aX = rows_map(m->aX, D, Md, 1);
  m->tm_aY = rows_map(m->aY, D, Md, 1);
  m->tm_aZ = rows_map(m->aZ, D, Md, 1);
  m->tm_qkv_d_out = rows_map(m->qkv_d, 3 * D, Md, 1);
  m->tm_qkv_d_attn = attn_map(m->qkv_d, 3 * D, S, T, B);
  m->tm_qkv_d_kv = attn_map(m->qkv_d, 3 * D, S, T, B, 64);
  m->tm_ao_d_attn = attn_map(m->ao_d, D, S, T, B);
  {
    uint64_t dq[4] = {static_cast<uint64_t>(3 * D), 1, Md, 1}, sq[3] = {3ull * D, 3ull * D, 3ull * D * Md};
    uint64_t dout[4] = {static_cast<uint64_t>(D), 1, Md, 1}, so[3] = {1ull * D, 1ull * D, 1ull * D * Md};
    uint32_t bq[4] = {64, 1, 128, 1}, bo[4] = {64, 1, static_cast<uint32_t>((128 / S) * S), 1};
    m->tm_qkv_d_spk = make_tmap_f16(m->qkv_d.p, 4, dq, sq, bq);
    uint32_t bkv[4] = {64, 1, 64, 1};
    m->tm_qkv_d_spk_kv = make_tmap_f16(m->qkv_d.p, 4, dq, sq, bkv);
    m->tm_ao_d_spk = make_tmap_f16(m->ao_d.p, 4, dout, so, bo);
  }
  m->tm_ao_d = rows_map(m->ao_d, D, Md, 1);
  m->tm_f_d_out = rows_map(m->f_d, c.dec_dim_feedforward, Md, 1);
  m->tm_f_d_in = m->tm_f_d_out;
  m->pB = B;
  m->pT = T;
  m->pS = S;
}

struct Launcher {
  fseend_fs_model* m;
  cudaStream_t st;
  int count = 0;
  template <class F>
  void run(const char* name, F&& f) {
    ProfEntry e;
    if (m->profiling) {
      e.name = name;
      CUDA_CHECK(cudaEventCreate(&e.a));
      CUDA_CHECK(cudaEventCreate(&e.b));
      CUDA_CHECK(cudaEventRecord(e.a, st));
    }
    f();
    ++count;
    if (m->profiling) {
      CUDA_CHECK(cudaEventRecord(e.b, st));
      m->prof_pending.push_back(e);
    }
  }
};

GemmParams flat_params(size_t rows, int N, int K, int mode) {
  GemmParams p{};
  p.rows_per_seq = static_cast<int>(rows);
  p.n_seq = 1;
  p.tiles_per_seq = static_cast<int>((rows + 127) / 128);
  p.n_tiles = N / 256;
  p.k_blocks = K / 64;
  p.taps = 1;
  p.tap_shift = 0;
  p.mode = mode;
  p.ln_eps = 1e-5f;
  return p;
}

FfnParams ffn_params(int rows_per_seq, int n_seq, int F, const FVec& b1, const FVec& b2, const FVec& g, const FVec& b,
                     float eps, const int* seq_len) {
  FfnParams p{};
  p.rows_per_seq = rows_per_seq;
  p.n_seq = n_seq;
  p.tiles_per_seq = (rows_per_seq + 127) / 128;
  p.F = F;
  p.ln_eps = eps;
  p.b1 = b1.f();
  p.b2 = b2.f();
  p.ln_g = g.f();
  p.ln_b = b.f();
  p.seq_len = seq_len;
  return p;
}

// T_force (0 = longest sequence of this call): pad to a common T, so that chunks of one host batch share a plan and
// write into one [B][T][S] output.
void forward_impl(fseend_fs_model* m, const float* x_packed, const int* ilens, int B, int S, float* logits,
                  float* emb_out, float* att_out, cudaStream_t st, int T_force = 0) {
  const fseend_fs_config& c = m->cfg;
  const int D = c.n_units;
  if (B < 1) throw std::invalid_argument("B must be >= 1");
  if (S < 1 || S > kMaxSlots) throw std::invalid_argument("max_nspks must be in [1,16]");
  int T = 0;
  long long total = 0;
  for (int b = 0; b < B; ++b) {
    if (ilens[b] < 1) throw std::invalid_argument("ilens must be >= 1");
    T = ilens[b] > T ? ilens[b] : T;
    total += ilens[b];
  }
  if (T_force) {
    if (T_force < T) throw std::invalid_argument("T_force below the longest sequence");
    T = T_force;
  }
  if (1ll * B * T * S * 768 >= (1ll << 31) * 8) throw std::invalid_argument("batch too large for one call");
  make_plan(m, B, T, S);
  const size_t Me = 1ull * B * T, Md = Me * S;

  // cu_seqlens + lens: pinned staging -> device (async on the same stream)
  const int slot = m->cu_slot;
  m->cu_slot = (slot + 1) % fseend_fs_model::kHostSlots;
  if (!m->cu_ev[slot]) CUDA_CHECK(cudaEventCreateWithFlags(&m->cu_ev[slot], cudaEventDisableTiming));
  else CUDA_CHECK(cudaEventSynchronize(m->cu_ev[slot]));   // the copies that last used this slot have completed
  int* cu = m->cu_host + slot * (2 * B + 1);
  cu[0] = 0;
  for (int b = 0; b < B; ++b) {
    cu[b + 1] = cu[b] + ilens[b];
    cu[B + 1 + b] = ilens[b];
  }
  CUDA_CHECK(cudaMemcpyAsync(m->cu_dev.p, cu, (B + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(m->len_dev.p, cu + B + 1, B * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaEventRecord(m->cu_ev[slot], st));

  Launcher L{m, st};
  CUtensorMap none = m->tm_hA;  // placeholder for unused descriptor arguments

  L.run("prep_input", [&] {
    launch_prep_input(x_packed, static_cast<const int*>(m->cu_dev.p), B, T, c.in_size, m->Kin, m->bn_scale.f(),
                      m->bn_shift.f(), static_cast<__half*>(m->x16.p), st);
  });
  // input projection + LN -> hA
  {
    GemmParams p = flat_params(Me, D, m->Kin, EPI_LN);
    p.bias = m->b_in.f();
    p.ln_g = m->g_in.f();
    p.ln_b = m->be_in.f();
    p.ln_eps = c.ln_eps;
    L.run("enc.gemm_in_ln", [&] { p.tmB_half = &m->w_in.tm128; launch_gemm(m->tm_x16, m->w_in.tm, none, m->tm_hA, p, st); });
  }
  // encoder layers: hA -> (attn) -> hB -> (ffn) -> hA
  for (int l = 0; l < c.enc_n_layers; ++l) {
    EncLayer& E = *m->enc[l];
    {
      GemmParams p = flat_params(Me, 3 * D, D, EPI_BIAS);
      p.bias = E.bqkv.f();
      L.run("enc.gemm_qkv", [&] { p.tmB_half = &E.wqkv.tm128; launch_gemm(m->tm_hA, E.wqkv.tm, none, m->tm_qkv_e_out, p, st); });
    }
    {
      AttnParams a{B, 1, T, c.n_heads, c.has_mask ? c.mask_delay : (1 << 28), 1.f / sqrtf(64.f), ATTN_CAUSAL, 128};
      L.run("enc.attn_causal", [&] { launch_attn(m->tm_qkv_e_attn, m->tm_qkv_e_kv, static_cast<__half*>(m->ao_e.p), a, st); });
    }
    {
      GemmParams p = flat_params(Me, D, D, EPI_LN);
      p.bias = E.bo.f();
      p.has_residual = 1;
      p.ln_g = E.g1.f();
      p.ln_b = E.be1.f();
      p.ln_eps = c.ln_eps;
      L.run("enc.gemm_out_ln", [&] { p.tmB_half = &E.wo.tm128; launch_gemm(m->tm_ao_e, E.wo.tm, m->tm_hA, m->tm_hB, p, st); });
    }
    const bool last = (l == c.enc_n_layers - 1);
    if (m->ffn_mode > 0) {
      // last layer: rows t >= ilens[b] become zeros = the reference's truncate + re-pad(0) (FS:model:38-39)
      FfnParams fp = ffn_params(last ? T : static_cast<int>(Me), last ? B : 1, c.enc_dim_feedforward, E.b1, E.b2, E.g2,
                                E.be2, c.ln_eps, last ? static_cast<const int*>(m->len_dev.p) : nullptr);
      const CUtensorMap& tx = last ? m->tm_hB_seq : m->tm_hB;
      const CUtensorMap& to = last ? m->tm_hconv_in : m->tm_hA;
      L.run("enc.ffn_fused", [&] {
        if (m->ffn_mode == 5) launch_ffn_pair(tx, E.w1.tm64, E.w2.tm128, to, fp, st);
        else launch_ffn(tx, E.w1.tm128, E.w2.tm128, to, fp, m->ffn_mode, st);
      });
    } else {
      {
        GemmParams p = flat_params(Me, c.enc_dim_feedforward, D, EPI_BIAS);
        p.bias = E.b1.f();
        p.relu = 1;
        L.run("enc.gemm_ffn1", [&] { p.tmB_half = &E.w1.tm128; launch_gemm(m->tm_hB, E.w1.tm, none, m->tm_f_e_out, p, st); });
      }
      {
        GemmParams p = flat_params(Me, D, c.enc_dim_feedforward, EPI_LN);
        p.bias = E.b2.f();
        p.has_residual = 1;
        p.ln_g = E.g2.f();
        p.ln_b = E.be2.f();
        p.ln_eps = c.ln_eps;
        // last layer: rows t >= ilens[b] become zeros = the reference's truncate + re-pad(0) (FS:model:38-39)
        if (l == c.enc_n_layers - 1) {
          p.rows_per_seq = T;
          p.n_seq = B;
          p.tiles_per_seq = (T + 127) / 128;
          p.seq_len = static_cast<const int*>(m->len_dev.p);
          CUtensorMap tmA = make_tmap_rows3d(m->f_e.p, c.enc_dim_feedforward, c.enc_dim_feedforward, T, B, 128);
          CUtensorMap tmR = make_tmap_rows3d(m->hB.p, D, D, T, B, 128);
          L.run("enc.gemm_ffn2_ln", [&] { p.tmB_half = &E.w2.tm128; launch_gemm(tmA, E.w2.tm, tmR, m->tm_hconv_in, p, st); });
        } else {
          L.run("enc.gemm_ffn2_ln", [&] { p.tmB_half = &E.w2.tm128; launch_gemm(m->tm_f_e_in, E.w2.tm, m->tm_hB, m->tm_hA, p, st); });
        }
      }
  
    }
  }
  if (c.enc_n_layers == 0) throw std::invalid_argument("enc_n_layers must be >= 1");
  // look-ahead Conv1d as 19 shifted GEMMs + bias + L2 norm -> emb16
  {
    GemmParams p{};
    p.rows_per_seq = T;
    p.n_seq = B;
    p.tiles_per_seq = (T + 127) / 128;
    p.n_tiles = 1;
    p.k_blocks = D / 64;
    p.taps = c.conv_kernel;
    p.tap_shift = -c.conv_padding;
    p.mode = EPI_L2;
    p.bias = m->b_conv.f();
    L.run("gemm_conv_l2", [&] { p.tmB_half = &m->w_conv.tm128; launch_gemm(m->tm_hconv_in, m->w_conv.tm, none, m->tm_emb_out, p, st); });
  }
  // attractor init -> aX [B][T][S][D]
  {
    GemmParams p = flat_params(Me, D, D, EPI_CONVERT);
    p.S = S;
    p.pe_proj = m->pe_proj.f();
    L.run("gemm_convert", [&] { p.tmB_half = &m->w_cvt.tm128; launch_gemm(m->tm_emb_in, m->w_cvt.tm, none, m->tm_cvt_out, p, st); });
  }
  // decoder layers: aX -> time attention -> aY -> speaker attention -> aZ -> FFN -> aX
  for (int l = 0; l < c.dec_n_layers; ++l) {
    DecLayer& Dl = *m->dec[l];
    {
      GemmParams p = flat_params(Md, 3 * D, D, EPI_BIAS);
      p.bias = Dl.bqkv1.f();
      L.run("dec.gemm_qkv1", [&] { p.tmB_half = &Dl.wqkv1.tm128; launch_gemm(m->tm_aX, Dl.wqkv1.tm, none, m->tm_qkv_d_out, p, st); });
    }
    {
      AttnParams a{B, S, T, c.n_heads, c.mask_delay, 1.f / sqrtf(64.f), ATTN_CAUSAL, 128};
      L.run("dec.attn_causal", [&] { launch_attn(m->tm_qkv_d_attn, m->tm_qkv_d_kv, static_cast<__half*>(m->ao_d.p), a, st); });
    }
    {
      GemmParams p = flat_params(Md, D, D, EPI_LN);
      p.bias = Dl.bo1.f();
      p.has_residual = 1;
      p.ln_g = Dl.g11.f();
      p.ln_b = Dl.be11.f();
      p.ln_eps = c.ln_eps;
      L.run("dec.gemm_out1_ln", [&] { p.tmB_half = &Dl.wo1.tm128; launch_gemm(m->tm_ao_d, Dl.wo1.tm, m->tm_aX, m->tm_aY, p, st); });
    }
    if (m->spk_mode == 2) {
      SpkFuseParams sp{static_cast<int>(Md), S, (128 / S) * S, 1.f / sqrtf(64.f), Dl.bqkv2.f(),
                       static_cast<__half*>(m->ao_d.p)};
      L.run("dec.spk_fused", [&] { launch_spkfuse(m->tm_aY, Dl.wqkv2.tm64, sp, st); });
    } else {
      {
        GemmParams p = flat_params(Md, 3 * D, D, EPI_BIAS);
        p.bias = Dl.bqkv2.f();
        L.run("dec.gemm_qkv2", [&] { p.tmB_half = &Dl.wqkv2.tm128; launch_gemm(m->tm_aY, Dl.wqkv2.tm, none, m->tm_qkv_d_out, p, st); });
      }
      if (m->spk_mode == 1) {
        AttnParams a{1, S, static_cast<int>(Md), c.n_heads, 0, 1.f / sqrtf(64.f), ATTN_BLOCKDIAG, (128 / S) * S};
        L.run("dec.spk_attn", [&] { launch_attn(m->tm_qkv_d_spk, m->tm_qkv_d_spk_kv, static_cast<__half*>(m->ao_d.p), a, st); });
      } else {
        L.run("dec.spk_attn", [&] {
          launch_spk_attn(static_cast<const __half*>(m->qkv_d.p), static_cast<__half*>(m->ao_d.p),
                          static_cast<int>(Me), S, 1.f / sqrtf(64.f), st);
        });
      }
    }
    {
      GemmParams p = flat_params(Md, D, D, EPI_LN);
      p.bias = Dl.bo2.f();
      p.has_residual = 1;
      p.ln_g = Dl.g21.f();
      p.ln_b = Dl.be21.f();
      p.ln_eps = c.ln_eps;
      L.run("dec.gemm_out2_ln", [&] { p.tmB_half = &Dl.wo2.tm128; launch_gemm(m->tm_ao_d, Dl.wo2.tm, m->tm_aY, m->tm_aZ, p, st); });
    }
    if (m->ffn_mode > 0) {
      FfnParams fp = ffn_params(static_cast<int>(Md), 1, c.dec_dim_feedforward, Dl.b1, Dl.b2, Dl.g22, Dl.be22, c.ln_eps,
                                nullptr);
      L.run("dec.ffn_fused", [&] {
        if (m->ffn_mode == 5) launch_ffn_pair(m->tm_aZ, Dl.w1.tm64, Dl.w2.tm128, m->tm_aX, fp, st);
        else launch_ffn(m->tm_aZ, Dl.w1.tm128, Dl.w2.tm128, m->tm_aX, fp, m->ffn_mode, st);
      });
    } else {
      {
        GemmParams p = flat_params(Md, c.dec_dim_feedforward, D, EPI_BIAS);
        p.bias = Dl.b1.f();
        p.relu = 1;
        L.run("dec.gemm_ffn1", [&] { p.tmB_half = &Dl.w1.tm128; launch_gemm(m->tm_aZ, Dl.w1.tm, none, m->tm_f_d_out, p, st); });
      }
      {
        GemmParams p = flat_params(Md, D, c.dec_dim_feedforward, EPI_LN);
        p.bias = Dl.b2.f();
        p.has_residual = 1;
        p.ln_g = Dl.g22.f();
        p.ln_b = Dl.be22.f();
        p.ln_eps = c.ln_eps;
        L.run("dec.gemm_ffn2_ln", [&] { p.tmB_half = &Dl.w2.tm128; launch_gemm(m->tm_f_d_in, Dl.w2.tm, m->tm_aZ, m->tm_aX, p, st); });
      }
  
    }
  }
  L.run("head", [&] {
    launch_head(static_cast<const __half*>(m->emb16.p), static_cast<const __half*>(m->aX.p), static_cast<int>(Me), S,
                logits, emb_out, att_out, st);
  });
  m->launches_last = L.count;
  CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ streaming
void stream_alloc(fseend_fs_stream* s, int cap) {
  const fseend_fs_config& c = s->m->cfg;
  const int D = c.n_units, B = s->B, S = s->S;
  auto grow = [&](DevBuf& buf, size_t n_seq, size_t es = 2) {
    DevBuf nb;
    nb.alloc(n_seq * cap * D * es);
    CUDA_CHECK(cudaMemset(nb.p, 0, nb.bytes));
    if (buf.p && s->cap > 0)
      CUDA_CHECK(cudaMemcpy2D(nb.p, 1ull * cap * D * es, buf.p, 1ull * s->cap * D * es, 1ull * s->cap * D * es, n_seq,
                              cudaMemcpyDeviceToDevice));
    std::swap(buf.p, nb.p);
    std::swap(buf.bytes, nb.bytes);
  };
  CUDA_CHECK(cudaDeviceSynchronize());
  if (s->rv) {
    for (auto& b : s->enc_k32) grow(*b, B, 4);
    for (auto& b : s->enc_v32) grow(*b, B, 4);
    for (auto& b : s->dec_k32) grow(*b, 1ull * B * S, 4);
    for (auto& b : s->dec_v32) grow(*b, 1ull * B * S, 4);
    grow(s->hist32, B, 4);
  } else {
    for (auto& b : s->enc_k) grow(*b, B);
    for (auto& b : s->enc_v) grow(*b, B);
    for (auto& b : s->dec_k) grow(*b, 1ull * B * S);
    for (auto& b : s->dec_v) grow(*b, 1ull * B * S);
    grow(s->hist, B);
  }
  CUDA_CHECK(cudaDeviceSynchronize());   // memset / copies ran on the legacy stream: order them before ANY caller stream
  s->cap = cap;
  if (!s->rv) s->tm_hist = make_tmap_rows3d(s->hist.p, D, D, cap, B, 128);
}

void stream_init(fseend_fs_stream* s, fseend_fs_model* m, int B, int S) {
  const fseend_fs_config& c = m->cfg;
  const int D = c.n_units;
  s->m = m;
  s->B = B;
  s->S = S;
  for (int l = 0; l < c.enc_n_layers; ++l) {
    s->enc_k.push_back(std::make_unique<DevBuf>());
    s->enc_v.push_back(std::make_unique<DevBuf>());
    s->enc_k32.push_back(std::make_unique<DevBuf>());
    s->enc_v32.push_back(std::make_unique<DevBuf>());
  }
  for (int l = 0; l < c.dec_n_layers; ++l) {
    s->dec_k.push_back(std::make_unique<DevBuf>());
    s->dec_v.push_back(std::make_unique<DevBuf>());
    s->dec_k32.push_back(std::make_unique<DevBuf>());
    s->dec_v32.push_back(std::make_unique<DevBuf>());
  }
  const size_t Rd = 1ull * B * S;
  s->rv = B * S <= 16;
  if (const char* e = getenv("FSEEND_STREAM_RV")) s->rv = s->rv && e[0] != '0';
  if (s->rv) {
    const size_t f = sizeof(float);
    const size_t Fmax = static_cast<size_t>(std::max(c.enc_dim_feedforward, c.dec_dim_feedforward));
    s->rX.alloc(1ull * B * m->Kin * f);
    s->rh0.alloc(1ull * B * D * f);
    s->rh1.alloc(1ull * B * D * f);
    s->rqkv.alloc(Rd * 3 * D * f);
    s->rao.alloc(Rd * D * f);
    s->rf.alloc(Rd * Fmax * f);
    s->ra0.alloc(Rd * D * f);
    s->ra1.alloc(Rd * D * f);
    s->ra2.alloc(Rd * D * f);
    s->rT1.alloc(Rd * D * f);
    s->remb.alloc(1ull * B * D * f);
    s->rY.alloc(1ull * B * D * f);
    s->epi_ctr.alloc(sizeof(unsigned int));
    CUDA_CHECK(cudaMemset(s->epi_ctr.p, 0, sizeof(unsigned int)));
  }
  s->x16.alloc(1ull * B * m->Kin * 2);
  s->h0.alloc(1ull * B * D * 2);
  s->h1.alloc(1ull * B * D * 2);
  s->qkv.alloc(Rd * 3 * D * 2);
  s->ao.alloc(Rd * D * 2);
  s->a0.alloc(Rd * D * 2);
  s->a1.alloc(Rd * D * 2);
  s->a2.alloc(Rd * D * 2);
  s->emb.alloc(1ull * B * D * 2);
  {
    std::vector<int> cu(B + 1);
    for (int i = 0; i <= B; ++i) cu[i] = i;
    s->cu.alloc((B + 1) * sizeof(int));
    CUDA_CHECK(cudaMemcpy(s->cu.p, cu.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice));
  }
  s->ctr.alloc(4 * sizeof(int));
  CUDA_CHECK(cudaMemset(s->ctr.p, 0, 4 * sizeof(int)));
  s->x_in.alloc(1ull * B * c.in_size * sizeof(float));
  s->y_out.alloc(1ull * B * S * sizeof(float));
  if (const char* e = getenv("FSEEND_STREAM_GRAPH")) s->use_graph = e[0] != '0';
  s->tm_x16 = rows_map(s->x16, m->Kin, B, 1);
  s->tm_h0 = rows_map(s->h0, D, B, 1);
  s->tm_h1 = rows_map(s->h1, D, B, 1);
  s->tm_qkv_e = rows_map(s->qkv, 3 * D, B, 1);
  s->tm_ao_e = rows_map(s->ao, D, B, 1);
  s->tm_emb_out = rows_map(s->emb, D, 1, B);     // one output row per sequence
  s->tm_emb_in = rows_map(s->emb, D, B, 1);
  {
    uint64_t dims[3] = {static_cast<uint64_t>(D), static_cast<uint64_t>(S), static_cast<uint64_t>(B)};
    uint64_t str[2] = {static_cast<uint64_t>(D), static_cast<uint64_t>(D) * S};
    uint32_t box[3] = {64, 1, 128};
    s->tm_cvt_out = make_tmap_f16(s->a0.p, 3, dims, str, box);
  }
  s->tm_a0 = rows_map(s->a0, D, Rd, 1);
  s->tm_a1 = rows_map(s->a1, D, Rd, 1);
  s->tm_a2 = rows_map(s->a2, D, Rd, 1);
  s->tm_qkv_d = rows_map(s->qkv, 3 * D, Rd, 1);
  s->tm_ao_d = rows_map(s->ao, D, Rd, 1);
  {
    uint64_t dq[4] = {static_cast<uint64_t>(3 * D), 1, Rd, 1}, sq[3] = {3ull * D, 3ull * D, 3ull * D * Rd};
    uint64_t d_o[4] = {static_cast<uint64_t>(D), 1, Rd, 1}, so[3] = {1ull * D, 1ull * D, 1ull * D * Rd};
    uint32_t bq[4] = {64, 1, 128, 1}, bo[4] = {64, 1, static_cast<uint32_t>((128 / S) * S), 1};
    s->tm_qkv_spk = make_tmap_f16(s->qkv.p, 4, dq, sq, bq);
    uint32_t bkv[4] = {64, 1, 64, 1};
    s->tm_qkv_spk_kv = make_tmap_f16(s->qkv.p, 4, dq, sq, bkv);
    s->tm_ao_spk = make_tmap_f16(s->ao.p, 4, d_o, so, bo);
  }
  stream_alloc(s, 1024);
}

// Small-row frame step (see fseend_fs_stream::rv): same arithmetic as stream_launch below with fp32 activations and
// caches; every product goes through the row-vector kernel (fp16 weights read once from L2, no tensor-core tiles).
void stream_launch_rv(fseend_fs_stream* s, const float* x_in, bool decode, float* y_out, cudaStream_t st) {
  fseend_fs_model* m = s->m;
  const fseend_fs_config& c = m->cfg;
  const int D = c.n_units, B = s->B, S = s->S;
  const int Rd = B * S;
  const float scale = 1.f / sqrtf(64.f);
  const int* enc_pos = static_cast<const int*>(s->ctr.p);
  const int* dec_pos = enc_pos + 1;
  auto F = [](DevBuf& b) { return static_cast<float*>(b.p); };
  // product (+ optional LayerNorm of the 256-wide result into `ln_out`, fused into the same launch)
  auto lin = [&](const float* A, int rows, const WMat& w, const float* bias, int act, const float* res, float* out,
                 const FVec* ln_g = nullptr, const FVec* ln_b = nullptr, float* ln_out = nullptr) {
    P32GemmParams p{};
    p.A = A;
    p.lda = w.K;
    p.a_seq_rows = rows;
    p.rows_per_seq = rows;
    p.n_seq = 1;
    p.k_blocks = w.K / 64;
    p.taps = 1;
    p.N = w.rows;
    p.bias = bias;
    p.act = act;
    p.alpha = 1.f;
    p.w_inv_scale = 1.f;
    p.residual = res;
    p.ldr = w.rows;
    p.out = out;
    p.ldo = w.rows;
    if (ln_out) {
      p.row_epi_counter = static_cast<unsigned int*>(s->epi_ctr.p);
      p.ln_g1 = ln_g->f();
      p.ln_b1 = ln_b->f();
      p.ln_out1 = ln_out;
      p.ln_eps = c.ln_eps;
    }
    if (!launch_p32_rowvec(static_cast<const __half*>(w.buf.p), nullptr, p, st))
      throw std::runtime_error("row-vector kernel rejected the shape");
  };
  if (x_in) {
    launch_p32_pad_input(x_in, static_cast<const int*>(s->cu.p), B, 1, c.in_size, m->Kin, F(s->rX), st, m->bn_scale.f(),
                         m->bn_shift.f(), -1.f);
    lin(F(s->rX), B, m->w_in, m->b_in.f(), P32_NONE, nullptr, F(s->rT1), &m->g_in, &m->be_in, F(s->rh0));
    for (int l = 0; l < c.enc_n_layers; ++l) {
      EncLayer& E = *m->enc[l];
      lin(F(s->rh0), B, E.wqkv, E.bqkv.f(), P32_NONE, nullptr, F(s->rqkv));
      launch_p32_step_attn(F(s->rqkv), F(*s->enc_k32[l]), F(*s->enc_v32[l]), B, s->cap, 0, scale, F(s->rao), st, enc_pos);
      lin(F(s->rao), B, E.wo, E.bo.f(), P32_NONE, F(s->rh0), F(s->rT1), &E.g1, &E.be1, F(s->rh1));
      lin(F(s->rh1), B, E.w1, E.b1.f(), P32_RELU, nullptr, F(s->rf));
      lin(F(s->rf), B, E.w2, E.b2.f(), P32_NONE, F(s->rh1), F(s->rT1), &E.g2, &E.be2, F(s->rh0));
    }
    launch_p32_hist_append(F(s->rh0), F(s->hist32), B, s->cap, 0, st, enc_pos);
  } else {
    launch_p32_hist_append(nullptr, F(s->hist32), B, s->cap, 0, st, enc_pos);
  }
  if (!decode) {
    launch_advance_counters(static_cast<int*>(s->ctr.p), 1, 0, 0, 0, st);
    return;
  }
  {
    const int K = c.conv_kernel, center = K / 2;
    P32GemmParams p{};
    p.A = F(s->hist32);
    p.lda = D;
    p.a_seq_rows = s->cap;
    p.rows_per_seq = 1;
    p.n_seq = B;
    p.k_blocks = D / 64;
    p.taps = K;
    p.tap_shift = -center;
    p.a_row_offset_dev = dec_pos;
    p.N = D;
    p.bias = m->b_conv.f();
    p.alpha = 1.f;
    p.w_inv_scale = 1.f;
    p.out = F(s->remb);
    p.ldo = D;
    p.row_epi_counter = static_cast<unsigned int*>(s->epi_ctr.p);     // L2 normalisation fused (last-CTA ticket)
    p.l2norm = 1;
    if (!launch_p32_rowvec(static_cast<const __half*>(m->w_conv.buf.p), nullptr, p, st))
      throw std::runtime_error("row-vector kernel rejected the conv shape");
  }
  lin(F(s->remb), B, m->w_cvt, nullptr, P32_NONE, nullptr, F(s->rY));
  launch_p32_convert(F(s->rY), m->pe_proj.f(), B, S, F(s->ra0), st);
  for (int l = 0; l < c.dec_n_layers; ++l) {
    DecLayer& Dl = *m->dec[l];
    lin(F(s->ra0), Rd, Dl.wqkv1, Dl.bqkv1.f(), P32_NONE, nullptr, F(s->rqkv));
    launch_p32_step_attn(F(s->rqkv), F(*s->dec_k32[l]), F(*s->dec_v32[l]), Rd, s->cap, 0, scale, F(s->rao), st, dec_pos);
    lin(F(s->rao), Rd, Dl.wo1, Dl.bo1.f(), P32_NONE, F(s->ra0), F(s->rT1), &Dl.g11, &Dl.be11, F(s->ra1));
    lin(F(s->ra1), Rd, Dl.wqkv2, Dl.bqkv2.f(), P32_NONE, nullptr, F(s->rqkv));
    launch_p32_spk_attn(F(s->rqkv), F(s->rao), B, S, scale, st);
    lin(F(s->rao), Rd, Dl.wo2, Dl.bo2.f(), P32_NONE, F(s->ra1), F(s->rT1), &Dl.g21, &Dl.be21, F(s->ra2));
    lin(F(s->ra2), Rd, Dl.w1, Dl.b1.f(), P32_RELU, nullptr, F(s->rf));
    lin(F(s->rf), Rd, Dl.w2, Dl.b2.f(), P32_NONE, F(s->ra2), F(s->rT1), &Dl.g22, &Dl.be22, F(s->ra0));
  }
  launch_p32_head(F(s->remb), F(s->ra0), B, S, y_out, nullptr, nullptr, st);
  launch_advance_counters(static_cast<int*>(s->ctr.p), 1, 1, 0, 0, st);
}

// The kernels of one frame.  Every per-frame index is read from the device counters (ctr[0] = encoder frame index,
// ctr[1] = decoder frame index), so the same launch sequence is valid for every frame and can be replayed from a CUDA
// graph.  x_in: staged frame (nullptr: flush step, zero conv input); decode: the look-ahead window is full.
void stream_launch(fseend_fs_stream* s, const float* x_in, bool decode, float* y_out, cudaStream_t st) {
  if (s->rv) {
    stream_launch_rv(s, x_in, decode, y_out, st);
    return;
  }
  fseend_fs_model* m = s->m;
  const fseend_fs_config& c = m->cfg;
  const int D = c.n_units, B = s->B, S = s->S;
  const float scale = 1.f / sqrtf(64.f);
  CUtensorMap none = s->tm_h0;
  const int* enc_pos = static_cast<const int*>(s->ctr.p);
  const int* dec_pos = enc_pos + 1;
  if (x_in) {
    launch_prep_input(x_in, static_cast<const int*>(s->cu.p), B, 1, c.in_size, m->Kin, m->bn_scale.f(),
                      m->bn_shift.f(), static_cast<__half*>(s->x16.p), st);
    {
      GemmParams p = flat_params(B, D, m->Kin, EPI_LN);
      p.bias = m->b_in.f();
      p.ln_g = m->g_in.f();
      p.ln_b = m->be_in.f();
      p.ln_eps = c.ln_eps;
      launch_gemm(s->tm_x16, m->w_in.tm, none, s->tm_h0, p, st);
    }
    for (int l = 0; l < c.enc_n_layers; ++l) {
      EncLayer& E = *m->enc[l];
      {
        GemmParams p = flat_params(B, 3 * D, D, EPI_BIAS);
        p.bias = E.bqkv.f();
        launch_gemm(s->tm_h0, E.wqkv.tm, none, s->tm_qkv_e, p, st);
      }
      // the streaming encoder attends over every past frame (causal by construction, FS:stream_mod:28-35)
      launch_step_attn(static_cast<const __half*>(s->qkv.p), static_cast<__half*>(s->enc_k[l]->p),
                       static_cast<__half*>(s->enc_v[l]->p), B, s->cap, 0, scale, static_cast<__half*>(s->ao.p), st,
                       enc_pos);
      {
        GemmParams p = flat_params(B, D, D, EPI_LN);
        p.bias = E.bo.f();
        p.has_residual = 1;
        p.ln_g = E.g1.f();
        p.ln_b = E.be1.f();
        p.ln_eps = c.ln_eps;
        launch_gemm(s->tm_ao_e, E.wo.tm, s->tm_h0, s->tm_h1, p, st);
      }
      FfnParams fp = ffn_params(B, 1, c.enc_dim_feedforward, E.b1, E.b2, E.g2, E.be2, c.ln_eps, nullptr);
      launch_ffn(s->tm_h1, E.w1.tm128, E.w2.tm128, s->tm_h0, fp, 3, st);
    }
    launch_hist_append(static_cast<const __half*>(s->h0.p), static_cast<__half*>(s->hist.p), B, s->cap, 0, st, enc_pos);
  } else {
    launch_hist_append(nullptr, static_cast<__half*>(s->hist.p), B, s->cap, 0, st, enc_pos);
  }
  if (!decode) {
    launch_advance_counters(static_cast<int*>(s->ctr.p), 1, 0, 0, 0, st);
    return;
  }
  const int K = c.conv_kernel, center = K / 2;
  {
    // output frame index cidx = ctr[1]; its window is hist[cidx - center .. cidx + center]
    GemmParams p{};
    p.rows_per_seq = 1;
    p.n_seq = B;
    p.tiles_per_seq = 1;
    p.n_tiles = 1;
    p.k_blocks = D / 64;
    p.taps = K;
    p.tap_shift = -center;
    p.a_row_offset = 0;
    p.a_row_offset_dev = dec_pos;
    p.mode = EPI_L2;
    p.bias = m->b_conv.f();
    launch_gemm(s->tm_hist, m->w_conv.tm, none, s->tm_emb_out, p, st);
  }
  {
    GemmParams p = flat_params(B, D, D, EPI_CONVERT);
    p.S = S;
    p.pe_proj = m->pe_proj.f();
    launch_gemm(s->tm_emb_in, m->w_cvt.tm, none, s->tm_cvt_out, p, st);
  }
  const size_t Rd = 1ull * B * S;
  for (int l = 0; l < c.dec_n_layers; ++l) {
    DecLayer& Dl = *m->dec[l];
    {
      GemmParams p = flat_params(Rd, 3 * D, D, EPI_BIAS);
      p.bias = Dl.bqkv1.f();
      launch_gemm(s->tm_a0, Dl.wqkv1.tm, none, s->tm_qkv_d, p, st);
    }
    launch_step_attn(static_cast<const __half*>(s->qkv.p), static_cast<__half*>(s->dec_k[l]->p),
                     static_cast<__half*>(s->dec_v[l]->p), static_cast<int>(Rd), s->cap, 0, scale,
                     static_cast<__half*>(s->ao.p), st, dec_pos);
    {
      GemmParams p = flat_params(Rd, D, D, EPI_LN);
      p.bias = Dl.bo1.f();
      p.has_residual = 1;
      p.ln_g = Dl.g11.f();
      p.ln_b = Dl.be11.f();
      p.ln_eps = c.ln_eps;
      launch_gemm(s->tm_ao_d, Dl.wo1.tm, s->tm_a0, s->tm_a1, p, st);
    }
    {
      SpkFuseParams sp{static_cast<int>(Rd), S, (128 / S) * S, scale, Dl.bqkv2.f(), static_cast<__half*>(s->ao.p)};
      launch_spkfuse(s->tm_a1, Dl.wqkv2.tm64, sp, st);
    }
    {
      GemmParams p = flat_params(Rd, D, D, EPI_LN);
      p.bias = Dl.bo2.f();
      p.has_residual = 1;
      p.ln_g = Dl.g21.f();
      p.ln_b = Dl.be21.f();
      p.ln_eps = c.ln_eps;
      launch_gemm(s->tm_ao_d, Dl.wo2.tm, s->tm_a1, s->tm_a2, p, st);
    }
    FfnParams fp = ffn_params(static_cast<int>(Rd), 1, c.dec_dim_feedforward, Dl.b1, Dl.b2, Dl.g22, Dl.be22, c.ln_eps,
                              nullptr);
    launch_ffn(s->tm_a2, Dl.w1.tm128, Dl.w2.tm128, s->tm_a0, fp, 3, st);
  }
  launch_head(static_cast<const __half*>(s->emb.p), static_cast<const __half*>(s->a0.p), B, S, y_out, nullptr, nullptr,
              st);
  launch_advance_counters(static_cast<int*>(s->ctr.p), 1, 1, 0, 0, st);
}

// One frame.  x_t: device fp32 [B][in_size], or nullptr for a flush step (the reference's dummy_conv_input).
// Returns 1 and writes logits [B][S] when the look-ahead conv has a full window, else 0.
// Steady-state frames replay an instantiated CUDA graph (33 kernel nodes): the per-frame cost is one graph launch plus
// two small device copies instead of 33 launches (FSEEND_STREAM_GRAPH=0 keeps the eager launches).
int stream_step(fseend_fs_stream* s, const float* x_t, float* logits, cudaStream_t st) {
  fseend_fs_model* m = s->m;
  const fseend_fs_config& c = m->cfg;
  const int B = s->B, S = s->S;
  if (B * S > 128) throw std::invalid_argument("streaming supports B * max_nspks <= 128");
  if (s->t + 1 > s->cap) stream_alloc(s, s->cap * 2);      // (synchronises; graphs are re-captured for the new buffers)
  const int center = c.conv_kernel / 2;
  const bool decode = s->t + 1 >= center + 1;
  float* x_in = x_t ? static_cast<float*>(s->x_in.p) : nullptr;
  if (x_t)
    CUDA_CHECK(cudaMemcpyAsync(x_in, x_t, 1ull * B * c.in_size * sizeof(float), cudaMemcpyDeviceToDevice, st));
  float* y_out = static_cast<float*>(s->y_out.p);

  if (s->graph_cap != s->cap) {      // caches were (re)allocated: captured pointers are stale
    if (s->graph_step) cudaGraphExecDestroy(s->graph_step);
    if (s->graph_flush) cudaGraphExecDestroy(s->graph_flush);
    s->graph_step = s->graph_flush = nullptr;
    s->graph_cap = s->cap;
  }
  cudaGraphExec_t* slot = x_t ? &s->graph_step : &s->graph_flush;
  // the first decoder steps run eagerly (lazy per-kernel attribute setup must not happen inside a capture)
  const bool graphable = s->use_graph && decode && s->eager_decodes >= 2;
  if (graphable && *slot == nullptr) {
    // captured on a private stream (the caller's may be the legacy default stream, which cannot capture); the
    // instantiated graph is launched on the caller's stream
    cudaGraph_t g = nullptr;
    if (!s->cap_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeThreadLocal));
    try {
      stream_launch(s, x_in, true, y_out, s->cap_stream);
    } catch (...) {
      cudaStreamEndCapture(s->cap_stream, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    CUDA_CHECK(cudaStreamEndCapture(s->cap_stream, &g));
    cudaError_t e = cudaGraphInstantiate(slot, g, 0);
    cudaGraphDestroy(g);
    CUDA_CHECK(e);
  }
  if (graphable) {
    CUDA_CHECK(cudaGraphLaunch(*slot, st));
  } else {
    stream_launch(s, x_in, decode, y_out, st);
    if (decode) s->eager_decodes += 1;
  }
  s->t += 1;
  CUDA_CHECK(cudaGetLastError());
  if (!decode) return 0;
  CUDA_CHECK(cudaMemcpyAsync(logits, y_out, 1ull * B * S * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 1;
}

void drain_profile(fseend_fs_model* m) {
  for (auto& e : m->prof_pending) {
    float ms = 0.f;
    cudaEventSynchronize(e.b);
    cudaEventElapsedTime(&ms, e.a, e.b);
    auto& acc = m->prof_acc[e.name];
    acc.first += ms;
    acc.second += 1;
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  m->prof_pending.clear();
}

template <class F>
int guarded(F&& f) {
  try {
    f();
    return FSEEND_OK;
  } catch (const std::invalid_argument& e) {
    set_last_error(e.what());
    return FSEEND_ERR_INVALID;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return FSEEND_ERR_CUDA;
  }
}

}  // namespace
}  // namespace fseend

// =============================================================================================== C ABI
extern "C" {

int fseend_version(void) { return FSEEND_VERSION; }
const char* fseend_last_error(void) { return g_last_error.c_str(); }

int fseend_device_ok(void) {
  // cudaDeviceGetAttribute is a cheap query (cudaGetDeviceProperties costs milliseconds per call, which dominated the
  // single-kernel entry points)
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int fseend_fs_create(const fseend_fs_config* cfg, int n_tensors, const char* const* names, const float* const* data,
                     const long long* numel, fseend_fs_model** out) {
  if (!cfg || !out || !names || !data || !numel) {
    set_last_error("null argument");
    return FSEEND_ERR_INVALID;
  }
  *out = nullptr;
  if (cfg->n_units != 256 || cfg->n_heads * 64 != cfg->n_units) {
    set_last_error("only n_units=256 with head_dim 64 is supported");
    return FSEEND_ERR_INVALID;
  }
  if (cfg->enc_dim_feedforward % 256 || cfg->dec_dim_feedforward % 256 || cfg->in_size < 1 || cfg->conv_kernel < 1 ||
      cfg->enc_n_layers < 1 || cfg->dec_n_layers < 0) {
    set_last_error("unsupported configuration (feed-forward widths must be multiples of 256)");
    return FSEEND_ERR_INVALID;
  }
  if (cfg->conv_kernel != 2 * cfg->conv_padding + 1) {
    // the forward assumes a 'same' convolution (T frames in, T frames out); the reference hard-codes padding = 9 with
    // kernel = 2 * conv_delay + 1 (FS:model:30), so any other pairing changes the output length there
    set_last_error("conv_kernel must equal 2 * conv_padding + 1 ('same' look-ahead convolution)");
    return FSEEND_ERR_INVALID;
  }
  if (!fseend_device_ok()) {
    set_last_error("fseend_b200 requires an sm_100 (B200) device; no CPU fallback exists");
    return FSEEND_ERR_NO_DEVICE;
  }
  auto* m = new fseend_fs_model();
  m->cfg = *cfg;
  int rc = guarded([&] {
    TensorTable tt;
    for (int i = 0; i < n_tensors; ++i) tt.t[names[i]] = {data[i], numel[i]};
    build_model(m, tt);
  });
  if (rc != FSEEND_OK) {
    delete m;
    // a missing tensor is reported distinctly
    if (rc == FSEEND_ERR_INVALID && g_last_error.find("state_dict") != std::string::npos) rc = FSEEND_ERR_MISSING;
    return rc;
  }
  *out = m;
  return FSEEND_OK;
}

void fseend_fs_destroy(fseend_fs_model* m) {
  if (!m) return;
  drain_profile(m);
  delete m;
}

int fseend_fs_forward(fseend_fs_model* m, const float* x_packed_dev, const int* ilens_host, int B, int max_nspks,
                      float* logits_dev, float* emb_dev, float* att_dev, void* stream) {
  if (!m || !x_packed_dev || !ilens_host || !logits_dev) {
    set_last_error("null argument");
    return FSEEND_ERR_INVALID;
  }
  return guarded([&] {
    forward_impl(m, x_packed_dev, ilens_host, B, max_nspks, logits_dev, emb_dev, att_dev,
                 static_cast<cudaStream_t>(stream));
  });
}

int fseend_fs_forward_host(fseend_fs_model* m, const float* x_packed_host, const int* ilens_host, int B,
                           int max_nspks, float* logits_host, float* emb_host, float* att_host) {
  if (!m || !x_packed_host || !ilens_host || !logits_host) {
    set_last_error("null argument");
    return FSEEND_ERR_INVALID;
  }
  return guarded([&] {
    long long total = 0;
    int T = 0;
    for (int b = 0; b < B; ++b) {
      total += ilens_host[b];
      T = ilens_host[b] > T ? ilens_host[b] : T;
    }
    const size_t xin = static_cast<size_t>(total) * m->cfg.in_size * sizeof(float);
    const size_t n_log = 1ull * B * T * max_nspks;
    const size_t n_emb = emb_host ? 1ull * B * T * m->cfg.n_units : 0;
    const size_t n_att = att_host ? 1ull * B * T * max_nspks * m->cfg.n_units : 0;
    // The batch is split into chunks of whole sequences that are processed back to back on a compute stream while the
    // copy stream is already fetching the next chunk's features: the PCIe transfer (the larger part of what a host-
    // buffer call adds to a device-buffer call) hides behind the kernels of the previous chunk.  All chunks are padded
    // to the batch's longest sequence, so they share one plan and fill one [B][T][S] result.
    int nc = m->host_chunks;
    // measured (B = 64, T = 500, S = 6, same box): 1 chunk 3.72 ms, 2 chunks 3.47 ms, 4 chunks 4.2 ms per call — small
    // chunks leave SMs idle in the encoder kernels, so only large batches are split, and only in two
    if (nc <= 0) nc = (B >= 32) ? 2 : 1;
    if (nc > fseend_fs_model::kMaxHostChunks) nc = fseend_fs_model::kMaxHostChunks;
    while (nc > 1 && B % nc) --nc;
    const int Bc = B / nc;
    // staging buffers live next to the plan; (re)allocated on growth (make_plan frees them on a shape change)
    make_plan(m, Bc, T, max_nspks);
    if (m->x_stage.bytes < xin) m->x_stage.alloc(xin);
    const size_t out_bytes = (n_log + n_emb + n_att) * sizeof(float);
    if (m->logits_stage.bytes < out_bytes) m->logits_stage.alloc(out_bytes);
    if (!m->compute_stream) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&m->compute_stream, cudaStreamNonBlocking));
      CUDA_CHECK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    }
    cudaStream_t st = m->compute_stream;
    float* dl = static_cast<float*>(m->logits_stage.p);
    float* de = emb_host ? dl + n_log : nullptr;
    float* da = att_host ? dl + n_log + n_emb : nullptr;
    size_t row0 = 0;
    std::vector<size_t> chunk_row0(nc);
    for (int c = 0; c < nc; ++c) {
      size_t rows = 0;
      for (int b = c * Bc; b < (c + 1) * Bc; ++b) rows += ilens_host[b];
      chunk_row0[c] = row0;
      const size_t off = row0 * m->cfg.in_size;
      CUDA_CHECK(cudaMemcpyAsync(static_cast<float*>(m->x_stage.p) + off, x_packed_host + off,
                                 rows * m->cfg.in_size * sizeof(float), cudaMemcpyHostToDevice, m->copy_stream));
      if (!m->chunk_ev[c]) CUDA_CHECK(cudaEventCreateWithFlags(&m->chunk_ev[c], cudaEventDisableTiming));
      CUDA_CHECK(cudaEventRecord(m->chunk_ev[c], m->copy_stream));
      row0 += rows;
    }
    for (int c = 0; c < nc; ++c) {
      CUDA_CHECK(cudaStreamWaitEvent(st, m->chunk_ev[c], 0));
      const size_t o = 1ull * c * Bc * T;
      forward_impl(m, static_cast<const float*>(m->x_stage.p) + chunk_row0[c] * m->cfg.in_size, ilens_host + c * Bc, Bc,
                   max_nspks, dl + o * max_nspks, de ? de + o * m->cfg.n_units : nullptr,
                   da ? da + o * max_nspks * m->cfg.n_units : nullptr, st, T);
    }
    CUDA_CHECK(cudaMemcpyAsync(logits_host, dl, n_log * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (emb_host) CUDA_CHECK(cudaMemcpyAsync(emb_host, de, n_emb * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (att_host) CUDA_CHECK(cudaMemcpyAsync(att_host, da, n_att * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

// Pipelined host-buffer forward: enqueue and return a ticket; fseend_fs_host_wait(ticket) blocks until that call's logits
// are in logits_host.  Two calls may be in flight: while call i runs its kernels on the compute stream, the copy stream
// already moves call i+1's features into the other staging buffer, so in steady state a step costs max(copy, compute)
// instead of copy + compute.  x_packed_host / logits_host must stay valid (and should be pinned) until the wait.
int fseend_fs_forward_host_async(fseend_fs_model* m, const float* x_packed_host, const int* ilens_host, int B,
                                 int max_nspks, float* logits_host, long long* ticket) {
  if (!m || !x_packed_host || !ilens_host || !logits_host || !ticket) {
    set_last_error("null argument");
    return FSEEND_ERR_INVALID;
  }
  return guarded([&] {
    if (B < 1) throw std::invalid_argument("B must be >= 1");
    long long total = 0;
    int T = 0;
    for (int b = 0; b < B; ++b) {
      if (ilens_host[b] < 1) throw std::invalid_argument("ilens must be >= 1");
      total += ilens_host[b];
      T = ilens_host[b] > T ? ilens_host[b] : T;
    }
    const size_t xin = static_cast<size_t>(total) * m->cfg.in_size * sizeof(float);
    const size_t n_log = 1ull * B * T * max_nspks;
    if (!m->compute_stream) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&m->compute_stream, cudaStreamNonBlocking));
      CUDA_CHECK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    }
    const long long tk = m->host_calls;
    const int slot = static_cast<int>(tk % fseend_fs_model::kHostDepth);
    if (!m->h2d_ev[slot]) {
      CUDA_CHECK(cudaEventCreateWithFlags(&m->h2d_ev[slot], cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&m->done_ev[slot], cudaEventDisableTiming));
    } else {
      // the call that used this slot last (ticket tk - kHostDepth) must have finished before its staging is reused
      CUDA_CHECK(cudaEventSynchronize(m->done_ev[slot]));
    }
    if (m->hx[slot].bytes < xin) m->hx[slot].alloc(xin);
    if (m->hl[slot].bytes < n_log * sizeof(float)) m->hl[slot].alloc(n_log * sizeof(float));
    CUDA_CHECK(cudaMemcpyAsync(m->hx[slot].p, x_packed_host, xin, cudaMemcpyHostToDevice, m->copy_stream));
    CUDA_CHECK(cudaEventRecord(m->h2d_ev[slot], m->copy_stream));
    cudaStream_t st = m->compute_stream;
    CUDA_CHECK(cudaStreamWaitEvent(st, m->h2d_ev[slot], 0));
    forward_impl(m, static_cast<const float*>(m->hx[slot].p), ilens_host, B, max_nspks, static_cast<float*>(m->hl[slot].p),
                 nullptr, nullptr, st);
    CUDA_CHECK(cudaMemcpyAsync(logits_host, m->hl[slot].p, n_log * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaEventRecord(m->done_ev[slot], st));
    m->host_calls = tk + 1;
    *ticket = tk;
  });
}

int fseend_fs_host_wait(fseend_fs_model* m, long long ticket) {
  if (!m) return FSEEND_ERR_INVALID;
  return guarded([&] {
    if (ticket < 0 || ticket >= m->host_calls) throw std::invalid_argument("unknown ticket");
    if (ticket < m->host_calls - fseend_fs_model::kHostDepth)
      return;   // older than the calls in flight: its slot was synchronised when it was reused
    const int slot = static_cast<int>(ticket % fseend_fs_model::kHostDepth);
    CUDA_CHECK(cudaEventSynchronize(m->done_ev[slot]));
  });
}

int fseend_fs_set_profiling(fseend_fs_model* m, int on) {
  if (!m) return FSEEND_ERR_INVALID;
  drain_profile(m);
  if (on && !m->profiling) m->prof_acc.clear();
  m->profiling = on != 0;
  return FSEEND_OK;
}

int fseend_fs_get_profile(fseend_fs_model* m, int max_entries, char (*names)[32], float* total_ms, int* launches) {
  if (!m) return FSEEND_ERR_INVALID;
  drain_profile(m);
  int i = 0;
  for (auto& kv : m->prof_acc) {
    if (i >= max_entries) break;
    strncpy(names[i], kv.first.c_str(), 31);
    names[i][31] = 0;
    total_ms[i] = static_cast<float>(kv.second.first);
    launches[i] = kv.second.second;
    ++i;
  }
  return i;
}

int fseend_fs_launches_per_forward(const fseend_fs_model* m) { return m ? m->launches_last : 0; }
size_t fseend_fs_workspace_bytes(const fseend_fs_model* m) { return m ? m->ws_bytes : 0; }

// ---------------------------------------------------------------------------- streaming entry points
int fseend_fs_stream_create(fseend_fs_model* m, int B, int max_nspks, fseend_fs_stream** out) {
  if (!m || !out) return FSEEND_ERR_INVALID;
  *out = nullptr;
  auto* s = new fseend_fs_stream();
  int rc = guarded([&] {
    if (B < 1 || max_nspks < 1 || max_nspks > 16 || B * max_nspks > 128)
      throw std::invalid_argument("streaming needs 1 <= B * max_nspks <= 128, max_nspks <= 16");
    if (!m->cfg.has_mask || m->cfg.mask_delay != 0)
      throw std::invalid_argument("streaming inference is causal: has_mask must be true and mask_delay 0");
    stream_init(s, m, B, max_nspks);
  });
  if (rc != FSEEND_OK) {
    delete s;
    return rc;
  }
  *out = s;
  return FSEEND_OK;
}

void fseend_fs_stream_destroy(fseend_fs_stream* s) { delete s; }

int fseend_fs_stream_step(fseend_fs_stream* s, const float* x_t_dev, float* logits_dev, int* produced, void* stream) {
  if (!s || !logits_dev || !produced) return FSEEND_ERR_INVALID;
  return guarded([&] { *produced = stream_step(s, x_t_dev, logits_dev, static_cast<cudaStream_t>(stream)); });
}

int fseend_fs_stream_frames(const fseend_fs_stream* s) { return s ? s->t : 0; }

// ---------------------------------------------------------------------------- single-kernel entry points
int fseend_op_gemm(const void* a_f16, int rows_per_seq, int n_seq, int K, const void* w_f16, int N, int taps,
                   int tap_shift, int mode, int relu, const float* bias, const void* residual_f16, const float* ln_g,
                   const float* ln_b, float ln_eps, const float* pe_proj, int S, const int* seq_len_dev,
                   void* out_f16, void* stream) {
  return guarded([&] {
    if (K % 64 || N % 256 || (mode != EPI_BIAS && N != 256)) throw std::invalid_argument("K%64, N%256 required");
    if (!fseend_device_ok()) throw std::invalid_argument("sm_100 device required");
    GemmParams p{};
    p.rows_per_seq = rows_per_seq;
    p.n_seq = n_seq;
    p.tiles_per_seq = (rows_per_seq + 127) / 128;
    p.n_tiles = N / 256;
    p.k_blocks = K / 64;
    p.taps = taps;
    p.tap_shift = tap_shift;
    p.mode = mode;
    p.relu = relu;
    p.has_residual = residual_f16 != nullptr;
    p.S = S;
    p.ln_eps = ln_eps;
    p.bias = bias;
    p.ln_g = ln_g;
    p.ln_b = ln_b;
    p.pe_proj = pe_proj;
    p.seq_len = seq_len_dev;
    CUtensorMap tmA = make_tmap_rows3d(a_f16, K, K, rows_per_seq, n_seq, 128);
    uint64_t wd[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(taps) * N};
    uint64_t ws[1] = {static_cast<uint64_t>(K)};
    uint32_t wb[2] = {64, 256};
    CUtensorMap tmB = make_tmap_f16(w_f16, 2, wd, ws, wb);
    uint32_t wb_half[2] = {64, 128};
    CUtensorMap tmBh = make_tmap_f16(w_f16, 2, wd, ws, wb_half);
    p.tmB_half = &tmBh;
    CUtensorMap tmO;
    if (mode == EPI_CONVERT) {
      uint64_t dims[3] = {256, static_cast<uint64_t>(S), static_cast<uint64_t>(rows_per_seq) * n_seq};
      uint64_t str[2] = {256, 256ull * S};
      uint32_t box[3] = {64, 1, 128};
      tmO = make_tmap_f16(out_f16, 3, dims, str, box);
    } else {
      tmO = make_tmap_rows3d(out_f16, N, N, rows_per_seq, n_seq, 128);
    }
    CUtensorMap tmR = residual_f16 ? make_tmap_rows3d(residual_f16, N, N, rows_per_seq, n_seq, 128) : tmO;
    launch_gemm(tmA, tmB, tmR, tmO, p, static_cast<cudaStream_t>(stream));
    CUDA_CHECK(cudaGetLastError());
  });
}

int fseend_fs_set_option(fseend_fs_model* m, const char* key, int value) {
  if (!m || !key) return FSEEND_ERR_INVALID;
  if (strcmp(key, "ffn") == 0 && value >= 0 && value <= 5) {
    m->ffn_mode = value;
    return FSEEND_OK;
  }
  if (strcmp(key, "spk") == 0 && value >= 0 && value <= 2) {
    m->spk_mode = value;
    return FSEEND_OK;
  }
  if (strcmp(key, "host_chunks") == 0 && value >= 0 && value <= fseend_fs_model::kMaxHostChunks) {
    m->host_chunks = value;
    return FSEEND_OK;
  }
  set_last_error(std::string("unknown option or bad value: ") + key);
  return FSEEND_ERR_INVALID;
}

int fseend_op_ffn(const void* x_f16, int rows_per_seq, int n_seq, const void* w1_f16, const float* b1,
                  const void* w2_f16, const float* b2, int F, const float* ln_g, const float* ln_b, float ln_eps,
                  const int* seq_len_dev, int cluster, void* out_f16, void* stream) {
  return guarded([&] {
    if (F % 128 || F < 128) throw std::invalid_argument("F must be a multiple of 128");
    if (cluster < 1 || cluster > 5) throw std::invalid_argument("variant must be 1..5");
    if (!fseend_device_ok()) throw std::invalid_argument("sm_100 device required");
    FfnParams p{};
    p.rows_per_seq = rows_per_seq;
    p.n_seq = n_seq;
    p.tiles_per_seq = (rows_per_seq + 127) / 128;
    p.F = F;
    p.ln_eps = ln_eps;
    p.b1 = b1;
    p.b2 = b2;
    p.ln_g = ln_g;
    p.ln_b = ln_b;
    p.seq_len = seq_len_dev;
    CUtensorMap tmX = make_tmap_rows3d(x_f16, 256, 256, rows_per_seq, n_seq, 128);
    CUtensorMap tmO = make_tmap_rows3d(out_f16, 256, 256, rows_per_seq, n_seq, 128);
    uint32_t box[2] = {64, 128};
    uint64_t d1[2] = {256, static_cast<uint64_t>(F)}, s1[1] = {256};
    uint64_t d2[2] = {static_cast<uint64_t>(F), 256}, s2[1] = {static_cast<uint64_t>(F)};
    uint32_t box64[2] = {64, 64};
    CUtensorMap tmW1 = make_tmap_f16(w1_f16, 2, d1, s1, cluster == 5 ? box64 : box);
    CUtensorMap tmW2 = make_tmap_f16(w2_f16, 2, d2, s2, box);
    if (cluster == 5) launch_ffn_pair(tmX, tmW1, tmW2, tmO, p, static_cast<cudaStream_t>(stream));
    else launch_ffn(tmX, tmW1, tmW2, tmO, p, cluster, static_cast<cudaStream_t>(stream));
    CUDA_CHECK(cudaGetLastError());
  });
}

int fseend_op_causal_attn(const void* qkv_f16, int B, int T, int S, int H, int mask_delay, float scale, void* out_f16,
                          void* stream) {
  return guarded([&] {
    if (H != 4) throw std::invalid_argument("H must be 4 (4 x 64 = 256)");
    if (!fseend_device_ok()) throw std::invalid_argument("sm_100 device required");
    uint64_t dq[4] = {768, static_cast<uint64_t>(S), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    uint64_t sq[3] = {768, 768ull * S, 768ull * S * T};
    uint64_t d_o[4] = {256, static_cast<uint64_t>(S), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    uint64_t so[3] = {256, 256ull * S, 256ull * S * T};
    uint32_t box[4] = {64, 1, 128, 1};
    CUtensorMap tq = make_tmap_f16(qkv_f16, 4, dq, sq, box);
    uint32_t bkv[4] = {64, 1, 64, 1};
    CUtensorMap tkv = make_tmap_f16(qkv_f16, 4, dq, sq, bkv);
    CUtensorMap to = make_tmap_f16(out_f16, 4, d_o, so, box);
    AttnParams a{B, S, T, H, mask_delay, scale, ATTN_CAUSAL, 128};
    launch_attn(tq, tkv, static_cast<__half*>(out_f16), a, static_cast<cudaStream_t>(stream));
    CUDA_CHECK(cudaGetLastError());
  });
}

int fseend_op_spk_attn_tc(const void* qkv_f16, int n_frames, int S, float scale, void* out_f16, void* stream) {
  return guarded([&] {
    if (S < 1 || S > 16) throw std::invalid_argument("S must be in [1,16]");
    if (!fseend_device_ok()) throw std::invalid_argument("sm_100 device required");
    const uint64_t rows = 1ull * n_frames * S;
    uint64_t dq[4] = {768, 1, rows, 1}, sq[3] = {768, 768, 768 * rows};
    uint64_t d_o[4] = {256, 1, rows, 1}, so[3] = {256, 256, 256 * rows};
    uint32_t bq[4] = {64, 1, 128, 1}, bo[4] = {64, 1, static_cast<uint32_t>((128 / S) * S), 1};
    CUtensorMap tq = make_tmap_f16(qkv_f16, 4, dq, sq, bq);
    uint32_t bkv[4] = {64, 1, 64, 1};
    CUtensorMap tkv = make_tmap_f16(qkv_f16, 4, dq, sq, bkv);
    CUtensorMap to = make_tmap_f16(out_f16, 4, d_o, so, bo);
    AttnParams a{1, S, static_cast<int>(rows), 4, 0, scale, ATTN_BLOCKDIAG, (128 / S) * S};
    launch_attn(tq, tkv, static_cast<__half*>(out_f16), a, static_cast<cudaStream_t>(stream));
    CUDA_CHECK(cudaGetLastError());
  });
}

int fseend_op_spk_attn(const void* qkv_f16, int n_frames, int S, float scale, void* out_f16, void* stream) {
  return guarded([&] {
    if (launch_spk_attn(static_cast<const __half*>(qkv_f16), static_cast<__half*>(out_f16), n_frames, S, scale,
                        static_cast<cudaStream_t>(stream)) != 0)
      throw std::invalid_argument("S must be in [1,16]");
    CUDA_CHECK(cudaGetLastError());
  });
}

int fseend_op_head(const void* emb_f16, const void* att_f16, int n_frames, int S, float* logits, float* emb_f32,
                   float* att_f32, void* stream) {
  return guarded([&] {
    launch_head(static_cast<const __half*>(emb_f16), static_cast<const __half*>(att_f16), n_frames, S, logits, emb_f32,
                att_f32, static_cast<cudaStream_t>(stream));
    CUDA_CHECK(cudaGetLastError());
  });
}

int fseend_op_prep_input(const float* x_packed, const int* cu_seqlens_dev, int B, int Tmax, int Din, int Kpad,
                         const float* scale, const float* shift, void* out_f16, void* stream) {
  return guarded([&] {
    if (Kpad % 2 || Kpad < Din) throw std::invalid_argument("Kpad must be even and >= Din");
    launch_prep_input(x_packed, cu_seqlens_dev, B, Tmax, Din, Kpad, scale, shift, static_cast<__half*>(out_f16),
                      static_cast<cudaStream_t>(stream));
    CUDA_CHECK(cudaGetLastError());
  });
}

int fseend_op_spk_qkv_attn(const void* x_f16, const void* w_f16, const float* bias, int n_frames, int S, float scale,
                           void* out_f16, void* stream) {
  return guarded([&] {
    if (S < 1 || S > 16) throw std::invalid_argument("S must be in [1,16]");
    if (!fseend_device_ok()) throw std::invalid_argument("sm_100 device required");
    const uint64_t rows = 1ull * n_frames * S;
    CUtensorMap tmX = make_tmap_rows3d(x_f16, 256, 256, rows, 1, 128);
    uint64_t dw[2] = {256, 768}, sw[1] = {256};
    uint32_t bw[2] = {64, 64};
    CUtensorMap tmW = make_tmap_f16(w_f16, 2, dw, sw, bw);
    SpkFuseParams sp{static_cast<int>(rows), S, (128 / S) * S, scale, bias, static_cast<__half*>(out_f16)};
    launch_spkfuse(tmX, tmW, sp, static_cast<cudaStream_t>(stream));
    CUDA_CHECK(cudaGetLastError());
  });
}

int fseend_op_label_prepare(const float* labels, int B, int T, int n_spk, int* perm, float* labels_out, void* stream) {
  return guarded([&] {
    if (B < 1 || T < 1) throw std::invalid_argument("label_prepare: need B, T >= 1");
    if (launch_label_prepare(labels, B, T, n_spk, perm, labels_out, static_cast<cudaStream_t>(stream)) != 0)
      throw std::invalid_argument("label_prepare: n_spk must be in [1, 14]");
    CUDA_CHECK(cudaGetLastError());
  });
}

size_t fseend_op_bce_loss_workspace_bytes(int B, int T) {
  return sizeof(float) * static_cast<size_t>(B) * static_cast<size_t>(bce_loss_chunks(T));
}

int fseend_op_bce_loss(const float* logits, int ld_logits, const float* target, int ld_target, int B, int T,
                       const int* lens_dev, const int* n_cls_dev, int label_delay, float* workspace, float* loss_dev,
                       void* stream) {
  return guarded([&] {
    if (B < 1 || T < 1 || label_delay < 0) throw std::invalid_argument("bce_loss: need B, T >= 1, label_delay >= 0");
    launch_bce_loss(logits, ld_logits, target, ld_target, B, T, lens_dev, n_cls_dev, label_delay, workspace, loss_dev,
                    static_cast<cudaStream_t>(stream));
    CUDA_CHECK(cudaGetLastError());
  });
}

int fseend_op_pit_costs(const float* logits, const float* labels, int B, int T, int C, const int* lens_dev,
                        int label_delay, int pad_term, double* cost_dev, void* stream) {
  return guarded([&] {
    if (launch_pit_costs(logits, labels, B, T, C, lens_dev, label_delay, pad_term, cost_dev,
                         static_cast<cudaStream_t>(stream)) != 0)
      throw std::invalid_argument("pit_costs: need B, T >= 1, 1 <= C <= 16, label_delay >= 0");
    CUDA_CHECK(cudaGetLastError());
  });
}

int fseend_op_splice_subsample(const float* feat, int T, int F, int context_size, int subsampling, float* out,
                               void* stream) {
  return guarded([&] {
    if (T < 0 || F < 1 || context_size < 0 || subsampling < 1)
      throw std::invalid_argument("splice_subsample: need T >= 0, F >= 1, context_size >= 0, subsampling >= 1");
    launch_splice_subsample(feat, T, F, context_size, subsampling, out, static_cast<cudaStream_t>(stream));
    CUDA_CHECK(cudaGetLastError());
  });
}

int fseend_op_decide_median(const float* pred, int T, int C, float threshold, int median, unsigned char* decisions,
                            void* stream) {
  return guarded([&] {
    if (T < 0 || C < 1) throw std::invalid_argument("decide_median: need T >= 0, C >= 1");
    if (median > 1 && (median % 2) == 0) throw std::invalid_argument("decide_median: the median width must be odd");
    launch_decide_median(pred, T, C, threshold, median, decisions, static_cast<cudaStream_t>(stream));
    CUDA_CHECK(cudaGetLastError());
  });
}

size_t fseend_op_embloss_workspace_bytes(int B, int T) {
  return sizeof(float) * static_cast<size_t>(embloss_num_partials(B, T));
}

int fseend_op_embloss(const float* emb_f32, const float* labels, const int* seq_len_dev, int B, int T, int S,
                      double divisor, float* workspace, float* loss_dev, void* stream) {
  return guarded([&] {
    if (B < 1 || T < 1 || S < 1 || S > 16) throw std::invalid_argument("embloss: need B,T >= 1 and 1 <= S <= 16");
    if (!(divisor > 0.0)) throw std::invalid_argument("embloss: divisor must be positive");
    if ((reinterpret_cast<uintptr_t>(emb_f32) & 15) != 0) throw std::invalid_argument("embloss: emb must be 16-byte aligned");
    if (!fseend_device_ok()) throw std::invalid_argument("sm_100 device required");
    EmbLossParams p{};
    p.emb = emb_f32;
    p.labels = labels;
    p.seq_len = seq_len_dev;
    p.partials = workspace;
    p.B = B;
    p.T = T;
    p.S = S;
    launch_embloss(p, divisor, loss_dev, static_cast<cudaStream_t>(stream));
    CUDA_CHECK(cudaGetLastError());
  });
}

}  // extern "C"

#include "ls_model.inc"
