"""Secondary measurements (not the headline bench line): LS-EEND batch throughput at BASELINE configs[2] shape,
FS-EEND and LS-EEND frame-by-frame latency.  Prints one JSON object."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch
from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
from nnet.model.streaming_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import StreamingTransformerEDADiarization
from nnet.model.onl_conformer_retention_enc_1dcnn_tfm_retention_enc_linear_non_autoreg_pos_enc_l2norm_emb_loss_mask import OnlineConformerRetentionDADiarization
from nnet.utils.copy_params import copy_params_from_masked_to_streaming

def ev_time(fn, warm=3, iters=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

out = {}
torch.manual_seed(0)
ls = OnlineConformerRetentionDADiarization(n_speakers=8, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
        dropout=0.1, max_seqlen=1000, recurrent_chunk_size=500, feed_forward_expansion_factor=4, dec_dim_feedforward=2048,
        conv_kernel_size=16).cuda().eval()
B, T, S = 16, 2000, 10
x = torch.randn(B * T, 345, device="cuda")
nat = ls.native()
ms = ev_time(lambda: nat.forward(x, [T] * B, S))
out["ls_batch_B16_T2000_S10"] = {"ms_per_forward": ms, "frames_per_s": B * T / ms * 1e3, "launches": nat.launches_per_forward}
# LS one-step latency (fused native step), B=1, S=10
st = ls.new_stream(1, 10)
xt = torch.randn(1, 345, device="cuda")
for _ in range(30): st.step(xt)
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 300
for _ in range(n): st.step(xt)
torch.cuda.synchronize()
lat = (time.perf_counter() - t0) / n
out["ls_one_step_B1_S10"] = {"ms_per_frame": lat * 1e3, "real_time_factor": lat / 0.1}
if os.environ.get("FSEEND_EXTRA_LONG"):
    # BASELINE configs[4]: one hour of audio (T = 36000 frames of 100 ms), single GPU, frame by frame incl. the flush
    st.reset()
    feats = torch.randn(36000, 345, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(36000): st.step(feats[t:t + 1])
    for _ in range(9): st.step(None)
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    out["ls_streaming_1hour_T36000_S10"] = {"seconds": tot, "ms_per_frame": tot / 36000 * 1e3, "real_time_factor": tot / 3600.0}
# FS streaming latency at t ~ 500 and t ~ 2000 (attention over the growing cache)
fs = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
        dropout=0.1, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048).cuda().eval()
sfs = StreamingTransformerEDADiarization(in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2, dropout=0.1,
        has_mask=True, max_seqlen=500, dec_dim_feedforward=2048).cuda().eval()
copy_params_from_masked_to_streaming(fs, sfs)
xt3 = torch.randn(1, 1, 345, device="cuda")
res = {}
torch.cuda.synchronize(); t0 = time.perf_counter()
for t in range(2000):
    sfs.test(xt3, 6)
    if t + 1 in (500, 2000):
        torch.cuda.synchronize(); res[t + 1] = (time.perf_counter() - t0)
out["fs_streaming_B1_S6"] = {"ms_per_frame_first_500": res[500] / 500 * 1e3, "ms_per_frame_avg_2000": res[2000] / 2000 * 1e3,
                              "real_time_factor_avg_2000": res[2000] / 2000 / 0.1}
print(json.dumps(out))
