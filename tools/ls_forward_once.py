"""One LS-EEND forward at the BASELINE configs[2] shape (B=16, T=2000, S=10) for ncu captures.
    python tools/ls_forward_once.py [fp32|fp16] [n_forwards]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch  # noqa: E402
from nnet.model.onl_conformer_retention_enc_1dcnn_tfm_retention_enc_linear_non_autoreg_pos_enc_l2norm_emb_loss_mask import (  # noqa: E402
    OnlineConformerRetentionDADiarization)

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
m = OnlineConformerRetentionDADiarization(
    n_speakers=8, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2, dropout=0.1, max_seqlen=1000,
    recurrent_chunk_size=500, feed_forward_expansion_factor=4, dec_dim_feedforward=2048, conv_kernel_size=16).cuda().eval()
m.set_precision(mode)
B, T, S = 16, 2000, 10
x = torch.randn(B * T, 345, device="cuda")
nat = m.native()
for _ in range(n):
    nat.forward(x, [T] * B, S)
torch.cuda.synchronize()
print(mode, "launches per forward", nat.launches_per_forward)
if os.environ.get("FSEEND_TIME"):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        nat.forward(x, [T] * B, S)
    b.record()
    torch.cuda.synchronize()
    print(mode, "ms per forward", a.elapsed_time(b) / 5)
