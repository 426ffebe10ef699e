"""N FS-EEND forwards at the bench shape (B=64, T=500, S=6) for ncu captures.   python tools/fs_forward_once.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch  # noqa: E402
from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
                                   dropout=0.1, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048).cuda().eval()
B, T, S = 64, 500, 6
x = torch.randn(B * T, 345, device="cuda")
nat = m.native()
for kv in filter(None, os.environ.get("FSEEND_OPTS", "").split(",")):
    k, v = kv.split("=")
    nat.set_option(k, int(v))
for _ in range(n):
    nat.forward(x, [T] * B, S)
torch.cuda.synchronize()
print("launches per forward", nat.launches_per_forward)
if os.environ.get("FSEEND_TIME"):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        nat.forward(x, [T] * B, S)
    b.record()
    torch.cuda.synchronize()
    print("ms per forward", a.elapsed_time(b) / 20)
