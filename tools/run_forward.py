"""Run a few B=64,T=500,S=6 forwards (profiling target for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch
from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Bn = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(0)
m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
                                   dropout=0.1, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048).cuda().eval()
x = torch.randn(Bn * 500, 345, device="cuda")
nat = m.native()
for _ in range(n):
    y, _, _ = nat.forward(x, [500] * Bn, 6)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
