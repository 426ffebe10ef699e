"""LS-EEND batch forward at the BASELINE configs[2] shape (B=16, T=2000, S=10): timing, or an ncu target with argv[1]=n."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch
from nnet.model.onl_conformer_retention_enc_1dcnn_tfm_retention_enc_linear_non_autoreg_pos_enc_l2norm_emb_loss_mask import OnlineConformerRetentionDADiarization
torch.manual_seed(0)
ls = OnlineConformerRetentionDADiarization(n_speakers=8, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
        dropout=0.1, max_seqlen=1000, recurrent_chunk_size=500, feed_forward_expansion_factor=4, dec_dim_feedforward=2048,
        conv_kernel_size=16).cuda().eval()
B, T, S = 16, 2000, 10
x = torch.randn(B * T, 345, device="cuda")
nat = ls.native()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if n:
    for _ in range(n): nat.forward(x, [T] * B, S)
    torch.cuda.synchronize()
else:
    for _ in range(5): nat.forward(x, [T] * B, S)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): nat.forward(x, [T] * B, S)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"LS forward B={B} T={T} S={S}: {ms:.3f} ms  {B*T/ms*1e3/1e6:.2f} M frames/s  launches {nat.launches_per_forward}")
