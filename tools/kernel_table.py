"""Per-kernel CUDA-event table of one FS-EEND forward at the bench shape (profiling mode of the library)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch  # noqa: E402
from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization  # noqa: E402

torch.manual_seed(0)
m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
                                   dropout=0.1, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048).cuda().eval()
B, T, S = 64, 500, 6
xs = [torch.randn(B * T, 345, device="cuda") for _ in range(4)]
nat = m.native()
for i in range(5):
    nat.forward(xs[i % 4], [T] * B, S)
nat.set_profiling(True)
for i in range(5):
    nat.forward(xs[i % 4], [T] * B, S)
torch.cuda.synchronize()
prof = nat.get_profile()
tot = sum(v[0] for v in prof.values()) / 5
print(f"sum of kernels {tot:.3f} ms")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:20s} {v[0] / v[1]:.4f} ms x {v[1] // 5}  = {100 * v[0] / 5 / tot:5.1f} %")
