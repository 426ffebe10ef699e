"""Per-kernel time vs batch size (number of tiles in flight): flat per-wave time => per-SM bound;
growing with concurrency => shared-resource (L2 / HBM / fabric) bound."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch
from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
torch.manual_seed(0)
m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
                                   dropout=0.1, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048).cuda().eval()
nat = m.native()
for B in (2, 4, 8, 12, 16, 24, 32, 64):
    x = torch.randn(B * 500, 345, device="cuda")
    for _ in range(3): nat.forward(x, [500] * B, 6)
    nat.set_profiling(True)
    for _ in range(5): nat.forward(x, [500] * B, 6)
    torch.cuda.synchronize()
    prof = nat.get_profile(); nat.set_profiling(False)
    tiles_dec = -(-B * 500 * 6 // 128)
    row = {k: v[0] / v[1] * 1e3 for k, v in prof.items()}
    print(f"B={B:3d} dec_tiles={tiles_dec:5d} waves={tiles_dec/148:5.2f} | ffn {row['dec.ffn_fused']:7.1f} us ({row['dec.ffn_fused']/max(1,-(-tiles_dec//148)):6.1f}/wave) | attn {row['dec.attn_causal']:6.1f} | spk {row['dec.spk_attn']:6.1f} | qkv1 {row['dec.gemm_qkv1']:6.1f} | out1 {row['dec.gemm_out1_ln']:6.1f}")
