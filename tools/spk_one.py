"""One fused speaker-attention launch at the bench shape (ncu target) + timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch
from fseend_b200 import native as N
F, S = 32000, 6
x = (torch.randn(F, S, 256, device="cuda")).half()
w = (torch.randn(768, 256, device="cuda") / 16).half()
b = torch.randn(768, device="cuda") * 0.1
for _ in range(3): o = N.op_spk_qkv_attn(x, w, b)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): o = N.op_spk_qkv_attn(x, w, b)
e.record(); torch.cuda.synchronize()
print("spk_fused us", a.elapsed_time(e) / 20 * 1e3)
