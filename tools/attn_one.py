"""One decoder-shaped causal attention launch (ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch
from fseend_b200 import native as N
B, T, S = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 500, 6)))
qkv = (torch.randn(B, T, S, 768, device="cuda") * 0.5).half()
for _ in range(3):
    o = N.op_causal_attn(qkv)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
