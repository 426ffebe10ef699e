"""L2-resident vs DRAM-streaming timing of the causal attention kernel (same shape, 1 vs 6 rotating buffers)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch
from fseend_b200 import native as N
w = torch.randn(8192, 8192, device="cuda").half()
for _ in range(200): w @ w
torch.cuda.synchronize()
def run(B, T, S, nbuf, iters=60):
    qkv = [(torch.randn(B, T, S, 768, device="cuda") * 0.5).half() for _ in range(nbuf)]
    best = 1e9
    for rep in range(3):
        for i in range(10): N.op_causal_attn(qkv[i % nbuf])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters): N.op_causal_attn(qkv[i % nbuf])
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters)
    return best * 1e3
for dbg in sys.argv[1:] or ["0"]:
    os.environ["FSEEND_ATTN_DBG"] = dbg
    for B in (8, 16):
        t1, t6 = run(B, 500, 6, 1), run(B, 500, 6, 12)
        print(f"dbg={dbg} B={B} T=500 S=6: L2-resident {t1:7.1f} us   streaming {t6:7.1f} us   (x{64//B}: {t1*64/B:6.1f} / {t6*64/B:6.1f})", flush=True)
