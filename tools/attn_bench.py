"""Micro-benchmark of the causal attention kernel on the decoder / encoder shapes (CUDA events, L2 flushed by rotating
buffers).  usage: python tools/attn_bench.py  (env switches are read per launch)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]
import torch
from fseend_b200 import native as N

_bufs = {}
def bench(B, T, S, iters=50, nbuf=3, reps=3):
    if (B, T, S) not in _bufs:
        _bufs[(B, T, S)] = [(torch.randn(B, T, S, 768, device="cuda") * 0.5).half() for _ in range(nbuf)]
    qkv = _bufs[(B, T, S)]
    for i in range(30): N.op_causal_attn(qkv[i % nbuf])
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters): N.op_causal_attn(qkv[i % nbuf])
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters)
    ms = best
    fl = 4.0 * 64 * T * (T + 1) / 2 * 4 * B * S
    return ms, fl / ms / 1e9

w = torch.randn(8192, 8192, device="cuda").half()
for _ in range(200): w @ w          # ramp the clocks
torch.cuda.synchronize()
for tag, env in [("default", {}), ("order0", {"FSEEND_ATTN_ORDER": "0"})] + \
        [(a, dict(kv.split("=") for kv in a.split(","))) for a in sys.argv[1:]]:
    for k, v in env.items(): os.environ[k] = v
    for shape in ([(64, 500, 6)] if len(sys.argv) > 1 else [(64, 500, 6), (64, 500, 1), (16, 2000, 1)]):
        ms, tf = bench(*shape)
        print(f"{tag:30s} B,T,S={shape}: {ms*1e3:8.1f} us  {tf:7.1f} TF/s causal-exact", flush=True)
    for k in env: os.environ.pop(k)
