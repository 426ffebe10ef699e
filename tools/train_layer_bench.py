"""Forward + backward of the FS-EEND encoder stack (4 post-norm layers, B x T x 256) on the native training kernels
(fseend_b200.autograd) against torch eager on the same GPU: fp32 (TF32 off — the arithmetic class the native path
matches) and TF32.  Prints one JSON line.  Not part of bench.py's headline: SURVEY §8f N1 is started, not complete."""
import argparse
import json
import os
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--native-only", action="store_true", help="one native forward+backward and exit (for ncu launch lists)")
    a = ap.parse_args()
    from fseend_b200.autograd import encoder_layer_forward
    torch.manual_seed(0)
    layers = nn.ModuleList(nn.TransformerEncoderLayer(256, 4, 2048, 0.0, batch_first=True) for _ in range(a.layers)).cuda()
    layers.train()
    x = torch.randn(a.batch, a.frames, 256, device="cuda", requires_grad=True)
    dy = torch.randn(a.batch, a.frames, 256, device="cuda")
    i = torch.arange(a.frames, device="cuda")
    mask = torch.zeros(a.frames, a.frames, device="cuda").masked_fill(i[None, :] > i[:, None], float("-inf"))

    def native():
        h = x
        for l in layers:
            h = encoder_layer_forward(l, h, 0)
        h.backward(dy)

    def eager():
        h = x
        for l in layers:
            h = l(h, src_mask=mask)
        h.backward(dy)

    def grads():
        return [p.grad.clone() for p in layers.parameters()] + [x.grad.clone()]

    def zero():
        for p in list(layers.parameters()) + [x]:
            p.grad = None

    if a.native_only:
        native()
        torch.cuda.synchronize()
        return
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def worst(ga, gb):
        return max(((a_.double() - b_.double()).abs().max() / b_.double().abs().max().clamp_min(1e-30)).item()
                   for a_, b_ in zip(ga, gb))

    zero(); native(); g_native = grads()
    zero(); eager(); g_fp32 = grads()
    # float64 reference (torch math path) for the accuracy columns
    import copy
    layers64 = copy.deepcopy(layers).double()
    x64 = x.detach().double().requires_grad_()
    h = x64
    for l in layers64:
        h = l(h, src_mask=mask.double())
    h.backward(dy.double())
    g_fp64 = [p.grad for p in layers64.parameters()] + [x64.grad]
    rel = worst(g_native, g_fp64)
    rel_fp32 = worst(g_fp32, g_fp64)
    del layers64, x64, h
    t_native = timed(lambda: (zero(), native()), a.steps, a.warmup)
    t_fp32 = timed(lambda: (zero(), eager()), a.steps, a.warmup)
    torch.backends.cuda.matmul.allow_tf32 = True
    zero(); eager(); g_tf32 = grads()
    rel_tf32 = worst(g_tf32, g_fp64)
    t_tf32 = timed(lambda: (zero(), eager()), a.steps, a.warmup)
    frames = a.batch * a.frames
    print(json.dumps({"what": "encoder stack fwd+bwd", "batch": a.batch, "frames": a.frames, "layers": a.layers,
                      "native_ms": round(t_native, 3), "torch_fp32_ms": round(t_fp32, 3), "torch_tf32_ms": round(t_tf32, 3),
                      "native_frames_per_s": round(frames / t_native * 1e3), "max_rel_grad_err_vs_fp64": {"native": rel, "torch_fp32": rel_fp32,
                                                                               "torch_tf32": rel_tf32}}))


if __name__ == "__main__":
    main()
