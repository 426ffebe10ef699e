"""One FS-EEND training step (forward + standard_loss + emb loss + backward + Adam) at BASELINE config 1's shape
(B=64 chunks of 500 frames, 345-d features, 4 enc + 2 dec layers) through the drop-in model — native forward/backward
kernels — against the same graph in plain torch eager ops on the same GPU (float32 with TF32 off, and TF32 on).
The torch arm is the stand-in graph of tests/test_train_graph_cpu.py (F.linear / F.layer_norm / explicit softmax
attention), i.e. what the reference's nn.TransformerEncoderLayer stack executes; the reference package itself is not on
the GPU box.  Prints one JSON line.  SURVEY §8f N1 is started: this is a secondary measurement, not the headline."""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fs-eend_b200")]


def _attn(qkv, delay):
    n, T, _ = qkv.shape
    q, k, v = (t.reshape(n, T, 4, 64).transpose(1, 2) for t in qkv.split(256, dim=-1))
    if delay >= T:
        o = F.scaled_dot_product_attention(q, k, v)
    elif delay == 0:
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
    else:
        i = torch.arange(T, device=qkv.device)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=i[None, :] <= i[:, None] + delay)
    return o.transpose(1, 2).reshape(n, T, 256)



def _causal4(qkv, delay):
    """[n, T, 768] or the decoder's interleaved [B, T, S, 768] (sequence (b, s) over T)."""
    if qkv.dim() == 3:
        return _attn(qkv, delay)
    B, T, S, _ = qkv.shape
    o = _attn(qkv.transpose(1, 2).reshape(B * S, T, 768), delay)
    return o.reshape(B, S, T, 256).transpose(1, 2)


class _Lin:
    @staticmethod
    def apply(x, w, b, act):
        y = F.linear(x, w, b)
        return torch.relu(y) if act == "relu" else y


class _AddLn:
    @staticmethod
    def apply(x, r, g, b, eps):
        return F.layer_norm(x if r is None else x + r, (256,), g, b, eps)


class _Ffn:
    @staticmethod
    def apply(x, w1, b1, w2, b2):
        return F.linear(torch.relu(F.linear(x, w1, b1)), w2, b2)


class _L2:
    apply = staticmethod(lambda x: x / torch.norm(x, dim=-1, keepdim=True))


class _Head:
    apply = staticmethod(lambda emb, att: (emb[:, :, None, :] * att).sum(-1))


def _bn(bn, x):
    return bn(x.transpose(1, 2)).transpose(1, 2).contiguous()


class _Causal:
    apply = staticmethod(lambda qkv, delay, p=0.0, seed=0: _causal4(qkv, delay))


class _Spk:
    apply = staticmethod(lambda qkv, p=0.0, seed=0: _attn(qkv, 1 << 20))


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


def measure(batch=64, frames=500, speakers=4, steps=5, warmup=2, native_only=False):
    """Returns the result dict (None with native_only: one native step for profilers)."""
    import fseend_b200.autograd as A
    import fseend_b200.train_graph as G
    from fseend_b200.loss import standard_loss
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    torch.manual_seed(0)
    m = OnlineTransformerDADiarization(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
                                       dropout=0.0, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048).cuda().train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-5)
    B, T, S = batch, frames, speakers
    src = [torch.randn(T, 345, device="cuda") for _ in range(B)]
    tgt = [(torch.rand(T, S, device="cuda") < 0.3).float() for _ in range(B)]
    # the labels the training step hands to model / loss carry silence + "no speaker" columns: S + 2 classes
    lab = [torch.cat([1 - t.max(-1, keepdim=True)[0], t, torch.zeros(T, 1, device="cuda")], -1) for t in tgt]
    lens = [T] * B
    last = {}

    def step():
        opt.zero_grad(set_to_none=True)
        out, emb_loss, _, _ = m(src, lab, lens)
        loss = standard_loss(out, lab, label_delay=0) + emb_loss
        loss.backward()
        opt.step()
        last["loss"] = loss

    if native_only:
        step()
        torch.cuda.synchronize()
        return None
    tf32_was = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    state0 = {k: v.clone() for k, v in m.state_dict().items()}
    torch.cuda.reset_peak_memory_stats()
    t_native = timed(step, steps, warmup)
    loss_native = last["loss"].item()
    peak_native = torch.cuda.max_memory_allocated() / 2**30
    # torch eager arm: same graph, stand-in ops
    saved = {(mod, n): getattr(mod, n) for mod in (A, G) for n in ("LinearFn", "AddLayerNormFn")}
    saved[(A, "CausalAttnFn")], saved[(A, "SpeakerAttnFn")], saved[(A, "FfnFn")] = A.CausalAttnFn, A.SpeakerAttnFn, A.FfnFn
    for n_ in ("L2NormFn", "HeadFn", "batch_norm_forward"):
        saved[(G, n_)] = getattr(G, n_)
    for mod in (A, G):
        mod.LinearFn, mod.AddLayerNormFn = _Lin, _AddLn
    A.CausalAttnFn, A.SpeakerAttnFn, A.FfnFn = _Causal, _Spk, _Ffn
    G.L2NormFn, G.HeadFn, G.batch_norm_forward = _L2, _Head, _bn
    res = {}
    try:
        for name, tf32 in (("torch_fp32", False), ("torch_tf32", True)):
            m.load_state_dict(state0)
            opt.state.clear()
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            res[name] = timed(step, steps, warmup)
            res[name + "_loss"] = last["loss"].item()
    finally:
        for (mod, n), v in saved.items():
            setattr(mod, n, v)
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32_was
    n_frames = B * T
    return {"workload": f"FS-EEND training step: fwd + standard_loss + emb loss + bwd + Adam, B={B} x T={T}, {S + 2} label classes, "
                        "4 enc + 2 dec layers, dropout 0",
            "status": "SURVEY 8f N1 STARTED: GEMMs / attention / LayerNorm / BatchNorm / L2 norms / head forward+backward are this "
                      "library's kernels; emb loss, layout copies, residual dropout and the optimizer are torch ops",
            "native_ms": round(t_native, 2), "native_frames_per_s": round(n_frames / t_native * 1e3),
            "torch_eager_fp32_ms": round(res["torch_fp32"], 2), "torch_eager_tf32_ms": round(res["torch_tf32"], 2),
            "loss_after_%d_steps" % (steps + warmup): {"native": loss_native, "torch_fp32": res["torch_fp32_loss"],
                                                        "torch_tf32": res["torch_tf32_loss"]},
            "native_peak_mem_gib": round(peak_native, 2),
            "torch_arm": "same graph with F.linear / F.layer_norm / SDPA (what nn.TransformerEncoderLayer executes), same GPU"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--speakers", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--native-only", action="store_true")
    a = ap.parse_args()
    r = measure(a.batch, a.frames, a.speakers, a.steps, a.warmup, a.native_only)
    if r is not None:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
