"""torch.autograd.Functions over the native training kernels (csrc/train_ops.cu, csrc/train_attn.cu).

SURVEY.md §8f "N1" — STARTED, not a training step: forward AND backward of Linear(+ReLU), residual + LayerNorm and
causal multi-head attention run in this library's kernels (fp32 tensors, split-precision tcgen05 products), composed
here into the reference's post-norm encoder layer (``nn.TransformerEncoderLayer`` as constructed at ``FS:model:147``
and looped at ``FS:fusion:129-131``).  Gradients are pinned against torch autograd in tests/test_train_ops_gpu.py.
Residual / FFN dropouts are torch's; the attention-probability dropout is the kernels' own (a counter-based hash mask
regenerated in the backward).  There is no optimizer or DDP wrapper of our own (torch's work unchanged).

No CPU path: every Function raises FseendError on non-CUDA tensors.
"""
from __future__ import annotations

import torch
from torch import nn

from . import native as N
from .native import FseendError, _check, _ptr, _stream, lib

_ws: dict = {}


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """One grow-only scratch buffer per device (kernels of one stream run in order, so it is reused by every call)."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=dev)
        _ws[key] = buf
    return buf


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.float32:
        raise FseendError(f"{what}: CUDA float32 tensor required (fseend_b200 has no CPU fallback)")
    return t.contiguous()


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b); x [..., K], W [N, K] (N % 128 == 0), act in {"none", "relu"}."""

    @staticmethod
    def forward(ctx, x, w, b, act: str = "none"):
        x2 = _f32c(x, "LinearFn x").reshape(-1, x.shape[-1])
        w = _f32c(w, "LinearFn w")
        rows, K = x2.shape
        Nn = w.shape[0]
        if w.shape[1] != K:
            raise FseendError("LinearFn: x / w shape mismatch")
        a = {"none": 0, "relu": 1}[act]
        y = torch.empty(rows, Nn, device=x.device, dtype=torch.float32)
        nb = lib().fseend_train_linear_workspace_bytes(rows, K, Nn)
        ws = _workspace(x.device, nb)
        with torch.cuda.device(x.device):
            _check(lib().fseend_train_linear_fwd(_ptr(x2), rows, K, _ptr(w), Nn, _ptr(None if b is None else _f32c(b, "b")), a,
                                                 _ptr(y), _ptr(ws), ws.numel(), _stream()))
        ctx.save_for_backward(x2, w, y if a == 1 else None)
        ctx.act, ctx.has_bias, ctx.xshape = a, b is not None, x.shape
        return y.view(*x.shape[:-1], Nn)

    @staticmethod
    def backward(ctx, dy):
        x2, w, y = ctx.saved_tensors
        rows, K = x2.shape
        Nn = w.shape[0]
        dy2 = _f32c(dy, "LinearFn dy").reshape(rows, Nn)
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        db = torch.empty(Nn, device=w.device, dtype=torch.float32) if ctx.has_bias else None
        nb = lib().fseend_train_linear_workspace_bytes(rows, K, Nn)
        ws = _workspace(w.device, nb)
        with torch.cuda.device(w.device):
            _check(lib().fseend_train_linear_bwd(_ptr(x2), _ptr(w), _ptr(y), _ptr(dy2), rows, K, Nn, ctx.act, 0, _ptr(dx),
                                                 _ptr(dw), _ptr(db), _ptr(ws), ws.numel(), _stream()))
        return (None if dx is None else dx.view(ctx.xshape)), dw, db, None


def _linear_fwd(x2, w, b, act):
    rows, K = x2.shape
    Nn = w.shape[0]
    y = torch.empty(rows, Nn, device=x2.device, dtype=torch.float32)
    ws = _workspace(x2.device, lib().fseend_train_linear_workspace_bytes(rows, K, Nn))
    _check(lib().fseend_train_linear_fwd(_ptr(x2), rows, K, _ptr(w), Nn, _ptr(b), act, _ptr(y), _ptr(ws), ws.numel(), _stream()))
    return y


def _linear_bwd(x2, w, dy2, relu_input, want_dx, want_db):
    rows, K = x2.shape
    Nn = w.shape[0]
    dx = torch.empty_like(x2) if want_dx else None
    dw = torch.empty_like(w)
    db = torch.empty(Nn, device=w.device, dtype=torch.float32) if want_db else None
    ws = _workspace(w.device, lib().fseend_train_linear_workspace_bytes(rows, K, Nn))
    _check(lib().fseend_train_linear_bwd(_ptr(x2), _ptr(w), None, _ptr(dy2), rows, K, Nn, 0, 1 if relu_input else 0, _ptr(dx),
                                         _ptr(dw), _ptr(db), _ptr(ws), ws.numel(), _stream()))
    return dx, dw, db


class FfnFn(torch.autograd.Function):
    """y = relu(x W1^T + b1) W2^T + b2 (``_ff_block``, FS:fusion:397-399 / nn.TransformerEncoderLayer).  One Function so
    that the ReLU's backward is folded into the epilogue of the down-projection's input-gradient product instead of a
    separate pass over the [rows, 2048] gradient."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        x2 = _f32c(x, "FfnFn x").reshape(-1, x.shape[-1])
        w1, b1, w2, b2 = (_f32c(t, "FfnFn parameter") for t in (w1, b1, w2, b2))
        if w1.shape[0] % 128 or w2.shape[0] % 128 or w2.shape[1] != w1.shape[0] or w1.shape[1] != x2.shape[1]:
            raise FseendError("FfnFn: shapes must be x [.., K], W1 [F, K], W2 [N, F] with F, N multiples of 128")
        with torch.cuda.device(x.device):
            h = _linear_fwd(x2, w1, b1, 1)
            y = _linear_fwd(h, w2, b2, 0)
        ctx.save_for_backward(x2, w1, w2, h)
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w1, w2, h = ctx.saved_tensors
        dy2 = _f32c(dy, "FfnFn dy").reshape(h.shape[0], w2.shape[0])
        with torch.cuda.device(x2.device):
            dh, dw2, db2 = _linear_bwd(h, w2, dy2, True, True, True)           # dh already carries the ReLU mask
            dx, dw1, db1 = _linear_bwd(x2, w1, dh, False, ctx.needs_input_grad[0], True)
        return (None if dx is None else dx.view(ctx.xshape)), dw1, db1, dw2, db2


class AddLayerNormFn(torch.autograd.Function):
    """y = LayerNorm(x + r) over the last dim (256); r may be None.  Gradient flows equally into x and r."""

    @staticmethod
    def forward(ctx, x, r, g, b, eps: float = 1e-5):
        x2 = _f32c(x, "AddLayerNormFn x").reshape(-1, 256)
        r2 = None if r is None else _f32c(r, "AddLayerNormFn r").reshape(-1, 256)
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        s = torch.empty_like(x2) if r2 is not None else None
        with torch.cuda.device(x.device):
            _check(lib().fseend_train_add_layernorm_fwd(_ptr(x2), _ptr(r2), _ptr(_f32c(g, "g")), _ptr(_f32c(b, "b")), rows,
                                                        float(eps), _ptr(s), _ptr(y), _stream()))
        ctx.save_for_backward(x2 if s is None else s, g)
        ctx.eps, ctx.has_r, ctx.shape = float(eps), r is not None, x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        s, g = ctx.saved_tensors
        rows = s.shape[0]
        dy2 = _f32c(dy, "AddLayerNormFn dy").reshape(rows, 256)
        dx = torch.empty_like(s)
        dg = torch.empty(256, device=s.device, dtype=torch.float32)
        db = torch.empty(256, device=s.device, dtype=torch.float32)
        nb = lib().fseend_train_layernorm_workspace_bytes(rows)
        ws = _workspace(s.device, nb)
        with torch.cuda.device(s.device):
            _check(lib().fseend_train_layernorm_bwd(_ptr(s), _ptr(g.contiguous()), _ptr(dy2), rows, ctx.eps, _ptr(dx), _ptr(dg),
                                                    _ptr(db), _ptr(ws), ws.numel(), _stream()))
        dxv = dx.view(ctx.shape)
        return dxv, (dxv if ctx.has_r else None), dg, db, None


class CausalAttnFn(torch.autograd.Function):
    """Causal 4-head attention on projected qkv (q | k | v); key j visible to query i iff j <= i + delay.

    qkv [n_seq, T, 768] -> [n_seq, T, 256], or the attractor decoder's interleaved layout qkv [B, T, S, 768] ->
    [B, T, S, 256], where every (b, s) is a sequence over T read with row stride S (no transposed copy)."""

    @staticmethod
    def forward(ctx, qkv, mask_delay: int = 0, dropout_p: float = 0.0, seed: int = 0):
        qkv = _f32c(qkv, "CausalAttnFn qkv")
        if qkv.dim() not in (3, 4) or qkv.shape[-1] != 768:
            raise FseendError("CausalAttnFn: qkv must be [n_seq, T, 768] or [B, T, S, 768]")
        inner = qkv.shape[2] if qkv.dim() == 4 else 1
        n, T = qkv.shape[0] * inner, qkv.shape[1]
        out = torch.empty(*qkv.shape[:-1], 256, device=qkv.device, dtype=torch.float32)
        lse = torch.empty(n, 4, T, device=qkv.device, dtype=torch.float32)
        with torch.cuda.device(qkv.device):
            _check(lib().fseend_train_attn_fwd(_ptr(qkv), n, T, inner, int(mask_delay), float(dropout_p), int(seed), _ptr(out),
                                               _ptr(lse), _stream()))
        ctx.save_for_backward(qkv, out, lse)
        ctx.delay, ctx.p, ctx.seed, ctx.inner = int(mask_delay), float(dropout_p), int(seed), inner
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        n, _, T = lse.shape
        dout = _f32c(dout, "CausalAttnFn dout")
        dqkv = torch.empty_like(qkv)
        dsum = torch.empty(lse.numel() + 16, device=lse.device, dtype=torch.float32)
        with torch.cuda.device(qkv.device):
            _check(lib().fseend_train_attn_bwd(_ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), n, T, ctx.inner, ctx.delay, ctx.p,
                                               ctx.seed, _ptr(dqkv), _ptr(dsum), _stream()))
        return dqkv, None, None, None


class SpeakerAttnFn(torch.autograd.Function):
    """qkv [n_frames, S, 768] -> [n_frames, S, 256]: unmasked 4-head attention over the speaker axis (FS:fusion:390)."""

    @staticmethod
    def forward(ctx, qkv, dropout_p: float = 0.0, seed: int = 0):
        qkv = _f32c(qkv, "SpeakerAttnFn qkv")
        if qkv.dim() != 3 or qkv.shape[-1] != 768 or qkv.shape[1] > 16:
            raise FseendError("SpeakerAttnFn: qkv must be [n_frames, S <= 16, 768]")
        n, S, _ = qkv.shape
        out = torch.empty(n, S, 256, device=qkv.device, dtype=torch.float32)
        with torch.cuda.device(qkv.device):
            _check(lib().fseend_train_spk_attn_fwd(_ptr(qkv), n, S, float(dropout_p), int(seed), _ptr(out), _stream()))
        ctx.save_for_backward(qkv)
        ctx.p, ctx.seed = float(dropout_p), int(seed)
        return out

    @staticmethod
    def backward(ctx, dout):
        (qkv,) = ctx.saved_tensors
        n, S, _ = qkv.shape
        dout = _f32c(dout, "SpeakerAttnFn dout")
        dqkv = torch.empty_like(qkv)
        with torch.cuda.device(qkv.device):
            _check(lib().fseend_train_spk_attn_bwd(_ptr(qkv), _ptr(dout), n, S, ctx.p, ctx.seed, _ptr(dqkv), _stream()))
        return dqkv, None, None


class L2NormFn(torch.autograd.Function):
    """y = x / ||x||_2 over the last dim (256), no eps (FS:model:41,43)."""

    @staticmethod
    def forward(ctx, x):
        x2 = _f32c(x, "L2NormFn x").reshape(-1, 256)
        y = torch.empty_like(x2)
        inv = torch.empty(x2.shape[0], device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _check(lib().fseend_train_l2norm_fwd(_ptr(x2), x2.shape[0], _ptr(y), _ptr(inv), _stream()))
        ctx.save_for_backward(y, inv)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        dy2 = _f32c(dy, "L2NormFn dy").reshape(-1, 256)
        dx = torch.empty_like(y)
        with torch.cuda.device(y.device):
            _check(lib().fseend_train_l2norm_bwd(_ptr(y), _ptr(inv), _ptr(dy2), y.shape[0], _ptr(dx), _stream()))
        return dx.view(dy.shape)


class HeadFn(torch.autograd.Function):
    """logits[b, t, s] = emb[b, t, :] . att[b, t, s, :]  (FS:model:60)."""

    @staticmethod
    def forward(ctx, emb, att):
        emb, att = _f32c(emb, "HeadFn emb"), _f32c(att, "HeadFn att")
        B, T, S, D = att.shape
        if D != 256 or tuple(emb.shape) != (B, T, D):
            raise FseendError("HeadFn: emb [B, T, 256], att [B, T, S, 256]")
        y = torch.empty(B, T, S, device=emb.device, dtype=torch.float32)
        with torch.cuda.device(emb.device):
            _check(lib().fseend_train_head_fwd(_ptr(emb), _ptr(att), B * T, S, _ptr(y), _stream()))
        ctx.save_for_backward(emb, att)
        return y

    @staticmethod
    def backward(ctx, dy):
        emb, att = ctx.saved_tensors
        B, T, S, _ = att.shape
        dy = _f32c(dy, "HeadFn dy")
        demb, datt = torch.empty_like(emb), torch.empty_like(att)
        with torch.cuda.device(emb.device):
            _check(lib().fseend_train_head_bwd(_ptr(emb), _ptr(att), _ptr(dy), B * T, S, _ptr(demb), _ptr(datt), _stream()))
        return demb, datt


class BatchNormTrainFn(torch.autograd.Function):
    """nn.BatchNorm1d in training mode over the rows of x [..., C]: returns (y, batch mean, biased batch variance)."""

    @staticmethod
    def forward(ctx, x, g, b, eps: float):
        Cn = x.shape[-1]
        x2 = _f32c(x, "BatchNormTrainFn x").reshape(-1, Cn)
        g, b = _f32c(g, "gamma"), _f32c(b, "beta")
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        stats = torch.empty(2 * Cn, device=x.device, dtype=torch.float32)
        ws = _workspace(x.device, lib().fseend_train_batchnorm_workspace_bytes(rows, Cn))
        with torch.cuda.device(x.device):
            _check(lib().fseend_train_batchnorm_fwd(_ptr(x2), _ptr(g), _ptr(b), rows, Cn, float(eps), _ptr(y), _ptr(stats),
                                                    _ptr(ws), ws.numel(), _stream()))
        ctx.save_for_backward(x2, g, stats)
        ctx.eps, ctx.shape = float(eps), x.shape
        ctx.mark_non_differentiable(stats)
        return y.view(x.shape), stats

    @staticmethod
    def backward(ctx, dy, _dstats):
        x2, g, stats = ctx.saved_tensors
        rows, Cn = x2.shape
        dy2 = _f32c(dy, "BatchNormTrainFn dy").reshape(rows, Cn)
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dg, db = torch.empty_like(g), torch.empty_like(g)
        ws = _workspace(x2.device, lib().fseend_train_batchnorm_workspace_bytes(rows, Cn))
        with torch.cuda.device(x2.device):
            _check(lib().fseend_train_batchnorm_bwd(_ptr(x2), _ptr(g), _ptr(stats), _ptr(dy2), rows, Cn, ctx.eps, _ptr(dx),
                                                    _ptr(dg), _ptr(db), _ptr(ws), ws.numel(), _stream()))
        return (None if dx is None else dx.view(ctx.shape)), dg, db, None


def batch_norm_forward(bn: nn.BatchNorm1d, x: torch.Tensor) -> torch.Tensor:
    """``bn(x.transpose(1, 2)).transpose(1, 2)`` for x [B, T, C] (FS:model:166).  Training mode: native kernels with batch
    statistics; the running statistics are updated as nn.BatchNorm1d does (momentum, unbiased variance,
    num_batches_tracked).  Eval mode inside a training graph (frozen statistics): torch's functional batch_norm."""
    if not bn.training:
        return torch.nn.functional.batch_norm(x.transpose(1, 2), bn.running_mean, bn.running_var, bn.weight, bn.bias, False,
                                              0.0, bn.eps).transpose(1, 2).contiguous()
    y, stats = BatchNormTrainFn.apply(x, bn.weight, bn.bias, bn.eps)
    if bn.track_running_stats:
        with torch.no_grad():
            Cn = x.shape[-1]
            rows = x.numel() // Cn
            bn.num_batches_tracked += 1
            mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            bn.running_mean.mul_(1 - mom).add_(stats[:Cn], alpha=mom)
            bn.running_var.mul_(1 - mom).add_(stats[Cn:] * (rows / max(rows - 1, 1)), alpha=mom)
    return y


def _seed() -> int:
    """A fresh 62-bit seed for one attention call, drawn from torch's CPU generator (reproducible under manual_seed)."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def _drop(x: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    return torch.nn.functional.dropout(x, p, True) if training and p > 0 else x


def _mha(sa: nn.MultiheadAttention, x: torch.Tensor, attend) -> torch.Tensor:
    """attend(qkv, p, seed) -> context; p = the module's attention-probability dropout when it is in train mode."""
    qkv = LinearFn.apply(x, sa.in_proj_weight, sa.in_proj_bias, "none")
    p = float(sa.dropout) if sa.training else 0.0
    ctx = attend(qkv, p, _seed() if p > 0 else 0)
    return LinearFn.apply(ctx, sa.out_proj.weight, sa.out_proj.bias, "none")


def _ffn(layer, x: torch.Tensor) -> torch.Tensor:
    p = float(layer.dropout.p) if layer.training else 0.0
    if p > 0:          # dropout sits between the two projections: the fused pair does not apply
        h = _drop(LinearFn.apply(x, layer.linear1.weight, layer.linear1.bias, "relu"), p, True)
        return LinearFn.apply(h, layer.linear2.weight, layer.linear2.bias, "none")
    return FfnFn.apply(x, layer.linear1.weight, layer.linear1.bias, layer.linear2.weight, layer.linear2.bias)


def fusion_layer_forward(layer, x: torch.Tensor, mask_delay: int = 0) -> torch.Tensor:
    """The reference's attractor-decoder layer, live path ``FS:fusion:356-376`` (post-norm), differentiable.

    x: [B, T, S, 256].  Time attention runs per (b, s) over T (causal), speaker attention per (b, t) over S (no mask);
    ``norm12`` is unused, as in the reference.  There are no layout copies (the reference transposes twice per layer,
    ``:358-365``); in train mode the residual
    dropouts (dropout11 / dropout21 / dropout2, ``:380-399``) are torch's, the attention-probability dropout is the
    kernels' own (hash mask)."""
    B, T, S, D = x.shape
    tr = layer.training
    # every op but the two attention cores is row-wise, so the tensor stays [B, T, S, D] throughout: the time attention
    # reads sequence (b, s) in place with row stride S, the speaker attention sees [B * T, S, .] as a view
    a = _mha(layer.self_attn1, x, lambda qkv, p, seed: CausalAttnFn.apply(qkv, mask_delay, p, seed))
    y = AddLayerNormFn.apply(_drop(a, layer.dropout11.p, tr), x, layer.norm11.weight, layer.norm11.bias, layer.norm11.eps)
    y = y.reshape(B * T, S, D)
    a = _mha(layer.self_attn2, y, lambda qkv, p, seed: SpeakerAttnFn.apply(qkv, p, seed))
    y = AddLayerNormFn.apply(_drop(a, layer.dropout21.p, tr), y, layer.norm21.weight, layer.norm21.bias, layer.norm21.eps)
    y = AddLayerNormFn.apply(_drop(_ffn(layer, y), layer.dropout2.p, tr), y, layer.norm22.weight, layer.norm22.bias,
                             layer.norm22.eps)
    return y.reshape(B, T, S, D)


def encoder_layer_forward(layer: nn.TransformerEncoderLayer, x: torch.Tensor, mask_delay: int = 0) -> torch.Tensor:
    """The reference's post-norm encoder layer (``FS:model:147``) on native kernels, differentiable.

    x: [n_seq, T, 256] (batch-first; the reference runs (T, B, 256) — the math is per sequence, the layout is ours).
    ``layer`` only provides the parameters (reference names: self_attn.in_proj_*, self_attn.out_proj, linear1/2, norm1/2)
    and the dropout probabilities (dropout1 / dropout / dropout2 and the attention's own)."""
    if layer.norm_first:
        raise FseendError("encoder_layer_forward: the reference layer is post-norm")
    tr = layer.training
    o = _mha(layer.self_attn, x, lambda qkv, p, seed: CausalAttnFn.apply(qkv, mask_delay, p, seed))
    x = AddLayerNormFn.apply(_drop(o, layer.dropout1.p, tr), x, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps)
    f = _drop(_ffn(layer, x), layer.dropout2.p, tr)
    return AddLayerNormFn.apply(f, x, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps)
