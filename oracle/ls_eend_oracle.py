"""CPU oracle for the LS-EEND hot path (TEST INFRASTRUCTURE — never imported by the product).

Functional restatement (plain torch CPU arithmetic) of the reference's Conformer-retention encoder +
retention attractor decoder, batch (chunkwise) and one-step (recurrent) forms.  File:line citations are relative
to /root/reference/LS-EEND/nnet:
  model/onl_conformer_retention_enc_1dcnn_tfm_retention_enc_linear_non_autoreg_pos_enc_l2norm_emb_loss_mask.py = LS:model
  modules/retention.py = LS:ret      modules/merge_retnet_layer.py = LS:fusion     conformer/*.py = LS:conf/*

Pinning: tests/golden/make_golden_ls.py imports the REAL reference (separate process: both reference trees use
the top-level package name ``nnet``) and stores its outputs; tests/test_oracle_ls.py checks this restatement
against them.  Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module.

The ``state_dict`` uses the reference's key names.  ``quant`` rounds every GEMM operand
(identity by default) to emulate tensor-core operand precision.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

from .fs_eend_oracle import _ident, layer_norm, linear, mha, pad_sequence, pe_table

Tensor = torch.Tensor
SD = Dict[str, Tensor]


class Cfg:
    """Constructor kwargs of OnlineConformerRetentionDADiarization (LS:model:15-32) with the values of
    LS-EEND/conf/spk_onl_conformer_retention_enc_dec_nonautoreg.yaml:27-40."""

    def __init__(self, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2, chunk=500, ff_expansion=4,
                 dec_ff=2048, conv_kernel=16, conv_delay=9):
        self.n_units, self.n_heads = n_units, n_heads
        self.enc_n_layers, self.dec_n_layers = enc_n_layers, dec_n_layers
        self.chunk, self.ff_expansion, self.dec_ff = chunk, ff_expansion, dec_ff
        self.conv_kernel, self.conv_delay = conv_kernel, conv_delay


def swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)                                   # LS:conf/activation.py:28-29


# ----------------------------------------------------------------------------- retention (a11)

def _group_norm(x: Tensor) -> Tensor:
    """LayerNorm(head_dim, eps=1e-6, elementwise_affine=False) over the last dim.  LS:ret:100."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-6)


def retention_chunkwise(x: Tensor, sd: SD, p: str, H: int, C: int, quant=_ident) -> Tensor:
    """MultiScaleRetention.forward(chunkwise_recurrent=True) with decay = log 1 and no rotation.
    LS:ret:196-228 + chunk_recurrent_forward :146-194 + RetNetRelPos.forward :29-47.   x: (N, T, D), T % C == 0."""
    N, T, D = x.shape
    hd = D // H
    q = linear(x, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"], quant)
    k = linear(x, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"], quant) * hd ** -0.5     # LS:ret:205
    v = linear(x, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"], quant)
    g = linear(x, sd[p + "g_proj.weight"], sd[p + "g_proj.bias"], quant)
    assert T % C == 0                                                                      # LS:ret:152
    nc = T // C
    qr = q.reshape(N, T, H, hd).transpose(1, 2).reshape(N, H, nc, C, hd).transpose(1, 2)   # (N, nc, H, C, hd)
    kr = k.reshape(N, T, H, hd).transpose(1, 2).reshape(N, H, nc, C, hd).transpose(1, 2)
    vr = v.reshape(N, nc, C, H, hd).transpose(2, 3)                                        # (N, nc, H, C, hd)
    # RetNetRelPos (chunkwise): mask[j][i] = 1/sqrt(j+1) for i <= j (decay = 0)              LS:ret:34-41
    idx = torch.arange(C, dtype=x.dtype)
    mask = torch.tril(torch.ones(C, C, dtype=x.dtype))
    scale = mask.sum(dim=-1, keepdim=True).sqrt()
    mask = mask / scale
    inner_decay = (1.0 / (scale / scale[-1]))                                              # (C,1)  LS:ret:44-46
    qk = (quant(qr) @ quant(kr).transpose(-1, -2)) * mask                                  # LS:ret:160-161
    inner_scale = qk.abs().sum(dim=-1, keepdim=True).clamp(min=1)                          # :162
    qk = qk / inner_scale
    inner_output = quant(qk) @ quant(vr)                                                   # :164
    kv = quant(kr).transpose(-1, -2) @ quant(vr * mask[-1, :, None])                       # :167  (N,nc,H,hd,hd)
    kv_state = x.new_zeros(N, H, hd, hd)
    kv_scale = x.new_ones(N, H, 1, 1)
    kv_rec, cross_scale = [], []
    for i in range(nc):                                                                    # :176-180
        kv_rec.append(kv_state / kv_scale)
        cross_scale.append(kv_scale)
        kv_state = kv_state + kv[:, i]
        kv_scale = kv_state.abs().sum(dim=-2, keepdim=True).max(dim=-1, keepdim=True).values.clamp(min=1)
    kv_rec = torch.stack(kv_rec, dim=1)
    cross_scale = torch.stack(cross_scale, dim=1)
    all_scale = torch.maximum(inner_scale, cross_scale)                                    # :185
    cross_output = quant(qr * inner_decay) @ quant(kv_rec)                                 # :189
    out = inner_output / (all_scale / inner_scale) + cross_output / (all_scale / cross_scale)   # :190
    out = out.transpose(2, 3)                                                              # (N, nc, C, H, hd)
    out = _group_norm(out).reshape(N, T, D)                                                # :222
    out = swish(g) * out                                                                   # :224
    return linear(out, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"], quant)


class RetState:
    """incremental_state of recurrent_forward (LS:ret:126-144): prev_key_value (N,H,hd,hd) and scale."""

    def __init__(self):
        self.kv: Optional[Tensor] = None
        self.scale: float = 1.0


def retention_step(x_t: Tensor, sd: SD, p: str, H: int, st: RetState, quant=_ident) -> Tensor:
    """MultiScaleRetention.forward(incremental_state=...) for one frame.  x_t: (N, 1, D).  LS:ret:126-144."""
    N, _, D = x_t.shape
    hd = D // H
    q = linear(x_t, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"], quant).reshape(N, H, hd)
    k = (linear(x_t, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"], quant) * hd ** -0.5).reshape(N, H, hd)
    v = linear(x_t, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"], quant).reshape(N, H, hd)
    g = linear(x_t, sd[p + "g_proj.weight"], sd[p + "g_proj.bias"], quant)
    kv = k[:, :, :, None] * v[:, :, None, :]
    if st.kv is not None:
        scale = st.scale + 1.0                                       # prev_scale * decay(=1) + 1
        kv = st.kv * (math.sqrt(st.scale) / math.sqrt(scale)) + kv / math.sqrt(scale)
    else:
        scale = 1.0
    st.kv, st.scale = kv, scale
    out = (quant(q)[:, :, :, None] * quant(kv)).sum(dim=2)           # (N, H, hd)   LS:ret:143
    out = _group_norm(out).reshape(N, 1, D)
    out = swish(g) * out
    return linear(out, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"], quant)


# ----------------------------------------------------------------------------- conformer block (a10)

def ffn_module(x: Tensor, sd: SD, p: str, quant=_ident) -> Tensor:
    """FeedForwardModule: LN -> Linear -> swish -> Linear.  LS:conf/feed_forward.py:47-54."""
    h = layer_norm(x, sd[p + "sequential.0.weight"], sd[p + "sequential.0.bias"])
    h = swish(linear(h, sd[p + "sequential.1.linear.weight"], sd[p + "sequential.1.linear.bias"], quant))
    return linear(h, sd[p + "sequential.4.linear.weight"], sd[p + "sequential.4.linear.bias"], quant)


def conv_module(x: Tensor, sd: SD, p: str, cache: Optional[Tensor] = None, quant=_ident):
    """ConformerConvModule (LS:conf/convolution.py:138-149; one-step :154-167).  x: (B, T, D).
    cache: (B, K-1, D) of previous GLU outputs (one-step) or None (batch, causal left zero padding).
    Returns (out (B,T,D), new_cache)."""
    h = layer_norm(x, sd[p + "sequential.0.weight"], sd[p + "sequential.0.bias"])
    w1 = sd[p + "sequential.2.conv.weight"][:, :, 0]                                    # (2D, D)
    h = linear(h, w1, sd[p + "sequential.2.conv.bias"], quant)
    D = x.shape[-1]
    u = h[..., :D] * torch.sigmoid(h[..., D:])                                           # GLU over channels
    wd = sd[p + "sequential.4.conv.weight"][:, 0, :]                                     # (D, K) depthwise, no bias
    K = wd.shape[-1]
    left = cache if cache is not None else u.new_zeros(u.shape[0], K - 1, D)
    up = torch.cat([left, u], dim=1)                                                     # causal: pad K-1 on the left
    T = u.shape[1]
    y = sum(up[:, k:k + T] * wd[:, k] for k in range(K))
    bn = p + "sequential.5."
    y = (y - sd[bn + "running_mean"]) / torch.sqrt(sd[bn + "running_var"] + 1e-5) * sd[bn + "weight"] + sd[bn + "bias"]
    y = swish(y)
    w2 = sd[p + "sequential.7.conv.weight"][:, :, 0]
    return linear(y, w2, sd[p + "sequential.7.conv.bias"], quant), up[:, -(K - 1):]


def conformer_block(x: Tensor, sd: SD, p: str, cfg: Cfg, quant=_ident) -> Tensor:
    """ConformerEncoderBlock.forward.  LS:conf/encoder.py:76-113."""
    s = p + "sequential."
    x = x + 0.5 * ffn_module(x, sd, s + "0.module.", quant)
    m = s + "1.module."
    xn = layer_norm(x, sd[m + "layer_norm.weight"], sd[m + "layer_norm.bias"])
    x = x + retention_chunkwise(xn, sd, m + "self_attn.", cfg.n_heads, cfg.chunk, quant)
    x = x + conv_module(x, sd, s + "2.module.", None, quant)[0]
    x = x + 0.5 * ffn_module(x, sd, s + "3.module.", quant)
    return layer_norm(x, sd[s + "4.weight"], sd[s + "4.bias"])


def encoder(sd: SD, src: Sequence[Tensor], cfg: Cfg, quant=_ident) -> Tensor:
    """EmbeddingEncoderModule.forward -> ConformerEncoder.forward.  LS:model:279-285, LS:conf/encoder.py:194-201."""
    x = pad_sequence(src, 0.0)
    T = x.shape[1]
    Tp = math.ceil(T / cfg.chunk) * cfg.chunk
    x = torch.nn.functional.pad(x, (0, 0, 0, Tp - T))
    e = "enc.encoder."
    x = linear(x, sd[e + "input_projection.linear.weight"], sd[e + "input_projection.linear.bias"], quant)
    x = layer_norm(x, sd[e + "layer_norm.weight"], sd[e + "layer_norm.bias"])
    for l in range(cfg.enc_n_layers):
        x = conformer_block(x, sd, f"{e}layers.{l}.", cfg, quant)
    return x


# ----------------------------------------------------------------------------- decoder (a12)

def fusion_layer(x: Tensor, sd: SD, p: str, cfg: Cfg, quant=_ident) -> Tensor:
    """LS TransformerEncoderFusionLayer live path (post-norm).  LS:fusion:233-253.   x: (B, T, S, D)."""
    B, T, S, D = x.shape
    y = x.transpose(1, 2).reshape(B * S, T, D)
    y = layer_norm(y + retention_chunkwise(y, sd, p + "self_attn1.", cfg.n_heads, cfg.chunk, quant),
                   sd[p + "norm11.weight"], sd[p + "norm11.bias"])
    y = y.reshape(B, S, T, D).transpose(1, 2).reshape(B * T, S, D)
    y = layer_norm(y + mha(y, sd, p + "self_attn2.", cfg.n_heads, None, quant),
                   sd[p + "norm21.weight"], sd[p + "norm21.bias"])
    h = torch.relu(linear(y, sd[p + "linear1.weight"], sd[p + "linear1.bias"], quant))
    y = layer_norm(y + linear(h, sd[p + "linear2.weight"], sd[p + "linear2.bias"], quant),
                   sd[p + "norm22.weight"], sd[p + "norm22.bias"])
    return y.reshape(B, T, S, D)


def attractor_init(sd: SD, emb: Tensor, S: int, quant=_ident) -> Tensor:
    B, T, D = emb.shape
    pe = sd["dec.pos_enc.pe"][0, :S].to(emb.dtype)
    cat = torch.cat([emb[:, :, None, :].expand(B, T, S, D), pe[None, None].expand(B, T, S, D)], dim=-1)
    return linear(cat, sd["dec.convert.weight"], sd["dec.convert.bias"], quant)


def decoder(sd: SD, emb: Tensor, S: int, cfg: Cfg, quant=_ident) -> Tensor:
    """MaskedTransformerDecoderModel.forward.  LS:model:215-220."""
    x = attractor_init(sd, emb, S, quant)
    for l in range(cfg.dec_n_layers):
        x = fusion_layer(x, sd, f"dec.layers.{l}.", cfg, quant)
    return x


def conv_l2(sd: SD, emb: Tensor, ilens: Sequence[int], cfg: Cfg, quant=_ident) -> Tensor:
    """LS:model:129-136: truncate, re-pad 0, pad to a chunk multiple, Conv1d(k=19, padding=conv_delay), L2."""
    emb = pad_sequence([e[:l] for e, l in zip(emb, ilens)], 0.0)
    T = emb.shape[1]
    Tp = math.ceil(T / cfg.chunk) * cfg.chunk
    emb = torch.nn.functional.pad(emb, (0, 0, 0, Tp - T))
    w, b = sd["cnn.weight"], sd["cnn.bias"]
    K, pad = w.shape[-1], cfg.conv_delay
    xp = torch.nn.functional.pad(emb, (0, 0, pad, pad))
    out = emb.new_zeros(emb.shape[0], Tp, w.shape[0]) + b
    for k in range(K):
        out = out + quant(xp[:, k:k + Tp]) @ quant(w[:, :, k]).transpose(0, 1)
    return out / torch.linalg.vector_norm(out, dim=-1, keepdim=True)


def test(sd: SD, src: Sequence[Tensor], ilens: Sequence[int], max_nspks: int, cfg: Cfg, quant=_ident):
    """OnlineConformerRetentionDADiarization.test.  LS:model:125-147."""
    emb = encoder(sd, src, cfg, quant)
    emb = conv_l2(sd, emb, ilens, cfg, quant)
    att = decoder(sd, emb, max_nspks, cfg, quant)
    att = att / torch.linalg.vector_norm(att, dim=-1, keepdim=True)
    y = (quant(emb)[:, :, None, :] * quant(att)).sum(dim=-1)
    return [o[:l] for o, l in zip(y, ilens)], [e[:l] for e, l in zip(emb, ilens)], [a[:l] for a, l in zip(att, ilens)]


def masked_emb_consistency_loss(emb: Tensor, tgt: Sequence[Tensor], ilens: Sequence[int], max_nspks: int) -> Tensor:
    """Length-masked emb-consistency loss.  LS:model:92-113: rows >= ilen are zeroed, the squared error is summed over
    (B, T, T) and divided by sum(ilen^2).  emb: (B, T, D) with T = max(ilens)."""
    mask = pad_sequence([torch.ones(l) for l in ilens], 0.0).unsqueeze(-1)
    e = emb * mask
    amap = e @ e.transpose(-1, -2)
    n = torch.linalg.vector_norm(e, dim=-1, keepdim=True)
    amap = amap / (n @ n.transpose(-1, -2) + 1e-6)
    tp = pad_sequence([torch.nn.functional.pad(t, (0, max_nspks - t.shape[1])) for t in tgt], 0.0)
    lmap = tp @ tp.transpose(-1, -2)
    tn = torch.linalg.vector_norm(tp, dim=-1, keepdim=True)
    lmap = lmap / (tn @ tn.transpose(-1, -2) + 1e-6)
    return ((amap - lmap) ** 2).sum() / float(sum(l * l for l in ilens))


def forward(sd: SD, src: Sequence[Tensor], tgt: Sequence[Tensor], ilens: Sequence[int], cfg: Cfg, quant=_ident):
    """OnlineConformerRetentionDADiarization.forward (eval-mode arithmetic).  LS:model:74-122."""
    n_speakers = [t.shape[1] for t in tgt]
    S = max(n_speakers)
    emb = encoder(sd, src, cfg, quant)
    T = max(ilens)
    emb = conv_l2(sd, emb, ilens, cfg, quant)
    att = decoder(sd, emb, S, cfg, quant)
    att = att / torch.linalg.vector_norm(att, dim=-1, keepdim=True)
    loss = masked_emb_consistency_loss(emb[:, :T], tgt, ilens, S)
    y = (quant(emb)[:, :, None, :] * quant(att)).sum(dim=-1)
    out = [o[:l, :n] for o, l, n in zip(y, ilens, n_speakers)]
    return out, loss, [e[:l] for e, l in zip(emb, ilens)], [a[:l, 1:n] for a, l, n in zip(att, ilens, n_speakers)]


# ----------------------------------------------------------------------------- one-step path (a13)

class StreamState:
    def __init__(self, cfg: Cfg):
        self.enc_ret = [RetState() for _ in range(cfg.enc_n_layers)]
        self.enc_conv: List[Optional[Tensor]] = [None] * cfg.enc_n_layers
        self.dec_ret = [RetState() for _ in range(cfg.dec_n_layers)]
        self.conv_buf: List[Tensor] = []
        self.t = 0


def enc_step(sd: SD, st: StreamState, x_t: Tensor, cfg: Cfg, quant=_ident) -> Tensor:
    """ConformerEncoder.forward_one_step.  LS:conf/encoder.py:223-228, block :115-123.   x_t: (B,1,Din)."""
    e = "enc.encoder."
    x = linear(x_t, sd[e + "input_projection.linear.weight"], sd[e + "input_projection.linear.bias"], quant)
    x = layer_norm(x, sd[e + "layer_norm.weight"], sd[e + "layer_norm.bias"])
    for l in range(cfg.enc_n_layers):
        s = f"{e}layers.{l}.sequential."
        x = x + 0.5 * ffn_module(x, sd, s + "0.module.", quant)
        m = s + "1.module."
        xn = layer_norm(x, sd[m + "layer_norm.weight"], sd[m + "layer_norm.bias"])
        x = x + retention_step(xn, sd, m + "self_attn.", cfg.n_heads, st.enc_ret[l], quant)
        c, st.enc_conv[l] = conv_module(x, sd, s + "2.module.", st.enc_conv[l], quant)
        x = x + c
        x = x + 0.5 * ffn_module(x, sd, s + "3.module.", quant)
        x = layer_norm(x, sd[s + "4.weight"], sd[s + "4.bias"])
    return x


def dec_step(sd: SD, st: StreamState, emb_t: Tensor, S: int, cfg: Cfg, quant=_ident) -> Tensor:
    """MaskedTransformerDecoderModel.forward_one_step (LS:model:235-243) + layer step (LS:fusion:255-276)."""
    B, _, D = emb_t.shape
    a = attractor_init(sd, emb_t, S, quant)                                  # (B,1,S,D)
    for l in range(cfg.dec_n_layers):
        p = f"dec.layers.{l}."
        x = a.transpose(1, 2).reshape(B * S, 1, D)
        x = layer_norm(x + retention_step(x, sd, p + "self_attn1.", cfg.n_heads, st.dec_ret[l], quant),
                       sd[p + "norm11.weight"], sd[p + "norm11.bias"])
        x = x.reshape(B, S, D)
        x = layer_norm(x + mha(x, sd, p + "self_attn2.", cfg.n_heads, None, quant),
                       sd[p + "norm21.weight"], sd[p + "norm21.bias"])
        h = torch.relu(linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"], quant))
        x = layer_norm(x + linear(h, sd[p + "linear2.weight"], sd[p + "linear2.bias"], quant),
                       sd[p + "norm22.weight"], sd[p + "norm22.bias"])
        a = x.reshape(B, 1, S, D)
    return a


def stream_all(sd: SD, x: Tensor, max_nspks: int, cfg: Cfg, quant=_ident) -> Tensor:
    """LS-EEND/streaming_infer_dia.py:52-97 (streaming_predict): per-frame encoder step, StreamingConv1d
    (left zero padding, output from frame conv_delay+1 on), decoder step, flush with conv_delay zero embeddings."""
    st = StreamState(cfg)
    K = 2 * cfg.conv_delay + 1
    w, b = sd["cnn.weight"], sd["cnn.bias"]
    outs = []

    def push(e):
        st.t += 1
        st.conv_buf.append(e)
        if len(st.conv_buf) > K:
            st.conv_buf.pop(0)
        if st.t < K // 2 + 1:
            return
        win = [torch.zeros_like(e)] * (K - len(st.conv_buf)) + st.conv_buf
        y = b.clone().expand(e.shape[0], 1, -1)
        for k in range(K):
            y = y + quant(win[k]) @ quant(w[:, :, k]).transpose(0, 1)
        emb = y / torch.linalg.vector_norm(y, dim=-1, keepdim=True)
        a = dec_step(sd, st, emb, max_nspks, cfg, quant)
        a = a / torch.linalg.vector_norm(a, dim=-1, keepdim=True)
        outs.append((quant(emb)[:, :, None, :] * quant(a)).sum(dim=-1))

    for t in range(x.shape[1]):
        push(enc_step(sd, st, x[:, t:t + 1], cfg, quant))
    for _ in range(cfg.conv_delay):
        push(x.new_zeros(x.shape[0], 1, cfg.n_units))
    return torch.cat(outs, dim=1)


# ----------------------------------------------------------------------------- synthetic state_dict

def random_state_dict(seed: int = 0, in_size: int = 345, cfg: Optional[Cfg] = None, trained_like: bool = True) -> SD:
    """Synthetic LS-EEND weights with the reference's key names/shapes."""
    cfg = cfg or Cfg()
    g = torch.Generator().manual_seed(seed)
    D, H = cfg.n_units, cfg.n_heads

    def U(*shape, a):
        return (torch.rand(*shape, generator=g) * 2 - 1) * a

    def N(*shape, s=1.0):
        return torch.randn(*shape, generator=g) * s

    sd: SD = {}

    def lin(name, out_f, in_f, gain=1.7):
        a = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = U(out_f, in_f, a=a * gain)
        sd[name + ".bias"] = U(out_f, a=a)

    def ln(name, n):
        sd[name + ".weight"] = 1 + U(n, a=0.5) if trained_like else torch.ones(n)
        sd[name + ".bias"] = N(n, s=0.1) if trained_like else torch.zeros(n)

    def ret(name, pos):
        sd[pos + ".angle"] = (1.0 / (10000 ** torch.linspace(0, 1, D // H // 2))).unsqueeze(-1).repeat(1, 2).flatten()
        sd[pos + ".decay"] = torch.zeros(H)
        for pj in ("q_proj", "k_proj", "v_proj", "g_proj", "out_proj"):
            lin(f"{name}.{pj}", D, D, gain=2.5)

    e = "enc.encoder"
    lin(e + ".input_projection.linear", D, in_size, gain=3.0)
    ln(e + ".layer_norm", D)
    F = D * cfg.ff_expansion
    for l in range(cfg.enc_n_layers):
        s = f"{e}.layers.{l}.sequential"
        for i in (0, 3):
            ln(f"{s}.{i}.module.sequential.0", D)
            lin(f"{s}.{i}.module.sequential.1.linear", F, D)
            lin(f"{s}.{i}.module.sequential.4.linear", D, F)
        ln(f"{s}.1.module.layer_norm", D)
        ret(f"{s}.1.module.self_attn", f"{s}.1.module.ret_pos")
        c = f"{s}.2.module.sequential"
        ln(c + ".0", D)
        sd[c + ".2.conv.weight"] = U(2 * D, D, 1, a=1.7 / math.sqrt(D))
        sd[c + ".2.conv.bias"] = U(2 * D, a=0.05)
        sd[c + ".4.conv.weight"] = U(D, 1, cfg.conv_kernel, a=1.0 / math.sqrt(cfg.conv_kernel) * 1.7)
        ln(c + ".5", D)
        sd[c + ".5.running_mean"] = N(D, s=0.2) if trained_like else torch.zeros(D)
        sd[c + ".5.running_var"] = 0.5 + torch.rand(D, generator=g) if trained_like else torch.ones(D)
        sd[c + ".5.num_batches_tracked"] = torch.tensor(0)
        sd[c + ".7.conv.weight"] = U(D, D, 1, a=1.7 / math.sqrt(D))
        sd[c + ".7.conv.bias"] = U(D, a=0.05)
        ln(f"{s}.4", D)
    lin("dec.encoder", D, in_size)
    ln("dec.encoder_norm", D)
    sd["dec.pos_enc.pe"] = pe_table(D)
    lin("dec.convert", D, 2 * D)
    for l in range(cfg.dec_n_layers):
        p = f"dec.layers.{l}"
        ret(p + ".self_attn1", p + ".ret_pos1")
        sd[p + ".self_attn2.in_proj_weight"] = U(3 * D, D, a=math.sqrt(6.0 / (4 * D)) * 1.5)
        sd[p + ".self_attn2.in_proj_bias"] = U(3 * D, a=0.05)
        lin(p + ".self_attn2.out_proj", D, D)
        lin(p + ".linear1", cfg.dec_ff, D)
        lin(p + ".linear2", D, cfg.dec_ff)
        for nm in ("norm11", "norm12", "norm21", "norm22"):
            ln(p + "." + nm, D)
    sd["cnn.weight"] = U(D, D, 2 * cfg.conv_delay + 1, a=1.7 / math.sqrt(D * (2 * cfg.conv_delay + 1)))
    sd["cnn.bias"] = U(D, a=0.05)
    return sd
