"""CPU oracle for the FS-EEND hot path (TEST INFRASTRUCTURE — never imported by the product).

A functional restatement, in plain torch CPU tensor arithmetic (matmul / exp / sum — no
``nn.MultiheadAttention``, no ``nn.TransformerEncoderLayer``), of the reference's
encoder + attractor-decoder forward.  Every function cites the reference file:line it follows
(paths relative to /root/reference).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module.

Pinning: ``tests/golden/make_golden.py`` imports the *real* reference ``nnet`` package in the
authoring container, runs it on seeded inputs and commits the outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks this restatement against those vectors (fp32, tol 2e-5) — the
reference itself ships no golden vectors for this path, so parity is pinned by
reference outputs generated here, not by reference-owned fixtures.

The ``state_dict`` consumed here uses the reference's own key names
(FS-EEND/nnet/model/onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm.py).

``quant`` hook: every GEMM operand passes through ``quant(x)`` (identity by default).  Tests use it
to emulate fp16 / bf16 / tf32 operand rounding with fp32 accumulation, to predict what a
tensor-core path can achieve before it is run on a GPU.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def _ident(x: Tensor) -> Tensor:
    return x


# bench.py's CPU arm sets this: attention then goes through torch's fused CPU SDPA kernel — the code path the
# reference's nn.MultiheadAttention takes under torch >= 2 — instead of the explicit
# softmax(QK^T)V below.  Same arithmetic (tests/test_oracle.py checks it), faster baseline.
USE_SDPA = False


class Cfg:
    """Hyper-parameters the reference passes as constructor kwargs (FS:model:11)."""

    def __init__(self, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
                 mask_delay=0, conv_delay=9, has_mask=True):
        self.n_units = n_units
        self.n_heads = n_heads
        self.enc_n_layers = enc_n_layers
        self.dec_n_layers = dec_n_layers
        self.mask_delay = mask_delay
        self.conv_delay = conv_delay
        self.has_mask = has_mask


# ----------------------------------------------------------------------------- primitives

def linear(x: Tensor, w: Tensor, b: Optional[Tensor], quant=_ident) -> Tensor:
    """y = x W^T + b  (torch.nn.Linear semantics)."""
    y = quant(x) @ quant(w).transpose(-1, -2)
    return y if b is None else y + b


def layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """torch.nn.LayerNorm: biased variance, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def batch_norm_eval(x: Tensor, sd: SD, prefix: str, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm1d in eval mode over the channel (last) dim of (B, T, C). FS:model:166."""
    mean, var = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    return (x - mean) / torch.sqrt(var + eps) * sd[prefix + "weight"] + sd[prefix + "bias"]


def causal_mask(T: int, mask_delay: int, dtype) -> Tensor:
    """Additive float mask, 0 where key j <= query i + mask_delay, -inf elsewhere. FS:model:152-155."""
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    m = torch.zeros(T, T, dtype=dtype)
    m[j > i + mask_delay] = float("-inf")
    return m


def mha(x: Tensor, sd: SD, prefix: str, n_heads: int, mask: Optional[Tensor], quant=_ident) -> Tensor:
    """nn.MultiheadAttention(x, x, x) self-attention, batch-first (N, L, E): packed in-proj
    (3E x E), q scaled by hd^-0.5, additive mask, softmax, PV, out-proj."""
    N, L, E = x.shape
    hd = E // n_heads
    qkv = linear(x, sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"], quant)
    q, k, v = qkv.split(E, dim=-1)
    q = q.reshape(N, L, n_heads, hd).transpose(1, 2)
    k = k.reshape(N, L, n_heads, hd).transpose(1, 2)
    v = v.reshape(N, L, n_heads, hd).transpose(1, 2)
    if USE_SDPA and quant is _ident:
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask)
    else:
        s = (quant(q) @ quant(k).transpose(-1, -2)) * (hd ** -0.5)
        if mask is not None:
            s = s + mask
        p = torch.softmax(s, dim=-1)
        o = quant(p) @ quant(v)                   # (N, H, L, hd)
    o = o.transpose(1, 2).reshape(N, L, E)
    return linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"], quant)


def ffn_relu(x: Tensor, sd: SD, prefix: str, quant=_ident) -> Tensor:
    """linear2(relu(linear1(x)))  — FS:fusion:397-399 / torch TransformerEncoderLayer._ff_block."""
    h = torch.relu(linear(x, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"], quant))
    return linear(h, sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"], quant)


# ----------------------------------------------------------------------------- encoder (a1, a2)

def pad_sequence(seqs: Sequence[Tensor], value: float) -> Tensor:
    T = max(s.shape[0] for s in seqs)
    out = seqs[0].new_full((len(seqs), T) + tuple(seqs[0].shape[1:]), value)
    for i, s in enumerate(seqs):
        out[i, : s.shape[0]] = s
    return out


def encoder(sd: SD, src: Sequence[Tensor], cfg: Cfg, quant=_ident) -> Tensor:
    """MaskedTransformerEncoderModel.forward, eval mode.  FS:model:162-188.
    pad(-1) -> BN -> Linear -> LN -> enc_n_layers x post-norm TransformerEncoderLayer."""
    x = pad_sequence(src, -1.0)                                        # FS:model:165
    x = batch_norm_eval(x, sd, "enc.bn.")                              # FS:model:166
    x = linear(x, sd["enc.encoder.weight"], sd["enc.encoder.bias"], quant)   # FS:model:173
    x = layer_norm(x, sd["enc.encoder_norm.weight"], sd["enc.encoder_norm.bias"])  # FS:model:174
    T = x.shape[1]
    mask = causal_mask(T, cfg.mask_delay, x.dtype) if cfg.has_mask else None
    for l in range(cfg.enc_n_layers):                                  # FS:fusion:129-131
        p = f"enc.transformer_encoder.layers.{l}."
        x = layer_norm(x + mha(x, sd, p + "self_attn.", cfg.n_heads, mask, quant),
                       sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        x = layer_norm(x + ffn_relu(x, sd, p, quant), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    return x


# ----------------------------------------------------------------------------- conv + L2 (a3)

def conv_l2(sd: SD, emb: Tensor, ilens: Sequence[int], cfg: Cfg, quant=_ident) -> Tensor:
    """Truncate to ilens, re-pad with 0, Conv1d(k=2*delay+1, padding=9), L2-normalise.
    FS:model:38-41 / 71-74."""
    B, T, D = emb.shape
    emb = pad_sequence([e[:l] for e, l in zip(emb, ilens)], 0.0)
    T = emb.shape[1]
    w, b = sd["cnn.weight"], sd["cnn.bias"]                            # (Dout, Din, K)
    K = w.shape[-1]
    pad = 9                                                             # FS:model:30 hard-codes padding=9
    xp = torch.nn.functional.pad(emb, (0, 0, pad, pad))
    Tout = T + 2 * pad - K + 1
    out = emb.new_zeros(B, Tout, w.shape[0]) + b
    for k in range(K):
        out = out + quant(xp[:, k:k + Tout]) @ quant(w[:, :, k]).transpose(0, 1)
    return out / torch.linalg.vector_norm(out, dim=-1, keepdim=True)


# ----------------------------------------------------------------------------- decoder (a4, a5)

def speaker_slot_pe(sd: SD, S: int) -> Tensor:
    """PositionalEncoding indexed by *speaker slot* (FS:model:218-224): pe[0, :S, :]."""
    return sd["dec.pos_enc.pe"][0, :S]


def attractor_init(sd: SD, emb: Tensor, S: int, quant=_ident) -> Tensor:
    """convert(cat[emb repeated over S, pe]).  FS:model:113-114.  Materialised as written."""
    B, T, D = emb.shape
    pe = speaker_slot_pe(sd, S).to(emb.dtype)
    cat = torch.cat([emb[:, :, None, :].expand(B, T, S, D), pe[None, None].expand(B, T, S, D)], dim=-1)
    return linear(cat, sd["dec.convert.weight"], sd["dec.convert.bias"], quant)


def fusion_layer(x: Tensor, sd: SD, p: str, cfg: Cfg, mask: Optional[Tensor], quant=_ident) -> Tensor:
    """TransformerEncoderFusionLayer live path (post-norm).  FS:fusion:356-376."""
    B, T, S, D = x.shape
    y = x.transpose(1, 2).reshape(B * S, T, D)                          # FS:fusion:358
    y = layer_norm(y + mha(y, sd, p + "self_attn1.", cfg.n_heads, mask, quant),
                   sd[p + "norm11.weight"], sd[p + "norm11.bias"])      # FS:fusion:363
    y = y.reshape(B, S, T, D).transpose(1, 2).reshape(B * T, S, D)      # FS:fusion:365
    y = layer_norm(y + mha(y, sd, p + "self_attn2.", cfg.n_heads, None, quant),
                   sd[p + "norm21.weight"], sd[p + "norm21.bias"])      # FS:fusion:372
    y = layer_norm(y + ffn_relu(y, sd, p, quant), sd[p + "norm22.weight"], sd[p + "norm22.bias"])  # :373
    return y.reshape(B, T, S, D)


def decoder(sd: SD, emb: Tensor, S: int, cfg: Cfg, quant=_ident) -> Tensor:
    """MaskedTransformerDecoderModel.forward.  FS:model:112-118."""
    x = attractor_init(sd, emb, S, quant)
    mask = causal_mask(emb.shape[1], cfg.mask_delay, emb.dtype)         # always masked (FS:model:116)
    for l in range(cfg.dec_n_layers):
        x = fusion_layer(x, sd, f"dec.attractor_decoder.layers.{l}.", cfg, mask, quant)
    return x


# ----------------------------------------------------------------------------- model (a3, a6, a7)

def logits_head(emb: Tensor, att: Tensor, quant=_ident) -> Tuple[Tensor, Tensor]:
    """L2-normalise attractors, y[b,t,s] = emb[b,t,:] . att[b,t,s,:].  FS:model:43,60 / 76,79."""
    att = att / torch.linalg.vector_norm(att, dim=-1, keepdim=True)
    y = (quant(emb)[:, :, None, :] * quant(att)).sum(dim=-1)
    return y, att


def test(sd: SD, src: Sequence[Tensor], ilens: Sequence[int], max_nspks: int, cfg: Cfg, quant=_ident,
         return_padded: bool = False):
    """OnlineTransformerDADiarization.test.  FS:model:67-84."""
    emb = encoder(sd, src, cfg, quant)
    emb = conv_l2(sd, emb, ilens, cfg, quant)
    att = decoder(sd, emb, max_nspks, cfg, quant)
    y, att = logits_head(emb, att, quant)
    if return_padded:
        return y, emb, att
    out = [o[:l] for o, l in zip(y, ilens)]
    embs = [e[:l] for e, l in zip(emb, ilens)]
    atts = [a[:l] for a, l in zip(att, ilens)]
    return out, embs, atts


def emb_consistency_loss(emb: Tensor, tgt: Sequence[Tensor], max_nspks: int) -> Tensor:
    """MSE(cos-sim map of emb, cos-sim map of labels), mean over B*T*T.  FS:model:46-57."""
    attn_map = emb @ emb.transpose(-1, -2)
    n = torch.linalg.vector_norm(emb, dim=-1, keepdim=True)
    attn_map = attn_map / (n @ n.transpose(-1, -2) + 1e-6)
    tp = [torch.nn.functional.pad(t, (0, max_nspks - t.shape[1])) for t in tgt]
    tp = pad_sequence(tp, 0.0)
    label_map = tp @ tp.transpose(-1, -2)
    tn = torch.linalg.vector_norm(tp, dim=-1, keepdim=True)
    label_map = label_map / (tn @ tn.transpose(-1, -2) + 1e-6)
    return ((attn_map - label_map) ** 2).mean()


def forward(sd: SD, src: Sequence[Tensor], tgt: Sequence[Tensor], ilens: Sequence[int], cfg: Cfg, quant=_ident):
    """OnlineTransformerDADiarization.forward (eval-mode arithmetic: dropout off).  FS:model:32-65."""
    n_speakers = [t.shape[1] for t in tgt]
    S = max(n_speakers)
    emb = encoder(sd, src, cfg, quant)
    emb = conv_l2(sd, emb, ilens, cfg, quant)
    att = decoder(sd, emb, S, cfg, quant)
    y, att = logits_head(emb, att, quant)
    loss = emb_consistency_loss(emb, tgt, S)
    out = [o[:l, :n] for o, l, n in zip(y, ilens, n_speakers)]
    embs = [e[:l] for e, l in zip(emb, ilens)]
    atts = [a[:l, 1:n] for a, l, n in zip(att, ilens, n_speakers)]
    return out, loss, embs, atts


# ----------------------------------------------------------------------------- streaming (a8)

class StreamState:
    """Caches of the frame-by-frame path.  The reference caches layer *inputs* and re-projects them
    every step (FS:stream_mod:28-35); arithmetic-wise that equals attending over the projected
    history, which is what is restated here."""

    def __init__(self, cfg: Cfg):
        self.enc_x: List[Optional[Tensor]] = [None] * cfg.enc_n_layers
        self.dec_x: List[Optional[Tensor]] = [None] * cfg.dec_n_layers
        self.conv_buf: List[Tensor] = []
        self.t = 0


def _inc_mha(x_t: Tensor, hist: Optional[Tensor], sd: SD, prefix: str, n_heads: int, quant=_ident):
    """IncrementalSelfAttention: query = x_t (N,1,E), keys/values = cat[hist, x_t].  FS:stream_mod:10-37."""
    kv = x_t if hist is None else torch.cat([hist, x_t], dim=1)
    N, L, E = kv.shape
    hd = E // n_heads
    w, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    q = linear(x_t, w[:E], b[:E], quant).reshape(N, 1, n_heads, hd).transpose(1, 2)
    k = linear(kv, w[E:2 * E], b[E:2 * E], quant).reshape(N, L, n_heads, hd).transpose(1, 2)
    v = linear(kv, w[2 * E:], b[2 * E:], quant).reshape(N, L, n_heads, hd).transpose(1, 2)
    p = torch.softmax((quant(q) @ quant(k).transpose(-1, -2)) * hd ** -0.5, dim=-1)
    o = (quant(p) @ quant(v)).transpose(1, 2).reshape(N, 1, E)
    return linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"], quant), kv


def stream_step(sd: SD, st: StreamState, x_t: Optional[Tensor], max_nspks: int, cfg: Cfg, quant=_ident):
    """StreamingTransformerEDADiarization.test for one frame, using the *masked* model's key names
    (the reference maps them with copy_params_from_masked_to_streaming, FS copy_params.py:7-62).
    x_t: (B,1,Din), or None for a flush step (dummy_conv_input=True, FS:stream_model:42-43).
    Returns (B,1,S) logits or None while the look-ahead conv has < center+1 frames (FS:stream_mod:163-166)."""
    D = cfg.n_units
    if x_t is None:
        B = st.conv_buf[-1].shape[0]
        e = st.conv_buf[-1].new_zeros(B, 1, D)
    else:
        x = batch_norm_eval(x_t, sd, "enc.bn.")
        x = layer_norm(linear(x, sd["enc.encoder.weight"], sd["enc.encoder.bias"], quant),
                       sd["enc.encoder_norm.weight"], sd["enc.encoder_norm.bias"])
        for l in range(cfg.enc_n_layers):                               # FS:stream_mod:117-127
            p = f"enc.transformer_encoder.layers.{l}."
            a, st.enc_x[l] = _inc_mha(x, st.enc_x[l], sd, p + "self_attn.", cfg.n_heads, quant)
            x = layer_norm(a + x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
            x = layer_norm(ffn_relu(x, sd, p, quant) + x, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        e = x
    # streaming conv ring buffer, FS:stream_mod:141-167
    K = 2 * cfg.conv_delay + 1
    st.t += 1
    st.conv_buf.append(e)
    if len(st.conv_buf) > K:
        st.conv_buf.pop(0)
    if st.t < K // 2 + 1:
        return None
    win = [torch.zeros_like(e)] * (K - len(st.conv_buf)) + st.conv_buf
    w, b = sd["cnn.weight"], sd["cnn.bias"]
    y = b.clone().expand(e.shape[0], 1, -1)
    for k in range(K):
        y = y + quant(win[k]) @ quant(w[:, :, k]).transpose(0, 1)
    emb = y / torch.linalg.vector_norm(y, dim=-1, keepdim=True)          # FS:stream_model:50
    # decoder step, FS:stream_mod:248-269 and 187-213
    B = emb.shape[0]
    S = max_nspks
    a = attractor_init(sd, emb, S, quant)                                # (B,1,S,D)
    for l in range(cfg.dec_n_layers):
        p = f"dec.attractor_decoder.layers.{l}."
        xt = a.transpose(1, 2).reshape(B * S, 1, D)
        o, st.dec_x[l] = _inc_mha(xt, st.dec_x[l], sd, p + "self_attn1.", cfg.n_heads, quant)
        xt = layer_norm(o + xt, sd[p + "norm11.weight"], sd[p + "norm11.bias"])
        xs = xt.reshape(B, S, D)
        xs = layer_norm(mha(xs, sd, p + "self_attn2.", cfg.n_heads, None, quant) + xs,
                        sd[p + "norm21.weight"], sd[p + "norm21.bias"])
        xs = layer_norm(ffn_relu(xs, sd, p, quant) + xs, sd[p + "norm22.weight"], sd[p + "norm22.bias"])
        a = xs.reshape(B, 1, S, D)
    y, _ = logits_head(emb, a, quant)
    return y                                                              # (B,1,S)


def stream_all(sd: SD, x: Tensor, max_nspks: int, cfg: Cfg, quant=_ident) -> Tensor:
    """Frame loop + flush of FS-EEND/streaming_infer_dia.py:77-86; returns (B, T, S) logits."""
    st = StreamState(cfg)
    outs = []
    for t in range(x.shape[1]):
        y = stream_step(sd, st, x[:, t:t + 1], max_nspks, cfg, quant)
        if y is not None:
            outs.append(y)
    for _ in range(cfg.conv_delay):
        outs.append(stream_step(sd, st, None, max_nspks, cfg, quant))
    return torch.cat(outs, dim=1)


# ----------------------------------------------------------------------------- synthetic state_dict

def pe_table(d_model: int, max_len: int = 5000) -> Tensor:
    """FS:model:209-216."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def random_state_dict(seed: int = 0, in_size: int = 345, n_units: int = 256, n_heads: int = 4,
                      enc_n_layers: int = 4, dec_n_layers: int = 2, ff: int = 2048, dec_ff: int = 2048,
                      conv_k: int = 19, trained_like: bool = True) -> SD:
    """Synthetic weights with the reference's key names/shapes (SURVEY.md §8b state_dict ABI).
    'trained_like' randomises LN/BN affines and running stats so folding bugs show."""
    g = torch.Generator().manual_seed(seed)
    D = n_units

    def U(*shape, a):
        return (torch.rand(*shape, generator=g) * 2 - 1) * a

    def N(*shape, s=1.0):
        return torch.randn(*shape, generator=g) * s

    sd: SD = {}

    def lin(name, out_f, in_f):
        a = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = U(out_f, in_f, a=a * 1.7)
        sd[name + ".bias"] = U(out_f, a=a) if trained_like else torch.zeros(out_f)

    def ln(name, n):
        if trained_like:
            sd[name + ".weight"] = 1 + U(n, a=0.5)
            sd[name + ".bias"] = N(n, s=0.1)
        else:
            sd[name + ".weight"] = torch.ones(n)
            sd[name + ".bias"] = torch.zeros(n)

    def attn(name):
        sd[name + ".in_proj_weight"] = U(3 * D, D, a=math.sqrt(6.0 / (4 * D)) * 1.5)
        sd[name + ".in_proj_bias"] = U(3 * D, a=0.05) if trained_like else torch.zeros(3 * D)
        lin(name + ".out_proj", D, D)

    ln("enc.bn", in_size)
    sd["enc.bn.running_mean"] = N(in_size) if trained_like else torch.zeros(in_size)
    sd["enc.bn.running_var"] = 0.5 + 1.5 * torch.rand(in_size, generator=g) if trained_like else torch.ones(in_size)
    sd["enc.bn.num_batches_tracked"] = torch.tensor(0)
    sd["enc.encoder.weight"] = U(D, in_size, a=0.1)
    sd["enc.encoder.bias"] = U(D, a=0.05) if trained_like else torch.zeros(D)
    ln("enc.encoder_norm", D)
    for l in range(enc_n_layers):
        p = f"enc.transformer_encoder.layers.{l}"
        attn(p + ".self_attn")
        lin(p + ".linear1", ff, D)
        lin(p + ".linear2", D, ff)
        ln(p + ".norm1", D)
        ln(p + ".norm2", D)
    sd["cnn.weight"] = U(D, D, conv_k, a=1.0 / math.sqrt(D * conv_k) * 1.7)
    sd["cnn.bias"] = U(D, a=0.05)
    lin("dec.encoder", D, in_size)          # dead parameters, kept for ABI (FS:model:99-100)
    ln("dec.encoder_norm", D)
    sd["dec.pos_enc.pe"] = pe_table(D)
    lin("dec.convert", D, 2 * D)
    for l in range(dec_n_layers):
        p = f"dec.attractor_decoder.layers.{l}"
        attn(p + ".self_attn1")
        attn(p + ".self_attn2")
        lin(p + ".linear1", dec_ff, D)
        lin(p + ".linear2", D, dec_ff)
        for nm in ("norm11", "norm12", "norm21", "norm22"):
            ln(p + "." + nm, D)
    return sd


def wide_state_dict(seed: int = 0, alpha: float = 1.5, S: int = 6) -> SD:
    """Synthetic weights whose logits use most of the cosine range: default random weights give logits with std ~0.02
    (range -0.2..0.1), so a 1e-3 absolute bound is ~5 % of the signal.  Here the attractor init passes the embedding
    through (convert = alpha * I + noise on the embedding half; the positional half maps slot s onto a designed offset
    of growing size) and the decoder layers are near-identity, so logits span alpha > 0: +0.15..+0.97 over the slots,
    alpha < 0: -0.97..-0.2."""
    sd = random_state_dict(seed=seed, trained_like=True)
    D = sd["dec.convert.weight"].shape[0]
    g = torch.Generator().manual_seed(1000 + seed)
    pe = sd["dec.pos_enc.pe"][0, :S]
    targets = torch.randn(S, D, generator=g)
    targets = targets / targets.norm(dim=1, keepdim=True)
    beta = torch.tensor([0.0, 0.15, 0.4, 0.8, 1.6, 4.0, 6.0, 8.0, 10.0, 12.0, 14.0, 16.0, 18.0, 20.0, 22.0, 24.0])[:S]
    targets = targets * (beta * abs(alpha))[:, None]
    M = targets.T @ torch.linalg.pinv(pe.T)                    # M pe_s = target_s
    W = sd["dec.convert.weight"].clone()
    W[:, :D] = alpha * torch.eye(D) + 0.1 * W[:, :D]
    W[:, D:] = M
    sd["dec.convert.weight"] = W
    sd["dec.convert.bias"] = sd["dec.convert.bias"] * 0.05
    n_dec = len({k.split(".")[3] for k in sd if k.startswith("dec.attractor_decoder.layers.")})
    for l in range(n_dec):
        p = f"dec.attractor_decoder.layers.{l}."
        for k in ("self_attn1.out_proj", "self_attn2.out_proj", "linear2"):
            sd[p + k + ".weight"] = sd[p + k + ".weight"] * 0.1
            sd[p + k + ".bias"] = sd[p + k + ".bias"] * 0.1
        for nm in ("norm11", "norm21", "norm22"):
            sd[p + nm + ".weight"] = torch.ones(D) + 0.05 * (sd[p + nm + ".weight"] - 1)
            sd[p + nm + ".bias"] = sd[p + nm + ".bias"] * 0.05
    return sd


def synthetic_features(B: int, T: int, in_size: int = 345, seed: int = 777, lens: Optional[Sequence[int]] = None):
    """x ~ N(0,1) float32 (T_i, 345) per item, seed 777."""
    g = torch.Generator().manual_seed(seed)
    lens = list(lens) if lens is not None else [T] * B
    return [torch.randn(l, in_size, generator=g) for l in lens], lens


def synthetic_labels(seed: int, lens: Sequence[int], n_spks: Sequence[int], p: float = 0.35) -> List[Tensor]:
    """0/1 speaker-activity matrices (T_i, n_spk_i) for the training-step goldens (tests/golden/make_golden_train.py)."""
    g = torch.Generator().manual_seed(seed + 1)
    return [(torch.rand(l, n, generator=g) < p).float() for l, n in zip(lens, n_spks)]

