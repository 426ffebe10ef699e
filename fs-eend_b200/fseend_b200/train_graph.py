"""Differentiable train-mode forward of the FS-EEND model (reference ``OnlineTransformerDADiarization.forward``,
FS-EEND/nnet/model/onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm.py:32-65) — SURVEY.md §8f N1, STARTED.

What runs where (DESIGN.md §7 has the table):
  * native kernels, forward AND backward (fseend_b200.autograd): every Linear (input projection, QKV / out projections,
    FFNs, the k=19 Conv1d as one GEMM over unfolded frames, the attractor ``convert``), every residual + LayerNorm, the
    causal time attention of encoder and decoder, the speaker-axis attention — >99 % of the step's FLOPs;
    also native: BatchNorm1d with batch statistics, the two L2 normalisations, the dot-product head;
  * torch CUDA ops (interim, small): the frame unfolding of the Conv1d, the length mask, the embedding-consistency loss (three batched matmuls), dropout masks on the
    residual branches, the optimizer.
Dropout (reference recipes train with 0.1): the residual / FFN dropouts are torch's functional dropout on the native
kernels' outputs; the attention-probability dropout inside nn.MultiheadAttention is implemented in the attention kernels
themselves (counter-based hash of (seed, sequence, head, query, key); the backward regenerates the mask).  The random
streams differ from torch's, the distribution does not.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.nn.utils.rnn import pad_sequence

from .autograd import (AddLayerNormFn, HeadFn, L2NormFn, LinearFn, batch_norm_forward, encoder_layer_forward,
                       fusion_layer_forward)
from .native import FseendError


def _require_device(dev):
    if dev.type != "cuda":
        raise FseendError("fseend_b200 runs on a CUDA sm_100 device only (move the model with .cuda())")


def fs_forward_train(model, src, tgt, ilens):
    """Returns (output list[(ilen, n_spk)], emb_consis_loss, emb list, attractors list) with autograd history."""
    enc, dec = model.enc, model.dec
    dev, dt = model.cnn.weight.device, model.cnn.weight.dtype
    _require_device(dev)
    lens = [int(l) for l in ilens]
    n_speakers = [t.shape[1] for t in tgt]
    S = max(n_speakers)
    # ---- encoder, reference :162-188
    x = pad_sequence([s.to(device=dev, dtype=dt) for s in src], batch_first=True, padding_value=-1.0)
    B, T, _ = x.shape
    x = batch_norm_forward(enc.bn, x)                                        # batch statistics incl. the -1 padding rows
    h = LinearFn.apply(x, enc.encoder.weight, enc.encoder.bias, "none")
    h = AddLayerNormFn.apply(h, None, enc.encoder_norm.weight, enc.encoder_norm.bias, enc.encoder_norm.eps)
    delay = enc.mask_delay if enc.has_mask else T
    for layer in enc.transformer_encoder.layers:
        h = encoder_layer_forward(layer, h, delay)
    # ---- truncate to ilens, re-pad with 0, Conv1d over time, L2 — reference :38-41
    Tm = max(lens)
    keep = (torch.arange(Tm, device=dev)[None, :] < torch.tensor(lens, device=dev)[:, None]).to(dt)
    emb = h[:, :Tm] * keep[..., None]
    K, pad = model.cnn.kernel_size[0], model.cnn.padding[0]
    D = emb.shape[-1]
    cols = F.pad(emb, (0, 0, pad, pad)).unfold(1, K, 1)                      # [B, Tout, D, K] view
    Tout = cols.shape[1]
    emb = LinearFn.apply(cols.reshape(B, Tout, D * K), model.cnn.weight.reshape(model.cnn.out_channels, D * K),
                         model.cnn.bias, "none")
    emb = L2NormFn.apply(emb)
    # ---- attractor decoder, reference :112-118: convert([emb ; pe_s]) = emb Wc[:, :D]^T + (pe_s Wc[:, D:]^T + b)
    Wc = dec.convert.weight
    pe = dec.pos_enc.pe[0, :S].to(dt)
    a0 = LinearFn.apply(emb, Wc[:, :D], None, "none")
    pp = LinearFn.apply(pe, Wc[:, D:], dec.convert.bias, "none")
    att = a0[:, :, None, :] + pp[None, None]
    for layer in dec.attractor_decoder.layers:
        att = fusion_layer_forward(layer, att, dec.mask_delay)
    att = L2NormFn.apply(att)
    # ---- embedding-consistency loss, reference :46-57 (torch; the (B, T, T) maps are 64 MB each at B=64, T=500)
    attn_map = emb @ emb.transpose(-1, -2)
    n = torch.norm(emb, dim=-1, keepdim=True)
    attn_map = attn_map / (n @ n.transpose(-1, -2) + 1e-6)
    tp = pad_sequence([F.pad(t.to(device=dev, dtype=dt), (0, S - t.shape[1])) for t in tgt], batch_first=True)
    tn = torch.norm(tp, dim=-1, keepdim=True)
    label_map = (tp @ tp.transpose(-1, -2)) / (tn @ tn.transpose(-1, -2) + 1e-6)
    emb_consis_loss = F.mse_loss(attn_map, label_map)
    # ---- head, reference :60-64
    y = HeadFn.apply(emb, att)
    output = [o[:l, :ns] for o, l, ns in zip(y, lens, n_speakers)]
    embs = [e[:l] for e, l in zip(emb, lens)]
    atts = [a[:l, 1:ns] for a, l, ns in zip(att, lens, n_speakers)]
    return output, emb_consis_loss, embs, atts


def standard_loss_train(ys, ts, label_delay=0):
    """Differentiable ``standard_loss`` (reference train/utils/loss.py:119-125): frame-weighted mean BCE-with-logits."""
    total = sum(F.binary_cross_entropy_with_logits(y[label_delay:], t[:len(t) - label_delay].to(y), reduction="mean")
                * (len(y) - label_delay) for y, t in zip(ys, ts))
    n_frames = sum(t.shape[0] for t in ts) - label_delay * len(ts)
    return total / n_frames
