"""Drop-in for the reference's FS-EEND/nnet/model/onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm.py.

Same class names, constructor kwargs, attributes (.enc .dec .cnn .delay .n_speakers), method signatures and
state_dict keys (checked against tests/golden/fs_state_dict_abi.txt, generated from the real reference).
The forward arithmetic is executed by the sm_100a shared library through its C ABI
(include/fseend_b200.h); the modules below only own the parameters.  No CPU path exists: calling
test()/forward() without the built library or off-GPU raises.
"""
import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from fseend_b200.native import NativeCacheMixin, op_embloss
from ..modules.merge_tfm_encoder import TransformerEncoder, TransformerEncoderFusionLayer


def _pad_labels(tgt, lens, T, max_nspks, dev):
    """list of (T_i, n_spk_i) activity matrices -> zero-padded fp32 [B, T, max_nspks] on the device."""
    labels = torch.zeros(len(tgt), T, max_nspks, device=dev, dtype=torch.float32)
    for b, t in enumerate(tgt):
        n = min(t.shape[0], T)
        labels[b, :n, :t.shape[1]] = t[:n].to(device=dev, dtype=torch.float32)
    return labels


class PositionalEncoding(nn.Module):
    """Sinusoid table indexed by *speaker slot* (reference :190-224); only the ``pe`` buffer matters."""

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class MaskedTransformerEncoderModel(nn.Module):
    """Parameters of the causal embedding encoder (reference :120-188): bn, encoder, encoder_norm,
    transformer_encoder.layers[i] (torch.nn.TransformerEncoderLayer parameter layout)."""

    def __init__(self, in_size, n_heads, n_units, n_layers, dim_feedforward=2048, dropout=0.5, has_mask=False,
                 max_seqlen=500, has_pos=False, mask_delay=0):
        super().__init__()
        self.in_size, self.n_heads, self.n_units, self.n_layers = in_size, n_heads, n_units, n_layers
        self.has_pos, self.has_mask, self.max_seqlen, self.mask_delay = has_pos, has_mask, max_seqlen, mask_delay
        self.dim_feedforward = dim_feedforward
        if has_pos:
            raise NotImplementedError("has_pos=True is not used by any reference config")
        self.bn = nn.BatchNorm1d(in_size)
        self.encoder = nn.Linear(in_size, n_units)
        self.encoder_norm = nn.LayerNorm(n_units)
        encoder_layers = nn.TransformerEncoderLayer(n_units, n_heads, dim_feedforward, dropout)
        self.transformer_encoder = TransformerEncoder(encoder_layers, n_layers)
        self.init_weights()

    def init_weights(self):
        initrange = 0.1
        self.encoder.bias.data.zero_()
        self.encoder.weight.data.uniform_(-initrange, initrange)


class MaskedTransformerDecoderModel(nn.Module):
    """Parameters of the attractor decoder (reference :87-118); ``encoder``/``encoder_norm`` are dead
    parameters kept for checkpoint compatibility."""

    def __init__(self, in_size, n_heads, n_units, n_layers, dim_feedforward, dropout=0.5, has_mask=False,
                 max_seqlen=500, has_pos=False, mask_delay=0):
        super().__init__()
        self.in_size, self.n_heads, self.n_units, self.n_layers = in_size, n_heads, n_units, n_layers
        self.has_pos, self.has_mask, self.max_seqlen, self.mask_delay = has_pos, has_mask, max_seqlen, mask_delay
        self.dim_feedforward = dim_feedforward
        self.encoder = nn.Linear(in_size, n_units)
        self.encoder_norm = nn.LayerNorm(n_units)
        self.pos_enc = PositionalEncoding(n_units, dropout)
        self.convert = nn.Linear(n_units * 2, n_units)
        decoder_layers = TransformerEncoderFusionLayer(n_units, n_heads, dim_feedforward, dropout, batch_first=True)
        self.attractor_decoder = TransformerEncoder(decoder_layers, n_layers)


class OnlineTransformerDADiarization(NativeCacheMixin, nn.Module):
    def __init__(self, n_speakers, in_size, n_units, n_heads, enc_n_layers, dec_n_layers, dropout, has_mask,
                 max_seqlen, dec_dim_feedforward, conv_delay=9, mask_delay=0, decom_kernel_size=64):
        super().__init__()
        self.n_speakers = n_speakers
        self.delay = conv_delay
        self.enc = MaskedTransformerEncoderModel(
            in_size, n_heads, n_units, enc_n_layers, dropout=dropout, has_mask=has_mask, max_seqlen=max_seqlen,
            mask_delay=mask_delay)
        self.dec = MaskedTransformerDecoderModel(
            in_size, n_heads, n_units, dec_n_layers, dim_feedforward=dec_dim_feedforward, dropout=dropout,
            has_mask=has_mask, max_seqlen=max_seqlen, mask_delay=mask_delay)
        self.cnn = nn.Conv1d(n_units, n_units, kernel_size=2 * conv_delay + 1, padding=9)
        self._native = None
        self._native_key = None

    # ------------------------------------------------------------------ native model management
    def _native_cfg(self):
        return dict(in_size=self.enc.in_size, n_units=self.enc.n_units, n_heads=self.enc.n_heads,
                    enc_n_layers=self.enc.n_layers, dec_n_layers=self.dec.n_layers,
                    enc_dim_feedforward=self.enc.dim_feedforward, dec_dim_feedforward=self.dec.dim_feedforward,
                    conv_kernel=self.cnn.kernel_size[0], conv_padding=self.cnn.padding[0],
                    mask_delay=self.enc.mask_delay, has_mask=bool(self.enc.has_mask),
                    bn_eps=self.enc.bn.eps, ln_eps=self.enc.encoder_norm.eps)

    def native(self):
        """The native model for the current parameter values (rebuilt when any parameter/buffer changed)."""
        from fseend_b200.native import FsModel
        return self._native_cached(lambda: FsModel(self._native_cfg(), self.state_dict()))

    def _pack(self, src: Sequence[Tensor], ilens: Sequence[int]):
        dev = self.cnn.weight.device
        if dev.type != "cuda":
            raise RuntimeError("fseend_b200 runs on a CUDA sm_100 device only (move the model with .cuda())")
        lens = [int(l) for l in ilens]
        rows = []
        for s, l in zip(src, lens):
            if l > s.shape[0]:
                raise ValueError("ilens exceeds the feature length")
            if l < s.shape[0] and (self.enc.mask_delay != 0 or not self.enc.has_mask):
                # with look-ahead, frames >= ilen would influence frames < ilen inside the encoder
                raise NotImplementedError("ilens < len(src) requires mask_delay == 0")
            rows.append(s[:l])
        x = torch.cat([r.to(device=dev, dtype=torch.float32) for r in rows], dim=0).contiguous()
        return x, lens

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def test(self, src, ilens, max_nspks=6):
        """Reference :67-84.  Returns (list[logits (ilen, max_nspks)], list[emb (ilen, D)], list[attractors])."""
        x, lens = self._pack(src, ilens)
        with self._on_device():
            y, emb, att = self.native().forward(x, lens, max_nspks, want_emb=True, want_att=True)
        output = [o[:l] for o, l in zip(y, lens)]
        emb = [e[:l] for e, l in zip(emb, lens)]
        attractors = [a[:l] for a, l in zip(att, lens)]
        return output, emb, attractors

    @torch.no_grad()
    def test_logits(self, src, ilens, max_nspks=6):
        """test() without materialising the fp32 embedding / attractor by-products (bench path)."""
        x, lens = self._pack(src, ilens)
        with self._on_device():
            y, _, _ = self.native().forward(x, lens, max_nspks)
        return [o[:l] for o, l in zip(y, lens)]

    def forward(self, src, tgt, ilens):
        """Reference :32-65.  Without gradients (eval, or no_grad): the inference pipeline of the shared library.  In
        train mode with gradients enabled: the differentiable path of fseend_b200.train_graph (native forward + backward
        kernels for the GEMMs / attention / LayerNorm; SURVEY §8f N1, started)."""
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from fseend_b200.train_graph import fs_forward_train
            with self._on_device():
                return fs_forward_train(self, src, tgt, ilens)
        with torch.no_grad(), self._on_device():
            n_speakers = [t.shape[1] for t in tgt]
            max_nspks = max(n_speakers)
            x, lens = self._pack(src, ilens)
            y, emb, att = self.native().forward(x, lens, max_nspks, want_emb=True, want_att=True)
            # embedding-consistency loss (reference :46-57) in one kernel (csrc/embloss.cu); padded rows t >= ilen
            # are part of the mean, as in the reference
            labels = _pad_labels(tgt, lens, y.shape[1], max_nspks, y.device)
            emb_consis_loss = op_embloss(emb, labels)
            output = [o[:l, :n] for o, l, n in zip(y, lens, n_speakers)]
            emb = [e[:l] for e, l in zip(emb, lens)]
            attractors = [a[:l, 1:n] for a, l, n in zip(att, lens, n_speakers)]
        return output, emb_consis_loss, emb, attractors
