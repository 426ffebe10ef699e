"""Drop-in for the reference's LS-EEND model file (same name under LS-EEND/nnet/model/): class names, constructor
kwargs, attributes (.n_units .delay .enc.encoder.layers .enc.encoder._conv_kernel_size .dec.layers .cnn) and
state_dict keys match (tests/golden/ls_state_dict_abi.txt, dumped from the real reference).  The arithmetic runs in
the sm_100a library through the C ABI (fseend_ls_*); there is no CPU path."""
import math

import torch
import torch.nn as nn
from torch import Tensor

from ..conformer.encoder import ConformerEncoder
from ..modules.merge_retnet_layer import TransformerEncoderFusionLayer


class PositionalEncoding(nn.Module):
    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class EmbeddingEncoderModule(nn.Module):
    def __init__(self, in_size, n_units, n_heads, n_layers, recurrent_chunk_size, feed_forward_expansion_factor=8,
                 conv_expansion_factor=2, dropout=0.1, conv_kernel_size=16, half_step_residual=True, max_seqlen=500):
        super().__init__()
        self.in_size, self.n_units, self.n_heads, self.n_layers = in_size, n_units, n_heads, n_layers
        self.max_seqlen, self.recurrent_chunk_size = max_seqlen, recurrent_chunk_size
        self.feed_forward_expansion_factor = feed_forward_expansion_factor
        self.encoder = ConformerEncoder(
            input_dim=in_size, encoder_dim=n_units, num_layers=n_layers, num_attention_heads=n_heads,
            feed_forward_expansion_factor=feed_forward_expansion_factor, conv_expansion_factor=conv_expansion_factor,
            feed_forward_dropout_p=dropout, attention_dropout_p=dropout, conv_dropout_p=dropout,
            conv_kernel_size=conv_kernel_size, half_step_residual=half_step_residual,
            recurrent_chunk_size=recurrent_chunk_size)


class MaskedTransformerDecoderModel(nn.Module):
    def __init__(self, in_size, n_heads, n_units, n_layers, recurrent_chunk_size, dim_feedforward, dropout=0.5,
                 max_seqlen=500, has_pos=False, mask_delay=0):
        super().__init__()
        self.in_size, self.n_heads, self.n_units, self.n_layers = in_size, n_heads, n_units, n_layers
        self.has_pos, self.max_seqlen, self.mask_delay = has_pos, max_seqlen, mask_delay
        self.dim_feedforward = dim_feedforward
        self.encoder = nn.Linear(in_size, n_units)          # dead parameters (never used by the reference either)
        self.encoder_norm = nn.LayerNorm(n_units)
        self.pos_enc = PositionalEncoding(n_units, dropout)
        self.convert = nn.Linear(n_units * 2, n_units)
        self.layers = nn.ModuleList([
            TransformerEncoderFusionLayer(n_units, n_heads, recurrent_chunk_size, dim_feedforward, dropout,
                                          batch_first=True) for _ in range(n_layers)])


class OnlineConformerRetentionDADiarization(nn.Module):
    def __init__(self, n_speakers, in_size, n_units, n_heads, enc_n_layers, dec_n_layers, dropout, max_seqlen,
                 recurrent_chunk_size: int = 500, feed_forward_expansion_factor: int = 8,
                 dec_dim_feedforward: int = 2048, conv_expansion_factor: int = 2, conv_kernel_size: int = 16,
                 half_step_residual: bool = True, conv_delay=9, mask_delay=0):
        super().__init__()
        if not half_step_residual or conv_expansion_factor != 2:
            raise NotImplementedError("fseend_b200 implements the published LS-EEND configuration")
        self.n_speakers, self.n_units, self.delay = n_speakers, n_units, conv_delay
        self.max_seqlen, self.recurrent_chunk_size = max_seqlen, recurrent_chunk_size
        self.enc = EmbeddingEncoderModule(
            in_size=in_size, n_units=n_units, n_heads=n_heads, n_layers=enc_n_layers,
            recurrent_chunk_size=recurrent_chunk_size, feed_forward_expansion_factor=feed_forward_expansion_factor,
            conv_expansion_factor=conv_expansion_factor, dropout=dropout, conv_kernel_size=conv_kernel_size,
            half_step_residual=half_step_residual, max_seqlen=max_seqlen)
        self.dec = MaskedTransformerDecoderModel(
            in_size, n_heads=n_heads, n_units=n_units, n_layers=dec_n_layers,
            recurrent_chunk_size=recurrent_chunk_size, dim_feedforward=dec_dim_feedforward, dropout=dropout,
            max_seqlen=max_seqlen, mask_delay=mask_delay)
        self.cnn = nn.Conv1d(n_units, n_units, kernel_size=2 * conv_delay + 1, padding=conv_delay)
        self._native = None
        self._native_key = None

    def _native_cfg(self):
        return dict(in_size=self.enc.in_size, n_units=self.n_units, n_heads=self.enc.n_heads,
                    enc_n_layers=self.enc.n_layers, dec_n_layers=self.dec.n_layers,
                    feed_forward_expansion_factor=self.enc.feed_forward_expansion_factor,
                    dec_dim_feedforward=self.dec.dim_feedforward,
                    conv_kernel_size=self.enc.encoder._conv_kernel_size,
                    recurrent_chunk_size=self.recurrent_chunk_size, conv_delay=self.delay)

    def native(self):
        from fseend_b200.native import LsModel
        tensors = list(self.parameters()) + list(self.buffers())
        key = (tuple((t.data_ptr(), t._version) for t in tensors), torch.cuda.current_device())
        if self._native is None or key != self._native_key:
            self._native = LsModel(self._native_cfg(), self.state_dict())
            self._native_key = key
        return self._native

    def _pack(self, src, ilens):
        dev = self.cnn.weight.device
        if dev.type != "cuda":
            raise RuntimeError("fseend_b200 runs on a CUDA sm_100 device only (move the model with .cuda())")
        lens = [int(l) for l in ilens]
        for s, l in zip(src, lens):
            if l > s.shape[0]:
                raise ValueError("ilens exceeds the feature length")
        # every LS-EEND sub-layer is causal, so frames >= ilen never influence frames < ilen: truncation is exact
        x = torch.cat([s[:l].to(device=dev, dtype=torch.float32) for s, l in zip(src, lens)], dim=0).contiguous()
        return x, lens

    @torch.no_grad()
    def test(self, src, ilens, max_nspks=6):
        """Reference :125-147.  Returns (list[logits (ilen, max_nspks)], list[emb (ilen, D)], list[attractors])."""
        x, lens = self._pack(src, ilens)
        y, emb, att = self.native().forward(x, lens, max_nspks, want_emb=True, want_att=True)
        return ([o[:l] for o, l in zip(y, lens)], [e[:l] for e, l in zip(emb, lens)],
                [a[:l] for a, l in zip(att, lens)])

    @torch.no_grad()
    def test_logits(self, src, ilens, max_nspks=6):
        x, lens = self._pack(src, ilens)
        y, _, _ = self.native().forward(x, lens, max_nspks)
        return [o[:l] for o, l in zip(y, lens)]

    def forward(self, src, tgt, ilens):
        raise NotImplementedError("LS-EEND training forward (masked emb-consistency loss, reference :74-122) and "
                                  "backward are SURVEY §8(f) N1/N2 — round 2")
