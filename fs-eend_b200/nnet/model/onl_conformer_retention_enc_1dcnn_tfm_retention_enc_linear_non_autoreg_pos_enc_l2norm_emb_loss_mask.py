"""Drop-in for the reference's LS-EEND model file (same name under LS-EEND/nnet/model/): class names, constructor
kwargs, attributes (.n_units .delay .enc.encoder.layers .enc.encoder._conv_kernel_size .dec.layers .cnn) and
state_dict keys match (tests/golden/ls_state_dict_abi.txt, dumped from the real reference).  The arithmetic runs in
the sm_100a library through the C ABI (fseend_ls_*); there is no CPU path."""
import math

import torch
import torch.nn as nn
from torch import Tensor

from fseend_b200.native import NativeCacheMixin

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


def _stream_for(owner, states, B, max_nspks):
    """The native one-step state bound to the caller-owned state list (the reference keeps its recurrent state in
    these dicts/tensors, LS-EEND/streaming_infer_dia.py:37-49; here they carry a handle to the device state)."""
    from fseend_b200.native import LsStream
    holder = states[0] if isinstance(states, (list, tuple)) and states and isinstance(states[0], dict) else None
    key = "_fseend_stream"
    st = holder.get(key) if holder is not None else getattr(owner, "_default_stream", None)
    native = owner.native()
    if st is None or st.model is not native or st.B != B or st.S != max_nspks:
        st = LsStream(native, B, max_nspks)
        if holder is not None:
            holder[key] = st
        else:
            owner._default_stream = st
    return st


class StreamingConv1d(nn.Module):
    """Frame-by-frame look-ahead Conv1d (reference model file :151-186): ``.conv`` holds the weights (loaded from
    model.cnn by the caller), ``.buffer``/``.t`` the window state.  The 19-tap window product runs as one tcgen05
    GEMM (K = 19*256) through the C ABI."""

    def __init__(self, in_channels, out_channels, kernel_size=19, precision="fp32"):
        super().__init__()
        from collections import deque
        self.precision = precision          # "fp32": split-precision GEMM (parity mode); "fp16": fp16 operands
        self.kernel_size = kernel_size
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, padding=0)
        self.buffer = deque(maxlen=kernel_size)
        self.center = kernel_size // 2
        self.t = 0
        self._w16 = None
        self._w_key = None

    @torch.no_grad()
    def forward(self, x_t):
        """x_t: (B, C, 1) -> (B, C_out, 1), or None for the first ``center`` frames."""
        from fseend_b200 import native as N
        self.t += 1
        self.buffer.append(x_t)
        if self.t < self.center + 1:
            return None
        pad = [torch.zeros_like(x_t)] * (self.kernel_size - len(self.buffer))
        win = torch.cat(pad + list(self.buffer), dim=2)                       # (B, C, K)
        w = self.conv.weight
        key = (w.data_ptr(), w._version, self.precision)
        if self._w16 is None or self._w_key != key:
            # out[b, co] = sum_{k, ci} W[co, ci, k] * win[b, ci, k]  ->  A [B, K*C] (k-major), W' [C_out, K*C]
            wk = w.detach().permute(0, 2, 1).reshape(w.shape[0], -1)
            self._w16 = N.P32Linear(wk) if self.precision == "fp32" else wk.to(torch.float16).contiguous()
            self._w_key = key
        bias = self.conv.bias.detach().float().contiguous()
        if self.precision == "fp32":
            a = win.permute(0, 2, 1).reshape(win.shape[0], -1).float().contiguous()
            return self._w16(a, bias=bias).unsqueeze(-1)
        a = win.permute(0, 2, 1).reshape(win.shape[0], -1).to(torch.float16).contiguous()
        y = N.op_gemm(a, self._w16, N.EPI_BIAS, bias=bias)
        return y.float().unsqueeze(-1)


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
        self._owner = None

    @torch.no_grad()
    def forward_one_step(self, x_t: Tensor, t: int, ret_states: list, conv_caches: list) -> Tensor:
        """Reference :291-293 -> conformer/encoder.py:223-228.  x_t: (B, 1, in_size) -> (B, 1, n_units).  The
        recurrent state lives on the device, bound to ``ret_states``; ``conv_caches`` is accepted for signature
        compatibility (its contents are not used)."""
        owner = self._owner()
        st = _stream_for(owner, ret_states, x_t.shape[0], getattr(owner, "_one_step_nspks", 4))
        dev = owner.cnn.weight.device
        return st.enc_step(x_t[:, 0].to(device=dev, dtype=torch.float32).contiguous(), t).unsqueeze(1)


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
        self._owner = None

    @torch.no_grad()
    def forward_one_step(self, emb_t: Tensor, t: int, max_nspks: int, ret_states: list) -> Tensor:
        """Reference :235-243.  emb_t: (B, 1, D) conv'ed + L2-normalised embedding -> (B, 1, max_nspks, D)
        un-normalised attractors of decoder frame t."""
        owner = self._owner()
        st = _stream_for(owner, ret_states, emb_t.shape[0], max_nspks)
        dev = owner.cnn.weight.device
        att = st.dec_step(emb_t[:, 0].to(device=dev, dtype=torch.float32).contiguous(), t)
        return att.unsqueeze(1)


class OnlineConformerRetentionDADiarization(NativeCacheMixin, nn.Module):
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
        self._precision = None
        import weakref
        self.enc._owner = weakref.ref(self)
        self.dec._owner = weakref.ref(self)

    def new_stream(self, batch_size: int = 1, max_nspks: int = 6):
        """Fused one-step driver (encoder + look-ahead conv + decoder + head per call): ``stream.step(x_t)`` with
        x_t (B, in_size) returns (B, max_nspks) logits of frame t - conv_delay or None; ``stream.step(None)`` flushes."""
        from fseend_b200.native import LsStream
        with self._on_device():
            return LsStream(self.native(), batch_size, max_nspks)

    def _native_cfg(self):
        return dict(in_size=self.enc.in_size, n_units=self.n_units, n_heads=self.enc.n_heads,
                    enc_n_layers=self.enc.n_layers, dec_n_layers=self.dec.n_layers,
                    feed_forward_expansion_factor=self.enc.feed_forward_expansion_factor,
                    dec_dim_feedforward=self.dec.dim_feedforward,
                    conv_kernel_size=self.enc.encoder._conv_kernel_size,
                    recurrent_chunk_size=self.recurrent_chunk_size, conv_delay=self.delay)

    def native(self):
        from fseend_b200.native import LsModel

        def build():
            nm = LsModel(self._native_cfg(), self.state_dict())
            if self._precision is not None:
                nm.set_precision(self._precision)
            return nm
        return self._native_cached(build)

    def set_precision(self, mode):
        """"fp32" (default; alias "parity"): fp32 activations + split-precision tcgen05 GEMMs — logits within 1e-3 of the
        reference on every frame.  "fp16": fp16-operand throughput mode (≈4x faster; median error 2-3e-4, isolated
        frames up to 7e-2: LS-EEND's eps = 1e-6 group norm amplifies operand rounding).  None: library default
        (FSEEND_LS_PRECISION).  Streams created afterwards inherit the mode."""
        if mode not in (None, "fp16", "fp32", "parity"):
            raise ValueError("precision must be 'fp16', 'fp32' / 'parity' or None")
        self._precision = mode
        if self._native is not None and mode is not None:
            self._native.set_precision(mode)

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
        with self._on_device():
            y, emb, att = self.native().forward(x, lens, max_nspks, want_emb=True, want_att=True)
        return ([o[:l] for o, l in zip(y, lens)], [e[:l] for e, l in zip(emb, lens)],
                [a[:l] for a, l in zip(att, lens)])

    @torch.no_grad()
    def test_logits(self, src, ilens, max_nspks=6):
        x, lens = self._pack(src, ilens)
        with self._on_device():
            y, _, _ = self.native().forward(x, lens, max_nspks)
        return [o[:l] for o, l in zip(y, lens)]

    def forward(self, src, tgt, ilens):
        """Reference :74-122 (eval-mode arithmetic): logits, the length-masked embedding-consistency loss (one kernel,
        csrc/embloss.cu), embeddings, attractors[1:n_spk].  Gradients are not produced (backward = SURVEY §8(f) N1)."""
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "fseend_b200 round 1 implements the forward hot path only; training backward is SURVEY §8(f) N1")
        from fseend_b200.native import op_embloss
        with torch.no_grad(), self._on_device():
            n_speakers = [t.shape[1] for t in tgt]
            max_nspks = max(n_speakers)
            x, lens = self._pack(src, ilens)
            y, emb, att = self.native().forward(x, lens, max_nspks, want_emb=True, want_att=True)
            seq_len = max(lens)
            dev = y.device
            labels = torch.zeros(len(tgt), seq_len, max_nspks, device=dev, dtype=torch.float32)
            for b, t in enumerate(tgt):
                n = min(t.shape[0], seq_len)
                labels[b, :n, :t.shape[1]] = t[:n].to(device=dev, dtype=torch.float32)
            lens_dev = torch.tensor(lens, device=dev, dtype=torch.int32)
            emb_consis_loss = op_embloss(emb[:, :seq_len].contiguous(), labels, seq_len=lens_dev,
                                         divisor=float(sum(l * l for l in lens)))
            output = [o[:l, :n] for o, l, n in zip(y, lens, n_speakers)]
            emb = [e[:l] for e, l in zip(emb, lens)]
            attractors = [a[:l, 1:n] for a, l, n in zip(att, lens, n_speakers)]
        return output, emb_consis_loss, emb, attractors
