"""Drop-in for the reference's FS-EEND/nnet/model/streaming_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm.py:
frame-by-frame FS-EEND (StreamingTransformerEDADiarization.test, reference :31-60) on the sm_100a kernels.

The reference keeps its caches in module attributes (grow-by-torch.cat of layer *inputs*, re-projected each step);
here the state is a native ``fseend_fs_stream`` (projected K/V caches, encoder history for the look-ahead conv) created
lazily at the first ``test()`` call and bound to that call's batch size and ``max_nspks``; ``reset()`` drops it
(the reference needs a new model instance per recording)."""
import torch
import torch.nn as nn
from torch import Tensor

from fseend_b200.native import NativeCacheMixin

from ..modules.streaming_tfm import StreamingAttractorDecoder, StreamingConv1d, StreamingEmbeddingEncoder


def streaming_to_masked_key(k: str) -> str:
    """Name of the masked model's tensor that a streaming-model tensor corresponds to (the inverse of the
    reference's copy table, FS-EEND/nnet/utils/copy_params.py:7-62)."""
    k = k.replace("enc.proj_norm.", "enc.encoder_norm.").replace("enc.proj.", "enc.encoder.")
    k = k.replace("enc.layers.", "enc.transformer_encoder.layers.").replace("self_attn.attention.", "self_attn.")
    k = k.replace("cnn.conv.", "cnn.")
    if k.startswith("dec.layers."):
        k = k.replace("dec.layers.", "dec.attractor_decoder.layers.")
        k = k.replace("temp_attn.attention.", "self_attn1.").replace("spk_attn.", "self_attn2.")
        k = k.replace(".norm1.", ".norm11.").replace(".norm2.", ".norm21.").replace(".norm3.", ".norm22.")
    return k


class StreamingTransformerEDADiarization(NativeCacheMixin, nn.Module):
    def __init__(self, in_size, n_units, n_heads, enc_n_layers, dec_n_layers, dropout, has_mask, max_seqlen,
                 dec_dim_feedforward, conv_delay=9, mask_delay=0, decom_kernel_size=64):
        super().__init__()
        self.delay = conv_delay
        self.n_units = n_units
        self.n_heads = n_heads
        self.enc = StreamingEmbeddingEncoder(in_size, n_units, n_heads, enc_n_layers,
                                             dim_feedforward=dec_dim_feedforward, dropout=dropout)
        self.cnn = StreamingConv1d(n_units, n_units, kernel_size=2 * conv_delay + 1)
        self.dec = StreamingAttractorDecoder(n_units, n_heads, dec_n_layers, dim_feedforward=dec_dim_feedforward,
                                             dropout=dropout)
        self._native = None
        self._native_key = None
        self._stream = None

    def _native_cfg(self):
        return dict(in_size=self.enc.in_size, n_units=self.n_units, n_heads=self.n_heads,
                    enc_n_layers=len(self.enc.layers), dec_n_layers=len(self.dec.layers),
                    enc_dim_feedforward=self.enc.dim_feedforward, dec_dim_feedforward=self.dec.dim_feedforward,
                    conv_kernel=self.cnn.kernel_size, conv_padding=self.cnn.center, mask_delay=0, has_mask=True,
                    bn_eps=self.enc.bn.eps, ln_eps=self.enc.proj_norm.eps)

    def native(self):
        from fseend_b200.native import FsModel
        return self._native_cached(
            lambda: FsModel(self._native_cfg(), {streaming_to_masked_key(k): v for k, v in self.state_dict().items()}))

    def reset(self):
        """Forget the streaming state (start a new recording)."""
        self._stream = None
        self.cnn.t = 0

    @torch.no_grad()
    def test(self, x_t: Tensor, max_nspks: int = 6, dummy_conv_input=False):
        """x_t: (B, 1, in_size) features of frame t.  Returns (B, 1, max_nspks) logits of frame t - conv_delay, or None
        for the first conv_delay calls.  With dummy_conv_input=True the encoder is skipped and a zero embedding is
        pushed into the look-ahead conv (the reference's end-of-recording flush)."""
        from fseend_b200.native import FsStream
        dev = self.cnn.conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("fseend_b200 runs on a CUDA sm_100 device only (move the model with .cuda())")
        B = x_t.shape[0]
        if self._stream is None:
            # the parameter check (native()) walks every tensor of the model: done when a recording starts, not once per
            # 100-ms frame — weights edited mid-recording take effect after reset() / invalidate_native()
            with self._on_device():
                self._stream = FsStream(self.native(), B, max_nspks)
            self.cnn.t = 0
        if self._stream.B != B or self._stream.S != max_nspks:
            raise ValueError("batch size / max_nspks changed mid-stream; call reset() first")
        with torch.cuda.device(dev):
            if dummy_conv_input:
                y = self._stream.step(None)
            else:
                assert x_t.shape[1] == 1, "Input should be a single time frame"
                y = self._stream.step(x_t[:, 0].to(device=dev, dtype=torch.float32).contiguous())
        self.cnn.t += 1
        return None if y is None else y.unsqueeze(1)
