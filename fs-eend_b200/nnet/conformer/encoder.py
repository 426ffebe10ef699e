"""ConformerEncoderBlock / ConformerEncoder containers (LS-EEND/nnet/conformer/encoder.py:33-123, 126-228).
Block.sequential = [Residual(FFN, 0.5), Residual(retention), Residual(conv module), Residual(FFN, 0.5), LayerNorm]."""
import torch.nn as nn

from .attention import MultiHeadedSelfRetentionModule
from .convolution import ConformerConvModule
from .feed_forward import FeedForwardModule
from .modules import Linear, ResidualConnectionModule, _no_forward


class ConformerEncoderBlock(nn.Module):
    def __init__(self, encoder_dim=512, num_attention_heads=8, feed_forward_expansion_factor=4,
                 conv_expansion_factor=2, feed_forward_dropout_p=0.1, attention_dropout_p=0.1, conv_dropout_p=0.1,
                 conv_kernel_size=31, half_step_residual=True, recurrent_chunk_size=500):
        super().__init__()
        self.feed_forward_residual_factor = 0.5 if half_step_residual else 1
        ff = lambda: FeedForwardModule(encoder_dim, feed_forward_expansion_factor, feed_forward_dropout_p)
        self.sequential = nn.Sequential(
            ResidualConnectionModule(ff(), module_factor=self.feed_forward_residual_factor),
            ResidualConnectionModule(MultiHeadedSelfRetentionModule(
                encoder_dim, num_attention_heads, recurrent_chunk_size, attention_dropout_p)),
            ResidualConnectionModule(ConformerConvModule(encoder_dim, conv_kernel_size, conv_expansion_factor,
                                                         conv_dropout_p)),
            ResidualConnectionModule(ff(), module_factor=self.feed_forward_residual_factor),
            nn.LayerNorm(encoder_dim),
        )

    forward = _no_forward


class ConformerEncoder(nn.Module):
    def __init__(self, input_dim=80, encoder_dim=512, num_layers=17, num_attention_heads=8,
                 feed_forward_expansion_factor=4, conv_expansion_factor=2, feed_forward_dropout_p=0.1,
                 attention_dropout_p=0.1, conv_dropout_p=0.1, conv_kernel_size=31, half_step_residual=True,
                 recurrent_chunk_size=500):
        super().__init__()
        self._conv_kernel_size = conv_kernel_size
        self.input_projection = Linear(input_dim, encoder_dim)
        self.layer_norm = nn.LayerNorm(encoder_dim)
        self.layers = nn.ModuleList([ConformerEncoderBlock(
            encoder_dim, num_attention_heads, feed_forward_expansion_factor, conv_expansion_factor,
            feed_forward_dropout_p, attention_dropout_p, conv_dropout_p, conv_kernel_size, half_step_residual,
            recurrent_chunk_size) for _ in range(num_layers)])

    forward = _no_forward
