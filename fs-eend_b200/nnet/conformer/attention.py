"""MultiHeadedSelfRetentionModule container (LS-EEND/nnet/conformer/attention.py:71-117): pre-LayerNorm + retention."""
import torch.nn as nn

from ..modules.retention import MultiScaleRetention, RetNetRelPos
from .modules import _no_forward


class MultiHeadedSelfRetentionModule(nn.Module):
    def __init__(self, d_model: int, num_heads: int, recurrent_chunk_size: int = 500, dropout_p: float = 0.1):
        super().__init__()
        self.layer_norm = nn.LayerNorm(d_model)
        self.ret_pos = RetNetRelPos(d_model, num_heads, recurrent_chunk_size=recurrent_chunk_size)
        self.self_attn = MultiScaleRetention(d_model, num_heads, value_factor=1)
        self.dropout = nn.Dropout(p=dropout_p)

    forward = _no_forward
