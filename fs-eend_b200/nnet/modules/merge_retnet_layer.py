"""LS-EEND attractor-decoder layer container (LS-EEND/nnet/modules/merge_retnet_layer.py:16-312): retention along
time (self_attn1 + ret_pos1), MHA along speakers (self_attn2), ReLU FFN, post-norm LayerNorms norm11/21/22
(norm12 is a dead parameter kept for checkpoint compatibility)."""
import torch.nn as nn

from .retention import MultiScaleRetention, RetNetRelPos


class TransformerEncoderFusionLayer(nn.Module):
    def __init__(self, d_model, nhead, recurrent_chunk_size=500, dim_feedforward=2048, dropout=0.1,
                 layer_norm_eps=1e-5, batch_first=False, norm_first=False):
        super().__init__()
        self.ret_pos1 = RetNetRelPos(d_model, nhead, recurrent_chunk_size=recurrent_chunk_size)
        self.self_attn1 = MultiScaleRetention(d_model, nhead, value_factor=1)
        self.self_attn2 = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=batch_first)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm_first = norm_first
        self.norm11 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm12 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm21 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm22 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.dropout11 = nn.Dropout(dropout)
        self.dropout21 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)

    def forward(self, *a, **k):
        raise RuntimeError("fseend_b200 layers are parameter containers; call the model's test()/forward()")
