"""Parameter containers for the FS-EEND layer stacks (API mirror of the reference's
FS-EEND/nnet/modules/merge_tfm_encoder.py: TransformerEncoder :17-139, TransformerEncoderFusionLayer :142-399).

The reference executes these layers with torch ops; here the modules only *hold* parameters under the
reference's names (state_dict ABI) — the arithmetic runs in the sm_100a kernels driven by
``nnet.model.*.OnlineTransformerDADiarization``.  Construction order matches the reference so a given RNG
seed yields the same initial weights.
"""
import copy

from torch import nn


class TransformerEncoderFusionLayer(nn.Module):
    """Attractor-decoder layer parameters: time-axis MHA (self_attn1), speaker-axis MHA (self_attn2), FFN,
    post-norm LayerNorms norm11 / norm21 / norm22 (norm12 exists but is dead, reference :214)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, layer_norm_eps=1e-5, batch_first=False):
        super().__init__()
        self.self_attn1 = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=batch_first)
        self.self_attn2 = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=batch_first)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm_first = False
        self.norm11 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm12 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm21 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm22 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.dropout11 = nn.Dropout(dropout)
        self.dropout21 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)

    def forward(self, *a, **k):
        raise RuntimeError("fseend_b200 layers are parameter containers; call the model's test()/forward()")


class TransformerEncoder(nn.Module):
    """Stack of ``num_layers`` deep copies of ``encoder_layer`` (all clones start identical, as in the reference)."""

    def __init__(self, encoder_layer, num_layers, norm=None):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = norm

    def forward(self, *a, **k):
        raise RuntimeError("fseend_b200 layers are parameter containers; call the model's test()/forward()")
