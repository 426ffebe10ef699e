"""FeedForwardModule container (LS-EEND/nnet/conformer/feed_forward.py:23-57): sequential = [LayerNorm, Linear(D, D*f),
Swish, Dropout, Linear(D*f, D), Dropout] — indices 0, 1, 4 carry parameters."""
import torch.nn as nn

from .modules import Linear, Swish, _no_forward


class FeedForwardModule(nn.Module):
    def __init__(self, encoder_dim: int = 512, expansion_factor: int = 4, dropout_p: float = 0.1):
        super().__init__()
        self.sequential = nn.Sequential(
            nn.LayerNorm(encoder_dim),
            Linear(encoder_dim, encoder_dim * expansion_factor, bias=True),
            Swish(),
            nn.Dropout(p=dropout_p),
            Linear(encoder_dim * expansion_factor, encoder_dim, bias=True),
            nn.Dropout(p=dropout_p),
        )

    forward = _no_forward
