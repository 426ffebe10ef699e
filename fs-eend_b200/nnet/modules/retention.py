"""Retention containers (LS-EEND/nnet/modules/retention.py: RetNetRelPos :13-59, MultiScaleRetention :82-228).
LS-EEND fixes the decay to 1 (buffer ``decay`` = log 1 = 0, :20) and disables the rotation (:209-213): retention is
causal linear attention with the chunk normalisation implemented in csrc/retention.cu."""
import torch
import torch.nn as nn


class RetNetRelPos(nn.Module):
    def __init__(self, embed_dim: int, num_heads: int, recurrent_chunk_size: int):
        super().__init__()
        angle = 1.0 / (10000 ** torch.linspace(0, 1, embed_dim // num_heads // 2))
        self.register_buffer("angle", angle.unsqueeze(-1).repeat(1, 2).flatten())
        self.register_buffer("decay", torch.log(torch.ones(num_heads)))
        self.recurrent_chunk_size = recurrent_chunk_size


class MultiScaleRetention(nn.Module):
    def __init__(self, embed_dim: int, num_heads: int, value_factor: int = 2, gate_fn: str = "swish"):
        super().__init__()
        if value_factor != 1 or gate_fn != "swish":
            raise NotImplementedError("LS-EEND uses value_factor=1 with a swish gate")
        self.factor, self.embed_dim, self.num_heads = value_factor, embed_dim, num_heads
        self.head_dim = self.key_dim = embed_dim // num_heads
        self.scaling = self.key_dim ** -0.5
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.k_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.v_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.g_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.group_norm = nn.LayerNorm(self.head_dim, eps=1e-6, elementwise_affine=False)
        self.reset_parameters()

    def reset_parameters(self):
        for proj in (self.q_proj, self.k_proj, self.v_proj, self.g_proj):
            nn.init.xavier_uniform_(proj.weight, gain=2 ** -2.5)
        nn.init.xavier_uniform_(self.out_proj.weight)
        nn.init.constant_(self.out_proj.bias, 0.0)
