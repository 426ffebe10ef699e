"""Parameter containers of the frame-by-frame FS-EEND model (API mirror of the reference's
FS-EEND/nnet/modules/streaming_tfm.py: IncrementalSelfAttention :10-37, StreamingTransformerEncoderLayer :39-79,
StreamingEmbeddingEncoder :82-129, StreamingConv1d :132-167, StreamingAttractorDecoderLayer :170-227,
StreamingAttractorDecoder :230-269).  They hold the parameters under the reference's names; the per-frame
arithmetic and the caches live in the native stream object (fseend_fs_stream)."""
import math

import torch
from torch import nn


def _container_forward(*a, **k):
    raise RuntimeError("fseend_b200 streaming modules are parameter containers; call the model's test()")


class IncrementalSelfAttention(nn.Module):
    def __init__(self, d_model, nhead):
        super().__init__()
        self.attention = nn.MultiheadAttention(embed_dim=d_model, num_heads=nhead, batch_first=True)

    forward = _container_forward


class StreamingTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation=None):
        super().__init__()
        self.self_attn = IncrementalSelfAttention(d_model, nhead)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)

    forward = _container_forward


class StreamingEmbeddingEncoder(nn.Module):
    def __init__(self, in_size, d_model, nhead, num_layers, dim_feedforward=2048, dropout=0.1, activation=None):
        super().__init__()
        self.in_size, self.d_model, self.nhead, self.dim_feedforward = in_size, d_model, nhead, dim_feedforward
        self.bn = nn.BatchNorm1d(in_size)
        self.proj = nn.Linear(in_size, d_model)
        self.proj_norm = nn.LayerNorm(d_model)
        self.layers = nn.ModuleList([
            StreamingTransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=dim_feedforward,
                                             dropout=dropout) for _ in range(num_layers)])
        self.cache = [{} for _ in range(num_layers)]   # kept for API shape; the K/V caches are device-resident
        self.init_weights()

    def init_weights(self):
        initrange = 0.1
        self.proj.bias.data.zero_()
        self.proj.weight.data.uniform_(-initrange, initrange)

    forward = _container_forward


class StreamingConv1d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=19):
        super().__init__()
        self.kernel_size = kernel_size
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, padding=0)
        self.center = kernel_size // 2
        self.t = 0          # frames pushed (mirrors the reference attribute; the ring buffer itself is on the device)
        self.buffer = None

    forward = _container_forward


class StreamingAttractorDecoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation=None):
        super().__init__()
        self.temp_attn = IncrementalSelfAttention(d_model, nhead)
        self.spk_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)

    forward = _container_forward


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


class StreamingAttractorDecoder(nn.Module):
    def __init__(self, d_model, nhead, num_layers, dim_feedforward=2048, dropout=0.1, activation=None):
        super().__init__()
        self.dim_feedforward = dim_feedforward
        self.pos_enc = PositionalEncoding(d_model, dropout)
        self.convert = nn.Linear(2 * d_model, d_model)
        self.layers = nn.ModuleList([
            StreamingAttractorDecoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=dim_feedforward,
                                           dropout=dropout) for _ in range(num_layers)])
        self.cache = [{} for _ in range(num_layers)]

    forward = _container_forward
