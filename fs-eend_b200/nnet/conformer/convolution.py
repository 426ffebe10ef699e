"""Conformer convolution-module containers (LS-EEND/nnet/conformer/convolution.py: DepthwiseConv1d :25-68,
PointwiseConv1d :71-108, ConformerConvModule :111-167).  sequential = [LayerNorm, Transpose, pointwise D->2D, GLU,
causal depthwise (k, no bias), BatchNorm1d, Swish, pointwise D->D, Dropout]."""
import torch.nn as nn

from .modules import GLU, Swish, Transpose, _no_forward


class DepthwiseConv1d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=False):
        super().__init__()
        assert out_channels % in_channels == 0
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, groups=in_channels, stride=stride,
                              padding=padding, bias=bias)

    forward = _no_forward


class PointwiseConv1d(nn.Module):
    def __init__(self, in_channels, out_channels, stride=1, padding=0, bias=True):
        super().__init__()
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=1, stride=stride, padding=padding, bias=bias)

    forward = _no_forward


class ConformerConvModule(nn.Module):
    def __init__(self, in_channels, kernel_size=31, expansion_factor=2, dropout_p=0.1):
        super().__init__()
        assert expansion_factor == 2, "Currently, Only Supports expansion_factor 2"
        self.kernel_size = kernel_size
        self.sequential = nn.Sequential(
            nn.LayerNorm(in_channels),
            Transpose(shape=(1, 2)),
            PointwiseConv1d(in_channels, in_channels * expansion_factor, stride=1, padding=0, bias=True),
            GLU(dim=1),
            DepthwiseConv1d(in_channels, in_channels, kernel_size, stride=1, padding=kernel_size - 1),
            nn.BatchNorm1d(in_channels),
            Swish(),
            PointwiseConv1d(in_channels, in_channels, stride=1, padding=0, bias=True),
            nn.Dropout(p=dropout_p),
        )

    forward = _no_forward
