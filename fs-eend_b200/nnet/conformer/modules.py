"""Containers mirroring LS-EEND/nnet/conformer/modules.py (ResidualConnectionModule :21-34, Linear :37-49) and the
activation markers of conformer/activation.py — parameter holders only; the arithmetic runs in the sm_100a kernels."""
import torch.nn as nn
import torch.nn.init as init


def _no_forward(self, *a, **k):
    raise RuntimeError("fseend_b200 LS-EEND modules are parameter containers; call the model's test()/forward()")


class Marker(nn.Module):
    """Parameter-free placeholder that keeps nn.Sequential indices identical to the reference's."""
    forward = _no_forward


class Swish(Marker):
    pass


class GLU(Marker):
    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim


class Transpose(Marker):
    def __init__(self, shape: tuple):
        super().__init__()
        self.shape = shape


class ResidualConnectionModule(nn.Module):
    def __init__(self, module: nn.Module, module_factor: float = 1.0, input_factor: float = 1.0):
        super().__init__()
        self.module = module
        self.module_factor = module_factor
        self.input_factor = input_factor

    forward = _no_forward


class Linear(nn.Module):
    """nn.Linear under the attribute name ``linear`` (xavier weight, zero bias — the reference's init)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)
        init.xavier_uniform_(self.linear.weight)
        if bias:
            init.zeros_(self.linear.bias)

    forward = _no_forward
