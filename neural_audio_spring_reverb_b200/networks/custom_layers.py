"""Parameter containers for the layers the fused kernels replace.

Same class names, constructor signatures and ``state_dict`` keys as the
reference (src/neural_audio_spring_reverb/networks/custom_layers.py:8-126), so
``load_state_dict(strict=True)`` of its checkpoints works.  The arithmetic of
``Conv1dCausal`` / ``FiLM`` / ``GatedAF`` / ``TanhAF`` is not executed layer by
layer: TCN/GCN.forward hands the whole block to one CUDA kernel
(csrc/generic_block.cu, csrc/tc_block.cu).  Calling a layer on its own is not
part of the accelerated path and raises.
"""
import torch.nn as nn
from torch import Tensor

_MSG = ("{name}.forward is fused into the block kernel of libnasr_b200; call TCN/GCN.forward "
        "(or the block's forward) instead - there is no per-layer / CPU path")


class FiLM(nn.Module):
    """Feature-wise linear modulation with eval-mode BatchNorm (custom_layers.py:8-42)."""

    def __init__(self, cond_dim: int, n_features: int, batch_norm: bool = True) -> None:
        super().__init__()
        self.num_features = n_features
        self.adaptor = nn.Linear(cond_dim, n_features * 2)
        if batch_norm is True:
            self.bn = nn.BatchNorm1d(n_features)

    def forward(self, x: Tensor, cond: Tensor) -> Tensor:
        raise RuntimeError(_MSG.format(name="FiLM"))


class Conv1dCausal(nn.Module):
    """Causal dilated Conv1d: left zero pad (k-1)*d, then Conv1d (custom_layers.py:45-88)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int,
                 dilation: int = 1, bias: bool = True) -> None:
        super().__init__()
        if stride != 1:
            raise ValueError("only stride 1 is supported (the reference never uses another)")
        self.padding = (kernel_size - 1) * dilation
        self.in_channels = in_channels
        self.conv = nn.Conv1d(in_channels, out_channels, (kernel_size,), (stride,), padding=0,
                              dilation=(dilation,), bias=bias)

    def forward(self, x: Tensor) -> Tensor:
        raise RuntimeError(_MSG.format(name="Conv1dCausal"))


class GatedAF(nn.Module):
    """tanh(x[:, :C]) * sigmoid(x[:, C:]) (custom_layers.py:91-111); fused in the GCN epilogue."""

    def forward(self, x: Tensor) -> Tensor:
        raise RuntimeError(_MSG.format(name="GatedAF"))


class TanhAF(nn.Module):
    """Final tanh of GCN (custom_layers.py:114-126); fused after out_net."""

    def forward(self, x: Tensor) -> Tensor:
        raise RuntimeError(_MSG.format(name="TanhAF"))
