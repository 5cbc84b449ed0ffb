"""TCN with FiLM conditioning - drop-in for the reference class
(src/neural_audio_spring_reverb/networks/tcn.py:36-164): same constructor,
attributes and ``state_dict`` keys; ``forward`` runs on libnasr_b200."""
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from .. import _native
from ._fused import FusedNetMixin
from .custom_layers import Conv1dCausal, FiLM


class TCNBlock(nn.Module):
    """conv -> FiLM (if cond_dim > 0) -> PReLU -> + res 1x1 (tcn.py:36-86)."""

    def __init__(self, in_ch: int, out_ch: int, kernel_size: int = 3, dilation: int = 1,
                 stride: int = 1, cond_dim: int = 0, activation=True):
        super().__init__()
        if not activation:
            raise ValueError("activation=False is not used by the reference TCN and is unsupported")
        self.in_ch = in_ch
        self.out_ch = out_ch
        self.kernel_size = kernel_size
        self.dilation = dilation
        self.stride = stride
        self.cond_dim = cond_dim
        self.conv = Conv1dCausal(in_channels=in_ch, out_channels=out_ch, kernel_size=kernel_size,
                                 stride=stride, dilation=dilation)
        if cond_dim > 0:
            self.film = FiLM(cond_dim=cond_dim, n_features=out_ch, batch_norm=True)
        if activation:
            self.act = nn.PReLU()
        self.res = nn.Conv1d(in_channels=in_ch, out_channels=out_ch, kernel_size=(1,), bias=False)

    def forward(self, x: Tensor, cond: Tensor) -> Tensor:
        raise RuntimeError("TCNBlock.forward: use TCN.block_forward(index, x, cond) - the block runs "
                           "as one fused kernel owned by the parent network's engine")


class TCN(FusedNetMixin, nn.Module):
    """Temporal convolutional network with conditioning module (tcn.py:89-164)."""

    _nasr_arch = _native.ARCH_TCN
    _nasr_final_tanh = False

    def __init__(self, n_channels: int, n_layers: int, dilation_growth: int, in_ch: int = 1,
                 out_ch: int = 1, kernel_size: int = 3, cond_dim: int = 0):
        super().__init__()
        self._nasr_build(TCNBlock, n_layers, n_channels, dilation_growth, in_ch, out_ch,
                         kernel_size, cond_dim)

    def forward(self, x: Tensor, cond: Optional[Tensor] = None) -> Tensor:
        """x [B, in_ch, T] fp32, cond [B, cond_dim] | None -> [B, out_ch, T] (tcn.py:150-155)."""
        return self._nasr_run(x, cond, chunk=False)
