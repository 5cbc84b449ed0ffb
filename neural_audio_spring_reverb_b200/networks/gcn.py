"""GCN (gated convolutional network) - drop-in for the reference class
(src/neural_audio_spring_reverb/networks/gcn.py:8-160): same constructor,
attributes and ``state_dict`` keys; ``forward`` runs on libnasr_b200."""
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from .. import _native
from ._fused import FusedNetMixin
from .custom_layers import Conv1dCausal, FiLM, GatedAF, TanhAF


class GCNBlock(nn.Module):
    """conv (-> 2C) -> FiLM -> tanh*sigmoid gate -> + res 1x1 (gcn.py:8-61)."""

    def __init__(self, in_ch: int, out_ch: int, kernel_size: int = 3, dilation: int = 1,
                 stride: int = 1, cond_dim: int = 0) -> None:
        super().__init__()
        self.in_ch = in_ch
        self.out_ch = out_ch
        self.kernel_size = kernel_size
        self.dilation = dilation
        self.stride = stride
        self.cond_dim = cond_dim
        self.conv = Conv1dCausal(in_channels=in_ch, out_channels=out_ch * 2, kernel_size=kernel_size,
                                 stride=stride, dilation=dilation)
        self.film = FiLM(cond_dim=cond_dim, n_features=out_ch * 2)
        self.gated_activation = GatedAF()
        self.res = nn.Conv1d(in_channels=in_ch, out_channels=out_ch, kernel_size=(1,), bias=False)

    def forward(self, x: Tensor, cond: Tensor) -> Tensor:
        raise RuntimeError("GCNBlock.forward: use GCN.block_forward(index, x, cond) - the block runs "
                           "as one fused kernel owned by the parent network's engine")


class GCN(FusedNetMixin, nn.Module):
    """Gated convolutional network with FiLM conditioning (gcn.py:64-160)."""

    _nasr_arch = _native.ARCH_GCN
    _nasr_final_tanh = True

    def __init__(self, in_ch: int = 1, out_ch: int = 1, n_blocks: int = 2, n_channels: int = 32,
                 dilation_growth: int = 8, kernel_size: int = 3, cond_dim: int = 3) -> None:
        super().__init__()
        print(f"Dilations: {[dilation_growth**i for i in range(n_blocks)]}")  # gcn.py:99
        self._nasr_build(GCNBlock, n_blocks, n_channels, dilation_growth, in_ch, out_ch,
                         kernel_size, cond_dim)
        self.af = TanhAF()

    def forward(self, x: Tensor, cond: Optional[Tensor] = None) -> Tensor:
        """x [B, in_ch, T], cond [B, cond_dim] -> tanh(out_net(blocks(x))) (gcn.py:140-147)."""
        return self._nasr_run(x, cond, chunk=False)
