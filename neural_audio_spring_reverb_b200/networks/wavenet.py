"""WaveNet - drop-in for the reference class
(src/neural_audio_spring_reverb/networks/wavenet.py:11-216).  Its `Conv1dStack` is
arithmetically the GCN block (wavenet.py:53-60 vs gcn.py:53-61); only the dilation
schedule differs (dilation_growth**s restarting in every block, wavenet.py:94), so the
forward runs on the same fused kernels with an explicit dilation list."""
from typing import Optional

import torch.nn as nn
from torch import Tensor

from .. import _native
from ._fused import FusedNetMixin
from .custom_layers import Conv1dCausal, FiLM, GatedAF, TanhAF


class Conv1dStack(nn.Module):
    """conv (-> 2C) -> FiLM -> tanh*sigmoid gate -> + res 1x1 (wavenet.py:11-60)."""

    def __init__(self, in_ch: int, out_ch: int, kernel_size: int, dilation: int, cond_dim: int) -> None:
        super().__init__()
        self.in_ch, self.out_ch, self.dilation = in_ch, out_ch, dilation
        self.conv = Conv1dCausal(in_channels=in_ch, out_channels=out_ch * 2, kernel_size=kernel_size,
                                 stride=1, dilation=dilation)
        self.film = FiLM(cond_dim=cond_dim, n_features=out_ch * 2)
        self.gated_activation = GatedAF()
        self.res = nn.Conv1d(in_channels=in_ch, out_channels=out_ch, kernel_size=(1,), bias=False)

    def forward(self, x: Tensor, cond: Tensor) -> Tensor:
        raise RuntimeError("Conv1dStack.forward: use WaveNet.block_forward(index, x, cond)")


class WaveNet1dBlock(nn.Module):
    """n_stacks Conv1dStacks with dilations dilation_growth**s (wavenet.py:63-116)."""

    def __init__(self, in_ch: int, out_ch: int, n_stacks: int, kernel_size: int, dilation_growth: int,
                 cond_dim: int) -> None:
        super().__init__()
        self.in_ch, self.out_ch, self.n_stacks = in_ch, out_ch, n_stacks
        self.kernel_size, self.dilation_growth, self.cond_dim = kernel_size, dilation_growth, cond_dim
        widths = [in_ch] + [out_ch] * n_stacks
        self.stacks = nn.ModuleList(
            Conv1dStack(widths[s], out_ch, kernel_size, dilation_growth**s, cond_dim) for s in range(n_stacks))

    def forward(self, x: Tensor, cond: Tensor) -> Tensor:
        raise RuntimeError("WaveNet1dBlock.forward: the stacks run as fused kernels of the parent WaveNet")


class WaveNet(FusedNetMixin, nn.Module):
    """wavenet.py:119-216: n_blocks x n_stacks gated stacks, out_net 1x1, tanh."""

    _nasr_arch = _native.ARCH_GCN
    _nasr_final_tanh = True

    def __init__(self, in_ch: int = 1, out_ch: int = 1, n_blocks: int = 2, n_stacks: int = 2,
                 n_channels: int = 32, kernel_size: int = 3, dilation_growth: int = 8, cond_dim: int = 3) -> None:
        super().__init__()
        self.in_ch, self.out_ch = in_ch, out_ch
        self.n_blocks, self.n_stacks, self.n_channels = n_blocks, n_stacks, n_channels
        self.kernel_size, self.dilation_growth, self.cond_dim = kernel_size, dilation_growth, cond_dim
        self.blocks = nn.ModuleList(
            WaveNet1dBlock(in_ch if b == 0 else n_channels, n_channels, n_stacks, kernel_size, dilation_growth, cond_dim)
            for b in range(n_blocks))
        self.out_net = nn.Conv1d(in_channels=n_channels, out_channels=out_ch, kernel_size=1, stride=1, padding=0,
                                 bias=False)
        self.af = TanhAF()

    # the engine sees one flat list of fused layers
    def _nasr_layers(self):
        return [st for blk in self.blocks for st in blk.stacks]

    def forward(self, x: Tensor, cond: Optional[Tensor] = None) -> Tensor:
        """x [B, in_ch, T], cond [B, cond_dim] -> tanh(out_net(stacks(x))) (wavenet.py:180-191)."""
        return self._nasr_run(x, cond, chunk=False)

    def calc_receptive_field(self) -> int:
        """1 + sum over all stacks of (k - 1) * dilation (wavenet.py:193-216)."""
        return 1 + sum((self.kernel_size - 1) * self.dilation_growth**s
                       for _ in range(self.n_blocks) for s in range(self.n_stacks))
