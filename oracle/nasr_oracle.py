"""ORACLE - test infrastructure, not product code.

CPU restatement of the reference's TCN / GCN eval-mode forward, op for op, on
the same ATen CPU operators the reference's own modules dispatch to
(F.pad, F.conv1d, F.linear, F.batch_norm, F.prelu, tanh, sigmoid).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file; nothing under neural_audio_spring_reverb_b200/ does.

Parity pin: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4/8c), so the oracle is pinned against the reference itself:
tests/golden/make_golden.py imports the reference from /root/reference, runs
its TCN/GCN classes (random-init configs and the 8 shipped checkpoints) and
commits the outputs under tests/golden/; tests/test_oracle.py checks this file
against those vectors.

Each function cites the reference lines it restates (paths relative to
/root/reference/src/neural_audio_spring_reverb/).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import Tensor

StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------- layers
def conv1d_causal(x: Tensor, weight: Tensor, bias: Optional[Tensor], dilation: int,
                  history: Optional[Tensor] = None) -> Tensor:
    """networks/custom_layers.py:85-88 (Conv1dCausal.forward): left pad (k-1)*d zeros,
    then Conv1d(padding=0, dilation=d).  With `history` the pad is the cached
    input tail instead of zeros (wrapper.py:23-30, PaddingCached.forward)."""
    pad = (weight.shape[-1] - 1) * dilation
    if history is None:
        x = F.pad(x, (pad, 0))
    else:
        x = torch.cat([history, x], dim=-1)
    return F.conv1d(x, weight, bias, stride=1, padding=0, dilation=dilation)


def film(x: Tensor, cond: Tensor, sd: StateDict, prefix: str, eps: float = 1e-5) -> Tensor:
    """networks/custom_layers.py:32-42 (FiLM.forward), eval mode: adaptor Linear,
    chunk into (g, b), BatchNorm1d with running stats, x * g + b."""
    h = F.linear(cond, sd[prefix + "adaptor.weight"], sd[prefix + "adaptor.bias"])
    g, b = torch.chunk(h, 2, dim=-1)
    g = g.unsqueeze(-1)
    b = b.unsqueeze(-1)
    x = F.batch_norm(x, sd[prefix + "bn.running_mean"], sd[prefix + "bn.running_var"],
                     sd[prefix + "bn.weight"], sd[prefix + "bn.bias"], training=False, eps=eps)
    return (x * g) + b


def gated_af(x: Tensor) -> Tensor:
    """networks/custom_layers.py:103-111 (GatedAF.forward)."""
    a, b = x.chunk(2, dim=1)
    return torch.tanh(a) * torch.sigmoid(b)


# --------------------------------------------------------------------------- blocks
def tcn_block(x: Tensor, cond: Optional[Tensor], sd: StateDict, i: int, dilation: int,
              history: Optional[Tensor] = None) -> Tensor:
    """networks/tcn.py:73-86 (TCNBlock.forward): conv -> FiLM (if present) -> PReLU ->
    + causal_crop(res(x_in)) (the crop is a no-op: equal lengths, tcn.py:28-33)."""
    p = f"blocks.{i}."
    y = conv1d_causal(x, sd[p + "conv.conv.weight"], sd[p + "conv.conv.bias"], dilation, history)
    if (p + "film.adaptor.weight") in sd:
        y = film(y, cond, sd, p + "film.")
    y = F.prelu(y, sd[p + "act.weight"])
    return y + F.conv1d(x, sd[p + "res.weight"])


def gcn_block(x: Tensor, cond: Tensor, sd: StateDict, i: int, dilation: int,
              history: Optional[Tensor] = None) -> Tensor:
    """networks/gcn.py:53-61 (GCNBlock.forward): conv(2C) -> FiLM -> gate -> + res(x_in)."""
    p = f"blocks.{i}."
    y = conv1d_causal(x, sd[p + "conv.conv.weight"], sd[p + "conv.conv.bias"], dilation, history)
    y = film(y, cond, sd, p + "film.")
    y = gated_af(y)
    return y + F.conv1d(x, sd[p + "res.weight"])


def flatten_wavenet_state(sd: StateDict) -> StateDict:
    """networks/wavenet.py:11-60,94-116: a WaveNet is n_blocks x n_stacks `Conv1dStack`s whose
    forward (wavenet.py:53-60) is line for line GCNBlock.forward (gcn.py:53-61); renaming
    blocks.b.stacks.s.* to blocks.(b*n_stacks+s).* lets the GCN restatement above run it.
    The caller passes the flat dilation list [g**s for each block for each stack]."""
    if not any(".stacks." in k for k in sd):
        return sd
    n_stacks = 1 + max(int(k.split(".")[3]) for k in sd if ".stacks." in k)
    out: StateDict = {}
    for k, v in sd.items():
        parts = k.split(".")
        if len(parts) > 3 and parts[0] == "blocks" and parts[2] == "stacks":
            out[".".join(["blocks", str(int(parts[1]) * n_stacks + int(parts[3]))] + parts[4:])] = v
        else:
            out[k] = v
    return out


def n_blocks_of(sd: StateDict) -> int:
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))


def is_gcn(sd: StateDict) -> bool:
    return "blocks.0.act.weight" not in sd


def kernel_size_of(sd: StateDict) -> int:
    return sd["blocks.0.conv.conv.weight"].shape[-1]


# --------------------------------------------------------------------------- networks
@torch.no_grad()
def forward(sd: StateDict, dilations: Sequence[int], x: Tensor, cond: Optional[Tensor],
            dtype: torch.dtype = torch.float32) -> Tensor:
    """networks/tcn.py:150-155 (TCN.forward) / networks/gcn.py:140-147 (GCN.forward):
    blocks in sequence, out_net 1x1, tanh for GCN only."""
    sd = {k: v.to(dtype) for k, v in flatten_wavenet_state(sd).items() if v.is_floating_point()}
    x = x.to(dtype)
    cond = None if cond is None else cond.to(dtype)
    gcn = is_gcn(sd)
    for i, d in enumerate(dilations):
        x = gcn_block(x, cond, sd, i, d) if gcn else tcn_block(x, cond, sd, i, d)
    x = F.conv1d(x, sd["out_net.weight"])
    return torch.tanh(x) if gcn else x


@torch.no_grad()
def block_forward(sd: StateDict, i: int, dilation: int, x: Tensor, cond: Optional[Tensor],
                  dtype: torch.dtype = torch.float32) -> Tensor:
    sd = {k: v.to(dtype) for k, v in sd.items() if v.is_floating_point()}
    cond = None if cond is None else cond.to(dtype)
    return (gcn_block if is_gcn(sd) else tcn_block)(x.to(dtype), cond, sd, i, dilation)


class StreamState:
    """wrapper.py:14-30 (PaddingCached): per conv, the last (k-1)*d input samples,
    zero-initialised, batch-resized on first use."""

    def __init__(self, sd: StateDict, dilations: Sequence[int], batch: int, dtype=torch.float32):
        k = kernel_size_of(sd)
        self.bufs: List[Tensor] = []
        for i, d in enumerate(dilations):
            cin = sd[f"blocks.{i}.conv.conv.weight"].shape[1]
            self.bufs.append(torch.zeros(batch, cin, (k - 1) * d, dtype=dtype))


@torch.no_grad()
def forward_chunk(sd: StateDict, dilations: Sequence[int], state: StreamState, x: Tensor,
                  cond: Optional[Tensor], dtype: torch.dtype = torch.float32) -> Tensor:
    """One streaming call of the reference's cached model (wrapper.py:33-57,131)."""
    sd = {k: v.to(dtype) for k, v in sd.items() if v.is_floating_point()}
    x = x.to(dtype)
    cond = None if cond is None else cond.to(dtype)
    gcn = is_gcn(sd)
    for i, d in enumerate(dilations):
        hist = state.bufs[i]
        full = torch.cat([hist, x], dim=-1)
        pad = hist.shape[-1]
        state.bufs[i] = full[..., full.shape[-1] - pad:] if pad > 0 else hist
        x = gcn_block(x, cond, sd, i, d, hist) if gcn else tcn_block(x, cond, sd, i, d, hist)
    x = F.conv1d(x, sd["out_net.weight"])
    return torch.tanh(x) if gcn else x


def receptive_field(kernel_size: int, dilations: Sequence[int]) -> int:
    """networks/tcn.py:157-164 / gcn.py:149-160."""
    rf = kernel_size
    for d in dilations[1:]:
        rf += (kernel_size - 1) * d
    return rf


# --------------------------------------------------------------------------- synthetic weights
def build_state(arch: str, n_blocks: int, n_channels: int, kernel_size: int, cond_dim: int,
                in_ch: int = 1, out_ch: int = 1, seed: int = 0) -> StateDict:
    """Deterministic synthetic parameters with the reference's state_dict keys and
    shapes (default-style uniform init, BatchNorm running
    stats and PReLU slopes randomised so that the fold is exercised).  Everything
    comes from one torch.Generator so the GPU box regenerates identical tensors."""
    g = torch.Generator().manual_seed(1000 + seed)

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    gcn = arch.upper() == "GCN"
    sd: StateDict = {}
    C = n_channels
    W = 2 * C if gcn else C
    has_film = gcn or cond_dim > 0
    for i in range(n_blocks):
        cin = in_ch if i == 0 else C
        p = f"blocks.{i}."
        bound = 1.0 / math.sqrt(cin * kernel_size)
        sd[p + "conv.conv.weight"] = uni((W, cin, kernel_size), bound)
        sd[p + "conv.conv.bias"] = uni((W,), bound)
        if has_film:
            lb = 1.0 / math.sqrt(max(cond_dim, 1))
            sd[p + "film.adaptor.weight"] = uni((2 * W, cond_dim), lb)
            sd[p + "film.adaptor.bias"] = uni((2 * W,), lb)
            sd[p + "film.bn.weight"] = 1.0 + uni((W,), 0.3)
            sd[p + "film.bn.bias"] = uni((W,), 0.3)
            sd[p + "film.bn.running_mean"] = torch.randn((W,), generator=g) * 0.5
            sd[p + "film.bn.running_var"] = 0.2 + torch.rand((W,), generator=g) * 1.8
            sd[p + "film.bn.num_batches_tracked"] = torch.tensor(100, dtype=torch.int64)
        if not gcn:
            sd[p + "act.weight"] = 0.05 + torch.rand((1,), generator=g) * 0.85
        sd[p + "res.weight"] = uni((C, cin, 1), 1.0 / math.sqrt(cin))
    sd["out_net.weight"] = uni((out_ch, C, 1), 1.0 / math.sqrt(C))
    return sd


def make_input(batch: int, in_ch: int, T: int, first_clip: int = 0) -> Tensor:
    """x ~ U(-1, 1), one generator per clip seeded 100 + clip index."""
    rows = []
    for b in range(batch):
        g = torch.Generator().manual_seed(100 + first_clip + b)
        rows.append(torch.rand((in_ch, T), generator=g) * 2 - 1)
    return torch.stack(rows)


# the five BASELINE.json configurations: constructor arguments of the network
CONFIGS = {
    "cfg1": dict(arch="TCN", n_blocks=4, n_channels=16, kernel_size=3, dilation_growth=2, cond_dim=2),
    "cfg2": dict(arch="TCN", n_blocks=10, n_channels=32, kernel_size=15, dilation_growth=2, cond_dim=2),
    "cfg3": dict(arch="GCN", n_blocks=10, n_channels=32, kernel_size=15, dilation_growth=2, cond_dim=2),
    "tcn-shipped": dict(arch="TCN", n_blocks=5, n_channels=32, kernel_size=3, dilation_growth=14, cond_dim=2),
    "gcn3-shipped": dict(arch="GCN", n_blocks=3, n_channels=64, kernel_size=3, dilation_growth=256, cond_dim=2),
}


def config_dilations(cfg: dict) -> List[int]:
    return [cfg["dilation_growth"] ** i for i in range(cfg["n_blocks"])]


def config_state(name: str, seed: int = 0) -> StateDict:
    c = CONFIGS[name]
    return build_state(c["arch"], c["n_blocks"], c["n_channels"], c["kernel_size"], c["cond_dim"], seed=seed)


# --------------------------------------------------------------------------- CPU timing
def time_forward(sd: StateDict, dilations: Sequence[int], x: Tensor, cond: Optional[Tensor],
                 repeats: int = 3, threads: Optional[int] = None) -> float:
    """Seconds for one forward, timed like inference.py:52-61 (perf_counter around the
    eval/no_grad call): one warm-up, best of `repeats`."""
    import os
    import time

    torch.set_num_threads(threads or os.cpu_count() or 1)
    forward(sd, dilations, x[..., : min(x.shape[-1], 4096)], cond)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        forward(sd, dilations, x, cond)
        best = min(best, time.perf_counter() - t0)
    return best
