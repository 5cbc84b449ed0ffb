"""Multi-GPU plumbing for the forward path: one process per GPU (torchrun), clips sharded
by rank, and a single broadcast of the weight blob.  There is no collective on the data
path: eval-mode clips are independent (BatchNorm uses running stats,
custom_layers.py:38-39; cond is per row)."""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of the clips owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n_clips, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def broadcast_weights(model, src: int = 0, device=None) -> None:
    """Every rank ends up with rank `src`'s parameters: ONE broadcast of the flat fp32 blob
    (cfg2: 150 890 floats = 0.6 MB) - NCCL over NVLink on GPUs, gloo in the CPU tests."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    blob = model.weight_blob()
    if device is not None:
        blob = blob.to(device)
    if dist.get_rank() != src:
        blob.zero_()
    dist.broadcast(blob, src=src)
    model.load_weight_blob(blob.cpu())


def max_over_ranks(seconds: float, device=None) -> float:
    """Timing convention of bench.py: a multi-GPU step takes as long as its slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
