"""Host-side placement of a rank's process next to its GPU.

One process per GPU (torchrun): the synchronous host-tensor forward (`nasr_forward_host`: H2D copy, launches, result
written to pinned host memory, stream sync) is latency-sensitive on the host side.  With eight ranks on one socket the
processes, their CUDA helper threads and the torch intra-op pools migrate over all cores and wake each other up; every
rank therefore pins itself to its own slice of the cores NVML reports as local to its GPU and keeps its thread pools
inside that slice.  `NASR_PIN=0` turns it off.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional


def _gpu_cpu_set(index: int) -> Optional[List[int]]:
    """CPUs NVML reports as ideal for GPU `index` (its NUMA node), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        return cpus or None
    except Exception:
        return None


def slice_for_rank(cpus: List[int], local_rank: int, local_world: int, sharing: List[int]) -> List[int]:
    """This rank's share of `cpus`: the ranks in `sharing` (local ranks whose GPUs have the same CPU set, in
    order) split it evenly; a rank never gets fewer than one core."""
    if local_rank not in sharing:
        sharing = sorted(set(sharing) | {local_rank})
    n = len(sharing)
    pos = sharing.index(local_rank)
    per = max(1, len(cpus) // n)
    begin = (pos * per) % len(cpus)
    out = cpus[begin:begin + per]
    return out or cpus[:1]


def pin_to_gpu(local_rank: int, local_world: int) -> Dict:
    """Pin the calling process (and size its torch thread pool) to its slice of the GPU-local cores."""
    info: Dict = dict(pinned=False)
    if os.environ.get("NASR_PIN", "1").lower() in ("0", "false", "no") or not hasattr(os, "sched_setaffinity"):
        info["reason"] = "disabled"
        return info
    try:
        allowed = sorted(os.sched_getaffinity(0))
        mine = _gpu_cpu_set(local_rank)
        cpus = [c for c in (mine or allowed) if c in allowed] or allowed
        sharing = []
        for r in range(local_world):
            other = _gpu_cpu_set(r)
            oc = [c for c in (other or allowed) if c in allowed] or allowed
            if oc == cpus:
                sharing.append(r)
        if local_world <= 1:
            part = cpus
        else:
            part = slice_for_rank(cpus, local_rank, local_world, sharing)
        os.sched_setaffinity(0, set(part))
        try:
            import torch
            torch.set_num_threads(max(1, len(part)))
        except Exception:
            pass
        info.update(pinned=True, cpus=f"{part[0]}-{part[-1]}" if part else "", n_cpus=len(part),
                    gpu_local_cpus=len(cpus), ranks_sharing=len(sharing))
    except Exception as exc:   # placement is an optimisation, never a failure
        info["reason"] = f"{type(exc).__name__}: {exc}"
    return info


def unpin() -> None:
    if hasattr(os, "sched_setaffinity"):
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except Exception:
            pass
