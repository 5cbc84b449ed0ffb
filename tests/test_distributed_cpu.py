"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (clip sharding and the
one-time weight broadcast). The data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import contextlib
    import io
    import neural_audio_spring_reverb_b200 as N
    from neural_audio_spring_reverb_b200.distributed import broadcast_weights, max_over_ranks, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1234 + rank)                      # ranks start with DIFFERENT weights
    with contextlib.redirect_stdout(io.StringIO()):
        models = [N.TCN(32, 10, 2, kernel_size=15, cond_dim=2), N.GCN(n_blocks=3, n_channels=8, dilation_growth=4, cond_dim=1)]
    sums = []
    for m in models:
        before = m.weight_blob().clone()
        broadcast_weights(m, src=0)
        after = m.weight_blob()
        sums.append((float(before.double().abs().sum()), float(after.double().abs().sum()), after.numel()))
        ref = after.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, after)                  # identical to rank 0's blob, bit for bit
        # and the module really holds them (state_dict round trip)
        assert torch.equal(m.state_dict()["out_net.weight"].reshape(-1), after[-m.out_net.weight.numel():])
    lo, hi = shard_range(513, rank, world)
    slow = max_over_ranks(1.0 + rank)
    out.put((rank, sums, (lo, hi), slow))
    dist.barrier()
    dist.destroy_process_group()


def test_weight_broadcast_and_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, sh0, t0), (r1, s1, sh1, t1) = res
    for (b0, a0, n0), (b1, a1, n1) in zip(s0, s1):
        assert n0 == n1 and a0 == a1 and b0 == a0 and b1 != a1   # rank 1 changed, rank 0 did not
    assert s0[0][2] == 150890 + 10 * 64 + 10 * 0 or s0[0][2] > 150890  # blob = params + BN running stats
    assert sh0 == (0, 257) and sh1 == (257, 513)
    assert t0 == t1 == 2.0


@pytest.mark.parametrize("n,world", [(512, 8), (513, 8), (7, 8), (1, 2), (0, 4)])
def test_shard_ranges_partition_the_batch(n, world):
    from neural_audio_spring_reverb_b200.distributed import shard_range
    edges = [shard_range(n, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for (a, b), (c, d) in zip(edges, edges[1:]):
        assert b == c and 0 <= b - a <= (n + world - 1) // world
