"""Dev: where the host-tensor forward loses time when N ranks share one host (e2e efficiency 0.47 at 8 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/e2e_scale.py

Every rank runs cfg2 (one 10 s clip) through model(x_host, cond_host) under four host set-ups - CPU pinning on / off x
zero-copy result on / off - and reports the per-call wall-time distribution; rank 0 prints the table of all ranks.
"""
import os
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from neural_audio_spring_reverb_b200 import hostaffinity  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    calls = int(os.environ.get("CALLS", "300"))
    rows = []
    for pin in (0, 1):
        for zc in (1, 0):
            if pin:
                aff = hostaffinity.pin_to_gpu(local, world)
            else:
                hostaffinity.unpin()
                torch.set_num_threads(os.cpu_count() or 1)
                aff = dict(pinned=False)
            os.environ["NASR_ZEROCOPY"] = str(zc)
            model, arch, kw, T = bench.build_model("cfg2")
            model = model.to(dev).eval()
            x = (torch.rand((1, 1, T)) * 2 - 1).pin_memory()
            c = torch.full((1, 2), 0.5).pin_memory()
            for _ in range(10):
                model(x, c)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            ts = []
            t_all0 = time.perf_counter()
            for _ in range(calls):
                t0 = time.perf_counter()
                model(x, c)
                ts.append(time.perf_counter() - t0)
            t_all = time.perf_counter() - t_all0
            ts.sort()
            rows.append(dict(rank=rank, pin=pin, zerocopy=zc, aff=aff.get("cpus"), med_us=1e6 * statistics.median(ts),
                             p10_us=1e6 * ts[len(ts) // 10], p90_us=1e6 * ts[(9 * len(ts)) // 10], max_us=1e6 * ts[-1],
                             msps=T * calls / t_all / 1e6))
            model.release_engine()
            if world > 1:
                dist.barrier()
    allrows = [None] * world
    if world > 1:
        dist.all_gather_object(allrows, rows)
    else:
        allrows = [rows]
    if rank == 0:
        print(f"world {world}, {calls} synchronous host-tensor forwards of cfg2 per set-up; per-call wall time in us")
        for v in range(4):
            r0 = allrows[0][v]
            total = sum(rr[v]["msps"] for rr in allrows)
            print(f"pin={r0['pin']} zerocopy={r0['zerocopy']}: aggregate {total / 1e3:.3f} G samples/s")
            for rr in allrows:
                q = rr[v]
                print(f"   rank {q['rank']} cpus {q['aff']}: median {q['med_us']:.0f}  p10 {q['p10_us']:.0f}  p90 {q['p90_us']:.0f}  "
                      f"max {q['max_us']:.0f}  -> {q['msps']:.0f} M samples/s")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
