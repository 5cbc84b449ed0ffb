"""Dev check of the accumulator-ring tcgen05 kernel (ring_block.cu) on a GPU box:
parity against the CPU oracle and per-block device times, ring path vs tap-gather path.

    python tools/ring_check.py [quick|full]
"""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

from oracle import nasr_oracle as O  # noqa: E402
from util import build_model, rel_err  # noqa: E402

DEV = "cuda:0"


def run(cname, B, T, path, check=True, cfg=None, sd=None):
    os.environ["NASR_PATH"] = path
    cfg = cfg or O.CONFIGS[cname]
    sd = sd or O.config_state(cname)
    m = build_model(cfg, sd, DEV)
    x = O.make_input(B, 1, T)
    cond = torch.linspace(0.1, 0.9, B * 2).view(B, 2)
    xd, cd = x.to(DEV), cond.to(DEV)
    y = m(xd, cd)
    torch.cuda.synchronize()
    eng = m._engine()
    paths = [eng.block_path(i) for i in range(cfg["n_blocks"])]
    err = None
    if check:
        dil = [cfg["dilation_growth"] ** i for i in range(cfg["n_blocks"])]
        ref = O.forward(sd, dil, x, cond)
        err = rel_err(y, ref)
    yd = torch.empty_like(y)
    best = None
    for _ in range(5):
        ms = eng.forward_profiled(xd.data_ptr(), yd.data_ptr(), B, T)
        best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
    # whole forward, events around 10 iterations
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        m(xd, cd)
    e0.record()
    for _ in range(10):
        m(xd, cd)
    e1.record()
    torch.cuda.synchronize()
    tot = e0.elapsed_time(e1) / 10
    print(f"{cname} B={B} T={T} path={path} paths={paths} err={err} total={tot:.3f} ms "
          f"({B * T / tot / 1e3:.1f} M samples/s) blocks(us)={[round(v * 1e3, 1) for v in best]}", flush=True)
    return y


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    t0 = time.time()
    run("cfg2", 1, 20000, "auto")
    run("cfg2", 2, 50000, "auto")
    if mode == "quick":
        return
    run("cfg2", 1, 480000, "tc", check=False)
    ya = run("cfg2", 1, 480000, "auto", check=True)
    run("cfg2", 8, 480000, "auto", check=False)
    run("cfg3", 1, 480000, "tc", check=False)
    run("cfg3", 1, 480000, "auto", check=True)
    run("tcn-shipped", 2, 100001, "auto")
    print(f"done in {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
