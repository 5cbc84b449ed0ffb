"""Dev: isolate the ring kernel on two-block nets (block 1 = ring kernel with dilation g)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

from oracle import nasr_oracle as O  # noqa: E402
from util import build_model  # noqa: E402

DEV = "cuda:0"


def case(arch, k, g, T, B=1, n_blocks=2):
    cfg = dict(arch=arch, n_blocks=n_blocks, n_channels=32, kernel_size=k, dilation_growth=g, cond_dim=2)
    sd = O.build_state(arch, n_blocks, 32, k, 2, seed=k * 31 + g)
    dil = [g ** i for i in range(n_blocks)]
    x = O.make_input(B, 1, T)
    cond = torch.linspace(0.1, 0.9, B * 2).view(B, 2)
    ref = O.forward(sd, dil, x, cond)
    out = {}
    for path in ("tc", "auto"):
        os.environ["NASR_PATH"] = path
        m = build_model(cfg, sd, DEV)
        y = m(x.to(DEV), cond.to(DEV)).cpu()
        paths = [m._engine().block_path(i) for i in range(n_blocks)]
        err = (y - ref).abs()[0, 0]
        scale = float(ref.abs().max())
        bad = (err > 1e-4 * scale).nonzero().flatten()
        msg = f"  {path}: paths={paths} max_rel={float(err.max()) / scale:.2e} bad={bad.numel()}/{T}"
        if bad.numel():
            msg += f" first={bad[:6].tolist()} last={int(bad[-1])}"
            gd = (128 // g) * g if g < 128 else g
            msg += f" bad%128 hist={torch.bincount(bad % 128, minlength=128)[:8].tolist()}.. bad%gd[:8]={torch.bincount(bad % gd)[:8].tolist()}"
        out[path] = msg
    print(f"{arch} k={k} g={g} T={T} n_blocks={n_blocks}")
    for p in out.values():
        print(p, flush=True)


if __name__ == "__main__":
    for k, g in ((1, 2), (2, 128), (3, 128), (15, 128), (15, 256), (3, 64), (3, 2), (15, 2), (15, 16), (3, 200)):
        case("TCN", k, g, 6000)
    case("TCN", 15, 128, 40000)
    case("TCN", 15, 4, 40000)
    case("GCN", 3, 128, 6000, n_blocks=3)
    case("GCN", 15, 8, 6000, n_blocks=3)
