"""Dev: isolate the tap-pass machinery of the ring kernel (engine.cu path 3) on tiny nets."""
import os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import nasr_oracle as O
from util import build_model, rel_err
DEV = "cuda:0"

def run(arch, k, g, taps, T, zero=None, B=1, show=False):
    os.environ["NASR_FORCE_PASSES"] = str(taps)
    cfg = dict(arch=arch, n_blocks=2, n_channels=32, kernel_size=k, dilation_growth=g, cond_dim=2)
    sd = O.build_state(arch, 2, 32, k, 2, seed=3)
    if zero is not None:
        w = sd["blocks.1.conv.conv.weight"]
        for j in zero: w[:, :, j] = 0
    m = build_model(cfg, sd, DEV); m.set_async(True)
    paths = [m._engine().block_path(i) for i in range(2)]
    x = O.make_input(B, 1, T); cond = torch.tensor([[0.2, 0.9]] * B)
    ref = O.forward(sd, [1, g], x, cond)
    y = m(x.to(DEV), cond.to(DEV)).cpu()
    err = (y - ref).abs()[0, 0]
    bad = (err > 1e-4 * ref.abs().max()).nonzero().flatten()
    print(f"{arch} k={k} d={g} taps/pass={taps} T={T} zero={zero} paths={paths} rel={rel_err(y, ref):.2e} "
          f"bad={len(bad)} first={bad[:8].tolist()} last={bad[-4:].tolist()}", flush=True)
    if show:
        print("   y  ", [round(float(v), 4) for v in y[0, 0, :10]])
        print("   ref", [round(float(v), 4) for v in ref[0, 0, :10]])
        # what would the result be with only one of the passes / a wrong shift?
        for name, zz in (("only newest tap", [0]), ("only oldest tap", [1]), ("no conv", [0, 1])):
            sd2 = {k2: v.clone() for k2, v in sd.items()}
            for j in zz: sd2["blocks.1.conv.conv.weight"][:, :, j] = 0
            r2 = O.forward(sd2, [1, g], x, cond)
            print(f"   vs '{name}': {rel_err(y, r2):.2e}")

run("TCN", 2, 1, 1, 1000, show=True)
run("TCN", 2, 1, 1, 1000, zero=[0, 1], show=True)
run("TCN", 2, 1, 1, 1000, zero=[0], show=True)
run("TCN", 2, 1, 1, 1000, zero=[1], show=True)
