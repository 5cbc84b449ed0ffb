"""Dev probe (CPU): how much accuracy does the split-precision product lose when the two
correction terms of  a*w = ah*wh + al*wh + ah*wl  run as fp8 tensor-core products
(kind::f8f6f4, twice the fp16 rate) instead of fp16 ones?

Everything is emulated in float64 with the operands rounded exactly as the device would
(round-to-nearest casts through torch's fp16 / float8 dtypes); the result is compared with
the float64 forward of the oracle on the same inputs.

    python tools/fp8_split_sim.py [T]
"""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import nasr_oracle as O  # noqa: E402
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from util import load_golden, golden_names  # noqa: E402

ACT_SCALE = 64.0
E5 = torch.float8_e5m2
E4 = torch.float8_e4m3fn


def rn(x, dt):
    return x.to(torch.float32).to(dt).to(torch.float64)


def split_act(a):
    A = a * ACT_SCALE
    ah = rn(A, torch.float16)
    al = rn(A - ah, torch.float16)
    return ah, al


def split_w(w):
    m = float(w.abs().max())
    S = 2.0 ** np.floor(np.log2(1023.99 / m)) if m > 0 else 1.0
    wh = rn(w * S, torch.float16)
    wl = rn(w * S - wh, torch.float16)
    return wh, wl, S


def qconv(x, w, d, mode):
    """causal dilated conv of x [B,Cin,T] with w [W,Cin,k] (no bias) under a product scheme."""
    pad = (w.shape[-1] - 1) * d

    def cv(xx, ww):
        return F.conv1d(F.pad(xx, (pad, 0)), ww, None, dilation=d)

    if mode == "exact":
        return cv(x, w)
    ah, al = split_act(x)
    wh, wl, S = split_w(w)
    if mode == "f16x3":
        r = cv(ah, wh) + cv(ah, wl) + cv(al, wh)
    elif mode == "f16x2a":      # drop the weight correction
        r = cv(ah, wh) + cv(al, wh)
    elif mode == "f16x1":
        r = cv(ah, wh)
    elif mode.startswith("f8"):
        # activations e5m2 (full fp16 dynamic range), weights e4m3; scales are powers of two
        # chosen so that both products carry the scale of ah*wh
        adt = E5
        wdt = E4 if "w5" not in mode else E5
        al8 = rn(al * 2.0 ** 10, adt)
        wh8 = rn(wh * 2.0 ** -10, wdt)
        ah8 = rn(ah * 0.5, adt)
        wl8 = rn(wl * 2.0, wdt)
        r = cv(ah, wh) + cv(al8, wh8) + cv(ah8, wl8)
        if "r" in mode[2:]:     # second fp8 piece of ah for the weight-correction term
            ar8 = rn((ah * 0.5 - ah8), adt)
            r = r + cv(ar8, wl8)
    else:
        raise ValueError(mode)
    return r / (S * ACT_SCALE)


def forward(sd, dil, x, cond, mode):
    sd = {k: v.double() for k, v in O.flatten_wavenet_state(sd).items() if v.is_floating_point()}
    x = x.double()
    cond = None if cond is None else cond.double()
    gcn = O.is_gcn(sd)
    for i, d in enumerate(dil):
        p = f"blocks.{i}."
        y = qconv(x, sd[p + "conv.conv.weight"], d, mode) + sd[p + "conv.conv.bias"].view(1, -1, 1)
        if (p + "film.adaptor.weight") in sd:
            y = O.film(y, cond, sd, p + "film.")
        y = O.gated_af(y) if gcn else F.prelu(y, sd[p + "act.weight"])
        x = y + qconv(x, sd[p + "res.weight"], 1, mode)
    x = F.conv1d(x, sd["out_net.weight"])
    return torch.tanh(x) if gcn else x


def rel(y, ref):
    return float(((y - ref).abs().flatten(1).max(1).values / ref.abs().flatten(1).max(1).values).max())


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 24000
    modes = ["f16x3", "f8", "f8w5", "f8r", "f16x2a", "f16x1"]
    cases = []
    for name in ("cfg2", "cfg3"):
        c = O.CONFIGS[name]
        for seed in (0, 1):
            cases.append((f"{name}/s{seed}", O.config_state(name, seed), O.config_dilations(c), c["cond_dim"]))
    for n in golden_names():
        if n.startswith("ckpt_") and not n.endswith("_cond"):
            meta, _, sd = load_golden(n)
            cfg = meta["cfg"]
            if cfg["n_channels"] != 32 and "--all" not in sys.argv:
                continue
            cases.append((n[5:30], sd, meta["dilations"], cfg["cond_dim"]))
    print(f"T={T}   rel err (max|y-ref|/max|ref|) vs float64 exact forward")
    print(f"{'case':28s} " + " ".join(f"{m:>9s}" for m in modes))
    for label, sd, dil, cd in cases:
        for lvl, gain in (("0dB", 1.0), ("-60dB", 1e-3)):
            x = O.make_input(2, 1, T) * gain
            cond = torch.tensor([[0.3, 0.8], [0.0, 1.0]]) if cd else None
            ref = forward(sd, dil, x, cond, "exact")
            errs = [rel(forward(sd, dil, x, cond, m), ref) for m in modes]
            print(f"{label + ' ' + lvl:28s} " + " ".join(f"{e:9.2e}" for e in errs), flush=True)


if __name__ == "__main__":
    main()
