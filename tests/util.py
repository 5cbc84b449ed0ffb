"""Helpers shared by the tests: golden fixture loading and model construction."""
import json
from pathlib import Path

import numpy as np
import torch

from oracle import nasr_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"
REL_TOL = 1e-4   # BASELINE.json north_star: max|y_gpu - y_ref| <= 1e-4 * max|y_ref| per clip (fp32)


def golden_names():
    return sorted(p.stem for p in GOLDEN.glob("*.npz"))


def load_golden(name):
    """-> (meta dict, y tensor, state_dict) ; weights come from the fixture (shipped
    checkpoints) or are regenerated from the seed and guarded by the stored checksum."""
    z = np.load(GOLDEN / f"{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    y = torch.from_numpy(z["y"])
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    cfg = meta["cfg"]
    if not sd:
        if name.startswith("ckpt_") and name.endswith("_cond"):
            z2 = np.load(GOLDEN / f"{name[:-5]}.npz")
            sd = {k[4:]: torch.from_numpy(z2[k]) for k in z2.files if k.startswith("sd::")}
        else:
            sd = O.build_state(cfg["arch"], cfg["n_blocks"], cfg["n_channels"], cfg["kernel_size"], cfg["cond_dim"],
                               in_ch=cfg.get("in_ch", 1), out_ch=cfg.get("out_ch", 1), seed=cfg.get("seed", 0))
    chk = float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if v.is_floating_point()))
    assert abs(chk - meta["weights_checksum"]) <= 1e-9 * max(1.0, abs(chk)), \
        f"{name}: regenerated weights differ from the ones the golden output was made with"
    return meta, y, sd


def golden_inputs(meta):
    cfg = meta["cfg"]
    x = O.make_input(meta["B"], cfg.get("in_ch", 1), meta["T"])
    cond = None
    if cfg["cond_dim"] > 0:
        cond = torch.tensor(meta["cond"], dtype=torch.float32).view(1, -1).repeat(meta["B"], 1)
    return x, cond


def build_model(cfg, sd, device):
    """Our drop-in module with the reference's constructor arguments."""
    import neural_audio_spring_reverb_b200 as N
    if cfg["arch"] == "WaveNet":
        m = N.WaveNet(in_ch=cfg.get("in_ch", 1), out_ch=cfg.get("out_ch", 1), n_blocks=cfg["n_blocks"],
                      n_stacks=cfg["n_stacks"], n_channels=cfg["n_channels"], kernel_size=cfg["kernel_size"],
                      dilation_growth=cfg["dilation_growth"], cond_dim=cfg["cond_dim"])
    elif cfg["arch"] == "TCN":
        m = N.TCN(cfg["n_channels"], cfg["n_blocks"], cfg["dilation_growth"], in_ch=cfg.get("in_ch", 1),
                  out_ch=cfg.get("out_ch", 1), kernel_size=cfg["kernel_size"], cond_dim=cfg["cond_dim"])
    else:
        m = N.GCN(in_ch=cfg.get("in_ch", 1), out_ch=cfg.get("out_ch", 1), n_blocks=cfg["n_blocks"],
                  n_channels=cfg["n_channels"], dilation_growth=cfg["dilation_growth"],
                  kernel_size=cfg["kernel_size"], cond_dim=cfg["cond_dim"])
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


def rel_err(y, ref):
    """max over clips of max|y - ref| / max|ref| (the north-star parity metric)."""
    y = y.detach().double().cpu()
    ref = ref.detach().double().cpu()
    num = (y - ref).abs().flatten(1).max(dim=1).values
    den = ref.abs().flatten(1).max(dim=1).values.clamp_min(1e-30)
    return float((num / den).max())
