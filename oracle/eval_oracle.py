"""TEST INFRASTRUCTURE - CPU restatement of the analysis steps behind the forward.

* MAE / ESR / DC of evaluate_model (reference src/neural_audio_spring_reverb/eval.py:38-40,118-121).  MAE is
  torch.nn.L1Loss; ESR and DC are auraloss.time.ESRLoss / DCLoss.  auraloss is a third-party dependency (pinned as
  auraloss==0.4.0 in the reference's wandb/*/files/requirements.txt) that is NOT installed here and not vendored under
  /root/reference, so its published formulas (auraloss 0.4.0, auraloss/time.py) are restated:
      ESR = mean over (batch, channel) of  sum_t (target - input)^2 / (sum_t target^2 + eps)
      DC  = mean over (batch, channel) of  (mean_t (target - input))^2 / (mean_t target^2 + eps),  eps = 1e-8
  **Parity unpinned** for ESR / DC (no auraloss to run, no golden values in the reference); MAE is pinned to torch.
  The mel-scaled MultiResolutionSTFTLoss of eval.py:41-49 is not restated.
* RT60 of measure_rt60 (reference tools/rt60.py:49-70), restated line by line in the reference's own dtypes (fp32
  cumulative sum) and in fp64; pinned by tests/golden/analysis/rt60_ref.npz, which tests/golden/make_golden_rt60.py
  produced by executing the reference's own source lines.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from typing import Dict

import numpy as np
import torch


def eval_metrics(pred: torch.Tensor, target: torch.Tensor, eps: float = 1e-8) -> Dict[str, float]:
    """pred / target: [B, C, T]; float64 arithmetic on the fp32 difference (eval.py:118-121)."""
    p, t = pred.detach().cpu().float(), target.detach().cpu().float()
    d = (t - p).double()
    td = t.double()
    mae = float(d.abs().mean())                                            # torch.nn.L1Loss, reduction "mean"
    esr = float(((d ** 2).sum(-1) / ((td ** 2).sum(-1) + eps)).mean())     # auraloss.time.ESRLoss
    dc = float((d.mean(-1) ** 2 / ((td ** 2).mean(-1) + eps)).mean())      # auraloss.time.DCLoss
    return {"eval/mae": mae, "eval/esr": esr, "eval/dc": dc}


def rt60_reference_dtypes(h: np.ndarray, fs: float, decay_db: float = 60.0) -> float:
    """tools/rt60.py:49-72 as written: float32 power, float32 np.cumsum, float32 log10."""
    x = np.asarray(h).astype("float32")
    power = x ** 2
    energy = np.cumsum(power[::-1])[::-1]
    try:
        i_nz = np.max(np.where(energy > 0)[0])
        energy = energy[:i_nz]
        energy_db = 10 * np.log10(energy)
        energy_db -= energy_db[0]
        i_5db = np.min(np.where(-5 - energy_db > 0)[0])
        t_5db = i_5db / fs
        i_decay = np.min(np.where(-decay_db - energy_db > 0)[0])
        t_decay = i_decay / fs
        return float((60 / decay_db) * (t_decay - t_5db))
    except Exception:
        return 0.0


def rt60_fp64(h: np.ndarray, fs: float, decay_db: float = 60.0) -> Dict[str, float]:
    """The same steps with the fp32 squares summed in float64 (what the device kernel computes)."""
    x = np.asarray(h).astype("float32")
    power = (x * x).astype(np.float64)
    energy = np.cumsum(power[::-1])[::-1]
    out = dict(rt60=0.0, i_5db=-1, i_decay=-1, i_nz=-1)
    nz = np.where(energy > 0)[0]
    if len(nz) == 0:
        return out
    i_nz = int(nz.max())
    out["i_nz"] = i_nz
    e = energy[:i_nz]
    if len(e) == 0:
        return out
    db = 10 * np.log10(e) - 10 * np.log10(e[0])
    w5 = np.where(-5 - db > 0)[0]
    wd = np.where(-decay_db - db > 0)[0]
    if len(w5):
        out["i_5db"] = int(w5.min())
    if len(wd):
        out["i_decay"] = int(wd.min())
    if len(w5) and len(wd):
        out["rt60"] = float((60 / decay_db) * (out["i_decay"] / fs - out["i_5db"] / fs))
    return out


def synthetic_ir(seed: int, n: int, fs: float, rt: float, tail_zeros: int = 0) -> np.ndarray:
    """Exponentially decaying noise whose energy drops 60 dB in `rt` seconds, optional all-zero tail."""
    g = np.random.default_rng(seed)
    t = np.arange(n) / fs
    h = g.standard_normal(n) * 10.0 ** (-3.0 * t / rt)
    if tail_zeros:
        h[-tail_zeros:] = 0.0
    return h.astype(np.float32)


RT60_CASES = [  # (seed, n, fs, rt, tail_zeros, decay_db)
    (1, 48000, 48000.0, 0.35, 0, 60.0),
    (2, 96000, 48000.0, 1.2, 1000, 60.0),
    (3, 32000, 16000.0, 0.8, 0, 30.0),
    (4, 20000, 16000.0, 5.0, 0, 60.0),      # never decays 60 dB inside the buffer: the reference's except branch
    (5, 479999, 48000.0, 2.5, 17, 60.0),
    (6, 1025, 48000.0, 0.004, 0, 20.0),
]
