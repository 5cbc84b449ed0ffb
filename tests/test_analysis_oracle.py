"""CPU: the analysis oracle (oracle/eval_oracle.py) against what the reference itself computes: RT60 values produced
by executing the reference's own source lines (tests/golden/analysis/rt60_ref.npz, made by
tests/golden/make_golden_rt60.py), MAE against torch.nn.L1Loss, and the defining properties of ESR / DC."""
from pathlib import Path

import numpy as np
import torch

from oracle import eval_oracle as E

GOLD = Path(__file__).resolve().parent / "golden" / "analysis" / "rt60_ref.npz"


def test_rt60_oracle_equals_reference_lines():
    z = np.load(GOLD)
    assert np.array_equal(z["cases"], np.array(E.RT60_CASES, dtype=np.float64))
    for (seed, n, fs, rt, tz, decay), ref in zip(E.RT60_CASES, z["rt60"]):
        h = E.synthetic_ir(seed, n, fs, rt, tz)
        assert E.rt60_reference_dtypes(h, fs, decay) == ref                 # same dtypes, same operations: bit equal
        assert abs(E.rt60_fp64(h, fs, decay)["rt60"] - ref) <= 2.0 / fs     # fp64 sums: at most a sample or two apart
    assert any(r == 0.0 for r in z["rt60"])                                 # the except branch is among the cases


def test_rt60_edge_cases():
    assert E.rt60_fp64(np.zeros(100, np.float32), 48000.0)["rt60"] == 0.0
    assert E.rt60_reference_dtypes(np.zeros(100, np.float32), 48000.0) == 0.0
    one = np.zeros(100, np.float32)
    one[0] = 1.0
    assert E.rt60_fp64(one, 48000.0)["rt60"] == 0.0 and E.rt60_reference_dtypes(one, 48000.0) == 0.0


def test_mae_is_l1loss_and_esr_dc_properties():
    g = torch.Generator().manual_seed(3)
    t = torch.randn(3, 1, 5000, generator=g)
    p = t + 0.1 * torch.randn(3, 1, 5000, generator=g)
    m = E.eval_metrics(p, t)
    assert abs(m["eval/mae"] - float(torch.nn.L1Loss()(p.double(), t.double()))) < 1e-7
    assert E.eval_metrics(t, t) == {"eval/mae": 0.0, "eval/esr": 0.0, "eval/dc": 0.0}
    z = E.eval_metrics(torch.zeros_like(t), t)
    assert abs(z["eval/esr"] - 1.0) < 1e-6                                  # predicting silence: error energy = target energy
    off = E.eval_metrics(t + 0.5, t)                                        # a pure DC offset
    want_dc = float((0.25 / ((t.double() ** 2).mean(-1) + 1e-8)).mean())
    assert abs(off["eval/dc"] - want_dc) < 1e-6 and abs(off["eval/mae"] - 0.5) < 1e-6
