"""Generates tests/golden/analysis/rt60_ref.npz by EXECUTING the reference's own source lines
(/root/reference/src/neural_audio_spring_reverb/tools/rt60.py:46-72; the function itself cannot be called: it uses
`wavfile` without importing it and imports matplotlib plotting).  Run in the build container only."""
import sys
import textwrap
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import eval_oracle as E  # noqa: E402

SRC = Path("/root/reference/src/neural_audio_spring_reverb/tools/rt60.py").read_text().splitlines()
body = SRC[45:72]            # "h = np.array(x)" ... "est_rt60 = np.array(0.0)"
assert body[0].strip() == "h = np.array(x)" and "est_rt60 = np.array(0.0)" in body[-1], (body[0], body[-1])
code = compile(textwrap.dedent("\n".join(body)), "rt60.py[46:72]", "exec")

vals = []
for seed, n, fs, rt, tz, decay in E.RT60_CASES:
    x = E.synthetic_ir(seed, n, fs, rt, tz).astype("float32")
    env = dict(np=np, x=x, fs=fs, decay_db=decay)
    # the reference hard-codes decay_db = 60 one line above the block; the block itself is generic in decay_db
    exec(code, env)
    vals.append(float(env["est_rt60"]))
    print(seed, n, fs, rt, tz, decay, "->", vals[-1])
np.savez(ROOT / "tests/golden/analysis/rt60_ref.npz", rt60=np.array(vals), cases=np.array(E.RT60_CASES, dtype=np.float64))
