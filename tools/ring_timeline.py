"""Dev: per-CTA timeline of one ring-kernel launch (NASR_RB_DBG=8)."""
import os, sys, ctypes
os.environ["NASR_RB_DBG"] = str(8 | (int(sys.argv[2]) if len(sys.argv) > 2 else 0))
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from oracle import nasr_oracle as O
from util import build_model
from neural_audio_spring_reverb_b200 import _native
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = 480000
cfg = O.CONFIGS["cfg2"]; sd = O.config_state("cfg2")
m = build_model(cfg, sd, "cuda:0")
x = torch.rand(B, 1, T, device="cuda:0") * 2 - 1
c = torch.full((B, 2), 0.5, device="cuda:0")
for _ in range(3):
    m(x, c)
torch.cuda.synchronize()
lib = _native.load_library()
buf = (ctypes.c_ulonglong * (16 * 160))()
n = lib.nasr_debug_ring_stamps(buf, 160)
a = np.frombuffer(buf, dtype=np.uint64).reshape(160, 16).astype(np.int64)
a = a[a[:, 0] > 0]
t0 = a[:, 0].min()
names = ["entry", "setup", "pdl_wait", "w_ready", "tile0", "warm_done", "mma_issued", "epi_first", "epi_last", "exit"]
print("dbg", os.environ["NASR_RB_DBG"], "B", B, "CTAs", len(a), "(times in us relative to the earliest CTA entry; last block of the forward)")
for i, nm in enumerate(names):
    col = (a[:, i] - t0) / 1e3
    print(f"{nm:12s} min {col.min():8.2f}  median {np.median(col):8.2f}  max {col.max():8.2f}")

sbuf = (ctypes.c_ulonglong * 256)()
lib.nasr_debug_ring_steps(sbuf)
st = np.frombuffer(sbuf, dtype=np.uint64).reshape(4, 64).astype(np.int64)
base = a[0, 0]
print("CTA 0 per step (us since its entry): tile requested | MMAs issued (d = delta to previous step) | epilogue drained | rows handed off")
prev = None
for q in range(64):
    if st[0, q] == 0 and st[3, q] == 0: continue
    f = lambda v: f"{(v - base) / 1e3:7.2f}" if v else "      -"
    d = f"{(st[0, q] - prev) / 1e3:5.2f}" if prev and st[0, q] else "    -"
    print(f"step {q - 14:3d}: {f(st[3, q])} | {f(st[0, q])} (d {d}) | {f(st[1, q])} | {f(st[2, q])}")
    if st[0, q]: prev = st[0, q]
