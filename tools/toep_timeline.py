"""Dev: per-tile timeline of CTA 0 of the tensor-core first block (NASR_TOEP_DBG=8 [+1 drain only, +2 no build])."""
import os, sys, ctypes
os.environ["NASR_TOEP_DBG"] = str(8 | (int(sys.argv[1]) if len(sys.argv) > 1 else 0))
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from oracle import nasr_oracle as O
from util import build_model
from neural_audio_spring_reverb_b200 import _native
cfg = O.CONFIGS["cfg2"]; sd = O.config_state("cfg2")
m = build_model(cfg, sd, "cuda:0")
x = torch.rand(1, 1, 480000, device="cuda:0") * 2 - 1
c = torch.full((1, 2), 0.5, device="cuda:0")
for _ in range(3): m(x, c)
torch.cuda.synchronize()
lib = _native.load_library()
lib.nasr_debug_toep_stamps.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 512)()
lib.nasr_debug_toep_stamps(buf, 512)
a = np.frombuffer(buf, dtype=np.uint64).reshape(64, 8).astype(np.int64)
t0 = a[0, 0]
print("dbg", os.environ["NASR_TOEP_DBG"], "- us since tile 0 was built; columns: built, stage_free, arrived, mma_ready, mma_issued, acc_seen, stored")
for q in range(26):
    if a[q, 0] == 0: break
    print(q, " ".join(f"{(v - t0) / 1e3:7.2f}" if v else "      -" for v in a[q, :7]))
