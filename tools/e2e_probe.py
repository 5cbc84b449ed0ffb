"""Dev: phase times of the host-tensor forward (NASR_E2E_DBG=1 makes the library print them)."""
import os, sys, time
os.environ["NASR_E2E_DBG"] = "1"
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
model, arch, kw, T = bench.build_model("cfg2")
model = model.to("cuda:0").eval()
xh = (torch.rand(1, 1, T) * 2 - 1).pin_memory(); ch = torch.full((1, 2), 0.5).pin_memory()
for _ in range(6):
    t0 = time.perf_counter(); y = model(xh, ch); dt = time.perf_counter() - t0
    print(f"python wall {dt * 1e6:.1f} us", file=sys.stderr)
