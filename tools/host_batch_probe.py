"""Dev: wall time of the host-tensor forward for a batch (pipelined slices vs one H2D / forward / D2H sequence)."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model, arch, kw, T = bench.build_model("cfg2")
dev = torch.device("cuda:0")
model = model.to(dev).eval()
x = (torch.rand(B, 1, T) * 2 - 1).pin_memory()
c = torch.full((B, 2), 0.5).pin_memory()
xd, cd = x.to(dev), c.to(dev)
model.set_async(True)
for _ in range(3): model(xd, cd)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): model(xd, cd)
torch.cuda.synchronize(); td = (time.perf_counter() - t0) / 5
model.set_async(None)
for _ in range(3): model(x, c)
ts = []
for _ in range(6):
    t0 = time.perf_counter(); y = model(x, c); ts.append(time.perf_counter() - t0)
yh = torch.empty(B, 1, T).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    xd.copy_(x, non_blocking=True); torch.cuda.synchronize()
th2d = (time.perf_counter() - t0) / 3
t0 = time.perf_counter()
for _ in range(3):
    yh.copy_(xd, non_blocking=True); torch.cuda.synchronize()
td2h = (time.perf_counter() - t0) / 3
t0 = time.perf_counter()
for _ in range(3): z = torch.empty((B, 1, T), dtype=torch.float32, pin_memory=True)
tal = (time.perf_counter() - t0) / 3
print(f"B={B} pipe={os.environ.get('NASR_HOST_PIPE','1')}: device-resident {td*1e3:.2f} ms; host path calls (ms) {[round(t*1e3,2) for t in ts]}; "
      f"H2D {th2d*1e3:.2f} ms ({B*T*4/th2d/1e9:.1f} GB/s), D2H {td2h*1e3:.2f} ms ({B*T*4/td2h/1e9:.1f} GB/s), pinned alloc {tal*1e3:.2f} ms")
