"""Dev tool: where does the end-to-end (host tensor) forward spend its time?"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench

model, arch, kw, T = bench.build_model("cfg2")
dev = torch.device("cuda:0")
model = model.to(dev).eval()
x = torch.rand(1, 1, T, device=dev) * 2 - 1
cond = torch.full((1, 2), 0.5, device=dev)
xh, ch = x.cpu().pin_memory(), cond.cpu().pin_memory()
eng = model._engine()

def wall(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6

def dev_sync():
    y = model(x, cond); torch.cuda.synchronize()
print("device path, sync each call      : %.1f us" % wall(dev_sync))
print("device path, no sync (launch cost): %.1f us" % wall(lambda: model(x, cond)))
print("host path (forward_host)          : %.1f us" % wall(lambda: model(xh, ch)))
s = torch.cuda.current_stream().cuda_stream
y = torch.empty(1, 1, T, device=dev)
def raw():
    eng.set_cond(cond.data_ptr(), 1, s); eng.forward(x.data_ptr(), y.data_ptr(), 1, T, s)
print("raw ctypes set_cond+forward nosync: %.1f us" % wall(raw))
def raw_sync():
    raw(); torch.cuda.synchronize()
print("raw ctypes set_cond+forward sync  : %.1f us" % wall(raw_sync))
yh = torch.empty(1, 1, T).pin_memory()
def copies():
    x.copy_(xh, non_blocking=True); yh.copy_(y, non_blocking=True); torch.cuda.synchronize()
print("H2D + D2H of 1.92 MB each, sync   : %.1f us" % wall(copies))
print("python overhead: _engine() key    : %.1f us" % wall(lambda: model._engine(), 200))
print("pinned empty alloc                : %.1f us" % wall(lambda: torch.empty((1, 1, T), pin_memory=True), 200))
ms = eng.forward_profiled(x.data_ptr(), y.data_ptr(), 1, T, s)
print("per-block device ms:", [round(m * 1e3) for m in ms], "sum %.0f us" % (sum(ms) * 1e3))
