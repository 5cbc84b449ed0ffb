"""Dev: where the Python side of the host-tensor forward spends its time."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
model, arch, kw, T = bench.build_model("cfg2")
model = model.to("cuda:0").eval()
xh = (torch.rand(1, 1, T) * 2 - 1).pin_memory(); ch = torch.full((1, 2), 0.5).pin_memory()
for _ in range(5): model(xh, ch)
def t(fn, n=200):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6
eng = model._engine(); dev = torch.device("cuda", eng.device)
print("model._nasr_check      %.1f us" % t(lambda: model._nasr_check(xh, ch)))
print("model._engine          %.1f us" % t(lambda: model._engine()))
print("torch.device           %.1f us" % t(lambda: torch.device("cuda", eng.device)))
print("current_stream         %.1f us" % t(lambda: torch.cuda.current_stream(dev).cuda_stream))
print("empty pinned           %.1f us" % t(lambda: torch.empty((1, 1, T), dtype=torch.float32, pin_memory=True)))
print("x checks               %.1f us" % t(lambda: (xh.dtype == torch.float32 and xh.is_contiguous() and not xh.requires_grad)))
y = torch.empty((1, 1, T), dtype=torch.float32, pin_memory=True)
s = torch.cuda.current_stream(dev).cuda_stream
print("eng.forward_host       %.1f us" % t(lambda: eng.forward_host(xh.data_ptr(), ch.data_ptr(), y.data_ptr(), 1, T, s), 30))
print("model._nasr_run        %.1f us" % t(lambda: model._nasr_run(xh, ch, False), 30))
print("model(x, c)            %.1f us" % t(lambda: model(xh, ch), 30))
with torch.no_grad():
    print("model(x, c) no_grad    %.1f us" % t(lambda: model(xh, ch), 30))
