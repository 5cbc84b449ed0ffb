"""Dev: throughput of chunked streaming (BASELINE cfg5: cfg2 network, 65 536-sample chunks, history carried)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import bench
import neural_audio_spring_reverb_b200 as N
model, arch, kw, T = bench.build_model("cfg2")
model = model.to("cuda:0").eval()
for chunk in (65536, 1024):
    total = 60 * 48000 if chunk == 65536 else 2 * 48000
    x = torch.rand(1, 1, total, device="cuda:0") * 2 - 1
    cond = torch.full((1, 2), 0.5, device="cuda:0")
    st = N.CachedStream(model)
    for s in range(0, 4 * chunk, chunk):
        st(x[..., s:s + chunk], cond)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(0, total, chunk):
        st(x[..., s:s + chunk], cond)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"streaming cfg2, chunk {chunk}: {total / ms / 1e3:.1f} M samples/s ({ms / (total / chunk) * 1e3:.0f} us per chunk, "
          f"RTF {ms / 1e3 / (total / 48000):.2e})")
