"""Dev: a few streaming chunks of cfg2 (for an ncu launch list: per-kernel durations of one chunk).
    python tools/stream_probe.py CHUNK N_CHUNKS"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import bench
import neural_audio_spring_reverb_b200 as N
chunk, n = int(sys.argv[1]), int(sys.argv[2])
model, arch, kw, T = bench.build_model("cfg2")
model = model.to("cuda:0").eval()
x = torch.rand(1, 1, chunk * n, device="cuda:0") * 2 - 1
cond = torch.full((1, 2), 0.5, device="cuda:0")
st = N.CachedStream(model)
for s in range(0, chunk * n, chunk):
    st(x[..., s:s + chunk], cond)
torch.cuda.synchronize()
