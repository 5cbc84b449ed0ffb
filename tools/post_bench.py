"""Dev: time the device post-processing (nasr_postprocess) next to the reference's lines (torchaudio on the host
cores, and torchaudio's own CUDA lfilter) on one 10 s clip reshaped to 16 rows, as make_inference does."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import neural_audio_spring_reverb_b200.inference as inf
from oracle import post_oracle as P

rows, T, sr = 16, 30000, 48000
torch.manual_seed(0)
y = torch.randn(rows, 1, T) * torch.exp(-torch.arange(T) / 5000.0) + 0.04
yd = y.cuda()
for _ in range(3):
    out = inf.postprocess(yd, sr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = inf.postprocess(yd, sr)
e1.record(); torch.cuda.synchronize()
ours = e0.elapsed_time(e1) / 20
t0 = time.perf_counter()
for _ in range(3):
    ref = P.postprocess_reference_fp32(y, sr)
cpu = (time.perf_counter() - t0) / 3 * 1e3
import torchaudio
def ta_cuda():
    x = yd.clone(); x /= x.abs().max(); x = torchaudio.functional.highpass_biquad(x, sr, 20.0); x = x.reshape(1, -1); x /= x.abs().max(); return x
ta_cuda(); torch.cuda.synchronize()
t0 = time.perf_counter(); r2 = ta_cuda(); torch.cuda.synchronize(); tac = (time.perf_counter() - t0) * 1e3
ref64 = P.postprocess(y, sr)
print(f"post-processing of {rows}x{T} samples: nasr_postprocess {ours * 1e3:.1f} us | reference lines on host (torchaudio CPU, "
      f"{torch.get_num_threads()} threads) {cpu:.2f} ms | torchaudio CUDA lfilter {tac:.2f} ms")
print(f"max|ours - fp64 oracle| {float((out.cpu() - ref64).abs().max()):.2e}; max|reference fp32 (CPU) - fp64 oracle| "
      f"{float((ref - ref64).abs().max()):.2e}; max|torchaudio CUDA - fp64 oracle| {float((r2.cpu() - ref64).abs().max()):.2e}")
