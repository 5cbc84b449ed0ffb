"""Dev: per-block time of the ring kernel on cfg2 under debug switches (env NASR_RB_DBG, NASR_LIB)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from oracle import nasr_oracle as O
from util import build_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 480000
cname = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
cfg = O.CONFIGS[cname]; sd = O.config_state(cname)
m = build_model(cfg, sd, "cuda:0")
x = torch.rand(B, 1, T, device="cuda:0") * 2 - 1
c = torch.full((B, 2), 0.5, device="cuda:0")
y = m(x, c); torch.cuda.synchronize()
eng = m._engine(); yd = torch.empty_like(y); best = None
for _ in range(5):
    ms = eng.forward_profiled(x.data_ptr(), yd.data_ptr(), B, T)
    best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
m.set_async(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
for _ in range(5): m(x, c)
for s_, e_ in ev:
    flush.zero_(); s_.record(); m(x, c); e_.record()
torch.cuda.synchronize()
tt = sorted(s_.elapsed_time(e_) for s_, e_ in ev)
print(f"  forward (events, L2 flushed): median {tt[len(tt)//2]*1e3:.1f} us, min {tt[0]*1e3:.1f} us -> {B*T/tt[len(tt)//2]/1e6:.3f} G samples/s; sum(blocks) {sum(best)*1e3:.1f} us")
print(f"{cname} B={B} T={T} dbg={os.environ.get('NASR_RB_DBG')} lib={os.path.basename(os.environ.get('NASR_LIB','default'))} blocks(us)={[round(v*1e3,1) for v in best]}", flush=True)
