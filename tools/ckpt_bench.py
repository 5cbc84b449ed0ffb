"""Dev / evidence: throughput of every shipped checkpoint shape (weights from tests/golden) on the tensor-core
path the engine picks by default, next to the fp32 FFMA kernels (NASR_PATH=fp32), with the parity of both
against the reference's golden vector.  One 10 s clip at the checkpoint's sample rate, device-resident,
CUDA events, L2 flushed between forwards.

    python tools/ckpt_bench.py [--iters 10] [--json out.json]
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from util import build_model, golden_inputs, golden_names, load_golden, rel_err  # noqa: E402

DEV = "cuda:0"


def timed(m, x, cond, iters, flush):
    for _ in range(3):
        m(x, cond)
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m(x, cond)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default="")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    rows = []
    names = [n for n in golden_names() if (n.startswith("ckpt_") and not n.endswith("_cond")) or n in ("synth_cfg1",)]
    if args.only:
        names = [n for n in names if args.only in n]
    for name in names:
        meta, y_ref, sd = load_golden(name)
        cfg = meta["cfg"]
        sr = cfg.get("sample_rate", 48000)
        T = 10 * sr
        gx, gcond = golden_inputs(meta)
        g = torch.Generator(device=DEV).manual_seed(1)
        x = torch.rand((1, 1, T), device=DEV, generator=g) * 2 - 1
        cond = torch.tensor([[0.3, 0.7]], device=DEV) if cfg["cond_dim"] else None
        row = dict(name=name.replace("ckpt_", ""), arch=cfg["arch"], C=cfg["n_channels"], k=cfg["kernel_size"],
                   dilations=meta.get("dilations"), T=T)
        ys = {}
        for mode in ("auto", "fp32"):
            os.environ["NASR_PATH"] = mode
            m = build_model(cfg, sd, DEV)
            m.set_async(True)
            y = m(gx.to(DEV), None if gcond is None else gcond.to(DEV))
            row[f"{mode}_parity"] = rel_err(y, y_ref)
            eng = m._engine()
            nb = len(meta.get("dilations") or []) or cfg["n_blocks"]
            row[f"{mode}_paths"] = [eng.block_path(i) for i in range(nb)]
            ms = timed(m, x, cond, args.iters, flush)
            row[f"{mode}_ms"] = ms
            row[f"{mode}_msps"] = T / ms / 1e3
            ys[mode] = m(x, cond)
            if mode == "auto":   # per-block device time in the pipelined forward (kernel-side stamps)
                yb = torch.empty((1, 1, T), device=DEV)
                st = torch.cuda.current_stream().cuda_stream
                if cond is not None:
                    eng.set_cond(cond.data_ptr(), 1, st)
                acc = None
                for _ in range(3):
                    flush.fill_(1.0)
                    msb = eng.forward_profiled(x.data_ptr(), yb.data_ptr(), 1, T, st)
                    acc = msb if acc is None else [p + q for p, q in zip(acc, msb)]
                row["auto_block_us"] = [round(v / 3 * 1e3, 1) for v in acc]
            row[f"{mode}_saturated"] = bool(m.saturated())
        os.environ["NASR_PATH"] = "auto"
        row["auto_vs_fp32_full"] = rel_err(ys["auto"], ys["fp32"])
        row["speedup"] = row["fp32_ms"] / row["auto_ms"]
        rows.append(row)
        print(f"{row['name']:46s} {row['arch']:7s} C={row['C']:<3d} k={row['k']:<3d} paths={row['auto_paths']} "
              f"auto {row['auto_ms']:.3f} ms ({row['auto_msps']:.0f} Msamples/s, parity {row['auto_parity']:.1e}) "
              f"fp32 {row['fp32_ms']:.3f} ms ({row['fp32_msps']:.0f} Ms/s, parity {row['fp32_parity']:.1e}) "
              f"x{row['speedup']:.2f}  full-clip auto vs fp32 {row['auto_vs_fp32_full']:.1e}  block us {row['auto_block_us']}", flush=True)
    if args.json:
        Path(args.json).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
