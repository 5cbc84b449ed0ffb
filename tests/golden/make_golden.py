"""Generate the golden vectors under tests/golden/ by running the REFERENCE itself.

Run once in the build container (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

For every case the reference's own TCN / GCN class (imported read-only from
/root/reference/src) is run in eval mode on the CPU in fp32 (and fp64 for the
noise-floor record) and the output is stored as a fixture:
  * synthetic-weight cases: weights come from oracle.build_state(seed) (regenerated
    on the GPU box from the seed; a checksum guards the regeneration), loaded into
    the reference class with load_state_dict(strict=True);
  * shipped checkpoints: the 4 TCN + 4 GCN .pt files under /root/reference/models;
    their fp32 parameters are stored in the fixture because the checkpoints do not
    travel to the GPU box.
The reference's streaming classes (wrapper.py:14-57) cannot be imported
(neutone_sdk is absent), so their source is exec'd from the file to record that
chunked == one-shot on the reference (bit-exact) for two cases.
"""
import ast
import json
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
sys.path.insert(0, str(REF / "src"))
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from neural_audio_spring_reverb.networks.tcn import TCN as RefTCN  # noqa: E402
from neural_audio_spring_reverb.networks.gcn import GCN as RefGCN  # noqa: E402
from neural_audio_spring_reverb.networks.wavenet import WaveNet as RefWaveNet  # noqa: E402
from neural_audio_spring_reverb.networks.custom_layers import Conv1dCausal  # noqa: E402
from oracle import nasr_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent
torch.set_num_threads(8)


def ref_streaming_namespace():
    """exec PaddingCached / Conv1dCached / replace_modules from the reference file."""
    src = (REF / "src/neural_audio_spring_reverb/wrapper.py").read_text()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef))
            and n.name in ("PaddingCached", "Conv1dCached", "replace_modules")]
    ns = {"torch": torch, "nn": torch.nn, "Tensor": torch.Tensor, "Conv1dCausal": Conv1dCausal}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "wrapper_subset", "exec"), ns)
    return ns


def build_ref(cfg):
    if cfg["arch"] == "WaveNet":
        return RefWaveNet(in_ch=cfg.get("in_ch", 1), out_ch=cfg.get("out_ch", 1), n_blocks=cfg["n_blocks"],
                          n_stacks=cfg["n_stacks"], n_channels=cfg["n_channels"], kernel_size=cfg["kernel_size"],
                          dilation_growth=cfg["dilation_growth"], cond_dim=cfg["cond_dim"])
    if cfg["arch"] == "TCN":
        return RefTCN(cfg["n_channels"], cfg["n_blocks"], cfg["dilation_growth"], in_ch=cfg.get("in_ch", 1),
                      out_ch=cfg.get("out_ch", 1), kernel_size=cfg["kernel_size"], cond_dim=cfg["cond_dim"])
    return RefGCN(in_ch=cfg.get("in_ch", 1), out_ch=cfg.get("out_ch", 1), n_blocks=cfg["n_blocks"],
                  n_channels=cfg["n_channels"], dilation_growth=cfg["dilation_growth"],
                  kernel_size=cfg["kernel_size"], cond_dim=cfg["cond_dim"])


def checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if v.is_floating_point()))


def run_case(name, cfg, sd, B, T, cond_vals, store_weights, check_stream=False):
    model = build_ref(cfg)
    model.load_state_dict(sd, strict=True)
    model.eval()
    x = O.make_input(B, cfg.get("in_ch", 1), T)
    cond = None
    if cfg["cond_dim"] > 0:
        cond = torch.tensor(cond_vals, dtype=torch.float32).view(1, -1).repeat(B, 1)
    with torch.no_grad():
        y = model(x, cond)
        y64 = model.double()(x.double(), None if cond is None else cond.double())
        model.float()
    noise = float((y.double() - y64).abs().max() / y64.abs().max())
    meta = dict(name=name, cfg=cfg, B=B, T=T, cond=cond_vals, weights_checksum=checksum(sd),
                ref_fp32_vs_fp64=noise, ymax=float(y.abs().max()),
                dilations=([int(d) for d in model.dilations] if hasattr(model, "dilations") else
                           [int(st.conv.conv.dilation[0]) for blk in model.blocks for st in blk.stacks]),
                rf=int(model.calc_receptive_field()),
                params=int(sum(p.numel() for p in model.parameters())))
    if check_stream:
        ns = ref_streaming_namespace()
        ns["replace_modules"](model)
        outs = []
        with torch.no_grad():
            for s in range(0, T, 1024):
                outs.append(model(x[..., s:s + 1024], cond))
        ys = torch.cat(outs, -1)
        meta["ref_stream1024_equals_oneshot_maxabs"] = float((ys - y).abs().max())
    arrays = dict(y=y.numpy().astype(np.float32), meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    if store_weights:
        for k, v in sd.items():
            arrays["sd::" + k] = v.numpy()
    np.savez_compressed(OUT / f"{name}.npz", **arrays)
    print(f"{name}: y{tuple(y.shape)} max|y|={meta['ymax']:.4f} fp32-vs-fp64={noise:.2e}"
          + (f" stream-vs-oneshot={meta['ref_stream1024_equals_oneshot_maxabs']:.1e}" if check_stream else ""))


def main():
    # ---- synthetic-weight cases (BASELINE.json config shapes, short clips) ----
    synth = [
        ("synth_cfg1", "cfg1", 2, 4096, [0.5, 0.5], True),
        ("synth_cfg2", "cfg2", 2, 20480, [0.5, 0.5], True),
        ("synth_cfg3_c000", "cfg3", 1, 18000, [0.0, 0.0], False),
        ("synth_cfg3_c025", "cfg3", 1, 18000, [0.25, 0.25], False),
        ("synth_cfg3_c050", "cfg3", 1, 18000, [0.5, 0.5], False),
        ("synth_cfg3_c075", "cfg3", 1, 18000, [0.75, 0.75], False),
        ("synth_cfg3_c100", "cfg3", 1, 18000, [1.0, 1.0], False),
        ("synth_tcn_shipped_shape", "tcn-shipped", 1, 90000, [0.3, 0.7], False),
        ("synth_gcn3_shipped_shape", "gcn3-shipped", 1, 140000, [0.3, 0.7], False),
    ]
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for name, cname, B, T, cond, stream in ([] if only else synth):
        cfg = dict(O.CONFIGS[cname])
        run_case(name, cfg, O.config_state(cname), B, T, cond, store_weights=False, check_stream=stream)
    # odd shapes: no FiLM, multi-channel I/O, channel counts that need padding
    odd = [
        ("synth_tcn_nofilm", dict(arch="TCN", n_blocks=3, n_channels=8, kernel_size=5, dilation_growth=3, cond_dim=0), 2, 3000, []),
        ("synth_tcn_io2", dict(arch="TCN", n_blocks=3, n_channels=10, kernel_size=4, dilation_growth=2, cond_dim=3, in_ch=2, out_ch=2), 2, 2500, [0.1, 0.9, 0.4]),
        ("synth_gcn_c6", dict(arch="GCN", n_blocks=2, n_channels=6, kernel_size=3, dilation_growth=5, cond_dim=1), 3, 2000, [0.6]),
        ("synth_gcn_k99", dict(arch="GCN", n_blocks=2, n_channels=32, kernel_size=99, dilation_growth=16, cond_dim=2), 1, 6000, [0.2, 0.8]),
    ]
    for name, cfg, B, T, cond in ([] if only else odd):
        sd = O.build_state(cfg["arch"], cfg["n_blocks"], cfg["n_channels"], cfg["kernel_size"], cfg["cond_dim"],
                           in_ch=cfg.get("in_ch", 1), out_ch=cfg.get("out_ch", 1), seed=7)
        cfg["seed"] = 7
        run_case(name, cfg, sd, B, T, cond, store_weights=False)

    # ---- the 8 shipped TCN / GCN checkpoints ----
    shipped = [
        ("models/TCN-egfxset-20240229-002014-48kHz.pt", 90000),
        ("models/TCN-springset-20240228-142838-16kHz.pt", 32000),
        ("models/kernel-99/TCN-99-egfxset-20240229-194858-48kHz.pt", 8192),
        ("models/kernel-99/TCN-99-springset-20240229-073453-16kHz.pt", 8192),
        ("models/GCN-3-egfxset-20240324-160003-48kHz.pt", 140000),
        ("models/GCN-springset-20240324-151439-16kHz.pt", 72000),
        ("models/kernel-99/GCN-99-egfxset-20240310-095201-48kHz.pt", 56000),
        ("models/kernel-99/GCN-99-springset-20240310-141620-16kHz.pt", 16000),
        # WaveNet: Conv1dStack == GCNBlock, same kernels, restarting dilations
        ("models/WaveNet-egfxset-20240229-010530-48kHz.pt", 48000),
        ("models/WaveNet-springset-20240228-145414-16kHz.pt", 46000),
        ("models/kernel-99/WaveNet-99-egfxset-20240229-214944-48kHz.pt", 8192),
        ("models/kernel-99/WaveNet-99-springset-20240229-080542-16kHz.pt", 8192),
    ]
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for rel, T in shipped:
        if only and only not in rel:
            continue
        ck = torch.load(REF / rel, map_location="cpu")
        c = ck["config_state_dict"]
        cfg = dict(arch=c["model_type"], n_blocks=c.get("n_layers", c.get("n_blocks")), n_channels=c["n_channels"],
                   kernel_size=c["kernel_size"], dilation_growth=c["dilation_growth"], cond_dim=c["cond_dim"],
                   in_ch=c["in_ch"], out_ch=c["out_ch"], batch_size=c["batch_size"], sample_rate=c["sample_rate"],
                   checkpoint=rel)
        if c["model_type"] == "WaveNet":
            cfg["n_stacks"] = c["n_stacks"]
        name = "ckpt_" + Path(rel).stem.replace("-", "_")
        run_case(name, cfg, ck["model_state_dict"], 1, T, [c.get("c0", 0.0), c.get("c1", 0.0)], store_weights=True)
        run_case(name + "_cond", cfg, ck["model_state_dict"], 1, min(T, 20000), [0.35, 0.8], store_weights=False)


if __name__ == "__main__":
    main()


def make_ir_signals():
    """tests/golden/ir/ir_signals_ref.npz: sweep, inverse filter and reference IR of the reference's own
    tools/ir_signals.py:generate_reference(0.05 s, 48 kHz) (float64), run in this container."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_ir_signals", "/root/reference/src/neural_audio_spring_reverb/tools/ir_signals.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    sweep, inv, ref_ir = ref.generate_reference(0.05, 48000)
    out = Path(__file__).resolve().parent / "ir"
    out.mkdir(exist_ok=True)
    np.savez_compressed(out / "ir_signals_ref.npz", sweep=sweep, inv=inv, ref_ir=ref_ir)
