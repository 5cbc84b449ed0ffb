"""CPU: host-side logic and the C-ABI surface (no compute calls; there is no GPU here)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

import neural_audio_spring_reverb_b200 as N
from neural_audio_spring_reverb_b200 import _native
from oracle import nasr_oracle as O
from util import build_model, golden_names, load_golden

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "nasr_b200.h").read_text()
    declared = set(re.findall(r"NASR_API[^;(]*?\b(nasr_\w+)\s*\(", header))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    lib = ctypes.CDLL(str(_native.LIB_PATH))
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert "sm_100a" in _native.version()


def test_model_desc_struct_matches_header_size():
    # 10 int32 + float + 64 int32
    assert ctypes.sizeof(_native.ModelDesc) == 4 * (10 + 1 + _native.NASR_MAX_BLOCKS)


def test_weight_count_matches_state_dict():
    lib = _native.load_library()
    for cname in ("cfg1", "cfg2", "cfg3", "gcn3-shipped"):
        cfg = O.CONFIGS[cname]
        m = build_model(cfg, O.config_state(cname), "cpu")
        d = _native.ModelDesc()
        d.arch = 1 if cfg["arch"] == "GCN" else 0
        d.n_blocks, d.in_ch, d.out_ch, d.n_channels = cfg["n_blocks"], 1, 1, cfg["n_channels"]
        d.kernel_size, d.cond_dim, d.has_film = cfg["kernel_size"], cfg["cond_dim"], 1
        for i in range(cfg["n_blocks"]):
            d.dilations[i] = cfg["dilation_growth"] ** i
        assert lib.nasr_weight_count(ctypes.byref(d)) == m.weight_blob().numel()
    params = sum(p.numel() for p in build_model(O.CONFIGS["cfg2"], O.config_state("cfg2"), "cpu").parameters())
    assert params == 150890   # cfg2 parameter count


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    m = build_model(O.CONFIGS["cfg1"], O.config_state("cfg1"), "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 64), torch.zeros(1, 2))
    with pytest.raises(RuntimeError):   # nasr_engine_create: no CUDA device
        _native.Engine(arch=0, n_blocks=1, in_ch=1, out_ch=1, n_channels=4, kernel_size=3, cond_dim=0,
                       has_film=False, final_tanh=False, dilations=[1], weights=torch.zeros(4 * 3 + 4 + 1 + 4 + 4).numpy(),
                       device=0)


def test_engine_create_rejects_bad_descriptors():
    with pytest.raises(ValueError):
        _native.Engine(arch=0, n_blocks=1, in_ch=1, out_ch=1, n_channels=4, kernel_size=3, cond_dim=0,
                       has_film=False, final_tanh=False, dilations=[1], weights=torch.zeros(5).numpy(), device=0)
    with pytest.raises(ValueError):
        _native.Engine(arch=7, n_blocks=1, in_ch=1, out_ch=1, n_channels=4, kernel_size=3, cond_dim=0,
                       has_film=False, final_tanh=False, dilations=[1], weights=torch.zeros(5).numpy(), device=0)


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("ckpt_") and not n.endswith("_cond")])
def test_strict_load_of_shipped_checkpoints(name):
    """state_dict keys/shapes equal the reference's, so model_utils.py:163 (strict) succeeds."""
    meta, _, sd = load_golden(name)
    m = build_model(meta["cfg"], sd, "cpu")
    assert set(m.state_dict().keys()) == set(sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    assert m.calc_receptive_field() == meta["rf"]
    assert sum(p.numel() for p in m.parameters()) == meta["params"]
    assert [layer.dilation for layer in m._nasr_layers()] == meta["dilations"]


def test_reference_attribute_surface():
    m = N.TCN(16, 4, 2, kernel_size=3, cond_dim=2)
    for attr in ("in_ch", "out_ch", "kernel_size", "cond_dim", "channels", "dilations", "n_blocks", "strides",
                 "blocks", "out_net"):
        assert hasattr(m, attr)
    assert m.dilations == [1, 2, 4, 8] and m.calc_receptive_field() == 31
    assert sum(p.numel() for p in m.parameters()) == 3732      # cfg1
    g = N.GCN(n_blocks=3, n_channels=64, dilation_growth=256, kernel_size=3, cond_dim=2)
    assert g.calc_receptive_field() == 131587
    assert not hasattr(N.TCN(8, 2, 2, cond_dim=0).blocks[0], "film")   # tcn.py:63-64
    assert hasattr(N.GCN(cond_dim=0).blocks[0], "film")                # gcn.py:45


def test_load_model_checkpoint_and_errors(tmp_path):
    import types
    meta, _, sd = load_golden("ckpt_GCN_springset_20240324_151439_16kHz")
    c = meta["cfg"]
    config = dict(name="GCN", model_type="GCN", cond_dim=2, c0=0.0, c1=0.0, in_ch=1, out_ch=1, n_channels=c["n_channels"],
                  n_blocks=c["n_blocks"], dilation_growth=c["dilation_growth"], kernel_size=c["kernel_size"],
                  sample_rate=16000, batch_size=64, lr=0.01)
    ck = tmp_path / "gcn.pt"
    torch.save({"label": "g", "timestamp": "0", "model_state_dict": sd, "optimizer_state_dict": {"a": 1},
                "scheduler_state_dict": None, "config_state_dict": config}, ck)
    model, opt, sched, cfg, rf, params = N.load_model_checkpoint(types.SimpleNamespace(checkpoint=str(ck), device="cpu"))
    assert isinstance(model, N.GCN) and opt == {"a": 1} and sched is None and rf == meta["rf"] and params == meta["params"]
    with pytest.raises(ValueError):
        N.initialize_model("cpu", dict(config, model_type="Transformer"))
    with pytest.raises(NotImplementedError):
        N.initialize_model("cpu", dict(config, model_type="LSTM"))
    bad = dict(sd)
    bad.pop("out_net.weight")
    torch.save({"model_state_dict": bad, "config_state_dict": config}, ck)
    with pytest.raises(RuntimeError):   # strict load_state_dict
        N.load_model_checkpoint(types.SimpleNamespace(checkpoint=str(ck), device="cpu"))
    torch.save({"model_state_dict": sd}, ck)
    with pytest.raises(KeyError):       # model_utils.py:160
        N.load_model_checkpoint(types.SimpleNamespace(checkpoint=str(ck), device="cpu"))


def test_weight_blob_order_matches_c_oracle_order():
    """The product's blob and the oracle's independently written blob agree element for element."""
    from oracle import c_oracle
    for cname in ("cfg1", "cfg3"):
        cfg = O.CONFIGS[cname]
        sd = O.config_state(cname)
        m = build_model(cfg, sd, "cpu")
        ob = c_oracle.blob_from_state(sd, cfg["n_blocks"], cfg["arch"] == "GCN")
        assert torch.equal(m.weight_blob(), torch.from_numpy(ob))


def test_ir_signals_match_reference_fixture():
    """tools/ir_signals.py against the arrays the reference's own generate_reference produced
    (tests/golden/ir/ir_signals_ref.npz, made by tests/golden/make_golden.py:make_ir_signals)."""
    import numpy as np
    from pathlib import Path
    from neural_audio_spring_reverb_b200.tools.ir_signals import generate_reference
    z = np.load(Path(__file__).resolve().parent / "golden" / "ir" / "ir_signals_ref.npz")
    sweep, inv, ref_ir = generate_reference(0.05, 48000, with_reference=True)
    assert np.array_equal(sweep, z["sweep"]) and np.array_equal(inv, z["inv"])
    assert np.allclose(ref_ir, z["ref_ir"], rtol=0, atol=1e-12)
    # sweep * inverse filter is (close to) an impulse: the peak dominates
    peak = np.abs(ref_ir).max()
    assert np.abs(ref_ir).argmax() == len(sweep) - 1 and np.median(np.abs(ref_ir)) < 0.05 * peak


def test_postprocess_argument_checks_without_a_device():
    """nasr_postprocess validates its arguments before touching the device; on a box without a GPU a valid call
    fails with NASR_ERR_CUDA instead of computing anything on the CPU."""
    import ctypes as C
    import numpy as np
    lib = _native.load_library()
    assert lib.nasr_postprocess_workspace_bytes(0, 100) == 0
    assert lib.nasr_postprocess_workspace_bytes(16, 30000) == 256 + 2 * 16 * 16 * ((30000 + 63) // 64)
    b = np.array([1.0, -2.0, 1.0], dtype=np.float32)
    a = np.array([1.0, -1.9, 0.9], dtype=np.float32)
    a0 = np.array([0.0, -1.9, 0.9], dtype=np.float32)
    dummy = C.c_void_p(256)      # non-NULL, never dereferenced on the host
    inval = _native.NASR_ERR_INVALID
    assert lib.nasr_postprocess(None, dummy, 1, 10, b.ctypes.data, a.ctypes.data, 1, dummy, 1 << 20, None) == inval
    assert lib.nasr_postprocess(dummy, dummy, 0, 10, b.ctypes.data, a.ctypes.data, 1, dummy, 1 << 20, None) == inval
    assert lib.nasr_postprocess(dummy, dummy, 1, 10, b.ctypes.data, a.ctypes.data, 1, dummy, 8, None) == inval      # workspace
    assert lib.nasr_postprocess(dummy, dummy, 1, 10, b.ctypes.data, a0.ctypes.data, 1, dummy, 1 << 20, None) == inval  # a0 == 0
    assert lib.nasr_postprocess(dummy, dummy, 1, 0, b.ctypes.data, a.ctypes.data, 1, dummy, 1 << 20, None) == _native.NASR_OK
    import torch
    if not torch.cuda.is_available():
        rc = lib.nasr_postprocess(dummy, dummy, 1, 10, b.ctypes.data, a.ctypes.data, 1, dummy, 1 << 20, None)
        assert rc == _native.NASR_ERR_CUDA


def test_wav_io_without_torchcodec(tmp_path):
    """inference.py's WAV reader / writer work without TorchCodec (torchaudio.load needs it in recent releases):
    24-bit PCM - the format of the reference's audio assets - and the float32 files the writer produces."""
    import wave
    import numpy as np
    import torch
    from neural_audio_spring_reverb_b200.inference import _read_wav, _write_wav
    rng = np.random.default_rng(0)
    ints = rng.integers(-(1 << 23), 1 << 23, size=4800, dtype=np.int64)
    raw = b"".join(int(v & 0xFFFFFF).to_bytes(3, "little") for v in ints)
    p24 = tmp_path / "in24.wav"
    with wave.open(str(p24), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(3)
        w.setframerate(48000)
        w.writeframes(raw)
    x, sr = _read_wav(str(p24))
    assert sr == 48000 and tuple(x.shape) == (1, 4800) and x.dtype == torch.float32
    assert np.allclose(x[0].numpy(), ints / float(1 << 23), atol=1e-7)
    out = tmp_path / "out.wav"
    y = torch.linspace(-1, 1, 1000).unsqueeze(0)
    _write_wav(str(out), y, 16000)
    z, sr2 = _read_wav(str(out))
    assert sr2 == 16000 and tuple(z.shape) == (1, 1000)
    assert float((z - y).abs().max()) <= 1.0 / 32768 + 1e-7     # torchaudio may store 16-bit PCM, scipy float32
