"""CPU: the oracle restatement against the golden vectors produced by the reference."""
import pytest
import torch

from oracle import nasr_oracle as O
from util import golden_inputs, golden_names, load_golden, rel_err

# long fixtures are covered by the GPU parity test; keep the CPU suite to a few minutes
CPU_CASES = [n for n in golden_names() if not n.startswith("ckpt_GCN_99_egfxset_20240310_095201_48kHz.")]


@pytest.mark.parametrize("name", CPU_CASES)
def test_oracle_matches_reference_golden(name):
    meta, y_ref, sd = load_golden(name)
    x, cond = golden_inputs(meta)
    torch.set_num_threads(8)
    y = O.forward(sd, meta["dilations"], x, cond)
    # same ATen operators as the reference -> agreement at the fp32 noise floor
    assert y.shape == y_ref.shape
    assert rel_err(y, y_ref) <= 2e-6, rel_err(y, y_ref)


@pytest.mark.parametrize("name", ["synth_cfg1", "synth_tcn_io2", "synth_gcn_c6"])
def test_oracle_fp64_is_within_reference_noise(name):
    meta, y_ref, sd = load_golden(name)
    x, cond = golden_inputs(meta)
    y64 = O.forward(sd, meta["dilations"], x, cond, dtype=torch.float64)
    assert rel_err(y64, y_ref) <= 1e-5


@pytest.mark.parametrize("name,chunk", [("synth_cfg1", 1024), ("synth_cfg1", 37), ("synth_gcn_c6", 300)])
def test_oracle_streaming_equals_oneshot(name, chunk):
    """wrapper.py's cached padding: chunked == one-shot (bit-exact in the reference)."""
    meta, y_ref, sd = load_golden(name)
    x, cond = golden_inputs(meta)
    st = O.StreamState(sd, meta["dilations"], meta["B"])
    outs = [O.forward_chunk(sd, meta["dilations"], st, x[..., s:s + chunk], cond) for s in range(0, meta["T"], chunk)]
    assert rel_err(torch.cat(outs, -1), y_ref) <= 2e-6


def test_known_answers_from_reference_logs():
    """logs/model_report.txt:68,98,148,179 - parameter counts and receptive fields."""
    meta, _, sd = load_golden("ckpt_TCN_egfxset_20240229_002014_48kHz")
    assert meta["params"] == 17989 and meta["rf"] == 82743
    assert O.receptive_field(3, meta["dilations"]) == 82743
    meta, _, sd = load_golden("ckpt_GCN_3_egfxset_20240324_160003_48kHz_cond")
    assert meta["params"] == 61312 and meta["rf"] == 131587
    assert O.receptive_field(3, meta["dilations"]) == 131587


def test_impulse_response_of_one_block_is_the_kernel():
    """A single TCN block without FiLM fed a unit impulse returns PReLU(w[:, 0, ::-1 taps] + b) + res."""
    sd = O.build_state("TCN", 1, 4, 5, 0, seed=3)
    x = torch.zeros(1, 1, 64)
    x[0, 0, 10] = 1.0
    y = O.block_forward(sd, 0, 3, x, None)
    w = sd["blocks.0.conv.conv.weight"][:, 0]          # [C, k]
    b = sd["blocks.0.conv.conv.bias"]
    a = sd["blocks.0.act.weight"]
    for j in range(5):
        t = 10 + (4 - j) * 3
        pre = w[:, j] + b
        exp = torch.where(pre > 0, pre, a * pre) + (sd["blocks.0.res.weight"][:, 0, 0] if t == 10 else 0)
        assert torch.allclose(y[0, :, t], exp, atol=1e-6)


@pytest.mark.parametrize("name", ["synth_cfg1", "synth_tcn_nofilm", "synth_tcn_io2", "synth_gcn_c6"])
def test_c_oracle_matches_reference_golden(name):
    """The independent plain-C restatement (fp64 accumulation) against the reference's output."""
    from oracle import c_oracle
    meta, y_ref, sd = load_golden(name)
    x, cond = golden_inputs(meta)
    T = min(meta["T"], 1500)
    y = c_oracle.forward(sd, meta["dilations"], x[..., :T].contiguous(), cond)
    assert rel_err(y, y_ref[..., :T]) <= 1e-5


def test_postprocess_oracle_pinned_to_torchaudio():
    """oracle/post_oracle.py (fp64 recursion) against the reference's own lines run here with torchaudio
    (inference.py:70-78).  The reference's fp32 recursion of the 20 Hz biquad carries ~5e-4 of rounding noise, which
    bounds how tightly the two can agree; the coefficients must match bit for bit."""
    import torchaudio
    from oracle import post_oracle as P
    b, a = P.highpass_coeffs(48000)
    x = torch.zeros(1, 8)
    x[0, 0] = 1.0
    # impulse response of torchaudio's filter vs the restated coefficients (exact in fp32 for the first taps)
    h = torchaudio.functional.highpass_biquad(x, 48000, 20.0)[0]
    assert abs(float(h[0]) - float(b[0] / a[0])) < 1e-7
    torch.manual_seed(0)
    for rows, T, sr in ((16, 3000, 48000), (8, 20000, 48000), (4, 12000, 16000)):
        y = torch.randn(rows, 1, T) * torch.exp(-torch.arange(T) / 6000.0) + 0.03
        got = P.postprocess(y, sr)
        ref = P.postprocess_reference_fp32(y, sr)
        assert got.shape == ref.shape == (1, rows * T)
        assert float((got - ref).abs().max()) <= 2e-3
        assert abs(float(got.abs().max()) - 1.0) < 1e-6


def test_reference_fp32_biquad_amplifies_a_1e7_perturbation_beyond_the_parity_bar():
    """Why make_inference is pinned to the fp64 restatement of the post-processing and only loosely to the reference's own
    fp32 lines (inference.py:70-78): torchaudio's sequential fp32 recursion of the 20 Hz biquad (poles at |z| ~ 0.998) is
    not reproducible below ~5e-4 - perturbing its input by 1e-7 (less than the fp32 noise of ANY two implementations of
    the network forward, the reference's own CPU and CUDA paths included) moves its output by more than the 1e-4 bar,
    while the same filter evaluated in fp64 moves by the size of the perturbation.  A bit-faithful fp32 mode (the exact
    operation order of torchaudio's CPU loop was reconstructed and matches it bit for bit on identical input) therefore
    cannot bring a GPU forward closer to the reference's post-processed samples than this."""
    import torchaudio
    torch.manual_seed(1)
    x = torch.randn(16, 30000) * 0.2
    x = x / x.abs().max()
    x2 = x + 1e-7 * torch.randn_like(x)
    y, y2 = (torchaudio.functional.highpass_biquad(v, 48000, 20.0) for v in (x, x2))
    d32 = float((y - y2).abs().max() / y.abs().max())
    z, z2 = (torchaudio.functional.highpass_biquad(v.double(), 48000, 20.0) for v in (x, x2))
    d64 = float((z - z2).abs().max() / z.abs().max())
    assert d32 > 1e-4, d32
    assert d64 < 2e-6, d64
    # the reference's fp32 result is itself that far from the exact (fp64) evaluation of its own filter
    assert 1e-4 < float((y.double() - z).abs().max() / z.abs().max()) < 2e-3
