"""TEST INFRASTRUCTURE ONLY - CPU restatement of the post-processing in the reference's
make_inference (src/neural_audio_spring_reverb/inference.py:70-78).  Only tests/, __graft_entry__.smoke()
and bench.py's CPU legs may import this; the product path never does.

    pred /= pred.abs().max()
    pred = torchaudio.functional.highpass_biquad(pred, sample_rate, 20)
    pred = pred.view(-1).unsqueeze(0);  pred /= pred.abs().max()

The filter itself lives in a third-party dependency that is not under /root/reference: torchaudio
(un-pinned in the reference's pyproject.toml; the author's runs used 2.1.2, this image has 2.11).
Its published algorithm, restated here:
  * highpass_biquad (torchaudio/functional/filtering.py): w0 = 2 pi f / sr, alpha = sin(w0) / (2 Q), Q = 0.707,
    b = [(1 + cos w0) / 2, -(1 + cos w0), (1 + cos w0) / 2], a = [1 + alpha, -2 cos w0, 1 - alpha], all computed
    in the waveform's dtype (fp32);
  * lfilter: coefficients normalised by a[0]; o[n] = sum_k b[k] x[n-k] - a[1] o[n-1] - a[2] o[n-2] from a zero
    state per row; output clamped to [-1, 1] (clamp=True default).
`postprocess` evaluates the recursion in float64 (scipy.signal.lfilter); `postprocess_reference_fp32` is the
reference's own fp32 path (torchaudio on the CPU), kept to pin the restatement (tests/test_oracle.py) and to
show the fp32 recursion's noise floor (~1e-3 for the 20 Hz filter at 48 kHz).
"""
import math

import numpy as np
import torch


def highpass_coeffs(sample_rate, cutoff_freq=20.0, Q=0.707):
    """(b, a) as float32 numpy arrays, computed with the same fp32 tensor ops as torchaudio's highpass_biquad."""
    dtype = torch.float32
    cutoff = torch.as_tensor(cutoff_freq, dtype=dtype)
    q = torch.as_tensor(Q, dtype=dtype)
    w0 = 2 * math.pi * cutoff / sample_rate
    alpha = torch.sin(w0) / 2.0 / q
    b0 = (1 + torch.cos(w0)) / 2
    b1 = -1 - torch.cos(w0)
    b2 = b0
    a0 = 1 + alpha
    a1 = -2 * torch.cos(w0)
    a2 = 1 - alpha
    b = torch.stack([b0, b1, b2]).to(dtype).numpy()
    a = torch.stack([a0, a1, a2]).to(dtype).numpy()
    return b, a


def postprocess(pred, sample_rate, cutoff_freq=20.0):
    """pred [rows, 1, T] or [rows, T] float32 -> [1, rows*T] float32; recursion in float64."""
    from scipy.signal import lfilter
    x = torch.as_tensor(pred, dtype=torch.float32)
    x = x / x.abs().max()                                    # fp32 division, as the reference does
    b, a = highpass_coeffs(sample_rate, cutoff_freq)
    b64, a64 = b.astype(np.float64) / float(a[0]), a.astype(np.float64) / float(a[0])
    y = lfilter(b64, a64, x.numpy().astype(np.float64), axis=-1)
    y = np.clip(y.astype(np.float32), -1.0, 1.0)
    y = torch.from_numpy(y).reshape(1, -1)
    return y / y.abs().max()


def postprocess_reference_fp32(pred, sample_rate, cutoff_freq=20.0):
    """The reference's own lines on the CPU (torchaudio fp32 recursion)."""
    import torchaudio
    x = torch.as_tensor(pred, dtype=torch.float32).clone()
    x /= x.abs().max()
    x = torchaudio.functional.highpass_biquad(x, sample_rate, cutoff_freq)
    x = x.reshape(-1).unsqueeze(0)
    x /= torch.max(torch.abs(x))
    return x


def deconvolve_direct(sweep_output, inverse_filter):
    """tools/ir_model.py:128-146 of the reference, restated with numpy / scipy in float64 (direct convolution)."""
    from scipy import signal
    a = np.asarray(sweep_output, dtype=np.float64).reshape(-1).copy()
    b = np.asarray(inverse_filter, dtype=np.float64).reshape(-1).copy()
    a = a - np.mean(a)
    a /= np.max(np.abs(a))
    b /= np.max(np.abs(b))
    ir = signal.convolve(a, b, mode="full", method="direct")
    ir = ir - np.mean(ir)
    ir /= np.max(np.abs(ir))
    return ir
