"""Analysis signals of the reference's IR measurement (src/neural_audio_spring_reverb/tools/ir_signals.py:53-116):
logarithmic sweep and its inverse filter.  Host-side numpy (O(N), a few milliseconds)."""
import numpy as np


def sweep_tone(sample_rate, duration, amplitude=-1.0, f0=20, f1=20000, inverse=False):
    """Exponential sine sweep f0 -> f1; `amplitude` in dB (ir_signals.py:53-84). inverse=True: time-reversed sweep
    with the 6 dB/octave envelope that makes sweep * inverse approximate an impulse."""
    gain = 10 ** (amplitude / 20)
    R = np.log(f1 / f0)
    t = np.arange(0, duration, 1.0 / sample_rate)
    out = np.sin((2.0 * np.pi * f0 * duration / R) * (np.exp(t * R / duration) - 1))
    if inverse:
        out = out[::-1] / np.exp(t * R / duration)
    return gain * out


def generate_reference(duration, sample_rate, decibels=-1.0, f0=20, with_reference=False):
    """(sweep, inverse_filter, reference IR) as ir_signals.py:87-116.  The reference IR (np.convolve of the two, O(N^2))
    is only computed on request: measure_model_ir discards it."""
    # the reference converts to linear gain here AND again inside sweep_tone (which expects dB): kept as is
    amplitude = 10 ** (decibels / 20)
    f1 = sample_rate / 2
    sweep = sweep_tone(sample_rate, duration, amplitude, f0=f0, f1=f1)
    inverse_filter = sweep_tone(sample_rate, duration, amplitude, f0=f0, f1=f1, inverse=True)
    reference = np.convolve(inverse_filter, sweep) if with_reference else None
    return sweep, inverse_filter, reference
