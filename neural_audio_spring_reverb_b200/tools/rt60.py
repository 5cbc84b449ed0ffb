"""`rt60` action: reverberation time of an impulse response by Schroeder's method
(src/neural_audio_spring_reverb/tools/rt60.py:9-91) on the device.

The reference reads the wav file, squares it, integrates backwards (np.cumsum of the reversed power), converts to dB
relative to the total energy and takes the first samples below -5 dB and -60 dB.  The integration and the two searches
run as one chunked reverse scan on the B200 (`nasr_rt60`, csrc/analysis.cu); the matplotlib plot of the reference is
not part of this package."""
from pathlib import Path

import numpy as np
import torch

from .. import _native


def rt60_of(h, sample_rate: float, decay_db: float = 60.0, device=None) -> dict:
    """RT60 of the impulse response `h` (array-like or tensor, any shape; flattened):
    dict(rt60 [s], i_5db, i_decay, i_nz), -1 for a crossing that does not exist (then rt60 = 0, rt60.py:71-72)."""
    if not isinstance(h, torch.Tensor):
        h = torch.as_tensor(np.asarray(h, dtype=np.float32))
    dev = torch.device(device) if device is not None else (h.device if h.is_cuda else torch.device("cuda:0"))
    return _native.rt60(h.to(dev, torch.float32).reshape(-1), sample_rate, decay_db)


def measure_rt60(args) -> float:
    """args.input = wav file of an impulse response (rt60.py:38-39 reads it with scipy.io.wavfile and casts the
    samples to float32 as they are); args.device optional.  Prints and returns the RT60 in seconds."""
    from scipy.io import wavfile

    print("RT60 measurement")
    fs, data = wavfile.read(args.input)
    x = np.asarray(data).astype("float32")
    res = rt60_of(x, float(fs), 60.0, getattr(args, "device", None))
    print(f"{Path(args.input).stem}: the RT60 is {res['rt60'] * 1000:.0f} ms")
    return res["rt60"]
