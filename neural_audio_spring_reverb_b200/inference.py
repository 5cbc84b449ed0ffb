"""`infer` entry point with the reference's signature and plumbing
(src/neural_audio_spring_reverb/inference.py:12-91); the forward runs on the
B200 engine."""
import os
import time
from pathlib import Path

import numpy as np
import torch

from .networks.model_utils import load_model_checkpoint


def _read_wav(path):
    """WAV -> (float32 tensor [channels, samples], sample_rate). torchaudio.load needs
    TorchCodec in recent releases; fall back to scipy (16/32-bit PCM, float)."""
    try:
        import torchaudio
        return torchaudio.load(path)
    except Exception:
        from scipy.io import wavfile
        sr, data = wavfile.read(path)       # 24-bit PCM (the reference's assets) arrives as int32, left-justified
        if data.dtype == np.int16:
            data = data.astype(np.float32) / 32768.0
        elif data.dtype == np.int32:
            data = data.astype(np.float32) / 2147483648.0
        elif data.dtype == np.uint8:
            data = (data.astype(np.float32) - 128.0) / 128.0
        else:
            data = data.astype(np.float32)
        if data.ndim == 1:
            data = data[None, :]
        else:
            data = data.T
        return torch.from_numpy(np.ascontiguousarray(data)), sr


def _write_wav(path, tensor, sample_rate):
    try:
        import torchaudio
        torchaudio.save(path, tensor, sample_rate=sample_rate)
    except Exception:
        from scipy.io import wavfile
        wavfile.write(path, int(sample_rate), tensor.squeeze(0).numpy().astype(np.float32))


def highpass_coeffs(sample_rate, cutoff_freq=20.0, Q=0.707):
    """(b, a) of torchaudio.functional.highpass_biquad, computed with the same fp32 tensor ops on the CPU
    (torchaudio/functional/filtering.py; the reference calls it at inference.py:73)."""
    import math
    dtype = torch.float32
    cutoff = torch.as_tensor(cutoff_freq, dtype=dtype)
    q = torch.as_tensor(Q, dtype=dtype)
    w0 = 2 * math.pi * cutoff / sample_rate
    alpha = torch.sin(w0) / 2.0 / q
    b0 = (1 + torch.cos(w0)) / 2
    b = torch.stack([b0, -1 - torch.cos(w0), b0]).to(dtype).numpy()
    a = torch.stack([1 + alpha, -2 * torch.cos(w0), 1 - alpha]).to(dtype).numpy()
    return b, a


def postprocess(pred: torch.Tensor, sample_rate: int, cutoff_freq: float = 20.0) -> torch.Tensor:
    """inference.py:70-78 on the device: normalise, 20 Hz high-pass biquad (lfilter semantics, clamp),
    flatten to [1, N], normalise - nasr_postprocess (fp64 chunked scan), result still on the GPU."""
    from . import _native
    b, a = highpass_coeffs(sample_rate, cutoff_freq)
    return _native.postprocess(pred, b, a, clamp=True).reshape(1, -1)


def make_inference(args) -> torch.Tensor:
    """Load args.checkpoint, run args.input (path or array-like [1, N]) through the model
    and return the processed signal [1, N] on the CPU (inference.py:12-91)."""
    model, _, _, config, rf, params = load_model_checkpoint(args)

    if isinstance(args.input, str):
        input, sample_rate = _read_wav(args.input)
    else:
        input = torch.as_tensor(args.input, dtype=torch.float32)

    batch_size = config["batch_size"]
    # the reference chops the signal into config["batch_size"] independent rows (inference.py:36-39)
    input = input.reshape(batch_size, 1, -1).to(args.device)

    if config["cond_dim"] > 0:
        c_values = [config.get(f"c{i}", 0.0) for i in range(config["cond_dim"])]
        c = torch.tensor(c_values, device=args.device).view(1, -1).repeat(batch_size, 1)
    else:
        c = None

    model.eval()
    with torch.no_grad():
        if input.is_cuda:
            torch.cuda.synchronize(input.device)
        start_time = time.perf_counter()
        pred = model(input, c)
        if pred.is_cuda:
            torch.cuda.synchronize(pred.device)  # the reference omits this; without it the RTF is a launch time
        duration = time.perf_counter() - start_time
        length_in_seconds = input.size(-1) / config["sample_rate"]
        print(f"RTF: {duration / length_in_seconds:.3f}")

    if not pred.is_cuda:
        pred = pred.to(args.device)       # host-tensor forward: the post-processing still runs on the device
    pred = postprocess(pred, config["sample_rate"], 20).cpu()

    if isinstance(args.input, str):
        file_name = Path(args.input).stem
        os.makedirs(f"{args.audio_dir}/processed", exist_ok=True)
        save_out = f"{args.audio_dir}/processed/{file_name}*{config['name']}.wav"
        _write_wav(save_out, pred, config["sample_rate"])
    return pred
