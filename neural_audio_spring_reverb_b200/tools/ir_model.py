"""`ir` action: impulse response of a trained model by sweep deconvolution
(src/neural_audio_spring_reverb/tools/ir_model.py:97-175).

The reference convolves the 240 000-sample model output with the 240 000-sample inverse filter with
scipy.signal.convolve(method="direct"): 5.8e10 multiply-adds on one host core.  Here the model runs on the B200
engine (make_inference) and the deconvolution is the same direct linear convolution as a hand-written fp64 kernel
(`nasr_convolve_full`, csrc/analysis.cu: ~15 ms for the 5 s sweep); the normalisations of ir_model.py:128-146 are done
in float64 as numpy does.  Plots (matplotlib) are not part of this package."""
from pathlib import Path

import numpy as np
import torch

from .. import _native
from ..inference import make_inference, _write_wav
from ..networks.model_utils import load_model_checkpoint
from .ir_signals import generate_reference


def deconvolve(sweep_output, inverse_filter, device) -> torch.Tensor:
    """ir_model.py:128-146 on the device, float64: remove the mean, peak-normalise both signals, full linear
    convolution, remove the mean, peak-normalise.  -> float64 tensor [N + M - 1] on `device`."""
    a = torch.as_tensor(np.asarray(sweep_output, dtype=np.float64).reshape(-1), device=device)
    b = torch.as_tensor(np.asarray(inverse_filter, dtype=np.float64).reshape(-1), device=device)
    a = a - a.mean()
    a = a / a.abs().max()
    b = b / b.abs().max()
    ir = _native.convolve_full(a, b)
    ir = ir - ir.mean()
    return ir / ir.abs().max()


def measure_model_ir(args) -> torch.Tensor:
    """args: checkpoint, device, duration, audio_dir (as the reference's CLI). Returns the IR [1, 2N-1] float32 on
    the CPU and saves it under audio_dir/IR_models/ like the reference."""
    print("Measure the impulse response of a trained model")
    model, _, _, config, _, _ = load_model_checkpoint(args)
    sweep, inverse_filter, _ = generate_reference(duration=args.duration, sample_rate=config["sample_rate"])
    args.input = sweep.reshape(1, -1)
    sweep_output = make_inference(args)                       # [1, N] CPU tensor (engine + device post-processing)
    ir = deconvolve(sweep_output.reshape(-1).numpy(), inverse_filter, args.device)
    ir_tensor = ir.to(torch.float32).unsqueeze(0).cpu()
    save_directory = Path(args.audio_dir) / "IR_models"
    save_directory.mkdir(parents=True, exist_ok=True)
    save_as = f"{save_directory}/{Path(args.checkpoint).stem}_IR.wav"
    _write_wav(save_as, ir_tensor, config["sample_rate"])
    print(f"Saved measured impulse response to {save_as}, sample rate: {config['sample_rate']}")
    return ir_tensor
