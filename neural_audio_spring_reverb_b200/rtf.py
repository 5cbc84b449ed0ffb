"""`rtf` entry point (src/neural_audio_spring_reverb/rtf.py:5-33): one second of
noise through make_inference."""
import os

import torch

from .inference import make_inference


def setup_dummy_args(args):
    args.input = torch.randn(1, 48000).to(args.device)
    args.batch_size = 1  # ignored, as in the reference: make_inference reads config["batch_size"]
    return args


def measure_rtf(args):
    args = setup_dummy_args(args)
    os.makedirs(args.audio_dir, exist_ok=True)
    pred = make_inference(args)
    print("Inference completed. Output tensor shape:", pred.shape)
    return pred
