"""CLI: the `infer`, `rtf`, `ir` and `rt60` actions of nafx-springrev
(src/neural_audio_spring_reverb/__main__.py:5-184) on the B200 engine.
The other actions (train, eval, download, wrap, ...) are not part of this package (the metrics of `eval` are
available as neural_audio_spring_reverb_b200.eval.evaluate_batch / evaluate_model for callers that bring a test set)."""
import argparse

import torch

ACTIONS = ["infer", "rtf", "ir", "rt60"]


def main(argv=None):
    parser = argparse.ArgumentParser(description="B200 TCN/GCN spring reverb inference")
    parser.add_argument("action", choices=ACTIONS)
    parser.add_argument("--audio_dir", type=str, default="audio")
    parser.add_argument("--device", type=str, default=None, help="cuda:N (default: cuda:0)")
    parser.add_argument("-c", "--checkpoint", type=str, default=None)
    parser.add_argument("-i", "--input", type=str, default=None)
    parser.add_argument("--duration", type=float, default=5.0, help="sweep length in seconds (ir)")
    args = parser.parse_args(argv)

    if args.device is None or args.device == "auto":
        if not torch.cuda.is_available():
            raise SystemExit("no CUDA device: this package has no CPU path")
        args.device = torch.device("cuda:0")
    else:
        args.device = torch.device(args.device)
    if args.action == "rt60":
        from .tools.rt60 import measure_rt60
        if args.input is None:
            parser.error("-i/--input is required for rt60")
        measure_rt60(args)
        return
    if args.checkpoint is None:
        parser.error("-c/--checkpoint is required")

    if args.action == "infer":
        from .inference import make_inference
        if args.input is None:
            parser.error("-i/--input is required for infer")
        make_inference(args)
    elif args.action == "rtf":
        from .rtf import measure_rtf
        measure_rtf(args)
    elif args.action == "ir":
        from .tools.ir_model import measure_model_ir
        measure_model_ir(args)


if __name__ == "__main__":
    main()
