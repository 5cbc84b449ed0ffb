"""Evaluation metrics of the reference's `eval` action on the device
(src/neural_audio_spring_reverb/eval.py:21-153).

The reference loads a test set, runs `pred = model(input, c)` per batch and scores it with torch.nn.L1Loss,
auraloss.time.ESRLoss, auraloss.time.DCLoss and a mel-scaled MultiResolutionSTFTLoss (eval.py:38-49,118-121), timing
each forward for a real-time factor (eval.py:103-116).  Datasets and wandb logging are outside this package; what is
here is the part behind the forward: MAE / ESR / DC as ONE fused reduction over (pred, target) on the B200
(`nasr_eval_metrics`, csrc/analysis.cu), so a prediction never has to leave the device to be scored.  ESR / DC follow
auraloss 0.4.0's published formulas (auraloss is not installed here: parity unpinned, see oracle/eval_oracle.py);
the MRSTFT metric is not built."""
import time
from typing import Dict, Iterable, Optional, Tuple

import torch
from torch import Tensor

from . import _native

METRICS = ("eval/mae", "eval/esr", "eval/dc")


def evaluate_batch(model, dry: Tensor, wet: Tensor, c: Optional[Tensor] = None) -> Tuple[Tensor, Dict[str, float]]:
    """One iteration of eval.py:98-121: forward + metrics, everything on the model's device.
    dry / wet: [B, 1, T]; returns (pred on the device, {"eval/mae", "eval/esr", "eval/dc"})."""
    dev = next(model.parameters()).device
    pred = model(dry.to(dev), None if c is None else c.to(dev))
    return pred, _native.eval_metrics(pred, wet.to(dev))


def evaluate_model(args, test_loader: Optional[Iterable] = None, model=None, config: Optional[dict] = None) -> Dict[str, float]:
    """The loop of eval.py:96-146 over an iterable of (dry, wet) batches.  The reference builds the loader from its
    datasets (eval.py:59-80), which this package does not ship: pass `test_loader` (and optionally an already loaded
    `model` / `config`; otherwise args.checkpoint is loaded like eval.py:28).  Returns the mean metrics and eval/rtf."""
    if test_loader is None:
        raise ValueError("evaluate_model needs a test_loader of (dry, wet) batches: the reference's datasets "
                         "(egfxset / springset, eval.py:59-80) are not part of this package")
    if model is None:
        from .networks.model_utils import load_model_checkpoint
        model, _, _, config, _, _ = load_model_checkpoint(args)
    config = config or {}
    dev = next(model.parameters()).device
    sample_rate = config.get("sample_rate", 48000)
    model.eval()
    results = {k: [] for k in METRICS}
    rtfs = []
    with torch.no_grad():
        for dry, wet in test_loader:
            c = None
            if config.get("cond_dim", getattr(model, "cond_dim", 0)) > 0:
                n = config.get("cond_dim", model.cond_dim)
                vals = [config.get(f"c{i}", 0.0) for i in range(n)]
                c = torch.tensor(vals, device=dev).view(1, -1).repeat(dry.shape[0], 1)      # eval.py:87-93
            t0 = time.perf_counter()
            pred, scores = evaluate_batch(model, dry, wet, c)
            torch.cuda.synchronize(dev)
            rtfs.append((time.perf_counter() - t0) / (dry.size(-1) / sample_rate))            # eval.py:110-116
            for k in METRICS:
                results[k].append(scores[k])
    out = {k: sum(v) / len(v) for k, v in results.items()}
    out["eval/rtf"] = sum(rtfs) / len(rtfs)
    return out
