"""Streaming front-end with the semantics of the reference's cached-padding
export (src/neural_audio_spring_reverb/wrapper.py:14-57,117-133): feeding a
signal chunk by chunk gives the same output as one forward over the whole
signal, because every block keeps the last (k-1)*d input samples."""
from typing import Optional

import torch
from torch import Tensor


class CachedStream(torch.nn.Module):
    """``CachedStream(model)(x, cond)`` is the streaming counterpart of
    ``replace_modules(model); model(x, cond)`` in the reference: state is zero
    after construction / ``reset()`` and is resized on the first call for the
    batch size it sees (wrapper.py:21-27)."""

    def __init__(self, model):
        super().__init__()
        self.model = model.eval()
        self._batch: Optional[int] = None

    def reset(self, batch_size: Optional[int] = None) -> None:
        self._batch = batch_size
        if batch_size is not None:
            self.model.reset_stream(batch_size)

    def forward(self, x: Tensor, cond: Optional[Tensor] = None) -> Tensor:
        assert x.ndim == 3  # wrapper.py:24
        if self._batch != x.shape[0]:
            self.reset(x.shape[0])
        return self.model.forward_chunk(x, cond)


def neutone_forward(stream: CachedStream, x: Tensor, film1: Tensor, film2: Tensor, depth: Tensor) -> Tensor:
    """do_forward_pass of the Neutone wrapper (wrapper.py:117-133): x [B, T] mono buffers,
    cond = stack([FiLM1, FiLM2], 1) * depth."""
    cond = torch.stack([film1, film2], dim=1) * depth
    cond = cond.expand(x.shape[0], 2)
    return stream(x.unsqueeze(1), cond).squeeze(1)
