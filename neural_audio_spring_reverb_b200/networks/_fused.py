"""Host side shared by TCN and GCN: packs the module's parameters into the
weight blob of the C ABI, owns the nasr_engine handle and routes forward /
streaming calls to it."""
from __future__ import annotations

import os
from typing import List, Optional

import torch
from torch import Tensor

from .. import _native


def _async_device_forward() -> bool:
    """NASR_ASYNC=1: `model(x_cuda)` only enqueues (no stream sync, no range check); the caller polls
    `model.saturated()`.  Default: the call waits for its result and silently redoes it on the fp32 kernels when an
    activation left the fp16 range of the tensor-core path - the reference forward has no input-range limit
    (tcn.py:150-155), so neither has the drop-in."""
    return os.environ.get("NASR_ASYNC", "0").lower() not in ("0", "", "false", "no")


def _path_from_env() -> int:
    v = os.environ.get("NASR_PATH", "auto").lower()
    if v in ("auto", "0", ""):
        return _native.PATH_AUTO
    if v in ("fp32", "ffma", "1"):
        return _native.PATH_FP32
    if v in ("tc", "gather", "2"):
        return _native.PATH_TC_GATHER
    raise ValueError(f"NASR_PATH={v!r}: expected 'auto', 'fp32' or 'tc'")


class FusedNetMixin:
    """Expects: self.blocks, self.out_net, self.in_ch, self.out_ch, self.kernel_size,
    self.cond_dim, self.channels, self.dilations and the class attributes below."""

    _nasr_arch: int = _native.ARCH_TCN
    _nasr_final_tanh: bool = False

    # ---- construction shared by TCN and GCN -------------------------------
    def _nasr_build(self, block_cls, n_blocks: int, n_channels: int, dilation_growth: int,
                    in_ch: int, out_ch: int, kernel_size: int, cond_dim: int) -> None:
        """Attributes and sub-module names of the reference networks (tcn.py:105-148,
        gcn.py:91-138): channels, dilations = growth**i, strides, blocks, out_net."""
        import torch.nn as nn

        self.in_ch, self.out_ch = in_ch, out_ch
        self.kernel_size, self.cond_dim = kernel_size, cond_dim
        self.channels = [n_channels] * n_blocks
        self.dilations = [dilation_growth**i for i in range(n_blocks)]
        self.n_blocks = n_blocks
        self.strides = [1] * n_blocks
        widths = [in_ch] + self.channels
        self.blocks = nn.ModuleList(
            block_cls(widths[i], widths[i + 1], kernel_size, self.dilations[i], 1, cond_dim)
            for i in range(n_blocks))
        self.out_net = nn.Conv1d(self.channels[-1], out_ch, kernel_size=(1,), stride=(1,), bias=False)

    def calc_receptive_field(self) -> int:
        """Receptive field in samples (tcn.py:157-164, gcn.py:149-160)."""
        assert all(s == 1 for s in self.strides)
        assert self.dilations[0] == 1
        return self.kernel_size + (self.kernel_size - 1) * sum(self.dilations[1:])

    # ---- engine life cycle -------------------------------------------------
    def _nasr_layers(self):
        """The fused layers in execution order (TCN/GCN: the blocks; WaveNet flattens its stacks)."""
        return list(self.blocks)

    def _nasr_has_film(self) -> bool:
        return hasattr(self._nasr_layers()[0], "film")

    def _nasr_tensors(self) -> List[Tensor]:
        """Parameters/buffers in the order nasr_weight_count() documents."""
        out: List[Tensor] = []
        film = self._nasr_has_film()
        for blk in self._nasr_layers():
            out += [blk.conv.conv.weight, blk.conv.conv.bias]
            if film:
                out += [blk.film.adaptor.weight, blk.film.adaptor.bias, blk.film.bn.weight,
                        blk.film.bn.bias, blk.film.bn.running_mean, blk.film.bn.running_var]
            if self._nasr_arch == _native.ARCH_TCN:
                out.append(blk.act.weight)
            out.append(blk.res.weight)
        out.append(self.out_net.weight)
        return out

    def weight_blob(self) -> torch.Tensor:
        """Flat fp32 CPU tensor: what nasr_engine_create consumes (and what a
        multi-GPU launch broadcasts once over NCCL)."""
        return torch.cat([t.detach().reshape(-1).to(device="cpu", dtype=torch.float32)
                          for t in self._nasr_tensors()])

    def load_weight_blob(self, blob: Tensor) -> None:
        """Inverse of weight_blob(): fill the parameters/buffers from a flat fp32 tensor
        (what the other ranks do after the one-time NCCL broadcast)."""
        tensors = self._nasr_tensors()
        need = sum(t.numel() for t in tensors)
        if blob.numel() != need:
            raise ValueError(f"weight blob has {blob.numel()} floats, expected {need}")
        off = 0
        with torch.no_grad():
            for t in tensors:
                n = t.numel()
                t.copy_(blob[off:off + n].reshape(t.shape).to(device=t.device, dtype=t.dtype))
                off += n

    def _engine(self) -> "_native.Engine":
        """The nasr_engine for the current parameters. Rebuilt when a parameter/buffer was
        modified in place (tensor version counters) or replaced (.to(), load_state_dict)."""
        cache = self.__dict__.get("_nasr_cache")
        if cache is not None:
            tensors, stamp, eng = cache
            if stamp == sum(t._version for t in tensors) + tensors[0].data_ptr():
                return eng
        tensors = self._nasr_tensors()
        dev = tensors[0].device
        if dev.type != "cuda":
            raise RuntimeError(
                f"{type(self).__name__}: parameters are on {dev}; the forward runs only on a B200 "
                "through libnasr_b200 (no CPU fallback) - move the model with .to('cuda:N')")
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        old = self.__dict__.pop("_nasr_cache", None)
        if old is not None:
            old[2].close()
        blob = self.weight_blob().numpy()
        layers = self._nasr_layers()
        eng = _native.Engine(
            arch=self._nasr_arch, n_blocks=len(layers), in_ch=self.in_ch, out_ch=self.out_ch,
            n_channels=layers[0].out_ch, kernel_size=self.kernel_size, cond_dim=self.cond_dim,
            has_film=self._nasr_has_film(), final_tanh=self._nasr_final_tanh,
            dilations=[l.dilation for l in layers], weights=blob, device=index, path=_path_from_env(),
            bn_eps=float(layers[0].film.bn.eps) if self._nasr_has_film() else 1e-5)
        stamp = sum(t._version for t in tensors) + tensors[0].data_ptr()
        self.__dict__["_nasr_cache"] = (tensors, stamp, eng)
        self.__dict__["_nasr_stream_B"] = None
        return eng

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() may replace the parameter tensors: drop the engine
        self.release_engine()
        return super()._apply(fn, *args, **kwargs)

    def release_engine(self) -> None:
        cache = self.__dict__.pop("_nasr_cache", None)
        if cache is not None:
            cache[2].close()

    # ---- argument checks shared by forward / forward_chunk -----------------
    def _nasr_check(self, x: Tensor, cond: Optional[Tensor]):
        assert x.ndim == 3  # (batch_size, in_ch, samples) - tcn.py:151
        if self.training:
            raise RuntimeError(f"{type(self).__name__} is inference-only here: call .eval() "
                               "(training / backward are out of scope of the B200 engine)")
        if x.shape[1] != self.in_ch:
            raise ValueError(f"expected {self.in_ch} input channels, got {x.shape[1]}")
        B = x.shape[0]
        if self._nasr_has_film() and self.cond_dim > 0:
            if cond is None:
                raise ValueError("cond is required when cond_dim > 0")
            if cond.ndim != 2 or cond.shape[0] != B or cond.shape[1] != self.cond_dim:
                raise ValueError(f"cond must be [{B}, {self.cond_dim}], got {tuple(cond.shape)}")
        else:
            cond = None
        return B, x.shape[2], cond

    def _nasr_run(self, x: Tensor, cond: Optional[Tensor], chunk: bool) -> Tensor:
        B, T, cond = self._nasr_check(x, cond)
        eng = self._engine()
        dev = torch.device("cuda", eng.device)
        if x.is_cuda:
            if x.device != dev:
                raise RuntimeError(f"input is on {x.device} but the model is on {dev}")
            xc = x.detach().contiguous().float()
            y = torch.empty((B, self.out_ch, T), device=dev, dtype=torch.float32)
            stream = torch.cuda.current_stream(dev).cuda_stream
            cptr = 0
            cc = None
            if cond is not None:
                cc = cond.detach().to(device=dev, dtype=torch.float32).contiguous()
                cptr = cc.data_ptr()
            # streaming: the same (unmodified) cond tensor chunk after chunk needs no new fold.  The tensor is kept alive
            # so that its address cannot be recycled for another cond behind our back.
            key = (eng, cptr, cc._version if cc is not None else 0, B, stream)
            if not (chunk and self.__dict__.get("_nasr_cond_key") == key):
                eng.set_cond(cptr, B, stream)
                self.__dict__["_nasr_cond_key"] = key if chunk else None
                self.__dict__["_nasr_cond_ref"] = cc if chunk else None
            if T > 0:
                if chunk:
                    eng.forward_chunk(xc.data_ptr(), y.data_ptr(), B, T, stream)
                elif self.__dict__.get("_nasr_async", None) or (self.__dict__.get("_nasr_async", None) is None
                                                                and _async_device_forward()):
                    eng.forward(xc.data_ptr(), y.data_ptr(), B, T, stream)
                else:
                    eng.forward_checked(xc.data_ptr(), y.data_ptr(), B, T, stream)
            return y
        # host tensors: the engine copies in, runs, copies out (e2e path of make_inference)
        xc = x if (x.dtype == torch.float32 and x.is_contiguous() and not x.requires_grad) \
            else x.detach().contiguous().float()
        cptr = 0
        if cond is not None:
            cc = cond if (cond.dtype == torch.float32 and cond.is_contiguous() and cond.device.type == "cpu"
                          and not cond.requires_grad) \
                else cond.detach().to(device="cpu", dtype=torch.float32).contiguous()
            cptr = cc.data_ptr()
        stream = torch.cuda.current_stream(dev).cuda_stream
        if chunk:
            xd = xc.to(dev, non_blocking=True)
            return self._nasr_run(xd, cond, True).cpu()
        # pinned result: the engine's last kernel writes it directly (zero-copy), see nasr_forward_host
        y = torch.empty((B, self.out_ch, T), dtype=torch.float32, pin_memory=True)
        self.__dict__["_nasr_cond_key"] = None          # the host path folds its own cond
        eng.forward_host(xc.data_ptr(), cptr, y.data_ptr(), B, T, stream)
        return y

    def set_async(self, flag: Optional[bool]) -> None:
        """True: device-tensor forwards only enqueue work (poll `saturated()` yourself); False: every forward checks
        the fp16-range flag and falls back to the fp32 kernels by itself; None: follow NASR_ASYNC (default: checked)."""
        self.__dict__["_nasr_async"] = flag

    def saturated(self) -> bool:
        """True if the last forward had to clamp an activation to the fp16 range of the tensor-core path
        (|a| * 64 > 65504).  Checked forwards (the default, and every host-tensor forward) have then already been
        redone on the fp32 kernels; after an asynchronous forward (`set_async(True)` / NASR_ASYNC=1, or a streaming
        chunk) the result is not trustworthy - run with NASR_PATH=fp32.  Synchronises."""
        eng = self._engine()
        dev = torch.device("cuda", eng.device)
        return eng.saturated(torch.cuda.current_stream(dev).cuda_stream)

    # ---- streaming (reference wrapper.py:14-57 semantics) ------------------
    def reset_stream(self, batch_size: int = 1) -> None:
        """Zero the per-block input history, like freshly built PaddingCached buffers."""
        eng = self._engine()
        dev = torch.device("cuda", eng.device)
        eng.stream_reset(batch_size, torch.cuda.current_stream(dev).cuda_stream)
        self.__dict__["_nasr_stream_B"] = batch_size

    def forward_chunk(self, x: Tensor, cond: Optional[Tensor] = None) -> Tensor:
        """Process the next chunk of a stream; history of the last (k-1)*d input
        samples of every block is carried across calls."""
        self._engine()   # may rebuild (parameters changed in place) and then forgets the stream: decide afterwards
        if self.__dict__.get("_nasr_stream_B") != x.shape[0]:
            self.reset_stream(x.shape[0])
        return self._nasr_run(x, cond, True)

    def block_forward(self, index: int, x: Tensor, cond: Optional[Tensor] = None) -> Tensor:
        """One block on its own, x [B, Cin, T] -> [B, C, T] (TCNBlock/GCNBlock.forward)."""
        if self.training:
            raise RuntimeError("inference-only: call .eval()")
        eng = self._engine()
        dev = torch.device("cuda", eng.device)
        blk = self._nasr_layers()[index]
        assert x.ndim == 3 and x.shape[1] == blk.in_ch
        B, _, T = x.shape
        xc = x.detach().to(dev, torch.float32).contiguous()
        y = torch.empty((B, blk.out_ch, T), device=dev, dtype=torch.float32)
        stream = torch.cuda.current_stream(dev).cuda_stream
        cptr = 0
        if self._nasr_has_film() and self.cond_dim > 0:
            cc = cond.detach().to(device=dev, dtype=torch.float32).contiguous()
            cptr = cc.data_ptr()
        self.__dict__["_nasr_cond_key"] = None
        eng.set_cond(cptr, B, stream)
        if T > 0:
            eng.block_forward(index, xc.data_ptr(), y.data_ptr(), B, T, stream)
        return y if x.is_cuda else y.cpu()

    def train(self, mode: bool = True):
        # nn.Module.train is kept so that .eval() works; forward refuses training mode
        return super().train(mode)
