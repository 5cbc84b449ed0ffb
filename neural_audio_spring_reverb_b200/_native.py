"""ctypes binding of libnasr_b200.so (the C ABI in include/nasr_b200.h).

This is the only compute path of the package: if the library is missing or no
sm_100 device is present, calls raise — there is no CPU / eager fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

NASR_MAX_BLOCKS = 64
NASR_OK, NASR_ERR_INVALID, NASR_ERR_CUDA, NASR_ERR_STATE, NASR_ERR_NOMEM = 0, -1, -2, -3, -4
ARCH_TCN, ARCH_GCN = 0, 1
PATH_AUTO, PATH_FP32, PATH_TC_GATHER = 0, 1, 2

# NASR_LIB selects a tuning variant built with `python -m ...build --out=...` (dev use only)
LIB_PATH = Path(os.environ.get("NASR_LIB") or Path(__file__).resolve().parent / "libnasr_b200.so")

# every symbol include/nasr_b200.h declares
EXPORTS = (
    "nasr_weight_count", "nasr_engine_create", "nasr_engine_destroy", "nasr_last_error",
    "nasr_set_cond", "nasr_forward", "nasr_forward_checked", "nasr_sat_fallbacks", "nasr_forward_profiled", "nasr_saturated", "nasr_forward_host", "nasr_stream_reset",
    "nasr_forward_chunk", "nasr_block_forward", "nasr_workspace_bytes",
    "nasr_receptive_field", "nasr_launch_count", "nasr_block_path", "nasr_version",
    "nasr_postprocess", "nasr_postprocess_workspace_bytes", "nasr_eval_metrics", "nasr_eval_metrics_workspace_bytes",
    "nasr_rt60", "nasr_rt60_workspace_bytes", "nasr_convolve_full", "nasr_debug_ring_stamps", "nasr_debug_ring_steps", "nasr_debug_ring_plan", "nasr_debug_toep_stamps",
)


class ModelDesc(C.Structure):
    _fields_ = [
        ("arch", C.c_int32), ("n_blocks", C.c_int32), ("in_ch", C.c_int32), ("out_ch", C.c_int32),
        ("n_channels", C.c_int32), ("kernel_size", C.c_int32), ("cond_dim", C.c_int32),
        ("has_film", C.c_int32), ("final_tanh", C.c_int32), ("path", C.c_int32),
        ("bn_eps", C.c_float), ("dilations", C.c_int32 * NASR_MAX_BLOCKS),
    ]


_lib = None


def load_library():
    """Load libnasr_b200.so; raise RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m neural_audio_spring_reverb_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for the TCN/GCN forward."
        )
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_LOCAL)
    vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_int64, C.c_void_p
    lib.nasr_weight_count.restype = C.c_size_t
    lib.nasr_weight_count.argtypes = [C.POINTER(ModelDesc)]
    lib.nasr_engine_create.restype = i32
    lib.nasr_engine_create.argtypes = [C.POINTER(ModelDesc), f32p, C.c_size_t, i32, C.POINTER(vp)]
    lib.nasr_engine_destroy.restype = None
    lib.nasr_engine_destroy.argtypes = [vp]
    lib.nasr_last_error.restype = C.c_char_p
    lib.nasr_last_error.argtypes = [vp]
    lib.nasr_set_cond.restype = i32
    lib.nasr_set_cond.argtypes = [vp, f32p, i32, vp]
    lib.nasr_forward.restype = i32
    lib.nasr_forward.argtypes = [vp, f32p, f32p, i32, i64, vp]
    lib.nasr_forward_checked.restype = i32
    lib.nasr_forward_checked.argtypes = [vp, f32p, f32p, i32, i64, vp, C.POINTER(C.c_int)]
    lib.nasr_sat_fallbacks.restype = i64
    lib.nasr_sat_fallbacks.argtypes = [vp]
    lib.nasr_forward_profiled.restype = i32
    lib.nasr_forward_profiled.argtypes = [vp, f32p, f32p, i32, i64, vp, C.c_void_p]
    lib.nasr_saturated.restype = i32
    lib.nasr_saturated.argtypes = [vp, vp]
    lib.nasr_forward_host.restype = i32
    lib.nasr_forward_host.argtypes = [vp, f32p, f32p, f32p, i32, i64, vp]
    lib.nasr_stream_reset.restype = i32
    lib.nasr_stream_reset.argtypes = [vp, i32, vp]
    lib.nasr_forward_chunk.restype = i32
    lib.nasr_forward_chunk.argtypes = [vp, f32p, f32p, i32, i64, vp]
    lib.nasr_block_forward.restype = i32
    lib.nasr_block_forward.argtypes = [vp, i32, f32p, f32p, i32, i64, vp]
    lib.nasr_workspace_bytes.restype = C.c_size_t
    lib.nasr_workspace_bytes.argtypes = [vp, i32, i64]
    lib.nasr_receptive_field.restype = i64
    lib.nasr_receptive_field.argtypes = [vp]
    lib.nasr_launch_count.restype = i64
    lib.nasr_launch_count.argtypes = [vp]
    lib.nasr_block_path.restype = i32
    lib.nasr_block_path.argtypes = [vp, i32]
    lib.nasr_version.restype = C.c_char_p
    lib.nasr_version.argtypes = []
    lib.nasr_postprocess_workspace_bytes.restype = C.c_size_t
    lib.nasr_postprocess_workspace_bytes.argtypes = [i32, i64]
    lib.nasr_postprocess.restype = i32
    lib.nasr_postprocess.argtypes = [f32p, f32p, i32, i64, C.c_void_p, C.c_void_p, i32, vp, C.c_size_t, vp]
    lib.nasr_eval_metrics_workspace_bytes.restype = C.c_size_t
    lib.nasr_eval_metrics_workspace_bytes.argtypes = [i32]
    lib.nasr_eval_metrics.restype = i32
    lib.nasr_eval_metrics.argtypes = [f32p, f32p, i32, i64, vp, vp, C.c_size_t, vp]
    lib.nasr_rt60_workspace_bytes.restype = C.c_size_t
    lib.nasr_rt60_workspace_bytes.argtypes = [i64]
    lib.nasr_rt60.restype = i32
    lib.nasr_rt60.argtypes = [f32p, i64, C.c_double, C.c_double, vp, vp, C.c_size_t, vp]
    lib.nasr_convolve_full.restype = i32
    lib.nasr_convolve_full.argtypes = [vp, i64, vp, i64, vp, vp]
    lib.nasr_debug_ring_plan.restype = i32
    lib.nasr_debug_ring_plan.argtypes = [i32, i32, i32, i32, i64, i64, i32, C.c_void_p]
    lib.nasr_debug_ring_steps.restype = i32
    lib.nasr_debug_ring_steps.argtypes = [C.c_void_p]
    lib.nasr_debug_ring_stamps.restype = i32
    lib.nasr_debug_ring_stamps.argtypes = [C.c_void_p, i32]
    _lib = lib
    return lib


def _raise(lib, handle, code: int, what: str):
    msg = lib.nasr_last_error(handle).decode("utf-8", "replace")
    text = f"{what}: {msg} (nasr_status {code})"
    if code == NASR_ERR_INVALID:
        raise ValueError(text)
    if code == NASR_ERR_NOMEM:
        raise MemoryError(text)
    raise RuntimeError(text)


class Engine:
    """Owning wrapper of one nasr_engine handle (one device, not thread-safe)."""

    def __init__(self, *, arch: int, n_blocks: int, in_ch: int, out_ch: int, n_channels: int,
                 kernel_size: int, cond_dim: int, has_film: bool, final_tanh: bool,
                 dilations, weights, device: int, path: int = PATH_AUTO, bn_eps: float = 1e-5):
        import numpy as np

        self._lib = load_library()
        self._h = C.c_void_p()
        if n_blocks > NASR_MAX_BLOCKS:
            raise ValueError(f"n_blocks {n_blocks} > {NASR_MAX_BLOCKS}")
        d = ModelDesc()
        d.arch, d.n_blocks, d.in_ch, d.out_ch = arch, n_blocks, in_ch, out_ch
        d.n_channels, d.kernel_size, d.cond_dim = n_channels, kernel_size, cond_dim
        d.has_film, d.final_tanh, d.path, d.bn_eps = int(has_film), int(final_tanh), path, bn_eps
        for i, dil in enumerate(dilations):
            if dil > 2**31 - 1:
                raise ValueError("dilation does not fit int32")
            d.dilations[i] = int(dil)
        self.desc = d
        w = np.ascontiguousarray(weights, dtype=np.float32)
        rc = self._lib.nasr_engine_create(C.byref(d), w.ctypes.data, w.size, device, C.byref(self._h))
        if rc != NASR_OK:
            _raise(self._lib, None, rc, "nasr_engine_create")
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.nasr_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc != NASR_OK:
            _raise(self._lib, self._h, rc, what)

    def set_cond(self, cond_ptr: int, B: int, stream: int = 0):
        self._ck(self._lib.nasr_set_cond(self._h, cond_ptr or None, B, stream or None), "nasr_set_cond")

    def forward(self, x_ptr: int, y_ptr: int, B: int, T: int, stream: int = 0):
        self._ck(self._lib.nasr_forward(self._h, x_ptr, y_ptr, B, T, stream or None), "nasr_forward")

    def forward_checked(self, x_ptr: int, y_ptr: int, B: int, T: int, stream: int = 0) -> bool:
        """Forward that is valid for any input range: waits for the stream and redoes the call on the fp32 kernels if
        an activation left the fp16 range of the tensor-core path. Returns True if it had to."""
        redone = C.c_int(0)
        self._ck(self._lib.nasr_forward_checked(self._h, x_ptr, y_ptr, B, T, stream or None, C.byref(redone)),
                 "nasr_forward_checked")
        return bool(redone.value)

    def sat_fallbacks(self) -> int:
        return int(self._lib.nasr_sat_fallbacks(self._h))

    def forward_profiled(self, x_ptr: int, y_ptr: int, B: int, T: int, stream: int = 0):
        """-> list of per-block device milliseconds (CUDA events on `stream`)."""
        ms = (C.c_float * self.desc.n_blocks)()
        self._ck(self._lib.nasr_forward_profiled(self._h, x_ptr, y_ptr, B, T, stream or None, ms),
                 "nasr_forward_profiled")
        return list(ms)

    def saturated(self, stream: int = 0) -> bool:
        rc = self._lib.nasr_saturated(self._h, stream or None)
        if rc < 0:
            _raise(self._lib, self._h, rc, "nasr_saturated")
        return bool(rc)

    def forward_host(self, x_ptr: int, cond_ptr: int, y_ptr: int, B: int, T: int, stream: int = 0):
        self._ck(self._lib.nasr_forward_host(self._h, x_ptr, cond_ptr or None, y_ptr, B, T, stream or None),
                 "nasr_forward_host")

    def stream_reset(self, B: int, stream: int = 0):
        self._ck(self._lib.nasr_stream_reset(self._h, B, stream or None), "nasr_stream_reset")

    def forward_chunk(self, x_ptr: int, y_ptr: int, B: int, T: int, stream: int = 0):
        self._ck(self._lib.nasr_forward_chunk(self._h, x_ptr, y_ptr, B, T, stream or None), "nasr_forward_chunk")

    def block_forward(self, block: int, x_ptr: int, y_ptr: int, B: int, T: int, stream: int = 0):
        self._ck(self._lib.nasr_block_forward(self._h, block, x_ptr, y_ptr, B, T, stream or None),
                 "nasr_block_forward")

    def workspace_bytes(self, B: int, T: int) -> int:
        return int(self._lib.nasr_workspace_bytes(self._h, B, T))

    def receptive_field(self) -> int:
        return int(self._lib.nasr_receptive_field(self._h))

    def launch_count(self) -> int:
        return int(self._lib.nasr_launch_count(self._h))

    def block_path(self, block: int) -> int:
        return int(self._lib.nasr_block_path(self._h, block))


def postprocess(y, b_coeffs, a_coeffs, clamp: bool = True):
    """inference.py:70-78 on the device: y [rows, T] (or [rows, 1, T]) fp32 CUDA tensor -> same shape, peak-normalised,
    filtered with lfilter(a_coeffs, b_coeffs) per row, peak-normalised again. b_coeffs / a_coeffs: 3 floats each."""
    import numpy as np
    import torch

    lib = load_library()
    if not y.is_cuda:
        raise RuntimeError("nasr_postprocess runs on the device only (no CPU path)")
    shape = y.shape
    yc = y.detach().reshape(-1, shape[-1]).contiguous().float()
    rows, T = yc.shape
    out = torch.empty_like(yc)
    if rows == 0 or T == 0:
        return out.reshape(shape)
    ws_bytes = int(lib.nasr_postprocess_workspace_bytes(rows, T))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=y.device)
    b = np.ascontiguousarray(np.asarray(b_coeffs, dtype=np.float32).reshape(3))
    a = np.ascontiguousarray(np.asarray(a_coeffs, dtype=np.float32).reshape(3))
    with torch.cuda.device(y.device):
        stream = torch.cuda.current_stream(y.device).cuda_stream
        rc = lib.nasr_postprocess(yc.data_ptr(), out.data_ptr(), rows, T, b.ctypes.data, a.ctypes.data, int(clamp),
                                  ws.data_ptr(), ws_bytes, stream or None)
    if rc != NASR_OK:
        raise (ValueError if rc == NASR_ERR_INVALID else RuntimeError)(f"nasr_postprocess failed (nasr_status {rc})")
    return out.reshape(shape)


def _status(rc: int, what: str):
    if rc != NASR_OK:
        raise (ValueError if rc == NASR_ERR_INVALID else RuntimeError)(f"{what} failed (nasr_status {rc})")


def eval_metrics(pred, target):
    """eval.py:38-40,118-121 on the device: pred / target [B, C, T] (or [rows, T]) fp32 CUDA tensors ->
    {"eval/mae", "eval/esr", "eval/dc"} as Python floats (one fused reduction over both tensors)."""
    import torch

    lib = load_library()
    if not (pred.is_cuda and target.is_cuda):
        raise RuntimeError("nasr_eval_metrics runs on the device only (no CPU path)")
    if pred.shape != target.shape:
        raise ValueError(f"pred {tuple(pred.shape)} and target {tuple(target.shape)} differ in shape")
    p = pred.detach().reshape(-1, pred.shape[-1]).contiguous().float()
    t = target.detach().reshape(-1, target.shape[-1]).to(p.device).contiguous().float()
    rows, T = p.shape
    out = torch.empty(3, dtype=torch.float64, device=p.device)
    ws_bytes = int(lib.nasr_eval_metrics_workspace_bytes(rows))
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=p.device)
    with torch.cuda.device(p.device):
        stream = torch.cuda.current_stream(p.device).cuda_stream
        _status(lib.nasr_eval_metrics(p.data_ptr(), t.data_ptr(), rows, T, out.data_ptr(), ws.data_ptr(), ws_bytes,
                                      stream or None), "nasr_eval_metrics")
    mae, esr, dc = out.cpu().tolist()
    return {"eval/mae": mae, "eval/esr": esr, "eval/dc": dc}


def rt60(h, sample_rate: float, decay_db: float = 60.0):
    """tools/rt60.py:49-70 on the device: h = impulse response (1-D fp32 CUDA tensor) ->
    dict(rt60 seconds, i_5db, i_decay, i_nz)."""
    import torch

    lib = load_library()
    if not h.is_cuda:
        raise RuntimeError("nasr_rt60 runs on the device only (no CPU path)")
    hc = h.detach().reshape(-1).contiguous().float()
    n = hc.numel()
    if n < 1:
        raise ValueError("empty impulse response")
    out = torch.empty(4, dtype=torch.float64, device=hc.device)
    ws_bytes = int(lib.nasr_rt60_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=hc.device)
    with torch.cuda.device(hc.device):
        stream = torch.cuda.current_stream(hc.device).cuda_stream
        _status(lib.nasr_rt60(hc.data_ptr(), n, float(sample_rate), float(decay_db), out.data_ptr(), ws.data_ptr(),
                              ws_bytes, stream or None), "nasr_rt60")
    r, i5, idec, inz = out.cpu().tolist()
    return dict(rt60=r, i_5db=int(i5), i_decay=int(idec), i_nz=int(inz))


def convolve_full(a, b):
    """scipy.signal.convolve(a, b, method="direct") on the device: 1-D float64 CUDA tensors -> [n + m - 1] float64."""
    import torch

    lib = load_library()
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("nasr_convolve_full runs on the device only (no CPU path)")
    ac = a.detach().reshape(-1).contiguous().double()
    bc = b.detach().reshape(-1).to(ac.device).contiguous().double()
    n, m = ac.numel(), bc.numel()
    if n < 1 or m < 1:
        raise ValueError("empty operand")
    out = torch.empty(n + m - 1, dtype=torch.float64, device=ac.device)
    with torch.cuda.device(ac.device):
        stream = torch.cuda.current_stream(ac.device).cuda_stream
        _status(lib.nasr_convolve_full(ac.data_ptr(), n, bc.data_ptr(), m, out.data_ptr(), stream or None),
                "nasr_convolve_full")
    return out


def version() -> str:
    return load_library().nasr_version().decode()
