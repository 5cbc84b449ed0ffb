"""ORACLE - test infrastructure. ctypes front-end of the plain-C restatement
(oracle/nasr_oracle.c). Only tests/ and __graft_entry__ (build + smoke check) use it."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
LIB = HERE / "libnasr_oracle.so"


class _Desc(C.Structure):
    _fields_ = [("arch", C.c_int), ("n_blocks", C.c_int), ("in_ch", C.c_int), ("out_ch", C.c_int),
                ("n_channels", C.c_int), ("kernel_size", C.c_int), ("cond_dim", C.c_int),
                ("has_film", C.c_int), ("dilations", C.POINTER(C.c_int))]


def build() -> Path:
    src = HERE / "nasr_oracle.c"
    if not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "libnasr_oracle.so"], check=True, capture_output=True)
    return LIB


def blob_from_state(sd, n_blocks: int, gcn: bool) -> np.ndarray:
    parts = []
    for i in range(n_blocks):
        p = f"blocks.{i}."
        keys = ["conv.conv.weight", "conv.conv.bias"]
        if (p + "film.adaptor.weight") in sd:
            keys += ["film.adaptor.weight", "film.adaptor.bias", "film.bn.weight", "film.bn.bias",
                     "film.bn.running_mean", "film.bn.running_var"]
        if not gcn:
            keys.append("act.weight")
        keys.append("res.weight")
        parts += [sd[p + k].reshape(-1).float() for k in keys]
    parts.append(sd["out_net.weight"].reshape(-1).float())
    return torch.cat(parts).numpy()


def forward(sd, dilations, x: torch.Tensor, cond) -> torch.Tensor:
    lib = C.CDLL(str(build()))
    gcn = "blocks.0.act.weight" not in sd
    n = len(dilations)
    w0 = sd["blocks.0.conv.conv.weight"]
    C_ = sd["out_net.weight"].shape[1]
    dil = (C.c_int * n)(*[int(d) for d in dilations])
    has_film = "blocks.0.film.adaptor.weight" in sd
    cd = sd["blocks.0.film.adaptor.weight"].shape[1] if has_film else 0
    d = _Desc(int(gcn), n, w0.shape[1], sd["out_net.weight"].shape[0], C_, w0.shape[2], cd, int(has_film), dil)
    blob = np.ascontiguousarray(blob_from_state(sd, n, gcn), dtype=np.float32)
    xa = np.ascontiguousarray(x.numpy(), dtype=np.float32)
    B, _, T = xa.shape
    ca = np.zeros((B, max(cd, 1)), np.float32) if cond is None else np.ascontiguousarray(cond.numpy(), np.float32)
    y = np.empty((B, d.out_ch, T), np.float32)
    lib.nasr_oracle_forward.restype = C.c_int
    lib.nasr_oracle_forward.argtypes = [C.POINTER(_Desc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int64]
    rc = lib.nasr_oracle_forward(C.byref(d), blob.ctypes.data, xa.ctypes.data, ca.ctypes.data, y.ctypes.data, B, T)
    if rc != 0:
        raise MemoryError("nasr_oracle_forward")
    return torch.from_numpy(y)
