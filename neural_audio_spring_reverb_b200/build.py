"""Build libnasr_b200.so (sm_100a only) in-tree with nvcc.

`python -m neural_audio_spring_reverb_b200.build` or `build_native()`.
The library is the product's only compute path; there is no fallback.
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libnasr_b200.so"
SOURCES = ["engine.cu", "generic_block.cu", "first_block.cu", "fold.cu", "tc_block.cu", "ring_block.cu", "toep_block.cu", "postproc.cu", "analysis.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libnasr_b200.so cannot be built")
    return nvcc


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "nasr_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_native(force: bool = False, verbose: bool = False, extra_flags=(), out: Path = None) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library.
    `extra_flags` / `out` build a tuning variant next to the default library (dev use)."""
    variant = out is not None
    out = Path(out) if variant else LIB_PATH
    if not variant and not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    obj_dir = PKG_DIR / "build" / (out.stem if variant else "default")
    obj_dir.mkdir(parents=True, exist_ok=True)
    srcs = [s for s in SOURCES if (CSRC / s).exists()]
    procs = []
    for s in srcs:
        obj = obj_dir / (s + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", str(CSRC / s), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for s, obj, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{log}")
        if verbose and log:
            print(log)
        objs.append(str(obj))
    tmp = Path(str(out) + ".tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    path = build_native(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_flags=extra,
                        out=Path(outs[0]) if outs else None)
    print(path)
