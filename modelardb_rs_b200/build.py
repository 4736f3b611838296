"""Builds libmodelardb_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libmodelardb_cuda.so")
SOURCES = ["mdb_cuda.cu"]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",           # belt and braces: the kernels already use explicit *_rn intrinsics
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmodelardb_cuda.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(_HERE), "include", "modelardb_cuda.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """`defines` / `out`: an experimental variant beside the product library (tuning runs on the GPU box select it
    with MODELARDB_CUDA_LIB, see _native.py); the product is always the default build."""
    if not force and out == LIB and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + ["-o", out] + [
        os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return out


if __name__ == "__main__":
    import sys
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a for a in sys.argv[1:] if not a.startswith("-D")]
    print(build_library(force=True, verbose=True, defines=defs, out=os.path.join(_HERE, outs[0]) if outs else LIB))
