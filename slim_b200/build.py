"""In-tree build of libslim.so (the CUDA engine + the C ABI) for sm_100a.

`python -m slim_b200.build` or `slim_b200.build.build()`.  The shared object lands in
slim_b200/lib/ so it travels with a snapshot of the repository (it is git-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libslim.so"
SOURCES = [CSRC / "engine.cu", CSRC / "api.cpp"]
HEADERS = [CSRC / "engine.h", CSRC / "gram.cuh", CSRC / "gram_batch.cuh", CSRC / "hybrid.cuh", CSRC / "predict.cuh", CSRC / "gather.cuh", CSRC / "fslim.cuh", PKG.parent / "include" / "slim.h", PKG.parent / "include" / "slim_b200.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--split-compile", "0",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared", "-ldl",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (needed to build slim_b200/lib/libslim.so)")


def stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not stale():
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    extra = os.environ.get("SLIMB200_NVCC_EXTRA", "").split()
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", str(LIB)] + [str(s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libslim.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
