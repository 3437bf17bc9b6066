"""Builds libplanet_b200.so (CUDA, sm_100a) in-tree.  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libplanet_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# -fmad=false: no FMA contraction — FP64 results must round exactly like the reference's JS doubles.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".h", ".cuh"))]
    out.append(os.path.join(HERE, "..", "include", "planet_b200.h"))
    out.append(os.path.join(HERE, "..", "include", "pb_detmath.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    cmd = [NVCC, *NVCC_FLAGS, "-o", SO, os.path.join(CSRC, "planet_b200.cu")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.check_call(cmd, cwd=CSRC)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
