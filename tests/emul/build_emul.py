"""TEST INFRASTRUCTURE ONLY.

Compiles the engine sources with g++ and -DPB_EMUL: every per-cell kernel functor is run by a
sequential loop and the ordered (sync-free dataflow) kernels are run in their logical order, so the
host orchestration and the per-cell logic can be checked against the oracle on a box without a GPU.
The python package never loads this library (see planet_heightmap_generation_b200/_lib.py); only
tests/ does, explicitly by path.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "planet_heightmap_generation_b200", "csrc")
SO = os.path.join(HERE, "libpb_hostemu_TESTONLY.so")


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    if not force and os.path.exists(SO) and all(os.path.getmtime(s) <= os.path.getmtime(SO) for s in srcs):
        return SO
    subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-DPB_EMUL",
                           "-shared", "-o", SO, os.path.join(CSRC, "planet_b200.cu")])
    return SO


if __name__ == "__main__":
    print(build(True))
