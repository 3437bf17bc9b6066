"""Runs the climate stack once on the bench planet (ncu target for the sweep kernels)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from planet_heightmap_generation_b200 import climate as cl  # noqa: E402
from planet_heightmap_generation_b200.engine import DeviceMesh  # noqa: E402
from planet_heightmap_generation_b200.sphere import synthetic_elevation  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
inp = bench.Inputs(cells)
elev = synthetic_elevation(inp.xyz, bench.SEED, 0.3)
dm = DeviceMesh(inp.mesh, inp.xyz)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    cl.computeClimate(dm, elev, inp.pio, inp.r_plate, bench.SEED)
print("climate done")
