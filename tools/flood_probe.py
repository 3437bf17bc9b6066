"""Runs priorityFloodCarve once on the bench planet (ncu target for the heap-flood kernel)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from planet_heightmap_generation_b200.engine import DeviceMesh  # noqa: E402
from planet_heightmap_generation_b200.terrain_post import priorityFloodCarve  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
from planet_heightmap_generation_b200.sphere import synthetic_elevation  # noqa: E402
mesh, xyz = bench.get_planet(cells)
elev = synthetic_elevation(xyz, bench.SEED, 0.3)
dm = DeviceMesh(mesh, xyz)
ocean = (elev <= 0).astype(np.uint8)
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    e = elev.copy()
    t = time.time()
    priorityFloodCarve(dm, e, ocean, 0.5)
    print(f"priorityFloodCarve {cells} cells: {time.time() - t:.3f} s")
