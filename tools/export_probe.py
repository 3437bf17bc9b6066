"""Times pb_export_map (exportMap, js/planet-mesh.js:1752-1950) on the GPU: device-pointer mode, CUDA events on torch's current
stream (the stream the library launches on in that mode), L2 flushed between runs.  Prints one JSON line.
usage: python tools/export_probe.py [cells=1000000] [width=4096] [steps=5]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planet_heightmap_generation_b200 import planet_mesh as pm  # noqa: E402
from planet_heightmap_generation_b200.engine import DeviceMesh  # noqa: E402
from planet_heightmap_generation_b200.sphere import synthetic_elevation  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
width = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dm = DeviceMesh.build_sphere(cells, 0.75, 42, device=0)
elev_h = synthetic_elevation(dm.r_xyz, 42, 0.3)
kop_h = (np.random.default_rng(3).integers(1, 31, elev_h.size) * (elev_h > 0)).astype(np.uint8)
elev, kop = torch.from_numpy(elev_h).cuda(), torch.from_numpy(kop_h).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for etype in ("biome", "heightmap"):
    for _ in range(2):
        px = pm.exportMapPixels(dm, etype, width, elev, kop)
    times = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        px = pm.exportMapPixels(dm, etype, width, elev, kop)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    out[etype] = {"ms": [round(t, 3) for t in times], "ms_median": float(np.median(times))}
dm.profile_start()
px, side = pm.exportMapPixels(dm, "biome", width, elev, kop, want_sides=True)
prof = dm.profile_stop()
host = pm.exportMapPixels(dm, "biome", width, elev_h, kop_h)            # host-pointer path, same bytes
ms = out["biome"]["ms_median"]
H = width // 2
# algorithmic bytes: raster 8 B/side + 4 B per covered pixel (owner) + memset 4 B/px; shade 4 + 4 + 4 B/px (owner, side → region, RGBA)
alg = 8 * 3 * dm.numTriangles + (4 + 4 + 12) * width * H
print(json.dumps({"what": "pb_export_map", "cells": cells + 1, "width": width, "height": H, "steps": steps, "types": out,
                  "mpix_per_s": width * H / ms / 1e3, "cells_per_s": (cells + 1) / ms * 1e3,
                  "algorithmic_bytes": alg, "algorithmic_gbs": alg / ms / 1e6,
                  "covered": float((side >= 0).float().mean()), "host_pointer_path_identical": bool((host == px.cpu().numpy()).all()),
                  "kernels": sorted(prof, key=lambda k: -k["ms"])[:8]}))
