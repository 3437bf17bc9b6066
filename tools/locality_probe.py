"""Experiment: how much faster is a smoothField sweep when cells are numbered along a space-filling curve?
Builds a Morton-ordered copy of the bench mesh (same graph, rows in the same neighbour order) and times both."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch  # noqa: E402
from planet_heightmap_generation_b200.climate_util import smoothField  # noqa: E402
from planet_heightmap_generation_b200.engine import DeviceMesh  # noqa: E402
from planet_heightmap_generation_b200.mesh import SphereMesh  # noqa: E402


def part1by2(x):
    x = x.astype(np.uint64) & np.uint64(0x1fffff)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1f00000000ffff)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def morton_perm(xyz):
    p = xyz.reshape(-1, 3).astype(np.float64)
    q = np.clip(((p + 1) * 0.5 * (1 << 20)).astype(np.int64), 0, (1 << 20) - 1)
    code = part1by2(q[:, 0]) | (part1by2(q[:, 1]) << np.uint64(1)) | (part1by2(q[:, 2]) << np.uint64(2))
    return np.argsort(code, kind="stable")


cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
passes = 100
mesh, xyz = bench.get_planet(cells)
N = mesh.numRegions
perm = morton_perm(xyz)                      # new index i holds old cell perm[i]
inv = np.empty(N, np.int64); inv[perm] = np.arange(N)
off, adj = mesh.adjOffset.astype(np.int64), mesh.adjList
deg = np.diff(off)
offP = np.zeros(N + 1, np.int64); np.cumsum(deg[perm], out=offP[1:])
# rows of the permuted mesh, same neighbour order
src_start = off[perm]
idx = np.repeat(src_start - offP[:-1], deg[perm]) + np.arange(offP[-1])
adjP = inv[adj[idx]].astype(np.int32)
meshP = SphereMesh.from_csr(offP.astype(np.int32), adjP)
xyzP = xyz.reshape(-1, 3)[perm].reshape(-1).copy()
span = np.abs(adjP - np.repeat(np.arange(N), deg[perm]))
span0 = np.abs(adj - np.repeat(np.arange(N), deg))
print(f"|nb - r| median/p90: original {np.median(span0):.0f}/{np.percentile(span0, 90):.0f}   morton {np.median(span):.0f}/{np.percentile(span, 90):.0f}")

field = np.random.default_rng(0).random(N).astype(np.float32)
for name, m, x, f in (("original", mesh, xyz, field), ("morton", meshP, xyzP, field[perm].copy())):
    dm = DeviceMesh(m, x)
    t = torch.from_numpy(f).cuda()
    smoothField(dm, t, 4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    smoothField(dm, t, passes)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:9s} {1e6 * dt / passes:8.1f} us per sweep  ({36.0 * N / (dt / passes) / 1e9:7.0f} GB/s algorithmic)")
    res = t.cpu().numpy()
    if name == "original":
        ref = res
    else:
        print("bit-identical after un-permuting:", bool((res.view(np.uint32) == ref[perm].view(np.uint32)).all()))
