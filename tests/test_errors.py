"""Error behaviour of the widened C ABI: bad arguments fail with a status and a message (pb_last_error), never a crash."""
import ctypes as C

import numpy as np
import pytest

from planet_heightmap_generation_b200 import PlanetB200Error
from planet_heightmap_generation_b200 import plates as pl
from planet_heightmap_generation_b200.engine import DeviceMesh
from planet_heightmap_generation_b200.mesh import SphereMesh
from planet_heightmap_generation_b200.worker import PlanetWorker
from tests.conftest import make_planet


def test_plate_pipeline_rejects_inconsistent_input(backend, oracle):
    mesh, xyz, nd, elev = make_planet(oracle, 3000)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    coarse = SphereMesh.from_csr(mesh.adjOffset[:201].copy(), mesh.adjList[:mesh.adjOffset[200]].copy())     # rows point outside 0..199
    with pytest.raises(PlanetB200Error, match="out of range"):
        pl.projectCoarsePlates(dm, xyz, coarse, xyz[:600], np.zeros(200, np.int32), 1, 10)
    with pytest.raises(ValueError):
        pl.projectCoarsePlates(dm, xyz, mesh, xyz[:30], np.zeros(mesh.numRegions, np.int32), 1, 10)
    rp = np.full(mesh.numRegions, -3, np.int32)
    with pytest.raises(PlanetB200Error, match="negative"):
        pl.smoothAndReconnectPlates(dm, rp, [1, 2], 1)
    with pytest.raises(PlanetB200Error):
        pl.generateCoarsePlates(dm, 1, 0, 3)                       # numPlates must be >= 1
    with pytest.raises(PlanetB200Error):
        pl.generateCoarsePlates(dm, 1, 8, 3, numCoarse=2)          # fewer than 4 coarse points
    dm.close()


def test_import_and_colour_entries_reject_bad_sizes(backend, oracle):
    mesh, xyz, nd, elev = make_planet(oracle, 3000)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    out = np.empty(mesh.numRegions, np.float32)
    px = np.zeros(4, np.uint8)
    rc = dm.lib.dll.pb_sample_heightmap(dm._mesh, px.ctypes.data, 0, 4, out.ctypes.data)
    assert rc != 0 and b"empty" in dm.lib.dll.pb_last_error()
    rc = dm.lib.dll.pb_region_colors(dm._mesh, 9, elev.ctypes.data, None, out.ctypes.data)
    assert rc != 0 and b"colour mode" in dm.lib.dll.pb_last_error()
    rc = dm.lib.dll.pb_mesh_get_triangles(dm._mesh, None, None)
    assert rc != 0
    dm.close()


def test_worker_reports_errors_as_messages(backend):
    w = PlanetWorker(lib=backend)
    r = w.onmessage(dict(cmd="generate", N=2000, P=8, jitter=0.75, nMag=0.4, numContinents=2))      # no seed
    assert r["type"] == "error" and "seed" in r["message"]
    r = w.onmessage(dict(cmd="importHeightmap", N=2000, jitter=0.75, grayscale=np.zeros(10, np.uint8), imageWidth=4, imageHeight=4, seed=1))
    assert r["type"] == "error" and "imageWidth" in r["message"]
    assert w.onmessage(dict(cmd="computeClimate"))["type"] == "error"
    w.close()
