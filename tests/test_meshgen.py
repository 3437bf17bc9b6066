"""Mesh construction on the device: pb_triangulate_sphere / pb_mesh_create_from_points against
the CPU checker oracle/mesh_hull.py (qhull convex hull + the SphereMesh constructor's circulation), bit-exact CSR."""
import ctypes as C

import numpy as np
import pytest

from planet_heightmap_generation_b200 import _lib
from planet_heightmap_generation_b200.engine import DeviceMesh
from planet_heightmap_generation_b200.sphere import sphere_points


def _hull(xyz):
    from oracle.mesh_hull import build_sphere_from_points
    return build_sphere_from_points(xyz)


def _check_same(dm, mesh):
    assert np.array_equal(dm.adjOffset, mesh.adjOffset)
    assert np.array_equal(dm.adjList, mesh.adjList)


@pytest.mark.parametrize("n,seed", [(30, 3.0), (500, 11.0), (3000, 42.0), (20000, 7.0)])
def test_fibonacci_sphere_matches_hull(backend, n, seed):
    xyz = sphere_points(n, 0.75, seed)
    mesh, xyz = _hull(xyz)
    dm = DeviceMesh.from_points(xyz, lib=backend)
    _check_same(dm, mesh)
    assert dm.numEdges == 6 * (n + 1) - 12
    dm.close()


@pytest.mark.parametrize("n,jitter,seed", [(1, 0.75, 1.0), (1000, 0.0, 5.0), (3000, 0.75, 42.0), (20000, 0.75, 0.5), (70000, 1.0, 123456.0)])
def test_fibonacci_points_match_oracle(backend, oracle, n, jitter, seed):
    """generateFibonacciSphere on the device (jump-ahead Park–Miller, deterministic asin/sin/cos) vs the oracle's
    sequential restatement: bit-exact, pole vertex appended."""
    ctx = C.c_void_p()
    backend.check(backend.dll.pb_context_create(0, C.byref(ctx)))
    out = np.empty(3 * (n + 1), np.float32)
    backend.check(backend.dll.pb_generate_fibonacci_sphere(ctx, n, C.c_double(jitter), C.c_double(seed), out.ctypes.data))
    want = oracle.fibonacci_sphere(n, jitter, seed)
    assert want.shape[0] == 3 * (n + 1) and tuple(want[-3:]) == (0.0, 0.0, 1.0)
    assert (out.view(np.uint32) == want.view(np.uint32)).all()
    backend.dll.pb_context_destroy(ctx)


def test_build_sphere_on_device(backend, oracle):
    """buildSphere = points + triangulation, both on the device; equals oracle points + hull checker."""
    from planet_heightmap_generation_b200.sphere import build_sphere
    got = build_sphere(5000, 0.75, 42.0, lib=backend)
    mesh, xyz = _hull(oracle.fibonacci_sphere(5000, 0.75, 42.0))
    assert (got["r_xyz"].view(np.uint32) == xyz.view(np.uint32)).all()
    _check_same(got["mesh"], mesh)
    got["mesh"].close()


def test_irregular_point_sets_match_hull(backend):
    """Uniform random points and a dense cluster: the grid block has to grow for some stars."""
    rng = np.random.default_rng(5)
    p = rng.normal(size=(20000, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    q = rng.normal(size=(4000, 3)) * 0.05 + np.array([0.3, 0.4, 0.85])
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    for pts in (p, np.concatenate([p[:4000], q])):
        mesh, xyz = _hull(pts.astype(np.float32).reshape(-1))
        dm = DeviceMesh.from_points(xyz, lib=backend)
        _check_same(dm, mesh)
        dm.close()


def test_mesh_from_points_runs_the_hot_path(backend, oracle):
    """A mesh built on the device is an ordinary pb_mesh: smoothField over it equals the oracle on the hull mesh."""
    from planet_heightmap_generation_b200.climate_util import smoothField
    from planet_heightmap_generation_b200.sphere import synthetic_elevation
    xyz = sphere_points(3000, 0.75, 42.0)
    mesh, xyz = _hull(xyz)
    dm = DeviceMesh.from_points(xyz, lib=backend)
    f = synthetic_elevation(xyz, 42, 0.3)
    want = f.copy()
    oracle.smooth_field(mesh, want, 3)
    smoothField(dm, f, 3)
    assert (f.view(np.uint32) == want.view(np.uint32)).all()
    nd = dm.computeNeighborDist()
    assert (nd.view(np.uint32) == oracle.neighbor_dist(mesh, xyz).view(np.uint32)).all()
    dm.close()


def test_triangulate_sphere_entry_and_errors(backend):
    lib = backend
    ctx = C.c_void_p()
    lib.check(lib.dll.pb_context_create(0, C.byref(ctx)))
    xyz = sphere_points(2000, 0.75, 1.0)
    n = xyz.size // 3
    off = np.empty(n + 1, np.int32)
    adj = np.empty(6 * n - 12, np.int32)
    lib.check(lib.dll.pb_triangulate_sphere(ctx, n, xyz.ctypes.data, off.ctypes.data, adj.ctypes.data))
    mesh, _ = _hull(xyz)
    assert np.array_equal(off, mesh.adjOffset) and np.array_equal(adj, mesh.adjList)
    # duplicate points cannot form a closed triangulation
    bad = xyz.copy()
    bad[3:6] = bad[0:3]
    rc = lib.dll.pb_triangulate_sphere(ctx, n, bad.ctypes.data, off.ctypes.data, adj.ctypes.data)
    assert rc != 0 and b"Delaunay" in lib.dll.pb_last_error()
    rc = lib.dll.pb_triangulate_sphere(ctx, 3, xyz.ctypes.data, off.ctypes.data, adj.ctypes.data)
    assert rc != 0
    lib.dll.pb_context_destroy(ctx)


@pytest.mark.gpu
def test_million_cell_mesh_on_device(cuda_lib):
    """1M cells: closed triangulated sphere (Euler), symmetric, and identical to the hull checker."""
    xyz = sphere_points(1_000_000, 0.75, 42.0)
    dm = DeviceMesh.from_points(xyz, lib=cuda_lib)
    n = dm.numRegions
    assert dm.numEdges == 6 * n - 12
    deg = np.diff(dm.adjOffset)
    assert deg.min() >= 3 and deg.max() <= 32
    src = np.repeat(np.arange(n, dtype=np.int64), deg)
    fwd = np.sort(src * n + dm.adjList)
    rev = np.sort(dm.adjList.astype(np.int64) * n + src)
    assert np.array_equal(fwd, rev)
    mesh, _ = _hull(xyz)
    _check_same(dm, mesh)
    dm.close()


@pytest.mark.parametrize("n,seed", [(30, 3.0), (3000, 42.0), (20000, 7.0)])
def test_triangles_and_halfedges_match_hull(backend, n, seed):
    """SphereMesh.triangles / .halfedges rebuilt on the device from the CSR rows, triangle centres and triangle
    elevations (what the worker's replies carry for the renderer) against the hull checker."""
    xyz = sphere_points(n, 0.75, seed)
    mesh, xyz = _hull(xyz)
    dm = DeviceMesh.from_points(xyz, lib=backend)
    tri, half = dm.trianglesAndHalfedges()
    assert dm.numTriangles == mesh.numTriangles
    assert np.array_equal(tri, mesh.triangles) and np.array_equal(half, mesh.halfedges)
    assert np.array_equal(dm.adjTriList(), mesh.adjTriList)
    # half-edge invariants (js/sphere-mesh.js: s_end_r(s) == s_begin_r(halfedges[s]))
    nxt = np.where(np.arange(tri.size) % 3 == 2, np.arange(tri.size) - 2, np.arange(tri.size) + 1)
    assert np.array_equal(half[half], np.arange(tri.size)) and np.array_equal(tri[nxt], tri[half])
    p = xyz.reshape(-1, 3).astype(np.float64)
    t = tri.reshape(-1, 3)
    want_c = (((p[t[:, 0]] + p[t[:, 1]]) + p[t[:, 2]]) / 3).astype(np.float32).reshape(-1)
    assert (dm.generateTriangleCenters().view(np.uint32) == want_c.view(np.uint32)).all()
    from planet_heightmap_generation_b200.sphere import synthetic_elevation
    elev = synthetic_elevation(xyz, 5, 0.3)
    e = elev.astype(np.float64)
    want_e = (((e[t[:, 0]] + e[t[:, 1]]) + e[t[:, 2]]) / 3).astype(np.float32)
    assert (dm.computeTriangleElevations(elev).view(np.uint32) == want_e.view(np.uint32)).all()
    dm.close()


def test_extreme_density_contrast_uses_the_retry_path(emu_lib):
    """9 of 10 points inside a 0.6° cap: the stars of the sparse regions do not close within the normal block limit and go
    through StarRetryK (block grown to the whole grid).  Host emulation only so far — the retry kernel has not run on a GPU."""
    rng = np.random.default_rng(16663996)
    p = rng.normal(size=(5000, 3)) * 0.01 + np.array([0, 0, 1.0])
    q = rng.normal(size=(500, 3))
    pts = np.concatenate([p, q])
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    mesh, xyz = _hull(pts.astype(np.float32).reshape(-1))
    import os
    from planet_heightmap_generation_b200 import PlanetB200Error
    os.environ["PB_MESH_NO_RETRY"] = "1"
    try:
        with pytest.raises(PlanetB200Error, match="could not be closed"):      # the normal block limit is not enough here
            DeviceMesh.from_points(xyz, lib=emu_lib)
    finally:
        del os.environ["PB_MESH_NO_RETRY"]
    dm = DeviceMesh.from_points(xyz, lib=emu_lib)
    _check_same(dm, mesh)
    dm.close()
