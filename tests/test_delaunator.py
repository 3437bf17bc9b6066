"""Delaunator-ordered sphere mesh (pb_mesh_create_delaunator, csrc/pb_delaunator.h): the reference's own neighbour order
(js/sphere-mesh.js:174-186 with delaunator@5.0.1).  Checked three ways: the numbering against an independent plain-Python
restatement of the published algorithm (oracle/delaunator_ref.py), the triangulation against qhull (oracle/mesh_hull.py: every
row must be a rotation of the checker's row), and the half-edge invariants of the arrays handed to the renderer.  The
algorithm runs on the host, so the same checks hold for both builds of the library."""
import numpy as np
import pytest

from tests.conftest import assert_bit_equal


def _points(oracle, n, jitter=0.75, seed=42):
    return oracle.fibonacci_sphere(n, jitter, seed)


@pytest.mark.parametrize("n,jitter,seed", [(12, 0.75, 1), (200, 0.0, 5), (700, 0.75, 42), (3000, 1.0, 7)])
def test_numbering_matches_python_restatement(backend, oracle, n, jitter, seed):
    from oracle.delaunator_ref import build_sphere_delaunator
    from planet_heightmap_generation_b200.engine import DeviceMesh
    pts = _points(oracle, n, jitter, seed)
    dm = DeviceMesh.from_points(pts, lib=backend, order="delaunator")
    tri, half, off, adj, adj_t = build_sphere_delaunator(pts)
    assert_bit_equal(dm.adjOffset, off, "adjOffset")
    assert_bit_equal(dm.adjList, adj, "adjList (row start = first side seen, js/sphere-mesh.js:102-106)")
    t, h = dm.trianglesAndHalfedges()
    assert_bit_equal(t, tri, "triangles")
    assert_bit_equal(h, half, "halfedges")
    assert_bit_equal(dm.adjTriList(), adj_t, "adjTriList")
    dm.close()


def test_same_triangulation_as_qhull_and_invariants(backend, oracle):
    from oracle.mesh_hull import build_sphere_from_points
    from planet_heightmap_generation_b200.engine import DeviceMesh
    pts = _points(oracle, 20000)
    dm = DeviceMesh.from_points(pts, lib=backend, order="delaunator")
    mesh, _ = build_sphere_from_points(pts)
    assert_bit_equal(dm.adjOffset, mesh.adjOffset, "degrees")
    starts_differ = 0
    for r in range(mesh.numRegions):
        a = dm.adjList[dm.adjOffset[r]:dm.adjOffset[r + 1]]
        b = mesh.adjList[mesh.adjOffset[r]:mesh.adjOffset[r + 1]]
        k = np.nonzero(b == a[0])[0]
        assert len(k) == 1 and (np.roll(b, -k[0]) == a).all(), f"row {r} is not a rotation of the checker's row"
        starts_differ += k[0] != 0
    assert starts_differ > mesh.numRegions // 2, "Delaunator's row starts differ from the canonical ones for most rows"
    t, h = dm.trianglesAndHalfedges()
    S = t.size
    assert S == 3 * (2 * mesh.numRegions - 4)
    assert (h >= 0).all() and (h[h] == np.arange(S)).all(), "every side has a twin"
    nxt = np.where(np.arange(S) % 3 == 2, np.arange(S) - 2, np.arange(S) + 1)
    assert (t[nxt] == t[h]).all(), "twin sides run in opposite directions"
    first = np.full(mesh.numRegions, S, np.int64)
    np.minimum.at(first, t, np.arange(S))
    assert (dm.adjList[dm.adjOffset[:-1]] == t[nxt[first]]).all(), "row start = end vertex of the first side of the region"
    dm.close()


def test_pipeline_parity_on_delaunator_mesh(backend, oracle):
    """The whole chain on the reference-ordered mesh: order-dependent stages (fills, BFS payloads, flood ties, f32 sums) follow
    the mesh's row order in the engine exactly as in the oracle."""
    from planet_heightmap_generation_b200 import climate as cl
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200.mesh import SphereMesh
    from planet_heightmap_generation_b200.sphere import synthetic_elevation, synthetic_plate_tables
    from planet_heightmap_generation_b200.terrain_post import runPostProcessing
    from tests.test_elevation_parity import _call
    from tests.test_terrain_post_parity import DEFAULT_SLIDERS
    pts = _points(oracle, 6000)
    dm = DeviceMesh.from_points(pts, lib=backend, order="delaunator")
    mesh = SphereMesh.from_csr(dm.adjOffset, dm.adjList)
    xyz = dm.r_xyz
    nd = oracle.neighbor_dist(mesh, xyz)
    elev0 = synthetic_elevation(xyz, 42, 0.3)
    r_plate, plates, seeds, r_super, sp = synthetic_plate_tables(xyz, elev0, 42)
    oe = oracle.Elevation(mesh, xyz)
    oe.assign(r_plate, plates, seeds, 42, 0.4, 42, 5, r_super, sp)
    got = _call(dm, xyz, r_plate, plates, seeds, 42, 0.4, 42, 5, r_super, sp)
    assert_bit_equal(got["r_elevation"], oe.get("r_elevation"), "r_elevation")
    want = oe.get("r_elevation")
    o_delta, o_ocean = oracle.run_post_processing(mesh, xyz, want, DEFAULT_SLIDERS, nd, 42, oe.get("hotspot"))
    e = got["r_elevation"].copy()
    runPostProcessing(dm, xyz, e, DEFAULT_SLIDERS, nd, 42, got["debugLayers"]["hotspot"])
    assert_bit_equal(e, want, "runPostProcessing")
    pio = {p for p, v in plates.items() if v["isOcean"]}
    oc = oracle.Climate(mesh, xyz)
    o_k = oc.run_all(want, pio, r_plate, 42)
    k = np.empty(mesh.numRegions, np.uint8)
    cl.computeClimate(dm, e, pio, r_plate, 42, 0.0, 0.0, 0.3, out_koppen=k)
    assert_bit_equal(k, o_k, "r_koppen")
    dm.close()
