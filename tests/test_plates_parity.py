"""Plate pipeline on the hi-res mesh: projectCoarsePlates, smoothAndReconnectPlates, buildSuperPlates
through the C ABI against the oracle's sequential restatement — integer fields bit-exact, super-plate tables bit-exact."""
import numpy as np
import pytest

from planet_heightmap_generation_b200 import plates as pl
from planet_heightmap_generation_b200.engine import DeviceMesh
from tests.conftest import make_planet

_coarse_cache = {}


def coarse_inputs(oracle, seed, P, NC=20000):
    """Stand-in for generateCoarsePlates (js/coarse-plates.js:19-39): the coarse Fibonacci mesh of buildSphere(NC, 0.75,
    makeRng(seed + 137)) with P plates grown as nearest-seed regions; plate id = seed region id."""
    key = (seed, P, NC)
    if key not in _coarse_cache:
        from oracle.mesh_hull import build_sphere_from_points
        cmesh, cxyz = build_sphere_from_points(oracle.fibonacci_sphere(NC, 0.75, seed + 137))
        rng = np.random.default_rng(int(seed) + 17 * P)
        seeds = rng.choice(NC, P, replace=False)
        pts = cxyz.reshape(-1, 3).astype(np.float64)
        warp = 1.0 + 0.3 * rng.random(P)
        crp = seeds[np.argmax((pts @ pts[seeds].T) * warp, axis=1)].astype(np.int32)
        crp[seeds] = seeds
        _coarse_cache[key] = (cmesh, cxyz, crp, [int(s) for s in seeds])
    return _coarse_cache[key]


def plate_tables(seeds, seed):
    rng = np.random.default_rng(seed)
    vec, ocean, dens = {}, set(), {}
    for s in seeds:
        p = rng.normal(size=3)
        vec[s] = {"pole": [float(v) for v in p / np.linalg.norm(p)], "omega": float((0.5 + rng.random() * 1.5) * (1 if rng.random() < 0.5 else -1))}
        if rng.random() < 0.6:
            ocean.add(s)
        dens[s] = float(2.4 + rng.random())
    return vec, ocean, dens


@pytest.mark.parametrize("n,P,seed", [(3000, 12, 7), (20000, 40, 42), (20000, 80, 3)])
def test_project_smooth_super_match_oracle(backend, oracle, n, P, seed):
    mesh, xyz, nd, elev = make_planet(oracle, n)
    cmesh, cxyz, crp, seeds = coarse_inputs(oracle, seed, P)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    want = oracle.project_coarse_plates(mesh, xyz, cmesh, cxyz, crp, seed, P)
    got = pl.projectCoarsePlates(dm, xyz, cmesh, cxyz, crp, seed, P)
    assert np.array_equal(got, want)
    assert np.array_equal(pl.projectCoarsePlates(dm, xyz, cmesh, cxyz, crp, seed, None), oracle.project_coarse_plates(mesh, xyz, cmesh, cxyz, crp, seed, None))

    before = want.copy()
    oracle.smooth_and_reconnect_plates(mesh, want, seeds, 3)
    pl.smoothAndReconnectPlates(dm, got, seeds, 3)
    assert (want != before).any()
    assert np.array_equal(got, want)

    vec, ocean, dens = plate_tables(seeds, seed)
    o_super, o_table = oracle.build_super_plates(mesh, want, {s: dict(isOcean=s in ocean, pole=vec[s]["pole"], omega=vec[s]["omega"], density=dens[s]) for s in seeds})
    res = pl.buildSuperPlates(dm, got, seeds, vec, ocean, dens)
    assert res["numSuperPlates"] == len(o_table) >= 2
    assert np.array_equal(res["r_superPlate"], o_super)
    for sp, t in o_table.items():
        assert tuple(res["superPlateVec"][sp]["pole"]) == tuple(t["pole"])
        assert res["superPlateVec"][sp]["omega"] == t["omega"]
        assert (sp in res["superPlateIsOcean"]) == t["isOcean"]
        assert res["superPlateDensity"][sp] == t["density"]
    dm.close()


def test_reconnect_repairs_fragments(backend, oracle):
    """Plates torn into fragments (salt-and-pepper noise + a severed strip): the largest component survives, orphans are
    re-assigned in the reference's scan / FIFO order; seed protection applies when r_plate[seed] == seed."""
    mesh, xyz, nd, elev = make_planet(oracle, 6000)
    n = mesh.numRegions
    rng = np.random.default_rng(9)
    seeds = [int(s) for s in rng.choice(n, 9, replace=False)]
    pts = xyz.reshape(-1, 3).astype(np.float64)
    rp = np.asarray(seeds, np.int32)[np.argmax(pts @ pts[seeds].T, axis=1)]
    noisy = rng.random(n) < 0.12
    rp[noisy] = rng.choice(seeds, int(noisy.sum()))
    rp[seeds] = seeds
    band = np.abs(pts[:, 2] - 0.2) < 0.03
    rp[band] = seeds[0]
    dm = DeviceMesh(mesh, xyz, lib=backend)
    for passes in (0, 1, 3):
        want, got = rp.copy(), rp.copy()
        oracle.smooth_and_reconnect_plates(mesh, want, seeds, passes)
        pl.smoothAndReconnectPlates(dm, got, seeds, passes)
        assert np.array_equal(got, want), passes
    dm.close()


def test_super_plates_edge_cases(backend, oracle):
    """Plates without a plateVec entry / density, and a rejected r_plate id."""
    from planet_heightmap_generation_b200 import PlanetB200Error
    mesh, xyz, nd, elev = make_planet(oracle, 3000)
    cmesh, cxyz, crp, seeds = coarse_inputs(oracle, 5, 16)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    rp = oracle.project_coarse_plates(mesh, xyz, cmesh, cxyz, crp, 5, 16)
    vec, ocean, dens = plate_tables(seeds, 5)
    for s in seeds[:5]:
        vec[s] = None
    dens[seeds[1]] = None
    table = {s: dict(isOcean=s in ocean, pole=None if vec[s] is None else vec[s]["pole"], omega=0.0 if vec[s] is None else vec[s]["omega"],
                     density=dens[s]) for s in seeds}
    o_super, o_table = oracle.build_super_plates(mesh, rp, table)
    res = pl.buildSuperPlates(dm, rp, seeds, {s: v for s, v in vec.items() if v is not None}, ocean, {s: d for s, d in dens.items() if d is not None})
    assert np.array_equal(res["r_superPlate"], o_super)
    for sp, t in o_table.items():
        assert tuple(res["superPlateVec"][sp]["pole"]) == tuple(t["pole"]) and res["superPlateVec"][sp]["omega"] == t["omega"]
        assert res["superPlateDensity"][sp] == t["density"]
    with pytest.raises(PlanetB200Error):
        pl.buildSuperPlates(dm, rp, seeds[:-1], vec, ocean, dens)
    dm.close()


@pytest.mark.gpu
def test_plate_pipeline_device_pointers_1M(cuda_lib, oracle):
    """1M cells, device-resident r_plate: same bits as the oracle."""
    import torch
    from planet_heightmap_generation_b200.sphere import sphere_points
    dm = DeviceMesh.from_points(sphere_points(1_000_000, 0.75, 42.0), lib=cuda_lib)
    from planet_heightmap_generation_b200.mesh import SphereMesh
    mesh = SphereMesh.from_csr(dm.adjOffset, dm.adjList)
    cmesh, cxyz, crp, seeds = coarse_inputs(oracle, 42, 40)
    want = oracle.project_coarse_plates(mesh, dm.r_xyz, cmesh, cxyz, crp, 42, 40)
    got = torch.empty(dm.numRegions, dtype=torch.int32, device="cuda")
    pl.projectCoarsePlates(dm, None, cmesh, cxyz, crp, 42, 40, out=got)
    assert np.array_equal(got.cpu().numpy(), want)
    oracle.smooth_and_reconnect_plates(mesh, want, seeds, 3)
    pl.smoothAndReconnectPlates(dm, got, seeds, 3)
    assert np.array_equal(got.cpu().numpy(), want)
    vec, ocean, dens = plate_tables(seeds, 42)
    o_super, o_table = oracle.build_super_plates(mesh, want, {s: dict(isOcean=s in ocean, pole=vec[s]["pole"], omega=vec[s]["omega"], density=dens[s]) for s in seeds})
    res = pl.buildSuperPlates(dm, got, seeds, vec, ocean, dens)
    assert np.array_equal(res["r_superPlate"].cpu().numpy(), o_super)
    assert res["numSuperPlates"] == len(o_table)
    dm.close()


@pytest.mark.parametrize("seed,P,cont,variety,cover,nc", [(42, 80, 4, 0.0, 0.3, 20000), (7, 12, 3, 0.5, 0.3, 4000), (123456, 40, 6, 1.0, 0.45, 6000),
                                                           (3, 5, 8, 0.0, 0.2, 3000), (99, 150, 2, 0.3, 0.3, 5000)])
def test_generate_coarse_plates_matches_oracle(backend, oracle, seed, P, cont, variety, cover, nc):
    """generateCoarsePlates = coarse buildSphere + generatePlates + assignOceanLand: plate ids, seed order, Euler poles,
    ocean flags and the worker's densities, bit for bit."""
    from planet_heightmap_generation_b200.sphere import park_miller
    mesh, xyz, nd, elev = make_planet(oracle, 3000)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    got = pl.generateCoarsePlates(dm, seed, P, cont, variety, cover, numCoarse=nc)
    want = oracle.generate_coarse_plates(seed, P, cont, variety, cover, n_coarse=nc)
    cm = got["coarseMesh"]
    assert np.array_equal(cm.adjOffset, want["coarseMesh"].adjOffset) and np.array_equal(cm.adjList, want["coarseMesh"].adjList)
    assert (got["coarse_xyz"].view(np.uint32) == want["coarse_xyz"].view(np.uint32)).all()
    assert got["coarsePlateSeeds"] == want["coarsePlateSeeds"] and len(got["coarsePlateSeeds"]) == min(P, nc + 1)
    assert np.array_equal(got["coarse_r_plate"], want["coarse_r_plate"])
    assert got["coarsePlateIsOcean"] == want["coarsePlateIsOcean"]
    for s in want["coarsePlateSeeds"]:
        assert got["coarsePlateVec"][s]["pole"] == want["coarsePlateVec"][s]["pole"]
        assert got["coarsePlateVec"][s]["omega"] == want["coarsePlateVec"][s]["omega"]
        d = park_miller(s + 777, 2)
        assert got["plateDensity"][s] == (3.0 + d[0] * 0.5 if s in want["coarsePlateIsOcean"] else 2.4 + d[1] * 0.5)
    # every region belongs to a seeded plate and the land share is near the requested coverage
    area = np.bincount(got["coarse_r_plate"], minlength=nc + 1)
    assert area[got["coarsePlateSeeds"]].sum() == nc + 1
    # the coarse mesh is an ordinary mesh of the same context: the projection runs straight from it
    r_plate = pl.projectCoarsePlates(dm, xyz, cm, got["coarse_xyz"], got["coarse_r_plate"], seed, P)
    assert np.array_equal(r_plate, oracle.project_coarse_plates(mesh, xyz, want["coarseMesh"], want["coarse_xyz"], want["coarse_r_plate"], seed, P))
    cm.close()
    dm.close()


@pytest.mark.parametrize("seed", [1, 2, 5, 11, 99, 1234, 31337, 777777])
@pytest.mark.parametrize("nc", [3000, 20000])
def test_projection_start_independent_over_seeds(emu_lib, oracle, seed, nc):
    """projectCoarsePlates (js/coarse-plates.js:51-117): the reference warm-starts every region's greedy walk on the coarse mesh
    from the previous region's result, the kernel starts every walk independently (DESIGN.md §5c argues that strict steepest
    ascent of the dot product on a Delaunay mesh ends at the same coarse site from any start).  The oracle keeps the warm start:
    identical r_plate over seeds and coarse resolutions is the evidence for that argument in f64, beyond the fixed-seed cases."""
    mesh, xyz, nd, elev = make_planet(oracle, 12000)
    P = 6 + seed % 37
    cmesh, cxyz, crp, seeds = coarse_inputs(oracle, seed, P, nc)
    dm = DeviceMesh(mesh, xyz, lib=emu_lib)
    want = oracle.project_coarse_plates(mesh, xyz, cmesh, cxyz, crp, seed, P)
    got = pl.projectCoarsePlates(dm, xyz, cmesh, cxyz, crp, seed, P)
    assert np.array_equal(got, want)
    assert len(np.unique(got)) >= min(P, 6) - 1
    dm.close()
