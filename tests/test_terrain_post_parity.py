"""Parity of the engine's terrain-post path against the oracle (bit-exact: every pass loads f32,
computes in FP64 in the reference's order and stores f32).  Runs on the host emulation of the
kernels in the CPU suite and on the CUDA library with -m gpu; both go through the C ABI."""
import numpy as np
import pytest

from tests.conftest import assert_bit_equal

DEFAULT_SLIDERS = dict(smoothing=0.10, glacialErosion=0.50, hydraulicErosion=0.50, thermalErosion=0.10,
                       ridgeSharpening=0.50, terrainWarp=0.75)


def _dm(backend, mesh, xyz):
    from planet_heightmap_generation_b200.engine import DeviceMesh
    return DeviceMesh(mesh, xyz, lib=backend)


def test_neighbor_dist(backend, oracle, planet_small):
    mesh, xyz, nd, _ = planet_small()
    assert_bit_equal(_dm(backend, mesh, xyz).computeNeighborDist(), nd, "neighborDist")


@pytest.mark.parametrize("passes", [0, 1, 2, 5])
def test_smooth_field(backend, oracle, planet_small, passes):
    from planet_heightmap_generation_b200.climate_util import smoothField
    mesh, xyz, nd, elev = planet_small()
    want = elev.copy(); oracle.smooth_field(mesh, want, passes)
    got = elev.copy(); smoothField(_dm(backend, mesh, xyz), got, passes)
    assert_bit_equal(got, want, "smoothField")


@pytest.mark.parametrize("hot", [False, True])
def test_warp_terrain(backend, oracle, planet_medium, hot):
    from planet_heightmap_generation_b200.terrain_post import warpTerrain
    mesh, xyz, nd, elev = planet_medium()
    hotspot = (np.abs(np.roll(elev, 17)) * 0.3).astype(np.float32) if hot else None
    want = elev.copy(); oracle.warp_terrain(mesh, want, xyz, 42, 0.75, hotspot)
    got = elev.copy(); warpTerrain(_dm(backend, mesh, xyz), got, xyz, 42, 0.75, hotspot)
    assert (want != elev).mean() > 0.5
    assert_bit_equal(got, want, "warpTerrain")


def test_smooth_sharpen_creep(backend, oracle, planet_small):
    from planet_heightmap_generation_b200 import terrain_post as tp
    mesh, xyz, nd, elev = planet_small()
    ocean = (elev <= 0).astype(np.uint8)
    dm = _dm(backend, mesh, xyz)
    for name, ofn, gfn, args in (
            ("smoothElevation", oracle.smooth_elevation, tp.smoothElevation, (3, 0.45)),
            ("sharpenRidges", oracle.sharpen_ridges, tp.sharpenRidges, (3, 0.04)),
            ("applySoilCreep", oracle.apply_soil_creep, tp.applySoilCreep, (3, 0.1125))):
        want = elev.copy(); ofn(mesh, want, ocean, *args)
        got = elev.copy(); gfn(dm, got, ocean, *args)
        assert (want != elev).any(), name
        assert_bit_equal(got, want, name)


@pytest.mark.parametrize("strength", [0.5, 0.85])
def test_priority_flood_carve(backend, oracle, planet_medium, strength):
    from planet_heightmap_generation_b200.terrain_post import priorityFloodCarve
    mesh, xyz, nd, elev = planet_medium()
    ocean = (elev <= 0).astype(np.uint8)
    want = elev.copy()
    o_drain, o_surf, o_open = oracle.priority_flood_carve(mesh, want, ocean, strength)
    got = elev.copy()
    drain, surf, openo = priorityFloodCarve(_dm(backend, mesh, xyz), got, ocean, strength, taps=True)
    assert ((o_surf - elev) > 1e-7).sum() > 20, "test planet must contain filled pits"
    assert_bit_equal(openo, o_open, "isOpenOcean")
    assert_bit_equal(drain, o_drain, "drainTo")
    assert_bit_equal(surf, o_surf, "surface")
    assert_bit_equal(got, want, "priorityFloodCarve elevation")


@pytest.mark.gpu
def test_priority_flood_carve_large(oracle, cuda_lib):
    """250k cells: deep flood trees (hundreds of hops), thousands of filled cells, a heap of thousands of entries."""
    from planet_heightmap_generation_b200.terrain_post import priorityFloodCarve
    from tests.conftest import make_planet
    mesh, xyz, nd, elev = make_planet(oracle, 250000)
    ocean = (elev <= 0).astype(np.uint8)
    want = elev.copy()
    o_drain, o_surf, o_open = oracle.priority_flood_carve(mesh, want, ocean, 0.85)
    got = elev.copy()
    drain, surf, openo = priorityFloodCarve(_dm(cuda_lib, mesh, xyz), got, ocean, 0.85, taps=True)
    assert_bit_equal(drain, o_drain, "drainTo")
    assert_bit_equal(surf, o_surf, "surface")
    assert_bit_equal(got, want, "priorityFloodCarve elevation")


@pytest.mark.parametrize("where", ["host", "device"])
def test_flood_option_host_and_device_match(backend, oracle, planet_medium, where):
    """Option flood=host (default: the serial heap pass on a host core, the rest on the GPU) and flood=device (the
    one-CTA kernel k_flood_heap; on the emulation the sequential form of that kernel) give the oracle's bits."""
    from planet_heightmap_generation_b200.terrain_post import priorityFloodCarve
    mesh, xyz, nd, elev = planet_medium()
    ocean = (elev <= 0).astype(np.uint8)
    want = elev.copy()
    o_drain, o_surf, o_open = oracle.priority_flood_carve(mesh, want, ocean, 0.5)
    dm = _dm(backend, mesh, xyz)
    dm.set_option("flood", where)
    got = elev.copy()
    drain, surf, openo = priorityFloodCarve(dm, got, ocean, 0.5, taps=True)
    assert_bit_equal(drain, o_drain, "drainTo")
    assert_bit_equal(surf, o_surf, "surface")
    assert_bit_equal(got, want, "elevation")
    with pytest.raises(Exception):
        dm.set_option("flood", "nowhere")


def test_flood_with_inland_sea_and_island(backend, oracle, planet_small):
    """Second-largest ocean component is an inland sea (not a flood seed); an island inside it is never flooded."""
    from planet_heightmap_generation_b200.terrain_post import priorityFloodCarve
    mesh, xyz, nd, elev = planet_small()
    p = xyz.reshape(-1, 3)
    # turn a land region into a lake with an island in the middle
    land = np.nonzero(elev > 0.2)[0]
    c = p[land[len(land) // 2]]
    d = p @ c
    elev[(d > 0.985) & (elev > 0)] = -0.05
    elev[d > 0.9985] = 0.3
    ocean = (elev <= 0).astype(np.uint8)
    want = elev.copy(); o = oracle.priority_flood_carve(mesh, want, ocean, 0.5)
    assert (o[2] != ocean).any(), "needs an ocean component that is not open ocean"
    got = elev.copy(); g = priorityFloodCarve(_dm(backend, mesh, xyz), got, ocean, 0.5, taps=True)
    assert_bit_equal(g[2], o[2], "isOpenOcean")
    assert_bit_equal(g[0], o[0], "drainTo")
    assert_bit_equal(got, want, "elevation")


ERODE_CASES = {
    "hydraulic": dict(hIters=6, K=0.0003, m=0.5, dt=1.0, tIters=0, talus=1.16, kThermal=0.015, gIters=0, g=0.0),
    "thermal": dict(hIters=0, K=0.0003, m=0.5, dt=1.0, tIters=3, talus=0.05, kThermal=0.15, gIters=0, g=0.0),
    "glacial": dict(hIters=0, K=0.0003, m=0.5, dt=1.0, tIters=0, talus=1.16, kThermal=0.015, gIters=4, g=1.0),
    "composite": dict(hIters=10, K=0.0003, m=0.5, dt=1.0, tIters=2, talus=0.08, kThermal=0.05, gIters=5, g=0.5),
    "m_not_half": dict(hIters=4, K=0.002, m=0.4, dt=1.0, tIters=0, talus=1.16, kThermal=0.015, gIters=0, g=0.0),
}


@pytest.mark.parametrize("case", list(ERODE_CASES))
def test_erode_composite(backend, oracle, planet_medium, case):
    from planet_heightmap_generation_b200.terrain_post import erodeComposite
    c = ERODE_CASES[case]
    mesh, xyz, nd, elev = planet_medium()
    ocean = (elev <= 0).astype(np.uint8)
    cap = 2 if c["hIters"] > 2 else -1
    want = elev.copy()
    o_t, o_f, o_l = oracle.erode_composite(mesh, want, xyz, ocean, c["hIters"], c["K"], c["m"], c["dt"], c["tIters"],
                                           c["talus"], c["kThermal"], c["gIters"], c["g"], nd, capture_iter=cap)
    got = elev.copy()
    taps = erodeComposite(_dm(backend, mesh, xyz), got, xyz, ocean, c["hIters"], c["K"], c["m"], c["dt"], c["tIters"],
                          c["talus"], c["kThermal"], c["gIters"], c["g"], nd, capture_iter=cap)
    assert (want != elev).mean() > 0.01, "case must change the terrain"
    if cap >= 0:
        assert_bit_equal(taps[2], o_l, "landCells order")
        assert_bit_equal(taps[0], o_t, "drainTarget (drainage receivers)")
        land = ocean == 0   # the reference also accumulates into ocean receivers but never reads them
        assert_bit_equal(taps[1][land], o_f[land], "flow")
    assert_bit_equal(got, want, f"erodeComposite[{case}]")


def test_erode_no_land_and_zero_iters(backend, oracle, planet_small):
    from planet_heightmap_generation_b200.terrain_post import erodeComposite
    mesh, xyz, nd, elev = planet_small()
    dm = _dm(backend, mesh, xyz)
    sea = -np.abs(elev) - 0.01
    got = sea.copy()
    erodeComposite(dm, got, xyz, np.ones_like(elev, np.uint8), 3, 0.0003, 0.5, 1.0, 1, 1.16, 0.015, 2, 0.5, nd)
    assert_bit_equal(got, sea, "all-ocean planet is untouched")
    got = elev.copy()
    erodeComposite(dm, got, xyz, (elev <= 0).astype(np.uint8), 0, 0.0003, 0.5, 1.0, 0, 1.16, 0.015, 0, 0.5, nd)
    assert_bit_equal(got, elev, "zero iterations is a no-op")


@pytest.mark.parametrize("hot", [False, True])
def test_run_post_processing_default_sliders(backend, oracle, planet_medium, hot):
    """BASELINE config 1 shape: default sliders ⇒ hIters 10, tIters 1, gIters 5, smooth 1, ridge 3, creep 3."""
    from planet_heightmap_generation_b200.terrain_post import runPostProcessing
    mesh, xyz, nd, elev = planet_medium()
    hotspot = (np.maximum(0, np.roll(elev, 5)) * 0.2).astype(np.float32) if hot else None
    want = elev.copy()
    o_delta, o_ocean = oracle.run_post_processing(mesh, xyz, want, DEFAULT_SLIDERS, nd, 42, hotspot)
    got = elev.copy()
    res = runPostProcessing(_dm(backend, mesh, xyz), xyz, got, DEFAULT_SLIDERS, nd, 42, hotspot)
    assert_bit_equal(res["r_isOcean"], o_ocean, "r_isOcean")
    assert_bit_equal(got, want, "runPostProcessing elevation")
    assert_bit_equal(res["dl_erosionDelta"], o_delta, "erosionDelta")
    assert [t["stage"] for t in res["postTiming"]] == ["Terrain warp", "Smoothing", "Erosion composite",
                                                      "Ridge sharpening", "Soil creep"]


def test_argument_errors(backend, planet_small):
    from planet_heightmap_generation_b200 import terrain_post as tp
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200._lib import PlanetB200Error
    mesh, xyz, nd, elev = planet_small()
    dm = _dm(backend, mesh, xyz)
    with pytest.raises(ValueError):
        tp.smoothElevation(dm, elev.astype(np.float64), (elev <= 0).astype(np.uint8), 1, 0.2)
    with pytest.raises(ValueError):
        tp.smoothElevation(dm, elev[:-1].copy(), (elev <= 0).astype(np.uint8), 1, 0.2)
    bad = mesh.adjList.copy(); bad[3] = mesh.numRegions + 5

    class M: pass
    m = M(); m.numRegions = mesh.numRegions; m.adjOffset = mesh.adjOffset; m.adjList = bad
    with pytest.raises(PlanetB200Error):
        DeviceMesh(m, xyz, lib=backend)


@pytest.mark.parametrize("flow", ["doubling", "ordered"])
def test_flow_accumulation_forms_match(backend, oracle, planet_medium, flow):
    """Hydraulic flow accumulation (js/terrain-post.js:604-611): subtree sizes by pointer doubling and the ordered dataflow give
    the oracle's drainage targets, flow and elevations bit for bit."""
    from planet_heightmap_generation_b200.terrain_post import erodeComposite
    mesh, xyz, nd, elev = planet_medium()
    ocean = (elev <= 0).astype(np.uint8)
    want = elev.copy()
    o_t, o_f, o_l = oracle.erode_composite(mesh, want, xyz, ocean, 6, 0.0003, 0.5, 1.0, 1, 1.16, 0.015, 2, 0.5, nd, capture_iter=4)
    dm = _dm(backend, mesh, xyz)
    dm.set_option("flow", flow)
    got = elev.copy()
    taps = erodeComposite(dm, got, xyz, ocean, 6, 0.0003, 0.5, 1.0, 1, 1.16, 0.015, 2, 0.5, nd, capture_iter=4)
    land = ocean == 0
    assert_bit_equal(taps[0], o_t, "drainTarget")
    assert_bit_equal(taps[1][land], o_f[land], "flow")
    assert_bit_equal(got, want, "erodeComposite")
