"""Parity at sizes where the reference changes behaviour: above 200 000 regions `assignElevation` and `findCollisions`
switch to 2 noise octaves (js/elevation.js:55, 457), and BASELINE config 2 asks for 50 stream-power iterations, which
puts the second priority flood at iteration 38 (js/terrain-post.js:446).  The whole chain runs on one 250 001-cell planet:
assignElevation (dual layer) → runPostProcessing(hIters = 50) → climate stack, every array compared bit for bit with the
oracle.  GPU only (the CPU suite covers the same code at ≤ 20 000 cells through the emulation)."""
import numpy as np
import pytest

from tests.conftest import assert_bit_equal, make_planet
from tests.test_climate_parity import F32_FIELDS
from tests.test_elevation_parity import _call

N_LARGE = 250_000
SLIDERS = dict(smoothing=0.10, glacialErosion=0.50, hydraulicErosion=0.50, thermalErosion=0.10,
               ridgeSharpening=0.50, terrainWarp=0.75)

_chain = {}


def _oracle_chain(oracle):
    """oracle side of the chain, computed once per session"""
    if _chain:
        return _chain
    from planet_heightmap_generation_b200.sphere import synthetic_plate_tables
    mesh, xyz, nd, elev0 = make_planet(oracle, N_LARGE)
    r_plate, plates, seeds, r_super, sp = synthetic_plate_tables(xyz, elev0, 42)
    oe = oracle.Elevation(mesh, xyz)
    oe.assign(r_plate, plates, seeds, 42, 0.4, 42, 5, r_super, sp)
    pre = oe.get("r_elevation")
    eroded = pre.copy()
    delta, ocean = oracle.run_post_processing(mesh, xyz, eroded, SLIDERS, nd, 42, oe.get("hotspot"), 50)
    _chain.update(mesh=mesh, xyz=xyz, nd=nd, r_plate=r_plate, plates=plates, seeds=seeds, r_super=r_super, sp=sp, oe=oe,
                  pre=pre, eroded=eroded, delta=delta, ocean=ocean)
    return _chain


@pytest.mark.gpu
def test_assign_elevation_two_octave_branch(cuda_lib, oracle):
    from planet_heightmap_generation_b200.elevation import DEBUG_LAYERS
    from planet_heightmap_generation_b200.engine import DeviceMesh
    c = _oracle_chain(oracle)
    assert c["mesh"].numRegions > 200000
    dm = DeviceMesh(c["mesh"], c["xyz"], lib=cuda_lib)
    got = _call(dm, c["xyz"], c["r_plate"], c["plates"], c["seeds"], 42, 0.4, 42, 5, c["r_super"], c["sp"])
    oe = c["oe"]
    for k in ("mountain_r", "coastline_r", "ocean_r"):
        assert_bit_equal(got[k], oe.get(k, np.uint8), k)
    assert_bit_equal(got["r_stress"], oe.get("r_stress"), "r_stress")
    for k in DEBUG_LAYERS:
        assert_bit_equal(got["debugLayers"][k], oe.get(k), "debug layer " + k)
    assert_bit_equal(got["r_elevation"], c["pre"], "r_elevation")


@pytest.mark.gpu
@pytest.mark.parametrize("flood", ["host", "device"])
def test_run_post_processing_50_iterations_mid_flood(cuda_lib, oracle, flood):
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200.terrain_post import runPostProcessing
    c = _oracle_chain(oracle)
    dm = DeviceMesh(c["mesh"], c["xyz"], lib=cuda_lib)
    dm.set_option("flood", flood)
    got = c["pre"].copy()
    res = runPostProcessing(dm, c["xyz"], got, SLIDERS, c["nd"], 42, c["oe"].get("hotspot"), hItersOverride=50)
    assert_bit_equal(res["r_isOcean"], c["ocean"], "r_isOcean")
    assert_bit_equal(got, c["eroded"], "runPostProcessing elevation (hIters 50, mid flood at 38)")
    assert_bit_equal(res["dl_erosionDelta"], c["delta"], "erosionDelta")


@pytest.mark.gpu
def test_climate_stack_large(cuda_lib, oracle):
    from planet_heightmap_generation_b200 import climate as cl
    from planet_heightmap_generation_b200.engine import DeviceMesh
    c = _oracle_chain(oracle)
    pio = {p for p, v in c["plates"].items() if v["isOcean"]}
    elev = c["eroded"]
    oc = oracle.Climate(c["mesh"], c["xyz"])
    o_koppen = oc.run_all(elev, pio, c["r_plate"], 42)
    dm = DeviceMesh(c["mesh"], c["xyz"], lib=cuda_lib)
    koppen = np.empty(c["mesh"].numRegions, np.uint8)
    wind, ocean, precip, temp, _ = cl.computeClimate(dm, elev, pio, c["r_plate"], 42, 0.0, 0.0, 0.3, out_koppen=koppen)
    for res, keys in ((wind, F32_FIELDS["wind"]), (ocean, F32_FIELDS["ocean"]), (precip, F32_FIELDS["precip"]), (temp, F32_FIELDS["temp"])):
        for k in keys:
            assert_bit_equal(res[k], oc.get(k), k)
    assert_bit_equal(wind["r_coastDistLand"], oc.get("r_coastDistLand", np.int32), "r_coastDistLand")
    assert_bit_equal(koppen, o_koppen, "r_koppen")
    assert len(np.unique(koppen)) >= 8


@pytest.mark.gpu
@pytest.mark.parametrize("etype", ["biome", "heightmap"])
def test_export_map_large(cuda_lib, oracle, etype):
    """exportMap (js/planet-mesh.js:1752-1950) of the eroded 250 001-cell planet at 2048 × 1024: ≈ 1.4 pixels per map triangle,
    the regime of the reference's own exports; owner side per pixel and RGBA bytes against the oracle."""
    from planet_heightmap_generation_b200 import planet_mesh as pm
    from planet_heightmap_generation_b200.engine import DeviceMesh
    c = _oracle_chain(oracle)
    elev = c["eroded"]
    koppen = (np.random.default_rng(3).integers(1, 31, elev.size) * (elev > 0)).astype(np.uint8)
    dm = DeviceMesh(c["mesh"], c["xyz"], lib=cuda_lib)
    got, got_side = pm.exportMapPixels(dm, etype, 2048, elev, koppen, want_sides=True)
    want, want_side = oracle.export_map(c["mesh"], c["xyz"], etype, 2048, elev, koppen)
    assert (got_side == want_side).all()
    assert (got == want).all()
    dm.close()
