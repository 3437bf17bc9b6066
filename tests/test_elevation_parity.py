"""Parity of assignElevation (js/elevation.js) against the oracle: collisions, stress propagation, the five
randomized distance fills, the capped BFS payloads, the harmonic-mean / noise synthesis, coastal roughening,
island arcs, hotspots and peak compression.  Every output and debug layer is compared bit for bit."""
import numpy as np
import pytest

from tests.conftest import assert_bit_equal, make_planet


def _inputs(oracle, n_cells, seed=42):
    from planet_heightmap_generation_b200.sphere import synthetic_plate_tables
    mesh, xyz, nd, elev = make_planet(oracle, n_cells, seed)
    r_plate, plates, seeds, r_super, sp = synthetic_plate_tables(xyz, elev, seed)
    return mesh, xyz, r_plate, plates, seeds, r_super, sp


def _call(dm, xyz, r_plate, plates, seeds, noise_seed, nmag, seed, spread, r_super=None, sp=None):
    from planet_heightmap_generation_b200.elevation import assignElevation
    pio = {p for p, v in plates.items() if v["isOcean"]}
    vec = {p: {"pole": v["pole"], "omega": v["omega"]} for p, v in plates.items()}
    dens = {p: v["density"] for p, v in plates.items()}
    spd = None
    if sp is not None:
        spd = {"r_superPlate": r_super, "superPlateIsOcean": {p for p, v in sp.items() if v["isOcean"]},
               "superPlateVec": {p: {"pole": v["pole"], "omega": v["omega"]} for p, v in sp.items()},
               "superPlateDensity": {p: v["density"] for p, v in sp.items()}}
    return assignElevation(dm, xyz, pio, r_plate, vec, seeds, noise_seed, nmag, seed, spread, dens, spd)


@pytest.mark.parametrize("n_cells,dual", [(3000, False), (20000, True), (20000, False)])
def test_assign_elevation(backend, oracle, n_cells, dual):
    from planet_heightmap_generation_b200.elevation import DEBUG_LAYERS
    from planet_heightmap_generation_b200.engine import DeviceMesh
    mesh, xyz, r_plate, plates, seeds, r_super, sp = _inputs(oracle, n_cells)
    oe = oracle.Elevation(mesh, xyz)
    oe.assign(r_plate, plates, seeds, 42, 0.4, 42, 5, r_super if dual else None, sp if dual else None)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    got = _call(dm, xyz, r_plate, plates, seeds, 42, 0.4, 42, 5, r_super if dual else None, sp if dual else None)
    for k in ("mountain_r", "coastline_r", "ocean_r"):
        assert_bit_equal(got[k], oe.get(k, np.uint8), k)
    assert_bit_equal(got["r_stress"], oe.get("r_stress"), "r_stress")
    for k in DEBUG_LAYERS:
        assert_bit_equal(got["debugLayers"][k], oe.get(k), "debug layer " + k)
    assert_bit_equal(got["r_elevation"], oe.get("r_elevation"), "r_elevation")
    E = got["r_elevation"]
    assert np.isfinite(E).all() and 0.1 < (E > 0).mean() < 0.6
    assert oe.get("domes").size >= 5, "hotspot chains must exist"
    assert (oe.get("hotspot") > 0).sum() > 0 and (oe.get("coastal") != 0).sum() > 100


def test_assign_elevation_other_seed_and_params(backend, oracle):
    from planet_heightmap_generation_b200.engine import DeviceMesh
    mesh, xyz, r_plate, plates, seeds, r_super, sp = _inputs(oracle, 8000, seed=7)
    oe = oracle.Elevation(mesh, xyz)
    oe.assign(r_plate, plates, seeds, 1234.5, 0.9, 77, 8, r_super, sp)
    got = _call(DeviceMesh(mesh, xyz, lib=backend), xyz, r_plate, plates, seeds, 1234.5, 0.9, 77, 8, r_super, sp)
    assert_bit_equal(got["r_elevation"], oe.get("r_elevation"), "r_elevation")
    assert_bit_equal(got["r_stress"], oe.get("r_stress"), "r_stress")


def test_pair_intensity_js_semantics(oracle):
    """getPairIntensity (js/elevation.js:44-53): the second multiply exceeds 2^53 and must round in double
    before ToUint32; values derived by evaluating the cited lines with Python's exact integers + float rounding."""
    def ref(a, b):
        lo, hi = min(a, b), max(a, b)
        def to_i32(v):
            v = int(v) & 0xFFFFFFFF
            return v - (1 << 32) if v >= (1 << 31) else v
        h = (to_i32(lo * 16807) ^ to_i32(hi * 48271)) & 0xFFFFFFFF
        x = (to_i32(h) >> 16) ^ to_i32(h)
        p = float(x) * float(0x45d9f3b)          # double product (may round)
        h = int(p) & 0xFFFFFFFF
        return 0.5 + (h % 10001) / 10000
    for a, b in [(3, 17), (19999, 12345), (0, 1), (18000, 18001), (7, 7), (123456, 654321)]:
        assert oracle.pair_intensity(a, b) == ref(a, b)


@pytest.mark.parametrize("kind", ["all_land", "all_ocean", "zero_omega"])
def test_assign_elevation_edge_cases(backend, oracle, kind):
    """No oceanic plate at all (empty ocean seed set → infinite ocean distances), no continental plate, and plates
    that do not move (no collisions → empty mountain set, hotspot chains skipped for driftLen < 1e-6)."""
    from planet_heightmap_generation_b200.engine import DeviceMesh
    mesh, xyz, r_plate, plates, seeds, r_super, sp = _inputs(oracle, 3000)
    plates = {p: dict(v) for p, v in plates.items()}
    for v in plates.values():
        if kind == "all_land":
            v["isOcean"] = False
        elif kind == "all_ocean":
            v["isOcean"] = True
        else:
            v["omega"] = 0.0
    oe = oracle.Elevation(mesh, xyz)
    oe.assign(r_plate, plates, seeds, 42, 0.4, 42, 5)
    got = _call(DeviceMesh(mesh, xyz, lib=backend), xyz, r_plate, plates, seeds, 42, 0.4, 42, 5)
    assert_bit_equal(got["r_elevation"], oe.get("r_elevation"), "r_elevation " + kind)
    assert_bit_equal(got["r_stress"], oe.get("r_stress"), "r_stress " + kind)
    for k in ("mountain_r", "coastline_r", "ocean_r"):
        assert_bit_equal(got[k], oe.get(k, np.uint8), k)
    if kind == "zero_omega":
        assert got["mountain_r"].sum() == 0 and oe.get("domes").size == 0


def test_assign_elevation_rejects_unknown_plate_id(backend, oracle):
    from planet_heightmap_generation_b200._lib import PlanetB200Error
    from planet_heightmap_generation_b200.engine import DeviceMesh
    mesh, xyz, r_plate, plates, seeds, r_super, sp = _inputs(oracle, 3000)
    bad = r_plate.copy(); bad[5] = 10 ** 6
    with pytest.raises(PlanetB200Error):
        _call(DeviceMesh(mesh, xyz, lib=backend), xyz, bad, plates, seeds, 42, 0.4, 42, 5)
