"""Parity of the climate stack (wind → ocean → precipitation → temperature → Köppen) against the oracle.
Every published field is compared bit for bit: the kernels load f32, compute in FP64 in the reference's
order and store f32; integer fields (BFS hop counts, masks, Köppen classes) must be identical.
Runs on the host emulation of the kernels in the CPU suite and on the CUDA library with -m gpu."""
import numpy as np
import pytest

from tests.conftest import assert_bit_equal, make_planet

F32_FIELDS = {
    "wind": ["r_lat", "r_lon", "r_sinLat", "r_eastX", "r_eastY", "r_eastZ", "r_northX", "r_northY", "r_northZ",
             "itczLons", "itczLatsSummer", "itczLatsWinter", "r_continentality", "r_plateContinentality",
             "r_pressure_summer", "r_pressure_winter", "r_wind_east_summer", "r_wind_north_summer", "r_wind_speed_summer",
             "r_wind_east_winter", "r_wind_north_winter", "r_wind_speed_winter"],
    "ocean": ["r_ocean_current_east_summer", "r_ocean_current_north_summer", "r_ocean_speed_summer", "r_ocean_warmth_summer",
              "r_ocean_current_east_winter", "r_ocean_current_north_winter", "r_ocean_speed_winter", "r_ocean_warmth_winter"],
    "precip": ["r_precip_summer", "r_rainshadow_summer", "r_precip_winter", "r_rainshadow_winter"],
    "temp": ["r_temperature_summer", "r_temperature_winter"],
}


def _setup(backend, oracle, n_cells, seed=42):
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200.sphere import synthetic_plates
    mesh, xyz, nd, elev = make_planet(oracle, n_cells, seed)
    r_plate, plate_is_ocean = synthetic_plates(xyz, elev, seed)
    return mesh, xyz, elev, r_plate, plate_is_ocean, DeviceMesh(mesh, xyz, lib=backend)


@pytest.mark.parametrize("n_cells", [3000, 20000])
def test_climate_stages(backend, oracle, n_cells):
    from planet_heightmap_generation_b200 import climate as cl
    mesh, xyz, elev, r_plate, pio, dm = _setup(backend, oracle, n_cells)
    oc = oracle.Climate(mesh, xyz)

    oc.wind(elev, pio, r_plate, 42)
    wind = cl.computeWind(dm, xyz, elev, pio, r_plate, 42)
    assert_bit_equal(wind["r_isLand"], oc.get("r_isLand", np.uint8), "r_isLand")
    assert_bit_equal(wind["r_coastDistLand"], oc.get("r_coastDistLand", np.int32), "r_coastDistLand (BFS hop counts)")
    assert_bit_equal(cl._state(dm).field("r_plateDist"), oc.get("r_plateDist", np.int32), "r_plateDist (BFS hop counts)")
    for k in F32_FIELDS["wind"]:
        assert_bit_equal(wind[k], oc.get(k), k)
    assert oc.get("r_coastDistLand", np.int32).max() > 5

    oc.ocean(elev)
    ocean = cl.computeOceanCurrents(dm, xyz, elev, wind)
    st = cl._state(dm)
    for k in ("r_oceanCoastDist", "r_westCoastDist", "r_eastCoastDist"):
        assert_bit_equal(st.field(k), oc.get(k, np.int32), k)
    for k in F32_FIELDS["ocean"]:
        assert_bit_equal(ocean[k], oc.get(k), k)

    oc.precipitation(elev, 0.0, 0.3)
    precip = cl.computePrecipitation(dm, xyz, elev, wind, ocean, 0.0, 0.3)
    for k in ("r_elevGradE", "r_elevGradN", "r_precip_complex_summer", "r_precip_heuristic_summer", "r_westCoast"):
        assert_bit_equal(st.field(k), oc.get(k), k)
    for k in F32_FIELDS["precip"]:
        assert_bit_equal(precip[k], oc.get(k), k)

    oc.temperature(elev, 0.0)
    temp = cl.computeTemperature(dm, xyz, elev, wind, ocean, precip, 0.0)
    for k in F32_FIELDS["temp"]:
        assert_bit_equal(temp[k], oc.get(k), k)

    want = oc.koppen(elev)
    got = cl.classifyKoppen(dm, elev, temp, precip)
    assert_bit_equal(got, want, "r_koppen")
    assert np.unique(want).size >= 8, "test planet must exercise several Köppen classes"


def test_compute_climate_one_call_with_offsets(backend, oracle):
    """handleComputeClimate path with non-default sliders (temperature +5, precipitation -0.4, land 0.55)."""
    from planet_heightmap_generation_b200 import climate as cl
    mesh, xyz, elev, r_plate, pio, dm = _setup(backend, oracle, 8000, seed=7)
    oc = oracle.Climate(mesh, xyz)
    want = oc.run_all(elev, pio, r_plate, 7, temperature_offset=5.0, precipitation_offset=-0.4, land_coverage=0.55)
    wind, ocean, precip, temp, got = cl.computeClimate(dm, elev, pio, r_plate, 7, 5.0, -0.4, 0.55)
    assert_bit_equal(got, want, "r_koppen")
    for k in F32_FIELDS["precip"]:
        assert_bit_equal(precip[k], oc.get(k), k)
    for k in F32_FIELDS["temp"]:
        assert_bit_equal(temp[k], oc.get(k), k)


def test_climate_edge_cases(backend, oracle):
    """All-ocean planet, no oceanic plates, and stage-order errors."""
    from planet_heightmap_generation_b200 import climate as cl
    from planet_heightmap_generation_b200._lib import PlanetB200Error
    mesh, xyz, elev, r_plate, pio, dm = _setup(backend, oracle, 3000)
    with pytest.raises(PlanetB200Error):
        cl.computePrecipitation(dm, xyz, elev, None, None)      # before computeWind
    sea = (-np.abs(elev) - 0.01).astype(np.float32)
    oc = oracle.Climate(mesh, xyz)
    want = oc.run_all(sea, set(), r_plate, 1)
    wind, ocean, precip, temp, got = cl.computeClimate(dm, sea, set(), r_plate, 1)
    assert_bit_equal(got, want, "r_koppen all-ocean")
    assert (got == 0).all()
    for k in ("r_precip_summer", "r_temperature_winter", "r_ocean_warmth_summer"):
        src = {"r_precip_summer": precip, "r_temperature_winter": temp, "r_ocean_warmth_summer": ocean}[k]
        assert_bit_equal(src[k], oc.get(k), k)


def test_gradients_and_masked_smoothing_properties(backend, oracle, planet_small):
    """Size-independent properties: a constant field has zero gradient and is a fixed point of every
    smoothing pass; masked smoothing never changes cells outside the mask."""
    from planet_heightmap_generation_b200.climate_util import smoothField
    from planet_heightmap_generation_b200.engine import DeviceMesh
    mesh, xyz, nd, elev = planet_small()
    dm = DeviceMesh(mesh, xyz, lib=backend)
    c = np.full(mesh.numRegions, 0.375, np.float32)
    got = c.copy()
    smoothField(dm, got, 7)
    assert_bit_equal(got, c, "constant field under smoothField")
