"""Per-region colour ramps: elevationToColor, biomeColor + smoothBiomeColors, heightmap / land heightmap /
land mask colours through the C ABI against the oracle's restatement, bit-exact Float32 buffers."""
import numpy as np
import pytest

from planet_heightmap_generation_b200.engine import DeviceMesh
from tests.conftest import make_planet


@pytest.mark.parametrize("mode", ["terrain", "biome", "biomeRaw", "koppen", "heightmap", "landheightmap", "landmask"])
def test_region_colors_match_oracle(backend, oracle, mode):
    mesh, xyz, nd, elev = make_planet(oracle, 6000)
    rng = np.random.default_rng(11)
    # cover every branch: deep ocean … above the last terrain knee, exact knees, every Köppen class incl. an out-of-table id
    elev = elev.copy()
    elev[:40] = np.asarray([-0.9, -0.5, -0.3, -0.1, -0.05, 0.0, 0.01, 0.03, 0.1, 0.25, 0.4, 0.5, 0.6, 0.75, 0.9, 0.95, 1.0, 1.2, 0.2, 0.7] * 2, np.float32)
    koppen = rng.integers(0, 31, mesh.numRegions).astype(np.uint8)
    koppen[:31] = np.arange(31)
    koppen[31] = 200
    elev[100:131] = 0.97          # snow / alpine zones for every class
    koppen[100:131] = np.arange(31)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    got = dm.regionColors(mode, elev, koppen if mode.startswith("biome") or mode == "koppen" else None)
    want = oracle.region_colors(mesh, mode, elev, koppen)
    assert got.dtype == np.float32 and got.shape == (3 * mesh.numRegions,)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    assert got.min() >= 0 and got.max() <= 1.0
    dm.close()


def test_region_colors_rejects_bad_arguments(backend, oracle):
    from planet_heightmap_generation_b200 import PlanetB200Error
    mesh, xyz, nd, elev = make_planet(oracle, 3000)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    with pytest.raises(PlanetB200Error):
        dm.regionColors("biome", elev, None)
    dm.close()
