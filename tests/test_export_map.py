"""exportMap (js/planet-mesh.js:1752-1950) through the C ABI against the oracle's restatement: the side that
owns every pixel (integer field) and the RGBA bytes of the ImageData, bit for bit, for every export type; then the PNG container."""
import numpy as np
import pytest

from planet_heightmap_generation_b200 import PlanetB200Error
from planet_heightmap_generation_b200 import planet_mesh as pm
from planet_heightmap_generation_b200.engine import DeviceMesh
from tests.conftest import make_planet


def _koppen(mesh, elev, seed=5):
    k = np.random.default_rng(seed).integers(1, 31, mesh.numRegions).astype(np.uint8)
    k[elev <= 0] = 0
    k[:3] = (200, 31, 30)          # ids outside the table take class 0's colour
    return k


@pytest.mark.parametrize("etype", ["colormap", "biome", "koppen", "heightmap", "landheightmap", "landmask"])
def test_export_map_matches_oracle(backend, oracle, etype):
    mesh, xyz, nd, elev = make_planet(oracle, 3000)
    koppen = _koppen(mesh, elev)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    got, got_side = pm.exportMapPixels(dm, etype, 512, elev, koppen, want_sides=True)
    want, want_side = oracle.export_map(mesh, xyz, etype, 512, elev, koppen)
    assert got.shape == (256, 512, 4) and got.dtype == np.uint8
    assert (got_side == want_side).all()
    assert (got == want).all()
    assert (got[..., 3] == 255).all()
    dm.close()


@pytest.mark.parametrize("cells,width", [(200, 1024), (3000, 64), (20000, 1024)])
def test_export_map_triangle_sizes(backend, oracle, cells, width):
    """Triangles of hundreds of pixels (few cells, wide image), sub-pixel triangles (many cells, tiny image) and the usual ≈ 1:1."""
    mesh, xyz, nd, elev = make_planet(oracle, cells)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    got, got_side = pm.exportMapPixels(dm, "colormap", width, elev, want_sides=True)
    want, want_side = oracle.export_map(mesh, xyz, "colormap", width, elev)
    assert (got_side == want_side).all()
    assert (got == want).all()
    dm.close()


def test_export_map_properties(backend, oracle):
    """Properties that do not need the oracle: the triangles of the sides tile the map (background only shows where the
    reference's pole triangles leave a gap); the side that owns a pixel begins at the region nearest to the pixel's direction
    or at one of its neighbours (the map cells are the centroid duals of the triangulation, not its Voronoi cells — with
    jitter 0.75 about a fifth of the area lies between the two); the land mask is black / white only and has the land area."""
    mesh, xyz, nd, elev = make_planet(oracle, 3000)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    W = 1024
    px, side = pm.exportMapPixels(dm, "landmask", W, elev, want_sides=True)
    H = W // 2
    assert (side >= 0).mean() > 0.97
    assert set(np.unique(px[..., :3])) <= {0, 255}
    # pixel centre → unit vector in the renderer's y-up frame (lon = atan2(x, z), lat = asin(y))
    jj, ii = np.mgrid[0:H, 0:W]
    lon = (-2 + 4 * (ii + 0.5) / W) * np.pi / 2
    lat = (1 - 2 * (jj + 0.5) / H) * np.pi / 2
    d = np.stack([np.cos(lat) * np.sin(lon), np.sin(lat), np.cos(lat) * np.cos(lon)], -1).reshape(-1, 3)
    band = np.abs(lat.reshape(-1)) < np.radians(75)          # lon/lat triangles bend away from geodesics near the poles
    own = mesh.triangles[np.maximum(side.reshape(-1), 0)]
    p = xyz.reshape(-1, 3).astype(np.float64)
    from scipy.spatial import cKDTree
    nearest = cKDTree(p).query(d)[1]
    covered = side.reshape(-1) >= 0
    assert (own == nearest)[band & covered].mean() > 0.75
    is_nb = np.zeros(own.shape, bool)
    for k in range(int(np.diff(mesh.adjOffset).max())):
        slot = mesh.adjOffset[nearest] + k
        valid = slot < mesh.adjOffset[nearest + 1]
        is_nb |= valid & (mesh.adjList[np.minimum(slot, mesh.adjList.size - 1)] == own)
    assert ((own == nearest) | is_nb)[band & covered].mean() > 0.999
    # land pixels ≈ land area (equirectangular pixels weighted by cos(lat))
    w = np.cos(lat)
    land_px = (w * (px[..., 0] == 255)).sum() / w.sum()
    land_cells = (elev > 0).mean()
    assert abs(land_px - land_cells) < 0.02
    dm.close()


def test_export_map_fallback_and_errors(backend, oracle):
    mesh, xyz, nd, elev = make_planet(oracle, 3000)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    a = pm.exportMapPixels(dm, "biome", 128, elev, None)          # no climate yet: the colour map (:1764-1765, 1788-1793)
    b = pm.exportMapPixels(dm, "colormap", 128, elev)
    assert (a == b).all()
    for bad in (0, 3, 131072):
        with pytest.raises(PlanetB200Error):
            pm.exportMapPixels(dm, "colormap", bad, elev)
    dm.close()


def test_png_round_trip_and_filenames(backend, oracle):
    mesh, xyz, nd, elev = make_planet(oracle, 3000)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    name, png = pm.exportMap(dm, "heightmap", 256, elev, seed="abc123")
    assert name == "orogen-heightmap-abc123.png"
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    assert (pm.decode_png(png) == pm.exportMapPixels(dm, "heightmap", 256, elev)).all()
    assert pm.exportFilename("biome", 7) == "orogen-satellite-7.png"
    assert pm.exportFilename("koppen", 7) == "orogen-climate-7.png"
    assert pm.exportFilename("landmask", 7) == "orogen-landmask-7.png"
    assert pm.exportFilename("landheightmap", 7) == "orogen-land-heightmap-7.png"
    assert pm.exportFilename("anything-else", 7) == "orogen-colormap-7.png"
    dm.close()
