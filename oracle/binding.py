"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding for oracle/liboracle.so (CPU restatement of the reference).  May be imported
only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    if force and os.path.exists(so):
        os.unlink(so)
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])   # no-op when up to date
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_cell_noise.restype = C.c_double
        _LIB.orc_cell_noise.argtypes = [C.c_double]
        _LIB.orc_percentile.restype = C.c_double
        _LIB.orc_climate_create.restype = C.c_void_p
        _LIB.orc_climate_get.restype = C.c_int64
        _LIB.orc_elev_create.restype = C.c_void_p
        _LIB.orc_elev_get.restype = C.c_int64
        _LIB.orc_pair_intensity.restype = C.c_double
    return _LIB


def _opt(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def rng(seed, n):
    out = np.empty(n, np.float64)
    lib().orc_rng(C.c_double(seed), C.c_int(n), _p(out, C.c_double))
    return out


def simplex_perm(seed):
    perm = np.empty(512, np.uint8)
    pm12 = np.empty(512, np.uint8)
    lib().orc_simplex_perm(C.c_double(seed), _p(perm, C.c_uint8), _p(pm12, C.c_uint8))
    return perm, pm12


def noise(seed, kind, xyz, octaves=5, persistence=2.0 / 3):
    xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
    out = np.empty(xyz.shape[0], np.float64)
    k = {"noise3D": 0, "fbm": 1, "ridgedFbm": 2}[kind]
    lib().orc_noise(C.c_double(seed), C.c_int(k), C.c_int(xyz.shape[0]), _p(xyz, C.c_double),
                    C.c_int(octaves), C.c_double(persistence), _p(out, C.c_double))
    return out


def cell_noise(r):
    return lib().orc_cell_noise(float(r))


_DET = {"exp": 0, "log": 1, "pow": 2, "atan": 3, "asin": 4, "atan2": 5, "sin": 6, "cos": 7, "tanh": 8}


def detmath(kind, x, y=None):
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros_like(x) if y is None else np.ascontiguousarray(y, np.float64)
    out = np.empty_like(x)
    lib().orc_detmath(C.c_int(_DET[kind]), C.c_int(x.size), _p(x, C.c_double), _p(y, C.c_double),
                      _p(out, C.c_double))
    return out


def fibonacci_sphere(n, jitter, seed):
    """js/sphere-mesh.js:9-37 + pole (0,0,1) appended (js/sphere-mesh.js:179-181)."""
    out = np.zeros(3 * (n + 1), np.float32)
    lib().orc_fibonacci_sphere(C.c_int(n), C.c_double(jitter), C.c_double(seed), _p(out, C.c_float))
    out[3 * n:] = (0, 0, 1)
    return out


def _mesh_args(mesh):
    return (C.c_int(mesh.numRegions), _p(mesh.adjOffset, C.c_int32), _p(mesh.adjList, C.c_int32))


def neighbor_dist(mesh, xyz):
    out = np.empty(mesh.adjList.shape[0], np.float32)
    lib().orc_neighbor_dist(*_mesh_args(mesh), _p(xyz, C.c_float), _p(out, C.c_float))
    return out


def warp_terrain(mesh, elev, xyz, seed, strength, hotspot=None):
    lib().orc_warp_terrain(*_mesh_args(mesh), _p(elev, C.c_float), _p(xyz, C.c_float), C.c_double(seed),
                           C.c_double(strength), _opt(hotspot, C.c_float))


def smooth_elevation(mesh, elev, is_ocean, iterations, strength):
    lib().orc_smooth_elevation(*_mesh_args(mesh), _p(elev, C.c_float), _p(is_ocean, C.c_uint8),
                               C.c_int(iterations), C.c_double(strength))


def priority_flood_carve(mesh, elev, is_ocean, carve_strength):
    n = mesh.numRegions
    drain = np.empty(n, np.int32)
    surf = np.empty(n, np.float32)
    openo = np.empty(n, np.uint8)
    lib().orc_priority_flood_carve(*_mesh_args(mesh), _p(elev, C.c_float), _p(is_ocean, C.c_uint8),
                                   C.c_double(carve_strength), _p(drain, C.c_int32), _p(surf, C.c_float),
                                   _p(openo, C.c_uint8))
    return drain, surf, openo


def erode_composite(mesh, elev, xyz, is_ocean, hIters, K, m, dt, tIters, talus, kThermal, gIters,
                    glacialStrength, neighborDist, capture_iter=-1):
    n = mesh.numRegions
    dt_ = np.full(n, -2, np.int32)
    fl = np.zeros(n, np.float32)
    lo = np.full(n, -1, np.int32)
    lib().orc_erode_composite(*_mesh_args(mesh), _p(elev, C.c_float), _p(xyz, C.c_float),
                              _p(is_ocean, C.c_uint8), C.c_int(hIters), C.c_double(K), C.c_double(m),
                              C.c_double(dt), C.c_int(tIters), C.c_double(talus), C.c_double(kThermal),
                              C.c_int(gIters), C.c_double(glacialStrength), _p(neighborDist, C.c_float),
                              C.c_int(capture_iter), _p(dt_, C.c_int32), _p(fl, C.c_float), _p(lo, C.c_int32))
    return dt_, fl, lo


def sharpen_ridges(mesh, elev, is_ocean, iterations, strength):
    lib().orc_sharpen_ridges(*_mesh_args(mesh), _p(elev, C.c_float), _p(is_ocean, C.c_uint8),
                             C.c_int(iterations), C.c_double(strength))


def apply_soil_creep(mesh, elev, is_ocean, iterations, strength):
    lib().orc_apply_soil_creep(*_mesh_args(mesh), _p(elev, C.c_float), _p(is_ocean, C.c_uint8),
                               C.c_int(iterations), C.c_double(strength))


def run_post_processing(mesh, xyz, elev, params, neighborDist, seed, hotspot=None, h_iters_override=-1):
    n = mesh.numRegions
    delta = np.empty(n, np.float32)
    is_ocean = np.empty(n, np.uint8)
    lib().orc_run_post_processing(
        *_mesh_args(mesh), _p(xyz, C.c_float), _p(elev, C.c_float),
        C.c_double(params["smoothing"]), C.c_double(params["glacialErosion"]),
        C.c_double(params["hydraulicErosion"]), C.c_double(params["thermalErosion"]),
        C.c_double(params["ridgeSharpening"]), C.c_double(params["terrainWarp"]),
        C.c_int(h_iters_override), _p(neighborDist, C.c_float), C.c_double(seed),
        _opt(hotspot, C.c_float), _p(delta, C.c_float), _p(is_ocean, C.c_uint8))
    return delta, is_ocean


def smooth_field(mesh, field, passes):
    lib().orc_smooth_field(*_mesh_args(mesh), _p(field, C.c_float), C.c_int(passes))


def percentile(arr, p):
    arr = np.ascontiguousarray(arr, np.float32)
    return lib().orc_percentile(_p(arr, C.c_float), C.c_int(arr.size), C.c_double(p))


# ---- climate stack (oracle/climate.cpp) ---------------------------------------------------------------
_KIND = {np.dtype(np.float32): 0, np.dtype(np.int32): 1, np.dtype(np.uint8): 2}


class Climate:
    """computeWind → computeOceanCurrents → computePrecipitation → computeTemperature → classifyKoppen
    (js/planet-worker.js:229-268) with every result field kept under the reference's key name."""

    def __init__(self, mesh, xyz):
        self.mesh = mesh
        self.xyz = np.ascontiguousarray(xyz, np.float32)
        self._off = np.ascontiguousarray(mesh.adjOffset, np.int32)
        self._adj = np.ascontiguousarray(mesh.adjList, np.int32)
        self._h = C.c_void_p(lib().orc_climate_create(C.c_int(mesh.numRegions), _p(self._off, C.c_int32),
                                                      _p(self._adj, C.c_int32), _p(self.xyz, C.c_float)))

    def __del__(self):
        try:
            lib().orc_climate_destroy(self._h)
        except Exception:
            pass

    def wind(self, elev, plate_is_ocean, r_plate, noise_seed, axial_tilt=23.5):
        ids = np.ascontiguousarray(sorted(plate_is_ocean), np.int32)
        r_plate = np.ascontiguousarray(r_plate, np.int32)
        lib().orc_climate_wind(self._h, _p(elev, C.c_float), _p(ids, C.c_int32), C.c_int(ids.size),
                               _p(r_plate, C.c_int32), C.c_double(noise_seed), C.c_double(axial_tilt))

    def ocean(self, elev):
        lib().orc_climate_ocean(self._h, _p(elev, C.c_float))

    def precipitation(self, elev, precipitation_offset=0.0, land_coverage=0.3):
        lib().orc_climate_precip(self._h, _p(elev, C.c_float), C.c_double(precipitation_offset),
                                 C.c_double(land_coverage))

    def temperature(self, elev, temperature_offset=0.0):
        lib().orc_climate_temperature(self._h, _p(elev, C.c_float), C.c_double(temperature_offset))

    def koppen(self, elev):
        lib().orc_climate_koppen(self._h, _p(elev, C.c_float))
        return self.get("r_koppen", np.uint8)

    def run_all(self, elev, plate_is_ocean, r_plate, noise_seed, temperature_offset=0.0, precipitation_offset=0.0,
                land_coverage=0.3):
        self.wind(elev, plate_is_ocean, r_plate, noise_seed)
        self.ocean(elev)
        self.precipitation(elev, precipitation_offset, land_coverage)
        self.temperature(elev, temperature_offset)
        return self.koppen(elev)

    def get(self, name, dtype=np.float32):
        kind = _KIND[np.dtype(dtype)]
        n = lib().orc_climate_get(self._h, name.encode(), C.c_int(kind), None, C.c_int64(0))
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype)
        lib().orc_climate_get(self._h, name.encode(), C.c_int(kind), out.ctypes.data_as(C.c_void_p), C.c_int64(n))
        return out


def compute_gradients(mesh, xyz, field, frames6):
    n = mesh.numRegions
    frames6 = np.ascontiguousarray(frames6, np.float32)
    ge, gn = np.empty(n, np.float32), np.empty(n, np.float32)
    lib().orc_compute_gradients(*_mesh_args(mesh), _p(xyz, C.c_float), _p(field, C.c_float), _p(frames6, C.c_float),
                                _p(ge, C.c_float), _p(gn, C.c_float))
    return ge, gn


def smooth_masked(mesh, field, mask, passes):
    lib().orc_smooth_masked(*_mesh_args(mesh), _p(field, C.c_float), _p(mask, C.c_uint8), C.c_int(passes))


# ---- elevation (oracle/elevation.cpp) --------------------------------------------------------------------
def plate_table_arrays(table):
    """table: dict pid -> dict(isOcean=bool, pole=(x,y,z), omega=float, density=float), insertion-ordered."""
    ids = np.ascontiguousarray(list(table.keys()), np.int32)
    oc = np.ascontiguousarray([1 if table[k]["isOcean"] else 0 for k in table], np.uint8)
    pole = np.ascontiguousarray([table[k]["pole"] for k in table], np.float64).reshape(-1)
    omega = np.ascontiguousarray([table[k]["omega"] for k in table], np.float64)
    dens = np.ascontiguousarray([table[k]["density"] for k in table], np.float64)
    return ids, oc, pole, omega, dens


class Elevation:
    """assignElevation (js/elevation.js:216-1391); every intermediate kept under a readable name."""

    def __init__(self, mesh, xyz):
        self.mesh = mesh
        self.xyz = np.ascontiguousarray(xyz, np.float32)
        self._off = np.ascontiguousarray(mesh.adjOffset, np.int32)
        self._adj = np.ascontiguousarray(mesh.adjList, np.int32)
        self._h = C.c_void_p(lib().orc_elev_create(C.c_int(mesh.numRegions), _p(self._off, C.c_int32),
                                                   _p(self._adj, C.c_int32), _p(self.xyz, C.c_float)))

    def __del__(self):
        try:
            lib().orc_elev_destroy(self._h)
        except Exception:
            pass

    def assign(self, r_plate, plates, plate_seeds, noise_seed, noise_mag, seed, spread, r_super=None, super_plates=None):
        r_plate = np.ascontiguousarray(r_plate, np.int32)
        ids, oc, pole, om, de = plate_table_arrays(plates)
        seeds = np.ascontiguousarray(list(plate_seeds), np.int32)
        if super_plates:
            r_super = np.ascontiguousarray(r_super, np.int32)
            sids, soc, spole, som, sde = plate_table_arrays(super_plates)
            lib().orc_elev_assign(self._h, _p(r_plate, C.c_int32), C.c_int(ids.size), _p(ids, C.c_int32), _p(oc, C.c_uint8),
                                  _p(pole, C.c_double), _p(om, C.c_double), _p(de, C.c_double), _p(seeds, C.c_int32),
                                  C.c_int(seeds.size), C.c_double(noise_seed), C.c_double(noise_mag), C.c_double(seed),
                                  C.c_double(spread), _p(r_super, C.c_int32), C.c_int(sids.size), _p(sids, C.c_int32),
                                  _p(soc, C.c_uint8), _p(spole, C.c_double), _p(som, C.c_double), _p(sde, C.c_double))
        else:
            lib().orc_elev_assign(self._h, _p(r_plate, C.c_int32), C.c_int(ids.size), _p(ids, C.c_int32), _p(oc, C.c_uint8),
                                  _p(pole, C.c_double), _p(om, C.c_double), _p(de, C.c_double), _p(seeds, C.c_int32),
                                  C.c_int(seeds.size), C.c_double(noise_seed), C.c_double(noise_mag), C.c_double(seed),
                                  C.c_double(spread), None, C.c_int(0), None, None, None, None, None)

    def get(self, name, dtype=np.float32):
        kind = _KIND[np.dtype(dtype)]
        n = lib().orc_elev_get(self._h, name.encode(), C.c_int(kind), None, C.c_int64(0))
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype)
        lib().orc_elev_get(self._h, name.encode(), C.c_int(kind), out.ctypes.data_as(C.c_void_p), C.c_int64(n))
        return out


def pair_intensity(a, b):
    return lib().orc_pair_intensity(C.c_int(a), C.c_int(b))


# ---- plate pipeline on the hi-res mesh (oracle/plates.cpp) --------------------------------------------------------
def project_coarse_plates(mesh, xyz, coarse_mesh, coarse_xyz, coarse_r_plate, seed, num_plates=None):
    """projectCoarsePlates (js/coarse-plates.js:51-117) → r_plate int32[N]."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    cxyz = np.ascontiguousarray(coarse_xyz, np.float32)
    coff = np.ascontiguousarray(coarse_mesh.adjOffset, np.int32)
    cadj = np.ascontiguousarray(coarse_mesh.adjList, np.int32)
    crp = np.ascontiguousarray(coarse_r_plate, np.int32)
    out = np.empty(mesh.numRegions, np.int32)
    lib().orc_project_coarse_plates(C.c_int(mesh.numRegions), _p(xyz, C.c_float), C.c_int(coarse_mesh.numRegions),
                                    _p(coff, C.c_int32), _p(cadj, C.c_int32), _p(cxyz, C.c_float), _p(crp, C.c_int32),
                                    C.c_double(seed), C.c_int(-1 if num_plates is None else int(num_plates)), _p(out, C.c_int32))
    return out


def smooth_and_reconnect_plates(mesh, r_plate, plate_seeds, num_passes):
    """smoothAndReconnectPlates (js/plates.js:241-348); r_plate mutated in place."""
    seeds = np.ascontiguousarray(list(plate_seeds), np.int32)
    lib().orc_smooth_and_reconnect_plates(*_mesh_args(mesh), _p(r_plate, C.c_int32), _p(seeds, C.c_int32),
                                          C.c_int(seeds.size), C.c_int(num_passes))
    return r_plate


def build_super_plates(mesh, r_plate, plates):
    """buildSuperPlates (js/super-plates.js:16-273).  plates: insertion-ordered dict pid → dict(isOcean, pole|None, omega,
    density|None).  Returns (r_superPlate, dict sp → dict(isOcean, pole, omega, density))."""
    r_plate = np.ascontiguousarray(r_plate, np.int32)
    ids = np.ascontiguousarray(list(plates.keys()), np.int32)
    has = np.ascontiguousarray([0 if plates[k].get("pole") is None else 1 for k in plates], np.uint8)
    pole = np.ascontiguousarray([plates[k]["pole"] if plates[k].get("pole") is not None else (0, 0, 0) for k in plates], np.float64).reshape(-1)
    om = np.ascontiguousarray([plates[k].get("omega", 0.0) for k in plates], np.float64)
    oc = np.ascontiguousarray([1 if plates[k]["isOcean"] else 0 for k in plates], np.uint8)
    de = np.ascontiguousarray([np.nan if plates[k].get("density") is None else plates[k]["density"] for k in plates], np.float64)
    cap = max(ids.size, 2)
    r_super = np.empty(mesh.numRegions, np.int32)
    sp_pole, sp_om = np.zeros(3 * cap), np.zeros(cap)
    sp_oc, sp_de = np.zeros(cap, np.uint8), np.zeros(cap)
    n = lib().orc_build_super_plates(*_mesh_args(mesh), _p(r_plate, C.c_int32), C.c_int(ids.size), _p(ids, C.c_int32),
                                     _p(has, C.c_uint8), _p(pole, C.c_double), _p(om, C.c_double), _p(oc, C.c_uint8),
                                     _p(de, C.c_double), _p(r_super, C.c_int32), _p(sp_pole, C.c_double), _p(sp_om, C.c_double),
                                     _p(sp_oc, C.c_uint8), _p(sp_de, C.c_double), C.c_int(cap))
    if n < 0:
        raise RuntimeError("super plate capacity")
    table = {sp: dict(isOcean=bool(sp_oc[sp]), pole=tuple(sp_pole[3 * sp:3 * sp + 3]), omega=float(sp_om[sp]), density=float(sp_de[sp]))
             for sp in range(n)}
    return r_super, table


def build_sphere(n, jitter, seed, mesh_order="canonical"):
    """buildSphere(N, jitter, makeRng(seed)) (js/sphere-mesh.js:174-186) → (SphereMesh, r_xyz).  mesh_order "canonical": qhull
    triangulation under the documented canonical numbering (mesh_hull.py); "delaunator": the numbering of the reference's own
    triangulator as restated in delaunator_ref.py (pure Python: seconds per 10 000 points)."""
    xyz = fibonacci_sphere(n, jitter, seed)
    if mesh_order == "canonical":
        from .mesh_hull import build_sphere_from_points
        return build_sphere_from_points(xyz)
    from planet_heightmap_generation_b200.mesh import SphereMesh
    from .delaunator_ref import build_sphere_delaunator
    tri, half, _, _, _ = build_sphere_delaunator(xyz)
    return SphereMesh(tri, half, n + 1), xyz


def generate_coarse_plates(seed, num_plates, num_continents, continent_size_variety=0.0, land_coverage=0.3, n_coarse=20000,
                           mesh_order="canonical"):
    """generateCoarsePlates (js/coarse-plates.js:19-39): coarse mesh of buildSphere(n_coarse, 0.75, makeRng(seed + 137)),
    generatePlates (js/plates.js:6-232) and assignOceanLand (js/ocean-land.js:7-238) on it.
    Returns dict(coarseMesh, coarse_xyz, coarse_r_plate, coarsePlateSeeds, coarsePlateVec, coarsePlateIsOcean)."""
    cmesh, cxyz = build_sphere(n_coarse, 0.75, seed + 137, mesh_order)
    n = cmesh.numRegions
    r_plate = np.empty(n, np.int32)
    seeds = np.zeros(num_plates, np.int32)
    pole, omega, oc = np.zeros(3 * num_plates), np.zeros(num_plates), np.zeros(num_plates, np.uint8)
    lib().orc_generate_coarse_plates.restype = C.c_int
    k = lib().orc_generate_coarse_plates(*_mesh_args(cmesh), _p(cxyz, C.c_float), C.c_double(seed), C.c_int(num_plates),
                                         C.c_int(num_continents), C.c_double(continent_size_variety), C.c_double(land_coverage),
                                         _p(r_plate, C.c_int32), _p(seeds, C.c_int32), _p(pole, C.c_double), _p(omega, C.c_double),
                                         _p(oc, C.c_uint8))
    ids = [int(s) for s in seeds[:k]]
    return dict(coarseMesh=cmesh, coarse_xyz=cxyz, coarse_r_plate=r_plate, coarsePlateSeeds=ids,
                coarsePlateVec={s: {"pole": [float(v) for v in pole[3 * i:3 * i + 3]], "omega": float(omega[i])} for i, s in enumerate(ids)},
                coarsePlateIsOcean={s for i, s in enumerate(ids) if oc[i]})


# ---- importHeightmap pieces (js/planet-worker.js:682-831) -------------------------------------------------------------
def sample_heightmap(mesh, xyz, grayscale, width, height):
    px = np.ascontiguousarray(grayscale, np.uint8)
    out = np.empty(mesh.numRegions, np.float32)
    lib().orc_sample_heightmap(C.c_int(mesh.numRegions), _p(np.ascontiguousarray(xyz, np.float32), C.c_float), _p(px, C.c_uint8),
                               C.c_int(width), C.c_int(height), _p(out, C.c_float))
    return out


def derive_synthetic_plates(mesh, elev):
    """→ (r_plate, plateSeeds in Set order, plateIsOcean set)"""
    out = np.empty(mesh.numRegions, np.int32)
    lib().orc_derive_synthetic_plates(*_mesh_args(mesh), _p(np.ascontiguousarray(elev, np.float32), C.c_float), _p(out, C.c_int32))
    seeds = [int(r) for r in np.nonzero(out == np.arange(mesh.numRegions))[0]]
    return out, seeds, {s for s in seeds if elev[s] <= 0}


def classify_imported(mesh, elev):
    m, c, o = (np.empty(mesh.numRegions, np.uint8) for _ in range(3))
    lib().orc_classify_imported(*_mesh_args(mesh), _p(np.ascontiguousarray(elev, np.float32), C.c_float), _p(m, C.c_uint8), _p(c, C.c_uint8),
                                _p(o, C.c_uint8))
    return m, c, o


# ---- colour ramps (js/color-map.js, js/planet-mesh.js:30-80) --------------------------------------------------------------
COLOR_MODES = {"terrain": 0, "biome": 1, "heightmap": 2, "landheightmap": 3, "landmask": 4, "biomeRaw": 5, "koppen": 6}


def region_colors(mesh, mode, elev, koppen=None):
    out = np.empty(3 * mesh.numRegions, np.float32)
    k = np.zeros(mesh.numRegions, np.uint8) if koppen is None else np.ascontiguousarray(koppen, np.uint8)
    lib().orc_region_colors(*_mesh_args(mesh), C.c_int(COLOR_MODES[mode]), _p(np.ascontiguousarray(elev, np.float32), C.c_float),
                            _p(k, C.c_uint8), _p(out, C.c_float))
    return out


# ---- equirectangular map export (js/planet-mesh.js:1752-1950) ---------------------------------------------------------------
EXPORT_TYPES = {"colormap": 0, "biome": 1, "heightmap": 2, "landheightmap": 3, "landmask": 4, "koppen": 6}


def triangle_centers(mesh, xyz):
    out = np.empty(3 * mesh.numTriangles, np.float32)
    lib().orc_triangle_centers(C.c_int(mesh.numTriangles), _p(mesh.triangles, C.c_int32), _p(np.ascontiguousarray(xyz, np.float32), C.c_float),
                               _p(out, C.c_float))
    return out


def export_map(mesh, xyz, export_type, width, elev, koppen=None):
    """(rgba uint8[height, width, 4], pixelSide int32[height, width]) — exportMap up to the ImageData; `mesh` needs triangles / halfedges."""
    h = width // 2
    rgba = np.empty((h, width, 4), np.uint8)
    side = np.empty((h, width), np.int32)
    k = np.zeros(mesh.numRegions, np.uint8) if koppen is None else np.ascontiguousarray(koppen, np.uint8)
    lib().orc_export_map(*_mesh_args(mesh), C.c_int(mesh.numSides), _p(mesh.triangles, C.c_int32), _p(mesh.halfedges, C.c_int32),
                         _p(np.ascontiguousarray(xyz, np.float32), C.c_float), C.c_int(EXPORT_TYPES[export_type]), C.c_int(width),
                         _p(np.ascontiguousarray(elev, np.float32), C.c_float), _p(k, C.c_uint8), _p(rgba, C.c_uint8), _p(side, C.c_int32))
    return rgba, side


def export_map_triangles(mesh, xyz, export_type, elev, koppen=None):
    """(posArr float32[9 * triCount], colArr float32[9 * triCount]) of exportMap's triangle loop (js/planet-mesh.js:1766-1846)"""
    pos = np.zeros(18 * mesh.numSides, np.float32)
    col = np.zeros(18 * mesh.numSides, np.float32)
    side = np.zeros(2 * mesh.numSides, np.int32)
    k = np.zeros(mesh.numRegions, np.uint8) if koppen is None else np.ascontiguousarray(koppen, np.uint8)
    lib().orc_export_map_triangles.restype = C.c_int
    n = lib().orc_export_map_triangles(*_mesh_args(mesh), C.c_int(mesh.numSides), _p(mesh.triangles, C.c_int32), _p(mesh.halfedges, C.c_int32),
                                       _p(np.ascontiguousarray(xyz, np.float32), C.c_float), C.c_int(EXPORT_TYPES[export_type]),
                                       _p(np.ascontiguousarray(elev, np.float32), C.c_float), _p(k, C.c_uint8), _p(pos, C.c_float), _p(col, C.c_float),
                                       _p(side, C.c_int32))
    return pos[:9 * n].copy(), col[:9 * n].copy()
