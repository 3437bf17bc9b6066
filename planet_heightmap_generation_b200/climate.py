"""Host-side mirror of the reference's climate stage functions (js/wind.js, js/ocean.js,
js/precipitation.js, js/temperature.js, js/koppen.js) — same names, argument order and result keys.

The result objects stay resident in HBM inside a `ClimateState` (the worker's W.cachedWind /
W.cachedOcean, js/planet-worker.js:291); a `Result` is a read-only mapping that fetches a field from
the device the first time it is indexed, as a numpy array (host mode) or a torch CUDA tensor (when the
stage was called with CUDA tensors).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import DeviceMesh

_NP_KIND = {0: np.float32, 1: np.int32, 2: np.uint8}

WIND_KEYS = tuple(
    [f"r_pressure_{s}" for s in ("summer", "winter")] + [f"r_wind_east_{s}" for s in ("summer", "winter")]
    + [f"r_wind_north_{s}" for s in ("summer", "winter")] + [f"r_wind_speed_{s}" for s in ("summer", "winter")]
    + ["itczLons", "itczLatsSummer", "itczLatsWinter", "r_lat", "r_lon", "r_sinLat", "r_isLand", "r_continentality",
       "r_coastDistLand", "r_plateContinentality", "r_eastX", "r_eastY", "r_eastZ", "r_northX", "r_northY", "r_northZ"])
OCEAN_KEYS = tuple(f"r_ocean_{k}_{s}" for s in ("summer", "winter") for k in ("current_east", "current_north", "speed", "warmth"))
PRECIP_KEYS = ("r_precip_summer", "r_rainshadow_summer", "r_precip_winter", "r_rainshadow_winter")
TEMP_KEYS = ("r_temperature_summer", "r_temperature_winter")


class ClimateState:
    """Device-resident climate results for one DeviceMesh."""

    def __init__(self, mesh: DeviceMesh):
        self.mesh = mesh
        self._h = C.c_void_p()
        mesh.lib.check(mesh.lib.dll.pb_climate_create(mesh._mesh, C.byref(self._h)))
        self._like = None

    def close(self):
        if getattr(self, "_h", None):
            self.mesh.lib.dll.pb_climate_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def field(self, name: str):
        kind, count = C.c_int32(), C.c_int64()
        lib = self.mesh.lib
        lib.check(lib.dll.pb_climate_field_info(self._h, name.encode(), C.byref(kind), C.byref(count)))
        like = self._like
        n = int(count.value)
        if like is not None and hasattr(like, "data_ptr"):
            import torch
            dt = {0: torch.float32, 1: torch.int32, 2: torch.uint8}[kind.value]
            out = torch.empty(n, dtype=dt, device=like.device)
            self.mesh._begin(out)
            ptr = out.data_ptr()
        else:
            out = np.empty(n, _NP_KIND[kind.value])
            self.mesh._begin(out)
            ptr = out.ctypes.data
        lib.check(lib.dll.pb_climate_get(self._h, name.encode(), ptr))
        return out


class Result:
    def __init__(self, state: ClimateState, keys, timing_key=None):
        self._state, self._keys, self._cache = state, tuple(keys), {}
        if timing_key:
            self._cache[timing_key] = []

    def __getitem__(self, k):
        if k not in self._cache:
            if k not in self._keys:
                raise KeyError(k)
            self._cache[k] = self._state.field(k)
        return self._cache[k]

    def keys(self):
        return self._keys

    def __contains__(self, k):
        return k in self._keys


def _state(mesh: DeviceMesh) -> ClimateState:
    st = getattr(mesh, "_climate", None)
    if st is None:
        st = mesh._climate = ClimateState(mesh)
    return st


def computeWind(mesh: DeviceMesh, r_xyz, r_elevation, plateIsOcean, r_plate, noise, axialTilt=23.5) -> Result:
    """js/wind.js:394-687.  `noise` is the seed of the worker's SimplexNoise instance (a number)."""
    st = _state(mesh)
    n = mesh.numRegions
    mesh._begin(r_elevation, r_plate)
    st._like = r_elevation
    ids = np.ascontiguousarray(sorted(int(p) for p in plateIsOcean), np.int32)
    mesh.lib.check(mesh.lib.dll.pb_compute_wind(
        st._h, mesh._ptr(r_elevation, "f32", n, "r_elevation"), ids.ctypes.data, int(ids.size),
        mesh._ptr(r_plate, "i32", n, "r_plate"), float(noise), float(axialTilt)))
    return Result(st, WIND_KEYS, "_windTiming")


def computeOceanCurrents(mesh: DeviceMesh, r_xyz, r_elevation, windResult) -> Result:
    """js/ocean.js:204-382"""
    st = _state(mesh)
    mesh._begin(r_elevation)
    st._like = r_elevation
    mesh.lib.check(mesh.lib.dll.pb_compute_ocean_currents(st._h, mesh._ptr(r_elevation, "f32", mesh.numRegions, "r_elevation")))
    return Result(st, OCEAN_KEYS, "_oceanTiming")


def computePrecipitation(mesh: DeviceMesh, r_xyz, r_elevation, windResult, oceanResult, precipitationOffset=0, landCoverage=0.3) -> Result:
    """js/precipitation.js:196-684"""
    st = _state(mesh)
    mesh._begin(r_elevation)
    st._like = r_elevation
    mesh.lib.check(mesh.lib.dll.pb_compute_precipitation(
        st._h, mesh._ptr(r_elevation, "f32", mesh.numRegions, "r_elevation"), float(precipitationOffset), float(landCoverage)))
    return Result(st, PRECIP_KEYS, "_precipTiming")


def computeTemperature(mesh: DeviceMesh, r_xyz, r_elevation, windResult, oceanResult, precipResult, temperatureOffset=0) -> Result:
    """js/temperature.js:69-237"""
    st = _state(mesh)
    mesh._begin(r_elevation)
    st._like = r_elevation
    mesh.lib.check(mesh.lib.dll.pb_compute_temperature(
        st._h, mesh._ptr(r_elevation, "f32", mesh.numRegions, "r_elevation"), float(temperatureOffset)))
    return Result(st, TEMP_KEYS, "_tempTiming")


def classifyKoppen(mesh: DeviceMesh, r_elevation, tempResult, precipResult, out=None):
    """js/koppen.js:67-288 → uint8 class ids (index into KOPPEN_CLASSES)."""
    st = _state(mesh)
    n = mesh.numRegions
    if out is None:
        out = mesh._new(r_elevation, "u8", n)
    mesh._begin(r_elevation, out)
    st._like = r_elevation
    mesh.lib.check(mesh.lib.dll.pb_classify_koppen(st._h, mesh._ptr(r_elevation, "f32", n, "r_elevation"),
                                                   mesh._ptr(out, "u8", n, "r_koppen")))
    return out


def computeClimate(mesh: DeviceMesh, r_elevation, plateIsOcean, r_plate, noise, temperatureOffset=0, precipitationOffset=0,
                   landCoverage=0.3, out_koppen=None):
    """The climate half of handleGenerate / handleComputeClimate (js/planet-worker.js:229-268, 579-672):
    all five stages in one C-ABI call.  Returns (windResult, oceanResult, precipResult, tempResult, koppen)."""
    st = _state(mesh)
    n = mesh.numRegions
    if out_koppen is None:
        out_koppen = mesh._new(r_elevation, "u8", n)
    mesh._begin(r_elevation, r_plate, out_koppen)
    st._like = r_elevation
    ids = np.ascontiguousarray(sorted(int(p) for p in plateIsOcean), np.int32)
    mesh.lib.check(mesh.lib.dll.pb_compute_climate(
        st._h, mesh._ptr(r_elevation, "f32", n, "r_elevation"), ids.ctypes.data, int(ids.size),
        mesh._ptr(r_plate, "i32", n, "r_plate"), float(noise), float(temperatureOffset), float(precipitationOffset),
        float(landCoverage), mesh._ptr(out_koppen, "u8", n, "r_koppen")))
    return (Result(st, WIND_KEYS), Result(st, OCEAN_KEYS), Result(st, PRECIP_KEYS), Result(st, TEMP_KEYS), out_koppen)


# js/koppen.js:19-51
KOPPEN_CODES = ("Ocean", "Af", "Am", "Aw", "BWh", "BWk", "BSh", "BSk", "Cfa", "Cfb", "Cfc", "Csa", "Csb", "Csc", "Cwa", "Cwb",
                "Cwc", "Dfa", "Dfb", "Dfc", "Dfd", "Dsa", "Dsb", "Dsc", "Dsd", "Dwa", "Dwb", "Dwc", "Dwd", "ET", "EF")
