"""Host-side mirror of js/elevation.js: assignElevation."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import ElevationResult, PlateTable
from .engine import DeviceMesh

DEBUG_LAYERS = ("base", "tectonic", "noise", "interior", "coastal", "ocean", "hotspot", "tecActivity", "margins", "backArc",
                "foldRidge", "orogenicPower")


def _table(plate_is_ocean, plate_vec, plate_density, order=None):
    ids = list(order) if order is not None else list(plate_vec.keys())
    keep = dict(
        ids=np.ascontiguousarray(ids, np.int32),
        oc=np.ascontiguousarray([1 if p in plate_is_ocean else 0 for p in ids], np.uint8),
        pole=np.ascontiguousarray([plate_vec[p]["pole"] for p in ids], np.float64).reshape(-1),
        omega=np.ascontiguousarray([plate_vec[p]["omega"] for p in ids], np.float64),
        dens=np.ascontiguousarray([plate_density[p] for p in ids], np.float64))
    t = PlateTable(len(ids), keep["ids"].ctypes.data, keep["oc"].ctypes.data, keep["pole"].ctypes.data,
                   keep["omega"].ctypes.data, keep["dens"].ctypes.data)
    return t, keep


def assignElevation(mesh: DeviceMesh, r_xyz, plateIsOcean, r_plate, plateVec, plateSeeds, noise, noiseMag, seed, spread,
                    plateDensity, superPlateData=None, debug: bool = True):
    """js/elevation.js:216-1391.  plateIsOcean: set of ids; plateVec: {pid: {"pole": (x,y,z), "omega": w}};
    plateSeeds: iterable of ids in the Set's insertion order; noise: the seed of the caller's SimplexNoise;
    plateDensity: {pid: density}; superPlateData: None or {"r_superPlate", "superPlateVec", "superPlateIsOcean",
    "superPlateDensity"}.  Returns the reference's result object with the region Sets as uint8 masks."""
    n = mesh.numRegions
    mesh._begin(r_plate, None if superPlateData is None else superPlateData["r_superPlate"])
    table, keep1 = _table(plateIsOcean, plateVec, plateDensity)
    seeds = np.ascontiguousarray(list(plateSeeds), np.int32)
    sp_ptr, keep2, r_super_ptr = None, None, None
    if superPlateData is not None:
        sp, keep2 = _table(superPlateData["superPlateIsOcean"], superPlateData["superPlateVec"], superPlateData["superPlateDensity"])
        sp_ptr = C.addressof(sp)
        r_super_ptr = mesh._ptr(superPlateData["r_superPlate"], "i32", n, "r_superPlate")
    out = {"r_elevation": mesh._new(r_plate, "f32", n), "r_stress": mesh._new(r_plate, "f32", n),
           "mountain_r": mesh._new(r_plate, "u8", n), "coastline_r": mesh._new(r_plate, "u8", n),
           "ocean_r": mesh._new(r_plate, "u8", n)}
    layers = {k: mesh._new(r_plate, "f32", n) for k in DEBUG_LAYERS} if debug else {}
    res = ElevationResult()
    for k in ("r_elevation", "r_stress", "mountain_r", "coastline_r", "ocean_r"):
        setattr(res, k, mesh._ptr(out[k], "f32" if k.startswith("r_") else "u8", n, k))
    for i, k in enumerate(DEBUG_LAYERS):
        res.debug[i] = mesh._ptr(layers[k], "f32", n, k) if debug else None
    mesh.lib.check(mesh.lib.dll.pb_assign_elevation(
        mesh._mesh, C.addressof(table), mesh._ptr(r_plate, "i32", n, "r_plate"), seeds.ctypes.data, int(seeds.size),
        float(noise), float(noiseMag), float(seed), float(spread), sp_ptr, r_super_ptr, C.addressof(res)))
    out["debugLayers"] = layers
    out["_timing"] = []
    return out
