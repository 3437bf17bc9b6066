"""Host mirror of the plate pipeline on the hi-res mesh: same names, argument order and
result keys as the reference's functions, over the C ABI (csrc/pb_plates.h).

  generateCoarsePlates       js/coarse-plates.js:19-39  (+ generatePlates js/plates.js:6-232, assignOceanLand js/ocean-land.js:7-238)
  projectCoarsePlates        js/coarse-plates.js:51-117
  smoothAndReconnectPlates   js/plates.js:241-348
  buildSuperPlates           js/super-plates.js:16-273

`r_plate` / `r_superPlate` may be numpy arrays (host pointer mode) or torch CUDA tensors (device pointer mode); the
coarse-mesh tables and the plate tables are small host-side objects, as in the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import PlateTable
from .engine import DeviceMesh


class SuperPlateTable(C.Structure):
    _fields_ = [("capacity", C.c_int32), ("numSuperPlates", C.c_int32), ("pole", C.c_void_p), ("omega", C.c_void_p),
                ("isOcean", C.c_void_p), ("density", C.c_void_p)]


def projectCoarsePlates(mesh: DeviceMesh, r_xyz, coarseMesh, coarse_xyz, coarse_r_plate, seed, numPlates=None, out=None):
    """→ r_plate (Int32Array semantics).  `r_xyz` is the mesh's own coordinate array (already resident; accepted for
    signature parity).  `out`: optional numpy / torch.cuda int32 array to fill."""
    n = mesh.numRegions
    if out is None:
        out = np.empty(n, np.int32)
    c_off = np.ascontiguousarray(coarseMesh.adjOffset, np.int32)
    c_adj = np.ascontiguousarray(coarseMesh.adjList, np.int32)
    c_xyz = np.ascontiguousarray(coarse_xyz, np.float32).reshape(-1)
    c_plate = np.ascontiguousarray(coarse_r_plate, np.int32)
    nc = int(coarseMesh.numRegions)
    if c_off.size != nc + 1 or c_xyz.size != 3 * nc or c_plate.size != nc:
        raise ValueError("coarse mesh arrays do not match coarseMesh.numRegions")
    mesh._begin(out)
    mesh.lib.check(mesh.lib.dll.pb_project_coarse_plates(
        mesh._mesh, nc, c_off.ctypes.data, c_adj.ctypes.data, c_xyz.ctypes.data, c_plate.ctypes.data, float(seed),
        -1 if numPlates is None else int(numPlates), mesh._ptr(out, "i32", n, "r_plate")))
    return out


def smoothAndReconnectPlates(mesh: DeviceMesh, r_plate, plateSeeds, numPasses):
    """r_plate is mutated in place (and returned)."""
    seeds = np.ascontiguousarray(list(plateSeeds), np.int32)
    mesh._begin(r_plate)
    mesh.lib.check(mesh.lib.dll.pb_smooth_and_reconnect_plates(
        mesh._mesh, mesh._ptr(r_plate, "i32", mesh.numRegions, "r_plate"), seeds.ctypes.data, int(seeds.size), int(numPasses)))
    return r_plate


def buildSuperPlates(mesh: DeviceMesh, r_plate, plateSeeds, plateVec, plateIsOcean, plateDensity, out=None):
    """→ {r_superPlate, superPlateVec, superPlateIsOcean, superPlateDensity, numSuperPlates}"""
    n = mesh.numRegions
    seeds = [int(s) for s in plateSeeds]
    ids = np.ascontiguousarray(seeds, np.int32)
    oc = np.ascontiguousarray([1 if s in plateIsOcean else 0 for s in seeds], np.uint8)
    pole = np.full(3 * len(seeds), np.nan)
    omega = np.zeros(len(seeds))
    dens = np.full(len(seeds), np.nan)
    for k, s in enumerate(seeds):
        pv = plateVec.get(s)
        if pv is not None and pv.get("pole") is not None:
            pole[3 * k:3 * k + 3] = pv["pole"]
            omega[k] = pv["omega"]
        if plateDensity.get(s) is not None:
            dens[k] = plateDensity[s]
    table = PlateTable(len(seeds), ids.ctypes.data, oc.ctypes.data, pole.ctypes.data, omega.ctypes.data, dens.ctypes.data)
    cap = max(len(seeds), 2)
    sp_pole, sp_omega, sp_oc, sp_dens = np.zeros(3 * cap), np.zeros(cap), np.zeros(cap, np.uint8), np.zeros(cap)
    sp = SuperPlateTable(cap, 0, sp_pole.ctypes.data, sp_omega.ctypes.data, sp_oc.ctypes.data, sp_dens.ctypes.data)
    if out is None:
        out = mesh._new(r_plate, "i32", n)
    mesh._begin(r_plate, out)
    mesh.lib.check(mesh.lib.dll.pb_build_super_plates(mesh._mesh, mesh._ptr(r_plate, "i32", n, "r_plate"), C.addressof(table),
                                                      mesh._ptr(out, "i32", n, "r_superPlate"), C.addressof(sp)))
    k = int(sp.numSuperPlates)
    return {"r_superPlate": out,
            "superPlateVec": {i: {"pole": [float(v) for v in sp_pole[3 * i:3 * i + 3]], "omega": float(sp_omega[i])} for i in range(k)},
            "superPlateIsOcean": {i for i in range(k) if sp_oc[i]},
            "superPlateDensity": {i: float(sp_dens[i]) for i in range(k)},
            "numSuperPlates": k}


class PlateTableOut(C.Structure):
    _fields_ = [("capacity", C.c_int32), ("n", C.c_int32), ("ids", C.c_void_p), ("isOcean", C.c_void_p), ("pole", C.c_void_p),
                ("omega", C.c_void_p), ("density", C.c_void_p)]


N_COARSE = 20000     # js/coarse-plates.js:11


def generateCoarsePlates(mesh: DeviceMesh, seed, numPlates, numContinents, continentSizeVariety=0.0, landCoverage=0.3,
                         numCoarse: int = N_COARSE):
    """→ {coarseMesh, coarse_xyz, coarse_r_plate, coarsePlateSeeds, coarsePlateVec, coarsePlateIsOcean} (+ plateDensity, the
    worker's per-plate draws, js/planet-worker.js:196-201).  `mesh` only supplies the GPU context (the reference's function
    takes none); coarseMesh is a DeviceMesh on the same context."""
    p = int(numPlates)
    n = int(numCoarse) + 1
    ids, oc = np.zeros(p, np.int32), np.zeros(p, np.uint8)
    pole, omega, dens = np.zeros(3 * p), np.zeros(p), np.zeros(p)
    table = PlateTableOut(p, 0, ids.ctypes.data, oc.ctypes.data, pole.ctypes.data, omega.ctypes.data, dens.ctypes.data)
    cxyz = np.empty(3 * n, np.float32)
    crp = np.empty(n, np.int32)
    handle = C.c_void_p()
    mesh.lib.check(mesh.lib.dll.pb_generate_coarse_plates(mesh._ctx, float(seed), p, int(numContinents), float(continentSizeVariety),
                                                          float(landCoverage), int(numCoarse), C.byref(handle), cxyz.ctypes.data,
                                                          crp.ctypes.data, C.addressof(table)))
    k = int(table.n)
    seeds = [int(s) for s in ids[:k]]
    return {"coarseMesh": DeviceMesh._adopt(mesh, handle, cxyz), "coarse_xyz": cxyz, "coarse_r_plate": crp,
            "coarsePlateSeeds": seeds,
            "coarsePlateVec": {s: {"pole": [float(v) for v in pole[3 * i:3 * i + 3]], "omega": float(omega[i])} for i, s in enumerate(seeds)},
            "coarsePlateIsOcean": {s for i, s in enumerate(seeds) if oc[i]},
            "plateDensity": {s: float(dens[i]) for i, s in enumerate(seeds)}}
