"""Host-side mirror of js/terrain-post.js and runPostProcessing (js/planet-worker.js:40-102).

Same function names, argument order and in-place mutation contract as the reference; `mesh` is a
`DeviceMesh`.  `r_xyz` and `neighborDist` arguments are accepted for signature compatibility — the
device mesh already holds both (they are functions of the mesh alone).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .engine import DeviceMesh, PostParams


def warpTerrain(mesh: DeviceMesh, r_elevation, r_xyz, seed, strength, r_hotspot=None):
    """js/terrain-post.js:233-309"""
    if strength <= 0:
        return
    n = mesh.numRegions
    mesh._begin(r_elevation, r_hotspot)
    mesh.lib.check(mesh.lib.dll.pb_warp_terrain(
        mesh._mesh, mesh._ptr(r_elevation, "f32", n, "r_elevation"), float(seed), float(strength),
        mesh._ptr(r_hotspot, "f32", n, "r_hotspot", optional=True)))


def smoothElevation(mesh: DeviceMesh, r_elevation, r_isOcean, iterations, strength):
    """js/terrain-post.js:317-354"""
    n = mesh.numRegions
    mesh._begin(r_elevation, r_isOcean)
    mesh.lib.check(mesh.lib.dll.pb_smooth_elevation(
        mesh._mesh, mesh._ptr(r_elevation, "f32", n, "r_elevation"), mesh._ptr(r_isOcean, "u8", n, "r_isOcean"),
        int(iterations), float(strength)))


def priorityFloodCarve(mesh: DeviceMesh, r_elevation, r_isOcean, carveStrength, taps: bool = False):
    """js/terrain-post.js:59-215.  taps=True also returns (drainTo, surface, isOpenOcean) after pass 1."""
    n = mesh.numRegions
    mesh._begin(r_elevation, r_isOcean)
    dt = sf = oo = None
    if taps:
        dt, sf, oo = mesh._new(r_elevation, "i32", n), mesh._new(r_elevation, "f32", n), mesh._new(r_elevation, "u8", n)
    mesh.lib.check(mesh.lib.dll.pb_priority_flood_carve(
        mesh._mesh, mesh._ptr(r_elevation, "f32", n, "r_elevation"), mesh._ptr(r_isOcean, "u8", n, "r_isOcean"),
        float(carveStrength), mesh._ptr(dt, "i32", n, "drainTo", True), mesh._ptr(sf, "f32", n, "surface", True),
        mesh._ptr(oo, "u8", n, "isOpenOcean", True)))
    return (dt, sf, oo) if taps else None


def erodeComposite(mesh: DeviceMesh, r_elevation, r_xyz, r_isOcean, hIters, K, m, dt, tIters, talusSlope, kThermal,
                   gIters=0, glacialStrength=0.0, neighborDist=None, capture_iter: int = -1):
    """js/terrain-post.js:369-707.  capture_iter >= 0 returns (drainTarget, flow, landOrder) of that
    hydraulic iteration, taken right before the implicit solve."""
    n = mesh.numRegions
    mesh._begin(r_elevation, r_isOcean)
    gIters = int(gIters or 0)
    glacialStrength = float(glacialStrength or 0)
    e = mesh._ptr(r_elevation, "f32", n, "r_elevation")
    o = mesh._ptr(r_isOcean, "u8", n, "r_isOcean")
    if capture_iter < 0:
        mesh.lib.check(mesh.lib.dll.pb_erode_composite(
            mesh._mesh, e, o, int(hIters), float(K), float(m), float(dt), int(tIters), float(talusSlope),
            float(kThermal), gIters, glacialStrength))
        return None
    d_t, fl, lo = mesh._new(r_elevation, "i32", n), mesh._new(r_elevation, "f32", n), mesh._new(r_elevation, "i32", n)
    mesh.lib.check(mesh.lib.dll.pb_erode_composite_debug(
        mesh._mesh, e, o, int(hIters), float(K), float(m), float(dt), int(tIters), float(talusSlope),
        float(kThermal), gIters, glacialStrength, int(capture_iter), mesh._ptr(d_t, "i32", n, "drainTarget"),
        mesh._ptr(fl, "f32", n, "flow"), mesh._ptr(lo, "i32", n, "landOrder")))
    return d_t, fl, lo


def sharpenRidges(mesh: DeviceMesh, r_elevation, r_isOcean, iterations, strength):
    """js/terrain-post.js:713-751"""
    n = mesh.numRegions
    mesh._begin(r_elevation, r_isOcean)
    mesh.lib.check(mesh.lib.dll.pb_sharpen_ridges(
        mesh._mesh, mesh._ptr(r_elevation, "f32", n, "r_elevation"), mesh._ptr(r_isOcean, "u8", n, "r_isOcean"),
        int(iterations), float(strength)))


def applySoilCreep(mesh: DeviceMesh, r_elevation, r_isOcean, iterations, strength):
    """js/terrain-post.js:758-794"""
    n = mesh.numRegions
    mesh._begin(r_elevation, r_isOcean)
    mesh.lib.check(mesh.lib.dll.pb_apply_soil_creep(
        mesh._mesh, mesh._ptr(r_elevation, "f32", n, "r_elevation"), mesh._ptr(r_isOcean, "u8", n, "r_isOcean"),
        int(iterations), float(strength)))


POST_STAGES = ("Terrain warp", "Smoothing", "Erosion composite", "Ridge sharpening", "Soil creep")


def runPostProcessing(mesh: DeviceMesh, r_xyz, r_elevation, params: dict, neighborDist, seed, r_hotspot=None,
                      hItersOverride: int = -1, out_erosionDelta=None, out_isOcean=None, timing: bool = True):
    """js/planet-worker.js:40-102.  Returns {dl_erosionDelta, postTiming, r_isOcean}; r_elevation is
    mutated in place.  `hItersOverride` lets BASELINE configs 2/5 ask for more stream-power iterations
    than the UI slider allows while keeping the slider-derived K."""
    n = mesh.numRegions
    delta = out_erosionDelta if out_erosionDelta is not None else mesh._new(r_elevation, "f32", n)
    ocean = out_isOcean if out_isOcean is not None else mesh._new(r_elevation, "u8", n)
    mesh._begin(r_elevation, r_hotspot, delta, ocean)
    p = PostParams(float(params.get("smoothing", 0)), float(params.get("glacialErosion", 0)),
                   float(params.get("hydraulicErosion", 0)), float(params.get("thermalErosion", 0)),
                   float(params.get("ridgeSharpening", 0)), float(params.get("terrainWarp", 0)), int(hItersOverride))
    mesh.lib.check(mesh.lib.dll.pb_run_post_processing(
        mesh._mesh, mesh._ptr(r_elevation, "f32", n, "r_elevation"), C.byref(p), float(seed),
        mesh._ptr(r_hotspot, "f32", n, "r_hotspot", True), mesh._ptr(delta, "f32", n, "erosionDelta"),
        mesh._ptr(ocean, "u8", n, "r_isOcean")))
    post = []
    if timing:
        ms = (C.c_double * 5)()
        mesh.lib.check(mesh.lib.dll.pb_last_post_timing(mesh._mesh, ms))
        post = [{"stage": s, "ms": float(v)} for s, v in zip(POST_STAGES, ms)]
    return {"dl_erosionDelta": delta, "postTiming": post, "r_isOcean": ocean}
