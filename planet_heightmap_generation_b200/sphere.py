"""Host-side generators for synthetic seeded planets (inputs of the hot path, not part of it).

`fibonacci_sphere` follows js/sphere-mesh.js:9-37 (generateFibonacciSphere with makeRng jitter) in
vectorised numpy; `build_sphere` adds the pole vertex and triangulates on the device.  numpy's
sin/cos/asin may differ from V8's in the last ulp of a double, which can flip an f32 store for a
handful of points — irrelevant here because the same r_xyz array is handed to every implementation
being compared.
"""
from __future__ import annotations

import numpy as np


_M = 2147483647


def park_miller(seed: float, n: int) -> np.ndarray:
    """First n outputs of makeRng(seed) (js/rng.js:3-6), vectorised by jump-ahead."""
    s0 = int(abs(np.floor(seed * 9301 + 49297)) % 2147483646) + 1
    out = np.empty(n, np.float64)
    B = 1 << 16
    mult = np.empty(B, np.uint64)
    a = 1
    for k in range(B):
        a = (a * 16807) % _M
        mult[k] = a
    s = s0
    for start in range(0, n, B):
        m = min(B, n - start)
        states = (np.uint64(s) * mult[:m]) % np.uint64(_M)
        out[start:start + m] = (states.astype(np.float64) - 1.0) / 2147483646.0
        s = int(states[m - 1])
    return out


def fibonacci_sphere(N: int, jitter: float, seed: float) -> np.ndarray:
    """r_xyz float32[3N] (js/sphere-mesh.js:9-37)."""
    k = np.arange(N, dtype=np.float64)
    s = 3.6 / np.sqrt(N)
    dlong = np.pi * (3 - np.sqrt(5.0))
    dz = 2.0 / N
    # z is accumulated by repeated subtraction in the reference; reproduce that rounding
    z = (1 - dz / 2) - np.concatenate(([0.0], np.cumsum(np.full(N - 1, dz))))
    lng = np.concatenate(([0.0], np.cumsum(np.full(N - 1, dlong))))
    r = np.sqrt(1 - z * z)
    lat = np.arcsin(z) * 180 / np.pi
    lon = lng * 180 / np.pi
    if jitter > 0:
        u = park_miller(seed, 4 * N).reshape(N, 4)
        jlat = u[:, 0] - u[:, 1]
        jlon = u[:, 2] - u[:, 3]
        nextz = np.maximum(-1, z - dz * 2 * np.pi * r / s)
        lat = lat + jitter * jlat * (lat - np.arcsin(nextz) * 180 / np.pi)
        lon = lon + jitter * jlon * (s / r * 180 / np.pi)
    latr = lat * np.pi / 180
    lonr = lon * np.pi / 180
    out = np.empty((N, 3), np.float32)
    out[:, 0] = np.cos(latr) * np.cos(lonr)
    out[:, 1] = np.cos(latr) * np.sin(lonr)
    out[:, 2] = np.sin(latr)
    return out.reshape(-1)


def sphere_points(N: int, jitter: float, seed: float) -> np.ndarray:
    """N Fibonacci points + the pole vertex (0,0,1) as id N (js/sphere-mesh.js:175, 179-183)."""
    xyz = np.empty(3 * (N + 1), np.float32)
    xyz[:3 * N] = fibonacci_sphere(N, jitter, seed)
    xyz[3 * N:] = (0, 0, 1)
    return xyz


def build_sphere(N: int, jitter: float, seed: float, device: int = 0, lib=None):
    """buildSphere (js/sphere-mesh.js:174-186) → {mesh, r_xyz}: the triangulation runs on the GPU
    (csrc/pb_meshgen.h) and the returned DeviceMesh carries adjOffset / adjList like the reference's mesh."""
    from .engine import DeviceMesh
    dm = DeviceMesh.build_sphere(N, jitter, seed, device=device, lib=lib)
    return {"mesh": dm, "r_xyz": dm.r_xyz}


def synthetic_elevation(r_xyz: np.ndarray, seed: int, land_fraction: float = 0.3) -> np.ndarray:
    """Seeded stand-in for assignElevation's output (`prePostElev`): a multi-octave random plane-wave
    field on the sphere, shifted so that `land_fraction` of the cells are above sea level, with
    continental relief of the reference's order of magnitude (land up to ≈1, ocean down to ≈-1)."""
    p = np.asarray(r_xyz, np.float32).reshape(-1, 3).astype(np.float64)
    rng = np.random.default_rng(seed)
    h = np.zeros(p.shape[0])
    amp_sum = 0.0
    for octave in range(7):
        f = 1.6 * 2.0 ** octave
        a = 0.62 ** octave
        for _ in range(4):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            h += a * np.sin(f * (p @ d) + rng.uniform(0, 2 * np.pi))
        amp_sum += 2 * a
    h /= amp_sum
    sea = np.quantile(h, 1 - land_fraction)
    h = h - sea
    h = np.where(h > 0, h / max(h.max(), 1e-9), h / max(-h.min(), 1e-9))
    land = h > 0
    h[land] = h[land] ** 0.8 * 0.9
    return h.astype(np.float32)


def synthetic_plates(r_xyz: np.ndarray, elevation: np.ndarray, seed: int, n_plates: int = 40):
    """Seeded stand-in for the plate pipeline's outputs that the climate stage reads: `r_plate`
    (per-cell plate id = a cell id, like the reference's plate seeds) and `plateIsOcean` (set of plate
    ids).  Plates are spherical Voronoi regions of random seed cells; a plate is oceanic when most of
    its cells are below sea level, so continental shelves and inland seas exist as in the reference."""
    p = np.asarray(r_xyz, np.float32).reshape(-1, 3)
    rng = np.random.default_rng(seed + 7919)
    seeds = np.sort(rng.choice(p.shape[0], size=n_plates, replace=False)).astype(np.int32)
    owner = np.empty(p.shape[0], np.int64)
    B = 1 << 18
    for s0 in range(0, p.shape[0], B):
        owner[s0:s0 + B] = np.argmax(p[s0:s0 + B] @ p[seeds].T, axis=1)
    r_plate = seeds[owner].astype(np.int32)
    below = np.bincount(owner, weights=(np.asarray(elevation) <= 0), minlength=n_plates)
    size = np.bincount(owner, minlength=n_plates)
    plate_is_ocean = {int(seeds[k]) for k in range(n_plates) if below[k] > 0.5 * size[k]}
    return r_plate, plate_is_ocean


def synthetic_plate_tables(r_xyz: np.ndarray, elevation: np.ndarray, seed: int, n_plates: int = 40, n_super: int = 10):
    """Seeded stand-ins for the plate pipeline's outputs read by assignElevation (js/planet-worker.js:176-216):
    r_plate, plate table {pid: isOcean, pole, omega, density} in plateSeeds order, and superPlateData
    (r_superPlate + super-plate table).  Poles are random unit vectors, omega in ±[0.5, 1.5], densities as the
    worker draws them (3.0–3.5 oceanic, 2.4–2.9 continental)."""
    r_plate, pio = synthetic_plates(r_xyz, elevation, seed, n_plates)
    rng = np.random.default_rng(seed + 104729)
    ids = [int(p) for p in np.unique(r_plate)]
    rng.shuffle(ids)                                   # plateSeeds is a Set: insertion order is not sorted
    plates = {}
    for pid in ids:
        pole = rng.normal(size=3)
        pole /= np.linalg.norm(pole)
        oc = pid in pio
        plates[pid] = dict(isOcean=oc, pole=tuple(float(v) for v in pole),
                           omega=float(rng.uniform(0.5, 1.5) * rng.choice([-1, 1])),
                           density=float((3.0 if oc else 2.4) + rng.uniform(0, 0.5)))
    # super plates: group plates by nearest of n_super random directions
    p = np.asarray(r_xyz, np.float32).reshape(-1, 3)
    centers = rng.normal(size=(n_super, 3))
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
    plate_center = {pid: p[pid].astype(np.float64) for pid in ids}
    group = {pid: int(np.argmax(centers @ plate_center[pid])) for pid in ids}
    super_plates = {}
    for g in sorted(set(group.values())):
        members = [pid for pid in ids if group[pid] == g]
        pole = rng.normal(size=3)
        pole /= np.linalg.norm(pole)
        oc = sum(plates[m]["isOcean"] for m in members) * 2 > len(members)
        super_plates[g] = dict(isOcean=oc, pole=tuple(float(v) for v in pole), omega=float(rng.uniform(0.5, 1.5) * rng.choice([-1, 1])),
                               density=float(np.mean([plates[m]["density"] for m in members])))
    lut = np.zeros(int(max(ids)) + 1, np.int32)
    for pid in ids:
        lut[pid] = group[pid]
    r_super = lut[r_plate].astype(np.int32)
    return r_plate, plates, ids, r_super, super_plates


def synthetic_coarse_plates(coarse_xyz: np.ndarray, seed: int, n_plates: int = 40, ocean_fraction: float = 0.7):
    """Seeded stand-in for generatePlates + assignOceanLand on the coarse mesh (js/coarse-plates.js:19-39): what
    projectCoarsePlates / smoothAndReconnectPlates / buildSuperPlates / assignElevation receive from the coarse stage.
    Returns (coarse_r_plate, plateSeeds list in Set order, plateVec, plateIsOcean, plateDensity).  Plates are weighted
    spherical Voronoi regions of random seed regions (plate id = seed region id); plates turn oceanic in random order
    until `ocean_fraction` of the area is ocean; poles are random unit vectors, omega in ±[0.5, 2.0) (js/plates.js:226-227),
    densities are the worker's own draws makeRng(r + 777) (js/planet-worker.js:196-201)."""
    p = np.asarray(coarse_xyz, np.float32).reshape(-1, 3).astype(np.float64)
    nc = p.shape[0]
    rng = np.random.default_rng(seed + 7919)
    seeds = [int(s) for s in rng.choice(nc, size=n_plates, replace=False)]
    weight = 1.0 + 0.35 * rng.random(n_plates)
    owner = np.argmax((p @ p[seeds].T) * weight, axis=1)
    owner[seeds] = np.arange(n_plates)
    coarse_r_plate = np.asarray(seeds, np.int32)[owner]
    area = np.bincount(owner, minlength=n_plates)
    plate_is_ocean, acc = set(), 0
    for k in rng.permutation(n_plates):
        if acc >= ocean_fraction * nc:
            break
        plate_is_ocean.add(seeds[k])
        acc += int(area[k])
    plate_vec, plate_density = {}, {}
    for s in seeds:
        pole = rng.normal(size=3)
        pole /= np.linalg.norm(pole)
        plate_vec[s] = {"pole": [float(v) for v in pole], "omega": float((0.5 + rng.random() * 1.5) * (-1 if rng.random() < 0.5 else 1))}
        d = park_miller(s + 777, 2)
        plate_density[s] = float(3.0 + d[0] * 0.5) if s in plate_is_ocean else float(2.4 + d[1] * 0.5)
    return coarse_r_plate, seeds, plate_vec, plate_is_ocean, plate_density
