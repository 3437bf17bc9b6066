"""Regenerates tests/golden/oracle_checksums.json: SHA-256 of every oracle output on a 3 000-cell seeded planet.

The reference has no golden vectors and cannot run here (no JS runtime), so these pins do not prove parity with
the reference; they freeze the oracle's behaviour so that an accidental change to oracle/ (or to pb_detmath.h)
is caught by `pytest -m "not gpu"`.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

SLIDERS = dict(smoothing=0.10, glacialErosion=0.50, hydraulicErosion=0.50, thermalErosion=0.10, ridgeSharpening=0.50,
               terrainWarp=0.75)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute():
    from oracle import binding as oracle
    from oracle.mesh_hull import build_sphere_from_points
    from planet_heightmap_generation_b200.sphere import synthetic_elevation, synthetic_plate_tables
    xyz = oracle.fibonacci_sphere(3000, 0.75, 42)
    mesh, xyz = build_sphere_from_points(xyz)
    out = {"r_xyz": sha(xyz), "adjOffset": sha(mesh.adjOffset), "adjList": sha(mesh.adjList)}
    nd = oracle.neighbor_dist(mesh, xyz)
    out["neighborDist"] = sha(nd)
    r_plate, plates, seeds, r_super, sp = synthetic_plate_tables(xyz, synthetic_elevation(xyz, 42, 0.3), 42)
    out["r_plate"] = sha(r_plate)
    oe = oracle.Elevation(mesh, xyz)
    oe.assign(r_plate, plates, seeds, 42, 0.4, 42, 5, r_super, sp)
    elev = oe.get("r_elevation")
    for k in ("r_elevation", "r_stress", "dist_mountain", "dist_ocean", "dist_coastline", "dist_coast", "dist_coast_land",
              "dBdry", "backArcDist", "hotspot", "coastal", "noise", "tectonic"):
        out["elevation." + k] = sha(oe.get(k))
    delta, ocean = oracle.run_post_processing(mesh, xyz, elev, SLIDERS, nd, 42, oe.get("hotspot"))
    out["post.r_elevation"] = sha(elev)
    out["post.erosionDelta"] = sha(delta)
    out["post.r_isOcean"] = sha(ocean)
    pio = {p for p, v in plates.items() if v["isOcean"]}
    oc = oracle.Climate(mesh, xyz)
    out["climate.r_koppen"] = sha(oc.run_all(elev, pio, r_plate, 42))
    for k in ("r_continentality", "r_pressure_summer", "r_wind_east_winter", "r_wind_speed_summer", "itczLatsSummer",
              "r_ocean_warmth_summer", "r_ocean_speed_winter", "r_precip_summer", "r_precip_winter", "r_rainshadow_summer",
              "r_temperature_summer", "r_temperature_winter"):
        out["climate." + k] = sha(oc.get(k))
    out["climate.r_coastDistLand"] = sha(oc.get("r_coastDistLand", np.int32))
    # equirectangular export of that planet (js/planet-mesh.js:1752-1950): owner side per pixel + RGBA, 256 x 128
    koppen = oc.run_all(elev, pio, r_plate, 42)
    for etype in ("colormap", "biome", "koppen", "heightmap", "landmask"):
        px, side = oracle.export_map(mesh, xyz, etype, 256, elev, koppen)
        out["export." + etype] = sha(px)
    out["export.pixelSide"] = sha(side)
    # plate pipeline (coarse stage on a 4 000-region coarse mesh, projection, smoothing, super plates)
    cp = oracle.generate_coarse_plates(42, 24, 3, 0.5, 0.3, n_coarse=4000)
    out["plates.coarse_xyz"] = sha(cp["coarse_xyz"])
    out["plates.coarse_adjList"] = sha(cp["coarseMesh"].adjList)
    out["plates.coarse_r_plate"] = sha(cp["coarse_r_plate"])
    out["plates.seeds"] = sha(np.asarray(cp["coarsePlateSeeds"], np.int32))
    out["plates.poles"] = sha(np.asarray([cp["coarsePlateVec"][s]["pole"] + [cp["coarsePlateVec"][s]["omega"]] for s in cp["coarsePlateSeeds"]]))
    out["plates.isOcean"] = sha(np.asarray([s in cp["coarsePlateIsOcean"] for s in cp["coarsePlateSeeds"]], np.uint8))
    rp = oracle.project_coarse_plates(mesh, xyz, cp["coarseMesh"], cp["coarse_xyz"], cp["coarse_r_plate"], 42, 24)
    out["plates.projected"] = sha(rp)
    oracle.smooth_and_reconnect_plates(mesh, rp, cp["coarsePlateSeeds"], 3)
    out["plates.smoothed"] = sha(rp)
    table = {s: dict(isOcean=s in cp["coarsePlateIsOcean"], pole=tuple(cp["coarsePlateVec"][s]["pole"]),
                     omega=cp["coarsePlateVec"][s]["omega"], density=2.5 + 0.01 * i) for i, s in enumerate(cp["coarsePlateSeeds"])}
    rs, spt = oracle.build_super_plates(mesh, rp, table)
    out["plates.r_superPlate"] = sha(rs)
    out["plates.superTable"] = sha(np.asarray([list(spt[k]["pole"]) + [spt[k]["omega"], spt[k]["density"], float(spt[k]["isOcean"])] for k in sorted(spt)]))
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_checksums.json")
    json.dump(compute(), open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)
