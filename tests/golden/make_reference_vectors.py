"""Generates tests/golden/reference_*.npz by EXECUTING THE UNMODIFIED REFERENCE (/root/reference/js/planet-worker.js and every
module it imports) under the minimal evaluator in tests/golden/minijs.py — the reference is browser JavaScript and this image
has no JavaScript runtime.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_reference_vectors.py            # all scenarios, a few minutes

What runs is the reference's own source, byte for byte: the worker's `self.onmessage` receives the same command objects the
web app posts (generate, reapply, computeClimate, editRecompute, importHeightmap) and the messages it posts back are stored.
Two things the reference pulls from outside its tree are supplied by the host, and the vectors inherit them:
  * `delaunator@5.0.1` (CDN import, js/planet-worker.js:17) → oracle/delaunator_ref.py, a restatement of the published
    algorithm.  Every array downstream of the mesh depends only on the reference's code GIVEN that triangulation;
  * `Math.sin/cos/exp/pow/…` → Python's libm (V8 uses an fdlibm port; include/pb_detmath.h is a third implementation).
tests/test_zz_reference_vectors.py compares the oracle, the host emulation and (under -m gpu) the CUDA library with these files.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_JS = "/root/reference/js"
CDN = "https://cdn.jsdelivr.net/npm/delaunator@5.0.1/+esm"

SLIDERS = dict(smoothing=0.10, hydraulicErosion=0.50, thermalErosion=0.10, ridgeSharpening=0.50, glacialErosion=0.50, terrainWarp=0.75)

SCENARIOS = {
    # name: list of commands posted to one worker, in order
    "A_600": [
        dict(cmd="generate", N=600, P=12, jitter=0.75, nMag=0.4, numContinents=3, continentSizeVariety=0.0, seed=42, **SLIDERS),
        dict(cmd="reapply", smoothing=0.3, glacialErosion=0.2, hydraulicErosion=0.7, thermalErosion=0.0, ridgeSharpening=0.2, terrainWarp=0.0,
             skipClimate=True),
        dict(cmd="computeClimate", temperatureOffset=2.5, precipitationOffset=-0.2),
        dict(cmd="editRecompute", nMag=0.55, EDIT=True, **SLIDERS),
    ],
    "B_2500": [
        dict(cmd="generate", N=2500, P=30, jitter=0.6, nMag=0.5, numContinents=4, continentSizeVariety=0.5, landCoverage=0.4,
             temperatureOffset=-1.5, precipitationOffset=0.15, toggledIndices=[1, 4], seed=7, smoothing=0.25, hydraulicErosion=0.8,
             thermalErosion=0.4, ridgeSharpening=0.3, glacialErosion=0.9, terrainWarp=0.4),
    ],
    "C_10000": [      # BASELINE config 1: 10k-cell sphere, default sliders
        dict(cmd="generate", N=10000, P=80, jitter=0.75, nMag=0.4, numContinents=4, continentSizeVariety=0.0, seed=42, **SLIDERS),
    ],
    "D_import_600": [
        dict(cmd="importHeightmap", N=600, jitter=0.75, IMAGE=(64, 32), seed=11, **SLIDERS),
    ],
    "G_200500": [     # above 200 000 regions assignElevation / findCollisions switch to 2 noise octaves (js/elevation.js:55, 457)
        dict(cmd="generate", N=200500, P=80, jitter=0.75, nMag=0.4, numContinents=4, continentSizeVariety=0.0, seed=42, skipClimate=True, **SLIDERS),
    ],
    "H_1000000": [    # the benchmark planet itself (bench.py: 1 000 001 cells, seed 42, slider defaults), climate skipped; ≈ 1.5 h
        dict(cmd="generate", N=1000000, P=80, jitter=0.75, nMag=0.4, numContinents=4, continentSizeVariety=0.0, seed=42, skipClimate=True, **SLIDERS),
    ],
    "J_100000": [     # full pipeline incl. the climate stack at 100 001 cells (≈ 300 graph sweeps per field family), block digests; ≈ 25 min
        dict(cmd="generate", N=100000, P=80, jitter=0.75, nMag=0.4, numContinents=4, continentSizeVariety=0.0, seed=42, **SLIDERS),
    ],
    "E_single_layer_400": [     # P < 8: no super plates (js/planet-worker.js:207), single-layer collisions
        dict(cmd="generate", N=400, P=6, jitter=0.75, nMag=0.4, numContinents=2, continentSizeVariety=0.0, seed=3, **SLIDERS),
    ],
}
# C keeps only the arrays BASELINE's configs name (the rest can be regenerated), G three full arrays plus SHA-256 digests of the
# others (fixture size); the other scenarios keep every array of every reply
KEEP_G = {"r_plate", "prePostElev", "r_elevation"}
KEEP_C = {"r_plate", "prePostElev", "r_elevation", "r_stress", "t_elevation", "r_wind_east_summer", "r_wind_north_winter",
          "r_ocean_warmth_summer", "r_precip_summer", "r_precip_winter", "r_temperature_summer", "r_temperature_winter",
          "debugLayers.erosionDelta", "debugLayers.koppen", "debugLayers.hotspot", "debugLayers.superPlates"}


def synthetic_image(w, h):
    yy, xx = np.mgrid[0:h, 0:w]
    g = 96 + 70 * np.sin(xx * 0.31) * np.cos(yy * 0.23) + 50 * np.sin((xx + 2 * yy) * 0.11)
    return np.clip(g, 0, 255).astype(np.uint8).ravel()


def make_interpreter():
    from oracle.delaunator_ref import delaunator
    from tests.golden import minijs as js

    def delaunator_ctor(args):
        flat = js.to_python(args[0])
        tri, half = delaunator(flat.tolist())
        o = js.JSObject()
        o.props["coords"] = args[0]
        o.props["triangles"] = js.from_python(np.asarray(tri, np.uint32))
        o.props["halfedges"] = js.from_python(np.asarray(half, np.int32))
        return o
    D = js.HostFunction(lambda this, args: js.throw_error("TypeError", "Class constructor Delaunator cannot be invoked without 'new'"),
                        "Delaunator", delaunator_ctor)
    it = js.Interpreter(REFERENCE_JS, host_modules={CDN: {"default": D}})
    _LAST["interp"] = it
    posted = []
    worker_self = js.JSObject()
    worker_self.props["postMessage"] = js.HostFunction(lambda this, args: (posted.append(js.to_python(args[0])), js.UNDEF)[1], "postMessage")
    it.globals["self"] = worker_self
    it.load("planet-worker.js")

    def post(message: dict):
        del posted[:]
        ev = js.JSObject(None, {"data": js.from_python(message)})
        js.call_function(worker_self.props["onmessage"], worker_self, [ev])
        replies = [m for m in posted if m.get("type") != "progress"]
        if len(replies) != 1:
            raise RuntimeError(f"expected one reply, got {[m.get('type') for m in replies]}")
        if replies[0]["type"] == "error":
            raise RuntimeError(f"the reference worker posted an error: {replies[0]['message']}")
        return replies[0]
    return post


def flatten(reply: dict):
    """(arrays, meta): numpy arrays by dotted key; everything else JSON-able"""
    arrays, meta = {}, {}
    for k, v in reply.items():
        if isinstance(v, np.ndarray):
            arrays[k] = v
        elif k in ("debugLayers", "windDebugLayers") and isinstance(v, dict):
            for kk, vv in v.items():
                if isinstance(vv, np.ndarray):
                    arrays[f"{k}.{kk}"] = vv
        elif k.startswith("_"):
            continue            # timings
        else:
            meta[k] = v
    return arrays, meta


def run_scenario(name):
    post = make_interpreter()
    suffix = ""
    if os.environ.get("REF_MATH") == "detmath":      # REF_MATH=detmath python … H_1000000 → reference_H_1000000_detmath.npz
        from tests.golden.fuzz_reference import use_detmath
        use_detmath(_LAST["interp"])
        suffix = "_detmath"
    out, metas, commands = {}, [], []
    last_done = None
    for i, cmd in enumerate(SCENARIOS[name]):
        cmd = dict(cmd)
        if cmd.pop("EDIT", False):
            # what the edit UI sends (js/edit-mode.js): the current ocean set with one plate toggled and one density changed
            seeds = [int(s) for s in last_done["plateSeeds"]]
            ocean = {int(s) for s in last_done["plateIsOcean"]}
            ocean ^= {seeds[2]}
            dens = {str(int(k)): float(v) for k, v in last_done["plateDensity"].items()}
            dens[str(seeds[0])] = 2.95
            cmd.update(plateIsOcean=sorted(ocean), plateDensity=dens)
        if "IMAGE" in cmd:
            w, h = cmd.pop("IMAGE")
            cmd.update(grayscale=synthetic_image(w, h), imageWidth=w, imageHeight=h)
        t0 = time.time()
        reply = post(cmd)
        print(f"  {name}[{i}] {cmd['cmd']}: {reply['type']} in {time.time() - t0:.1f} s", flush=True)
        if reply["type"] == "done":
            last_done = reply
        arrays, meta = flatten(reply)
        if name.startswith("C_"):
            arrays = {k: v for k, v in arrays.items() if k in KEEP_C}
        if name.startswith("G_"):
            import hashlib
            meta["sha256"] = {k: hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() for k, v in arrays.items()}
            arrays = {k: v for k, v in arrays.items() if k in KEEP_G}
        if name[0] in "HJ":
            # 1M cells: no array is stored; per array one 8-byte digest per block of 4096 elements (a difference is localised to blocks)
            import hashlib
            blocks = {}
            for k, v in arrays.items():
                b = np.ascontiguousarray(v)
                blocks[k] = np.frombuffer(b"".join(hashlib.sha256(b[i:i + 4096].tobytes()).digest()[:8] for i in range(0, b.size, 4096)), np.uint8)
            arrays = {"blocks." + k: v for k, v in blocks.items()}
        for k, v in arrays.items():
            out[f"{i}/{k}"] = v
        metas.append(meta)
        commands.append({k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in cmd.items()})
    out["__meta__"] = np.frombuffer(json.dumps({"commands": commands, "replies": metas}, default=lambda o: o.tolist()).encode(), np.uint8)
    path = os.path.join(HERE, f"reference_{name}{suffix}.npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KB, {len(out) - 1} arrays)")


def _extract(src: str, start: str, end: str) -> str:
    a = src.index(start)
    return src[a:src.index(end, a)]


def run_render_scenario():
    """F_render_600: the pure pieces of js/planet-mesh.js — smoothBiomeColors, heightmapColor, landHeightmapColor, landMaskColor,
    koppenColor and the triangle loop of exportMap (:1766-1846) — evaluated on the planet of scenario A.  planet-mesh.js itself
    cannot be loaded (it imports three.js and the DOM-bound scene), so the text of those functions is cut out of the reference
    file HERE, at generation time, and evaluated as a module next to it; nothing of it is stored in the repository."""
    from tests.golden import minijs as js
    post = make_interpreter()
    reply = post(dict(SCENARIOS["A_600"][0]))
    interp = _LAST["interp"]          # the interpreter make_interpreter() just built
    worker = interp.modules[os.path.realpath(os.path.join(REFERENCE_JS, "planet-worker.js"))]
    W = worker.env.v["W"]
    pm = open(os.path.join(REFERENCE_JS, "planet-mesh.js"), encoding="utf-8").read()
    body = _extract(pm, "function smoothBiomeColors(", "// Grayscale heightmap")
    body += _extract(pm, "function heightmapColor(", "// Diverging color map")
    body += _extract(pm, "function koppenColor(", "// Plate colours")
    export_map = pm[pm.index("export async function exportMap("):]
    loop = _extract(export_map, "    const { numSides } = mesh;\n    const PI = Math.PI;", "    const geo = new THREE.BufferGeometry();")
    module = ("import { elevationToColor, elevToHeightKm, biomeColor } from './color-map.js';\n"
              "import { KOPPEN_CLASSES } from './koppen.js';\n" + body +
              "function mapTriangles(mesh, r_xyz, t_xyz, r_elevation, type, koppenArr, biomeSmoothed) {\n" + loop +
              "    return { posArr, colArr, triCount };\n}\n"
              "export { smoothBiomeColors, heightmapColor, landHeightmapColor, landMaskColor, koppenColor, mapTriangles, elevationToColor };\n")
    mod = interp.load(os.path.join(REFERENCE_JS, "__planet_mesh_extract__.js"), source=module)
    fn = lambda n: mod.env.v[mod.exports[n]]          # noqa: E731
    mesh, r_xyz = js.get_prop(W, "mesh"), js.get_prop(W, "r_xyz")
    elev = js.from_python(reply["r_elevation"])
    koppen = js.from_python(reply["debugLayers"]["koppen"])
    t_xyz = js.from_python(reply["t_xyz"])
    n = reply["r_elevation"].size
    out = {}
    smoothed = js.call_function(fn("smoothBiomeColors"), js.UNDEF, [mesh, koppen, elev])
    out["colors.biome"] = js.to_python(smoothed)
    for mode, name in (("terrain", "elevationToColor"), ("heightmap", "heightmapColor"), ("landheightmap", "landHeightmapColor"), ("landmask", "landMaskColor")):
        rgb = np.empty(3 * n, np.float32)
        for r in range(n):
            rgb[3 * r:3 * r + 3] = js.to_python(js.call_function(fn(name), js.UNDEF, [float(reply["r_elevation"][r])]))
        out["colors." + mode] = rgb
    ids = list(range(0, 33)) + [200]
    out["koppenColor.ids"] = np.asarray(ids, np.int32)
    out["koppenColor.rgb"] = np.asarray([js.to_python(js.call_function(fn("koppenColor"), js.UNDEF, [float(i)])) for i in ids], np.float64).ravel()
    for etype in ("biome", "heightmap", "colormap", "koppen"):
        res = js.call_function(fn("mapTriangles"), js.UNDEF, [mesh, r_xyz, t_xyz, elev, etype, koppen, smoothed])
        cnt = int(js.get_prop(res, "triCount"))
        out[f"triangles.{etype}.pos"] = js.to_python(js.get_prop(res, "posArr"))[:9 * cnt]
        out[f"triangles.{etype}.col"] = js.to_python(js.get_prop(res, "colArr"))[:9 * cnt]
    path = os.path.join(HERE, "reference_F_render_600.npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KB, {len(out)} arrays)")


def run_stage_scenario():
    """I_post50_2500: the five exported stage functions of js/terrain-post.js called one by one on scenario B's planet with
    BASELINE config 2's parameters — hIters = 50 passed directly (the sliders only reach 20), K = 0.0003, m = 0.5, dt = 1,
    tIters = 1, gIters = 5 — which puts the second priority flood at iteration 38 (:446).  Plus smoothField and percentile of
    js/climate-util.js.  The elevation after every stage is stored."""
    from tests.golden import minijs as js
    post = make_interpreter()
    reply = post(dict(SCENARIOS["B_2500"][0]))
    interp = _LAST["interp"]
    worker = interp.modules[os.path.realpath(os.path.join(REFERENCE_JS, "planet-worker.js"))]
    W = worker.env.v["W"]
    mesh, r_xyz, nd = js.get_prop(W, "mesh"), js.get_prop(W, "r_xyz"), js.get_prop(W, "neighborDist")
    tp = lambda n: interp.get_export("terrain-post.js", n)          # noqa: E731
    elev = js.from_python(reply["prePostElev"].copy())
    hot = js.from_python(reply["debugLayers"]["hotspot"])
    out = {"in.prePostElev": reply["prePostElev"], "in.hotspot": reply["debugLayers"]["hotspot"], "in.triangles": reply["triangles"],
           "in.halfedges": reply["halfedges"], "in.r_xyz": reply["r_xyz"]}
    t0 = time.time()
    js.call_function(tp("warpTerrain"), js.UNDEF, [mesh, elev, r_xyz, 7.0, 0.75, hot])
    out["1.warpTerrain"] = js.to_python(elev)
    is_ocean = js.from_python((out["1.warpTerrain"] <= 0).astype(np.uint8))
    js.call_function(tp("smoothElevation"), js.UNDEF, [mesh, elev, is_ocean, 1.0, 0.25])
    out["2.smoothElevation"] = js.to_python(elev)
    js.call_function(tp("erodeComposite"), js.UNDEF, [mesh, elev, r_xyz, is_ocean, 50.0, 0.0003, 0.5, 1.0, 1.0, 1.16, 0.015, 5.0, 0.5, nd])
    out["3.erodeComposite"] = js.to_python(elev)
    js.call_function(tp("sharpenRidges"), js.UNDEF, [mesh, elev, is_ocean, 3.0, 0.04])
    out["4.sharpenRidges"] = js.to_python(elev)
    js.call_function(tp("applySoilCreep"), js.UNDEF, [mesh, elev, is_ocean, 3.0, 0.1125])
    out["5.applySoilCreep"] = js.to_python(elev)
    field = js.from_python(reply["r_precip_summer"].copy())
    js.call_function(interp.get_export("climate-util.js", "smoothField"), js.UNDEF, [mesh, field, 7.0])
    out["6.smoothField7"] = js.to_python(field)
    out["7.percentiles"] = np.asarray([js.call_function(interp.get_export("climate-util.js", "percentile"), js.UNDEF,
                                                        [js.from_python(reply["r_elevation"]), p]) for p in (0.0, 0.05, 0.5, 0.95, 0.97, 0.999)], np.float64)
    print(f"  stage functions in {time.time() - t0:.0f} s")
    path = os.path.join(HERE, "reference_I_post50_2500.npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KB, {len(out)} arrays)")


_LAST = {}


if __name__ == "__main__":
    if not os.path.isdir(REFERENCE_JS):
        sys.exit(f"{REFERENCE_JS} not found: the vectors can only be regenerated where the reference is present")
    names = sys.argv[1:] or [n for n in SCENARIOS if n[0] not in "GHJ"] + ["F_render_600", "I_post50_2500"]      # G ≈ 15 min, H ≈ 1.5 h: ask for them by name
    for n in names:
        print(n, flush=True)
        if n == "F_render_600":
            run_render_scenario()
        elif n == "I_post50_2500":
            run_stage_scenario()
        else:
            run_scenario(n)
