"""Parity against THE REFERENCE ITSELF.  tests/golden/reference_*.npz hold the replies of the unmodified reference worker
(/root/reference/js/planet-worker.js + every module it imports) to the web app's own commands — generate, reapply,
computeClimate, editRecompute, importHeightmap — produced in the build container by executing that source under the
minimal ECMAScript evaluator tests/golden/minijs.py (tests/golden/make_reference_vectors.py; the image has no JavaScript
runtime).  The same commands are replayed here through

  * the engine's worker mirror in the reference's neighbour order (`mesh_order="delaunator"`): host emulation of the kernels in the
    CPU suite, the CUDA library under -m gpu, and
  * the oracle's restatement of the handlers (tests/test_worker.py:OracleWorker),

and every array of every reply is compared: integer fields (triangles, halfedges, r_plate, Köppen classes, region sets) must be
identical, and so must the Float32 fields — bit for bit.  The committed vectors need no tolerance at all; the checker allows two
last-bit flips per array within BASELINE's 1e-4 relative because three `Math.*` implementations are in play (Python's libm for the
vectors, include/pb_detmath.h here, V8's fdlibm port in a browser) and the reference amplifies the last bit of one `Math.sin` to
1e-7 at the cell a hotspot dome is centred on (tests/golden/fuzz_reference.py measures it: with the same `Math` on
both sides no value differs).  Inherited by the vectors: the triangulation comes from oracle/delaunator_ref.py (delaunator@5.0.1 is
a CDN import of the reference, not part of its tree)."""
import json
import os

import numpy as np
import pytest

from planet_heightmap_generation_b200.worker import PlanetWorker

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENARIOS = ["A_600", "B_2500", "C_10000", "D_import_600", "E_single_layer_400", "G_200500"]
SET_KEYS = ("mountain_r", "coastline_r", "ocean_r")
REL_TOL = 1e-4            # BASELINE north_star: "float elevation/climate within 1e-4 relative"
MAX_ULP_FLIPS = 2         # Float32 elements per array allowed to differ at all (see the module docstring)


def load(name):
    z = np.load(os.path.join(GOLDEN, f"reference_{name}.npz"))
    meta = json.loads(bytes(z["__meta__"]).decode())
    replies = []
    for i, rmeta in enumerate(meta["replies"]):
        arrays = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(f"{i}/")}
        replies.append((rmeta, arrays))
    return meta["commands"], replies


def command_for(cmd):
    c = dict(cmd)
    if "grayscale" in c:
        c["grayscale"] = np.asarray(c["grayscale"], np.uint8)
    if "plateDensity" in c:
        c["plateDensity"] = {int(k): v for k, v in c["plateDensity"].items()}
    return c


def lookup(reply, dotted):
    v = reply
    for part in dotted.split("."):
        if v is None:
            return None
        v = v.get(part)
    return v


def check_array(where, got, want, stats):
    assert got is not None, f"{where}: missing in the reply"
    got = np.asarray(got)
    assert got.shape == want.shape, f"{where}: shape {got.shape} vs reference {want.shape}"
    if want.dtype.kind == "f":
        g, w = got.astype(np.float32), want.astype(np.float32)
        differ = (g.view(np.uint32) != w.view(np.uint32)) & ~((g == 0) & (w == 0))
        n = int(differ.sum())
        stats["float_elements"] += g.size
        stats["float_differing"] += n
        if n:
            rel = float(np.max(np.abs(g[differ].astype(np.float64) - w[differ]) / np.maximum(np.abs(w[differ]), 1e-3)))
            stats["worst"] = max(stats["worst"], rel)
            assert rel <= REL_TOL, f"{where}: {n} of {g.size} differ, worst relative {rel:.3g}"
            assert n <= MAX_ULP_FLIPS, f"{where}: {n} of {g.size} Float32 values differ from the reference (allowed {MAX_ULP_FLIPS})"
    else:
        bad = got.astype(np.int64) != want.astype(np.int64)
        stats["int_elements"] += got.size
        assert not bad.any(), f"{where}: {int(bad.sum())} of {got.size} integer values differ from the reference"


def check_reply(name, i, reply, rmeta, arrays, stats):
    assert reply["type"] == rmeta["type"], f"{name}[{i}]: {reply}"
    for key, want in arrays.items():
        check_array(f"{name}[{i}].{key}", lookup(reply, key), want, stats)
    for key in SET_KEYS:
        if key in rmeta:
            assert sorted(int(r) for r in reply[key]) == sorted(int(r) for r in rmeta[key]), f"{name}[{i}].{key}"
    if "plateSeeds" in rmeta:
        assert [int(s) for s in reply["plateSeeds"]] == [int(s) for s in rmeta["plateSeeds"]], "plateSeeds (Set order)"
        assert sorted(int(s) for s in reply["plateIsOcean"]) == sorted(int(s) for s in rmeta["plateIsOcean"]), "plateIsOcean"
    for key in ("plateDensity", "plateDensityLand", "plateDensityOcean"):
        if rmeta.get(key):
            assert {int(k): v for k, v in reply[key].items()} == {int(k): v for k, v in rmeta[key].items()}, key
    if rmeta.get("plateVec"):
        # Euler poles are doubles computed with sin / cos / acos: the three Math implementations differ in the last ulp of a double
        for pid, pv in rmeta["plateVec"].items():
            mine = reply["plateVec"][int(pid)]
            if isinstance(pv, dict):
                assert np.allclose(mine["pole"], pv["pole"], rtol=1e-13, atol=1e-15) and abs(mine["omega"] - pv["omega"]) <= 1e-13 * abs(pv["omega"]), f"plateVec[{pid}]"
            else:                         # importHeightmap posts zero vectors (js/planet-worker.js:838-840)
                assert list(mine if not isinstance(mine, dict) else mine["pole"]) == list(pv), f"plateVec[{pid}]"
    for key in ("numRegions", "seed", "nMag", "skipClimate"):
        if key in rmeta:
            assert reply[key] == rmeta[key], key
    # the 200 501-cell scenario stores three arrays in full and SHA-256 digests of all of them
    import hashlib
    for key, digest in (rmeta.get("sha256") or {}).items():
        v = lookup(reply, key)
        assert v is not None, key
        got = hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest()
        if got != digest and key not in arrays:
            raise AssertionError(f"{name}[{i}].{key}: SHA-256 differs from the reference's array")
        stats["sha_checked"] = stats.get("sha_checked", 0) + 1


def test_vectors_are_complete():
    for name in SCENARIOS:
        commands, replies = load(name)
        assert len(commands) == len(replies) >= 1
        assert all(r[0]["type"] in ("done", "reapplyDone", "climateDone", "editDone") for r in replies)
    commands, replies = load("A_600")
    assert [c["cmd"] for c in commands] == ["generate", "reapply", "computeClimate", "editRecompute"]
    assert len(replies[0][1]) >= 50 and replies[0][1]["r_elevation"].size == 601
    assert {"triangles", "halfedges", "r_xyz", "t_xyz", "r_plate", "prePostElev", "r_elevation", "r_stress", "debugLayers.koppen"} <= set(replies[0][1])


@pytest.mark.parametrize("name", SCENARIOS)
def test_engine_replays_the_reference_worker(backend, name):
    commands, replies = load(name)
    w = PlanetWorker(lib=backend, mesh_order="delaunator")
    stats = dict(float_elements=0, float_differing=0, int_elements=0, worst=0.0)
    for i, (cmd, (rmeta, arrays)) in enumerate(zip(commands, replies)):
        reply = w.onmessage(command_for(cmd))
        check_reply(name, i, reply, rmeta, arrays, stats)
    w.close()
    assert stats["float_elements"] > 5000 and stats["int_elements"] > 400
    print(f"{name}: {stats}")


@pytest.mark.parametrize("name", ["A_600", "B_2500", "E_single_layer_400", "J_100000"])
def test_oracle_replays_the_reference_worker(oracle, name):
    """The oracle's handlers (the checker every other parity test trusts) against the reference's own replies."""
    from tests.test_worker import OracleWorker
    commands, replies = load(name)
    ow = OracleWorker(oracle, mesh_order="delaunator")
    stats = dict(float_elements=0, float_differing=0, int_elements=0, worst=0.0)
    cmd, (rmeta, ref) = commands[0], replies[0]
    elev, delta, koppen = ow.generate(cmd)
    got = {"r_xyz": ow.xyz, "triangles": ow.mesh.triangles, "halfedges": ow.mesh.halfedges, "r_plate": ow.r_plate, "prePostElev": ow.pre,
           "r_elevation": elev, "r_stress": ow.oe.get("r_stress"), "t_xyz": oracle.triangle_centers(ow.mesh, ow.xyz),
           "debugLayers.erosionDelta": delta, "debugLayers.koppen": koppen}
    for k in ("base", "tectonic", "noise", "interior", "coastal", "ocean", "hotspot", "tecActivity", "margins", "backArc", "foldRidge", "orogenicPower"):
        got["debugLayers." + k] = ow.oe.get(k)
    for k in ("r_wind_east_summer", "r_wind_north_summer", "r_wind_east_winter", "r_wind_north_winter", "itczLons", "itczLatsSummer",
              "itczLatsWinter", "r_ocean_current_east_summer", "r_ocean_current_north_winter", "r_ocean_speed_summer", "r_ocean_speed_winter",
              "r_ocean_warmth_summer", "r_ocean_warmth_winter", "r_precip_summer", "r_precip_winter", "r_temperature_summer", "r_temperature_winter"):
        got[k] = ow.clim.get(k)
    if name[0] == "J":        # block digests instead of arrays (100 001 cells)
        import hashlib
        for k, v in got.items():
            v = np.ascontiguousarray(v)
            if k in ("triangles", "halfedges", "r_plate"):
                v = v.astype(np.int32)
            d = np.frombuffer(b"".join(hashlib.sha256(v[i:i + 4096].tobytes()).digest()[:8] for i in range(0, v.size, 4096)), np.uint8)
            assert (d == ref["blocks." + k]).all(), f"oracle {name}.{k}: block digests differ from the reference's array"
            stats["float_elements"] += v.size
    else:
        for k, v in got.items():
            check_array(f"oracle {name}.{k}", v, ref[k], stats)
    assert [int(s) for s in rmeta["plateSeeds"]] == ow.seeds
    assert sorted(int(s) for s in rmeta["plateIsOcean"]) == sorted(ow.pio)
    assert {int(k): v for k, v in rmeta["plateDensity"].items()} == ow.dens
    for key in SET_KEYS:
        assert sorted(int(r) for r in rmeta[key]) == [int(r) for r in np.nonzero(ow.oe.get(key, np.uint8))[0]], key
    print(f"oracle {name}: {stats}")


def _planet_A():
    """mesh, r_xyz and the final arrays of scenario A's generate reply"""
    from planet_heightmap_generation_b200.mesh import SphereMesh
    _, replies = load("A_600")
    ref = replies[0][1]
    mesh = SphereMesh(ref["triangles"], ref["halfedges"], int(replies[0][0]["numRegions"]))
    return mesh, ref["r_xyz"], ref["r_elevation"], ref["debugLayers.koppen"]


def test_colour_ramps_match_the_reference(backend, oracle):
    """elevationToColor (js/color-map.js:116-125), smoothBiomeColors over biomeColor, heightmapColor, landHeightmapColor,
    landMaskColor, koppenColor (js/planet-mesh.js:30-80, 175-178) evaluated from the reference's source: Float32 colour buffers."""
    from planet_heightmap_generation_b200.engine import DeviceMesh
    z = np.load(os.path.join(GOLDEN, "reference_F_render_600.npz"))
    mesh, xyz, elev, koppen = _planet_A()
    dm = DeviceMesh(mesh, xyz, lib=backend)
    stats = dict(float_elements=0, float_differing=0, int_elements=0, worst=0.0)
    for mode in ("terrain", "biome", "heightmap", "landheightmap", "landmask"):
        want = z["colors." + mode]
        check_array(f"oracle colours {mode}", oracle.region_colors(mesh, mode, elev, koppen), want, stats)
        check_array(f"engine colours {mode}", dm.regionColors(mode, elev, koppen if mode == "biome" else None), want, stats)
    ids = z["koppenColor.ids"]
    k = np.zeros(mesh.numRegions, np.uint8)
    k[:ids.size] = ids.astype(np.uint8)
    want = z["koppenColor.rgb"].astype(np.float32)
    check_array("oracle koppenColor", oracle.region_colors(mesh, "koppen", elev, k)[:3 * ids.size], want, stats)
    check_array("engine koppenColor", dm.regionColors("koppen", elev, k)[:3 * ids.size], want, stats)
    assert stats["float_differing"] == 0
    dm.close()


@pytest.mark.parametrize("etype", ["biome", "heightmap", "colormap", "koppen"])
def test_export_triangles_match_the_reference(oracle, etype):
    """The triangle loop of exportMap (js/planet-mesh.js:1766-1846) run from the reference's source on scenario A's planet: the
    Float32 position and colour buffers of the oracle's export (the pixels of the engine are compared with the oracle's in
    tests/test_export_map.py; the rasteriser between the two is WebGL in the reference, a stated rule here)."""
    z = np.load(os.path.join(GOLDEN, "reference_F_render_600.npz"))
    mesh, xyz, elev, koppen = _planet_A()
    pos, col = oracle.export_map_triangles(mesh, xyz, etype, elev, koppen)
    stats = dict(float_elements=0, float_differing=0, int_elements=0, worst=0.0)
    check_array(f"posArr {etype}", pos, z[f"triangles.{etype}.pos"], stats)
    check_array(f"colArr {etype}", col, z[f"triangles.{etype}.col"], stats)
    assert stats["float_differing"] == 0 and stats["float_elements"] > 60000


@pytest.mark.parametrize("name", ["H_1000000", "H_1000000_detmath", "J_100000"])
def test_engine_reproduces_the_reference_block_digests(backend, name):
    """Planets too large to store.  H: the planet bench.py times — 1 000 001 cells, seed 42, slider defaults — generated by the
    reference worker with the climate skipped (820 graph sweeps at that size are beyond the evaluator).  J: the whole pipeline
    including the climate stack at 100 001 cells.  Per array one 8-byte digest per block of 4096 elements is stored; every block of
    every array — integer and Float32 alike — must match, with the one documented exception below.  H_1000000_detmath is the same
    reference run with include/pb_detmath.h behind the evaluator's Math.*: there every one of the 10 027 blocks must match."""
    import hashlib
    path = os.path.join(GOLDEN, f"reference_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"reference_{name}.npz has not been generated (an hour under the evaluator)")
    commands, replies = load(name)
    rmeta, blocks = replies[0]
    w = PlanetWorker(lib=backend, mesh_order="delaunator")
    reply = w.onmessage(command_for(commands[0]))
    assert reply["type"] == "done", reply
    assert [int(s) for s in reply["plateSeeds"]] == [int(s) for s in rmeta["plateSeeds"]]
    assert sorted(int(s) for s in reply["plateIsOcean"]) == sorted(int(s) for s in rmeta["plateIsOcean"])
    for key in SET_KEYS:
        assert sorted(int(r) for r in reply[key]) == sorted(int(r) for r in rmeta[key]), key
    checked, report = 0, {}
    for key, want in blocks.items():
        arr = key.split(".", 1)[1]
        v = np.ascontiguousarray(lookup(reply, arr))
        got = np.frombuffer(b"".join(hashlib.sha256(v[i:i + 4096].tobytes()).digest()[:8] for i in range(0, v.size, 4096)), np.uint8)
        assert got.size == want.size, f"{arr}: {v.size} elements"
        differing = int((got.reshape(-1, 8) != want.reshape(-1, 8)).any(axis=1).sum())
        if differing:
            report[arr] = (differing, want.size // 8)
        checked += 1
    w.close()
    assert checked >= 20
    # J matches in every block.  On H one block of one array differs: debugLayers.hotspot holds the uplift of the cell a dome is
    # centred on, whose last Float32 bit follows the last bit of one Math.sin (libm produced the vectors,
    # pb_detmath.h runs here) — the sum r_elevation[r] += uplift absorbs it, every other array is identical.
    allowed = {"H_1000000": {"debugLayers.hotspot": 1}}.get(name, {})
    for arr, (differing, total) in report.items():
        assert differing <= allowed.get(arr, 0), f"{name}.{arr}: {differing} of {total} blocks of 4096 elements differ from the reference's array"


def test_stage_functions_with_50_stream_power_iterations(backend, oracle):
    """js/terrain-post.js's five exported stage functions (:233, 317, 369, 713, 758) called directly by the generator with BASELINE
    config 2's parameters — hIters = 50 (second priority flood at iteration 38, :446), K 0.0003, m 0.5, dt 1, tIters 1, gIters 5 —
    plus smoothField / percentile of js/climate-util.js (:5, 103).  Every stage starts from the reference's output of the stage
    before, in the engine and in the oracle; then the whole chain end to end."""
    from planet_heightmap_generation_b200 import terrain_post as tp
    from planet_heightmap_generation_b200.climate_util import smoothField
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200.mesh import SphereMesh
    z = np.load(os.path.join(GOLDEN, "reference_I_post50_2500.npz"))
    xyz = z["in.r_xyz"]
    mesh = SphereMesh(z["in.triangles"], z["in.halfedges"], xyz.size // 3)
    nd = oracle.neighbor_dist(mesh, xyz)
    dm = DeviceMesh(mesh, xyz, lib=backend)
    hot = z["in.hotspot"]
    is_ocean = (z["1.warpTerrain"] <= 0).astype(np.uint8)
    stats = dict(float_elements=0, float_differing=0, int_elements=0, worst=0.0)
    stages = [
        ("1.warpTerrain", "in.prePostElev", lambda e: tp.warpTerrain(dm, e, xyz, 7, 0.75, hot), lambda e: oracle.warp_terrain(mesh, e, xyz, 7, 0.75, hot)),
        ("2.smoothElevation", "1.warpTerrain", lambda e: tp.smoothElevation(dm, e, is_ocean, 1, 0.25), lambda e: oracle.smooth_elevation(mesh, e, is_ocean, 1, 0.25)),
        ("3.erodeComposite", "2.smoothElevation", lambda e: tp.erodeComposite(dm, e, xyz, is_ocean, 50, 0.0003, 0.5, 1, 1, 1.16, 0.015, 5, 0.5, nd),
         lambda e: oracle.erode_composite(mesh, e, xyz, is_ocean, 50, 0.0003, 0.5, 1, 1, 1.16, 0.015, 5, 0.5, nd)),
        ("4.sharpenRidges", "3.erodeComposite", lambda e: tp.sharpenRidges(dm, e, is_ocean, 3, 0.04), lambda e: oracle.sharpen_ridges(mesh, e, is_ocean, 3, 0.04)),
        ("5.applySoilCreep", "4.sharpenRidges", lambda e: tp.applySoilCreep(dm, e, is_ocean, 3, 0.1125), lambda e: oracle.apply_soil_creep(mesh, e, is_ocean, 3, 0.1125)),
    ]
    chain_engine, chain_oracle = z["in.prePostElev"].copy(), z["in.prePostElev"].copy()
    for name, src, engine_fn, oracle_fn in stages:
        a, b = z[src].copy(), z[src].copy()
        engine_fn(a)
        oracle_fn(b)
        check_array("engine " + name, a, z[name], stats)
        check_array("oracle " + name, b, z[name], stats)
        engine_fn(chain_engine)
        oracle_fn(chain_oracle)
    check_array("engine chain", chain_engine, z["5.applySoilCreep"], stats)
    check_array("oracle chain", chain_oracle, z["5.applySoilCreep"], stats)
    assert (z["3.erodeComposite"] != z["2.smoothElevation"]).mean() > 0.2          # the erosion did something
    _, replies = load("B_2500")
    f1, f2 = replies[0][1]["r_precip_summer"].copy(), replies[0][1]["r_precip_summer"].copy()
    smoothField(dm, f1, 7)
    oracle.smooth_field(mesh, f2, 7)
    check_array("engine smoothField", f1, z["6.smoothField7"], stats)
    check_array("oracle smoothField", f2, z["6.smoothField7"], stats)
    elev = replies[0][1]["r_elevation"]
    got = [oracle.percentile(elev, p) for p in (0.0, 0.05, 0.5, 0.95, 0.97, 0.999)]
    assert got == z["7.percentiles"].tolist()
    assert stats["float_differing"] == 0
    dm.close()
