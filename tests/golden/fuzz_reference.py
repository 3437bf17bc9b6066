"""Randomised comparison of the engine with THE REFERENCE ITSELF (build container only: needs /root/reference).

Every round draws (N, seed, plates, continents, variety, land coverage, jitter, offsets, sliders incl. the 0 / 1 extremes), posts
generate → reapply → editRecompute → computeClimate to the unmodified reference worker running under tests/golden/minijs.py and to
the engine's worker mirror (host emulation of the kernels, reference neighbour order), and compares every array of every reply
with the criteria of tests/test_zz_reference_vectors.py (integers identical; Float32 identical up to two last-bit flips per array
and 1e-4 relative).  Usage:  python tests/golden/fuzz_reference.py [rounds] [first_seed]
Environment: FUZZ_NMIN / FUZZ_NMAX (cells), FUZZ_IMPORT=1 (importHeightmap rounds only), FUZZ_MATH=detmath (the evaluator's Math.* is
include/pb_detmath.h from the start: every Float32 value must then be identical).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from planet_heightmap_generation_b200._lib import Library  # noqa: E402
from planet_heightmap_generation_b200.worker import PlanetWorker  # noqa: E402
from tests.emul.build_emul import build  # noqa: E402
from tests.golden.make_reference_vectors import _LAST, flatten, make_interpreter  # noqa: E402
from tests.test_zz_reference_vectors import check_reply  # noqa: E402

SLIDER_KEYS = ("smoothing", "glacialErosion", "hydraulicErosion", "thermalErosion", "ridgeSharpening", "terrainWarp")
NRANGE = (int(os.environ.get("FUZZ_NMIN", 300)), int(os.environ.get("FUZZ_NMAX", 2500)))


def use_detmath(interp=None):
    """replaces Math.exp/log/pow/atan/asin/acos/atan2/sin/cos/tanh of the evaluator just built by include/pb_detmath.h (through the
    oracle library): with the SAME elementary functions on both sides any remaining difference would be algorithmic"""
    import ctypes as C
    from oracle import binding as ob
    from tests.golden import minijs as js
    lib = ob.lib()

    def det(kind):
        x, y, o = (C.c_double * 1)(), (C.c_double * 1)(), (C.c_double * 1)()

        def f(this, args):
            x[0] = js.to_num(args[0]) if args else float("nan")
            y[0] = js.to_num(args[1]) if len(args) > 1 else 0.0
            lib.orc_detmath(kind, 1, x, y, o)
            return o[0]
        return f
    m = (interp or _LAST["interp"]).globals["Math"]
    for name, kind in (("exp", 0), ("log", 1), ("pow", 2), ("atan", 3), ("asin", 4), ("atan2", 5), ("sin", 6), ("cos", 7), ("tanh", 8)):
        m.props[name] = js.HostFunction(det(kind), name)
    asin = det(4)
    m.props["acos"] = js.HostFunction(lambda this, args: 3.141592653589793 * 0.5 - asin(this, args), "acos")


def replay(gen, commands_fn, lib, detmath, stats):
    """one planet through both workers; raises AssertionError on a difference"""
    post = make_interpreter()
    if detmath:
        use_detmath()
    w = PlanetWorker(lib=lib, mesh_order="delaunator")
    try:
        ref = post(dict(gen))
        mine = w.onmessage(dict(gen))
        arrays, meta = flatten(ref)
        check_reply("generate", 0, mine, meta, arrays, stats)
        for i, c in enumerate(commands_fn(ref)):
            cj = dict(c)
            if "plateDensity" in cj:
                cj["plateDensity"] = {str(kk): v for kk, v in cj["plateDensity"].items()}
            r_ref = post(cj)
            r_mine = w.onmessage(dict(c))
            arrays, meta = flatten(r_ref)
            check_reply(c["cmd"], i + 1, r_mine, meta, arrays, stats)
    finally:
        w.close()


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = Library(build())
    bad = lastbit = 0
    totals = dict(float_elements=0, float_differing=0, int_elements=0, worst=0.0)
    for k in range(rounds):
        rng = np.random.default_rng(first + k)
        sliders = lambda: {s: float(rng.choice([0.0, 1.0, np.round(rng.random(), 2)], p=[0.15, 0.15, 0.7])) for s in SLIDER_KEYS}   # noqa: E731
        gen = dict(cmd="generate", N=int(rng.integers(*NRANGE)), P=int(rng.choice([2, 3, 5, 8, 12, 20, 40, 80])), jitter=float(rng.choice([0.0, 0.5, 0.75, 1.0])),
                   nMag=float(np.round(rng.random() * 0.8, 2)), numContinents=int(rng.integers(1, 9)), continentSizeVariety=float(rng.choice([0, 0.5, 1.0])),
                   temperatureOffset=float(rng.choice([0, -3, 4])), precipitationOffset=float(rng.choice([0, -0.3, 0.3])),
                   landCoverage=float(rng.choice([0.05, 0.15, 0.3, 0.5, 0.85])), seed=int(rng.integers(0, 16777216)), **sliders())
        if rng.random() < 0.3:
            gen["toggledIndices"] = [int(rng.integers(0, gen["P"]))]
        imported = os.environ.get("FUZZ_IMPORT") == "1" or rng.random() < 0.15
        if imported:          # importHeightmap (js/planet-worker.js:771-942) of a random smooth image with black oceans
            iw = int(rng.choice([32, 64, 100]))
            ih = iw // 2
            yy, xx = np.mgrid[0:ih, 0:iw]
            img = 120 + 80 * np.sin(xx * rng.uniform(0.1, 0.5) + rng.uniform(0, 6)) * np.cos(yy * rng.uniform(0.1, 0.6)) + 30 * rng.random((ih, iw))
            img[np.sin(xx * rng.uniform(0.05, 0.3) + yy * rng.uniform(0.05, 0.3)) > rng.uniform(-0.3, 0.5)] = 0
            gen = dict(cmd="importHeightmap", N=gen["N"], jitter=gen["jitter"], grayscale=np.clip(np.round(img), 0, 255).astype(np.uint8).ravel(),
                       imageWidth=iw, imageHeight=ih, seed=gen["seed"], temperatureOffset=gen["temperatureOffset"],
                       precipitationOffset=gen["precipitationOffset"], **{k: gen[k] for k in SLIDER_KEYS})
        t0 = time.time()

        def commands_fn(ref, case=first + k, gen=gen):
            r2 = np.random.default_rng(7 + case)          # the follow-up commands are a function of the case number: both replays get the same
            sl = lambda: {s: float(r2.choice([0.0, 1.0, np.round(r2.random(), 2)], p=[0.15, 0.15, 0.7])) for s in SLIDER_KEYS}   # noqa: E731
            seeds = [int(s) for s in ref["plateSeeds"]]
            ocean = {int(s) for s in ref["plateIsOcean"]} ^ {seeds[int(r2.integers(0, len(seeds)))]}
            dens = {int(kk): float(v) for kk, v in ref["plateDensity"].items()}
            dens[seeds[0]] = float(np.round(2.4 + r2.random(), 3))
            if ref.get("type") == "done" and gen["cmd"] == "importHeightmap":
                return [dict(cmd="reapply", skipClimate=bool(r2.random() < 0.5), **sl()), dict(cmd="computeClimate", temperatureOffset=1.5, precipitationOffset=0.1)]
            return [dict(cmd="reapply", skipClimate=bool(r2.random() < 0.5), **sl()),
                    dict(cmd="editRecompute", plateIsOcean=sorted(ocean), plateDensity=dens, nMag=float(np.round(r2.random() * 0.6, 2)), **sl()),
                    dict(cmd="computeClimate", temperatureOffset=1.5, precipitationOffset=0.1)]
        stats = dict(float_elements=0, float_differing=0, int_elements=0, worst=0.0)
        try:
            replay(gen, commands_fn, lib, os.environ.get("FUZZ_MATH") == "detmath", stats)
            verdict = "ok"
        except AssertionError as e:
            # same planet with pb_detmath inside the evaluator: a difference that disappears is the elementary functions' last bit
            # (the reference derives a rift angle from rounding noise at the cell a hotspot dome is centred on)
            s2 = dict(float_elements=0, float_differing=0, int_elements=0, worst=0.0)
            try:
                replay(gen, commands_fn, lib, True, s2)
                verdict = f"LAST-BIT (libm vs pb_detmath; 0 of {s2['float_elements']} floats differ with the same Math): {str(e)[:160]}"
                lastbit += 1
            except AssertionError as e2:
                verdict = f"MISMATCH (also with the same Math): {str(e2)[:300]}"
                bad += 1
        except Exception as e:       # an evaluator gap or a worker error: report, do not hide
            verdict = f"ERROR: {type(e).__name__}: {str(e)[:300]}"
            bad += 1
        for kk in ("float_elements", "float_differing", "int_elements"):
            totals[kk] += stats[kk]
        totals["worst"] = max(totals["worst"], stats["worst"])
        print(f"round {first + k}: {gen['cmd']} N={gen['N']} P={gen.get('P')} cont={gen.get('numContinents')} var={gen.get('continentSizeVariety')} land={gen.get('landCoverage')} "
              f"jitter={gen['jitter']} seed={gen['seed']} toggled={gen.get('toggledIndices')} floats={stats['float_elements']} "
              f"differing={stats['float_differing']} ints={stats['int_elements']} [{time.time() - t0:.0f} s] -> {verdict}", flush=True)
    print(f"rounds {rounds}, failures {bad}, last-bit-only {lastbit}, totals (libm runs) {totals}")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
