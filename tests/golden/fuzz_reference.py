"""Randomised comparison of the engine with THE REFERENCE ITSELF (build container only: needs /root/reference).

Every round draws (N, seed, plates, continents, variety, land coverage, jitter, offsets, sliders incl. the 0 / 1 extremes), posts
generate → reapply → editRecompute → computeClimate to the unmodified reference worker running under tests/golden/minijs.py and to
the engine's worker mirror (host emulation of the kernels, reference neighbour order), and compares every array of every reply
with the criteria of tests/test_zz_reference_vectors.py (integers identical; Float32 identical up to two last-bit flips per array
and 1e-4 relative).  Usage:  python tests/golden/fuzz_reference.py [rounds] [first_seed]
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
from tests.golden.make_reference_vectors import flatten, make_interpreter  # noqa: E402
from tests.test_zz_reference_vectors import check_reply  # noqa: E402

SLIDER_KEYS = ("smoothing", "glacialErosion", "hydraulicErosion", "thermalErosion", "ridgeSharpening", "terrainWarp")
NRANGE = (int(os.environ.get("FUZZ_NMIN", 300)), int(os.environ.get("FUZZ_NMAX", 2500)))


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = Library(build())
    bad = 0
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
        t0 = time.time()
        post = make_interpreter()
        w = PlanetWorker(lib=lib, mesh_order="delaunator")
        stats = dict(float_elements=0, float_differing=0, int_elements=0, worst=0.0)
        what = "generate"
        try:
            ref = post(dict(gen))
            mine = w.onmessage(dict(gen))
            arrays, meta = flatten(ref)
            check_reply("generate", 0, mine, meta, arrays, stats)
            seeds = [int(s) for s in ref["plateSeeds"]]
            commands = [dict(cmd="reapply", skipClimate=bool(rng.random() < 0.5), **sliders())]
            ocean = {int(s) for s in ref["plateIsOcean"]} ^ {seeds[int(rng.integers(0, len(seeds)))]}
            dens = {int(kk): float(v) for kk, v in ref["plateDensity"].items()}
            dens[seeds[0]] = float(np.round(2.4 + rng.random(), 3))
            commands.append(dict(cmd="editRecompute", plateIsOcean=sorted(ocean), plateDensity=dens, nMag=float(np.round(rng.random() * 0.6, 2)), **sliders()))
            commands.append(dict(cmd="computeClimate", temperatureOffset=1.5, precipitationOffset=0.1))
            for i, c in enumerate(commands):
                what = c["cmd"]
                cj = dict(c)
                if "plateDensity" in cj:
                    cj["plateDensity"] = {str(kk): v for kk, v in cj["plateDensity"].items()}
                r_ref = post(cj)
                r_mine = w.onmessage(dict(c))
                arrays, meta = flatten(r_ref)
                check_reply(what, i + 1, r_mine, meta, arrays, stats)
            verdict = "ok"
        except AssertionError as e:
            verdict = f"MISMATCH in {what}: {str(e)[:300]}"
            bad += 1
        except Exception as e:       # an evaluator gap or a worker error: report, do not hide
            verdict = f"ERROR in {what}: {type(e).__name__}: {str(e)[:300]}"
            bad += 1
        w.close()
        for kk in ("float_elements", "float_differing", "int_elements"):
            totals[kk] += stats[kk]
        totals["worst"] = max(totals["worst"], stats["worst"])
        print(f"round {first + k}: N={gen['N']} P={gen['P']} cont={gen['numContinents']} var={gen['continentSizeVariety']} land={gen['landCoverage']} "
              f"jitter={gen['jitter']} seed={gen['seed']} toggled={gen.get('toggledIndices')} floats={stats['float_elements']} "
              f"differing={stats['float_differing']} ints={stats['int_elements']} [{time.time() - t0:.0f} s] -> {verdict}", flush=True)
    print(f"rounds {rounds}, failures {bad}, totals {totals}")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
