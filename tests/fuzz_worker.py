"""Randomised end-to-end comparison of the engine (host emulation build of the kernel sources — test infrastructure) with
the oracle: the worker's generate → reapply → editRecompute → computeClimate sequence on random (N, seed, P, sliders).
Run on a box without a GPU:  python tests/fuzz_worker.py [rounds] [first_seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import binding as oracle  # noqa: E402
from planet_heightmap_generation_b200._lib import Library  # noqa: E402
from planet_heightmap_generation_b200.worker import PlanetWorker  # noqa: E402
from tests.emul.build_emul import build  # noqa: E402
from tests.test_worker import SLIDER_KEYS, OracleWorker, same  # noqa: E402


NRANGE = (int(os.environ.get("FUZZ_NMIN", 1500)), int(os.environ.get("FUZZ_NMAX", 12000)))


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    oracle.build()
    lib = Library(build())
    bad = 0
    for k in range(rounds):
        rng = np.random.default_rng(first + k)
        sl = {s: float(rng.choice([0.0, 1.0, np.round(rng.random(), 2)], p=[0.15, 0.15, 0.7])) for s in SLIDER_KEYS}
        msg = dict(cmd="generate", N=int(rng.integers(*NRANGE)), P=int(rng.choice([2, 3, 5, 8, 12, 20, 40, 80, 150])), jitter=float(rng.choice([0.0, 0.5, 0.75, 1.0])),
                   nMag=float(np.round(rng.random() * 0.8, 2)), numContinents=int(rng.integers(1, 12)), continentSizeVariety=float(rng.choice([0, 0.5, 1.0])),
                   temperatureOffset=float(rng.choice([0, -3, 4])), precipitationOffset=float(rng.choice([0, -0.3, 0.3])),
                   landCoverage=float(rng.choice([0.05, 0.15, 0.3, 0.5, 0.85])), seed=int(rng.integers(0, 16777216)), **sl)
        w, ow = PlanetWorker(lib=lib), OracleWorker(oracle)
        r = w.onmessage(dict(msg))
        ok = r["type"] == "done"
        if ok:
            elev, delta, koppen = ow.generate(msg)
            ok = same(r["r_plate"], ow.r_plate) and same(r["prePostElev"], ow.pre) and same(r["r_elevation"], elev) and \
                same(r["debugLayers"]["koppen"], koppen) and same(r["r_stress"], ow.oe.get("r_stress")) and \
                all(same(r[f], ow.clim.get(f)) for f in ("r_wind_east_summer", "r_ocean_warmth_winter", "r_precip_summer", "r_temperature_winter"))
        if ok:
            sl2 = {s: float(np.round(rng.random(), 2)) for s in SLIDER_KEYS}
            r2 = w.onmessage(dict(cmd="reapply", **sl2))
            e2, d2, k2 = ow.reapply(dict(sl2), msg["temperatureOffset"], msg["precipitationOffset"], msg["landCoverage"])
            ok = r2["type"] == "reapplyDone" and same(r2["r_elevation"], e2) and same(r2["erosionDelta"], d2) and same(r2["windDebugLayers"]["koppen"], k2)
        if ok:
            pio = set(ow.pio)
            flip = ow.seeds[int(rng.integers(0, len(ow.seeds)))]
            pio.symmetric_difference_update({flip})
            dens = dict(ow.dens)
            dens[flip] = float(2.4 + rng.random())
            sl3 = {s: float(np.round(rng.random(), 2)) for s in SLIDER_KEYS}
            m3 = dict(cmd="editRecompute", plateIsOcean=sorted(pio), plateDensity=dens, nMag=float(np.round(rng.random() * 0.6, 2)), **sl3)
            r3 = w.onmessage(m3)
            e3, d3, k3 = ow.edit(m3, msg["temperatureOffset"], msg["precipitationOffset"], msg["landCoverage"])
            ok = r3["type"] == "editDone" and same(r3["r_elevation"], e3) and same(r3["debugLayers"]["koppen"], k3) and same(r3["prePostElev"], ow.pre)
        if ok:
            r4 = w.onmessage(dict(cmd="computeClimate", temperatureOffset=1.5, precipitationOffset=0.1))
            k4 = ow.climate(ow.final, 1.5, 0.1, msg["landCoverage"], recompute_wind=False)
            ok = r4["type"] == "climateDone" and same(r4["climateDebugLayers"]["koppen"], k4)
        print(f"round {k}: N={msg['N']} P={msg['P']} jitter={msg['jitter']} seed={msg['seed']} sliders={sl} -> {'ok' if ok else 'MISMATCH ' + str(r.get('message', ''))}", flush=True)
        bad += not ok
        w.close()
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
