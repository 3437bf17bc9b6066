"""Four planets in flight (one PlanetWorker per Python thread, two of them the same planet) plus the worker threads inside every
assignElevation call (five randomized fills, per-plate stress propagation) on a ThreadSanitizer build of the host-emulation
library — the host threading of the engine is the same code in the CUDA build.  Build and run:

  g++ -x c++ -std=c++17 -O1 -g -fPIC -ffp-contract=off -DPB_EMUL -fsanitize=thread -fno-omit-frame-pointer -shared \
      -o /tmp/libpb_hostemu_tsan.so planet_heightmap_generation_b200/csrc/planet_b200.cu
  LD_PRELOAD=$(gcc -print-file-name=libtsan.so) python tools/tsan_emulation.py

Last run (end of round 2): no ThreadSanitizer report."""
import os
import sys, threading, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planet_heightmap_generation_b200._lib import Library
from planet_heightmap_generation_b200.worker import PlanetWorker
lib = Library('/tmp/libpb_hostemu_tsan.so')
S = dict(smoothing=0.1, glacialErosion=0.5, hydraulicErosion=0.5, thermalErosion=0.1, ridgeSharpening=0.5, terrainWarp=0.75)
def one(seed, out):
    w = PlanetWorker(lib=lib)
    r = w.onmessage(dict(cmd="generate", N=6000, P=24, jitter=0.75, nMag=0.4, numContinents=3, seed=seed, **S))
    assert r["type"] == "done", r
    r2 = w.onmessage(dict(cmd="reapply", **{k: 0.3 for k in S}))
    out[seed] = (r["r_elevation"].copy(), r2["r_elevation"].copy())
    w.close()
res = {}
one(1, res)                       # single context: fills + propagation workers inside one call
ts = [threading.Thread(target=one, args=(s, res)) for s in (1, 2, 3, 1)]      # four planets in flight, two of them the same planet
for t in ts: t.start()
for t in ts: t.join()
print("done", sorted(res))
