"""Runs the worker flows (generate → reapply → editRecompute → computeClimate → exportMap, importHeightmap; both mesh orders,
400 … 20 000 cells) on an AddressSanitizer + UBSan build of the host-emulation library: the engine's host orchestration (staging,
host-serial stages, C ABI argument handling) is the same code in the CUDA build.  Build and run:

  g++ -x c++ -std=c++17 -O1 -g -fPIC -ffp-contract=off -DPB_EMUL -fsanitize=address,undefined -fno-omit-frame-pointer -shared \
      -o /tmp/libpb_hostemu_asan.so planet_heightmap_generation_b200/csrc/planet_b200.cu
  LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
      python tools/asan_emulation.py
(libstdc++ is preloaded as well so that ASan's __cxa_throw interceptor finds the real function when the library raises a C++ exception.)

Last run (end of round 2): clean; so are the four flows of tests/test_napi_addon.py with napi_host.cc + the addon compiled with the
same flags against that library (error paths included)."""
import os
import sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planet_heightmap_generation_b200._lib import Library
from planet_heightmap_generation_b200.worker import PlanetWorker
from planet_heightmap_generation_b200 import planet_mesh as pm
lib = Library('/tmp/libpb_hostemu_asan.so')
S = dict(smoothing=0.1, glacialErosion=0.5, hydraulicErosion=0.5, thermalErosion=0.1, ridgeSharpening=0.5, terrainWarp=0.75)
for order in ("canonical", "delaunator"):
    for N, P in ((400, 6), (3000, 12), (20000, 80)):
        w = PlanetWorker(lib=lib, mesh_order=order)
        r = w.onmessage(dict(cmd="generate", N=N, P=P, jitter=0.75, nMag=0.4, numContinents=3, seed=7 + N, toggledIndices=[1], **S))
        assert r["type"] == "done", r
        r2 = w.onmessage(dict(cmd="reapply", **{k: 0.3 for k in S}))
        assert r2["type"] == "reapplyDone", r2
        pio = sorted(set(r["plateIsOcean"]) ^ {r["plateSeeds"][0]})
        r3 = w.onmessage(dict(cmd="editRecompute", plateIsOcean=pio, plateDensity=r["plateDensity"], nMag=0.3, **S))
        assert r3["type"] == "editDone", r3
        r4 = w.onmessage(dict(cmd="computeClimate", temperatureOffset=1.0))
        assert r4["type"] == "climateDone", r4
        name, png = w.exportMap(r3, "biome", 256)
        print(order, N, "ok", len(png), flush=True)
        w.close()
img = (np.random.default_rng(1).random(64 * 32) * 255).astype(np.uint8)
w = PlanetWorker(lib=lib)
r = w.onmessage(dict(cmd="importHeightmap", N=2000, jitter=0.75, grayscale=img, imageWidth=64, imageHeight=32, seed=3, **S))
assert r["type"] == "done", r
print("import ok")
