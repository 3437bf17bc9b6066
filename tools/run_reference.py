"""Posts worker commands to the UNMODIFIED reference worker (/root/reference/js/planet-worker.js) running under the evaluator
tests/golden/minijs.py and writes every array of every reply to an .npz file (build container only).

  python tools/run_reference.py out.npz '{"cmd": "generate", "N": 600, "P": 12, "jitter": 0.75, "nMag": 0.4, "numContinents": 3,
      "seed": 42, "smoothing": 0.1, "hydraulicErosion": 0.5, "thermalErosion": 0.1, "ridgeSharpening": 0.5, "glacialErosion": 0.5,
      "terrainWarp": 0.75}' '{"cmd": "reapply", "smoothing": 0.3, …}'

Arrays are stored as "<reply index>/<key>" (debug layers as "<i>/debugLayers.<name>"), everything else as JSON under "__meta__" —
the layout of tests/golden/reference_*.npz."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.golden.make_reference_vectors import flatten, make_interpreter  # noqa: E402

if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    post = make_interpreter()
    out, metas, commands = {}, [], []
    for i, text in enumerate(sys.argv[2:]):
        cmd = json.loads(text)
        t0 = time.time()
        reply = post(cmd)
        arrays, meta = flatten(reply)
        print(f"[{i}] {cmd['cmd']} → {reply['type']} ({len(arrays)} arrays, {time.time() - t0:.1f} s)")
        out.update({f"{i}/{k}": v for k, v in arrays.items()})
        metas.append(meta)
        commands.append(cmd)
    out["__meta__"] = np.frombuffer(json.dumps({"commands": commands, "replies": metas}, default=lambda o: o.tolist()).encode(), np.uint8)
    np.savez_compressed(sys.argv[1], **out)
    print("wrote", sys.argv[1])
