"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the oracle on the
host cores) prints one JSON line with the keys the driver reads, the same `config` object the GPU arm would print for the same
flags, and the `e2e` / `cpu_baseline` objects the tier asks of that arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cells", "3000", "--steps", "2", "--warmup", "1",
                          "--in-flight", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "voronoi_cells_per_sec" and d["unit"] == "cells/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] >= 1 and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
    sys.path.insert(0, ROOT)
    import argparse

    import bench
    args = argparse.Namespace(cells=3000, hiters=50, workload="full", flood="", multi_gpu="auto", gpus=1)
    assert d["config"] == bench.bench_config(args, 1), "both arms must print the same config for the same flags"
    assert d["config"]["cells_per_planet"] == 3001 and "workload" in d["config"] and "model" not in d["config"]


def test_algorithmic_bytes_follow_the_survey_table():
    sys.path.insert(0, ROOT)
    import bench
    n, e, land = 1_000_001, 5_999_994, 274_655
    assert bench.algorithmic_bytes("pb::SmoothFieldK", n, e, land) == 4 * (n + 1) + 4 * e + 8 * n or \
        abs(bench.algorithmic_bytes("pb::SmoothFieldK", n, e, land) - 36 * n) < 100          # Jacobi sweep: 28 B CSR + 4 + 4 per cell
    assert bench.algorithmic_bytes("pb::k_carve_lift", n, e, land) == 30 * land               # carve / enforce 30·L
    assert bench.algorithmic_bytes("pb::no_such_kernel", n, e, land) in (None, 0)
