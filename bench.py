#!/usr/bin/env python
"""bench.py — Voronoi cells/s of the per-cell hot path (BASELINE.json: full pipeline elevation→erosion→climate).

A "step" is one pass over a 1 000 001-cell Fibonacci-sphere Voronoi mesh (seed 42, sliders at the
reference's UI defaults, P = 40 plates in 10 super-plates):

  --workload full (default, BASELINE configs[2]) = elevation + post + climate, i.e. handleGenerate
                     (js/planet-worker.js:136-339) from assignElevation on
  --workload post    (configs[1])  reapply-style pass (js/planet-worker.js:341-358): clone the pre-erosion
                     elevation, runPostProcessing = warp → smooth → erodeComposite (50 stream-power iterations,
                     5 glacial, 1 thermal, two priority floods) → ridge sharpening → soil creep → erosionDelta
  --workload climate computeWind → computeOceanCurrents → computePrecipitation → computeTemperature →
                     classifyKoppen on the eroded elevation (js/planet-worker.js:229-268)
  --workload elevation   assignElevation alone (js/elevation.js:216-1391)

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--cells C] [--workload W]

own arm       value  = cells/s with the planet resident in HBM (device-pointer mode of the C ABI),
                       CUDA-event timed, max over ranks; N>1 = N independent planets (replicas, weak).
              e2e    = the same pass through the host-pointer C ABI (pinned host buffers in, every result
                       array of the reference's reply message back in host memory), wall clock.
              roofline / cpu_baseline as the task contract asks.
reference arm the CPU oracle (oracle/, single thread like the reference's single Web Worker) on the
              same workload.  The reference itself is JavaScript and no JS runtime exists in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SLIDERS = dict(smoothing=0.10, glacialErosion=0.50, hydraulicErosion=0.50, thermalErosion=0.10,
               ridgeSharpening=0.50, terrainWarp=0.75)
SEED = 42
METRIC = "voronoi_cells_per_sec"
UNIT = "cells/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------
# synthetic planet (cached on local disk: both arms and every rank use the same mesh)
# ---------------------------------------------------------------------------------------------------
def get_planet(cells: int, seed: int = SEED, on_device: int | None = None):
    """Mesh + r_xyz of the seeded planet.  Own arm (`on_device` = GPU index): triangulated on that GPU by the
    product path (DeviceMesh.from_points).  Reference arm: the CPU checker oracle/mesh_hull.py, cached on disk."""
    from planet_heightmap_generation_b200.mesh import SphereMesh
    if on_device is not None:
        from planet_heightmap_generation_b200.engine import DeviceMesh
        t = time.time()
        dm = DeviceMesh.build_sphere(cells, 0.75, seed, device=on_device)
        mesh, xyz = SphereMesh.from_csr(dm.adjOffset, dm.adjList), dm.r_xyz.copy()
        dm.close()
        log(f"[bench] {cells}-cell sphere generated and triangulated on the GPU in {time.time() - t:.2f}s (first call, includes context creation)")
        return mesh, xyz
    from oracle.mesh_hull import build_sphere_from_points
    cache_dir = os.path.join(tempfile.gettempdir(), "planet_b200_cache")
    path = os.path.join(cache_dir, f"mesh_{cells}_{seed}.npz")
    if os.path.exists(path):
        try:
            z = np.load(path)
            return SphereMesh.from_csr(z["adjOffset"], z["adjList"]), z["r_xyz"]
        except Exception:
            pass
    t = time.time()
    from oracle import binding as oracle
    oracle.build()
    mesh, xyz = build_sphere_from_points(oracle.fibonacci_sphere(cells, 0.75, seed))
    log(f"[bench] built {cells}-cell mesh in {time.time() - t:.1f}s")
    try:
        os.makedirs(cache_dir, exist_ok=True)
        fd, tmp = tempfile.mkstemp(dir=cache_dir, suffix=".npz")
        os.close(fd)
        np.savez(tmp, adjOffset=mesh.adjOffset, adjList=mesh.adjList, r_xyz=xyz)
        os.replace(tmp, path)
    except Exception as e:   # cache is best effort
        log(f"[bench] mesh cache not written: {e}")
    return mesh, xyz


NMAG, SPREAD = 0.40, 5       # Roughness slider default (index.html) and the worker's fixed spread (planet-worker.js:138)


# index.html slider defaults: 80 plates, 4 continents, size variety 0, land coverage 0.30
NUM_PLATES, NUM_CONTINENTS, SIZE_VARIETY, LAND_COVERAGE, N_COARSE = 80, 4, 0.0, 0.30, 20000


class Inputs:
    """The seeded planet both arms work on: (N, jitter 0.75, seed) → mesh + r_xyz.  Everything downstream — coarse
    plates, r_plate, super plates, elevation … — is computed by the arm itself from the seed and the slider defaults."""

    def __init__(self, cells: int, seed: int = SEED, on_device: int | None = None):
        self.mesh, self.xyz = get_planet(cells, SEED, on_device)


def workload_name(cells, hiters, workload):
    elev = "assignElevation (nMag 0.40, spread 5, super plates on)"
    post = (f"runPostProcessing with default sliders, hIters={hiters} K=0.0003 m=0.5 tIters=1 gIters=5, smooth 1, "
            f"ridge 3, creep 3")
    clim = "computeWind+computeOceanCurrents+computePrecipitation+computeTemperature+classifyKoppen (default offsets)"
    plat = (f"generateCoarsePlates ({NUM_PLATES} plates, {NUM_CONTINENTS} continents, land {LAND_COVERAGE}, {N_COARSE}-region coarse mesh) + "
            f"projectCoarsePlates + smoothAndReconnectPlates(3) + buildSuperPlates")
    tri = "buildSphere (Fibonacci points with seeded jitter + pole, spherical Delaunay adjacency in the SphereMesh constructor's order)"
    what = {"post": post, "climate": clim, "elevation": elev, "mesh": tri, "plates": plat,
            "elevation+post": " then ".join([elev, post]),
            "full": " then ".join([tri, plat, elev, post, clim])}[workload]
    return f"{cells + 1}-cell Fibonacci sphere (jitter 0.75, seed {SEED}), {what}"


def bench_config(args, world):
    """`config` of the JSON line — the same dict in both arms (it depends on the command line only)."""
    return {"workload": workload_name(args.cells, args.hiters, args.workload), "cells_per_planet": args.cells + 1,
            "seed": SEED, "hiters": args.hiters,
            "multi_gpu": "single" if world == 1 else MULTI_GPU_TEXT[args.multi_gpu],
            "l2": "256 MiB buffer written between steps (inside the timed region)",
            "flood": args.flood or "host",
            "inputs": "only (N, jitter, seed) and the slider defaults: points, mesh adjacency, coarse plates, r_plate and "
                      "super plates are rebuilt inside every step",
            "cpu_arm": "oracle/ (C++ -O2 restatement of the reference's JS, one thread like its single Web Worker); it starts from "
                       "the finished mesh: mesh construction is in the GPU arm's step only (its CPU checker is qhull, not the "
                       "reference's Delaunator)"}


MULTI_GPU_TEXT = {
    "replicas": "replicas (one copy of the planet per GPU, no data-path collective)",
    "shards": "cell-range shards, one-cell halo: every rank holds the planet, the Jacobi / propagation sweep loops of the climate "
              "stack run on contiguous cell-id ranges with a peer-memory halo exchange per sweep; class S/R stages replicated",
}


CLIMATE_REPLY_F32 = (  # the per-cell arrays of the worker's climateDone reply (js/planet-worker.js:635-653)
    [f"r_wind_{d}_{s}" for s in ("summer", "winter") for d in ("east", "north")]
    + [f"r_ocean_{k}_{s}" for s in ("summer", "winter") for k in ("current_east", "current_north", "speed", "warmth")]
    + ["r_precip_summer", "r_precip_winter", "r_temperature_summer", "r_temperature_winter"])


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  Default: NVML from a background thread of this process (a few
    cheap queries every 250 ms); BENCH_SAMPLER=smi uses an `nvidia-smi -lms` child process instead (the form the profiling recipe
    shows — on some boxes its queries were seen to stall the launching process for hundreds of milliseconds), BENCH_SAMPLER=none
    switches sampling off."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.mode = os.environ.get("BENCH_SAMPLER", "nvml")
        self.proc = None
        self.path = None
        self.thread = None
        self.samples = []
        self.stop_flag = False

    def _nvml_loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((sm, mx, [k for k, b in bits.items() if r & b]))
                time.sleep(0.25)
        except Exception as e:      # sampling is evidence, not the measurement
            self.samples.append(("error", str(e)))

    def start(self):
        if self.mode == "none":
            return
        if self.mode == "nvml":
            import threading
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=fd,
                                         stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "sampler": self.mode}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            good = [x for x in self.samples if x and x[0] != "error"]
            if good:
                reasons = sorted({r for x in good for r in x[2]})
                out.update(sm_mhz=float(np.median([x[0] for x in good])), sm_max_mhz=good[0][1], reasons=reasons, samples=len(good))
            else:
                out["error"] = [x[1] for x in self.samples if x and x[0] == "error"][:1]
            return out
        if not self.proc:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------
# CPU legs (the oracle is the checker / baseline; never the product path)
# ---------------------------------------------------------------------------------------------------
def oracle_step_seconds(inp, hiters, steps, warmup, workload, budget_s=None, ready=None, keep=None):
    """Times the oracle on the same workload.  Returns (per-step seconds, per-stage seconds of the last step).
    `budget_s` bounds the CPU work: once it is used up the loop stops after the next timed pass (warm-up passes that no
    longer fit are skipped), so a large --steps cannot turn the reference arm into a run of many minutes.
    `keep` (a dict) receives the result arrays of the last pass — the checker side of the bench's parity block."""
    from oracle import binding as oracle
    oracle.build()
    mesh, xyz = inp.mesh, inp.xyz
    nd = oracle.neighbor_dist(mesh, xyz)
    oe = oracle.Elevation(mesh, xyz)
    clim = oracle.Climate(mesh, xyz)

    pstate = {}

    def plates():
        from planet_heightmap_generation_b200.sphere import park_miller
        cp = oracle.generate_coarse_plates(SEED, NUM_PLATES, NUM_CONTINENTS, SIZE_VARIETY, LAND_COVERAGE, N_COARSE)
        seeds, pio = cp["coarsePlateSeeds"], cp["coarsePlateIsOcean"]
        rp = oracle.project_coarse_plates(mesh, xyz, cp["coarseMesh"], cp["coarse_xyz"], cp["coarse_r_plate"], SEED, NUM_PLATES)
        oracle.smooth_and_reconnect_plates(mesh, rp, seeds, 3)
        table = {}
        for s in seeds:                                       # plate densities: js/planet-worker.js:196-201
            d = park_miller(s + 777, 2)
            table[s] = dict(isOcean=s in pio, pole=tuple(cp["coarsePlateVec"][s]["pole"]), omega=cp["coarsePlateVec"][s]["omega"],
                            density=float(3.0 + d[0] * 0.5) if s in pio else float(2.4 + d[1] * 0.5))
        rs, sp = oracle.build_super_plates(mesh, rp, table)
        pstate.update(r_plate=rp, r_super=rs, super=sp, seeds=seeds, pio=pio, table=table)

    def elevation():
        oe.assign(pstate["r_plate"], pstate["table"], pstate["seeds"], SEED, NMAG, SEED, SPREAD, pstate["r_super"], pstate["super"])
        return oe.get("r_elevation"), oe.get("hotspot")

    plates()

    if workload == "mesh":
        from oracle.mesh_hull import build_sphere_from_points
        times = []
        for i in range(warmup + steps):
            t = time.perf_counter()
            build_sphere_from_points(oracle.fibonacci_sphere(mesh.numRegions - 1, 0.75, SEED))
            if i >= warmup:
                times.append(time.perf_counter() - t)
        return times, {"mesh_s": times[-1]}
    pre = hot = eroded = None
    if workload in ("post", "climate"):
        pre, hot = elevation()
        pre, hot = pre.copy(), hot.copy()
    if workload == "climate":
        eroded = pre.copy()
        oracle.run_post_processing(mesh, xyz, eroded, SLIDERS, nd, SEED, hot, hiters)
    if ready is not None:
        ready()                      # preparation done: several of these run side by side and start their passes together
    times, stages = [], {}
    began, over = time.perf_counter(), False
    for i in range(warmup + steps):
        t = time.perf_counter()
        e, h = (pre.copy(), hot) if pre is not None else (None, None)
        if workload in ("full", "plates"):
            plates()
        t0 = time.perf_counter()
        if workload in ("full", "elevation", "elevation+post"):
            e, h = elevation()
        t1 = time.perf_counter()
        if keep is not None and e is not None:
            keep["prePostElev"] = e.copy()
        if workload in ("full", "post", "elevation+post"):
            o_delta, o_ocean = oracle.run_post_processing(mesh, xyz, e, SLIDERS, nd, SEED, h, hiters)
            if keep is not None:
                keep.update(erosionDelta=o_delta, r_isOcean=o_ocean)
        t2 = time.perf_counter()
        if workload in ("full", "climate"):
            o_koppen = clim.run_all(eroded if workload == "climate" else e, pstate["pio"], pstate["r_plate"], SEED)
            if keep is not None:
                keep["r_koppen"] = o_koppen
                for k in CLIMATE_REPLY_F32:
                    keep[k] = clim.get(k)
        t3 = time.perf_counter()
        if keep is not None:
            keep.update(r_plate=pstate["r_plate"].copy(), r_superPlate=np.asarray(pstate["r_super"]).copy())
            if e is not None:
                keep["r_elevation"] = e if workload != "climate" else eroded
        if i >= warmup or over:
            times.append(t3 - t)
        stages = {"plates_s": t0 - t, "elevation_s": t1 - t0, "post_s": t2 - t1, "climate_s": t3 - t2}
        if budget_s is not None and (time.perf_counter() - began) + (t3 - t) > budget_s:
            if times:
                break
            over = True            # budget gone during warm-up: the next pass is the one timed pass
    return times, stages


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inp = Inputs(args.cells)
    n = inp.mesh.numRegions
    times, stages = oracle_step_seconds(inp, args.hiters, args.steps, min(args.warmup, 1), args.workload, budget_s=120.0)
    total = float(np.sum(times))
    v = n * len(times) / total
    sample = (f"{args.workload} workload, {len(times)} timed passes of {n} cells (of {args.steps} requested; the arm stops after "
              f"≈120 s of CPU work), oracle/ C++ -O2, 1 thread")
    # the counterpart of the own arm's `throughput_in_flight`: the same number of planets at once, one host thread each
    in_flight = None
    if args.workload == "full" and args.in_flight > 1:
        import threading
        b = min(args.in_flight, os.cpu_count() or 1)
        errs, mark = [], {}
        gate = threading.Barrier(b, action=lambda: mark.setdefault("t0", time.perf_counter()))

        def one():
            try:
                oracle_step_seconds(inp, args.hiters, 1, 0, args.workload, ready=gate.wait)
            except Exception as e:
                errs.append(e)
                gate.abort()

        threads = [threading.Thread(target=one) for _ in range(b)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        secs = time.perf_counter() - mark.get("t0", time.perf_counter())
        if not errs:
            in_flight = {"planets_in_flight": b, "cores": b, "value": n * b / secs, "unit": UNIT, "seconds": secs,
                         "note": "one oracle pipeline per host thread, one pass each"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1000 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
        "note": "reference is browser JavaScript (one Web Worker, single thread); no JS runtime in this image, "
                "so this arm times oracle/ — the C++ -O2 restatement of the same functions — on one host core",
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "cpu": cpu_model(), "host_cores": os.cpu_count(), "stages_last_step": stages},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "throughput_in_flight": in_flight,
    }), flush=True)


# ---------------------------------------------------------------------------------------------------
# algorithmic bytes per launch of the kernels that can dominate (DESIGN.md §5 gives the derivation)
# ---------------------------------------------------------------------------------------------------
def algorithmic_bytes(name: str, N: int, E: int, land: int):
    """Compulsory bytes of one launch (every distinct array a pass must touch counted once,
    neighbour gathers assumed to hit cache).  Keyed by kernel name; None = not tabulated."""
    csr = 4 * (N + 1) + 4 * E
    lf = land / max(N, 1)
    table = {
        # order 4 + target 4 + cellDist 4 + flow 4 + elev r/w 8 + isOcean 1 per land cell
        "pb::SolveK": 25 * land,
        # order 4 + pos 4 + target 4 + contrib r/w 8 per land cell
        "pb::AccumulateK": 20 * land,
        # CSR + elev 4 + isOcean 1 + ndist (land rows) + drainTarget 4 + cellDist 4
        "pb::ReceiversK": csr + 5 * N + int(4 * E * lf) + 8 * land,
        # per flooded land cell: CSR row (4 + 4·6) + visited 6·1 + elev 6·4 + surface r/w 8 + drainTo 4 + heap entry r/w 16
        "pb::k_flood_heap": 86 * land,
        # per filled cell: surface 4 + elev r/w 8 + depth 4 + path reads; tabulated per land cell as in SURVEY §8d (30·L)
        "pb::k_carve_lift": 30 * land,
        # sorted f64 coordinates 24 + key 4 + id 4 + degree/offset 4 + row written 4·6 (candidates assumed cached)
        "pb::StarK": 60 * N,
        "pb::SmoothFieldK": csr + 8 * N,
        "pb::SmoothMaskedK": csr + 9 * N,
        "pb::DiffuseWarmthK": csr + 12 * N,
        # offsets + isLand + src + dst for every cell; adj + weight for land rows
        "pb::ShadowSweepK": 4 * (N + 1) + 9 * N + int(8 * E * lf),
        # the same sweep over compacted land rows: counted with the row-indexed form's bytes so that the two are comparable
        "pb::ShadowLandK": 4 * (N + 1) + 9 * N + int(8 * E * lf),
        # CSR + xyz 12 + wind3d 12 + windE/N 8 + height 4 + src 4 + dst 4 + mask 1
        "pb::AdvectK": csr + 45 * N,
        "cub::DeviceRadixSort::SortPairs": 4 * 16 * land,
    }
    return table.get(name)


# ---------------------------------------------------------------------------------------------------
# several planets in flight on one GPU (one context + stream + host thread per planet)
# ---------------------------------------------------------------------------------------------------
def pipelined_throughput(args, mesh, xyz, local, planets: int, steps: int):
    """The step is latency-bound (one dependent chain of heap pops on one SM, host-serial fills), so a GPU that
    generates many planets keeps several in flight: every planet has its own pb_context, CUDA stream and host thread
    and runs the same full pipeline as the single-planet step.  Returns seconds for `steps` steps of every planet
    (host wall clock around a device-wide synchronize on both sides)."""
    import threading

    import torch
    from planet_heightmap_generation_b200 import climate as cl
    from planet_heightmap_generation_b200 import plates as pl
    from planet_heightmap_generation_b200.elevation import assignElevation
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200.terrain_post import runPostProcessing
    dev = torch.device("cuda", local)
    N, E = mesh.numRegions, int(mesh.adjList.shape[0])

    class Job:
        def __init__(self):
            self.dm = DeviceMesh(mesh, xyz, device=local)
            if args.flood:
                self.dm.set_option("flood", args.flood)
            self.stream = torch.cuda.Stream(device=dev)
            i32, f32, u8 = torch.int32, torch.float32, torch.uint8
            self.xyz = torch.empty(3 * N, dtype=f32, device=dev)
            self.off, self.adj = torch.empty(N + 1, dtype=i32, device=dev), torch.empty(E, dtype=i32, device=dev)
            self.r_plate, self.r_super = torch.empty(N, dtype=i32, device=dev), torch.empty(N, dtype=i32, device=dev)
            self.delta = torch.empty(N, dtype=f32, device=dev)
            self.ocean, self.koppen = torch.empty(N, dtype=u8, device=dev), torch.empty(N, dtype=u8, device=dev)
            self.error = None

        def step(self):
            dm = self.dm
            dm.generateFibonacciSphere(args.cells, 0.75, SEED, out=self.xyz)
            dm.triangulateSphere(self.xyz, self.off, self.adj)
            cp = pl.generateCoarsePlates(dm, SEED, NUM_PLATES, NUM_CONTINENTS, SIZE_VARIETY, LAND_COVERAGE, N_COARSE)
            seeds, vec, pio, dens = cp["coarsePlateSeeds"], cp["coarsePlateVec"], cp["coarsePlateIsOcean"], cp["plateDensity"]
            pl.projectCoarsePlates(dm, None, cp["coarseMesh"], cp["coarse_xyz"], cp["coarse_r_plate"], SEED, NUM_PLATES, out=self.r_plate)
            cp["coarseMesh"].close()
            pl.smoothAndReconnectPlates(dm, self.r_plate, seeds, 3)
            sp = pl.buildSuperPlates(dm, self.r_plate, seeds, vec, pio, dens, out=self.r_super)
            res = assignElevation(dm, None, pio, self.r_plate, vec, seeds, SEED, NMAG, SEED, SPREAD, dens, sp)
            elev = res["r_elevation"]
            runPostProcessing(dm, None, elev, SLIDERS, None, SEED, res["debugLayers"]["hotspot"], hItersOverride=args.hiters,
                              out_erosionDelta=self.delta, out_isOcean=self.ocean, timing=False)
            cl.computeClimate(dm, elev, pio, self.r_plate, SEED, 0.0, 0.0, 0.3, out_koppen=self.koppen)
            self.elev = elev

        def run(self, k):
            try:
                torch.cuda.set_device(local)
                with torch.cuda.stream(self.stream):
                    for _ in range(k):
                        self.step()
                    self.stream.synchronize()
            except Exception as e:      # surfaced by the caller
                self.error = e

    jobs = [Job() for _ in range(planets)]

    def run_all(k):
        threads = [threading.Thread(target=j.run, args=(k,)) for j in jobs]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for j in jobs:
            if j.error:
                raise j.error

    run_all(1)                                   # warm-up: allocations, caches
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_all(steps)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    first = jobs[0]
    agree = all(bool((j.elev == first.elev).all().item()) and bool((j.koppen == first.koppen).all().item()) for j in jobs[1:])
    ref = (first.elev.clone(), first.koppen.clone())
    for j in jobs:
        j.dm.close()
    return secs, agree, ref


# ---------------------------------------------------------------------------------------------------
# cell-range sharded sweep loops: strong scaling of ONE planet's climate stack across the ranks
# ---------------------------------------------------------------------------------------------------
def build_eroded_planet(dm, cells, hiters):
    """(elev, pio, r_plate) of the seeded planet on `dm`: plates → assignElevation (→ runPostProcessing when hiters >= 0), all on
    the device"""
    import torch
    from planet_heightmap_generation_b200 import plates as pl
    from planet_heightmap_generation_b200.elevation import assignElevation
    from planet_heightmap_generation_b200.terrain_post import runPostProcessing
    dev = torch.device("cuda", torch.cuda.current_device())
    N = cells + 1
    r_plate = torch.empty(N, dtype=torch.int32, device=dev)
    r_super = torch.empty(N, dtype=torch.int32, device=dev)
    cp = pl.generateCoarsePlates(dm, SEED, NUM_PLATES, NUM_CONTINENTS, SIZE_VARIETY, LAND_COVERAGE, N_COARSE)
    seeds, vec, pio, dens = cp["coarsePlateSeeds"], cp["coarsePlateVec"], cp["coarsePlateIsOcean"], cp["plateDensity"]
    pl.projectCoarsePlates(dm, None, cp["coarseMesh"], cp["coarse_xyz"], cp["coarse_r_plate"], SEED, NUM_PLATES, out=r_plate)
    cp["coarseMesh"].close()
    pl.smoothAndReconnectPlates(dm, r_plate, seeds, 3)
    sp = pl.buildSuperPlates(dm, r_plate, seeds, vec, pio, dens, out=r_super)
    res = assignElevation(dm, None, pio, r_plate, vec, seeds, SEED, NMAG, SEED, SPREAD, dens, sp)
    elev = res["r_elevation"]
    if hiters >= 0:
        runPostProcessing(dm, None, elev, SLIDERS, None, SEED, res["debugLayers"]["hotspot"], hItersOverride=hiters, timing=False)
    return elev, pio, r_plate


def sharded_climate_probe(args, rank, world, local, cells, steps=2):
    """Supplementary line of an N-GPU run in replicas mode: the climate stack of ONE `cells`-cell planet (BASELINE config 4
    size by default) with its sweep loops sharded by cell-id range over the N ranks, against the same stack unsharded on one
    GPU, results compared bit for bit.  Every rank holds the planet (replicated), so no rank is idle in either run."""
    import torch
    import torch.distributed as dist
    from planet_heightmap_generation_b200 import climate as cl
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200.sharded import SweepShardGroup
    dev = torch.device("cuda", local)
    t_setup = time.perf_counter()
    dm = DeviceMesh.build_sphere(cells, 0.75, SEED, device=local)
    N = cells + 1
    group = SweepShardGroup(dm, rank, world, min_cells=0)
    group.set_min_cells(1 << 62)                       # the planet itself is built unsharded
    elev, pio, r_plate = build_eroded_planet(dm, cells, -1)    # climate of the pre-erosion elevation: the probe times sweeps, not erosion
    koppen = torch.empty(N, dtype=torch.uint8, device=dev)
    setup_s = time.perf_counter() - t_setup

    def run(k):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            out = cl.computeClimate(dm, elev, pio, r_plate, SEED, 0.0, 0.0, 0.3, out_koppen=koppen)
        e1.record()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / k, out

    run(1)
    ms_one, out = run(steps)
    st = cl._state(dm)
    keys = ("r_precip_summer", "r_precip_winter", "r_temperature_summer", "r_temperature_winter", "r_ocean_warmth_summer")
    def field(k):
        v = st.field(k)
        return v.clone() if isinstance(v, torch.Tensor) else torch.from_numpy(np.array(v, copy=True))

    ref = {k: field(k) for k in keys}
    ref_koppen = koppen.clone()
    group.set_min_cells(0)
    before = group.info()
    run(1)
    ms_sh, out = run(steps)
    info = group.info()
    same = bool((koppen == ref_koppen).all().item()) and all(bool((field(k) == ref[k]).all().item()) for k in keys)
    flag = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    group.close()
    dm.close()
    sweeps = (info["sweeps_sharded"] - before["sweeps_sharded"]) // (steps + 1)
    return {"what": f"climate stack of one {N}-cell planet (820 → {sweeps} graph sweeps per pass at this size), sweep loops sharded by "
                    f"contiguous cell-id range over {world} GPUs with a one-cell halo pushed into peer memory by the sweep kernels "
                    f"(CUDA IPC over NVLink), ranges all-gathered at the end of every loop; everything else replicated",
            "cells": N, "n_gpus": world, "ms_per_pass_unsharded_1gpu": ms_one, "ms_per_pass_sharded": ms_sh,
            "speedup": ms_one / ms_sh, "cells_per_s_sharded": N / (ms_sh / 1e3), "sweeps_per_pass": int(sweeps),
            "loops_per_pass": int((info["loops_sharded"] - before["loops_sharded"]) // (steps + 1)),
            "halo_bytes_per_sweep_rank0": info["halo_bytes_per_sweep"], "adjacent_ranks_rank0": info["adjacent_ranks"],
            "rows_rank0": info["hi"] - info["lo"], "bit_identical_to_unsharded_on_every_rank": bool(flag.item()),
            "setup_s": setup_s,
            "limiter": "the replicated part of the stack (pointwise kernels, 5 BFS, 6 radix sorts for the percentiles, ITCZ bins) "
                       "does not shrink with N; the sharded sweeps are bounded by the per-sweep flag round trip over NVLink"}



# ---------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from planet_heightmap_generation_b200 import climate as cl
    from planet_heightmap_generation_b200 import plates as pl
    from planet_heightmap_generation_b200.elevation import DEBUG_LAYERS, assignElevation
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200.terrain_post import runPostProcessing

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = args.workload
    do_elev, do_post, do_clim = wl in ("full", "elevation", "elevation+post"), wl in ("full", "post", "elevation+post"), wl in ("full", "climate")
    do_mesh = wl in ("full", "mesh")
    do_plates = wl in ("full", "plates")
    # replicas: every rank processes its own copy of the same seeded planet (identical work per GPU, so the
    # N-GPU numbers are a clean weak-scaling series)
    inp = Inputs(args.cells, SEED, on_device=local)
    mesh, xyz = inp.mesh, inp.xyz
    N, E = mesh.numRegions, int(mesh.adjList.shape[0])
    dm = DeviceMesh(mesh, xyz, device=local)
    shards_mode = world > 1 and args.multi_gpu == "shards"
    group = None
    if shards_mode:
        from planet_heightmap_generation_b200.sharded import SweepShardGroup
        group = SweepShardGroup(dm, rank, world, min_cells=None if args.shard_min_cells < 0 else args.shard_min_cells)
    planets = 1 if shards_mode else world          # strong scaling: every rank works on the same planet
    xyz_t = torch.empty(3 * N, dtype=torch.float32, device=dev)
    off_t = torch.empty(N + 1, dtype=torch.int32, device=dev)
    adj_t = torch.empty(E, dtype=torch.int32, device=dev)
    if args.flood:
        dm.set_option("flood", args.flood)

    r_plate = torch.empty(N, dtype=torch.int32, device=dev)
    r_super = torch.empty(N, dtype=torch.int32, device=dev)
    delta = torch.empty(N, dtype=torch.float32, device=dev)
    ocean = torch.empty(N, dtype=torch.uint8, device=dev)
    koppen = torch.empty(N, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    state = {}

    def plate_stage(rp, rs):
        """generateCoarsePlates → projectCoarsePlates → smoothAndReconnectPlates → buildSuperPlates (planet-worker.js:160-211)"""
        cp = pl.generateCoarsePlates(dm, SEED, NUM_PLATES, NUM_CONTINENTS, SIZE_VARIETY, LAND_COVERAGE, N_COARSE)
        seeds, vec, pio, dens = cp["coarsePlateSeeds"], cp["coarsePlateVec"], cp["coarsePlateIsOcean"], cp["plateDensity"]
        pl.projectCoarsePlates(dm, None, cp["coarseMesh"], cp["coarse_xyz"], cp["coarse_r_plate"], SEED, NUM_PLATES, out=rp)
        cp["coarseMesh"].close()
        pl.smoothAndReconnectPlates(dm, rp, seeds, 3)
        sp = pl.buildSuperPlates(dm, rp, seeds, vec, pio, dens, out=rs)
        return dict(seeds=seeds, vec=vec, pio=pio, dens=dens, super=sp)

    def plates_device():
        state.update(plate_stage(r_plate, r_super))

    def elevation_device():
        res = assignElevation(dm, None, state["pio"], r_plate, state["vec"], state["seeds"], SEED, NMAG, SEED, SPREAD, state["dens"],
                              state["super"])
        state["elev"], state["hotspot"] = res["r_elevation"], res["debugLayers"]["hotspot"]

    def post_device():
        runPostProcessing(dm, None, state["elev"], SLIDERS, None, SEED, state["hotspot"], hItersOverride=args.hiters,
                          out_erosionDelta=delta, out_isOcean=ocean, timing=False)

    # untimed preparation of the inputs the chosen workload starts from
    plates_device()
    elevation_device()
    pre = state["elev"].clone()
    land = int((pre > 0).sum().item())
    if wl in ("mesh", "plates"):
        state["elev"] = pre
    if wl == "climate":
        post_device()

    def step_device():
        flush.zero_()
        if do_mesh:
            dm.generateFibonacciSphere(args.cells, 0.75, SEED, out=xyz_t)
            dm.triangulateSphere(xyz_t, off_t, adj_t)
        if do_plates:
            plates_device()
        if do_elev:
            elevation_device()
        elif do_post:
            state["elev"] = pre.clone()   # handleReapply clones W.prePostElev before post-processing (planet-worker.js:353)
        if do_post:
            post_device()
        if do_clim:
            cl.computeClimate(dm, state["elev"], state["pio"], r_plate, SEED, 0.0, 0.0, 0.3, out_koppen=koppen)
        if step_sync:
            torch.cuda.synchronize()      # diagnostic (BENCH_STEP_SYNC=1): no run-ahead of the next step's host stages

    step_sync = os.environ.get("BENCH_STEP_SYNC", "0") != "0"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler starts before the last warm-up steps: the first nvidia-smi queries of a fresh process take tens of
    # milliseconds of driver time each and must not fall into the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup - 1, 0)):
        step_device()
    # last warm-up step is fully profiled: it names the dominant kernel (largest total device time)
    dm.profile_start(None)
    step_device()
    rows = sorted(dm.profile_stop(), key=lambda p: -p["ms"])
    barrier()
    dominant = args.dominant or (rows[0]["name"] if rows else "")
    sweep_kernel = "pb::SmoothFieldK" if do_clim else "pb::SolveK"
    breakdown = None
    if rank == 0:
        tot = sum(p["ms"] for p in rows)
        breakdown = [{"name": p["name"], "launches": p["launches"], "ms": round(p["ms"], 3)} for p in rows[:14]]
        log(f"[bench] per-kernel device time of one warm-up step (events around every launch, sum {tot:.1f} ms):")
        for p in rows[:30]:
            log(f"   {p['ms']:10.3f} ms  {p['launches']:6d}x  {p['name']}")

    launches0 = dm.launch_count()
    # events around the launches of the dominant kernel and of the most-launched sweep kernel only
    dm.profile_start(dominant if dominant == sweep_kernel else "|".join([dominant, sweep_kernel]))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    import gc
    gc.collect()
    gc.disable()        # a generation-2 collection of the interpreter (torch + numpy + scipy are loaded) costs 0.1-0.3 s and would land
                        # in a random timed step; it is re-enabled after the timed regions
    ev0.record()
    marks = []
    for _ in range(args.steps):
        step_device()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append(e)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    ms_steps = [round(a.elapsed_time(b), 1) for a, b in zip([ev0] + marks[:-1], marks)]
    prof = dm.profile_stop()
    launches = dm.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = N * planets * args.steps / (ms_total / 1000.0)

    # ---- end to end through the host-pointer C ABI ------------------------------------------------
    elev_dev_final = state["elev"].clone()
    mesh_same = True
    if do_mesh:   # the adjacency rebuilt inside the timed steps is the one the mesh was created from
        mesh_same = bool((xyz_t.cpu() == torch.from_numpy(xyz)).all().item()) and \
            bool((off_t.cpu() == torch.from_numpy(mesh.adjOffset)).all().item()) and \
            bool((adj_t.cpu() == torch.from_numpy(mesh.adjList)).all().item())
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_xyz = torch.empty(3 * N, dtype=torch.float32).pin_memory()
    h_off = torch.empty(N + 1, dtype=torch.int32).pin_memory()
    h_adj = torch.empty(E, dtype=torch.int32).pin_memory()
    np_xyz, np_off, np_adj = h_xyz.numpy(), h_off.numpy(), h_adj.numpy()
    r_plate_dev_final, r_super_dev_final = r_plate.clone(), r_super.clone()
    h_plate, h_super, h_pre, h_hot0 = pin(r_plate.cpu().numpy()), pin(r_super.cpu().numpy()), pin(pre.cpu().numpy()), pin(state["hotspot"].cpu().numpy())
    h_elev = torch.empty(N, dtype=torch.float32).pin_memory()
    h_delta = torch.empty(N, dtype=torch.float32).pin_memory()
    h_ocean = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_koppen = torch.empty(N, dtype=torch.uint8).pin_memory()
    np_delta, np_ocean, np_koppen, np_plate, np_super = h_delta.numpy(), h_ocean.numpy(), h_koppen.numpy(), h_plate.numpy(), h_super.numpy()
    if wl == "climate":
        h_elev.copy_(elev_dev_final.cpu())
    # mesh: r_xyz goes host → device for the triangulation; plates: r_plate in/out between the three calls
    h2d = (12 * N if do_mesh else 0) + (4 * N + 4 * N if do_plates else 0) + (8 * N if do_elev else 0) + \
          ((8 * N if not do_elev else 0) if do_post else 0) + (8 * N if do_clim else 0)
    d2h = (12 * N + 4 * (N + 1) + 4 * E if do_mesh else 0) + (4 * N + 4 * N + 4 * N if do_plates else 0) + ((4 + 4 + 3 + 4 * len(DEBUG_LAYERS)) * N if do_elev else 0) + (9 * N if do_post else 0) + \
          ((4 * len(CLIMATE_REPLY_F32) + 1) * N + 3 * 4 * 360 if do_clim else 0)
    reply, host = {}, dict(state, super=dict(state["super"], r_superPlate=np_super))

    host_stage = {}

    def step_host(collect=False):
        flush.zero_()
        np_elev, np_hot = h_elev.numpy(), h_hot0.numpy()
        tt = [time.perf_counter()]
        if do_mesh:
            dm.generateFibonacciSphere(args.cells, 0.75, SEED, out=np_xyz)
            dm.triangulateSphere(np_xyz, np_off, np_adj)
        tt.append(time.perf_counter())
        if do_plates:
            host.update(plate_stage(np_plate, np_super))
        tt.append(time.perf_counter())
        if do_elev:
            res = assignElevation(dm, None, host["pio"], np_plate, host["vec"], host["seeds"], SEED, NMAG, SEED, SPREAD, host["dens"],
                                  host["super"])
            np_elev, np_hot = res["r_elevation"], res["debugLayers"]["hotspot"]
        elif do_post:
            h_elev.copy_(h_pre)
        tt.append(time.perf_counter())
        if collect and (do_elev or do_post):
            host["prePostElev"] = np.array(np_elev, copy=True)
        if do_post:
            runPostProcessing(dm, None, np_elev, SLIDERS, None, SEED, np_hot, hItersOverride=args.hiters,
                              out_erosionDelta=np_delta, out_isOcean=np_ocean, timing=False)
        tt.append(time.perf_counter())
        if do_clim:
            w, o, p, t, _ = cl.computeClimate(dm, np_elev, host["pio"], np_plate, SEED, 0.0, 0.0, 0.3, out_koppen=np_koppen)
            for res, keys in ((w, CLIMATE_REPLY_F32[:4] + ["itczLons", "itczLatsSummer", "itczLatsWinter"]),
                              (o, CLIMATE_REPLY_F32[4:12]), (p, CLIMATE_REPLY_F32[12:14]), (t, CLIMATE_REPLY_F32[14:])):
                for k in keys:
                    reply[k] = res[k]     # device → host copy of every array of the climateDone message
        tt.append(time.perf_counter())
        host["elev"] = np_elev
        host_stage.update(mesh_ms=1e3 * (tt[1] - tt[0]), plates_ms=1e3 * (tt[2] - tt[1]), elevation_ms=1e3 * (tt[3] - tt[2]),
                          post_ms=1e3 * (tt[4] - tt[3]), climate_ms=1e3 * (tt[5] - tt[4]))

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    gc.enable()
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = N * planets * args.steps / e2e_s
    # host-pointer and device-pointer passes agree bit for bit
    if do_plates:
        mesh_same = mesh_same and bool((h_plate == r_plate_dev_final.cpu()).all().item()) and bool((h_super == r_super_dev_final.cpu()).all().item())
    same = (wl in ("mesh", "plates") or bool((torch.from_numpy(np.ascontiguousarray(host["elev"])) == elev_dev_final.cpu()).all().item())) and \
        (not do_clim or bool((h_koppen == koppen.cpu()).all().item())) and mesh_same and \
        (not do_mesh or (bool((h_xyz == torch.from_numpy(xyz)).all().item()) and
                         bool((h_off == torch.from_numpy(mesh.adjOffset)).all().item()) and bool((h_adj == torch.from_numpy(mesh.adjList)).all().item())))

    # ---- roofline (events recorded inside the timed region) ---------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy), of measured" if "hbm_gbs" in peaks else "fallback 6650 GB/s, of fallback"

    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (1M cells); only
    # quoted when the bench runs the size the capture was taken at
    ncu_traffic = {"pb::k_carve_lift": (42.2e6, "profiles/r02_kernels_ncu.md (same seeded planet: 41.25 MB read + 0.97 MB written)"),
                   "pb::SolveK": (35.7e6, "profiles/r02_kernels_ncu.md (34.08 MB read + 1.65 MB written)"),
                   "pb::SmoothFieldK": (32.0e6, "profiles/r01_sweeps_ncu.md"),
                   "pb::StarK": (133.2e6, "profiles/r01_mesh_plates_ncu.md")}

    def roof(kernel):
        dom = [p for p in prof if p["name"] == kernel]
        if not dom:
            return None
        d = dom[0]
        ab = algorithmic_bytes(d["name"], N, E, land)
        avg_ms = d["ms"] / d["launches"]
        out = {"bound": "hbm", "kernel": d["name"], "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
               "traffic": None, "algorithmic_bytes_per_launch": ab, "avg_launch_ms": avg_ms,
               "launches_timed": d["launches"], "share_of_step": d["ms"] / ms_total, "peak_source": peak_src}
        if ab:
            out["achieved"] = ab / (avg_ms * 1e-3) / 1e9
            out["frac"] = out["achieved"] / peak
        if d["name"] in ncu_traffic and args.cells == 1_000_000:
            out["traffic"], out["traffic_source"] = ncu_traffic[d["name"]]
        return out

    roofline = roof(dominant)
    roofline_sweep = roof(sweep_kernel) if sweep_kernel != dominant else None

    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        keep = {}
        times, stages = oracle_step_seconds(inp, args.hiters, 1, 0, wl, keep=keep)
        cpu_baseline = {"value": N / times[0], "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"one full pass of the same {N}-cell workload ({times[0]:.1f} s), oracle/ C++ -O2, 1 thread"
                                  + ("; mesh construction is not in the CPU figure (its checker is qhull, not the reference's "
                                     "Delaunator: `--workload mesh` times it separately)" if wl == "full" else
                                     " (qhull convex hull + SphereMesh constructor, scipy)" if wl == "mesh" else ""),
                        "cpu": cpu_model(), "host_cores": os.cpu_count(), "stages": stages}
        # ---- parity on the benchmark's own planet: every array of the timed arm's reply against the oracle's, bit for bit ----
        step_host(collect=True)
        got = {"r_plate": np_plate if do_plates else r_plate_dev_final.cpu().numpy(),
               "r_superPlate": np_super if do_plates else r_super_dev_final.cpu().numpy()}
        if do_elev or do_post:
            got["prePostElev"] = host.get("prePostElev")
        if do_elev or do_post or do_clim:
            got["r_elevation"] = host["elev"]
        if do_post:
            got.update(erosionDelta=np_delta, r_isOcean=np_ocean)
        if do_clim:
            got["r_koppen"] = np_koppen
            got.update({k: reply[k] for k in CLIMATE_REPLY_F32})
        fields = {}
        for k, want in keep.items():
            if k not in got or got[k] is None:
                continue
            a, b_ = np.ascontiguousarray(got[k]), np.ascontiguousarray(want)
            if a.shape != b_.shape or a.dtype != b_.dtype:
                fields[k] = int(max(a.size, b_.size))
                continue
            fields[k] = int(((a.view(np.uint32) != b_.view(np.uint32)) if a.dtype == np.float32 else (a != b_)).sum())
        parity = {"vs": "oracle", "cells": N, "fields": fields, "mismatches": int(sum(fields.values())),
                  "compare": "bit-exact (f32 compared as uint32), arrays of one extra untimed pass through the host-pointer C ABI "
                             "against one pass of oracle/ on the same (N, seed, sliders)"}
        log(f"[bench] parity vs oracle at {N} cells: {parity['mismatches']} mismatching cells over {len(fields)} arrays")

    # ---- the main line is complete here; the supplementary legs below run under a watchdog ----------------------------
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "ms_steps": ms_steps, "higher_is_better": True, "scaling": "strong" if shards_mode else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "impl": "b200", "config": bench_config(args, world), "land_cells": land, "parity": parity,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": d2h * world, "ms_per_step": 1000 * e2e_s / args.steps,
                "matches_device_path": same, "stages_last_step_ms": {k: round(v, 2) for k, v in host_stage.items()}},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_sweep": roofline_sweep,
        "cpu_baseline": cpu_baseline, "sharded_sweeps": None, "throughput_in_flight": None, "kernel_breakdown": breakdown, "library": dm.lib.version,
    }
    import threading

    def give_up():
        # a supplementary leg did not come back (e.g. a peer rank died): the measured line still goes out, every rank leaves
        if rank == 0:
            line["supplementary"] = f"a supplementary leg did not finish within {args.extras_timeout} s and was abandoned"
            print(json.dumps(line), flush=True)
        os._exit(0)

    watchdog = threading.Timer(args.extras_timeout, give_up)
    watchdog.daemon = True
    watchdog.start()

    # ---- several planets in flight (supplementary: `value` above is one planet at a time) ----------------------
    in_flight = None
    if wl == "full" and args.in_flight > 1 and world == 1:
        secs, agree, ref = pipelined_throughput(args, mesh, xyz, local, args.in_flight, args.steps)
        if world > 1:
            t = torch.tensor([secs], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        in_flight = {"planets_in_flight_per_gpu": args.in_flight, "value": N * world * args.in_flight * args.steps / secs, "unit": UNIT,
                     "seconds": secs, "steps_per_planet": args.steps,
                     "matches_single_planet_path": bool(agree and (ref[0] == elev_dev_final).all().item() and (ref[1] == koppen).all().item()),
                     "note": "one pb_context + CUDA stream + host thread per planet, same full pipeline per planet; host wall clock "
                             "around device synchronizes"}

    line["throughput_in_flight"] = in_flight

    # ---- the data path that shards: sweep loops by cell-id range (pb_shardsweep.h) -------------------------------
    sharding = None
    if shards_mode:
        info = group.info()
        # the same climate pass unsharded on this GPU must give the same bits
        ident = None
        if do_clim:
            ref_k = koppen.clone()
            group.set_min_cells(1 << 62)
            cl.computeClimate(dm, state["elev"], state["pio"], r_plate, SEED, 0.0, 0.0, 0.3, out_koppen=koppen)
            torch.cuda.synchronize()
            ok = torch.tensor([1 if bool((koppen == ref_k).all().item()) else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            ident = bool(ok.item())
        sharding = {"rows_rank0": info["hi"] - info["lo"], "adjacent_ranks_rank0": info["adjacent_ranks"],
                    "halo_bytes_per_sweep_rank0": info["halo_bytes_per_sweep"], "sharded_sweeps_total_rank0": info["sweeps_sharded"],
                    "sharded_loops_total_rank0": info["loops_sharded"], "loops_sharded": info["active"], "min_cells": info["min_cells"],
                    "koppen_bit_identical_to_unsharded_on_every_rank": ident,
                    "limiter": "Amdahl: the class S/R stages (host-serial heap flood and randomized fills, dependency-latency-bound "
                               "solve / accumulate, stable sort) and the pointwise / BFS / sort kernels run replicated on every rank; "
                               "only the Jacobi / propagation sweep loops shrink with N"}
        dist.barrier()
        group.close()
    elif world > 1 and args.shard_probe_cells > 0 and wl == "full":
        sharding = sharded_climate_probe(args, rank, world, local, args.shard_probe_cells)

    line["sharded_sweeps"] = sharding
    watchdog.cancel()

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and parity["mismatches"]:
        raise SystemExit(f"bench.py: the timed arm's results differ from the oracle: {parity['fields']}")


# ---------------------------------------------------------------------------------------------------
# sharded sweeps: the one data path that has a real exchange step (cell-range shards + one-cell halo per sweep)
# ---------------------------------------------------------------------------------------------------
def run_sharded_sweeps(args):
    """`--workload sharded-sweeps`: smoothField (js/climate-util.js:5-25) on a planet whose sweep loops are sharded by
    contiguous cell-id ranges across the ranks (csrc/pb_shardsweep.h: halo values stored into peer memory by the sweep
    kernel, all-gather at the end of the loop).  Strong scaling of one planet: value = cells × sweeps per second."""
    import torch
    import torch.distributed as dist
    from planet_heightmap_generation_b200.climate_util import smoothField
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200.sharded import SweepShardGroup
    from planet_heightmap_generation_b200.sphere import synthetic_elevation
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    dm = DeviceMesh.build_sphere(args.cells, 0.75, SEED, device=local)
    N = args.cells + 1
    f0 = torch.from_numpy(synthetic_elevation(dm.r_xyz, SEED, 0.3)).to(dev)
    group = SweepShardGroup(dm, rank, world, min_cells=0) if world > 1 else None
    f = f0.clone()
    sweeps = args.sweeps

    def step():
        f.copy_(f0)
        smoothField(dm, f, sweeps)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    info = group.info() if group else None
    if rank == 0:
        per_sweep_us = 1000 * ms / (args.steps * sweeps)
        print(json.dumps({
            "metric": "cell_sweeps_per_sec_sharded_smoothField", "value": N * sweeps * args.steps / (ms / 1000), "unit": "cell-sweeps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "b200",
            "config": {"workload": f"{N}-cell sphere, {sweeps} smoothField sweeps per step, cell-range shards, one-cell halo pushed into "
                                   f"peer memory by the sweep kernel (CUDA IPC over NVLink), all-gather at the end of the loop",
                       "us_per_sweep": per_sweep_us, "shards": info},
            "algorithmic_GBps": 36.0 * N / (per_sweep_us * 1e-6) / 1e9}), flush=True)
    if group:
        barrier()
        group.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    # exactly one line on stdout: libraries (NCCL's version banner, …) that write to fd 1 are sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="full", choices=["full", "post", "climate", "elevation", "elevation+post", "mesh", "plates", "sharded-sweeps"])
    ap.add_argument("--sweeps", type=int, default=100, help="sweeps per step of --workload sharded-sweeps")
    ap.add_argument("--flood", default="", choices=["", "device", "host"],
                    help="engine option: where the serial heap pass of priorityFloodCarve runs (default host)")
    ap.add_argument("--multi-gpu", default="auto", choices=["auto", "replicas", "shards"], dest="multi_gpu",
                    help="N>1: independent planets per GPU (weak) or ONE planet whose sweep loops are sharded by cell-id range "
                         "(strong); auto = shards from 4M cells up (below that a sweep is shorter than the halo flag round trip)")
    ap.add_argument("--shard-min-cells", type=int, default=-1, dest="shard_min_cells",
                    help="engine threshold below which sweep loops stay unsharded (-1: engine default, 0: always shard)")
    ap.add_argument("--extras-timeout", type=float, default=240.0, dest="extras_timeout",
                    help="seconds the supplementary legs (planets in flight, sharded probe) may take before the main line is printed without them")
    ap.add_argument("--shard-probe-cells", type=int, default=10_000_000, dest="shard_probe_cells",
                    help="N>1, replicas mode: size of the supplementary sharded climate measurement (0: skip)")
    ap.add_argument("--in-flight", type=int, default=4, dest="in_flight",
                    help="planets kept in flight per GPU for the supplementary throughput figure of the full workload (0/1: skip)")
    ap.add_argument("--cells", type=int, default=1_000_000)
    ap.add_argument("--hiters", type=int, default=50)
    ap.add_argument("--dominant", default="", help="kernel whose launches are event-timed for the roofline "
                                                   "(default: the kernel with the largest device time)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.multi_gpu == "auto":
        args.multi_gpu = "shards" if args.cells >= 4_000_000 else "replicas"
    if args.warmup < 3 and args.impl == "b200":
        log("[bench] note: fewer than 3 warm-up steps requested")
    if args.workload == "sharded-sweeps":
        run_sharded_sweeps(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
