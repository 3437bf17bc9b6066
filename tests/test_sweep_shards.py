"""Cell-range sharding of the sweep loops (csrc/pb_shardsweep.h): every rank holds the whole planet, the Jacobi /
propagation loops run on contiguous cell-id ranges with a peer-memory halo exchange per sweep and an all-gather per loop.
The sharded climate stack must reproduce the oracle (= the unsharded pass) bit for bit on every rank.

CPU suite: the ranks are THREADS of one process over the host emulation of the kernels (the "IPC handles" are plain
pointers there), world 2 and 3.  -m gpu: one process per GPU (CUDA IPC over NVLink), skipped on a 1-GPU box."""
import os
import socket
import threading

import numpy as np
import pytest

from tests.conftest import assert_bit_equal, make_planet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = ["r_wind_east_summer", "r_ocean_warmth_winter", "r_ocean_current_east_summer", "r_precip_summer", "r_precip_winter",
         "r_rainshadow_summer", "r_temperature_summer", "r_temperature_winter", "r_continentality"]


class ThreadExchange:
    """all_gather between the rank threads of one process"""

    def __init__(self, world):
        self.world, self.slots, self.bar = world, [None] * world, threading.Barrier(world)

    def make(self, rank):
        def exchange(mine):
            self.slots[rank] = mine
            self.bar.wait()
            out = list(self.slots)
            self.bar.wait()
            return out
        return exchange


def _rank_body(lib, rank, world, exchange, mesh, xyz, elev, r_plate, pio, passes, out, device=0):
    from planet_heightmap_generation_b200 import climate as cl
    from planet_heightmap_generation_b200.climate_util import smoothField
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200.sharded import SweepShardGroup
    if exchange is None:             # one process per rank: torch.distributed carries the handles
        import torch.distributed as dist

        def exchange(mine):
            everyone = [None] * world
            dist.all_gather_object(everyone, mine)
            return everyone
    dm = DeviceMesh(mesh, xyz, device=device, lib=lib)
    grp = SweepShardGroup(dm, rank, world, exchange=exchange, min_cells=0)
    f = elev.copy()
    smoothField(dm, f, passes)
    smoothField(dm, f, 1)            # a second loop: buffer reuse + barrier epochs
    koppen = np.empty(mesh.numRegions, np.uint8)
    wind, ocean, precip, temp, _ = cl.computeClimate(dm, elev, pio, r_plate, 42, 0.0, 0.0, 0.3, out_koppen=koppen)
    st = cl._state(dm)
    res = {"smooth": f, "koppen": koppen, "info": grp.info()}
    for k in CHECK:
        res[k] = st.field(k)
    exchange(b"done")                # nobody frees its buffers while a peer may still store into them
    grp.close()
    out[rank] = res


def _oracle_side(oracle, n_cells, passes):
    from planet_heightmap_generation_b200.sphere import synthetic_plates
    mesh, xyz, nd, elev = make_planet(oracle, n_cells)
    r_plate, pio = synthetic_plates(xyz, elev, 42)
    want = elev.copy()
    oracle.smooth_field(mesh, want, passes)
    oracle.smooth_field(mesh, want, 1)
    oc = oracle.Climate(mesh, xyz)
    koppen = oc.run_all(elev, pio, r_plate, 42)
    return mesh, xyz, elev, r_plate, pio, want, oc, koppen


def _compare(results, world, want, oc, koppen, n):
    for rank in range(world):
        res = results[rank]
        assert res is not None, f"rank {rank} did not finish"
        assert_bit_equal(res["smooth"], want, f"rank {rank}: sharded smoothField")
        for k in CHECK:
            assert_bit_equal(res[k], oc.get(k), f"rank {rank}: {k}")
        assert_bit_equal(res["koppen"], koppen, f"rank {rank}: r_koppen")
        info = res["info"]
        assert info["active"] and info["hi"] - info["lo"] in (n // world, n // world + 1)
        assert info["sweeps_sharded"] > 100 and 0 < info["halo_bytes_per_sweep"] < 4 * n // 4
        assert info["adjacent_ranks"] >= 1


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_climate_threads_emulation(emu_lib, oracle, world):
    mesh, xyz, elev, r_plate, pio, want, oc, koppen = _oracle_side(oracle, 6000, 5)
    ex = ThreadExchange(world)
    out, errs = [None] * world, []

    def run(rank):
        try:
            _rank_body(emu_lib, rank, world, ex.make(rank), mesh, xyz, elev, r_plate, pio, 5, out)
        except Exception as e:       # pragma: no cover
            errs.append(e)
            ex.bar.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errs, errs
    _compare(out, world, want, oc, koppen, mesh.numRegions)


def test_shard_ranges_and_send_lists(emu_lib, oracle):
    """The pole vertex (last id) touches the lowest ids, so the last range is adjacent to the first; halo bytes are O(√N)."""
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200.sharded import SweepShardGroup
    mesh, xyz, nd, elev = make_planet(oracle, 6000)
    world = 4
    ex = ThreadExchange(world)
    infos = [None] * world

    def run(rank):
        dm = DeviceMesh(mesh, xyz, lib=emu_lib)
        g = SweepShardGroup(dm, rank, world, exchange=ex.make(rank), min_cells=0)
        infos[rank] = g.info()
        ex.make(rank)(b"done")
        g.close()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert all(i is not None for i in infos)
    assert infos[0]["lo"] == 0 and infos[-1]["hi"] == mesh.numRegions
    for a, b in zip(infos, infos[1:]):
        assert a["hi"] == b["lo"]
    assert infos[0]["adjacent_ranks"] == 2 and infos[-1]["adjacent_ranks"] == 2      # neighbours + the pole link
    assert infos[1]["adjacent_ranks"] == 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gpu_worker(rank, world, port, n_cells, passes, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import binding as oracle
    from planet_heightmap_generation_b200 import build as b
    from planet_heightmap_generation_b200._lib import Library
    from planet_heightmap_generation_b200.sphere import synthetic_plates
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # only carries the IPC handles
    lib = Library(b.build())
    mesh, xyz, nd, elev = make_planet(oracle, n_cells)
    r_plate, pio = synthetic_plates(xyz, elev, 42)
    out = [None] * world
    _rank_body(lib, rank, world, None, mesh, xyz, elev, r_plate, pio, passes, out, device=rank)
    res = out[rank]
    np.savez(os.path.join(out_dir, f"rank_{rank}.npz"), smooth=res["smooth"], koppen=res["koppen"],
             info=np.array([res["info"][k] for k in ("lo", "hi", "adjacent_ranks", "halo_bytes_per_sweep", "sweeps_sharded", "loops_sharded", "active", "min_cells")]),
             **{k: res[k] for k in CHECK})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_climate_peer_memory(world, oracle, tmp_path):
    """One process per GPU, CUDA IPC peer stores + flags over NVLink inside the sweep kernels."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    n_cells, passes = 40000, 7
    mesh, xyz, elev, r_plate, pio, want, oc, koppen = _oracle_side(oracle, n_cells, passes)
    mp.spawn(_gpu_worker, args=(world, _free_port(), n_cells, passes, str(tmp_path)), nprocs=world, join=True)
    results = []
    for r in range(world):
        z = np.load(tmp_path / f"rank_{r}.npz")
        res = {k: z[k] for k in CHECK}
        res.update(smooth=z["smooth"], koppen=z["koppen"])
        names = ("lo", "hi", "adjacent_ranks", "halo_bytes_per_sweep", "sweeps_sharded", "loops_sharded", "active", "min_cells")
        res["info"] = dict(zip(names, [int(v) for v in z["info"]]))
        results.append(res)
    _compare(results, world, want, oc, koppen, mesh.numRegions)
