"""N>1 path on CPU: two gloo ranks smooth a field over a cell-range-sharded mesh with one-cell halo exchange
and must reproduce the single-mesh oracle bit for bit.  The per-rank compute runs on the host emulation of the
kernels (test infrastructure); on a GPU box the same code runs on NCCL with the CUDA library (-m gpu)."""
import os
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_cells, passes, backend, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import binding as oracle
    from planet_heightmap_generation_b200._lib import Library
    from planet_heightmap_generation_b200.engine import DeviceMesh
    from planet_heightmap_generation_b200.sharded import HaloExchanger, Shard, smoothFieldSharded
    from tests.conftest import make_planet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if backend == "gloo":
        from tests.emul.build_emul import build
        lib, device = Library(build()), None
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        from planet_heightmap_generation_b200 import build as b
        lib, device = Library(b.build()), torch.device("cuda", rank % torch.cuda.device_count())
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    mesh, xyz, nd, elev = make_planet(oracle, n_cells)
    sh = Shard(mesh, xyz, world, rank)
    dm = DeviceMesh(sh.mesh, sh.r_xyz, device=0 if device is None else device.index, lib=lib)
    field = torch.from_numpy(sh.scatter(elev))
    if device is not None:
        field = field.to(device)
    ex = HaloExchanger(sh, device)
    smoothFieldSharded(dm, ex, field, passes)
    if device is not None:
        torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"own_{rank}.npy"), sh.owned(field).cpu().numpy())
    np.save(os.path.join(out_dir, f"halo_{rank}.npy"), np.array([sh.halo.size, len(sh.send), len(sh.recv)]))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, n_cells, passes, backend, tmp_path):
    import torch.multiprocessing as mp
    from oracle import binding as oracle
    from tests.conftest import make_planet
    mp.spawn(_worker, args=(world, _free_port(), n_cells, passes, backend, str(tmp_path)), nprocs=world, join=True)
    mesh, xyz, nd, elev = make_planet(oracle, n_cells)
    want = elev.copy()
    oracle.smooth_field(mesh, want, passes)
    got = np.concatenate([np.load(tmp_path / f"own_{r}.npy") for r in range(world)])
    assert got.shape == want.shape
    assert (got.view(np.uint32) == want.view(np.uint32)).all(), "sharded smoothField differs from the single-mesh oracle"
    for r in range(world):
        halo, nsend, nrecv = np.load(tmp_path / f"halo_{r}.npy")
        assert 0 < halo < 0.2 * n_cells and nsend >= 1 and nrecv >= 1


@pytest.mark.parametrize("world,passes", [(2, 5), (3, 4)])
def test_sharded_smooth_field_gloo(world, passes, tmp_path):
    _run(world, 6000, passes, "gloo", tmp_path)


def test_shard_plan_properties():
    """Every cell is owned exactly once; halos are exactly the out-of-range neighbours; send/recv lists pair up."""
    from oracle import binding as oracle
    from planet_heightmap_generation_b200.sharded import Shard
    from tests.conftest import make_planet
    mesh, xyz, nd, elev = make_planet(oracle, 6000)
    world = 4
    shards = [Shard(mesh, xyz, world, r) for r in range(world)]
    assert sum(s.nOwn for s in shards) == mesh.numRegions
    for s in shards:
        for p, (a, b) in s.recv.items():
            want = s.halo[a - s.nOwn:b - s.nOwn]
            peer = shards[p]
            assert s.rank in peer.send
            assert (peer.send[s.rank] + peer.lo == want).all()
        # the pole vertex (last id) touches the lowest ids: the last shard exchanges with shard 0
    assert 0 in shards[-1].recv and (world - 1) in shards[0].recv


@pytest.mark.gpu
def test_sharded_smooth_field_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, 20000, 6, "nccl", tmp_path)
