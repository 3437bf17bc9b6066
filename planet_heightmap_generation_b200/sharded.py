"""Cell-range sharding of the mesh with one-cell halos.

The reference is single-threaded; there is no collective to mirror.  The mesh shards by contiguous
cell-id ranges (Fibonacci ids advance in z, so a range is a z-band whose neighbours lie within ±5·√N ids:
a shard exchanges halos with its two adjacent shards, plus the shard that owns the pole vertex N-1, which
touches the lowest ids).  Every rank holds

  * its owned cells as local ids [0, nOwn) and the halo cells it reads as local ids [nOwn, nLocal),
  * a local CSR whose owned rows keep the reference's neighbour ORDER (f64 sums are order-dependent) with
    column ids remapped to local ids; halo rows are empty,
  * per peer: the owned cells that peer reads (send list) and the slice of the halo it fills (recv slice).

A Jacobi sweep (smoothField, js/climate-util.js:5-25) is then: one local sweep over the owned rows + one
halo exchange.  The exchange uses torch.distributed point-to-point ops — NCCL over NVLink between GPUs,
gloo in the CPU test-suite — and carries 4 bytes per halo cell per sweep (≈ 3.5·√N cells per cut).
"""
from __future__ import annotations

import numpy as np

from .mesh import SphereMesh


class Shard:
    def __init__(self, mesh, r_xyz, world: int, rank: int):
        N = int(mesh.numRegions)
        off = np.asarray(mesh.adjOffset, np.int64)
        adj = np.asarray(mesh.adjList, np.int64)
        self.world, self.rank, self.N = world, rank, N
        bounds = [(N * k) // world for k in range(world + 1)]
        self.bounds = bounds
        lo, hi = bounds[rank], bounds[rank + 1]
        self.lo, self.hi, self.nOwn = lo, hi, hi - lo
        rows = adj[off[lo]:off[hi]]
        outside = rows[(rows < lo) | (rows >= hi)]
        self.halo = np.unique(outside)                                   # global ids, ascending
        self.nLocal = self.nOwn + self.halo.size
        # local CSR (owned rows in the reference's neighbour order, halo rows empty)
        local_cols = np.where((rows >= lo) & (rows < hi), rows - lo, self.nOwn + np.searchsorted(self.halo, rows))
        l_off = np.empty(self.nLocal + 1, np.int32)
        l_off[:self.nOwn + 1] = (off[lo:hi + 1] - off[lo]).astype(np.int32)
        l_off[self.nOwn + 1:] = l_off[self.nOwn]
        self.mesh = SphereMesh.from_csr(l_off, local_cols.astype(np.int32))
        xyz = np.asarray(r_xyz, np.float32).reshape(-1, 3)
        self.r_xyz = np.concatenate([xyz[lo:hi], xyz[self.halo]]).reshape(-1).copy()
        # recv: the halo ids owned by peer p form one contiguous slice (halo is sorted, ranges are contiguous)
        owner_cut = np.searchsorted(self.halo, bounds)
        self.recv = {p: (self.nOwn + int(owner_cut[p]), self.nOwn + int(owner_cut[p + 1]))
                     for p in range(world) if p != rank and owner_cut[p + 1] > owner_cut[p]}
        # send: owned cells that peer p reads = the symmetric question asked from p's side (the graph is
        # undirected: if p reads my cell c, then c has a neighbour in p's range)
        self.send = {}
        owner_of_col = np.searchsorted(bounds, rows, side="right") - 1
        row_of = np.repeat(np.arange(lo, hi), np.diff(off[lo:hi + 1]))
        for p in range(world):
            if p == rank:
                continue
            cells = np.unique(row_of[owner_of_col == p])
            if cells.size:
                self.send[p] = (cells - lo).astype(np.int64)              # local ids, ascending global order
        self.halo_bytes_per_sweep = 4 * int(self.halo.size)

    def scatter(self, global_field: np.ndarray) -> np.ndarray:
        """local array (owned + halo) of a global per-cell field"""
        return np.concatenate([global_field[self.lo:self.hi], global_field[self.halo]])

    def owned(self, local_field):
        return local_field[:self.nOwn]


class HaloExchanger:
    """Fills the halo part of a local field from the owning ranks (torch.distributed p2p)."""

    def __init__(self, shard: Shard, device=None):
        import torch
        self.shard = shard
        self.device = device
        self.send_idx = {p: torch.as_tensor(ix, device=device) for p, ix in shard.send.items()}
        self.send_buf, self.recv_buf = {}, {}

    def exchange(self, field):
        """field: 1-D torch tensor of nLocal elements (CPU for gloo, CUDA for NCCL); in place."""
        import torch
        import torch.distributed as dist
        sh = self.shard
        ops = []
        for p, ix in self.send_idx.items():
            buf = self.send_buf.get((p, field.dtype))
            if buf is None:
                buf = self.send_buf[(p, field.dtype)] = torch.empty(ix.numel(), dtype=field.dtype, device=field.device)
            torch.index_select(field, 0, ix, out=buf)
            ops.append(dist.P2POp(dist.isend, buf, p))
        for p, (a, b) in sh.recv.items():
            ops.append(dist.P2POp(dist.irecv, field[a:b], p))            # contiguous slice: received in place
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


def smoothFieldSharded(dm_local, exchanger: HaloExchanger, field, passes: int):
    """`passes` sweeps of smoothField over the sharded mesh; `field` is the local (owned + halo) torch tensor,
    updated in place.  dm_local = DeviceMesh(shard.mesh, shard.r_xyz).  Halo values must be current on entry."""
    from .climate_util import smoothField
    arg = field if field.is_cuda else field.numpy()
    for _ in range(int(passes)):
        smoothField(dm_local, arg, 1)
        exchanger.exchange(field)
    return field


class SweepShardGroup:
    """Cell-range sharding of the sweep loops of one planet across the ranks (csrc/pb_shardsweep.h).

    Every rank builds the SAME DeviceMesh (the whole planet) and attaches a group to it; afterwards smoothField and the
    sweep loops inside computeOceanCurrents / computePrecipitation / computeTemperature / computeWind compute only this rank's
    cell-id range, exchange the one-cell halo by peer-memory stores inside the sweep kernels and all-gather the ranges at the
    end of every loop.  Setup exchanges the CUDA IPC handles once (`exchange`: bytes → list of every rank's bytes; default
    torch.distributed.all_gather_object); the sweeps never touch the host.  All ranks must issue the same calls."""

    def __init__(self, dm, rank: int, world: int, exchange=None, min_cells=None):
        import ctypes as C
        self.dm, self.rank, self.world = dm, rank, world
        lib = dm.lib
        self._h = C.c_void_p()
        lib.check(lib.dll.pb_sweep_shards_create(dm._mesh, rank, world, C.byref(self._h)))
        if min_cells is not None:
            lib.check(lib.dll.pb_sweep_shards_set_min_cells(self._h, int(min_cells)))
        buf = (C.c_ubyte * 192)()
        lib.check(lib.dll.pb_sweep_shards_export(self._h, buf))
        if exchange is None:
            import torch.distributed as dist

            def exchange(mine):
                everyone = [None] * world
                dist.all_gather_object(everyone, mine)
                return everyone
        handles = exchange(bytes(buf))
        for p in range(world):
            if p == rank:
                continue
            hb = (C.c_ubyte * 192).from_buffer_copy(handles[p])
            lib.check(lib.dll.pb_sweep_shards_connect(self._h, p, hb))
        exchange(b"connected")          # nobody starts sweeping before every rank has mapped every buffer

    def set_min_cells(self, n: int):
        """planets below `n` cells keep their sweep loops unsharded (0: always shard; a huge value: never)"""
        self.dm.lib.check(self.dm.lib.dll.pb_sweep_shards_set_min_cells(self._h, int(n)))

    def info(self):
        import ctypes as C
        out = (C.c_int64 * 8)()
        self.dm.lib.check(self.dm.lib.dll.pb_sweep_shards_info(self._h, out))
        return dict(lo=int(out[0]), hi=int(out[1]), adjacent_ranks=int(out[2]), halo_bytes_per_sweep=int(out[3]),
                    sweeps_sharded=int(out[4]), loops_sharded=int(out[5]), active=bool(out[6]), min_cells=int(out[7]))

    def close(self):
        if getattr(self, "_h", None):
            self.dm.lib.dll.pb_sweep_shards_destroy(self._h)
            self._h = None
