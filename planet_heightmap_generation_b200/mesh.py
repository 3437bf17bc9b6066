"""Host-side mesh container (mirror of the SphereMesh class, js/sphere-mesh.js:94-146).

The triangulation itself is done on the device (`DeviceMesh.from_points`, csrc/pb_meshgen.h); this module only holds
the CSR container the hot-path functions take ({numRegions, adjOffset, adjList}) and the constructor logic that turns
(triangles, halfedges) into adjacency lists (first side seen per region, `s = next(halfedges[s])` circulation,
js/sphere-mesh.js:102-145), which the CPU checker in oracle/mesh_hull.py reuses.
"""
from __future__ import annotations

import numpy as np


class SphereMesh:
    """Same public fields as the reference's SphereMesh (js/sphere-mesh.js:94-146)."""

    def __init__(self, triangles: np.ndarray, halfedges: np.ndarray, num_regions: int):
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int32)
        self.halfedges = np.ascontiguousarray(halfedges, dtype=np.int32)
        self.numRegions = int(num_regions)
        self.numSides = int(self.triangles.shape[0])
        self.numTriangles = self.numSides // 3
        self.adjOffset, self.adjList, self.adjTriList = _build_csr(
            self.triangles, self.halfedges, self.numRegions)

    @classmethod
    def from_csr(cls, adj_offset: np.ndarray, adj_list: np.ndarray) -> "SphereMesh":
        """Hot-path functions need only {numRegions, adjOffset, adjList}."""
        m = cls.__new__(cls)
        m.triangles = m.halfedges = m.adjTriList = None
        m.adjOffset = np.ascontiguousarray(adj_offset, dtype=np.int32)
        m.adjList = np.ascontiguousarray(adj_list, dtype=np.int32)
        m.numRegions = int(m.adjOffset.shape[0] - 1)
        m.numSides = int(m.adjList.shape[0])
        m.numTriangles = m.numSides // 3
        return m


def _next_side(s: np.ndarray) -> np.ndarray:
    return np.where(s % 3 == 2, s - 2, s + 1)


def _build_csr(triangles, halfedges, num_regions):
    num_sides = triangles.shape[0]
    sides = np.arange(num_sides, dtype=np.int64)
    # _r_s[r] = first side whose begin vertex is r  (js/sphere-mesh.js:102-106)
    r_s = np.full(num_regions, -1, dtype=np.int64)
    order = np.argsort(triangles, kind="stable")
    tri_sorted = triangles[order]
    first = np.ones(num_sides, dtype=bool)
    first[1:] = tri_sorted[1:] != tri_sorted[:-1]
    r_s[tri_sorted[first]] = order[first]
    if (r_s < 0).any():
        raise ValueError("mesh has regions with no incident side")
    sigma = _next_side(halfedges.astype(np.int64))  # s -> next(halfedges[s])
    end_r = triangles[_next_side(sides)]            # s_end_r(s)
    # walk every region's cycle in lock-step
    cols_r, cols_t = [], []
    cur = r_s.copy()
    alive = np.ones(num_regions, dtype=bool)
    deg = np.zeros(num_regions, dtype=np.int32)
    while alive.any():
        idx = np.nonzero(alive)[0]
        c = cur[idx]
        cols_r.append((idx, deg[idx].copy(), end_r[c], (c // 3)))
        deg[idx] += 1
        nxt = sigma[c]
        cur[idx] = nxt
        alive[idx] = nxt != r_s[idx]
        if len(cols_r) > 64:
            raise ValueError("degenerate mesh: vertex degree > 64")
    adj_offset = np.zeros(num_regions + 1, dtype=np.int32)
    np.cumsum(deg, out=adj_offset[1:])
    adj_list = np.empty(int(adj_offset[-1]), dtype=np.int32)
    adj_tri = np.empty(int(adj_offset[-1]), dtype=np.int32)
    for idx, k, nb, t in cols_r:
        pos = adj_offset[idx] + k
        adj_list[pos] = nb
        adj_tri[pos] = t
    return adj_offset, adj_list, adj_tri
