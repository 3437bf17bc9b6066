"""Host-side sphere mesh construction (mirror of js/sphere-mesh.js:94-186).

The reference builds its mesh with the external Delaunator 5.0.1 (stereographic 2-D
Delaunay + pole closure, js/sphere-mesh.js:174-186).  Delaunator is not part of the
reference tree and no JS runtime exists here, so the triangulation comes from the 3-D
convex hull of the (re-normalised) points, which is the same spherical Delaunay
triangulation.  Triangle numbering — and therefore CSR neighbour *order* — is made
canonical instead of Delaunator-faithful: every triangle is rotated so its smallest
vertex id comes first, triangles are sorted lexicographically, and are oriented
counter-clockwise seen from outside.  The `SphereMesh` constructor logic (first side
seen per region, `s = next(halfedges[s])` circulation, js/sphere-mesh.js:102-145) is then
applied unchanged.  Mesh construction is an input to the hot path, not part of it
(SURVEY.md §8f rank 1).
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


def triangulate_sphere(xyz: np.ndarray):
    """Spherical Delaunay of unit vectors → canonical (triangles, halfedges)."""
    from scipy.spatial import ConvexHull

    pts = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    pts = pts / np.linalg.norm(pts, axis=1, keepdims=True)
    n = pts.shape[0]
    hull = ConvexHull(pts)
    tri = hull.simplices.astype(np.int64)
    if np.unique(tri).shape[0] != n:
        raise ValueError("convex hull dropped points (duplicates?)")
    # orient CCW seen from outside: det[a,b,c] > 0 because the origin is inside the hull
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    det = np.einsum("ij,ij->i", a, np.cross(b, c))
    flip = det < 0
    tri[flip, 1], tri[flip, 2] = tri[flip, 2].copy(), tri[flip, 1].copy()
    # canonical numbering: smallest vertex first, then lexicographic triangle order
    k = np.argmin(tri, axis=1)
    rows = np.arange(tri.shape[0])
    tri = np.stack([tri[rows, k], tri[rows, (k + 1) % 3], tri[rows, (k + 2) % 3]], axis=1)
    order = np.lexsort((tri[:, 2], tri[:, 1], tri[:, 0]))
    tri = tri[order]
    triangles = tri.reshape(-1)
    # halfedges: side a->b pairs with side b->a
    num_sides = triangles.shape[0]
    s = np.arange(num_sides, dtype=np.int64)
    beg = triangles
    end = triangles[_next_side(s)]
    key_fwd = beg * n + end
    key_rev = end * n + beg
    order_f = np.argsort(key_fwd, kind="stable")
    pos = np.searchsorted(key_fwd[order_f], key_rev)
    if (pos >= num_sides).any() or (key_fwd[order_f][pos] != key_rev).any():
        raise ValueError("hull is not a closed manifold")
    halfedges = order_f[pos]
    return triangles.astype(np.int32), halfedges.astype(np.int32)


def build_sphere_from_points(r_xyz: np.ndarray):
    """r_xyz: float32 [(N+1)*3] including the pole vertex (0,0,1) as last row
    (js/sphere-mesh.js:179-185).  Returns (SphereMesh, r_xyz float32 flat)."""
    r_xyz = np.ascontiguousarray(r_xyz, dtype=np.float32).reshape(-1, 3)
    triangles, halfedges = triangulate_sphere(r_xyz)
    mesh = SphereMesh(triangles, halfedges, r_xyz.shape[0])
    return mesh, r_xyz.reshape(-1).copy()
