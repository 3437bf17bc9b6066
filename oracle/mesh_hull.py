"""ORACLE — test infrastructure, never imported by the product path.

CPU checker for mesh construction (buildSphere, js/sphere-mesh.js:174-186): the reference triangulates the
stereographic projection with the external Delaunator 5.0.1 and closes the pole, which is the convex hull of the N+1
unit vectors.  Delaunator is not part of the reference tree and no JS runtime exists here, so the hull comes from
qhull (scipy.spatial.ConvexHull) on the f64-normalised points.  Triangle numbering — and therefore CSR neighbour
*order* — is canonical instead of Delaunator-faithful: every triangle is rotated so its smallest vertex id comes
first, triangles are sorted lexicographically and oriented counter-clockwise seen from outside; the SphereMesh
constructor logic (js/sphere-mesh.js:102-145) is then applied unchanged.  The canonical numbering is this repository's own
(there is nothing in the reference to pin it against); the reference-ordered mesh is oracle/delaunator_ref.py.
"""
from __future__ import annotations

import numpy as np

from planet_heightmap_generation_b200.mesh import SphereMesh, _next_side


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
