"""TEST INFRASTRUCTURE — checker for the Delaunator-ordered sphere mesh (csrc/pb_delaunator.h).

A plain-Python restatement of the published sweep-hull algorithm of delaunator@5.0.1, the reference's external mesh
dependency (js/sphere-mesh.js:177, loaded from a CDN at js/planet-worker.js:17; not vendored under /root/reference), followed
by the reference's own stereographicProjection / addPoleToMesh / SphereMesh constructor (js/sphere-mesh.js:41-146).
Written independently of the engine's C++ (exact orientation sign through rational arithmetic instead of floating-point
expansions) and only used by tests, on small inputs (pure Python loops).

PARITY UNPINNED for the numbering: the library is a CDN import of the reference (not in its tree), so the triangle numbering
produced here cannot be compared with the original's output; it follows the published source (index.js of the 5.0.1 tag) from
memory.  The reference vectors (tests/golden/make_reference_vectors.py) hand THIS triangulator to the reference's buildSphere.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

EPSILON = 2.0 ** -52


def orient2d(ax, ay, bx, by, cx, cy):
    """sign-exact (robust-predicates' convention: positive when a, b, c are counter-clockwise with the y axis pointing down)"""
    det = (ay - cy) * (bx - cx) - (ax - cx) * (by - cy)
    if abs(det) > 1e-12 * (abs((ay - cy) * (bx - cx)) + abs((ax - cx) * (by - cy))):
        return det
    f = Fraction
    d = (f(ay) - f(cy)) * (f(bx) - f(cx)) - (f(ax) - f(cx)) * (f(by) - f(cy))
    return 1.0 if d > 0 else -1.0 if d < 0 else 0.0


def in_circle(ax, ay, bx, by, cx, cy, px, py):
    dx, dy, ex, ey, fx, fy = ax - px, ay - py, bx - px, by - py, cx - px, cy - py
    ap, bp, cp = dx * dx + dy * dy, ex * ex + ey * ey, fx * fx + fy * fy
    return dx * (ey * cp - bp * fy) - dy * (ex * cp - bp * fx) + ap * (ex * fy - ey * fx) < 0


def _dist(ax, ay, bx, by):
    dx, dy = ax - bx, ay - by
    return dx * dx + dy * dy


def _circum(ax, ay, bx, by, cx, cy):
    dx, dy, ex, ey = bx - ax, by - ay, cx - ax, cy - ay
    bl, cl = dx * dx + dy * dy, ex * ex + ey * ey
    den = dx * ey - dy * ex
    d = 0.5 / den if den != 0 else math.inf
    return (ey * bl - dy * cl) * d, (dx * cl - ex * bl) * d


def _pseudo_angle(dx, dy):
    p = dx / (abs(dx) + abs(dy))
    return (3 - p if dy > 0 else 1 + p) / 4


def _quicksort(ids, dists, left, right):
    if right - left <= 20:
        for i in range(left + 1, right + 1):
            temp = ids[i]
            td = dists[temp]
            j = i - 1
            while j >= left and dists[ids[j]] > td:
                ids[j + 1] = ids[j]
                j -= 1
            ids[j + 1] = temp
        return
    median = (left + right) >> 1
    i, j = left + 1, right
    ids[median], ids[i] = ids[i], ids[median]
    if dists[ids[left]] > dists[ids[right]]:
        ids[left], ids[right] = ids[right], ids[left]
    if dists[ids[i]] > dists[ids[right]]:
        ids[i], ids[right] = ids[right], ids[i]
    if dists[ids[left]] > dists[ids[i]]:
        ids[left], ids[i] = ids[i], ids[left]
    temp = ids[i]
    td = dists[temp]
    while True:
        i += 1
        while dists[ids[i]] < td:
            i += 1
        j -= 1
        while dists[ids[j]] > td:
            j -= 1
        if j < i:
            break
        ids[i], ids[j] = ids[j], ids[i]
    ids[left + 1] = ids[j]
    ids[j] = temp
    if right - i + 1 >= j - left:
        _quicksort(ids, dists, i, right)
        _quicksort(ids, dists, left, j - 1)
    else:
        _quicksort(ids, dists, left, j - 1)
        _quicksort(ids, dists, i, right)


def delaunator(coords):
    """coords: flat [x0, y0, x1, y1, …] doubles → (triangles, halfedges) as lists, Delaunator 5.0.1 numbering"""
    c = [float(v) for v in coords]
    n = len(c) >> 1
    max_tri = max(2 * n - 5, 0)
    tri = [0] * (3 * max_tri)
    half = [0] * (3 * max_tri)
    hash_size = math.ceil(math.sqrt(n))
    hull_prev, hull_next, hull_tri = [0] * n, [0] * n, [0] * n
    hull_hash = [-1] * hash_size
    ids = list(range(n))
    xs, ys = c[0::2], c[1::2]
    cx, cy = (min(xs) + max(xs)) / 2, (min(ys) + max(ys)) / 2
    i0 = min(range(n), key=lambda i: (_dist(cx, cy, xs[i], ys[i]), i))
    i0x, i0y = xs[i0], ys[i0]
    i1, best = None, math.inf
    for i in range(n):
        if i == i0:
            continue
        d = _dist(i0x, i0y, xs[i], ys[i])
        if d < best and d > 0:
            i1, best = i, d
    i1x, i1y = xs[i1], ys[i1]
    i2, min_radius = None, math.inf
    for i in range(n):
        if i == i0 or i == i1:
            continue
        x, y = _circum(i0x, i0y, i1x, i1y, xs[i], ys[i])
        r = x * x + y * y
        if r < min_radius:
            i2, min_radius = i, r
    if min_radius == math.inf:
        raise ValueError("collinear input")
    i2x, i2y = xs[i2], ys[i2]
    if orient2d(i0x, i0y, i1x, i1y, i2x, i2y) < 0:
        i1, i2, i1x, i1y, i2x, i2y = i2, i1, i2x, i2y, i1x, i1y
    ox, oy = _circum(i0x, i0y, i1x, i1y, i2x, i2y)
    ccx, ccy = i0x + ox, i0y + oy
    dists = [_dist(xs[i], ys[i], ccx, ccy) for i in range(n)]
    _quicksort(ids, dists, 0, n - 1)

    st = {"len": 0, "hull_start": i0}

    def hash_key(x, y):
        return math.floor(_pseudo_angle(x - ccx, y - ccy) * hash_size) % hash_size

    def link(a, b):
        half[a] = b
        if b != -1:
            half[b] = a

    def add_triangle(a, b, cc, ha, hb, hc):
        t = st["len"]
        tri[t], tri[t + 1], tri[t + 2] = a, b, cc
        link(t, ha)
        link(t + 1, hb)
        link(t + 2, hc)
        st["len"] += 3
        return t

    def legalize(a):
        stack = []
        ar = 0
        while True:
            b = half[a]
            a0 = a - a % 3
            ar = a0 + (a + 2) % 3
            if b == -1:
                if not stack:
                    break
                a = stack.pop()
                continue
            b0 = b - b % 3
            al = a0 + (a + 1) % 3
            bl = b0 + (b + 2) % 3
            p0, pr, pl, p1 = tri[ar], tri[a], tri[al], tri[bl]
            if in_circle(xs[p0], ys[p0], xs[pr], ys[pr], xs[pl], ys[pl], xs[p1], ys[p1]):
                tri[a] = p1
                tri[b] = p0
                hbl = half[bl]
                if hbl == -1:
                    e = st["hull_start"]
                    while True:
                        if hull_tri[e] == bl:
                            hull_tri[e] = a
                            break
                        e = hull_prev[e]
                        if e == st["hull_start"]:
                            break
                link(a, hbl)
                link(b, half[ar])
                link(ar, bl)
                br = b0 + (b + 1) % 3
                if len(stack) < 512:
                    stack.append(br)
            else:
                if not stack:
                    break
                a = stack.pop()
        return ar

    hull_next[i0] = hull_prev[i2] = i1
    hull_next[i1] = hull_prev[i0] = i2
    hull_next[i2] = hull_prev[i1] = i0
    hull_tri[i0], hull_tri[i1], hull_tri[i2] = 0, 1, 2
    hull_hash[hash_key(i0x, i0y)] = i0
    hull_hash[hash_key(i1x, i1y)] = i1
    hull_hash[hash_key(i2x, i2y)] = i2
    add_triangle(i0, i1, i2, -1, -1, -1)

    xp = yp = None
    for k in range(n):
        i = ids[k]
        x, y = xs[i], ys[i]
        if k > 0 and abs(x - xp) <= EPSILON and abs(y - yp) <= EPSILON:
            continue
        xp, yp = x, y
        if i in (i0, i1, i2):
            continue
        start = 0
        key = hash_key(x, y)
        for j in range(hash_size):
            start = hull_hash[(key + j) % hash_size]
            if start != -1 and start != hull_next[start]:
                break
        start = hull_prev[start]
        e = start
        while True:
            q = hull_next[e]
            if not orient2d(x, y, xs[e], ys[e], xs[q], ys[q]) >= 0:
                break
            e = q
            if e == start:
                e = -1
                break
        if e == -1:
            continue
        t = add_triangle(e, i, hull_next[e], -1, -1, hull_tri[e])
        hull_tri[i] = legalize(t + 2)
        hull_tri[e] = t
        nx = hull_next[e]
        while True:
            q = hull_next[nx]
            if not orient2d(x, y, xs[nx], ys[nx], xs[q], ys[q]) < 0:
                break
            t = add_triangle(nx, i, q, hull_tri[i], -1, hull_tri[nx])
            hull_tri[i] = legalize(t + 2)
            hull_next[nx] = nx
            nx = q
        if e == start:
            while True:
                q = hull_prev[e]
                if not orient2d(x, y, xs[q], ys[q], xs[e], ys[e]) < 0:
                    break
                t = add_triangle(q, i, e, -1, hull_tri[e], hull_tri[q])
                legalize(t + 2)
                hull_tri[q] = t
                hull_next[e] = e
                e = q
        st["hull_start"] = hull_prev[i] = e
        hull_next[e] = hull_prev[nx] = i
        hull_next[i] = nx
        hull_hash[hash_key(x, y)] = i
        hull_hash[hash_key(xs[e], ys[e])] = e
    return tri[:st["len"]], half[:st["len"]]


def build_sphere_delaunator(r_xyz):
    """buildSphere (js/sphere-mesh.js:174-186) for the N + 1 points of r_xyz (pole last).
    Returns (triangles, halfedges, adjOffset, adjList, adjTriList) as int32 arrays."""
    p = np.ascontiguousarray(r_xyz, np.float32).reshape(-1, 3)
    num_regions = p.shape[0]
    n = num_regions - 1
    flat = []
    for i in range(n):
        z = float(p[i, 2])
        denom = max(1e-12, 1 - z)
        flat += [float(p[i, 0]) / denom, float(p[i, 1]) / denom]
    tri, half = delaunator(flat)
    num_sides = len(tri)

    def nxt(s):
        return s - 2 if s % 3 == 2 else s + 1

    unpaired = [s for s in range(num_sides) if half[s] == -1]
    point_to_side = {}
    for s in unpaired:
        point_to_side[tri[s]] = s
    nu = len(unpaired)
    nt = tri + [0] * (3 * nu)
    nh = half + [0] * (3 * nu)
    s = unpaired[-1]
    for i in range(nu):
        ns = num_sides + 3 * i
        nh[s] = ns
        nh[ns] = s
        nt[ns], nt[ns + 1], nt[ns + 2] = nt[nxt(s)], nt[s], n
        k = num_sides + (3 * i + 4) % (3 * nu)
        nh[ns + 2] = k
        nh[k] = ns + 2
        s = point_to_side[nt[nxt(s)]]
    r_s = [-1] * num_regions
    for s2, r in enumerate(nt):
        if r_s[r] == -1:
            r_s[r] = s2
    off, adj, adj_t = [0], [], []
    for r in range(num_regions):
        s0 = r_s[r]
        if s0 != -1:
            s2 = s0
            while True:
                adj.append(nt[nxt(s2)])
                adj_t.append(s2 // 3)
                s2 = nxt(nh[s2])
                if s2 == s0:
                    break
        off.append(len(adj))
    as32 = lambda a: np.asarray(a, np.int32)
    return as32(nt), as32(nh), as32(off), as32(adj), as32(adj_t)
