// pb_delaunator.h — Delaunator-faithful sphere mesh (option mesh order = "delaunator").
//
// The reference builds its mesh with the external library delaunator@5.0.1 (+ robust-predicates' orient2d), loaded from a
// CDN (js/planet-worker.js:17, js/sphere-mesh.js:177): Fibonacci points → stereographic projection → `new Delaunator(flat)` →
// addPoleToMesh → SphereMesh.  Delaunator's triangle NUMBERING decides where every region's neighbour row starts
// (`_r_s[r]` = first side whose begin vertex is r, js/sphere-mesh.js:102-106) and therefore every first-wins tie-break, BFS
// payload, randomized fill and f32 accumulation order downstream.  The device mesh builder (pb_meshgen.h) produces the same
// triangulation with a canonical numbering; this file restates the published sweep-hull algorithm of Delaunator 5.0.1 —
// seed triangle by smallest circumradius, points sorted by distance from its circumcentre, angular hull hash, stack-based
// legalisation with the floating-point inCircle, exact-sign orient2d — so that a planet can be generated with the reference's
// own neighbour order: same seed ⇒ same planet as the web app.  It is a serial host algorithm (≈ 1 s per million points);
// the default mesh order stays the 4 ms device builder.
//
// Nothing of the library is present under /root/reference (CDN dependency) and no JavaScript runtime exists in the build
// image, so the numbering cannot be pinned against the original; tests compare the triangulation with qhull, the numbering
// with an independent restatement in Python (oracle/delaunator_ref.py), and check the half-edge invariants.
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector>
#include <stdexcept>
#include <algorithm>

namespace pb {
namespace delaunator {

// ---- exact-sign orient2d with robust-predicates' convention: (ay - cy)(bx - cx) - (ax - cx)(by - cy) ----
inline void two_sum(double a, double b, double& x, double& y) { x = a + b; const double bv = x - a; y = (a - (x - bv)) + (b - bv); }
inline void two_prod(double a, double b, double& x, double& y) { x = a * b; y = fma(a, b, -x); }
// e: non-overlapping expansion of n components (increasing magnitude); adds b, returns the new length
inline int grow_expansion(double* e, int n, double b) {
    double q = b;
    int m = 0;
    for (int i = 0; i < n; i++) {
        double s, r;
        two_sum(q, e[i], s, r);
        if (r != 0) e[m++] = r;
        q = s;
    }
    if (q != 0 || m == 0) e[m++] = q;
    return m;
}
inline double orient2d(double ax, double ay, double bx, double by, double cx, double cy) {
    const double detleft = (ay - cy) * (bx - cx);
    const double detright = (ax - cx) * (by - cy);
    const double det = detleft - detright;
    double detsum;
    if (detleft > 0) { if (detright <= 0) return det; detsum = detleft + detright; }
    else if (detleft < 0) { if (detright >= 0) return det; detsum = -detleft - detright; }
    else return det;
    const double errbound = 3.3306690738754716e-16 * detsum;        // (3 + 16 eps) eps, eps = 2^-53
    if (det >= errbound || -det >= errbound) return det;
    // exact: det = ay bx - ay cx - cy bx - ax by + ax cy + cx by   (the cy cx terms cancel)
    const double pa[6] = {ay, -ay, -cy, -ax, ax, cx};
    const double pb_[6] = {bx, cx, bx, by, cy, by};
    double e[16];
    int n = 0;
    for (int k = 0; k < 6; k++) {
        double hi, lo;
        two_prod(pa[k], pb_[k], hi, lo);
        n = grow_expansion(e, n, lo);
        n = grow_expansion(e, n, hi);
    }
    return e[n - 1];          // the most significant component carries the sign
}

inline double dist2(double ax, double ay, double bx, double by) { const double dx = ax - bx, dy = ay - by; return dx * dx + dy * dy; }
inline bool in_circle(double ax, double ay, double bx, double by, double cx, double cy, double px, double py) {
    const double dx = ax - px, dy = ay - py, ex = bx - px, ey = by - py, fx = cx - px, fy = cy - py;
    const double ap = dx * dx + dy * dy, bp = ex * ex + ey * ey, cp = fx * fx + fy * fy;
    return dx * (ey * cp - bp * fy) - dy * (ex * cp - bp * fx) + ap * (ex * fy - ey * fx) < 0;
}
inline double circumradius(double ax, double ay, double bx, double by, double cx, double cy) {
    const double dx = bx - ax, dy = by - ay, ex = cx - ax, ey = cy - ay;
    const double bl = dx * dx + dy * dy, cl = ex * ex + ey * ey, d = 0.5 / (dx * ey - dy * ex);
    const double x = (ey * bl - dy * cl) * d, y = (dx * cl - ex * bl) * d;
    return x * x + y * y;
}
inline void circumcenter(double ax, double ay, double bx, double by, double cx, double cy, double& ox, double& oy) {
    const double dx = bx - ax, dy = by - ay, ex = cx - ax, ey = cy - ay;
    const double bl = dx * dx + dy * dy, cl = ex * ex + ey * ey, d = 0.5 / (dx * ey - dy * ex);
    ox = ax + (ey * bl - dy * cl) * d;
    oy = ay + (dx * cl - ex * bl) * d;
}
inline double pseudo_angle(double dx, double dy) {
    const double p = dx / (fabs(dx) + fabs(dy));
    return (dy > 0 ? 3 - p : 1 + p) / 4;
}

struct Delaunator {
    const double* coords; int n;
    std::vector<uint32_t> triangles; std::vector<int32_t> halfedges;
    std::vector<uint32_t> hullPrev, hullNext, hullTri, ids;
    std::vector<int32_t> hullHash;
    std::vector<double> dists;
    int hashSize = 0, hullStart = 0, trianglesLen = 0;
    double cx = 0, cy = 0;
    uint32_t edgeStack[512];

    int hash_key(double x, double y) const {
        // Math.floor(pseudoAngle * hashSize) % hashSize; a NaN angle (point on the centre) gives NaN % n → index NaN: the
        // reference then reads undefined; such input does not occur for distinct points
        return (int)((long long)floor(pseudo_angle(x - cx, y - cy) * hashSize) % hashSize);
    }
    void link(int a, int b) { halfedges[a] = b; if (b != -1) halfedges[b] = a; }
    int add_triangle(int i0, int i1, int i2, int a, int b, int c) {
        const int t = trianglesLen;
        triangles[t] = i0; triangles[t + 1] = i1; triangles[t + 2] = i2;
        link(t, a); link(t + 1, b); link(t + 2, c);
        trianglesLen += 3;
        return t;
    }
    int legalize(int a) {
        int i = 0, ar = 0;
        for (;;) {
            const int b = halfedges[a];
            const int a0 = a - a % 3;
            ar = a0 + (a + 2) % 3;
            if (b == -1) {
                if (i == 0) break;
                a = (int)edgeStack[--i];
                continue;
            }
            const int b0 = b - b % 3;
            const int al = a0 + (a + 1) % 3;
            const int bl = b0 + (b + 2) % 3;
            const uint32_t p0 = triangles[ar], pr = triangles[a], pl = triangles[al], p1 = triangles[bl];
            const bool illegal = in_circle(coords[2 * p0], coords[2 * p0 + 1], coords[2 * pr], coords[2 * pr + 1],
                                           coords[2 * pl], coords[2 * pl + 1], coords[2 * p1], coords[2 * p1 + 1]);
            if (illegal) {
                triangles[a] = p1;
                triangles[b] = p0;
                const int hbl = halfedges[bl];
                if (hbl == -1) {      // edge swapped on the other side of the hull: fix the hull's triangle reference
                    uint32_t e = (uint32_t)hullStart;
                    do {
                        if (hullTri[e] == (uint32_t)bl) { hullTri[e] = (uint32_t)a; break; }
                        e = hullPrev[e];
                    } while (e != (uint32_t)hullStart);
                }
                link(a, hbl);
                link(b, halfedges[ar]);
                link(ar, bl);
                const int br = b0 + (b + 1) % 3;
                if (i < 512) edgeStack[i++] = (uint32_t)br;
            } else {
                if (i == 0) break;
                a = (int)edgeStack[--i];
            }
        }
        return ar;
    }
    static void swap_ids(std::vector<uint32_t>& a, long long i, long long j) { const uint32_t t = a[i]; a[i] = a[j]; a[j] = t; }
    void quicksort(long long left, long long right) {
        if (right - left <= 20) {
            for (long long i = left + 1; i <= right; i++) {
                const uint32_t temp = ids[i];
                const double tempDist = dists[temp];
                long long j = i - 1;
                while (j >= left && dists[ids[j]] > tempDist) { ids[j + 1] = ids[j]; j--; }
                ids[j + 1] = temp;
            }
        } else {
            const long long median = (left + right) >> 1;
            long long i = left + 1, j = right;
            swap_ids(ids, median, i);
            if (dists[ids[left]] > dists[ids[right]]) swap_ids(ids, left, right);
            if (dists[ids[i]] > dists[ids[right]]) swap_ids(ids, i, right);
            if (dists[ids[left]] > dists[ids[i]]) swap_ids(ids, left, i);
            const uint32_t temp = ids[i];
            const double tempDist = dists[temp];
            for (;;) {
                do i++; while (dists[ids[i]] < tempDist);
                do j--; while (dists[ids[j]] > tempDist);
                if (j < i) break;
                swap_ids(ids, i, j);
            }
            ids[left + 1] = ids[j];
            ids[j] = temp;
            if (right - i + 1 >= j - left) { quicksort(i, right); quicksort(left, j - 1); }
            else { quicksort(left, j - 1); quicksort(i, right); }
        }
    }

    Delaunator(const double* c, int nPoints) : coords(c), n(nPoints) {
        if (n < 3) throw std::invalid_argument("Delaunator needs at least three points");
        const int maxTriangles = std::max(2 * n - 5, 0);
        triangles.assign((size_t)maxTriangles * 3, 0); halfedges.assign((size_t)maxTriangles * 3, 0);
        hashSize = (int)ceil(sqrt((double)n));
        hullPrev.assign(n, 0); hullNext.assign(n, 0); hullTri.assign(n, 0); hullHash.assign(hashSize, -1);
        ids.resize(n); dists.resize(n);
        double minX = INFINITY, minY = INFINITY, maxX = -INFINITY, maxY = -INFINITY;
        for (int i = 0; i < n; i++) {
            const double x = coords[2 * i], y = coords[2 * i + 1];
            if (x < minX) minX = x;
            if (y < minY) minY = y;
            if (x > maxX) maxX = x;
            if (y > maxY) maxY = y;
            ids[i] = (uint32_t)i;
        }
        const double bx = (minX + maxX) / 2, by = (minY + maxY) / 2;
        int i0 = -1, i1 = -1, i2 = -1;
        { double minDist = INFINITY; for (int i = 0; i < n; i++) { const double d = dist2(bx, by, coords[2 * i], coords[2 * i + 1]); if (d < minDist) { i0 = i; minDist = d; } } }
        if (i0 < 0) throw std::invalid_argument("Delaunator: no finite point");
        const double i0x = coords[2 * i0], i0y = coords[2 * i0 + 1];
        { double minDist = INFINITY; for (int i = 0; i < n; i++) { if (i == i0) continue; const double d = dist2(i0x, i0y, coords[2 * i], coords[2 * i + 1]); if (d < minDist && d > 0) { i1 = i; minDist = d; } } }
        if (i1 < 0) throw std::invalid_argument("Delaunator: all points coincide");
        double i1x = coords[2 * i1], i1y = coords[2 * i1 + 1];
        double minRadius = INFINITY;
        for (int i = 0; i < n; i++) {
            if (i == i0 || i == i1) continue;
            const double r = circumradius(i0x, i0y, i1x, i1y, coords[2 * i], coords[2 * i + 1]);
            if (r < minRadius) { i2 = i; minRadius = r; }
        }
        if (minRadius == INFINITY || i2 < 0) throw std::invalid_argument("Delaunator: all points are collinear");
        double i2x = coords[2 * i2], i2y = coords[2 * i2 + 1];
        if (orient2d(i0x, i0y, i1x, i1y, i2x, i2y) < 0) {
            const int i = i1; const double x = i1x, y = i1y;
            i1 = i2; i1x = i2x; i1y = i2y;
            i2 = i; i2x = x; i2y = y;
        }
        circumcenter(i0x, i0y, i1x, i1y, i2x, i2y, cx, cy);
        for (int i = 0; i < n; i++) dists[i] = dist2(coords[2 * i], coords[2 * i + 1], cx, cy);
        quicksort(0, n - 1);

        hullStart = i0;
        hullNext[i0] = hullPrev[i2] = (uint32_t)i1;
        hullNext[i1] = hullPrev[i0] = (uint32_t)i2;
        hullNext[i2] = hullPrev[i1] = (uint32_t)i0;
        hullTri[i0] = 0; hullTri[i1] = 1; hullTri[i2] = 2;
        hullHash[hash_key(i0x, i0y)] = i0;
        hullHash[hash_key(i1x, i1y)] = i1;
        hullHash[hash_key(i2x, i2y)] = i2;
        trianglesLen = 0;
        add_triangle(i0, i1, i2, -1, -1, -1);

        double xp = 0, yp = 0;
        for (int k = 0; k < n; k++) {
            const int i = (int)ids[k];
            const double x = coords[2 * i], y = coords[2 * i + 1];
            if (k > 0 && fabs(x - xp) <= 2.220446049250313e-16 && fabs(y - yp) <= 2.220446049250313e-16) continue;   // near-duplicate
            xp = x; yp = y;
            if (i == i0 || i == i1 || i == i2) continue;
            int start = 0;
            { const int key = hash_key(x, y);
              for (int j = 0; j < hashSize; j++) { start = hullHash[(key + j) % hashSize]; if (start != -1 && (uint32_t)start != hullNext[start]) break; } }
            start = (int)hullPrev[start];
            int e = start, q;
            for (;;) {
                q = (int)hullNext[e];
                if (!(orient2d(x, y, coords[2 * e], coords[2 * e + 1], coords[2 * q], coords[2 * q + 1]) >= 0)) break;
                e = q;
                if (e == start) { e = -1; break; }
            }
            if (e == -1) continue;
            int t = add_triangle(e, i, (int)hullNext[e], -1, -1, (int)hullTri[e]);
            hullTri[i] = (uint32_t)legalize(t + 2);
            hullTri[e] = (uint32_t)t;
            int nx = (int)hullNext[e];
            for (;;) {
                q = (int)hullNext[nx];
                if (!(orient2d(x, y, coords[2 * nx], coords[2 * nx + 1], coords[2 * q], coords[2 * q + 1]) < 0)) break;
                t = add_triangle(nx, i, q, (int)hullTri[i], -1, (int)hullTri[nx]);
                hullTri[i] = (uint32_t)legalize(t + 2);
                hullNext[nx] = (uint32_t)nx;
                nx = q;
            }
            if (e == start) {
                for (;;) {
                    q = (int)hullPrev[e];
                    if (!(orient2d(x, y, coords[2 * q], coords[2 * q + 1], coords[2 * e], coords[2 * e + 1]) < 0)) break;
                    t = add_triangle(q, i, e, -1, (int)hullTri[e], (int)hullTri[q]);
                    legalize(t + 2);
                    hullTri[q] = (uint32_t)t;
                    hullNext[e] = (uint32_t)e;
                    e = q;
                }
            }
            hullStart = e;
            hullPrev[i] = (uint32_t)e;
            hullNext[e] = hullPrev[nx] = (uint32_t)i;
            hullNext[i] = (uint32_t)nx;
            hullHash[hash_key(x, y)] = i;
            hullHash[hash_key(coords[2 * e], coords[2 * e + 1])] = e;
        }
        triangles.resize((size_t)trianglesLen); halfedges.resize((size_t)trianglesLen);
    }
};

// buildSphere (js/sphere-mesh.js:174-186) from the N + 1 points (pole last): stereographic projection :41-53, Delaunator,
// addPoleToMesh :56-91, SphereMesh constructor :95-146.  Outputs the closed triangle / half-edge arrays and the CSR with the
// per-edge inner triangles.
struct SphereMeshHost {
    std::vector<int> triangles, halfedges, adjOffset, adjList, adjTri;
};
inline void build_sphere(const float* r_xyz, int numRegions, SphereMeshHost& out) {
    const int N = numRegions - 1;          // the last point is the pole
    std::vector<double> flat(2 * (size_t)N);
    for (int i = 0; i < N; i++) {
        const double z = r_xyz[3 * i + 2];
        const double denom = std::max(1e-12, 1 - z);
        flat[2 * i] = (double)r_xyz[3 * i] / denom;
        flat[2 * i + 1] = (double)r_xyz[3 * i + 1] / denom;
    }
    Delaunator d(flat.data(), N);
    const int numSides = (int)d.triangles.size();
    auto next = [](int s) { return (s % 3 == 2) ? s - 2 : s + 1; };
    int numUnpaired = 0, firstUnpaired = -1;
    std::vector<int> pointToSide(numRegions, -1);
    for (int s = 0; s < numSides; s++)
        if (d.halfedges[s] == -1) { numUnpaired++; pointToSide[d.triangles[s]] = s; firstUnpaired = s; }
    std::vector<int>& nt = out.triangles; std::vector<int>& nh = out.halfedges;
    nt.assign((size_t)numSides + 3 * (size_t)numUnpaired, 0); nh.assign(nt.size(), 0);
    for (int s = 0; s < numSides; s++) { nt[s] = (int)d.triangles[s]; nh[s] = d.halfedges[s]; }
    for (int i = 0, s = firstUnpaired; i < numUnpaired; i++, s = pointToSide[nt[next(s)]]) {
        const int ns = numSides + 3 * i;
        nh[s] = ns; nh[ns] = s;
        nt[ns] = nt[next(s)]; nt[ns + 1] = nt[s]; nt[ns + 2] = N;
        const int k = numSides + (3 * i + 4) % (3 * numUnpaired);
        nh[ns + 2] = k; nh[k] = ns + 2;
    }
    const int S = (int)nt.size();
    std::vector<int> r_s(numRegions, -1);
    for (int s = 0; s < S; s++) { const int r = nt[s]; if (r_s[r] == -1) r_s[r] = s; }
    out.adjOffset.assign((size_t)numRegions + 1, 0);
    for (int r = 0; r < numRegions; r++) {
        int cnt = 0;
        const int s0 = r_s[r];
        if (s0 != -1) {
            int s = s0;
            do { cnt++; s = next(nh[s]); if (cnt > S) throw std::runtime_error("Delaunator mesh: open circulation"); } while (s != s0);
        }
        out.adjOffset[r + 1] = out.adjOffset[r] + cnt;
    }
    out.adjList.assign((size_t)out.adjOffset[numRegions], 0); out.adjTri.assign(out.adjList.size(), 0);
    for (int r = 0; r < numRegions; r++) {
        const int s0 = r_s[r];
        if (s0 == -1) continue;
        int s = s0, idx = out.adjOffset[r];
        do { out.adjList[idx] = nt[next(s)]; out.adjTri[idx] = s / 3; idx++; s = next(nh[s]); } while (s != s0);
    }
}

}  // namespace delaunator
}  // namespace pb
