// pb_meshgen.h — spherical Delaunay adjacency on the device (the "mesh→" prefix of the path).
//
// Replaces buildSphere's triangulation + the SphereMesh constructor (js/sphere-mesh.js:94-146, 174-186): the
// reference runs the external Delaunator 5.0.1 on the stereographic projection and then closes the pole, which yields
// the convex hull of the N+1 unit vectors; its per-region neighbour list is the half-edge circulation
// `s = next(halfedges[s])` started at the first side whose begin vertex is the region.  Here every region computes
// its own star, independently of all others:
//   1. points are normalised in f64 and sorted by the key of a uniform 3-D grid (CUB radix sort);
//   2. one thread per region gift-wraps the hull faces around its point: the nearest point is a neighbour; from a hull
//      edge (r, q) the next neighbour t is the candidate that has every other candidate below the plane (r, q, t).
//      Only orientation determinants are used (f64 with a static error bound, double-double when the bound fails), so
//      the stars of different regions agree with each other;
//   3. candidates come from the (2L+1)³ grid block around the region; the star is accepted when the circumscribed cap
//      of every face lies inside the searched block (2·ρ ≤ L·cellsize), otherwise L grows;
//   4. the ring is written in the reference's circulation order for a canonical triangle numbering (each triangle
//      rotated to start at its smallest vertex, triangles in lexicographic order, counter-clockwise seen from
//      outside — the numbering planet_heightmap_generation_b200/mesh.py uses, Delaunator's own numbering being
//      unavailable: SURVEY §8c), which is a purely local rule: start at the triangle with the smallest canonical
//      tuple, then walk clockwise.
// A count pass, an exclusive scan and a fill pass produce the CSR arrays; a last pass checks that the adjacency is
// symmetric and that the edge count is 6n-12 (Euler), so an inconsistent predicate cannot go unnoticed.
#pragma once
#include "pb_prims.h"

namespace pb {

struct DD { double hi, lo; };
PB_DEV DD dd_two_sum(double a, double b) { double s = a + b, bb = s - a; return {s, (a - (s - bb)) + (b - bb)}; }
PB_DEV DD dd_two_prod(double a, double b) { double p = a * b; return {p, fma(a, b, -p)}; }
PB_DEV DD dd_add(DD a, DD b) {
    DD s = dd_two_sum(a.hi, b.hi);
    double lo = s.lo + (a.lo + b.lo);
    double hi = s.hi + lo;
    return {hi, lo - (hi - s.hi)};
}
PB_DEV DD dd_neg(DD a) { return {-a.hi, -a.lo}; }
PB_DEV DD dd_mul(DD a, DD b) {
    DD p = dd_two_prod(a.hi, b.hi);
    double lo = p.lo + (a.hi * b.lo + a.lo * b.hi);
    double hi = p.hi + lo;
    return {hi, lo - (hi - p.hi)};
}
PB_DEV DD dd_diff(double a, double b) { return dd_two_sum(a, -b); }

struct P3 { double x, y, z; };

// sign of det[q-r, b-r, c-r]: > 0 when c lies above the plane through r, q, b oriented by (q-r)×(b-r)
PB_DEV int orient_sign(const P3& r, const P3& q, const P3& b, const P3& c) {
    const double ax = q.x - r.x, ay = q.y - r.y, az = q.z - r.z;
    const double bx = b.x - r.x, by = b.y - r.y, bz = b.z - r.z;
    const double cx = c.x - r.x, cy = c.y - r.y, cz = c.z - r.z;
    const double m1 = by * cz, m2 = bz * cy, m3 = bz * cx, m4 = bx * cz, m5 = bx * cy, m6 = by * cx;
    const double det = ax * (m1 - m2) + ay * (m3 - m4) + az * (m5 - m6);
    const double perm = fabs(ax) * (fabs(m1) + fabs(m2)) + fabs(ay) * (fabs(m3) + fabs(m4)) + fabs(az) * (fabs(m5) + fabs(m6));
    const double bound = 4e-15 * perm;
    if (det > bound) return 1;
    if (det < -bound) return -1;
    // double-double re-evaluation (differences exact, products and sums to ~1e-31 relative)
    const DD Ax = dd_diff(q.x, r.x), Ay = dd_diff(q.y, r.y), Az = dd_diff(q.z, r.z);
    const DD Bx = dd_diff(b.x, r.x), By = dd_diff(b.y, r.y), Bz = dd_diff(b.z, r.z);
    const DD Cx = dd_diff(c.x, r.x), Cy = dd_diff(c.y, r.y), Cz = dd_diff(c.z, r.z);
    const DD t1 = dd_mul(Ax, dd_add(dd_mul(By, Cz), dd_neg(dd_mul(Bz, Cy))));
    const DD t2 = dd_mul(Ay, dd_add(dd_mul(Bz, Cx), dd_neg(dd_mul(Bx, Cz))));
    const DD t3 = dd_mul(Az, dd_add(dd_mul(Bx, Cy), dd_neg(dd_mul(By, Cx))));
    const DD d = dd_add(dd_add(t1, t2), t3);
    const double v = d.hi + d.lo;
    if (fabs(v) <= 1e-28 * perm) return 0;
    return v > 0 ? 1 : -1;
}

struct GridSpec {
    int G;             // cells per axis
    double inv;        // G / 2
    double cell;       // 2 / G
    PB_DEV int coord(double v) const { int i = (int)((v + 1.0) * inv); return i < 0 ? 0 : (i >= G ? G - 1 : i); }
    PB_DEV uint32_t key(int ix, int iy, int iz) const { return ((uint32_t)ix * (uint32_t)G + (uint32_t)iy) * (uint32_t)G + (uint32_t)iz; }
};

// normalised f64 coordinates (x / sqrt((x²+y²)+z²), the order numpy's norm uses) and the grid key of every point
struct MeshGenKeyK {
    const float* xyz; GridSpec g; uint32_t* key; int* id;
    PB_DEV void operator()(int i) const {
        const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        const double n = sqrt((x * x + y * y) + z * z);
        key[i] = g.key(g.coord(x / n), g.coord(y / n), g.coord(z / n));
        id[i] = i;
    }
};
struct MeshGenGatherK {
    const float* xyz; const int* sid; double* sx; double* sy; double* sz;
    PB_DEV void operator()(int j) const {
        const int i = sid[j];
        const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        const double n = sqrt((x * x + y * y) + z * z);
        sx[j] = x / n; sy[j] = y / n; sz[j] = z / n;
    }
};

constexpr int kMaxRing = 32;
constexpr int kMaxBlock = 12;
constexpr int kRowCache = 12;

struct StarK {
    int n; GridSpec g;
    const uint32_t* skey; const int* sid; const double* sx; const double* sy; const double* sz;
    int* deg;              // count pass: deg[id]          (adj == nullptr)
    int* rows;             // count pass: the finished row of every region of degree <= kRowCache, stride kRowCache
    const int* off; int* adj;   // fill pass (only rows longer than kRowCache are recomputed)
    int* fail;             // [0] stars that could not be closed, [1] exact degeneracies met

    PB_DEV int lower_bound(uint32_t k) const {
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (skey[mid] < k) lo = mid + 1; else hi = mid; }
        return lo;
    }
    PB_DEV P3 pt(int j) const { return {sx[j], sy[j], sz[j]}; }

    // the candidates of a block: per (x, y) column the z-run [z0, z1] is one contiguous range of the sorted keys
    struct Block {
        int x0, x1, y0, y1, z0, z1, nCol;
        int lo[25], hi[25];          // cached ranges when the block has at most 25 columns (L <= 2)
    };
    PB_DEV void open_block(Block& b, int ix, int iy, int iz, int L) const {
        b.x0 = ix - L < 0 ? 0 : ix - L; b.x1 = ix + L > g.G - 1 ? g.G - 1 : ix + L;
        b.y0 = iy - L < 0 ? 0 : iy - L; b.y1 = iy + L > g.G - 1 ? g.G - 1 : iy + L;
        b.z0 = iz - L < 0 ? 0 : iz - L; b.z1 = iz + L > g.G - 1 ? g.G - 1 : iz + L;
        b.nCol = (b.x1 - b.x0 + 1) * (b.y1 - b.y0 + 1);
        if (b.nCol <= 25) {
            int k = 0;
            for (int cx = b.x0; cx <= b.x1; cx++) for (int cy = b.y0; cy <= b.y1; cy++, k++) {
                b.lo[k] = lower_bound(g.key(cx, cy, b.z0)); b.hi[k] = lower_bound(g.key(cx, cy, b.z1) + 1u);
            }
        }
    }
    template <class F>
    PB_DEV void for_candidates(const Block& b, const F& f) const {
        if (b.nCol <= 25) {
            for (int k = 0; k < b.nCol; k++) for (int c = b.lo[k]; c < b.hi[k]; c++) f(c);
        } else {
            for (int cx = b.x0; cx <= b.x1; cx++) for (int cy = b.y0; cy <= b.y1; cy++) {
                const int lo = lower_bound(g.key(cx, cy, b.z0)), hi = lower_bound(g.key(cx, cy, b.z1) + 1u);
                for (int c = lo; c < hi; c++) f(c);
            }
        }
    }

    // ring of sorted indices, counter-clockwise seen from outside; returns the degree or -1 (block too small / not closed)
    PB_DEV int wrap(int j, const P3& r, int ix, int iy, int iz, int L, int* ring, bool whole) const {
        Block blk;
        open_block(blk, ix, iy, iz, L);
        // nearest point: always a Delaunay neighbour
        int q0 = -1; double best = 1e300;
        for_candidates(blk, [&](int c) {
            if (c == j) return;
            const double dx = sx[c] - r.x, dy = sy[c] - r.y, dz = sz[c] - r.z;
            const double d2 = (dx * dx + dy * dy) + dz * dz;
            if (d2 < best || (d2 == best && sid[c] < sid[q0])) { best = d2; q0 = c; }
        });
        if (q0 < 0) return -1;
        const double reach = (double)L * g.cell * (1.0 - 1e-9);
        if (!whole && best > reach * reach) return -1;
        int d = 0, q = q0;
        while (true) {
            if (d >= kMaxRing) return -1;
            ring[d++] = q;
            const P3 pq = pt(q);
            int t = -1; P3 ptc{0, 0, 0};
            for_candidates(blk, [&](int c) {
                if (c == j || c == q) return;
                const P3 pc = pt(c);
                if (t < 0) { t = c; ptc = pc; return; }
                const int s = orient_sign(r, pq, ptc, pc);
                if (s > 0) { t = c; ptc = pc; }
                else if (s == 0) {
                    atomic_add(fail + 1, 1);
                    // four co-circular points: keep the candidate nearer to r (either choice is a Delaunay triangulation)
                    const double ex = pc.x - r.x, ey = pc.y - r.y, ez = pc.z - r.z, fx = ptc.x - r.x, fy = ptc.y - r.y, fz = ptc.z - r.z;
                    if ((ex * ex + ey * ey) + ez * ez < (fx * fx + fy * fy) + fz * fz) { t = c; ptc = pc; }
                }
            });
            if (t < 0) return -1;
            // the cap cut off by the plane (r, q, t) must lie inside the searched block: 2·circumradius <= L·cell
            if (!whole) {
                const double ax = pq.x - r.x, ay = pq.y - r.y, az = pq.z - r.z, bx = ptc.x - r.x, by = ptc.y - r.y, bz = ptc.z - r.z;
                const double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
                const double cx2 = pq.x - ptc.x, cy2 = pq.y - ptc.y, cz2 = pq.z - ptc.z;
                const double la = (ax * ax + ay * ay) + az * az, lb = (bx * bx + by * by) + bz * bz, lc = (cx2 * cx2 + cy2 * cy2) + cz2 * cz2;
                const double n2 = (nx * nx + ny * ny) + nz * nz;
                // (2ρ)² = la·lb·lc / |n|²
                if (!(la * lb * lc <= reach * reach * n2)) return -1;
            }
            if (t == q0) return d;
            q = t;
        }
    }

    PB_DEV void operator()(int j) const {
        const P3 r = pt(j);
        const int ix = g.coord(r.x), iy = g.coord(r.y), iz = g.coord(r.z);
        if (adj && deg[sid[j]] <= kRowCache) return;     // fill pass called directly: cached rows are handled by StarFillK
        int ring[kMaxRing];
        int d = -1;
        for (int L = 1; L <= kMaxBlock && d < 0; L++) {
            const bool whole = L >= g.G - 1;
            d = wrap(j, r, ix, iy, iz, L, ring, whole);
            if (whole) break;
        }
        const int me = sid[j];
        if (d < 3) {
            atomic_add(fail, 1);
            if (!adj) deg[me] = 0;
            return;
        }
        if (!adj) deg[me] = d;
        // canonical start: the incident triangle (me, ring[i], ring[i+1]) whose rotation-to-smallest-vertex tuple is smallest
        int ids[kMaxRing];
        for (int i = 0; i < d; i++) ids[i] = sid[ring[i]];
        int bi = 0, b0 = 0, b1 = 0, b2 = 0;
        for (int i = 0; i < d; i++) {
            const int u = ids[i], v = ids[i + 1 == d ? 0 : i + 1];
            int t0, t1, t2;
            if (me < u && me < v) { t0 = me; t1 = u; t2 = v; }
            else if (u < v) { t0 = u; t1 = v; t2 = me; }
            else { t0 = v; t1 = me; t2 = u; }
            if (i == 0 || t0 < b0 || (t0 == b0 && (t1 < b1 || (t1 == b1 && t2 < b2)))) { bi = i; b0 = t0; b1 = t1; b2 = t2; }
        }
        // first side = me → ring[bi]; `s = next(halfedges[s])` then visits the ring clockwise (js/sphere-mesh.js:133-143)
        int* row;
        if (adj) row = adj + off[me];
        else if (d <= kRowCache) row = rows + (size_t)me * kRowCache;
        else return;
        for (int k = 0; k < d; k++) { int i = bi - k; if (i < 0) i += d; row[k] = ids[i]; }
    }
};

// fill pass: rows cached by the count pass are copied, longer ones recomputed
struct StarFillK {
    StarK star; const int* pos;    // pos[id] = sorted index of region id
    PB_DEV void operator()(int me) const {
        const int d = star.deg[me];
        if (d <= kRowCache) {
            const int* src = star.rows + (size_t)me * kRowCache;
            int* row = star.adj + star.off[me];
            for (int k = 0; k < d; k++) row[k] = src[k];
        } else {
            star(pos[me]);
        }
    }
};
struct MeshGenInvertK {
    const int* sid; int* pos;
    PB_DEV void operator()(int j) const { pos[sid[j]] = j; }
};

// Second chance for the stars StarK could not close within its block limit (point sets whose density varies by orders of
// magnitude): the block keeps growing until it is the whole grid.  Launched only when the count pass reports failures, so
// evenly spread points — every planet the reference generates — never run it.  mode 0: degree; mode 1: row into adj.
struct StarRetryK {
    StarK star; const int* pos; uint8_t* retried; int mode;
    PB_DEV void operator()(int me) const {
        if (mode == 0 ? star.deg[me] != 0 : !retried[me]) return;
        const int j = pos[me];
        const P3 r = star.pt(j);
        const GridSpec& g = star.g;
        const int ix = g.coord(r.x), iy = g.coord(r.y), iz = g.coord(r.z);
        int ring[kMaxRing];
        int d = -1;
        for (int L = kMaxBlock + 1; d < 0; L += (L >> 1)) {
            const bool whole = L >= g.G - 1;
            d = star.wrap(j, r, ix, iy, iz, whole ? g.G : L, ring, whole);
            if (whole) break;
        }
        if (d < 3) { if (mode == 0) atomic_add(star.fail, 1); return; }
        if (mode == 0) { star.deg[me] = d; retried[me] = 1; return; }
        int ids[kMaxRing];
        for (int i = 0; i < d; i++) ids[i] = star.sid[ring[i]];
        int bi = 0, b0 = 0, b1 = 0, b2 = 0;                       // same canonical start as StarK
        for (int i = 0; i < d; i++) {
            const int u = ids[i], v = ids[i + 1 == d ? 0 : i + 1];
            int t0, t1, t2;
            if (me < u && me < v) { t0 = me; t1 = u; t2 = v; }
            else if (u < v) { t0 = u; t1 = v; t2 = me; }
            else { t0 = v; t1 = me; t2 = u; }
            if (i == 0 || t0 < b0 || (t0 == b0 && (t1 < b1 || (t1 == b1 && t2 < b2)))) { bi = i; b0 = t0; b1 = t1; b2 = t2; }
        }
        int* row = star.adj + star.off[me];
        for (int k = 0; k < d; k++) { int i = bi - k; if (i < 0) i += d; row[k] = ids[i]; }
    }
};

// adjacency must be symmetric: every edge r→q has its twin q→r
struct MeshGenSymmetryK {
    int n; const int* off; const int* adj; int* bad;
    PB_DEV void operator()(int r) const {
        for (int i = off[r]; i < off[r + 1]; i++) {
            const int q = adj[i];
            bool found = false;
            for (int k = off[q]; k < off[q + 1]; k++) found |= adj[k] == r;
            if (!found) atomic_add(bad, 1);
        }
    }
};

// generateFibonacciSphere (js/sphere-mesh.js:9-37) + the pole vertex buildSphere appends (:179-183).  z and lng are the
// reference's running sums (z -= dz, lng += dlong: every step is rounded, so they are accumulated serially on the
// host); the four jitter draws of point k are Park–Miller outputs 4k+1 … 4k+4 (js/rng.js:3-6), reached by jump-ahead:
// state_m = s0 · 16807^m mod (2^31 - 1).
struct FibonacciK {
    int N; double jitter, s, dz; const double* z; const double* lng; unsigned long long s0; float* xyz;
    static PB_DEV unsigned long long mulmod(unsigned long long a, unsigned long long b) { return (a * b) % 2147483647ull; }
    PB_DEV void operator()(int k) const {
        if (k == N) { xyz[3 * k] = 0.0f; xyz[3 * k + 1] = 0.0f; xyz[3 * k + 2] = 1.0f; return; }
        const double zz = z[k];
        const double r = sqrt(1 - zz * zz);
        double latDeg = pb_asin(zz) * 180 / PB_PI;
        double lonDeg = lng[k] * 180 / PB_PI;
        if (jitter > 0) {
            unsigned long long st = s0, base = 16807ull, e = 4ull * (unsigned long long)k;
            while (e) { if (e & 1ull) st = mulmod(st, base); base = mulmod(base, base); e >>= 1; }
            double u[4];
            for (int i = 0; i < 4; i++) { st = mulmod(st, 16807ull); u[i] = (double)(st - 1ull) / 2147483646.0; }
            const double jLat = u[0] - u[1], jLon = u[2] - u[3];
            double nextZ = zz - dz * 2 * PB_PI * r / s;
            if (nextZ < -1) nextZ = -1;
            latDeg += jitter * jLat * (latDeg - pb_asin(nextZ) * 180 / PB_PI);
            lonDeg += jitter * jLon * (s / r * 180 / PB_PI);
        }
        const double latR = latDeg * PB_PI / 180, lonR = lonDeg * PB_PI / 180;
        xyz[3 * k] = (float)(pb_cos(latR) * pb_cos(lonR));
        xyz[3 * k + 1] = (float)(pb_cos(latR) * pb_sin(lonR));
        xyz[3 * k + 2] = (float)pb_sin(latR);
    }
};

struct FibonacciSphere {
    DevBuf<double> z, lng;
    std::vector<double> hz, hlng;
    int cachedN = 0;
    // dXyz: device f32[3(N+1)]
    void generate(const Exec& ex, int N, double jitter, double seed, float* dXyz) {
        if (N < 1) throw Error("numPoints must be positive");
        const double dz = 2.0 / N, dlong = PB_PI * (3 - sqrt(5.0));
        if (cachedN != N) {                      // the two running sums depend on N only
            hz.resize(N); hlng.resize(N);
            double l = 0, zz = 1 - dz / 2;
            for (int k = 0; k < N; k++, zz -= dz) { hz[k] = zz; hlng[k] = l; l += dlong; }
            dev_copy(z.ensure(N), hz.data(), sizeof(double) * (size_t)N, 0, ex.stream);
            dev_copy(lng.ensure(N), hlng.data(), sizeof(double) * (size_t)N, 0, ex.stream);
            cachedN = N;
        }
        const unsigned long long s0 = (unsigned long long)(fmod(fabs(floor(seed * 9301.0 + 49297.0)), 2147483646.0) + 1.0);
        ex.for_each(N + 1, FibonacciK{N, jitter, 3.6 / sqrt((double)N), dz, z.p, lng.p, s0, dXyz});
    }
};

struct SphereTriangulator {
    DevBuf<uint32_t> key;
    DevBuf<int> sid, deg, fail, rows, pos;
    DevBuf<double> sx, sy, sz;
    DevBuf<uint8_t> scanTemp, retried;
    Prims prims;

    static GridSpec grid_for(int n) {
        // cell ≈ 2.5 mean spacings, at most 1024 cells per axis (30-bit keys)
        const double spacing = sqrt(4.0 * 3.14159265358979323846 / (double)n);
        const char* e = getenv("PB_MESH_CELL");             // tuning knob: cell size in mean spacings
        const double mult = e ? atof(e) : 2.5;
        int G = (int)floor(2.0 / (mult * spacing));
        if (G < 1) G = 1;
        if (G > 1024) G = 1024;
        return GridSpec{G, 0.5 * (double)G, 2.0 / (double)G};
    }

    void exclusive_scan(const Exec& ex, const int* in, int* out, int count) {
        launch_stats().launches++;
        ProfScope ps(ex.prof, "cub::DeviceScan::ExclusiveSum", ex.stream);
#if PB_CUDA
        size_t bytes = 0;
        PB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, count, ex.stream));
        scanTemp.ensure(bytes);
        PB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(scanTemp.p, bytes, in, out, count, ex.stream));
#else
        int acc = 0;
        for (int i = 0; i < count; i++) { const int v = in[i]; out[i] = acc; acc += v; }
#endif
    }

    // dXyz: device f32[3n].  dOff: device int[n+1], dAdj: device int[6n-12].  Throws when the result is not a closed
    // triangulated sphere.
    void build(const Exec& ex, int n, const float* dXyz, int* dOff, int* dAdj) {
        if (n < 4) throw Error("a sphere mesh needs at least 4 points");
        const GridSpec g = grid_for(n);
        key.ensure(n); sid.ensure(n); deg.ensure((size_t)n + 1); fail.ensure(4);
        sx.ensure(n); sy.ensure(n); sz.ensure(n);
        dev_memset(fail.p, 0, 4 * sizeof(int), ex.stream);
        ex.for_each(n, MeshGenKeyK{dXyz, g, key.p, sid.p});
        prims.sort_pairs(ex, key.p, sid.p, n, false, 30);
        ex.for_each(n, MeshGenGatherK{dXyz, sid.p, sx.p, sy.p, sz.p});
        dev_memset(deg.p + n, 0, sizeof(int), ex.stream);
        rows.ensure((size_t)n * kRowCache); pos.ensure(n);
        ex.for_each(n, MeshGenInvertK{sid.p, pos.p});
        ex.for_each(n, StarK{n, g, key.p, sid.p, sx.p, sy.p, sz.p, deg.p, rows.p, nullptr, nullptr, fail.p});
        exclusive_scan(ex, deg.p, dOff, n + 1);
        int h[4] = {0, 0, 0, 0}, total = 0;
        dev_copy(h, fail.p, 2 * sizeof(int), 1, ex.stream);
        dev_copy(&total, dOff + n, sizeof(int), 1, ex.stream);
        stream_sync(ex.stream);
        bool anyRetried = false;
        if (h[0] && !getenv("PB_MESH_NO_RETRY")) {     // some blocks never got large enough: grow them up to the whole grid
            anyRetried = true;
            dev_memset(retried.ensure(n), 0, (size_t)n, ex.stream);
            dev_memset(fail.p, 0, sizeof(int), ex.stream);
            ex.for_each(n, StarRetryK{StarK{n, g, key.p, sid.p, sx.p, sy.p, sz.p, deg.p, rows.p, nullptr, nullptr, fail.p}, pos.p, retried.p, 0});
            exclusive_scan(ex, deg.p, dOff, n + 1);
            dev_copy(h, fail.p, 2 * sizeof(int), 1, ex.stream);
            dev_copy(&total, dOff + n, sizeof(int), 1, ex.stream);
            stream_sync(ex.stream);
        }
        if (h[0]) throw Error("spherical Delaunay: " + std::to_string(h[0]) + " region stars could not be closed (duplicate points or a region with more than 32 neighbours)");
        if (total != 6 * n - 12) throw Error("spherical Delaunay: edge count " + std::to_string(total) + " != 6n-12 (degenerate point set)");
        ex.for_each(n, StarFillK{StarK{n, g, key.p, sid.p, sx.p, sy.p, sz.p, deg.p, rows.p, dOff, dAdj, fail.p}, pos.p});
        if (anyRetried)
            ex.for_each(n, StarRetryK{StarK{n, g, key.p, sid.p, sx.p, sy.p, sz.p, deg.p, rows.p, dOff, dAdj, fail.p}, pos.p, retried.p, 1});
        ex.for_each(n, MeshGenSymmetryK{n, dOff, dAdj, fail.p + 2});
        dev_copy(h, fail.p, 3 * sizeof(int), 1, ex.stream);
        stream_sync(ex.stream);
        if (h[2]) throw Error("spherical Delaunay: " + std::to_string(h[2]) + " edges have no twin (inconsistent predicates)");
    }
};

}  // namespace pb

// ---- triangles / half-edges of a mesh from its CSR rows (SphereMesh.triangles / .halfedges, js/sphere-mesh.js:94-100) ----
// A row lists a region's neighbours clockwise (seen from outside), so the counter-clockwise triangle between two
// consecutive entries is (r, row[k+1], row[k]).  In the canonical numbering every triangle starts at its smallest
// vertex and triangles are in lexicographic order: region r owns the triangles in which it is the smallest vertex, sorted
// by (second, third) — a count pass, a scan and a fill pass give `triangles`; the twin of side a→b is found in the block
// of the smallest vertex of the triangle on the other side of the edge.
namespace pb {

struct TriCountK {
    Csr g; int* cnt;
    PB_DEV void operator()(int r) const {
        const int s = g.off[r], d = g.off[r + 1] - s;
        int c = 0;
        for (int k = 0; k < d; k++) { const int b = g.adj[s + (k + 1 == d ? 0 : k + 1)], cc = g.adj[s + k]; if (r < b && r < cc) c++; }
        cnt[r] = c;
    }
};
struct TriFillK {
    Csr g; const int* start; int* tri;
    PB_DEV void operator()(int r) const {
        const int s = g.off[r], d = g.off[r + 1] - s;
        int bs[32], cs[32], n = 0;
        for (int k = 0; k < d; k++) {
            const int b = g.adj[s + (k + 1 == d ? 0 : k + 1)], c = g.adj[s + k];
            if (!(r < b && r < c)) continue;
            int at = n++;
            while (at > 0 && (bs[at - 1] > b || (bs[at - 1] == b && cs[at - 1] > c))) { bs[at] = bs[at - 1]; cs[at] = cs[at - 1]; at--; }
            bs[at] = b; cs[at] = c;
        }
        int* out = tri + 3 * (size_t)start[r];
        for (int i = 0; i < n; i++) { out[3 * i] = r; out[3 * i + 1] = bs[i]; out[3 * i + 2] = cs[i]; }
    }
};
struct HalfedgeK {
    Csr g; const int* start; const int* tri; int* half; int* bad;
    PB_DEV void operator()(int side) const {
        const int t = side / 3, e = side - 3 * t;
        const int a = tri[3 * t + e], b = tri[3 * t + (e == 2 ? 0 : e + 1)];
        // the third vertex of the triangle on the other side of a→b: the entry after b in a's (clockwise) row
        const int s = g.off[a], d = g.off[a + 1] - s;
        int p = -1;
        for (int k = 0; k < d; k++) if (g.adj[s + k] == b) { p = k; break; }
        if (p < 0) { atomic_add(bad, 1); half[side] = -1; return; }
        const int x = g.adj[s + (p + 1 == d ? 0 : p + 1)];
        // that triangle is (a, x, b) counter-clockwise; rotate to its smallest vertex and look it up in that vertex's block
        int v0 = a, v1 = x, v2 = b;
        if (x < a && x < b) { v0 = x; v1 = b; v2 = a; } else if (b < a && b < x) { v0 = b; v1 = a; v2 = x; }
        int found = -1;
        for (int u = start[v0]; u < start[v0 + 1]; u++) if (tri[3 * u + 1] == v1 && tri[3 * u + 2] == v2) { found = u; break; }
        if (found < 0) { atomic_add(bad, 1); half[side] = -1; return; }
        const int pos = v0 == b ? 0 : (v1 == b ? 1 : 2);          // the twin side starts at b
        half[side] = 3 * found + pos;
    }
};
// _adjTriList (js/sphere-mesh.js:128-143): for the k-th neighbour b of region a, the triangle that holds side a→b, i.e.
// (a, b, c) counter-clockwise with c = the entry before b in a's clockwise row
struct AdjTriK {
    Csr g; const int* start; const int* tri; int* adjTri; int* bad;
    PB_DEV void operator()(int a) const {
        const int s = g.off[a], d = g.off[a + 1] - s;
        for (int k = 0; k < d; k++) {
            const int b = g.adj[s + k], c = g.adj[s + (k == 0 ? d - 1 : k - 1)];
            int v0 = a, v1 = b, v2 = c;
            if (b < a && b < c) { v0 = b; v1 = c; v2 = a; } else if (c < a && c < b) { v0 = c; v1 = a; v2 = b; }
            int found = -1;
            for (int u = start[v0]; u < start[v0 + 1]; u++) if (tri[3 * u + 1] == v1 && tri[3 * u + 2] == v2) { found = u; break; }
            if (found < 0) atomic_add(bad, 1);
            adjTri[s + k] = found;
        }
    }
};
// generateTriangleCenters (js/sphere-mesh.js:206-219) and computeTriangleElevations (js/planet-worker.js:29-37)
struct TriCentersK {
    const int* tri; const float* xyz; float* t_xyz;
    PB_DEV void operator()(int t) const {
        const int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
        for (int k = 0; k < 3; k++) t_xyz[3 * t + k] = (float)((((double)xyz[3 * a + k] + (double)xyz[3 * b + k]) + (double)xyz[3 * c + k]) / 3);
    }
};
struct TriElevationK {
    const int* tri; const float* elev; float* t_elev;
    PB_DEV void operator()(int t) const {
        t_elev[t] = (float)((((double)elev[tri[3 * t]] + (double)elev[tri[3 * t + 1]]) + (double)elev[tri[3 * t + 2]]) / 3);
    }
};

struct MeshTriangles {
    DevBuf<int> cnt, start, tri, half, flag, adjTri;
    DevBuf<uint8_t> scanTemp;
    int T = 0;
    // builds (once per mesh) the device copies; tri / half are 3T ints
    void build(const Exec& ex, Csr g, int N) {
        if (T) return;
        cnt.ensure((size_t)N + 1); start.ensure((size_t)N + 1); flag.ensure(2);
        dev_memset(cnt.p + N, 0, sizeof(int), ex.stream);
        dev_memset(flag.p, 0, 2 * sizeof(int), ex.stream);
        ex.for_each(N, TriCountK{g, cnt.p});
        launch_stats().launches++;
#if PB_CUDA
        size_t bytes = 0;
        PB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.p, start.p, N + 1, ex.stream));
        scanTemp.ensure(bytes);
        PB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(scanTemp.p, bytes, cnt.p, start.p, N + 1, ex.stream));
#else
        { int acc = 0; for (int i = 0; i <= N; i++) { const int v = cnt.p[i]; start.p[i] = acc; acc += v; } }
#endif
        int total = 0;
        dev_copy(&total, start.p + N, sizeof(int), 1, ex.stream);
        stream_sync(ex.stream);
        if (total != 2 * N - 4) throw Error("mesh is not a closed triangulated sphere: " + std::to_string(total) + " triangles for " + std::to_string(N) + " regions");
        tri.ensure(3 * (size_t)total); half.ensure(3 * (size_t)total);
        ex.for_each(N, TriFillK{g, start.p, tri.p});
        ex.for_each(3 * total, HalfedgeK{g, start.p, tri.p, half.p, flag.p});
        int bad = 0;
        dev_copy(&bad, flag.p, sizeof(int), 1, ex.stream);
        stream_sync(ex.stream);
        if (bad) throw Error("mesh rows are not consistent circulations: " + std::to_string(bad) + " sides without a twin");
        adjTri.ensure((size_t)3 * total);            // E = 3T directed edges
        ex.for_each(N, AdjTriK{g, start.p, tri.p, adjTri.p, flag.p + 1});
        dev_copy(&bad, flag.p + 1, sizeof(int), 1, ex.stream);
        stream_sync(ex.stream);
        if (bad) throw Error("mesh rows are not consistent circulations: " + std::to_string(bad) + " edges without an inner triangle");
        T = total;
    }
};

}  // namespace pb
