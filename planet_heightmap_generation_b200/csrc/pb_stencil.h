// pb_stencil.h — class-P per-cell passes over the CSR neighbour graph (kernel families K1, K3, K7):
// Jacobi / masked Jacobi / bilateral / sharpen / creep sweeps, neighbour distances, domain warp.
// Every pass loads f32, computes in FP64 with the reference's operation order and stores f32, which
// makes it bit-identical to the JS typed-array semantics.
#pragma once
#include "pb_platform.h"
#include "pb_noise.h"

namespace pb {

// One 16-byte word per row next to the CSR: up to eight neighbours as signed 16-bit id deltas (nb - r), in adjacency order,
// zero-padded (a cell is never its own neighbour).  Fibonacci ids advance in z, so every neighbour of a cell lies within
// ±5·√N ids — 16 bits are enough up to tens of millions of cells; rows that do not fit (more than eight
// neighbours, a larger delta: the pole vertex N-1 touches the lowest ids) carry the escape mark and are read from the CSR.
// A sweep thread then fetches its whole row with ONE coalesced 128-bit load instead of two offsets + six to eight strided
// 32-bit loads: 16 B per row instead of 28 B, and an order of magnitude fewer L1 wavefronts (the sweeps are LSU-bound,
// profiles/r01_sweeps_ncu.md).
struct alignas(16) PackedRow { uint32_t w[4]; };
#define PB_PACK_ESCAPE 0x7fffu

struct Csr {
    int N;
    const int* off;   // [N+1]
    const int* adj;   // [E]
    const PackedRow* pack;   // [N] or nullptr
};

struct PackRowsK {
    Csr g; PackedRow* out;
    PB_DEV void operator()(int r) const {
        const int b = g.off[r], deg = g.off[r + 1] - b;
        PackedRow p; p.w[0] = PB_PACK_ESCAPE; p.w[1] = p.w[2] = p.w[3] = 0;
        bool ok = deg <= 8;
        uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < deg && ok; k++) {
            const int d = g.adj[b + k] - r;
            if (d == 0 || d > 32766 || d < -32767) ok = false;
            else h[k] = (uint32_t)d & 0xffffu;
        }
        if (ok) for (int k = 0; k < 4; k++) p.w[k] = h[2 * k] | (h[2 * k + 1] << 16);
        out[r] = p;
    }
};

PB_DEV double or_default(double a, double b) { return (a == 0.0 || a != a) ? b : a; }  // JS `a || b`

// js/sphere-mesh.js:191-203 — chord length per directed edge
struct NeighborDistK {
    Csr g; const float* xyz; float* out;
    PB_DEV void operator()(int r) const {
        const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
        for (int i = g.off[r], e = g.off[r + 1]; i < e; i++) {
            const int nb = g.adj[i];
            const double dx = x - xyz[3 * nb], dy = y - xyz[3 * nb + 1], dz = z - xyz[3 * nb + 2];
            out[i] = (float)sqrt(dx * dx + dy * dy + dz * dz);
        }
    }
};

// Row gather with memory-level parallelism.  A sweep thread that walks its row as `sum += src[adj[i]]` serialises
// two dependent load latencies per neighbour (≈ 12 per cell): at full occupancy that, not bandwidth, bounds the
// sweep (profiles/r01_sweeps_ncu.md).  The helpers below issue all neighbour-id loads of a row at once and then all
// value gathers at once (rows of up to PB_ROW_FAST neighbours — every row of a spherical Delaunay mesh in practice;
// longer rows take the plain loop), so a cell costs two load latencies.  The additions still run in adjacency
// order, which keeps the f64 sums bit-identical.
#define PB_ROW_FAST 8
struct RowIds {
    int nb[PB_ROW_FAST];
    PB_DEV void load(int r, int deg, const int* ids) {       // deg <= PB_ROW_FAST
#pragma unroll
        for (int k = 0; k < PB_ROW_FAST; k++) nb[k] = k < deg ? ids[k] : r;
    }
    // whole row from the packed word; false: escape row (take the CSR path)
    PB_DEV bool load_packed(const PackedRow* pack, int r, int& deg) { return load_word(pack + r, r, deg); }
    // the packed word of row r stored at `word` (compacted row lists keep their own array of words)
    PB_DEV bool load_word(const PackedRow* word, int r, int& deg) {
#if PB_CUDA
        const uint4 q = __ldg((const uint4*)word);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#else
        const uint32_t* w = word->w;
#endif
        if ((w[0] & 0xffffu) == PB_PACK_ESCAPE) return false;
        int n = 0;
#pragma unroll
        for (int k = 0; k < PB_ROW_FAST; k++) {
            const int d = (int)(int16_t)(uint16_t)(w[k >> 1] >> (16 * (k & 1)));
            nb[k] = r + d;
            n += d != 0;
        }
        deg = n;
        return true;
    }
};

// js/climate-util.js:5-25 — one Laplacian sweep  dst = (src[r] + Σ src[nb]) / (deg + 1)
struct SmoothFieldK {
    Csr g; const float* src; float* dst;
    PB_DEV void operator()(int r) const {
        if (g.pack) {
            RowIds row; int deg;
            if (row.load_packed(g.pack, r, deg)) { gathered(r, deg, row); return; }
        }
        const int b = g.off[r];
        this->row(r, b, g.off[r + 1] - b, g.adj + b);
    }
    PB_DEV void gathered(int r, int deg, const RowIds& row) const {
        double sum = src[r];
        float v[PB_ROW_FAST];
#pragma unroll
        for (int k = 0; k < PB_ROW_FAST; k++) v[k] = src[row.nb[k]];
#pragma unroll
        for (int k = 0; k < PB_ROW_FAST; k++) if (k < deg) sum += v[k];
        dst[r] = (float)(sum / (deg + 1));
    }
    PB_DEV void row(int r, int b, int deg, const int* ids) const {
        (void)b;
        if (deg <= PB_ROW_FAST) { RowIds rw; rw.load(r, deg, ids); gathered(r, deg, rw); return; }
        double sum = src[r];
        for (int i = 0; i < deg; i++) sum += src[ids[i]];
        dst[r] = (float)(sum / (deg + 1));
    }
};

// any row functor over a compacted row list: the lanes of a warp all carry work (rows outside the list are not visited)
// Item functors (kItems): operator()(i) handles item i of a compacted list, row_of(i) is its row, by_row(r) handles row r when
// the caller only knows the row (the boundary rows of a sharded sweep).
template <class F>
struct OverRowsK {
    static constexpr bool kItems = true;
    const int* rows; F f;
    PB_DEV void operator()(int i) const { f(rows[i]); }
    PB_DEV int row_of(int i) const { return rows[i]; }
    PB_DEV void by_row(int r) const { f(r); }
};

// js/planet-worker.js:51-54
struct IsOceanK {
    const float* elev; uint8_t* isOcean;
    PB_DEV void operator()(int r) const { isOcean[r] = elev[r] <= 0 ? 1 : 0; }
};

// cell classes used by smoothElevation (:323-329) and applySoilCreep (:763-771):
// 0 ocean, 1 coastal land (has an ocean neighbour), 2 interior land
struct CellClassK {
    Csr g; const uint8_t* isOcean; uint8_t* cls;
    PB_DEV void operator()(int r) const {
        if (isOcean[r]) { cls[r] = 0; return; }
        uint8_t c = 2;
        for (int i = g.off[r], e = g.off[r + 1]; i < e; i++)
            if (isOcean[g.adj[i]]) { c = 1; break; }
        cls[r] = c;
    }
};

// js/terrain-post.js:331-353 — bilateral sweep; coastal land is locked, ocean cells DO move
struct BilateralK {
    Csr g; const float* src; float* dst; const uint8_t* cls; double strength;
    PB_DEV void operator()(int r) const {
        if (cls[r] == 1) { dst[r] = src[r]; return; }
        const double h = src[r];
        double wSum = 0, hSum = 0;
        for (int i = g.off[r], e = g.off[r + 1]; i < e; i++) {
            const double nh = src[g.adj[i]];
            const double w = 1.0 / (1.0 + fabs(nh - h) * 8.0);
            wSum += w;
            hSum += nh * w;
        }
        if (wSum > 0) {
            const double avg = hSum / wSum;
            dst[r] = (float)(h + (avg - h) * strength);
        } else dst[r] = (float)h;
    }
};

// js/terrain-post.js:727-750
struct SharpenK {
    Csr g; const float* src; float* dst; const float* original; const uint8_t* isOcean; double strength;
    PB_DEV void operator()(int r) const {
        const float hf = src[r];
        if (isOcean[r]) { dst[r] = hf; return; }
        const double h = hf;
        double sum = 0;
        const int b = g.off[r], e = g.off[r + 1];
        for (int i = b; i < e; i++) sum += src[g.adj[i]];
        if (e == b) { dst[r] = hf; return; }
        const double avg = sum / (e - b);
        if (h > avg) {
            double hn = h + (h - avg) * strength;
            const double cap = (double)original[r] * 1.5;
            if (hn > cap) hn = cap;
            dst[r] = (float)hn;
        } else dst[r] = hf;
    }
};

// js/terrain-post.js:776-793 — Laplacian over land neighbours, interior land only
struct CreepK {
    Csr g; const float* src; float* dst; const uint8_t* cls; double strength;
    PB_DEV void operator()(int r) const {
        const float hf = src[r];
        if (cls[r] != 2) { dst[r] = hf; return; }
        const double h = hf;
        double sum = 0; int count = 0;
        for (int i = g.off[r], e = g.off[r + 1]; i < e; i++) {
            const int nb = g.adj[i];
            if (cls[nb] != 0) { sum += src[nb]; count++; }
        }
        if (count == 0) { dst[r] = hf; return; }
        const double avg = sum / count;
        dst[r] = (float)(h + (avg - h) * strength);
    }
};

// js/terrain-post.js:690-706 — one masked Jacobi step (0.3) on glaciated land
struct GlacialBlendK {
    Csr g; const float* src; float* dst; const uint8_t* isOcean; const float* glacIdx;
    PB_DEV void operator()(int r) const {
        const float hf = src[r];
        if (isOcean[r] || !(glacIdx[r] > 0)) { dst[r] = hf; return; }
        double sum = 0; int count = 0;
        for (int i = g.off[r], e = g.off[r + 1]; i < e; i++) {
            const int nb = g.adj[i];
            if (!isOcean[nb]) { sum += src[nb]; count++; }
        }
        if (count > 0) {
            const double avg = sum / count;
            dst[r] = (float)((double)hf + (avg - (double)hf) * 0.3);
        } else dst[r] = hf;
    }
};

// js/terrain-post.js:245-289 — FBM displacement in the tangent plane, then a greedy walk over the
// CSR graph towards the displaced point.  The inner loop bounds are those of the cell the outer
// iteration STARTED on while `cur` moves (:276-284) — kept on purpose.
struct WarpWalkK {
    Csr g; const float* xyz; const float* elev; float* out; Simplex noise; double maxAmp;
    PB_DEV void operator()(int r) const {
        const double px = xyz[3 * r], py = xyz[3 * r + 1], pz = xyz[3 * r + 2];
        double ex = -pz, ez = px;
        const double ey = 0;
        const double elen = sqrt(ex * ex + ez * ez);
        if (elen > 1e-10) { ex /= elen; ez /= elen; } else { ex = 1; ez = 0; }
        const double nx = py * ez, ny = pz * ex - px * ez, nz = -py * ex;
        const double nlen = or_default(sqrt(nx * nx + ny * ny + nz * nz), 1.0);
        const double nnx = nx / nlen, nny = ny / nlen, nnz = nz / nlen;
        const double freq = 4;
        const double d1 = noise.fbm(px * freq, py * freq, pz * freq, 5) * maxAmp;
        const double d2 = noise.fbm(px * freq + 31.7, py * freq + 47.3, pz * freq + 19.1, 5) * maxAmp;
        double wx = px + ex * d1 + nnx * d2;
        double wy = py + ey * d1 + nny * d2;
        double wz = pz + ez * d1 + nnz * d2;
        const double wlen = or_default(sqrt(wx * wx + wy * wy + wz * wz), 1.0);
        wx /= wlen; wy /= wlen; wz /= wlen;
        int cur = r;
        double best = wx * px + wy * py + wz * pz;
        for (;;) {
            bool moved = false;
            for (int i = g.off[cur], e = g.off[cur + 1]; i < e; i++) {
                const int nb = g.adj[i];
                const double dot = wx * xyz[3 * nb] + wy * xyz[3 * nb + 1] + wz * xyz[3 * nb + 2];
                if (dot > best) { best = dot; cur = nb; moved = true; }
            }
            if (!moved) break;
        }
        out[r] = elev[cur];
    }
};

// js/terrain-post.js:294-308
struct WarpBlendK {
    float* elev; const float* warped; const float* hotspot; double warpBias;
    PB_DEV void operator()(int r) const {
        const double orig = elev[r];
        const double w = warped[r];
        double bias = warpBias;
        if (hotspot) {
            double hf = fabs((double)hotspot[r]) / or_default(fabs(orig), 1.0);
            if (hf > 1.0) hf = 1.0;    // Math.min(1, x): NaN cannot occur (denominator is never 0)
            bias *= 1.0 - 0.8 * hf;
        }
        if (w > orig) elev[r] = (float)(orig + (w - orig) * bias);
        else elev[r] = (float)(w + (orig - w) * (1.0 - bias));
    }
};

struct CopyF32K {
    const float* src; float* dst;
    PB_DEV void operator()(int i) const { dst[i] = src[i]; }
};

struct SubF32K {  // js/planet-worker.js:96-99  erosionDelta = after − before
    const float* a; const float* b; float* out;
    PB_DEV void operator()(int i) const { out[i] = (float)((double)a[i] - (double)b[i]); }
};

}  // namespace pb
