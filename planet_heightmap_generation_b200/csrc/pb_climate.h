// pb_climate.h — per-cell kernels of the climate stack (kernel families K1, K2, K3, K6, K8, K15):
//   js/wind.js, js/ocean.js, js/precipitation.js, js/heuristic-precip.js, js/temperature.js,
//   js/koppen.js, js/climate-util.js, js/color-map.js:7-12.
// Every pass loads f32, computes in FP64 in the reference's operation order and stores f32, so the
// results are bit-identical to the JS typed-array semantics; transcendentals go through
// include/pb_detmath.h.  Hop-count BFS fields are level-synchronous frontier expansions (hop counts
// do not depend on queue order).  The only order-dependent reductions — the ITCZ cap samples — are
// accumulated in the reference's (lat bin, lon bin, cell id) order.
#pragma once
#include "pb_platform.h"
#include "pb_stencil.h"
#include "pb_noise.h"
#include "pb_prims.h"
#include "pb_flood.h"

namespace pb {

#define PB_DEG (PB_PI / 180)
#define PB_RAD (180 / PB_PI)

// Math.min / Math.max (NaN-propagating; the signed-zero rule cannot matter for the uses below)
PB_DEV double jmin(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
PB_DEV double jmax(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }

PB_DEV double elev_to_height_km(double elev) {          // js/color-map.js:7-12
    if (elev <= 0) return elev * 10;
    const double t = jmin(elev, 1);
    const double t2 = t * t;
    return 6 * t2 * t2 * (5 - 4 * t);
}
PB_DEV double smoothstep(double e0, double e1, double x) {   // js/wind.js:76-80
    if (e0 == e1) return x >= e1 ? 1 : 0;
    const double t = jmax(0, jmin(1, (x - e0) / (e1 - e0)));
    return t * t * (3 - 2 * t);
}

// ---- ITCZ spline (js/wind.js:11-72) and the 360-sample lookup (js/climate-util.js:29-43) -----------
#define PB_ITCZ_NLON 72
#define PB_ITCZ_SAMPLES 360
struct SplineDev { double xs[PB_ITCZ_NLON], ys[PB_ITCZ_NLON], b[PB_ITCZ_NLON], c[PB_ITCZ_NLON], d[PB_ITCZ_NLON]; };

PB_DEV double evaluate_spline(const SplineDev* sp, double lon) {
    const int n = PB_ITCZ_NLON;
    const double period = 2 * PB_PI;
    const double x0 = sp->xs[0];
    const double t = fmod(fmod(lon - x0, period) + period, period) + x0;
    int seg = 0;
    for (int i = 0; i < n; i++) {
        const double lo = sp->xs[i];
        const double hi = i < n - 1 ? sp->xs[i + 1] : x0 + period;
        if (t >= lo && t < hi) { seg = i; break; }
    }
    const double dx = t - sp->xs[seg];
    return sp->ys[seg] + sp->b[seg] * dx + sp->c[seg] * dx * dx + sp->d[seg] * dx * dx * dx;
}
PB_DEV double itcz_lookup(const float* lats, double lon) {
    const int n = PB_ITCZ_SAMPLES;
    const double step = (2 * PB_PI) / n;
    const double lonStart = -PB_PI + step * 0.5;
    double fi = (lon - lonStart) / step;
    fi = fmod(fmod(fi, (double)n) + n, (double)n);
    const double i0 = floor(fi);
    const int i1 = ((int)i0 + 1) % n;
    const double frac = fi - i0;
    return (double)lats[(int)i0] * (1 - frac) + (double)lats[i1] * frac;
}

// ---- wind.js step 0 (:418-443) -------------------------------------------------------------------------
struct WindPrecomputeK {
    const float* xyz; const float* elev;
    float *lat, *lon, *sinLat, *cosLat; uint8_t* isLand; uint8_t* isOcean;
    float *eX, *eY, *eZ, *nX, *nY, *nZ;
    PB_DEV void operator()(int r) const {
        const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
        lat[r] = (float)pb_asin(jmax(-1, jmin(1, y)));
        lon[r] = (float)pb_atan2(x, z);
        sinLat[r] = (float)y;
        cosLat[r] = (float)or_default(sqrt(1 - y * y), 0.01);
        const uint8_t land = elev[r] > 0 ? 1 : 0;
        isLand[r] = land; isOcean[r] = land ? 0 : 1;
        double ex = z, ey = 0, ez = -x;
        double elen = sqrt(ex * ex + ez * ez);
        if (elen < 1e-10) { ex = 1; ez = 0; elen = 1; }
        ex /= elen; ez /= elen;
        double nx = y * ez - z * ey;
        double ny = z * ex - x * ez;
        double nz = x * ey - y * ex;
        const double nlen = or_default(sqrt(nx * nx + ny * ny + nz * nz), 1.0);
        nx /= nlen; ny /= nlen; nz /= nlen;
        eX[r] = (float)ex; eY[r] = (float)ey; eZ[r] = (float)ez;
        nX[r] = (float)nx; nY[r] = (float)ny; nZ[r] = (float)nz;
    }
};

// ---- geo index (:88-119): bin id per cell; the stable sort by bin reproduces the counting sort ------
#define PB_LAT_BINS 36
#define PB_LON_BINS 72
struct GeoBinK {
    const float* lat; const float* lon; uint32_t* bin; int* cell;
    PB_DEV void operator()(int r) const {
        const int latBin = (int)jmax(0, jmin(PB_LAT_BINS - 1, floor(((double)lat[r] + PB_PI / 2) / PB_PI * PB_LAT_BINS)));
        const int lonBin = (int)jmax(0, jmin(PB_LON_BINS - 1, floor(((double)lon[r] + PB_PI) / (2 * PB_PI) * PB_LON_BINS)));
        bin[r] = (uint32_t)(latBin * PB_LON_BINS + lonBin);
        cell[r] = r;
    }
};
struct BinOffsetK {   // binOffset[b] = first position whose bin >= b   (b in 0..numBins)
    const uint32_t* sortedBin; int n; int* binOffset;
    PB_DEV void operator()(int b) const {
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sortedBin[mid] < (uint32_t)b) lo = mid + 1; else hi = mid; }
        binOffset[b] = lo;
    }
};

// one ITCZ cap sample (:127-164): sample s = ((season*72 + i)*4 + d); out[2s] = landFrac, out[2s+1] = avgElev
struct ItczSampleGeom {
    double lat, lon, radius, cosRadius, sinLat0, cosLat0;
    int bMin, bMax, lMin, lMax;
};
PB_DEV ItczSampleGeom itcz_sample_geom(int s) {
    ItczSampleGeom g;
    const int d = s & 3, i = (s >> 2) % PB_ITCZ_NLON, season = (s >> 2) / PB_ITCZ_NLON;
    const double sign = season == 0 ? 1 : -1;
    g.lon = -PB_PI + (i + 0.5) * (2 * PB_PI / PB_ITCZ_NLON);
    g.lat = (5 + 5 * d) * sign * PB_DEG;
    g.radius = 20 * PB_DEG;
    const double latMin = g.lat - g.radius, latMax = g.lat + g.radius;
    g.bMin = (int)jmax(0, floor((latMin + PB_PI / 2) / PB_PI * PB_LAT_BINS));
    g.bMax = (int)jmin(PB_LAT_BINS - 1, floor((latMax + PB_PI / 2) / PB_PI * PB_LAT_BINS));
    const double cosLat = or_default(pb_cos(g.lat), 0.01);
    const double lonSpan = g.radius / cosLat;
    g.lMin = (int)floor((g.lon - lonSpan + PB_PI) / (2 * PB_PI) * PB_LON_BINS);
    g.lMax = (int)floor((g.lon + lonSpan + PB_PI) / (2 * PB_PI) * PB_LON_BINS);
    g.cosRadius = pb_cos(g.radius);
    g.sinLat0 = pb_sin(g.lat); g.cosLat0 = pb_cos(g.lat);
    return g;
}
struct ItczSampleArgs {
    const int* binOffset; const int* cells; const float* lon; const float* sinLat; const float* cosLat;
    const float* elev; const uint8_t* isLand; double* out;
};
// per-cell part of the sample: returns pass flag and the value added to elevSum
PB_DEV bool itcz_cell(const ItczSampleArgs& a, const ItczSampleGeom& g, int r, double* val, bool* land) {
    const double dlon = (double)a.lon[r] - g.lon;
    const double cosDist = g.sinLat0 * (double)a.sinLat[r] + g.cosLat0 * (double)a.cosLat[r] * pb_cos(dlon);
    if (!(cosDist >= g.cosRadius)) return false;
    *land = a.isLand[r] != 0;
    *val = jmax(0, (double)a.elev[r]);
    return true;
}
#if !PB_CUDA
struct ItczSampleSerialK {   // emulation: one logical thread per sample
    ItczSampleArgs a;
    void operator()(int s) const {
        const ItczSampleGeom g = itcz_sample_geom(s);
        double landCount = 0, totalCount = 0, elevSum = 0;
        for (int bi = g.bMin; bi <= g.bMax; bi++)
            for (int li = g.lMin; li <= g.lMax; li++) {
                const int lj = ((li % PB_LON_BINS) + PB_LON_BINS) % PB_LON_BINS;
                const int bin = bi * PB_LON_BINS + lj;
                for (int k = a.binOffset[bin]; k < a.binOffset[bin + 1]; k++) {
                    double v; bool land;
                    if (itcz_cell(a, g, a.cells[k], &v, &land)) { totalCount++; if (land) landCount++; elevSum += v; }
                }
            }
        a.out[2 * s] = totalCount == 0 ? 0 : landCount / totalCount;
        a.out[2 * s + 1] = totalCount == 0 ? 0 : elevSum / totalCount;
    }
};
#else
// CUDA: one CTA per sample.  Threads evaluate the cap predicate for 256 cells at a time; the f64
// elevation sum is order-dependent, so thread 0 accumulates the passing non-zero terms in order
// (adding +0.0 to a non-negative sum is the identity, so zero terms are skipped).
#define PB_ITCZ_THREADS 256
__global__ void __launch_bounds__(PB_ITCZ_THREADS) k_itcz_sample(ItczSampleArgs a) {
    const int s = blockIdx.x;
    const ItczSampleGeom g = itcz_sample_geom(s);
    __shared__ double sVal[PB_ITCZ_THREADS];
    __shared__ unsigned sMask[PB_ITCZ_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int landCount = 0, totalCount = 0;
    double elevSum = 0;
    for (int bi = g.bMin; bi <= g.bMax; bi++)
        for (int li = g.lMin; li <= g.lMax; li++) {
            const int lj = ((li % PB_LON_BINS) + PB_LON_BINS) % PB_LON_BINS;
            const int bin = bi * PB_LON_BINS + lj;
            const int kb = a.binOffset[bin], ke = a.binOffset[bin + 1];
            for (int base = kb; base < ke; base += PB_ITCZ_THREADS) {
                const int k = base + tid;
                bool pass = false, land = false;
                double v = 0;
                if (k < ke) pass = itcz_cell(a, g, a.cells[k], &v, &land);
                totalCount += __syncthreads_count(pass);
                landCount += __syncthreads_count(pass && land);
                const unsigned m = __ballot_sync(0xffffffffu, pass && v > 0);
                sVal[tid] = v;
                if (lane == 0) sMask[warp] = m;
                __syncthreads();
                if (tid == 0) {
                    for (int w = 0; w < PB_ITCZ_THREADS / 32; w++) {
                        unsigned mm = sMask[w];
                        while (mm) { const int q = __ffs(mm) - 1; mm &= mm - 1; elevSum += sVal[w * 32 + q]; }
                    }
                }
                __syncthreads();
            }
        }
    if (tid == 0) {
        a.out[2 * s] = totalCount == 0 ? 0 : (double)landCount / (double)totalCount;
        a.out[2 * s + 1] = totalCount == 0 ? 0 : elevSum / (double)totalCount;
    }
}
#endif

// computeITCZ (:174-232) after the samples + buildPeriodicSpline (:11-55): 72-point host-sized work,
// kept on the device so the stage needs no host round trip.  One logical thread per season.
struct ItczFinishK {
    const double* samples; SplineDev* splines;
    PB_DEV void operator()(int season) const {
        const int NUM_LON = PB_ITCZ_NLON;
        const double sign = season == 0 ? 1 : -1;
        SplineDev* sp = splines + season;
        double lats[PB_ITCZ_NLON], tmp[PB_ITCZ_NLON], h[PB_ITCZ_NLON], alpha[PB_ITCZ_NLON];
        for (int i = 0; i < NUM_LON; i++) {
            sp->xs[i] = -PB_PI + (i + 0.5) * (2 * PB_PI / NUM_LON);
            double landSum = 0, elevSum = 0, n = 0;
            for (int d = 0; d < 4; d++) {
                const int s = (season * NUM_LON + i) * 4 + d;
                landSum += samples[2 * s]; elevSum += samples[2 * s + 1]; n++;
            }
            const double avgLand = landSum / n, avgElev = elevSum / n;
            const double landPull = jmin(1, avgLand * 2);
            const double itczDeg = 5 + landPull * 15 - elev_to_height_km(avgElev) * 1.5;
            const double clampedDeg = jmax(5, jmin(20, itczDeg));
            lats[i] = clampedDeg * sign * PB_DEG;
        }
        for (int pass = 0; pass < 3; pass++) {
            for (int i = 0; i < NUM_LON; i++) {
                const int p = (i - 1 + NUM_LON) % NUM_LON, n = (i + 1) % NUM_LON;
                tmp[i] = 0.25 * lats[p] + 0.5 * lats[i] + 0.25 * lats[n];
            }
            for (int i = 0; i < NUM_LON; i++) lats[i] = tmp[i];
        }
        const double clampMin = (sign > 0 ? 5 : -20) * PB_DEG, clampMax = (sign > 0 ? 20 : -5) * PB_DEG;
        for (int i = 0; i < NUM_LON; i++) lats[i] = jmax(clampMin, jmin(clampMax, lats[i]));
        const int n = NUM_LON;
        const double period = 2 * PB_PI;
        for (int i = 0; i < n; i++) {
            const int next = (i + 1) % n;
            h[i] = fmod(sp->xs[next] - sp->xs[i] + period, period);
            if (h[i] == 0) h[i] = period / n;
        }
        for (int i = 0; i < n; i++) {
            const int prev = (i - 1 + n) % n, next = (i + 1) % n;
            alpha[i] = (3 / h[i]) * (lats[next] - lats[i]) - (3 / h[prev]) * (lats[i] - lats[prev]);
        }
        for (int i = 0; i < n; i++) sp->c[i] = 0;
        for (int iter = 0; iter < 20; iter++)
            for (int i = 0; i < n; i++) {
                const int prev = (i - 1 + n) % n, next = (i + 1) % n;
                sp->c[i] = (alpha[i] - h[prev] * sp->c[prev] - h[i] * sp->c[next]) / (2 * (h[prev] + h[i]));
            }
        for (int i = 0; i < n; i++) {
            const int next = (i + 1) % n;
            sp->ys[i] = lats[i];
            sp->b[i] = (lats[next] - lats[i]) / h[i] - h[i] * (sp->c[next] + 2 * sp->c[i]) / 3;
            sp->d[i] = (sp->c[next] - sp->c[i]) / (3 * h[i]);
        }
    }
};
struct ItczTableK {   // :656-668
    const SplineDev* splines; float* lons; float* latsSummer; float* latsWinter;
    PB_DEV void operator()(int i) const {
        const double lon = -PB_PI + (i + 0.5) * (2 * PB_PI / PB_ITCZ_SAMPLES);
        lons[i] = (float)lon;
        latsSummer[i] = (float)evaluate_spline(splines + 0, lon);
        latsWinter[i] = (float)evaluate_spline(splines + 1, lon);
    }
};

// ---- BFS hop counts (wind.js:510-538, 559-586; ocean.js:58-80) -------------------------------------------
// frontier expansion, one launch per level; three rotating counters: level L reads cnt[L%3], appends to
// cnt[(L+1)%3] and clears cnt[(L+2)%3].
struct BfsLevelK {
    Csr g; const uint8_t* passable; int* dist; const int* frontier; int* next; int* cnt; int level;
    PB_DEV void block0() const { cnt[(level + 2) % 3] = 0; }
    PB_DEV void operator()(int i) const {
        const int r = frontier[i];
        const int d = level + 1;
        for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
            const int nb = g.adj[j];
            if (passable[nb] && dist[nb] == -1 && atomic_cas(dist + nb, -1, d) == -1) next[atomic_add(cnt + (level + 1) % 3, 1)] = nb;
        }
    }
};
#if PB_CUDA
}  // namespace pb
#include <cooperative_groups.h>
namespace pb {
// All levels of up to three independent BFS in a single cooperative launch: one grid-wide barrier per level for all of them
// together instead of one kernel launch per level and BFS (a BFS at 1M cells has 100–300 levels of a few thousand cells each —
// barrier-latency bound, so the two BFS of computeWind and the three of computeOceanCurrents each share their barriers).
struct BfsMulti { int k; const uint8_t* passable[3]; int* dist[3]; int* fa[3]; int* fb[3]; int* cnt[3]; };
__global__ void __launch_bounds__(256) k_bfs_persistent(Csr g, BfsMulti B) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int level = 0;; level++) {
        bool any = false;
        for (int b = 0; b < B.k; b++) {
            int* cnt = B.cnt[b];
            const int n = ld_volatile(cnt + level % 3);
            if (gtid == 0) cnt[(level + 2) % 3] = 0;
            if (n == 0) continue;
            any = true;
            const int d = level + 1;
            int* slot = cnt + (level + 1) % 3;
            const int* cur = (level & 1) ? B.fb[b] : B.fa[b];
            int* nxt = (level & 1) ? B.fa[b] : B.fb[b];
            const uint8_t* passable = B.passable[b];
            int* dist = B.dist[b];
            for (int i = gtid; i < n; i += stride) {
                const int r = __ldcg(cur + i);      // written by other SMs in the previous level: read through L2
                for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
                    const int nb = g.adj[j];
                    if (passable[nb] && ld_volatile(dist + nb) == -1 && atomicCAS(dist + nb, -1, d) == -1) nxt[atomicAdd(slot, 1)] = nb;
                }
            }
        }
        if (!any) break;
        grid.sync();
    }
}
#endif

// seeds: land cells touching the main ocean component (wind.js:514-523)
struct LandCoastSeedK {
    Csr g; const uint8_t* isLand; const uint8_t* isOcean; const int* parent; const unsigned long long* best; int* dist; uint8_t* flag;
    PB_DEV void operator()(int r) const {
        int d = -1;
        if (isLand[r])
            for (int j = g.off[r], e = g.off[r + 1]; j < e; j++)
                if (is_main_component(g.adj[j])) { d = 0; break; }
        dist[r] = d; flag[r] = d == 0;
    }
    PB_DEV bool is_main_component(int c) const {
        if (!isOcean[c]) return false;
        const unsigned long long b = *best;
        if (b == 0) return false;
        const int mainRoot = (int)(0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFull));
        int x = c, p = parent[x];
        while (p != x) { x = p; p = parent[x]; }
        return x == mainRoot;
    }
};
struct PlateTableK { const int* ids; uint8_t* table; PB_DEV void operator()(int i) const { table[ids[i]] = 1; } };
struct ContPlateK {   // contPlate[r] = !plateIsOcean.has(r_plate[r])
    const int* r_plate; const uint8_t* table; int tableSize; uint8_t* contPlate;
    PB_DEV void operator()(int r) const {
        const int p = r_plate[r];
        contPlate[r] = (p >= 0 && p < tableSize && table[p]) ? 0 : 1;
    }
};
struct MaskBoundarySeedK {   // dist 0 where mask[r] and some neighbour is outside the mask (wind.js:561-573)
    Csr g; const uint8_t* mask; int* dist; uint8_t* flag;
    PB_DEV void operator()(int r) const {
        int d = -1;
        if (mask[r])
            for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) if (!mask[g.adj[j]]) { d = 0; break; }
        dist[r] = d; flag[r] = d == 0;
    }
};
struct ContinentalityK {   // :541-549, :587-592
    const uint8_t* mask; const int* dist; float* out; double avgEdgeKm;
    PB_DEV void operator()(int r) const {
        out[r] = (mask[r] && dist[r] >= 0) ? (float)smoothstep(0, 2000, dist[r] * avgEdgeKm) : 0.0f;
    }
};

// ---- pressure (:239-301), gradient (:306-339), wind (:343-378) ---------------------------------------------
struct PressureK {
    const float* lat; const float* lon; const SplineDev* spline; int summer; const float* cont; const float* elev;
    Simplex noise; const float* xyz; float* out;
    PB_DEV static double gauss(double q) { return pb_exp(-0.5 * (q * q)); }
    PB_DEV void operator()(int r) const {
        const double la = lat[r], lo = lon[r], landFrac = cont[r];
        const double itczLat = evaluate_spline(spline, lo);
        const double latDeg = la * PB_RAD;
        const double seasonSign = summer ? 1 : -1;
        double p = 1013;
        const double dItcz = (la - itczLat) * PB_RAD;
        p -= 15 * gauss(dItcz / 8);
        const double shiftDeg = seasonSign * 5;
        const double nhSubHigh = 30 + shiftDeg;
        const double shSubHigh = -(30 - shiftDeg);
        const double highIntensity = 12 * (1 - 0.3 * landFrac);
        p += highIntensity * gauss((latDeg - nhSubHigh) / 10);
        p += highIntensity * gauss((latDeg - shSubHigh) / 10);
        p -= 10 * gauss((latDeg - 60) / 10);
        p -= 10 * gauss((latDeg + 60) / 10);
        p += 8 * gauss((latDeg - 85) / 8);
        p += 8 * gauss((latDeg + 85) / 8);
        const double continentalScale = smoothstep(0.2, 0.5, landFrac);
        if (continentalScale > 0.001) {
            const double absLatDeg = fabs(la) * PB_RAD;
            const double latFactor = absLatDeg < 15 ? 0
                : absLatDeg < 30 ? 0.75 * smoothstep(15, 30, absLatDeg)
                : absLatDeg < 45 ? 0.75 + 0.25 * smoothstep(30, 45, absLatDeg)
                : absLatDeg < 60 ? 1
                : absLatDeg < 90 ? smoothstep(90, 60, absLatDeg)
                : 0;
            const bool isSummerHemisphere = (seasonSign > 0 && la > 0) || (seasonSign < 0 && la < 0);
            if (isSummerHemisphere) p -= 10 * latFactor * continentalScale;
            else p += 14 * latFactor * continentalScale;
        }
        p -= 3 * elev_to_height_km(jmax(0, (double)elev[r]));
        p += noise.fbm((double)xyz[3 * r] * 2, (double)xyz[3 * r + 1] * 2, (double)xyz[3 * r + 2] * 2, 3) * 2;
        out[r] = (float)p;
    }
};
struct GradientsK {
    Csr g; const float* xyz; const float* P; const float *eX, *eY, *eZ, *nX, *nY, *nZ; float* gradE; float* gradN;
    PB_DEV void operator()(int r) const {
        const double px = xyz[3 * r], py = xyz[3 * r + 1], pz = xyz[3 * r + 2];
        const double ex = eX[r], ey = eY[r], ez = eZ[r], nx = nX[r], ny = nY[r], nz = nZ[r];
        const double pHere = P[r];
        double sumEP = 0, sumEE = 0, sumNP = 0, sumNN = 0;
        for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
            const int nb = g.adj[j];
            const double dx = (double)xyz[3 * nb] - px, dy = (double)xyz[3 * nb + 1] - py, dz = (double)xyz[3 * nb + 2] - pz;
            const double de = dx * ex + dy * ey + dz * ez;
            const double dn = dx * nx + dy * ny + dz * nz;
            const double dp = (double)P[nb] - pHere;
            sumEP += de * dp; sumEE += de * de; sumNP += dn * dp; sumNN += dn * dn;
        }
        gradE[r] = (float)(sumEE > 1e-12 ? sumEP / sumEE : 0);
        gradN[r] = (float)(sumNN > 1e-12 ? sumNP / sumNN : 0);
    }
};
struct PressureToWindK {
    const float* gradE; const float* gradN; const float* sinLat; float* windE; float* windN; float* speed; uint32_t* key;
    PB_DEV void operator()(int r) const {
        const double sin5 = pb_sin(5 * PB_DEG);
        const double pgfE = -(double)gradE[r], pgfN = -(double)gradN[r];
        const double sl = sinLat[r];
        const double geoAngle = 70 * PB_DEG * smoothstep(0, sin5, fabs(sl));
        const double frictionAngle = 20 * PB_DEG;
        const double sign = sl >= 0 ? -1 : 1;
        const double totalAngle = sign * (geoAngle - frictionAngle);
        double sinA, cosA;
        pb_sincos(totalAngle, &sinA, &cosA);
        const double we = (pgfE * cosA - pgfN * sinA) * 0.6;
        const double wn = (pgfE * sinA + pgfN * cosA) * 0.6;
        windE[r] = (float)we; windN[r] = (float)wn;
        const float sp = (float)sqrt(we * we + wn * wn);
        speed[r] = sp;
        key[r] = f32_sort_key(sp);
    }
};
// percentile (js/climate-util.js:103-110) after an ascending sort of the order-preserving keys:
// value at index floor(n*p), `|| 1`.  n comes from the device when nDev != nullptr.
struct PercentilePickK {
    const uint32_t* sortedKeys; int n; const int* nDev; double p; double* out;
    PB_DEV void operator()(int) const {
        const int cnt = nDev ? *nDev : n;
        if (cnt == 0) { *out = 1; return; }
        const int k = (int)floor((double)cnt * p);
        const uint32_t u = sortedKeys[k];
        const uint32_t bits = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        float f;
#if PB_CUDA
        f = __uint_as_float(bits);
#else
        memcpy(&f, &bits, 4);
#endif
        *out = or_default((double)f, 1.0);
    }
};
struct NormalizeMin1K {   // x = min(1, x / *scale)
    float* x; const double* scale;
    PB_DEV void operator()(int r) const { x[r] = (float)jmin(1, (double)x[r] / *scale); }
};
struct PressureDevK { const float* p; float* out; PB_DEV void operator()(int r) const { out[r] = (float)((double)p[r] - 1013); } };

// ---- masked Jacobi sweeps ----------------------------------------------------------------------------------------
// smoothOcean (ocean.js:168-189) and the land-only west-coast smoothing (heuristic-precip.js:154-166):
// cells outside the mask keep their value (copyOutside) or become 0.
struct SmoothMaskedK {
    Csr g; const uint8_t* mask; const float* src; float* dst; int zeroOutside;
    PB_DEV void operator()(int r) const {
        if (!mask[r]) { dst[r] = zeroOutside ? 0.0f : src[r]; return; }
        if (g.pack) {
            RowIds row; int deg;
            if (row.load_packed(g.pack, r, deg)) { gathered(r, deg, row); return; }
        }
        const int b = g.off[r];
        this->row(r, b, g.off[r + 1] - b, g.adj + b);
    }
    PB_DEV void gathered(int r, int deg, const RowIds& row) const {
        double sum = src[r]; int count = 1;
        float v[PB_ROW_FAST]; uint8_t mk[PB_ROW_FAST];
#pragma unroll
        for (int k = 0; k < PB_ROW_FAST; k++) { v[k] = src[row.nb[k]]; mk[k] = mask[row.nb[k]]; }
#pragma unroll
        for (int k = 0; k < PB_ROW_FAST; k++) if (k < deg && mk[k]) { sum += v[k]; count++; }
        dst[r] = (float)(sum / count);
    }
    PB_DEV void row(int r, int b, int deg, const int* ids) const {
        (void)b;
        if (!mask[r]) { dst[r] = zeroOutside ? 0.0f : src[r]; return; }
        if (deg <= PB_ROW_FAST) { RowIds rw; rw.load(r, deg, ids); gathered(r, deg, rw); return; }
        double sum = src[r]; int count = 1;
        for (int j = 0; j < deg; j++) { const int nb = ids[j]; if (mask[nb]) { sum += src[nb]; count++; } }
        dst[r] = (float)(sum / count);
    }
};
// diffuseOceanWarmth sweep (temperature.js:33-51): cells with plate continentality >= 0.95 keep their value
struct DiffuseWarmthK {
    Csr g; const float* pcont; const float* src; float* dst;
    PB_DEV void operator()(int r) const {
        if ((double)pcont[r] >= 0.95) { dst[r] = src[r]; return; }
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
        if ((double)pcont[r] >= 0.95) { dst[r] = src[r]; return; }
        if (deg <= PB_ROW_FAST) { RowIds rw; rw.load(r, deg, ids); gathered(r, deg, rw); return; }
        double sum = src[r];
        for (int j = 0; j < deg; j++) sum += src[ids[j]];
        dst[r] = (float)(sum / (deg + 1));
    }
};

// ---- ocean.js ----------------------------------------------------------------------------------------------------
struct OceanCoastSeedK {   // computeCoastFields seeds (:21-55)
    Csr g; const float* xyz; const uint8_t* isOcean; const float *eX, *eY, *eZ;
    int* coastDist; int* westDist; int* eastDist; uint8_t* fCoast; uint8_t* fWest; uint8_t* fEast;
    PB_DEV void operator()(int r) const {
        int c = -1, w = -1, ea = -1;
        if (isOcean[r]) {
            double lx = 0, ly = 0, lz = 0;
            bool has = false;
            for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
                const int nb = g.adj[j];
                if (!isOcean[nb]) {
                    has = true;
                    lx += (double)xyz[3 * nb] - (double)xyz[3 * r];
                    ly += (double)xyz[3 * nb + 1] - (double)xyz[3 * r + 1];
                    lz += (double)xyz[3 * nb + 2] - (double)xyz[3 * r + 2];
                }
            }
            if (has) {
                c = 0;
                const double normalE = lx * (double)eX[r] + ly * (double)eY[r] + lz * (double)eZ[r];
                if (normalE < -0.2) w = 0;
                else if (normalE > 0.2) ea = 0;
                else { if (normalE <= 0) w = 0; else ea = 0; }
            }
        }
        coastDist[r] = c; westDist[r] = w; eastDist[r] = ea;
        fCoast[r] = c == 0; fWest[r] = w == 0; fEast[r] = ea == 0;
    }
};
struct CircumpolarBinsK {   // hasCircumpolarChannel (:88-111) for both hemispheres: bins[h*72 + bin] = 1
    const float* lat; const float* lon; const uint8_t* isOcean; int* bins;
    PB_DEV void operator()(int r) const {
        if (!isOcean[r]) return;
        const double la = lat[r];
        double bin = floor((((double)lon[r] + PB_PI) / (2 * PB_PI)) * 72);
        bin = fmod(fmod(bin, 72.0) + 72, 72.0);
        const double t = 60 * PB_DEG, w = 5 * PB_DEG;
        if (!(la < t - w || la > t + w)) bins[(int)bin] = 1;
        if (!(la < -t - w || la > -t + w)) bins[72 + (int)bin] = 1;
    }
};
struct CircumpolarFlagK {
    const int* bins; int* flags;
    PB_DEV void operator()(int h) const {
        int all = 1;
        for (int i = 0; i < 72; i++) if (!bins[h * 72 + i]) all = 0;
        flags[h] = all;
    }
};
struct OceanCurrentsK {   // :266-333
    const float* lat; const float* lon; const uint8_t* isOcean; const float* itczLats; const int* westDist; const int* eastDist;
    const int* circ; double seasonalShiftDeg, coastThreshold; float* curE; float* curN;
    PB_DEV void operator()(int r) const {
        if (!isOcean[r]) { curE[r] = 0; curN[r] = 0; return; }
        const double la = lat[r];
        const double absLatDeg = fabs(la) / PB_DEG;
        const double hemisphereSign = la >= 0 ? 1 : -1;
        const double bandLatDeg = fabs(la / PB_DEG - seasonalShiftDeg);
        const double itczLat = itcz_lookup(itczLats, lon[r]);
        const double distFromItcz = fabs(la - itczLat) / PB_DEG;
        double baseE;
        if (distFromItcz < 3) baseE = 1 - 2 * smoothstep(0, 3, distFromItcz);
        else if (bandLatDeg < 30) baseE = -1;
        else if (bandLatDeg < 35) baseE = -1 + 2 * smoothstep(30, 35, bandLatDeg);
        else if (bandLatDeg < 58) baseE = 1;
        else if (bandLatDeg < 65) baseE = 1 - 1.5 * smoothstep(58, 65, bandLatDeg);
        else baseE = -0.5;
        float cE = (float)baseE, cN = 0.0f;
        const double wDist = westDist[r], eDist = eastDist[r];
        if (wDist >= 0 && wDist < coastThreshold) {
            const double t = 1 - wDist / coastThreshold;
            const double strength = t * t * 2.0;
            cN = (float)((double)cN + hemisphereSign * strength);
            cE = (float)((double)cE * (1 - t * t * 0.7));
        }
        if (eDist >= 0 && eDist < coastThreshold) {
            const double t = 1 - eDist / coastThreshold;
            const double strength = t * t * 0.8;
            cN = (float)((double)cN - hemisphereSign * strength);
            cE = (float)((double)cE * (1 - t * t * 0.5));
        }
        const bool isCircumpolar = (la > 0 && circ[0]) || (la < 0 && circ[1]);
        if (isCircumpolar && absLatDeg >= 55 && absLatDeg <= 75) {
            const double cStrength = 1 - fabs(absLatDeg - 65) / 10;
            cE = (float)((double)cE * (1 - cStrength) + 1.5 * cStrength);
            cN = (float)((double)cN * (1 - cStrength * 0.8));
        }
        curE[r] = cE; curN[r] = cN;
    }
};
struct ZeroOutsideK { const uint8_t* mask; float* a; float* b; PB_DEV void operator()(int r) const { if (!mask[r]) { a[r] = 0; b[r] = 0; } } };
struct WarmthK {   // classifyWarmth (:120-164)
    const uint8_t* isOcean; const float* lat; const int* westDist; const int* eastDist; double fadeRange, seasonalShiftDeg; float* warmth;
    PB_DEV void operator()(int r) const {
        if (!isOcean[r]) { warmth[r] = 0; return; }
        const double bandLatDeg = fabs((double)lat[r] / PB_DEG - seasonalShiftDeg);
        double cellSign;
        if (bandLatDeg < 28) cellSign = 1;
        else if (bandLatDeg < 35) cellSign = 1 - 2 * smoothstep(28, 35, bandLatDeg);
        else if (bandLatDeg < 55) cellSign = -1;
        else if (bandLatDeg < 65) cellSign = -1 + 2 * smoothstep(55, 65, bandLatDeg);
        else cellSign = 1;
        const double wDist = westDist[r], eDist = eastDist[r];
        double warm = 0;
        if (wDist >= 0 && wDist < fadeRange) { const double t = 1 - wDist / fadeRange; warm += cellSign * t * t; }
        if (eDist >= 0 && eDist < fadeRange) { const double t = 1 - eDist / fadeRange; warm -= cellSign * t * t; }
        warmth[r] = (float)jmax(-1, jmin(1, warm));
    }
};
struct OceanSpeedK {   // :359-365: speed + selection key (ocean cells with speed > 0 only) + their count
    const float* curE; const float* curN; const uint8_t* isOcean; float* speed; uint32_t* key; int* count;
    PB_DEV void operator()(int r) const {
        const double e = curE[r], n = curN[r];
        const float sp = (float)sqrt(e * e + n * n);
        speed[r] = sp;
        if (isOcean[r] && sp > 0) { key[r] = f32_sort_key(sp); atomic_add(count, 1); }
        else key[r] = 0xFFFFFFFFu;
    }
};

// ---- heuristic-precip.js ----------------------------------------------------------------------------------------------
PB_DEV double zonal_base(double d) {   // :16-38
    if (d < 5) return 1.0;
    else if (d < 10) return 1.0 - 0.65 * smoothstep(5, 10, d);
    else if (d < 33) return 0.35 - 0.33 * smoothstep(10, 28, d);
    else if (d < 55) return 0.02 + 0.48 * smoothstep(33, 55, d);
    else if (d < 70) return 0.5 - 0.2 * smoothstep(55, 70, d);
    else return 0.3 - 0.2 * smoothstep(70, 90, d);
}
PB_DEV void heuristic_wind(double dist, bool north, double* we, double* wn) {   // :52-86
    const double hemiSign = north ? 1 : -1;
    if (dist < 5) { *we = 0; *wn = -hemiSign * 0.1; }
    else if (dist < 30) {
        const double s = smoothstep(5, 15, dist) * (1 - smoothstep(25, 32, dist));
        *we = -s * 0.8; *wn = -hemiSign * s * 0.3;
    } else if (dist < 60) {
        const double s = smoothstep(30, 40, dist) * (1 - smoothstep(55, 65, dist));
        *we = s * 0.9; *wn = hemiSign * s * 0.25;
    } else {
        const double s = smoothstep(60, 70, dist);
        *we = -s * 0.4; *wn = -hemiSign * s * 0.15;
    }
}
struct WestCoastSeedK {   // :131-150
    Csr g; const float* xyz; const uint8_t* isLand; const int* coastDistLand; const float *eX, *eY, *eZ; float* westCoast;
    PB_DEV void operator()(int r) const {
        float out = 0;
        if (isLand[r] && coastDistLand[r] == 0) {
            double dotE = 0; int count = 0;
            for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
                const int nb = g.adj[j];
                if (!isLand[nb]) {
                    const double dx = (double)xyz[3 * nb] - (double)xyz[3 * r];
                    const double dy = (double)xyz[3 * nb + 1] - (double)xyz[3 * r + 1];
                    const double dz = (double)xyz[3 * nb + 2] - (double)xyz[3 * r + 2];
                    dotE += dx * (double)eX[r] + dy * (double)eY[r] + dz * (double)eZ[r];
                    count++;
                }
            }
            if (count > 0) out = dotE < 0 ? 1.0f : -1.0f;
        }
        westCoast[r] = out;
    }
};
struct HeuristicPrecipK {   // :183-259
    const float* lat; const float* lon; const uint8_t* isLand; const float* cont; const float* elev; const float* itczLats;
    const float* westCoast; const float* gradE; const float* gradN; const int* coastDistLand; int isSummer; double avgEdgeKm; float* precip;
    PB_DEV void operator()(int r) const {
        const double la = lat[r];
        const double itczLat = itcz_lookup(itczLats, lon[r]) * 0.3;
        const double signedDist = la - itczLat;
        const double distFromItczDeg = fabs(signedDist) / PB_DEG;
        const bool isNorthOfItcz = signedDist > 0;
        const double zonal = zonal_base(distFromItczDeg);
        const double absLatDeg = fabs(la) / PB_DEG;
        const bool inSummerHemi = isSummer ? (la >= 0) : (la < 0);
        double seasonMod = inSummerHemi ? 1.1 : 0.9;
        if (inSummerHemi && absLatDeg > 22 && absLatDeg < 45) {
            const double medSuppress = smoothstep(22, 30, absLatDeg) * (1 - smoothstep(38, 45, absLatDeg));
            const double strength = 0.15 + (double)westCoast[r] * 0.20;
            seasonMod *= (1 - medSuppress * jmax(0, strength));
        }
        double contMod = 1.0;
        const double c = isLand[r] ? (double)cont[r] : 0;
        if (c > 0) contMod = 1.0 - c * c * 0.65;
        double oroMod = 1.0;
        if (isLand[r] && elev[r] > 0) {
            double we, wn;
            heuristic_wind(distFromItczDeg, isNorthOfItcz, &we, &wn);
            const double windDotGrad = we * (double)gradE[r] + wn * (double)gradN[r];
            if (windDotGrad > 0) oroMod = 1.0 + jmin(1, windDotGrad * 15) * 0.6;
            else {
                const double heightKm = elev_to_height_km(jmax(0, (double)elev[r]));
                const double heightScale = jmin(1, heightKm / 3);
                const double shadow = jmin(1, -windDotGrad * 18);
                oroMod = jmax(0.3, 1.0 - shadow * 0.7 * heightScale);
            }
        }
        double distMod = 1.0;
        if (isLand[r] && coastDistLand[r] > 0) {
            const double distKm = coastDistLand[r] * avgEdgeKm;
            if (distKm > 2000) distMod = jmax(0.03, 1 - smoothstep(2000, 3000, distKm));
        }
        precip[r] = (float)jmax(0.05, zonal * seasonMod * contMod * oroMod * distMod);
    }
};

// ---- precipitation.js ---------------------------------------------------------------------------------------------------
struct BlendElevK { float* smoothed; const float* elev; PB_DEV void operator()(int r) const { smoothed[r] = (float)((double)smoothed[r] * 0.6 + (double)elev[r] * 0.4); } };
struct HeightKmK { const float* elev; float* out; PB_DEV void operator()(int r) const { out[r] = (float)elev_to_height_km(jmax(0, (double)elev[r])); } };
// heuristic wind (heuristic-precip.js:90-108) + 50/50 blend (:263-270) + 3-D wind (:273-281)
struct WindBlendK {
    const float* lat; const float* lon; const float* itczLats; const float* rawE; const float* rawN;
    const float *eX, *eY, *eZ, *nX, *nY, *nZ; float* windE; float* windN; float* wX; float* wY; float* wZ;
    PB_DEV void operator()(int r) const {
        const double la = lat[r];
        const double itczLat = itcz_lookup(itczLats, lon[r]) * 0.3;
        const double signedDist = la - itczLat;
        double hwe, hwn;
        heuristic_wind(fabs(signedDist) / PB_DEG, signedDist > 0, &hwe, &hwn);
        const float hE = (float)hwe, hN = (float)hwn;
        const float wE = (float)(0.5 * (double)rawE[r] + 0.5 * (double)hE);
        const float wN = (float)(0.5 * (double)rawN[r] + 0.5 * (double)hN);
        windE[r] = wE; windN[r] = wN;
        const double we = wE, wn = wN;
        wX[r] = (float)(we * (double)eX[r] + wn * (double)nX[r]);
        wY[r] = (float)(we * (double)eY[r] + wn * (double)nY[r]);
        wZ[r] = (float)(we * (double)eZ[r] + wn * (double)nZ[r]);
    }
};
struct ConvergenceK {   // :19-52
    Csr g; const float* xyz; const float* wX; const float* wY; const float* wZ; float* out;
    PB_DEV void operator()(int r) const {
        const double wdx = wX[r], wdy = wY[r], wdz = wZ[r];
        double conv = 0; int count = 0;
        for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
            const int nb = g.adj[j];
            const double dx = (double)xyz[3 * nb] - (double)xyz[3 * r];
            const double dy = (double)xyz[3 * nb + 1] - (double)xyz[3 * r + 1];
            const double dz = (double)xyz[3 * nb + 2] - (double)xyz[3 * r + 2];
            conv -= ((double)wX[nb] + wdx) * dx + ((double)wY[nb] + wdy) * dy + ((double)wZ[nb] + wdz) * dz;
            count++;
        }
        out[r] = (float)(count > 0 ? conv / count : 0);
    }
};
struct MoistureInitK {   // :69-109
    Csr g; const float* xyz; const uint8_t* isLand; const int* coastDistLand; const float* warmth;
    const float* wX; const float* wY; const float* wZ; float* moisture;
    PB_DEV void operator()(int r) const {
        if (!isLand[r]) { moisture[r] = (float)(0.4 + 0.35 * jmax(0, (double)warmth[r])); return; }
        float out = 0;
        if (coastDistLand[r] == 0) {
            double warmthSum = 0, ox = 0, oy = 0, oz = 0; int oceanCount = 0;
            for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
                const int nb = g.adj[j];
                if (!isLand[nb]) {
                    oceanCount++;
                    warmthSum += warmth[nb];
                    ox += (double)xyz[3 * nb] - (double)xyz[3 * r];
                    oy += (double)xyz[3 * nb + 1] - (double)xyz[3 * r + 1];
                    oz += (double)xyz[3 * nb + 2] - (double)xyz[3 * r + 2];
                }
            }
            if (oceanCount > 0) {
                const double avgWarmth = warmthSum / oceanCount;
                const double windDotOcean = (double)wX[r] * ox + (double)wY[r] * oy + (double)wZ[r] * oz;
                const double onshore = windDotOcean < 0 ? 1.0 : 0.25;
                const double warmthFactor = 0.5 + 0.5 * jmax(-0.8, jmin(1, avgWarmth));
                out = (float)(onshore * warmthFactor);
            }
        }
        moisture[r] = out;
    }
};
struct AdvectK {   // one sweep of :118-179
    Csr g; const float* xyz; const uint8_t* isLand; const float* windE; const float* windN;
    const float* wX; const float* wY; const float* wZ; const float* heightKm; const float* src; float* dst;
    double depletionBase; int maxHops;
    PB_DEV void operator()(int r) const {
        const float s = src[r];
        if (!isLand[r]) { dst[r] = s; return; }
        const double we = windE[r], wn = windN[r];
        if (we * we + wn * wn < 1e-6) { dst[r] = s; return; }
        double upM = 0, upW = 0, upH = 0;
        const double heightHere = heightKm[r];
        const double px = xyz[3 * r], py = xyz[3 * r + 1], pz = xyz[3 * r + 2];
        for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
            const int nb = g.adj[j];
            const double dx = px - (double)xyz[3 * nb], dy = py - (double)xyz[3 * nb + 1], dz = pz - (double)xyz[3 * nb + 2];
            const double dot = (double)wX[nb] * dx + (double)wY[nb] * dy + (double)wZ[nb] * dz;
            if (dot > 0) { upM += (double)src[nb] * dot; upH += (double)heightKm[nb] * dot; upW += dot; }
        }
        if (upW > 0) {
            const double incoming = upM / upW;
            const double upwindHeight = upH / upW;
            const double heightGain = jmax(0, heightHere - upwindHeight);
            const double normalizedGain = heightGain * maxHops;
            const double elevDepletion = jmin(0.8, normalizedGain * 0.55);
            const double depletion = depletionBase + elevDepletion;
            const double carried = incoming * jmax(0, 1 - depletion);
            dst[r] = (float)jmax((double)s, carried);
        } else dst[r] = s;
    }
};
struct PrecipParams {
    int summer, maxHops;
    double avgEdgeKm, avgEdgeRad, precipitationOffset, landCoverage;
};
struct PrecipMechanismsK {   // :307-487 plus the rain-shadow seed :501-513
    const float* lat; const float* lon; const float* elev; const uint8_t* isLand; const float* cont; const float* itczLats;
    const float* moisture; const float* convergence; const float* windE; const float* windN; const float* gradE; const float* gradN;
    const float* pressureDev; const int* coastDistLand; const float* heightKm; PrecipParams P; float* precip; float* rainShadow;
    PB_DEV void operator()(int r) const {
        const double la = lat[r];
        const double absLatDeg = fabs(la) / PB_DEG;
        const double el = elev[r];
        const bool land = isLand[r] != 0;
        const int cdl = coastDistLand[r];
        double p = moisture[r];
        const double itczLat = itcz_lookup(itczLats, lon[r]);
        const double distFromItcz = fabs(la - itczLat) / PB_DEG;
        const double c = land ? (double)cont[r] : 0;
        if (distFromItcz < 15) {
            const double itczStrength = smoothstep(15, 0, distFromItcz);
            const double coreBoost = distFromItcz < 5 ? 1.5 : 1.0;
            p = p * (1 + itczStrength * coreBoost) + itczStrength * 0.3;
        }
        const double conv = convergence[r];
        if (conv > 0) {
            const double convStrength = jmin(1, (conv / P.avgEdgeRad) * 0.055);
            p = p * (1 + convStrength * 1.2) + convStrength * (double)moisture[r] * 0.4;
        }
        const double we = windE[r], wn = windN[r];
        const double windDotGrad = we * (double)gradE[r] + wn * (double)gradN[r];
        if (land && el > 0) {
            if (windDotGrad > 0) p += jmin(1, windDotGrad * 15) * 1.0;
            else p *= jmax(0.02, 1 - jmin(1, -windDotGrad * 18) * 0.95);
        }
        const double pDev = pressureDev[r];
        const bool inLocalSummer = P.summer ? (la >= 0) : (la < 0);
        const double subtropCenter = inLocalSummer ? 30 : 24;
        const double subtropWidth = inLocalSummer ? 16 : 12;
        double subtropPeak = inLocalSummer ? 0.50 : 0.30;
        if (land && inLocalSummer) {
            const double polewardWind = la >= 0 ? wn : -wn;
            if (polewardWind > 0) {
                const double coastDist = cdl >= 0 ? cdl : P.maxHops;
                const double coastProximity = 1 - smoothstep(0, P.maxHops * 0.4, coastDist);
                const double monsoonRelief = smoothstep(0, 0.15, polewardWind) * coastProximity;
                subtropPeak *= (1 - monsoonRelief * 0.7);
            }
        }
        const double subtropDist = fabs(absLatDeg - subtropCenter);
        const double latBandSuppression = subtropDist < subtropWidth ? smoothstep(subtropWidth, 0, subtropDist) * subtropPeak : 0;
        double pressureMod;
        if (pDev > 0) pressureMod = smoothstep(0, 12, pDev) * 0.25;
        else pressureMod = -smoothstep(0, 15, -pDev) * 0.2;
        const double totalSuppression = jmax(0, latBandSuppression + pressureMod);
        if (totalSuppression > 0) p *= jmax(0.05, 1 - totalSuppression);
        else p *= (1 - totalSuppression);
        if (absLatDeg > 40) {
            const double polarStrength = smoothstep(40, 70, absLatDeg);
            const double coastDist = cdl < 0 ? P.maxHops : cdl;
            const double inlandFade = 1 - smoothstep(0, P.maxHops, coastDist);
            const double polarBase = polarStrength * 0.10;
            const double polarCoastal = polarStrength * 0.20 * inlandFade;
            p += polarBase + polarCoastal;
            p *= (1 + polarStrength * 0.15);
        }
        if (land && c > 0) p *= jmax(0.03, 1 - c * c * 0.55);
        const double hk = heightKm[r];
        if (land && hk > 1.5) {
            const double leeCoastHops = jmax(2, floor(200 / P.avgEdgeKm + 0.5));
            if (windDotGrad < -0.01 && cdl >= 0 && cdl < leeCoastHops) p += 0.15 * jmin(1, hk / 5);
        }
        if (!land) {
            const double highPressureFade = pDev > 0 ? smoothstep(0, 12, pDev) : 0;
            const double oceanBase = 0.15 * (1 - highPressureFade);
            p = jmax(p, oceanBase);
        }
        if (land && cdl > 0) {
            const double distKm = cdl * P.avgEdgeKm;
            if (distKm > 2000) p *= jmax(0.03, 1 - smoothstep(2000, 3000, distKm));
        }
        const double precipMult = 1 + P.precipitationOffset * 0.5;
        double finalPrecip = p * precipMult;
        if (P.landCoverage > 0.4) {
            const double t = (P.landCoverage - 0.4) / 0.6;
            finalPrecip *= 1 - t * t * 0.98;
        }
        precip[r] = (float)jmax(0, finalPrecip);
        // rain-shadow seed
        float rs = 0;
        if (land && el > 0 && !(hk < 0.8)) {
            const double heightScale = jmin(1, (hk - 0.5) / 2.5);
            if (windDotGrad > 0) rs = (float)(jmin(1, windDotGrad * 20) * heightScale);
            else if (windDotGrad < 0) rs = (float)(-jmin(1, -windDotGrad * 18) * heightScale);
        }
        rainShadow[r] = rs;
    }
};
// wind-aligned edge weights (:520-547), kept in CSR position: weight 0 = edge not in the list
struct EdgeWeightsK {
    Csr g; const float* xyz; const uint8_t* isLand; const float* wX; const float* wY; const float* wZ; float* upWt; float* dnWt;
    PB_DEV void operator()(int r) const {
        const bool land = isLand[r] != 0;
        const double px = xyz[3 * r], py = xyz[3 * r + 1], pz = xyz[3 * r + 2];
        const double wx = wX[r], wy = wY[r], wz = wZ[r];
        for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
            float up = 0, dn = 0;
            if (land) {
                const int nb = g.adj[j];
                const double dx = px - (double)xyz[3 * nb], dy = py - (double)xyz[3 * nb + 1], dz = pz - (double)xyz[3 * nb + 2];
                const double upDot = (double)wX[nb] * dx + (double)wY[nb] * dy + (double)wZ[nb] * dz;
                if (upDot > 0) up = (float)upDot;
                const double dnDot = -(wx * dx + wy * dy + wz * dz);
                if (dnDot > 0) dn = (float)dnDot;
            }
            upWt[j] = up; dnWt[j] = dn;
        }
    }
};
// one propagation sweep (:555-571 with sign = -1, keep = min; :582-598 with sign = +1, keep = max)
struct ShadowSweepK {
    Csr g; const uint8_t* isLand; const float* wt; const float* src; float* dst; double keepFactor; int sign;
    PB_DEV void operator()(int r) const {
        const float s = src[r];
        if (!isLand[r]) { dst[r] = s; return; }
        const int b = g.off[r];
        if (g.pack) {
            RowIds row; int deg;
            if (row.load_packed(g.pack, r, deg)) { gathered(r, s, b, deg, row); return; }
        }
        this->row(r, b, g.off[r + 1] - b, g.adj + b);
    }
    PB_DEV void finish(int r, float s, double val, double w) const {
        if (w > 0) {
            const double carried = (val / w) * keepFactor;
            dst[r] = (float)(sign < 0 ? jmin((double)s, carried) : jmax((double)s, carried));
        } else dst[r] = s;
    }
    PB_DEV void gathered(int r, float s, int b, int deg, const RowIds& row) const {
        double val = 0, w = 0;
        float wk[PB_ROW_FAST], v[PB_ROW_FAST];
#pragma unroll
        for (int k = 0; k < PB_ROW_FAST; k++) { wk[k] = k < deg ? wt[b + k] : 0.0f; v[k] = src[row.nb[k]]; }
#pragma unroll
        for (int k = 0; k < PB_ROW_FAST; k++)
            if (wk[k] > 0) {
                const double vv = v[k];
                if (sign < 0 ? (vv < 0) : (vv > 0)) { val += vv * (double)wk[k]; w += (double)wk[k]; }
            }
        finish(r, s, val, w);
    }
    PB_DEV void row(int r, int b, int deg, const int* ids) const {
        const float s = src[r];
        if (!isLand[r]) { dst[r] = s; return; }
        if (deg <= PB_ROW_FAST) { RowIds rw; rw.load(r, deg, ids); gathered(r, s, b, deg, rw); return; }
        double val = 0, w = 0;
        for (int j = 0; j < deg; j++) {
            const float wj = wt[b + j];
            if (wj > 0) {
                const double v = src[ids[j]];
                if (sign < 0 ? (v < 0) : (v > 0)) { val += v * (double)wj; w += (double)wj; }
            }
        }
        finish(r, s, val, w);
    }
};
// The same sweep over a COMPACTED list of the land rows.  Fibonacci ids advance by the golden angle in longitude, so land and
// ocean rows alternate pseudo-randomly in id order: the row-indexed kernel above pulls in almost every cache line of the
// offsets, packed rows and edge weights although only the land rows (≈ 27 %) use them — 500 MB of DRAM traffic per sweep at
// 10M cells for 262 MB of algorithmic bytes (profiles/r02_sweeps_ncu.md).  The reference compacts its land edge lists for the
// same reason (js/precipitation.js:520-547).  Here every land row owns one contiguous 48-byte record — its packed neighbour
// word and eight edge weights — and ocean rows are never visited: both ping-pong buffers start as the initial field, and an
// ocean cell keeps that value through the whole propagation.
struct LandPackK { const int* landRow; const PackedRow* pack; PackedRow* out; PB_DEV void operator()(int i) const { out[i] = pack[landRow[i]]; } };
struct LandWeightsK {
    Csr g; const int* landRow; const float* wt; float* out;
    PB_DEV void operator()(int i) const {
        const int r = landRow[i];
        const int b = g.off[r], deg = g.off[r + 1] - b;
        for (int k = 0; k < PB_ROW_FAST; k++) out[(size_t)PB_ROW_FAST * i + k] = k < deg ? wt[b + k] : 0.0f;
    }
};
struct LandIndexK { const int* landRow; int* landIndex; PB_DEV void operator()(int i) const { landIndex[landRow[i]] = i; } };
struct LandRangeK {      // out[0], out[1] = first land item with row >= lo, >= hi (the item range of a rank's cell-id range)
    const int* landRow; int n; int lo, hi; int* out;
    PB_DEV int lower(int key) const { int a = 0, b = n; while (a < b) { const int m = (a + b) >> 1; if (landRow[m] < key) a = m + 1; else b = m; } return a; }
    PB_DEV void operator()(int) const { out[0] = lower(lo); out[1] = lower(hi); }
};
struct ShadowLandK {
    static constexpr bool kItems = true;
    Csr g; const int* landRow; const PackedRow* landPack; const float* landWt; const float* wtCsr; const uint8_t* isLand;
    const float* src; float* dst; double keepFactor; int sign; const int* landIndex;
    PB_DEV int row_of(int i) const { return landRow[i]; }
    PB_DEV void by_row(int r) const { const int i = landIndex[r]; if (i >= 0) (*this)(i); }
    PB_DEV void operator()(int i) const {
        const int r = landRow[i];
        RowIds row; int deg;
        if (!row.load_word(landPack + i, r, deg)) {       // escape row: the CSR form
            const int b = g.off[r];
            ShadowSweepK{g, isLand, wtCsr, src, dst, keepFactor, sign}.row(r, b, g.off[r + 1] - b, g.adj + b);
            return;
        }
        const float s = src[r];
        float wk[PB_ROW_FAST], v[PB_ROW_FAST];
#if PB_CUDA
        const float4 w0 = __ldg((const float4*)(landWt + (size_t)PB_ROW_FAST * i)), w1 = __ldg((const float4*)(landWt + (size_t)PB_ROW_FAST * i) + 1);
        wk[0] = w0.x; wk[1] = w0.y; wk[2] = w0.z; wk[3] = w0.w; wk[4] = w1.x; wk[5] = w1.y; wk[6] = w1.z; wk[7] = w1.w;
#else
        for (int k = 0; k < PB_ROW_FAST; k++) wk[k] = landWt[(size_t)PB_ROW_FAST * i + k];
#endif
#pragma unroll
        for (int k = 0; k < PB_ROW_FAST; k++) v[k] = src[row.nb[k]];
        double val = 0, w = 0;
#pragma unroll
        for (int k = 0; k < PB_ROW_FAST; k++)
            if (wk[k] > 0) {
                const double vv = v[k];
                if (sign < 0 ? (vv < 0) : (vv > 0)) { val += vv * (double)wk[k]; w += (double)wk[k]; }
            }
        ShadowSweepK{g, isLand, wtCsr, src, dst, keepFactor, sign}.finish(r, s, val, w);
    }
};
struct KeepExtremeK {   // :572-574 / :599-601
    const float* src; float* field; int sign;
    PB_DEV void operator()(int r) const { if (sign < 0 ? (src[r] < field[r]) : (src[r] > field[r])) field[r] = src[r]; }
};
struct MergeShadowK { const float* shadow; const float* windward; float* out; PB_DEV void operator()(int r) const { out[r] = shadow[r] < 0 ? shadow[r] : windward[r]; } };
struct ApplyShadowK {   // :616-627
    const uint8_t* isLand; const float* rainShadow; float* precip;
    PB_DEV void operator()(int r) const {
        if (!isLand[r]) return;
        const double rs = rainShadow[r];
        if (rs < -0.01) precip[r] = (float)((double)precip[r] * jmax(0.02, 1 - jmin(1, -rs * 2.25) * 0.92));
        else if (rs > 0.01) precip[r] = (float)((double)precip[r] + rs * 1.2);
    }
};
struct BlendPrecipK {   // :652-654 + selection key
    const float* complex_; const float* heur; float* out; uint32_t* key;
    PB_DEV void operator()(int r) const {
        const float v = (float)(0.5 * (double)complex_[r] + 0.5 * (double)heur[r]);
        out[r] = v; key[r] = f32_sort_key(v);
    }
};
struct NormalizeCapK {   // :658-676
    float* x; const double* scale; const uint8_t* isLand; const float* cont;
    PB_DEV void operator()(int r) const {
        float v = (float)jmin(1, (double)x[r] / *scale);
        if (isLand[r] && (double)cont[r] > 0.5) {
            const double t = smoothstep(0.5, 1.0, cont[r]);
            const double cap = 1.0 - t * 0.80;
            v = (float)jmin((double)v, cap);
        }
        x[r] = v;
    }
};

// ---- temperature.js ----------------------------------------------------------------------------------------------------------
struct CoastalSeedK { const uint8_t* isLand; const float* warmth; float* out; PB_DEV void operator()(int r) const { out[r] = isLand[r] ? 0.0f : warmth[r]; } };
struct TemperatureK {   // :106-212
    const float* lat; const float* lon; const uint8_t* isLand; const float* elev; const float* cont; const float* pcont;
    const float* itczLats; const float* warmth; const float* speed; const float* precip; const float* coastalWarmth;
    int summer; double temperatureOffset; float* temp;
    PB_DEV void operator()(int r) const {
        const double la = lat[r];
        const bool land = isLand[r] != 0;
        const double el = elev[r], c = cont[r], pc = pcont[r];
        const double tropicalHW = 13;
        const double maxDist = 90 - tropicalHW;
        const double itczLat = itcz_lookup(itczLats, lon[r]);
        const double distItcz = fabs(la - itczLat) / PB_DEG;
        const double tItcz = jmax(0, distItcz - tropicalHW) / maxDist;
        const double T_itcz = 28 - 47 * pb_pow(tItcz, 1.4);
        const double flatItczLat = (summer ? 5 : -5) * PB_DEG;
        const double distFlat = fabs(la - flatItczLat) / PB_DEG;
        const double tFlat = jmax(0, distFlat - tropicalHW) / maxDist;
        const double T_flat = 28 - 47 * pb_pow(tFlat, 1.4);
        const double absLatDeg = fabs(la) / PB_DEG;
        const double blend = smoothstep(45, 90, absLatDeg);
        double T = T_itcz * (1 - blend) + T_flat * blend;
        const double moisture = precip[r];
        const double lapse = 4.5 + 4.8 * (1 - moisture);
        if (land && el > 0) T -= lapse * elev_to_height_km(el);
        if (!land) T += (double)warmth[r] * jmin(1, (double)speed[r] * 2) * 16;
        else {
            const double cw = coastalWarmth[r];
            if (fabs(cw) > 0.001) T += cw * (1 - smoothstep(0, 0.95, pc)) * 20;
        }
        if (moisture > 0.5) T *= (1 - smoothstep(0.5, 1.0, moisture) * 0.15);
        else if (moisture < 0.3) T *= (1 + smoothstep(0.3, 0.0, moisture) * 0.15);
        {
            const double distAnn = fabs(la) / PB_DEG;
            const double tAnn = jmax(0, distAnn - tropicalHW) / maxDist;
            const double T_annual = 28 - 47 * pb_pow(tAnn, 1.4);
            const double T_ann_adj = land && el > 0 ? T_annual - lapse * elev_to_height_km(el) : T_annual;
            const double deviation = T - T_ann_adj;
            const double seasonalBoost = 12 * smoothstep(10, 55, distAnn) * (1 - smoothstep(75, 90, distAnn));
            const bool isLocalSummer = summer ? (la >= 0) : (la < 0);
            const double seasonSign = isLocalSummer ? 1 : -1;
            const double boostedDeviation = deviation + seasonSign * seasonalBoost;
            const double maritimeFactor = 0.50 + c * 0.70;
            T = T_ann_adj + boostedDeviation * maritimeFactor;
        }
        T += temperatureOffset;
        temp[r] = (float)T;
    }
};
struct TempNormalizeK { float* t; PB_DEV void operator()(int r) const { t[r] = (float)jmax(0, jmin(1, ((double)t[r] - (-45.0)) / 90.0)); } };

// ---- koppen.js:67-288 ------------------------------------------------------------------------------------------------------------
struct KoppenK {
    const float* elev; const float* tSummer; const float* tWinter; const float* pSummer; const float* pWinter; uint8_t* out;
    PB_DEV void operator()(int r) const {
        enum { Ocean, Af, Am, Aw, BWh, BWk, BSh, BSk, Cfa, Cfb, Cfc, Csa, Csb, Csc, Cwa, Cwb, Cwc, Dfa, Dfb, Dfc, Dfd,
               Dsa, Dsb, Dsc, Dsd, Dwa, Dwb, Dwc, Dwd, ET, EF };
        if (elev[r] <= 0) { out[r] = Ocean; return; }
        const double Ts = -45 + jmax(0, jmin(1, (double)tSummer[r])) * 90;
        const double Tw = -45 + jmax(0, jmin(1, (double)tWinter[r])) * 90;
        const double Thot = jmax(Ts, Tw), Tcold = jmin(Ts, Tw);
        const double Tann = (Ts + Tw) / 2;
        const double Tshoulder = Thot - (Thot - Tcold) * (2.0 / 6);
        const bool localSummerIsSim = Ts >= Tw;
        const double Ps = jmax(0, (double)pSummer[r]) * 1000, Pw = jmax(0, (double)pWinter[r]) * 1000;
        const double Pann = Ps + Pw;
        const double PsummerLocal = localSummerIsSim ? Ps : Pw;
        const double PwinterLocal = localSummerIsSim ? Pw : Ps;
        const double PsMonthLocal = PsummerLocal / 6, PwMonthLocal = PwinterLocal / 6;
        const double Pdry = jmin(PsMonthLocal, PwMonthLocal);
        int band;   // 0 A, 1 C, 2 D
        if (Thot < 0) { out[r] = EF; return; }
        else if (Thot < 10) { out[r] = ET; return; }
        else if (Tcold >= 18) band = 0;
        else if (Tcold >= 0) band = 1;
        else band = 2;
        double Pthresh;
        const double summerFrac = Pann > 0 ? PsummerLocal / Pann : 0.5;
        if (summerFrac >= 0.7) Pthresh = 20 * Tann + 280;
        else if (summerFrac <= 0.3) Pthresh = 20 * Tann;
        else Pthresh = 20 * Tann + 140;
        Pthresh = jmax(0, Pthresh);
        if (Pann < Pthresh) {
            const bool isHot = Tann >= 18;
            if (Pann < Pthresh * 0.5) out[r] = isHot ? BWh : BWk;
            else out[r] = isHot ? BSh : BSk;
            return;
        }
        int pat;   // 0 f, 1 s, 2 w
        const bool localSummerDrier = PsummerLocal < PwinterLocal;
        if (localSummerDrier && PsMonthLocal < 50 && PsMonthLocal < PwMonthLocal / 2) pat = 1;
        else if (!localSummerDrier && PwMonthLocal < PsMonthLocal / 10) pat = 2;
        else pat = 0;
        int tl;    // 0 a, 1 b, 2 c, 3 d
        if (Thot >= 22) tl = 0;
        else if (Tshoulder >= 10) tl = 1;
        else if (Tcold >= -38) tl = 2;
        else tl = 3;
        if (band == 0) {
            if (Pdry >= 60) out[r] = Af;
            else if (Pann >= 25 * (100 - Pdry)) out[r] = Am;
            else out[r] = Aw;
            return;
        }
        if (band == 1) { out[r] = tl < 3 ? (uint8_t)(Cfa + 3 * pat + tl) : (uint8_t)Cfb; return; }
        out[r] = (uint8_t)(Dfa + 4 * pat + tl);
    }
};

}  // namespace pb
