// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).  The reference has no tests or golden
// vectors and no JS runtime exists in this image; this file restates js/elevation.js line by line and is pinned against that
// source executed under tests/golden/minijs.py (tests/test_zz_reference_vectors.py: elevation, stress, sets, debug layers).
//
//   plateVelocityAt :11-21, findCollisions :27-122, propagateStress :127-159,
//   assignDistanceField :164-189, assignElevation :216-1391
// Single thread, double arithmetic with f32 typed-array stores (every `r_elevation[r] += x` rounds to
// f32, exactly like the Float32Array it is), Sets iterate in insertion order, transcendentals through
// include/pb_detmath.h.
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>
#include "js_semantics.h"
#include "noise.h"

namespace {

typedef std::vector<float> F32;
typedef std::vector<int32_t> I32;
typedef std::vector<uint8_t> U8;
const float INF32 = INFINITY;

// JS object keyed by plate id: {pid: {pole, omega}}, {pid: density}, plus a Set of oceanic ids
struct Plates {
    std::unordered_map<int, int> index;
    std::vector<double> pole, omega, density;
    std::vector<uint8_t> isOcean;
    Plates(int n, const int32_t* ids, const uint8_t* oc, const double* p, const double* om, const double* de) {
        for (int k = 0; k < n; k++) index[ids[k]] = k;
        pole.assign(p, p + 3 * n); omega.assign(om, om + n); density.assign(de, de + n); isOcean.assign(oc, oc + n);
    }
    int find(int id) const { auto it = index.find(id); return it == index.end() ? -1 : it->second; }
    bool ocean(int id) const { const int k = find(id); return k >= 0 && isOcean[k]; }
    double dens(int id) const { const int k = find(id); return k >= 0 ? density[k] : NAN; }
};

// insertion-ordered Set of cell ids
struct CellSet {
    std::vector<int> items;
    U8 in;
    explicit CellSet(int N = 0) : in(N, 0) {}
    bool has(int r) const { return in[r] != 0; }
    void add(int r) { if (!in[r]) { in[r] = 1; items.push_back(r); } }
};

// js/elevation.js:11-21
void plateVelocityAt(const Plates& P, int k, double x, double y, double z, double v[3]) {
    const double px = P.pole[3 * k], py = P.pole[3 * k + 1], pz = P.pole[3 * k + 2], omega = P.omega[k];
    v[0] = omega * (py * z - pz * y);
    v[1] = omega * (pz * x - px * z);
    v[2] = omega * (px * y - py * x);
}

// js/elevation.js:44-53 (the cache only memoises)
double getPairIntensity(int a, int b) {
    const double lo = std::min(a, b), hi = std::max(a, b);
    uint32_t h = (uint32_t)(js::to_int32(lo * 16807) ^ js::to_int32(hi * 48271));
    const int32_t x = (js::to_int32((double)h) >> 16) ^ js::to_int32((double)h);
    h = js::to_uint32((double)x * (double)0x45d9f3b);
    return 0.5 + (double)(h % 10001u) / 10000;
}

struct Collisions {
    CellSet mountain_r, coastline_r, ocean_r;
    F32 r_stress, r_subductFactor;
    std::vector<int8_t> r_boundaryType;
    U8 r_bothOcean, r_hasOcean;
};

// js/elevation.js:27-122
Collisions findCollisions(const OMesh& mesh, const float* xyz, const Plates& P, const int32_t* r_plate, const SimplexNoise& noise) {
    const int N = mesh.N;
    const double dt = 1e-2 / js::max(1, std::sqrt(N / 10000.0));
    Collisions c;
    c.mountain_r = CellSet(N); c.coastline_r = CellSet(N); c.ocean_r = CellSet(N);
    c.r_stress.assign(N, 0.f); c.r_subductFactor.assign(N, 0.5f); c.r_boundaryType.assign(N, 0);
    c.r_bothOcean.assign(N, 0); c.r_hasOcean.assign(N, 0);
    const int undulOctaves = N > 200000 ? 2 : 3;
    for (int r = 0; r < N; r++) {
        const int myPlate = r_plate[r];
        double bestComp = -INFINITY, bestNormalComp = 0;
        int best = -1;
        for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
            const int nb = mesh.adjList[ni];
            if (myPlate != r_plate[nb]) {
                const int ri3 = 3 * r, ni3 = 3 * nb;
                const double dx = (double)xyz[ri3] - xyz[ni3], dy = (double)xyz[ri3 + 1] - xyz[ni3 + 1], dz = (double)xyz[ri3 + 2] - xyz[ni3 + 2];
                const double dBefore = std::sqrt(dx * dx + dy * dy + dz * dz);
                double v1[3], v2[3];
                plateVelocityAt(P, P.find(myPlate), xyz[ri3], xyz[ri3 + 1], xyz[ri3 + 2], v1);
                plateVelocityAt(P, P.find(r_plate[nb]), xyz[ni3], xyz[ni3 + 1], xyz[ni3 + 2], v2);
                const double ax = xyz[ri3] + v1[0] * dt, ay = xyz[ri3 + 1] + v1[1] * dt, az = xyz[ri3 + 2] + v1[2] * dt;
                const double bx = xyz[ni3] + v2[0] * dt, by = xyz[ni3 + 1] + v2[1] * dt, bz = xyz[ni3 + 2] + v2[2] * dt;
                const double adx = ax - bx, ady = ay - by, adz = az - bz;
                const double dAfter = std::sqrt(adx * adx + ady * ady + adz * adz);
                const double comp = dBefore - dAfter;
                if (comp > bestComp) {
                    bestComp = comp; best = nb;
                    const double rvx = v1[0] - v2[0], rvy = v1[1] - v2[1], rvz = v1[2] - v2[2];
                    const double bnLen = js::or_default(dBefore, 1);
                    bestNormalComp = -(rvx * dx + rvy * dy + rvz * dz) / bnLen;
                }
            }
        }
        if (best != -1) {
            const bool collided = bestComp > 0.75 * dt;
            const bool rOcean = P.ocean(myPlate), nOcean = P.ocean(r_plate[best]);
            c.r_bothOcean[r] = (rOcean && nOcean) ? 1 : 0;
            c.r_hasOcean[r] = (rOcean || nOcean) ? 1 : 0;
            const double thresh = 0.3 * dt;
            if (bestNormalComp > thresh) c.r_boundaryType[r] = 1;
            else if (bestNormalComp < -thresh) c.r_boundaryType[r] = 2;
            else c.r_boundaryType[r] = 3;
            if (collided) c.r_stress[r] = js::f32((bestComp / dt) * getPairIntensity(myPlate, r_plate[best]));
            const double densityDiff = P.dens(myPlate) - P.dens(r_plate[best]);
            const double baseFactor = 0.5 + 0.5 * pb_tanh(densityDiff * 8);
            const double densityContrast = std::fabs(densityDiff);
            const double undulationStrength = pb_exp(-densityContrast * 12);
            const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
            const double undulation = noise.fbm(x * 6, y * 6, z * 6, undulOctaves) * 0.4 * undulationStrength;
            c.r_subductFactor[r] = js::f32(js::max(0, js::min(1, baseFactor + undulation)));
            if (rOcean && nOcean) (collided ? c.coastline_r : c.ocean_r).add(r);
            else if (!rOcean && !nOcean) {
                if (collided) { if (c.r_subductFactor[r] < 0.55) c.mountain_r.add(r); else c.coastline_r.add(r); }
            } else (collided ? c.mountain_r : c.coastline_r).add(r);
        }
    }
    return c;
}

// js/elevation.js:127-159
void propagateStress(const OMesh& mesh, F32& r_stress, F32& r_subductFactor, const int32_t* r_plate, const Plates& P,
                     double decayFactor, double subductDecayFactor, int numPasses) {
    std::vector<int> frontier;
    for (int r = 0; r < mesh.N; r++) if (r_stress[r] > 0.01) frontier.push_back(r);
    for (int pass = 0; pass < numPasses && !frontier.empty(); pass++) {
        std::vector<int> next;
        for (size_t fi = 0; fi < frontier.size(); fi++) {
            const int r = frontier[fi];
            const int plate = r_plate[r];
            if (P.ocean(plate)) continue;
            const float sf = r_subductFactor[r];
            const double effDecay = sf > 0.5 ? subductDecayFactor : decayFactor;
            const double propagated = r_stress[r] * effDecay;
            if (propagated < 0.005) continue;
            for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                const int nb = mesh.adjList[ni];
                if (r_plate[nb] == plate && propagated > r_stress[nb]) {
                    r_stress[nb] = js::f32(propagated);
                    r_subductFactor[nb] = sf;
                    next.push_back(nb);
                }
            }
        }
        frontier.swap(next);
    }
}

// js/elevation.js:164-189
F32 assignDistanceField(const OMesh& mesh, const std::vector<int>& seeds, const U8& isStop, double seed) {
    RandInt randInt(seed);
    F32 r_dist(mesh.N, INF32);
    std::vector<int> queue;
    for (int r : seeds) { queue.push_back(r); r_dist[r] = 0; }
    for (size_t qi = 0; qi < queue.size(); qi++) {
        const size_t pos = qi + (size_t)randInt((double)(queue.size() - qi));
        const int cur = queue[pos];
        queue[pos] = queue[qi];
        for (int ni = mesh.adjOffset[cur]; ni < mesh.adjOffset[cur + 1]; ni++) {
            const int nb = mesh.adjList[ni];
            if (r_dist[nb] == INF32 && !isStop[nb]) { r_dist[nb] = js::f32((double)r_dist[cur] + 1); queue.push_back(nb); }
        }
    }
    return r_dist;
}

struct Dome {
    double x, y, z, strength, baseStrength, sigma; int chainIndex, chainLength; double dx, dy, dz, ux, uy, uz, vx, vy, vz;
    std::vector<double> riftAngles;
    double cosThreshPeak, invS2, swellSigma, swellStrength, cosThreshSwell, invS2Swell, driftStretch, calderaSigma, calderaDepth, invS2Caldera, ageFactor;
    bool hasCaldera;
};

}  // namespace

struct OracleElevation {
    OMesh mesh; int N; const float* xyz;
    std::map<std::string, F32> f;
    std::map<std::string, I32> i;
    std::map<std::string, U8> u;
    OracleElevation(const OMesh& m, const float* x) : mesh(m), N(m.N), xyz(x) {}

    // js/elevation.js:216-1391
    void assignElevation(const Plates& P, const int32_t* r_plate, const int32_t* plateSeeds, int nSeeds, double noiseSeed,
                         double noiseMag, double seed, double spread, const Plates* SP, const int32_t* r_superPlate) {
        const SimplexNoise noise(noiseSeed);
        F32 E(N, 0.f);
        F32 dl_base(N, 0.f), dl_tectonic(N, 0.f), dl_noise(N, 0.f), dl_interior(N, 0.f), dl_coastal(N, 0.f), dl_ocean(N, 0.f),
            dl_hotspot(N, 0.f), dl_tecActivity(N, 0.f), dl_margins(N, 0.f), dl_backArc(N, 0.f), dl_foldRidge(N, 0.f), dl_orogenicPower(N, 0.f);

        Collisions smallCol = findCollisions(mesh, xyz, P, r_plate, noise);
        const bool hasSuperPlates = SP != nullptr;
        Collisions superCol;
        if (hasSuperPlates) superCol = findCollisions(mesh, xyz, *SP, r_superPlate, noise);

        CellSet mountain_r(N), coastline_r(N), ocean_r(N);
        F32 r_stress, r_subductFactor;
        std::vector<int8_t> r_boundaryType;
        U8 r_bothOcean, r_hasOcean;
        const double SMALL_W = 0.05, SUPER_W = 0.95;
        if (!hasSuperPlates) {
            mountain_r = smallCol.mountain_r; coastline_r = smallCol.coastline_r; ocean_r = smallCol.ocean_r;
            r_stress = smallCol.r_stress; r_subductFactor = smallCol.r_subductFactor; r_boundaryType = smallCol.r_boundaryType;
            r_bothOcean = smallCol.r_bothOcean; r_hasOcean = smallCol.r_hasOcean;
        } else {
            for (int r : superCol.mountain_r.items) mountain_r.add(r);
            for (int r : smallCol.mountain_r.items) mountain_r.add(r);
            for (int r : superCol.ocean_r.items) ocean_r.add(r);
            for (int r : smallCol.ocean_r.items) ocean_r.add(r);
            for (int r : superCol.coastline_r.items) if (!mountain_r.has(r)) coastline_r.add(r);
            for (int r : smallCol.coastline_r.items) if (!mountain_r.has(r) && !coastline_r.has(r)) coastline_r.add(r);
            r_stress.assign(N, 0.f);
            {
                double maxSuperStress = 0;
                for (int r = 0; r < N; r++) if (superCol.r_stress[r] > maxSuperStress) maxSuperStress = superCol.r_stress[r];
                const double invMax = maxSuperStress > 1e-6 ? 1 / maxSuperStress : 0;
                for (int r = 0; r < N; r++) {
                    const double sS = smallCol.r_stress[r], sP = superCol.r_stress[r];
                    const double proximity = js::min(1, sP * invMax * 3);
                    const double effectiveSmallW = SMALL_W * (SMALL_W + (1 - SMALL_W) * proximity);
                    r_stress[r] = js::f32(effectiveSmallW * sS + SUPER_W * sP);
                }
            }
            r_subductFactor.assign(N, 0.f);
            for (int r = 0; r < N; r++) {
                const double wS = SMALL_W * smallCol.r_stress[r], wP = SUPER_W * superCol.r_stress[r];
                const double total = wS + wP;
                if (total > 1e-6) r_subductFactor[r] = js::f32((wS * smallCol.r_subductFactor[r] + wP * superCol.r_subductFactor[r]) / total);
                else r_subductFactor[r] = js::f32(SMALL_W * smallCol.r_subductFactor[r] + SUPER_W * superCol.r_subductFactor[r]);
            }
            r_boundaryType.assign(N, 0);
            for (int r = 0; r < N; r++) {
                const double wS = SMALL_W * smallCol.r_stress[r], wP = SUPER_W * superCol.r_stress[r];
                r_boundaryType[r] = wS > wP ? smallCol.r_boundaryType[r] : superCol.r_boundaryType[r];
            }
            r_bothOcean.assign(N, 0); r_hasOcean.assign(N, 0);
            for (int r = 0; r < N; r++) {
                r_bothOcean[r] = smallCol.r_bothOcean[r] | superCol.r_bothOcean[r];
                r_hasOcean[r] = smallCol.r_hasOcean[r] | superCol.r_hasOcean[r];
            }
        }

        const double scaleFactor = std::sqrt(N / 10000.0);
        const double baseDecay = 0.5 + spread * 0.04;
        const double decayFactor = pb_pow(baseDecay, 1 / scaleFactor);
        const double subductBaseDecay = baseDecay * 0.45;
        const double subductDecayFactor = pb_pow(subductBaseDecay, 1 / scaleFactor);
        const int numPasses = (int)js::max(1, js::round(spread * 3 * scaleFactor));
        if (!hasSuperPlates) {
            propagateStress(mesh, r_stress, r_subductFactor, r_plate, P, decayFactor, subductDecayFactor, numPasses);
        } else {
            F32 smallStress(smallCol.r_stress), smallSubduct(smallCol.r_subductFactor);
            propagateStress(mesh, smallStress, smallSubduct, r_plate, P, decayFactor, subductDecayFactor, numPasses);
            F32 superStress(superCol.r_stress), superSubduct(superCol.r_subductFactor);
            propagateStress(mesh, superStress, superSubduct, r_superPlate, *SP, decayFactor, subductDecayFactor, numPasses);
            for (int r = 0; r < N; r++) r_stress[r] = js::f32(SMALL_W * smallStress[r] + SUPER_W * superStress[r]);
            for (int r = 0; r < N; r++) {
                const double wS = SMALL_W * smallStress[r], wP = SUPER_W * superStress[r];
                const double total = wS + wP;
                if (total > 1e-6) r_subductFactor[r] = js::f32((wS * smallSubduct[r] + wP * superSubduct[r]) / total);
            }
        }

        // plate representatives (:368-382)
        {
            std::unordered_map<int, int> plateRep;
            for (int r = 0; r < N; r++) {
                const int pid = r_plate[r];
                if (!plateRep.count(pid) && !mountain_r.has(r) && !coastline_r.has(r) && !ocean_r.has(r)) plateRep[pid] = r;
            }
            for (int k = 0; k < nSeeds; k++) {
                const int pid = plateSeeds[k];
                auto it = plateRep.find(pid);
                if (it != plateRep.end()) (P.ocean(pid) ? ocean_r : coastline_r).add(it->second);
            }
        }
        CellSet stress_mountain_r(N);
        for (int r : mountain_r.items) if (r_subductFactor[r] < 0.55) stress_mountain_r.add(r);
        U8 stop_r(N, 0);
        for (int r : stress_mountain_r.items) stop_r[r] = 1;
        for (int r : coastline_r.items) stop_r[r] = 1;
        for (int r : ocean_r.items) stop_r[r] = 1;

        F32 dist_mountain = assignDistanceField(mesh, stress_mountain_r.items, ocean_r.in, seed + 1);
        F32 dist_ocean = assignDistanceField(mesh, ocean_r.items, coastline_r.in, seed + 2);
        F32 dist_coastline = assignDistanceField(mesh, coastline_r.items, stop_r, seed + 3);

        U8 r_isOcean(N, 0);
        for (int r = 0; r < N; r++) if (P.ocean(r_plate[r])) r_isOcean[r] = 1;
        CellSet coastSeeds(N);
        for (int r = 0; r < N; r++)
            if (!r_isOcean[r])
                for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++)
                    if (r_isOcean[mesh.adjList[ni]]) { coastSeeds.add(mesh.adjList[ni]); break; }
        F32 dist_coast = assignDistanceField(mesh, coastSeeds.items, U8(N, 0), seed + 4);
        CellSet landCoastSeeds(N);
        for (int r = 0; r < N; r++) {
            if (r_isOcean[r]) continue;
            for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++)
                if (r_isOcean[mesh.adjList[ni]]) { landCoastSeeds.add(r); break; }
        }
        F32 dist_coast_land = assignDistanceField(mesh, landCoastSeeds.items, r_isOcean, seed + 5);

        const double interiorBand = js::max(4, js::round(16 * scaleFactor));
        const double tectonicReach = js::max(6, js::round(20 * scaleFactor));

        double maxStress = 0;
        {
            std::vector<float> stressVals;
            for (int r = 0; r < N; r++) {
                if (r_stress[r] > 0.01) stressVals.push_back(r_stress[r]);
                if (r_stress[r] > maxStress) maxStress = r_stress[r];
            }
            if (!stressVals.empty()) {
                std::sort(stressVals.begin(), stressVals.end());
                const size_t k = std::min(stressVals.size() - 1, (size_t)std::floor(stressVals.size() * 0.97));
                maxStress = stressVals[k];
            }
            if (maxStress < 0.01) maxStress = 1;
        }
        const double eps = 1e-3, warpScale = 0.4;
        const int warpOctaves = N > 200000 ? 2 : 3;
        const double plateauStart = js::max(2, js::round(3 * scaleFactor));

        // coast-boundary BFS (:464-509)
        std::vector<int> coastBdry;
        for (int r = 0; r < N; r++) {
            const uint8_t rOc = r_isOcean[r];
            for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++)
                if (r_isOcean[mesh.adjList[ni]] != rOc) { coastBdry.push_back(r); break; }
        }
        const double maxCD = js::max(8, js::round(8 * scaleFactor));
        F32 dBdry(N, js::f32(maxCD + 1)), coastStressMax(N, 0.f), coastSubductMax(N, 0.f);
        U8 coastConvergent(N, 0);
        for (int r : coastBdry) {
            dBdry[r] = 0;
            coastStressMax[r] = js::f32(js::min(1, r_stress[r] / maxStress));
            coastSubductMax[r] = r_subductFactor[r];
            coastConvergent[r] = r_boundaryType[r] == 1 ? 1 : 0;
        }
        for (size_t qi = 0; qi < coastBdry.size();) {
            const int r = coastBdry[qi++];
            const double nd = (double)dBdry[r] + 1;
            if (nd > maxCD) continue;
            for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                const int nr = mesh.adjList[ni];
                if (nd < dBdry[nr]) {
                    dBdry[nr] = js::f32(nd);
                    coastStressMax[nr] = coastStressMax[r]; coastSubductMax[nr] = coastSubductMax[r]; coastConvergent[nr] = coastConvergent[r];
                    coastBdry.push_back(nr);
                } else if (nd == dBdry[nr] && coastStressMax[r] > coastStressMax[nr]) {
                    coastStressMax[nr] = coastStressMax[r]; coastSubductMax[nr] = coastSubductMax[r]; coastConvergent[nr] = coastConvergent[r];
                }
            }
        }
        // capped FIFO BFS helper for rift / ridge / fracture (no payload)
        auto cappedBfs = [&](F32& dist, std::vector<int>& q, double cap, int mode) {   // mode 0 rift, 1 ocean-only
            for (size_t qi = 0; qi < q.size();) {
                const int r = q[qi++];
                const double nd = (double)dist[r] + 1;
                if (nd > cap) continue;
                const int plate = r_plate[r];
                for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                    const int nr = mesh.adjList[ni];
                    const bool ok = mode == 0 ? (r_plate[nr] == plate && !r_isOcean[nr]) : (r_isOcean[nr] != 0);
                    if (nd < dist[nr] && ok) { dist[nr] = js::f32(nd); q.push_back(nr); }
                }
            }
        };
        const double riftHalfWidth = js::max(2, js::round(4 * scaleFactor));
        F32 riftDist(N, INF32);
        { std::vector<int> q; for (int r = 0; r < N; r++) if (r_boundaryType[r] == 2 && !r_hasOcean[r]) { q.push_back(r); riftDist[r] = 0; } cappedBfs(riftDist, q, riftHalfWidth, 0); }
        const SimplexNoise riftNoise(seed + 419);
        const double ridgeHalfWidth = js::max(2, js::round(4 * scaleFactor));
        F32 ridgeDist(N, INF32);
        { std::vector<int> q; for (int r = 0; r < N; r++) if (r_boundaryType[r] == 2 && r_bothOcean[r]) { q.push_back(r); ridgeDist[r] = 0; } cappedBfs(ridgeDist, q, ridgeHalfWidth, 1); }
        const double fractureHalfWidth = js::max(2, js::round(3 * scaleFactor));
        F32 fractureDist(N, INF32);
        { std::vector<int> q; for (int r = 0; r < N; r++) if (r_boundaryType[r] == 3 && r_bothOcean[r]) { q.push_back(r); fractureDist[r] = 0; } cappedBfs(fractureDist, q, fractureHalfWidth, 1); }
        // back-arc BFS with payload (:601-631)
        const double baStart = js::max(1, js::round(2 * scaleFactor)), baPeak = js::max(2, js::round(3 * scaleFactor)), baEnd = js::max(3, js::round(5 * scaleFactor));
        F32 backArcDist(N, INF32), backArcStress(N, 0.f);
        {
            std::vector<int> q;
            for (int r = 0; r < N; r++)
                if (r_boundaryType[r] == 1 && r_hasOcean[r] && r_subductFactor[r] < 0.50) {
                    q.push_back(r); backArcDist[r] = 0; backArcStress[r] = js::f32(js::min(1, r_stress[r] / maxStress));
                }
            for (size_t qi = 0; qi < q.size();) {
                const int r = q[qi++];
                const double nd = (double)backArcDist[r] + 1;
                if (nd > baEnd) continue;
                const int plate = r_plate[r];
                for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                    const int nr = mesh.adjList[ni];
                    if (nd < backArcDist[nr] && r_plate[nr] == plate) { backArcDist[nr] = js::f32(nd); backArcStress[nr] = backArcStress[r]; q.push_back(nr); }
                }
            }
        }
        const SimplexNoise foldNoise(seed + 557);
        auto add = [&](int r, double v) { E[r] = js::f32((double)E[r] + v); };

        // main per-cell loop (:638-973)
        for (int r = 0; r < N; r++) {
            const bool isOceanPlate = r_isOcean[r] != 0;
            const double sfAsym = r_subductFactor[r];
            const double asymmetry = 1.0 + (sfAsym - 0.5) * 0.8;
            const double a = dist_mountain[r] * asymmetry + eps;
            const double b = dist_ocean[r] + eps;
            const double c = dist_coastline[r] + eps;
            const double BASE_SCALE = 0.6;
            if (a == INFINITY && b == INFINITY) E[r] = js::f32(0.1 * BASE_SCALE);
            else E[r] = js::f32((1 / a - 1 / b) / (1 / a + 1 / b + 1 / c) * BASE_SCALE);
            dl_base[r] = E[r];
            const double stressNorm = js::min(1, r_stress[r] / maxStress);
            const int btype = r_boundaryType[r];
            const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
            const double wx = x + warpScale * noise.fbm(x + 5.3, y + 1.7, z + 3.1, warpOctaves);
            const double wy = y + warpScale * noise.fbm(x + 8.1, y + 2.9, z + 7.3, warpOctaves);
            const double wz = z + warpScale * noise.fbm(x + 1.4, y + 6.2, z + 4.8, warpOctaves);
            const double rawOro = noise.noise3D(x * 1.5 + 33.7, y * 1.5 + 11.2, z * 1.5 + 22.9);
            const double shaped = rawOro >= 0 ? std::sqrt(rawOro) : -std::sqrt(-rawOro);
            const double orogenicPower = js::max(0, js::min(1, 0.5 + 0.5 * shaped));
            dl_orogenicPower[r] = js::f32(orogenicPower - 0.5);

            if (!isOceanPlate) {
                const double sf = r_subductFactor[r];
                const float elevBefore = E[r];
                if (sf > 0.5 && E[r] > 0) { const double suppression = (sf - 0.5) * 2; E[r] = js::f32((double)E[r] * (1 - suppression * 0.42)); }
                if (stressNorm > 0.01) {
                    const double stressMag = stressNorm * stressNorm * 0.55 * orogenicPower;
                    const double uplift = stressMag * (1 - sf), depress = stressMag * 0.4 * sf;
                    const double heightVar = 0.60 + 0.8 * noise.fbm(x * 8 + 13.7, y * 8 + 9.2, z * 8 + 4.5, 3);
                    add(r, (uplift - depress) * heightVar);
                }
                if (stressNorm > 0 && stressNorm < 0.10) { const double forelandT = stressNorm / 0.10; add(r, -(0.06 * (1 - forelandT))); }
                {
                    const double rd = riftDist[r];
                    if (rd != INFINITY) {
                        const double floorEnd = js::max(1, js::round(1.5 * scaleFactor)), shoulderEnd = js::max(2, js::round(2.5 * scaleFactor));
                        double riftEffect = 0;
                        if (rd <= 0.5) { riftEffect = -0.15; riftEffect += riftNoise.ridgedFbm(x * 8, y * 8, z * 8, 3) * 0.04; }
                        else if (rd <= floorEnd) { const double t = rd / floorEnd; riftEffect = -0.12 * (1 - t * 0.3); riftEffect += riftNoise.ridgedFbm(x * 8, y * 8, z * 8, 3) * 0.03 * (1 - t); }
                        else if (rd <= shoulderEnd) { const double t = (rd - floorEnd) / (shoulderEnd - floorEnd); riftEffect = 0.03 * (1 - t); }
                        else if (riftHalfWidth > shoulderEnd) {
                            const double t = (rd - shoulderEnd) / (riftHalfWidth - shoulderEnd);
                            const double fadeT = js::min(1, t);
                            const double fade = fadeT * fadeT * (3 - 2 * fadeT);
                            riftEffect = 0.03 * (1 - fade) * 0.2;
                        }
                        add(r, riftEffect);
                    }
                }
                {
                    const double bad = backArcDist[r];
                    if (bad != INFINITY && bad >= baStart) {
                        const double dMtn = dist_mountain[r];
                        const double orogenyFactor = (dMtn != INFINITY && dMtn < bad) ? js::max(0, dMtn / bad) : 1.0;
                        double baEffect = 0;
                        if (bad <= baPeak) { const double t = (bad - baStart) / js::max(1, baPeak - baStart); const double s = t * t * (3 - 2 * t); baEffect = -0.10 * backArcStress[r] * s * orogenyFactor; }
                        else if (bad <= baEnd) { const double t = (bad - baPeak) / js::max(1, baEnd - baPeak); const double s = t * t * (3 - 2 * t); baEffect = -0.10 * backArcStress[r] * (1 - s) * orogenyFactor; }
                        add(r, baEffect);
                        dl_backArc[r] = js::f32(baEffect);
                    }
                }
                dl_tectonic[r] = js::f32((double)E[r] - (double)elevBefore);
                const double dMtn = dist_mountain[r];
                const double rawProximity = (dMtn == INFINITY || dMtn >= tectonicReach) ? 0 : (1 - dMtn / tectonicReach);
                const double tectonicActivity = js::max(stressNorm, rawProximity * rawProximity);
                dl_tecActivity[r] = js::f32(tectonicActivity);
                {
                    const int k = P.find(r_plate[r]);
                    const double foldActivity = tectonicActivity * tectonicActivity;
                    if (k >= 0 && foldActivity > 0.01) {
                        const double ppx = P.pole[3 * k], ppy = P.pole[3 * k + 1], ppz = P.pole[3 * k + 2];
                        const double uu = x * ppx + y * ppy + z * ppz;
                        const double phaseWarp = foldNoise.fbm(x * 3 + 55.3, y * 3 + 33.7, z * 3 + 17.2, 2) * 0.08;
                        const double phase = (uu + phaseWarp) * 30 * PB_PI;
                        const double ridge = 1 - std::fabs(pb_sin(phase));
                        const double foldCentered = ridge - 0.36;
                        const double ampMod = 0.6 + 0.4 * foldNoise.fbm(x * 4 + 88.1, y * 4 + 62.3, z * 4 + 41.7, 2);
                        const double elevBoost = 1 + 4 * js::max(0, E[r]);
                        const double foldAmp = foldActivity * js::max(0, 1 - sf * 1.5) * noiseMag * 0.8 * elevBoost;
                        const double foldContrib = foldCentered * foldAmp * ampMod;
                        add(r, foldContrib);
                        dl_foldRidge[r] = js::f32(foldContrib);
                    }
                }
                const bool isPlateauZone = sf < 0.45 && dMtn != INFINITY && dMtn > plateauStart;
                const double blend = js::min(1, stressNorm * 3);
                const double smoothNoise = noise.fbm(wx, wy, wz) * noiseMag;
                const double ridgedNoise = noise.ridgedFbm(wx, wy, wz) * noiseMag * 1.5;
                const double noiseVal = smoothNoise * (1 - blend) + ridgedNoise * blend;
                const double detailNoise = noise.fbm(wx * 4 + 22.1, wy * 4 + 6.8, wz * 4 + 15.4, 4, 0.5) * noiseMag * 0.5;
                const double noiseActivity = js::min(1, stressNorm * 4);
                const double plateauSuppress = isPlateauZone ? js::max(0.30, 1 - tectonicActivity * 0.60) : 1.0;
                const double noiseScale = (0.25 + 0.75 * noiseActivity) * plateauSuppress;
                const double fineNoise = noise.fbm(wx * 8 + 41.7, wy * 8 + 13.2, wz * 8 + 27.9, 3, 0.5) * noiseMag * 0.25;
                const double fineScale = std::sqrt(noiseScale);
                const double totalNoise = (noiseVal + detailNoise) * noiseScale + fineNoise * fineScale;
                add(r, totalNoise);
                dl_noise[r] = js::f32(totalNoise);
                {
                    const double currentElev = E[r];
                    if (currentElev > 0.12) {
                        const double elevExcess = currentElev - 0.12;
                        const double dissectVal = noise.fbm(wx * 16 + 71.3, wy * 16 + 44.8, wz * 16 + 29.1, 3, 0.5);
                        const double dissectAmp = std::sqrt(elevExcess) * stressNorm * noiseMag * 0.4;
                        const double dissectContrib = dissectVal * dissectAmp;
                        add(r, dissectContrib);
                        dl_noise[r] = js::f32((double)dl_noise[r] + dissectContrib);
                    }
                }
                {
                    const double currentElev = E[r];
                    if (currentElev > 0.65 && stressNorm > 0.2) {
                        const double excess = currentElev - 0.65;
                        const double peakNoise = noise.ridgedFbm(wx * 24 + 91.3, wy * 24 + 55.7, wz * 24 + 38.2, 3, 0.5);
                        const double spike = js::max(0, peakNoise - 0.45);
                        const double peakContrib = spike * excess * stressNorm * 1.2;
                        add(r, peakContrib);
                        dl_noise[r] = js::f32((double)dl_noise[r] + peakContrib);
                    }
                }
                const double lcd = dist_coast_land[r];
                if (lcd < INFINITY) {
                    const double tDown = js::min(lcd / interiorBand, 1);
                    const double sDown = tDown * tDown * (3 - 2 * tDown);
                    const double tUp = js::min(lcd / (interiorBand * 0.4), 1);
                    const double sUp = tUp * tUp * (3 - 2 * tUp);
                    const double interiorUplift = 0.06 + tectonicActivity * 0.16;
                    const double baseBias = -0.08 * (1 - sDown) + interiorUplift * sUp;
                    const double mod = 1.0 + 0.2 * noise.fbm(x * 2 + 19.3, y * 2 + 7.6, z * 2 + 13.1, 2);
                    const double bias = baseBias * mod;
                    add(r, bias);
                    dl_interior[r] = js::f32(bias);
                }
                if (isPlateauZone && tectonicActivity > 0.1) {
                    const double plateauBoost = 0.025 * tectonicActivity * (1 - sf);
                    add(r, plateauBoost);
                    dl_interior[r] = js::f32((double)dl_interior[r] + plateauBoost);
                }
            } else {
                const double dc = dist_coast[r];
                double oceanBase;
                if (dc < 5) oceanBase = -0.04 - 0.06 * (dc / 5);
                else if (dc < 12) oceanBase = -0.10 - 0.25 * ((dc - 5) / 7);
                else oceanBase = -0.35 + noise.fbm(x * 2, y * 2, z * 2, 3) * 0.03;
                E[r] = js::f32(js::min(E[r], oceanBase));
                dl_ocean[r] = E[r];
                const bool isActiveMargin = coastConvergent[r] == 1;
                dl_margins[r] = js::f32(isActiveMargin ? 0.8 : 0.2);
                if (ridgeDist[r] != INF32 && ridgeDist[r] <= ridgeHalfWidth) dl_margins[r] = 1.0f;
                if (fractureDist[r] != INF32 && fractureDist[r] <= fractureHalfWidth) dl_margins[r] = -0.5f;
                const float elevBeforeOcTec = E[r];
                const double rd = ridgeDist[r];
                if (rd != INFINITY && rd <= ridgeHalfWidth) {
                    const double t = rd / ridgeHalfWidth;
                    const double ridgeFade = (1 - t) * (1 - t);
                    const double ridgeNoise = noise.ridgedFbm(x * 3, y * 3, z * 3, 4);
                    add(r, (0.12 * ridgeNoise + 0.06) * ridgeFade);
                }
                const double fd = fractureDist[r];
                if (fd != INFINITY && fd <= fractureHalfWidth) { const double ft = fd / fractureHalfWidth; add(r, -(0.03 * (1 - ft))); }
                if (btype == 1) add(r, -(0.15 + 0.15 * stressNorm));
                {
                    const double bad = backArcDist[r];
                    if (bad != INFINITY && bad >= baStart) {
                        const double dMtn = dist_mountain[r];
                        const double orogenyFactor = (dMtn != INFINITY && dMtn < bad) ? js::max(0, dMtn / bad) : 1.0;
                        double baEffect = 0;
                        if (bad <= baPeak) { const double t = (bad - baStart) / js::max(1, baPeak - baStart); const double s = t * t * (3 - 2 * t); baEffect = -0.10 * backArcStress[r] * s * orogenyFactor; }
                        else if (bad <= baEnd) { const double t = (bad - baPeak) / js::max(1, baEnd - baPeak); const double s = t * t * (3 - 2 * t); baEffect = -0.10 * backArcStress[r] * (1 - s) * orogenyFactor; }
                        add(r, baEffect);
                        dl_backArc[r] = js::f32(baEffect);
                    }
                }
                dl_tectonic[r] = js::f32((double)E[r] - (double)elevBeforeOcTec);
                const double oceanNoise = noise.fbm(wx, wy, wz) * noiseMag * 0.3;
                add(r, oceanNoise);
                dl_noise[r] = js::f32(oceanNoise);
            }
        }

        // coastal roughening (:978-1050)
        {
            const double coastRoughenDist = js::max(8, js::round(8 * scaleFactor));
            const SimplexNoise cNoise(seed + 77), cNoise2(seed + 133), cNoise3(seed + 211);
            const double islandReach = js::max(4, js::round(4 * scaleFactor));
            for (int r = 0; r < N; r++) {
                if (dBdry[r] > coastRoughenDist) continue;
                const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
                const double t = dBdry[r] / coastRoughenDist;
                const double sn = js::min(1, js::max(coastStressMax[r], r_stress[r] / maxStress));
                const bool isSubductingOcean = r_isOcean[r] && coastConvergent[r] && coastSubductMax[r] > 0.45;
                const double subSup = isSubductingOcean ? js::min(1, (coastSubductMax[r] - 0.45) / 0.55) : 0;
                const float elevBeforeCoast = E[r];
                const bool isPassiveCoast = !coastConvergent[r];
                const double falloff1 = (1 - t) * (1 - t);
                const double stressAmp1 = 1 + sn * 5;
                const double coastFreq = isPassiveCoast ? 12 : 18;
                const double coastAmp = isPassiveCoast ? 0.08 : 0.12;
                const double n1 = cNoise.fbm(x * coastFreq + 3.7, y * coastFreq + 7.1, z * coastFreq + 2.3, 5, 0.55);
                double coastNoise1 = n1 * coastAmp * falloff1 * stressAmp1;
                if (subSup > 0 && coastNoise1 > 0) coastNoise1 *= (1 - subSup);
                add(r, coastNoise1);
                const double warpReach = isPassiveCoast ? 1.2 : 1.5;
                const double falloffW = js::max(0, 1 - t * warpReach);
                if (falloffW > 0) {
                    const double warpAmt = 0.35 * falloffW * (1 + sn * 2);
                    const double dwx = cNoise3.fbm(x * 6 + 11.3, y * 6 + 4.7, z * 6 + 8.2, 3, 0.6) * warpAmt;
                    const double dwy = cNoise3.fbm(x * 6 + 2.9, y * 6 + 9.4, z * 6 + 1.6, 3, 0.6) * warpAmt;
                    const double dwz = cNoise3.fbm(x * 6 + 7.5, y * 6 + 0.3, z * 6 + 5.9, 3, 0.6) * warpAmt;
                    const double origN = noise.fbm(x, y, z) * noiseMag;
                    const double warpN = noise.fbm(x + dwx, y + dwy, z + dwz) * noiseMag;
                    double warpDelta = (warpN - origN) * falloffW;
                    if (subSup > 0 && warpDelta > 0) warpDelta *= (1 - subSup);
                    add(r, warpDelta);
                }
                if (r_isOcean[r] && dBdry[r] > 0 && dBdry[r] <= islandReach && subSup < 0.3) {
                    const double islandN = cNoise2.fbm(x * 35 + 5.1, y * 35 + 9.3, z * 35 + 2.7, 4, 0.5);
                    const double threshold = 0.25 - sn * 0.2;
                    if (islandN > threshold) {
                        const double excess = (islandN - threshold) / (1 - threshold);
                        const double distFade = 1 - (dBdry[r] / islandReach);
                        double bump = excess * excess * 0.18 * (1 + sn * 2) * distFade;
                        bump *= (1 - subSup / 0.3);
                        add(r, bump);
                    }
                }
                dl_coastal[r] = js::f32((double)dl_coastal[r] + ((double)E[r] - (double)elevBeforeCoast));
            }
        }

        // island arcs (:1054-1107)
        {
            const SimplexNoise arcNoise(seed + 307);
            const double maxArcDist = js::max(5, js::round(5 * scaleFactor));
            std::vector<int> q;
            F32 arcDist(N, js::f32(maxArcDist + 1)), arcStress(N, 0.f);
            for (int r = 0; r < N; r++)
                if (r_boundaryType[r] == 1 && r_bothOcean[r] && r_subductFactor[r] < 0.45) {
                    q.push_back(r); arcDist[r] = 0; arcStress[r] = js::f32(js::min(1, r_stress[r] / maxStress));
                }
            for (size_t aq = 0; aq < q.size();) {
                const int r = q[aq++];
                const double nd = (double)arcDist[r] + 1;
                if (nd > maxArcDist) continue;
                const int plate = r_plate[r];
                for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                    const int nr = mesh.adjList[ni];
                    if (nd < arcDist[nr] && r_plate[nr] == plate && r_isOcean[nr]) { arcDist[nr] = js::f32(nd); arcStress[nr] = arcStress[r]; q.push_back(nr); }
                }
            }
            for (int r = 0; r < N; r++) {
                const double d = arcDist[r];
                if (d < 1 || d > maxArcDist) continue;
                const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
                const double peakDist = js::max(1.5, 1.5 * scaleFactor), sigma = js::max(1.5, 1.5 * scaleFactor);
                const double qd = (d - peakDist) / sigma;
                const double distWeight = pb_exp(-0.5 * (qd * qd));
                const double n = arcNoise.ridgedFbm(x * 4, y * 4, z * 4, 4, 2.0, 0.5, 1.0);
                if (n > 0.30) {
                    const double excess = (n - 0.30) / (1 - 0.30);
                    const double uplift = excess * excess * 0.55 * distWeight * (0.5 + arcStress[r]);
                    add(r, uplift);
                    dl_coastal[r] = js::f32((double)dl_coastal[r] + uplift);
                }
            }
            f["arcDist"] = arcDist;
        }

        // hotspots (:1115-1373)
        {
            const int NUM_HOTSPOTS = 5, CHAIN_LENGTH = 6;
            const double CHAIN_DECAY = 0.75, CHAIN_SPACING = 0.06, DOME_SIGMA = 0.006, DOME_STRENGTH = 0.60, SWELL_SIGMA_MULT = 2, SWELL_STR_MULT = 0.10;
            Rng hsRng(seed + 999);
            const SimplexNoise hsNoise(seed + 501), hsNoise2(seed + 502), hsNoise3(seed + 503);
            std::vector<Dome> domes;
            auto frame = [](double px, double py, double pz, double dx, double dy, double dz, Dome& o) {
                const double dd = dx * px + dy * py + dz * pz;
                double ux = dx - dd * px, uy = dy - dd * py, uz = dz - dd * pz;
                const double uLen = js::or_default(std::sqrt(ux * ux + uy * uy + uz * uz), 1);
                ux /= uLen; uy /= uLen; uz /= uLen;
                o.ux = ux; o.uy = uy; o.uz = uz;
                o.vx = py * uz - pz * uy; o.vy = pz * ux - px * uz; o.vz = px * uy - py * ux;
            };
            RandInt hsRandInt(seed + 1001);
            for (int h = 0; h < NUM_HOTSPOTS; h++) {
                const double hStrength = DOME_STRENGTH * (0.4 + hsRng.next() * 1.2);
                const double hSigma = DOME_SIGMA * (0.4 + hsRng.next() * 1.2);
                const double hDecay = CHAIN_DECAY + (hsRng.next() - 0.5) * 0.35;
                const int hLength = (int)js::max(3, CHAIN_LENGTH + js::round((hsRng.next() - 0.5) * 10));
                const int centerR = (int)hsRandInt((double)N);
                const double hx = xyz[3 * centerR], hy = xyz[3 * centerR + 1], hz = xyz[3 * centerR + 2];
                const int plate = r_plate[centerR];
                const int pk = P.find(plate);
                if (pk < 0) continue;
                double drift[3];
                plateVelocityAt(P, pk, hx, hy, hz, drift);
                const double driftLen = std::sqrt(drift[0] * drift[0] + drift[1] * drift[1] + drift[2] * drift[2]);
                if (driftLen < 1e-6) continue;
                drift[0] /= driftLen; drift[1] /= driftLen; drift[2] /= driftLen;
                const double oceanBoost = P.ocean(plate) ? 1.8 : 1.0;
                const double baseRiftAngle = hsNoise3.noise3D(hx * 10, hy * 10, hz * 10) * PB_PI;
                auto riftAngles = [&](int ci, int cl) {
                    std::vector<double> v;
                    if (ci == 0) v = {baseRiftAngle, baseRiftAngle + PB_PI * 0.6, baseRiftAngle - PB_PI * 0.6};
                    else if (ci == 1) v = {baseRiftAngle, baseRiftAngle + PB_PI};
                    else if (ci <= (int)std::floor(cl * 0.4)) v = {baseRiftAngle};
                    return v;
                };
                Dome d0{};
                d0.x = hx; d0.y = hy; d0.z = hz; d0.strength = hStrength * oceanBoost; d0.baseStrength = hStrength; d0.sigma = hSigma;
                d0.chainIndex = 0; d0.chainLength = hLength; d0.dx = drift[0]; d0.dy = drift[1]; d0.dz = drift[2];
                frame(hx, hy, hz, drift[0], drift[1], drift[2], d0);
                d0.riftAngles = riftAngles(0, hLength);
                domes.push_back(d0);
                double perpX = drift[1] * hz - drift[2] * hy, perpY = drift[2] * hx - drift[0] * hz, perpZ = drift[0] * hy - drift[1] * hx;
                const double perpLen = js::or_default(std::sqrt(perpX * perpX + perpY * perpY + perpZ * perpZ), 1);
                perpX /= perpLen; perpY /= perpLen; perpZ /= perpLen;
                double cx = hx, cy = hy, cz = hz, str = hStrength * oceanBoost, baseStr = hStrength;
                for (int c = 0; c < hLength; c++) {
                    const int ci = c + 1;
                    const double decayJitter = hDecay * (0.7 + hsRng.next() * 0.6);
                    str *= decayJitter; baseStr *= decayJitter;
                    const double stepSpacing = CHAIN_SPACING * (0.3 + hsRng.next() * 1.4);
                    const double ageBroadening = 1.0 + ci * 0.06;
                    const double stepSigma = hSigma * (0.5 + hsRng.next() * 1.0) * ageBroadening;
                    const double wobble = (hsRng.next() - 0.5) * 0.8;
                    const double ddx = -drift[0] + perpX * wobble, ddy = -drift[1] + perpY * wobble, ddz = -drift[2] + perpZ * wobble;
                    const double dot = ddx * cx + ddy * cy + ddz * cz;
                    double tx = ddx - dot * cx, ty = ddy - dot * cy, tz = ddz - dot * cz;
                    const double tLen = std::sqrt(tx * tx + ty * ty + tz * tz);
                    if (tLen < 1e-6) break;
                    tx /= tLen; ty /= tLen; tz /= tLen;
                    const double cosA = pb_cos(stepSpacing), sinA = pb_sin(stepSpacing);
                    cx = cx * cosA + tx * sinA; cy = cy * cosA + ty * sinA; cz = cz * cosA + tz * sinA;
                    const double nL = std::sqrt(cx * cx + cy * cy + cz * cz);
                    cx /= nL; cy /= nL; cz /= nL;
                    Dome dc{};
                    dc.x = cx; dc.y = cy; dc.z = cz; dc.strength = str; dc.baseStrength = baseStr; dc.sigma = stepSigma;
                    dc.chainIndex = ci; dc.chainLength = hLength; dc.dx = drift[0]; dc.dy = drift[1]; dc.dz = drift[2];
                    frame(cx, cy, cz, drift[0], drift[1], drift[2], dc);
                    dc.riftAngles = riftAngles(ci, hLength);
                    domes.push_back(dc);
                }
            }
            for (Dome& dm : domes) {
                dm.cosThreshPeak = pb_cos(dm.sigma * 5.5);
                dm.invS2 = -0.5 / (dm.sigma * dm.sigma);
                const double swSigma = dm.sigma * SWELL_SIGMA_MULT;
                dm.swellSigma = swSigma;
                dm.swellStrength = dm.baseStrength * SWELL_STR_MULT;
                dm.cosThreshSwell = pb_cos(swSigma * 3);
                dm.invS2Swell = -0.5 / (swSigma * swSigma);
                dm.driftStretch = 1.0 / 1.4;
                dm.hasCaldera = dm.chainIndex <= 1 && dm.strength > 0.15;
                dm.calderaSigma = dm.sigma * 0.25;
                dm.calderaDepth = dm.strength * 0.20;
                dm.invS2Caldera = -0.5 / (dm.calderaSigma * dm.calderaSigma);
                dm.ageFactor = dm.chainLength > 0 ? (double)dm.chainIndex / dm.chainLength : 0;
            }
            for (int r = 0; r < N; r++) {
                const double rx = xyz[3 * r], ry = xyz[3 * r + 1], rz = xyz[3 * r + 2];
                bool nearSwell = false, nearPeak = false;
                for (const Dome& dm : domes) {
                    const double cdot = dm.x * rx + dm.y * ry + dm.z * rz;
                    if (cdot > dm.cosThreshSwell) { nearSwell = true; if (cdot > dm.cosThreshPeak) { nearPeak = true; break; } }
                }
                if (!nearSwell) continue;
                double shapeWarpSq = 1.0;
                if (nearPeak) {
                    const double ws = 8;
                    const double wx = hsNoise2.fbm(rx * ws + 5.1, ry * ws + 3.7, rz * ws + 9.2, 2, 0.5) * 0.4;
                    const double wy = hsNoise2.fbm(rx * ws + 11.3, ry * ws + 7.1, rz * ws + 2.9, 2, 0.5) * 0.4;
                    const double wz = hsNoise2.fbm(rx * ws + 1.7, ry * ws + 13.5, rz * ws + 6.4, 2, 0.5) * 0.4;
                    const double shapeWarp = 1.0 + 0.40 * hsNoise.fbm((rx + wx) * 20 + 3.2, (ry + wy) * 20 + 7.8, (rz + wz) * 20 + 1.5, 4, 0.5);
                    shapeWarpSq = shapeWarp * shapeWarp;
                }
                double totalUplift = 0, totalSwellUplift = 0, weightedAge = 0, ageWeightSum = 0;
                for (const Dome& dm : domes) {
                    const double dot = dm.x * rx + dm.y * ry + dm.z * rz;
                    if (dot > dm.cosThreshSwell) { const double swAngleSq = 2 * (1 - dot); totalSwellUplift += dm.swellStrength * pb_exp(swAngleSq * dm.invS2Swell); }
                    if (dot < dm.cosThreshPeak) continue;
                    const double offX = rx - dot * dm.x, offY = ry - dot * dm.y, offZ = rz - dot * dm.z;
                    const double parComp = offX * dm.ux + offY * dm.uy + offZ * dm.uz;
                    const double perpComp = offX * dm.vx + offY * dm.vy + offZ * dm.vz;
                    const double stretchedParSq = (parComp * dm.driftStretch) * (parComp * dm.driftStretch);
                    const double angleSq = stretchedParSq + perpComp * perpComp;
                    double gauss = pb_exp(angleSq * shapeWarpSq * dm.invS2);
                    if (!dm.riftAngles.empty() && gauss > 0.01) {
                        const double angle = pb_atan2(perpComp, parComp);
                        double maxRift = 0;
                        for (double ra : dm.riftAngles) {
                            double da = angle - ra;
                            da = da - js::round(da / (2 * PB_PI)) * 2 * PB_PI;
                            const double c2 = pb_cos(da);
                            const double riftFactor = c2 * c2 * c2 * c2;
                            if (riftFactor > maxRift) maxRift = riftFactor;
                        }
                        gauss *= (1.0 + 0.5 * maxRift);
                    }
                    const double peakUplift = dm.strength * gauss;
                    totalUplift += peakUplift;
                    weightedAge += dm.ageFactor * peakUplift;
                    ageWeightSum += peakUplift;
                    if (dm.hasCaldera) totalUplift -= dm.calderaDepth * pb_exp(angleSq * dm.invS2Caldera);
                }
                const double combinedUplift = totalSwellUplift + totalUplift;
                if (combinedUplift > 0.001) {
                    const double age = ageWeightSum > 0 ? weightedAge / ageWeightSum : 0;
                    const double texBase = 0.7 * hsNoise.ridgedFbm(rx * 12, ry * 12, rz * 12, 4, 2.0, 0.5, 1.0);
                    const double texDetail = 0.3 * hsNoise.ridgedFbm(rx * 30, ry * 30, rz * 30, 3, 2.0, 0.5, 1.0);
                    const double texRaw = texBase + texDetail;
                    const double texMin = 0.4 + age * 0.3, texMax = 1.2 - age * 0.2;
                    const double volc = texMin + (texMax - texMin) * texRaw;
                    const double uplift = totalSwellUplift + js::max(0, totalUplift) * volc;
                    add(r, uplift);
                    dl_hotspot[r] = js::f32(uplift);
                }
            }
            F32 dd;
            for (const Dome& dm : domes) { dd.push_back((float)dm.x); dd.push_back((float)dm.y); dd.push_back((float)dm.z); dd.push_back((float)dm.strength); dd.push_back((float)dm.sigma); }
            f["domes"] = dd;
        }

        for (int r = 0; r < N; r++) if (E[r] > 0) E[r] = js::f32(pb_pow(E[r], 0.92));

        f["r_elevation"] = E; f["r_stress"] = r_stress; f["r_subductFactor"] = r_subductFactor;
        f["dist_mountain"] = dist_mountain; f["dist_ocean"] = dist_ocean; f["dist_coastline"] = dist_coastline;
        f["dist_coast"] = dist_coast; f["dist_coast_land"] = dist_coast_land;
        f["dBdry"] = dBdry; f["coastStressMax"] = coastStressMax; f["coastSubductMax"] = coastSubductMax;
        f["riftDist"] = riftDist; f["ridgeDist"] = ridgeDist; f["fractureDist"] = fractureDist; f["backArcDist"] = backArcDist; f["backArcStress"] = backArcStress;
        f["base"] = dl_base; f["tectonic"] = dl_tectonic; f["noise"] = dl_noise; f["interior"] = dl_interior; f["coastal"] = dl_coastal;
        f["ocean"] = dl_ocean; f["hotspot"] = dl_hotspot; f["tecActivity"] = dl_tecActivity; f["margins"] = dl_margins; f["backArc"] = dl_backArc;
        f["foldRidge"] = dl_foldRidge; f["orogenicPower"] = dl_orogenicPower;
        u["mountain_r"] = mountain_r.in; u["coastline_r"] = coastline_r.in; u["ocean_r"] = ocean_r.in;
        u["coastConvergent"] = coastConvergent; u["r_bothOcean"] = r_bothOcean; u["r_hasOcean"] = r_hasOcean;
        I32 bt(N); for (int r = 0; r < N; r++) bt[r] = r_boundaryType[r];
        i["r_boundaryType"] = bt;
        i["mountain_order"] = I32(mountain_r.items.begin(), mountain_r.items.end());
        i["coastline_order"] = I32(coastline_r.items.begin(), coastline_r.items.end());
        i["ocean_order"] = I32(ocean_r.items.begin(), ocean_r.items.end());
    }
};

extern "C" {

void* orc_elev_create(int N, const int32_t* off, const int32_t* adj, const float* xyz) { return new OracleElevation(OMesh{N, off, adj}, xyz); }
void orc_elev_destroy(void* h) { delete (OracleElevation*)h; }
// plate tables: ids, isOcean (u8), pole (3 doubles each), omega, density — one row per plate.
// plateSeeds: plate ids in the Set's insertion order.  nSuper == 0 → no superPlateData.
void orc_elev_assign(void* h, const int32_t* r_plate, int nPlates, const int32_t* ids, const uint8_t* isOcean, const double* pole,
                     const double* omega, const double* density, const int32_t* plateSeeds, int nSeeds, double noiseSeed,
                     double noiseMag, double seed, double spread, const int32_t* r_superPlate, int nSuper, const int32_t* sIds,
                     const uint8_t* sIsOcean, const double* sPole, const double* sOmega, const double* sDensity) {
    Plates P(nPlates, ids, isOcean, pole, omega, density);
    if (nSuper > 0) {
        Plates SP(nSuper, sIds, sIsOcean, sPole, sOmega, sDensity);
        ((OracleElevation*)h)->assignElevation(P, r_plate, plateSeeds, nSeeds, noiseSeed, noiseMag, seed, spread, &SP, r_superPlate);
    } else ((OracleElevation*)h)->assignElevation(P, r_plate, plateSeeds, nSeeds, noiseSeed, noiseMag, seed, spread, nullptr, nullptr);
}
int64_t orc_elev_get(void* h, const char* name, int kind, void* out, int64_t cap) {
    OracleElevation* c = (OracleElevation*)h;
    const void* p = nullptr;
    int64_t n = 0, es = 4;
    if (kind == 0) { auto it = c->f.find(name); if (it == c->f.end()) return -1; p = it->second.data(); n = (int64_t)it->second.size(); }
    else if (kind == 1) { auto it = c->i.find(name); if (it == c->i.end()) return -1; p = it->second.data(); n = (int64_t)it->second.size(); }
    else { auto it = c->u.find(name); if (it == c->u.end()) return -1; p = it->second.data(); n = (int64_t)it->second.size(); es = 1; }
    if (out && cap >= n) std::memcpy(out, p, (size_t)(n * es));
    return n;
}
double orc_pair_intensity(int a, int b) { return getPairIntensity(a, b); }

}  // extern "C"
