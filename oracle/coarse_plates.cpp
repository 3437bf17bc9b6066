// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).  Pinned against the reference's own source run under tests/golden/minijs.py (tests/test_zz_reference_vectors.py).
// Sequential restatement of the coarse stage that feeds the plate pipeline:
//   generatePlates     js/plates.js:6-232      (farthest-point seeds, round-robin weighted growth, Euler poles)
//   assignOceanLand    js/ocean-land.js:7-238  (continent seeding / growth on the plate graph, trapped seas)
// generateCoarsePlates (js/coarse-plates.js:19-39) = buildSphere(20000, 0.75, makeRng(seed + 137)) + these two; the
// mesh part is oracle/mesh_hull.py, the glue is oracle/binding.py:generate_coarse_plates.
// JS containers keep their iteration order: Set / Array → std::vector in insertion order; objects keyed by plate id
// are only looked up, never iterated, so std::map / dense arrays stand in for them.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

#include "js_semantics.h"
#include "noise.h"

extern "C" void orc_smooth_and_reconnect_plates(int N, const int32_t* off, const int32_t* adj, int32_t* r_plate, const int32_t* seeds,
                                                int nSeeds, int numPasses);

namespace {

struct Top3 {
    int r0 = -1, r1 = -1, r2 = -1; double d0 = -1, d1 = -1, d2 = -1;
    void offer(int r, double d) {
        if (d > d2) {
            if (d > d0) { r2 = r1; d2 = d1; r1 = r0; d1 = d0; r0 = r; d0 = d; }
            else if (d > d1) { r2 = r1; d2 = d1; r1 = r; d1 = d; }
            else { r2 = r; d2 = d; }
        }
    }
    int valid() const { return (r0 != -1) + (r1 != -1) + (r2 != -1); }
    int pick(int k) const { return k == 0 ? r0 : k == 1 ? r1 : r2; }
};

// js/plates.js:6-232
void generate_plates(int N, const int32_t* off, const int32_t* adj, const float* xyz, int numPlates, double seed,
                     std::vector<int32_t>& r_plate, std::vector<int>& plateSeeds, std::vector<double>& pole, std::vector<double>& omega) {
    r_plate.assign(N, -1);
    Rng rng(seed + 0.5);
    RandInt randInt(seed);
    std::vector<uint8_t> isSeed(N, 0);
    std::vector<float> minDist(N, INFINITY);
    auto dist_to = [&](int r, double sx, double sy, double sz) { return 1 - (xyz[3 * r] * sx + xyz[3 * r + 1] * sy + xyz[3 * r + 2] * sz); };
    auto add_seed = [&](int r) { plateSeeds.push_back(r); isSeed[r] = 1; };

    const int firstSeed = (int)randInt(N);
    add_seed(firstSeed);
    {
        const double sx = xyz[3 * firstSeed], sy = xyz[3 * firstSeed + 1], sz = xyz[3 * firstSeed + 2];
        for (int r = 0; r < N; r++) minDist[r] = js::f32(dist_to(r, sx, sy, sz));
        minDist[firstSeed] = 0;
    }
    while ((int)plateSeeds.size() < numPlates && (int)plateSeeds.size() < N) {
        Top3 t;
        for (int r = 0; r < N; r++) if (!isSeed[r]) t.offer(r, minDist[r]);
        int valid = t.valid();
        if (!valid) break;
        const int newSeed = t.pick((int)randInt(valid));
        add_seed(newSeed);
        const double nx = xyz[3 * newSeed], ny = xyz[3 * newSeed + 1], nz = xyz[3 * newSeed + 2];
        if ((int)plateSeeds.size() < numPlates) {
            Top3 u;
            for (int r = 0; r < N; r++) {
                const double d = dist_to(r, nx, ny, nz);
                if (d < minDist[r]) minDist[r] = js::f32(d);
                if (isSeed[r]) continue;
                u.offer(r, minDist[r]);
            }
            valid = u.valid();
            if (!valid) break;
            const int newSeed2 = u.pick((int)randInt(valid));
            add_seed(newSeed2);
            const double mx = xyz[3 * newSeed2], my = xyz[3 * newSeed2 + 1], mz = xyz[3 * newSeed2 + 2];
            for (int r = 0; r < N; r++) { const double d = dist_to(r, mx, my, mz); if (d < minDist[r]) minDist[r] = js::f32(d); }
        } else {
            for (int r = 0; r < N; r++) { const double d = dist_to(r, nx, ny, nz); if (d < minDist[r]) minDist[r] = js::f32(d); }
        }
    }

    const double lowPlateT = js::max(0, js::min(1, (80 - numPlates) / 60.0));
    const double rateMin = 0.7 - 0.4 * lowPlateT, rateRange = 2.3 + 2.4 * lowPlateT;
    const double dirBase = 0.15 + 0.25 * lowPlateT, dirScale = 0.25 + 0.25 * lowPlateT;
    const int P = (int)plateSeeds.size();
    std::vector<double> growthRate(P), dirStrength(P), gdir(3 * (size_t)P);
    for (int k = 0; k < P; k++) {
        const int center = plateSeeds[k];
        const double a = rng.next(), b = rng.next();
        growthRate[k] = rateMin + a * b * rateRange;
        const double px = xyz[3 * center], py = xyz[3 * center + 1], pz = xyz[3 * center + 2];
        double pLen = std::sqrt(px * px + py * py + pz * pz); if (pLen == 0 || pLen != pLen) pLen = 1;
        const double nx = px / pLen, ny = py / pLen, nz = pz / pLen;
        const double rx = rng.next() - 0.5, ry = rng.next() - 0.5, rz = rng.next() - 0.5;
        const double d = rx * nx + ry * ny + rz * nz;
        const double tx = rx - d * nx, ty = ry - d * ny, tz = rz - d * nz;
        double tLen = std::sqrt(tx * tx + ty * ty + tz * tz); if (tLen == 0 || tLen != tLen) tLen = 1;
        gdir[3 * k] = tx / tLen; gdir[3 * k + 1] = ty / tLen; gdir[3 * k + 2] = tz / tLen;
        dirStrength[k] = js::min(0.85, rng.next() * (dirBase + dirScale / growthRate[k]));
    }
    std::vector<std::vector<int>> frontier(P);
    std::vector<double> areaCount(P, 1);
    for (int k = 0; k < P; k++) { r_plate[plateSeeds[k]] = plateSeeds[k]; frontier[k].push_back(plateSeeds[k]); }
    long long remaining = (long long)N - P;
    const double COMPACT_WEIGHT = 0.3 - 0.22 * lowPlateT;
    const double expectedArea = js::max(1, (double)(N - P) / numPlates);
    const double areaGovernorMult = 2.0 + 2.0 * lowPlateT;
    const double invNumRegions = 1.0 / N;
    while (remaining > 0) {
        bool anyProgress = false;
        for (int k = 0; k < P; k++) {
            std::vector<int>& fr = frontier[k];
            if (fr.empty()) continue;
            const int pid = plateSeeds[k];
            const double rate = growthRate[k], d0 = gdir[3 * k], d1 = gdir[3 * k + 1], d2 = gdir[3 * k + 2];
            const double dirStr = dirStrength[k], dirStrHalf = dirStr * 0.5;
            double steps = js::max(1, std::ceil(rate * (0.5 + rng.next())));
            if (areaCount[k] > expectedArea * areaGovernorMult) steps = js::max(1, std::ceil(steps * 0.5));
            const double expectedChordDist = std::sqrt((areaCount[k] != 0 ? areaCount[k] : 1) * invNumRegions / PB_PI) * 2;
            const double compactThreshold = expectedChordDist * 1.8;
            const double sx = xyz[3 * pid], sy = xyz[3 * pid + 1], sz = xyz[3 * pid + 2];
            for (int s = 0; s < steps && !fr.empty(); s++) {
                int bestIdx = 0; double bestScore = -INFINITY;
                const int samples = (int)js::min((double)fr.size(), 3 + std::floor(dirStr * 5));
                for (int i = 0; i < samples; i++) {
                    const int idx = (int)randInt((double)fr.size());
                    const int cell = fr[idx];
                    const double dx = xyz[3 * cell] - sx, dy = xyz[3 * cell + 1] - sy, dz = xyz[3 * cell + 2] - sz;
                    const double dLenSq = dx * dx + dy * dy + dz * dz;
                    double dLen = std::sqrt(dLenSq); if (dLen == 0 || dLen != dLen) dLen = 1;
                    const double alignment = (dx * d0 + dy * d1 + dz * d2) / dLen;
                    const double excess = js::max(0, dLenSq * 0.5 - compactThreshold);
                    const double compactPenalty = excess * (COMPACT_WEIGHT * 4);
                    const double score = alignment * dirStr + rng.next() * (1 - dirStrHalf) - compactPenalty;
                    if (score > bestScore) { bestScore = score; bestIdx = idx; }
                }
                const int current = fr[bestIdx];
                fr[bestIdx] = fr.back();
                fr.pop_back();
                for (int j = off[current], e = off[current + 1]; j < e; j++) {
                    const int nb = adj[j];
                    if (r_plate[nb] == -1) { r_plate[nb] = pid; fr.push_back(nb); areaCount[k]++; remaining--; anyProgress = true; }
                }
            }
        }
        if (!anyProgress) break;
    }
    for (bool orphans = true; orphans;) {
        orphans = false;
        for (int r = 0; r < N; r++) {
            if (r_plate[r] != -1) continue;
            for (int j = off[r], e = off[r + 1]; j < e; j++)
                if (r_plate[adj[j]] != -1) { r_plate[r] = r_plate[adj[j]]; orphans = true; break; }
        }
    }
    std::vector<int32_t> seeds32(plateSeeds.begin(), plateSeeds.end());
    orc_smooth_and_reconnect_plates(N, off, adj, r_plate.data(), seeds32.data(), P, (int)js::round(3 - 2 * lowPlateT));
    pole.resize(3 * (size_t)P); omega.resize(P);
    for (int k = 0; k < P; k++) {
        const double theta = rng.next() * 2 * PB_PI;
        const double cosP = 2 * rng.next() - 1;
        const double sinP = std::sqrt(1 - cosP * cosP);
        pole[3 * k] = sinP * pb_cos(theta); pole[3 * k + 1] = sinP * pb_sin(theta); pole[3 * k + 2] = cosP;
        const double mag = 0.5 + rng.next() * 1.5;
        omega[k] = mag * (rng.next() < 0.5 ? -1 : 1);
    }
}

// js/ocean-land.js:7-238.  Returns isOcean per plate in plateSeeds order.
std::vector<uint8_t> assign_ocean_land(int N, const int32_t* off, const int32_t* adj, const int32_t* r_plate, const std::vector<int>& plateIds,
                                       const float* xyz, double seed, int numContinents, double variety, double landCoverage) {
    Rng rng(seed + 42);
    const int P = (int)plateIds.size();
    std::map<int, int> idx;
    for (int k = 0; k < P; k++) idx[plateIds[k]] = k;
    std::vector<double> area(P, 0), cx(P, 0), cy(P, 0), cz(P, 0), perim(P, 0), compact(P, 0);
    for (int r = 0; r < N; r++) {
        auto it = idx.find(r_plate[r]);
        if (it == idx.end()) continue;
        const int k = it->second;
        area[k]++; cx[k] += xyz[3 * r]; cy[k] += xyz[3 * r + 1]; cz[k] += xyz[3 * r + 2];
    }
    for (int k = 0; k < P; k++) { const double a = area[k] != 0 ? area[k] : 1; cx[k] /= a; cy[k] /= a; cz[k] /= a; }
    std::vector<std::vector<int>> padj(P);
    for (int r = 0; r < N; r++) {
        const int me = idx.at(r_plate[r]);
        bool boundary = false;
        for (int ni = off[r], e = off[r + 1]; ni < e; ni++) {
            const int np = r_plate[adj[ni]];
            if (np != r_plate[r]) {
                const int o = idx.at(np);
                if (std::find(padj[me].begin(), padj[me].end(), o) == padj[me].end()) padj[me].push_back(o);
                boundary = true;
            }
        }
        if (boundary) perim[me]++;
    }
    double maxCompact = 0;
    for (int k = 0; k < P; k++) {
        compact[k] = std::sqrt(area[k] != 0 ? area[k] : 1) / (perim[k] != 0 ? perim[k] : 1);
        if (compact[k] > maxCompact) maxCompact = compact[k];
    }
    if (maxCompact > 0) for (int k = 0; k < P; k++) compact[k] /= maxCompact;
    const double targetLandArea = landCoverage * N;
    const int effectiveNum = std::min(numContinents, P);
    std::vector<int> contSeeds;
    std::vector<uint8_t> chosen(P, 0);
    struct Cand { int k; double score; };
    const int first = (int)std::floor(rng.next() * P);
    contSeeds.push_back(first); chosen[first] = 1;
    for (int s = 1; s < effectiveNum; s++) {
        std::vector<Cand> cands;
        for (int k = 0; k < P; k++) {
            if (chosen[k]) continue;
            double minD = INFINITY;
            for (int ex : contSeeds) {
                const double dx = cx[k] - cx[ex], dy = cy[k] - cy[ex], dz = cz[k] - cz[ex];
                const double d = dx * dx + dy * dy + dz * dz;
                if (d < minD) minD = d;
            }
            const double rawAreaFactor = std::sqrt((double)N / P) / std::sqrt(area[k] != 0 ? area[k] : 1);
            const double areaFactor = 1 + (rawAreaFactor - 1) * (1 - variety * 0.5);
            const double comp = 0.3 + 0.7 * compact[k];
            cands.push_back({k, minD * areaFactor * comp});
        }
        if (cands.empty()) break;
        std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return b.score - a.score < 0; });
        const int topK = std::min((int)cands.size(), 3);
        const Cand pick = cands[(int)std::floor(rng.next() * topK)];
        contSeeds.push_back(pick.k); chosen[pick.k] = 1;
    }
    double seedArea = 0;
    for (int k : contSeeds) seedArea += area[k];
    while (contSeeds.size() > 1 && seedArea > targetLandArea) {
        size_t maxIdx = 0;
        for (size_t i = 1; i < contSeeds.size(); i++) if (area[contSeeds[i]] > area[contSeeds[maxIdx]]) maxIdx = i;
        seedArea -= area[contSeeds[maxIdx]];
        chosen[contSeeds[maxIdx]] = 0;
        contSeeds.erase(contSeeds.begin() + (long)maxIdx);
    }
    std::vector<int> continent(P, -1);            // -1 ↔ undefined
    for (size_t c = 0; c < contSeeds.size(); c++) continent[contSeeds[c]] = (int)c;
    double landArea = seedArea;
    const double growTarget = targetLandArea * 0.9;
    const int numC = (int)contSeeds.size();
    std::vector<double> cTarget(numC, 0), cArea(numC, 0);
    for (int c = 0; c < numC; c++) cArea[c] = area[contSeeds[c]];
    if (variety > 0 && numC > 1) {
        std::vector<double> w;
        for (int c = 0; c < numC; c++) w.push_back(pb_exp((rng.next() - 0.5) * variety * 2.5));
        double total = 0;
        for (double v : w) total = total + v;
        for (int c = 0; c < numC; c++) cTarget[c] = growTarget * w[c] / total;
    } else {
        const double equal = growTarget / std::max(numC, 1);
        for (int c = 0; c < numC; c++) cTarget[c] = equal;
    }
    bool progress = true;
    while (progress && landArea < growTarget) {
        progress = false;
        for (int c = 0; c < numC && landArea < growTarget; c++) {
            if (cArea[c] >= cTarget[c]) continue;
            std::vector<Cand> cands;
            for (int k = 0; k < P; k++) {
                if (continent[k] != -1) continue;
                bool self = false, other = false; int same = 0;
                for (int a : padj[k]) {
                    const int ac = continent[a];
                    if (ac == c) { self = true; same++; }
                    else if (ac != -1) { other = true; break; }
                }
                if (self && !other) cands.push_back({k, same + compact[k] * 3 + rng.next() * 0.5});
            }
            if (cands.empty()) continue;
            std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return b.score - a.score < 0; });
            const int topK = std::min((int)cands.size(), 3);
            const Cand pick = cands[(int)std::floor(rng.next() * topK)];
            continent[pick.k] = c; cArea[c] += area[pick.k]; landArea += area[pick.k]; progress = true;
        }
    }
    std::vector<std::vector<int>> comps;
    std::vector<uint8_t> vis(P, 0);
    for (int k = 0; k < P; k++) {
        if (continent[k] != -1 || vis[k]) continue;
        std::vector<int> comp{k};
        vis[k] = 1;
        for (size_t qi = 0; qi < comp.size(); qi++)
            for (int a : padj[comp[qi]]) if (continent[a] == -1 && !vis[a]) { vis[a] = 1; comp.push_back(a); }
        comps.push_back(comp);
    }
    auto comp_area = [&](const std::vector<int>& c) { double a = 0; for (int p : c) a += area[p]; return a; };
    size_t mainIdx = 0;
    for (size_t i = 1; i < comps.size(); i++) if (comp_area(comps[i]) > comp_area(comps[mainIdx])) mainIdx = i;
    const double absorbCap = targetLandArea * 1.1;
    for (size_t i = 0; i < comps.size(); i++) {
        if (i == mainIdx) continue;
        std::vector<int> bordering;
        for (int op : comps[i]) {
            for (int a : padj[op])
                if (continent[a] != -1 && std::find(bordering.begin(), bordering.end(), continent[a]) == bordering.end()) bordering.push_back(continent[a]);
            if (bordering.size() > 1) break;
        }
        if (bordering.size() == 1) {
            const double ca = comp_area(comps[i]);
            if (landArea + ca <= absorbCap) { for (int op : comps[i]) continent[op] = bordering[0]; landArea += ca; }
        }
    }
    std::vector<uint8_t> isOcean(P);
    for (int k = 0; k < P; k++) isOcean[k] = continent[k] == -1;
    return isOcean;
}

}  // namespace

extern "C" {
// Returns the number of plates actually seeded.  Output arrays hold numPlates entries (3·numPlates for pole).
int orc_generate_coarse_plates(int N, const int32_t* off, const int32_t* adj, const float* xyz, double seed, int numPlates,
                               int numContinents, double continentSizeVariety, double landCoverage, int32_t* r_plate, int32_t* seeds,
                               double* pole, double* omega, uint8_t* isOcean) {
    std::vector<int32_t> rp;
    std::vector<int> ps;
    std::vector<double> pl, om;
    generate_plates(N, off, adj, xyz, numPlates, seed, rp, ps, pl, om);
    std::vector<uint8_t> oc = assign_ocean_land(N, off, adj, rp.data(), ps, xyz, seed, numContinents, continentSizeVariety, landCoverage);
    std::copy(rp.begin(), rp.end(), r_plate);
    for (size_t k = 0; k < ps.size(); k++) {
        seeds[k] = ps[k]; omega[k] = om[k]; isOcean[k] = oc[k];
        pole[3 * k] = pl[3 * k]; pole[3 * k + 1] = pl[3 * k + 1]; pole[3 * k + 2] = pl[3 * k + 2];
    }
    return (int)ps.size();
}
}
