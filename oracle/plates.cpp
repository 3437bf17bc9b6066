// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).  Pinned against the reference's own source run under tests/golden/minijs.py (tests/test_zz_reference_vectors.py).
// Sequential restatement of the plate pipeline on the hi-res mesh:
//   projectCoarsePlates        js/coarse-plates.js:51-117
//   smoothAndReconnectPlates   js/plates.js:241-348
//   buildSuperPlates           js/super-plates.js:16-273
// JS containers are restated with their iteration order: Set / Object-by-insertion → std::vector in insertion order.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "js_semantics.h"
#include "noise.h"

namespace {

// js/coarse-plates.js:51-117
void project_coarse_plates(int N, const float* r_xyz, int NC, const int32_t* cOff, const int32_t* cAdj, const float* coarse_xyz,
                           const int32_t* coarse_r_plate, double seed, int numPlates /* <0: null */, int32_t* r_plate) {
    SimplexNoise noise(seed + 999);
    const double coarseEdgeRad = PB_PI / std::sqrt((double)NC);
    const double lowPlateT = numPlates >= 0 ? js::max(0, js::min(1, (80 - numPlates) / 60.0)) : 0;
    const double perturbAmp = coarseEdgeRad * (1.5 + 1.0 * lowPlateT);
    const double BASE_FREQ = 8;
    const int MAX_WALK = (int)std::ceil(std::sqrt((double)NC));
    int cur = 0;
    for (int r = 0; r < N; r++) {
        const double ox = r_xyz[3 * r], oy = r_xyz[3 * r + 1], oz = r_xyz[3 * r + 2];
        double dx = 0, dy = 0, dz = 0, amp = perturbAmp, freq = BASE_FREQ;
        for (int oct = 0; oct < 4; oct++) {
            dx += noise.noise3D(ox * freq, oy * freq, oz * freq) * amp;
            dy += noise.noise3D(ox * freq + 100, oy * freq + 100, oz * freq + 100) * amp;
            dz += noise.noise3D(ox * freq + 200, oy * freq + 200, oz * freq + 200) * amp;
            amp *= 0.5;
            freq *= 2;
        }
        double px = ox + dx, py = oy + dy, pz = oz + dz;
        double len = std::sqrt(px * px + py * py + pz * pz);
        if (len == 0 || len != len) len = 1;               // `|| 1`
        px /= len; py /= len; pz /= len;
        double bestDot = px * coarse_xyz[3 * cur] + py * coarse_xyz[3 * cur + 1] + pz * coarse_xyz[3 * cur + 2];
        bool improved = true;
        int steps = 0;
        while (improved && steps < MAX_WALK) {
            improved = false;
            steps++;
            for (int i = cOff[cur], iEnd = cOff[cur + 1]; i < iEnd; i++) {   // bounds fixed at loop entry, cur may move
                const int nb = cAdj[i];
                const double d = px * coarse_xyz[3 * nb] + py * coarse_xyz[3 * nb + 1] + pz * coarse_xyz[3 * nb + 2];
                if (d > bestDot) { bestDot = d; cur = nb; improved = true; }
            }
        }
        if (steps >= MAX_WALK) {
            for (int c = 0; c < NC; c++) {
                const double d = px * coarse_xyz[3 * c] + py * coarse_xyz[3 * c + 1] + pz * coarse_xyz[3 * c + 2];
                if (d > bestDot) { bestDot = d; cur = c; }
            }
        }
        r_plate[r] = coarse_r_plate[cur];
    }
}

// js/plates.js:241-348
void smooth_and_reconnect(int N, const int32_t* off, const int32_t* adj, int32_t* r_plate, const int32_t* seeds, int nSeeds, int numPasses) {
    std::vector<uint8_t> isSeed(N, 0);
    for (int k = 0; k < nSeeds; k++) {
        const int pid = seeds[k];
        if (pid < N && r_plate[pid] == pid) isSeed[pid] = 1;
    }
    int maxDeg = 0;
    for (int r = 0; r < N; r++) maxDeg = std::max(maxDeg, off[r + 1] - off[r]);
    std::vector<int32_t> cntPlates(maxDeg + 1);
    std::vector<int> cntValues(maxDeg + 1);
    for (int pass = 0; pass < numPasses; pass++) {
        const double threshold = pass == 0 ? 0.4 : 0.5;
        for (int r = 0; r < N; r++) {
            const int rStart = off[r], rEnd = off[r + 1], deg = rEnd - rStart;
            int nDistinct = 0;
            for (int j = rStart; j < rEnd; j++) {
                const int p = r_plate[adj[j]];
                bool found = false;
                for (int k = 0; k < nDistinct; k++) if (cntPlates[k] == p) { cntValues[k]++; found = true; break; }
                if (!found) { cntPlates[nDistinct] = p; cntValues[nDistinct] = 1; nDistinct++; }
            }
            int bestPlate = r_plate[r], bestCount = 0;
            for (int k = 0; k < nDistinct; k++) if (cntValues[k] > bestCount) { bestCount = cntValues[k]; bestPlate = cntPlates[k]; }
            if (bestCount > deg * threshold && !isSeed[r]) r_plate[r] = bestPlate;
        }
    }
    // reconnect: keep the largest component of every plate (first found wins ties), re-assign the rest by BFS
    std::vector<uint8_t> visited(N, 0), inMain(N, 0);
    std::map<int, std::vector<int>> bestComponent;
    std::vector<int> bfs;
    for (int r = 0; r < N; r++) {
        if (visited[r]) continue;
        const int pid = r_plate[r];
        bfs.clear();
        bfs.push_back(r);
        visited[r] = 1;
        for (size_t qi = 0; qi < bfs.size(); qi++)
            for (int ni = off[bfs[qi]], e = off[bfs[qi] + 1]; ni < e; ni++) {
                const int nb = adj[ni];
                if (!visited[nb] && r_plate[nb] == pid) { visited[nb] = 1; bfs.push_back(nb); }
            }
        auto it = bestComponent.find(pid);
        if (it == bestComponent.end() || bfs.size() > it->second.size()) bestComponent[pid] = bfs;
    }
    for (auto& kv : bestComponent) for (int r : kv.second) inMain[r] = 1;
    std::vector<int> queue;
    for (int r = 0; r < N; r++) {
        if (inMain[r]) continue;
        for (int ni = off[r], e = off[r + 1]; ni < e; ni++)
            if (inMain[adj[ni]]) { r_plate[r] = r_plate[adj[ni]]; inMain[r] = 1; queue.push_back(r); break; }
    }
    for (size_t qi = 0; qi < queue.size(); qi++) {
        const int r = queue[qi];
        for (int ni = off[r], e = off[r + 1]; ni < e; ni++) {
            const int nb = adj[ni];
            if (!inMain[nb]) { r_plate[nb] = r_plate[r]; inMain[nb] = 1; queue.push_back(nb); }
        }
    }
}

// js/super-plates.js:16-273.  Plate tables are parallel arrays in plateSeeds order; hasVec[k] == 0 ↔ `!pv || !pv.pole`,
// density NaN ↔ undefined.  Outputs: r_superPlate[N], and per super plate pole[3], omega, isOcean, density.
int build_super_plates(int N, const int32_t* off, const int32_t* adj, const int32_t* r_plate, int nSeeds, const int32_t* seeds,
                       const uint8_t* hasVec, const double* pole, const double* omega, const uint8_t* isOcean, const double* density,
                       int32_t* r_superPlate, double* spPole, double* spOmega, uint8_t* spIsOcean, double* spDensity, int spCap) {
    std::map<int, int> idx;                                   // plate seed id → position in plateSeeds order
    for (int k = 0; k < nSeeds; k++) idx[seeds[k]] = k;
    std::vector<double> plateArea(nSeeds, 0);
    for (int r = 0; r < N; r++) plateArea[idx.at(r_plate[r])]++;
    // plateNeighbors: Set per plate, insertion order = first occurrence in the (r, ni) scan
    std::vector<std::vector<int>> nbrs(nSeeds);
    for (int r = 0; r < N; r++) {
        const int me = idx.at(r_plate[r]);
        for (int ni = off[r], e = off[r + 1]; ni < e; ni++) {
            const int np = r_plate[adj[ni]];
            if (np != r_plate[r]) {
                const int o = idx.at(np);
                if (std::find(nbrs[me].begin(), nbrs[me].end(), o) == nbrs[me].end()) nbrs[me].push_back(o);
            }
        }
    }
    // components of same-type plates
    std::vector<uint8_t> vis(nSeeds, 0);
    std::vector<std::vector<int>> components;
    for (int p = 0; p < nSeeds; p++) {
        if (vis[p]) continue;
        std::vector<int> comp, queue{p};
        vis[p] = 1;
        for (size_t head = 0; head < queue.size(); head++) {
            const int cur = queue[head];
            comp.push_back(cur);
            for (int nb : nbrs[cur]) if (!vis[nb] && isOcean[nb] == isOcean[p]) { vis[nb] = 1; queue.push_back(nb); }
        }
        components.push_back(comp);
    }
    const int numPlates = nSeeds;
    const int target = (int)js::max(2, js::min(20, js::round(numPlates / 4.0)));
    std::vector<int> plateToSuper(nSeeds, -1);
    int nextSuper = 0;
    const double INF = INFINITY;
    for (const auto& comp : components) {
        const int k = (int)js::max(1, js::round((double)target * (double)comp.size() / (double)numPlates));
        if (k <= 1) {
            const int sp = nextSuper++;
            for (int p : comp) plateToSuper[p] = sp;
            continue;
        }
        std::vector<uint8_t> inComp(nSeeds, 0);
        for (int p : comp) inComp[p] = 1;
        std::vector<std::vector<int>> localAdj(nSeeds);
        std::vector<double> edgeWeight(nSeeds, 0), dist(nSeeds, INF);
        for (int p : comp) {
            for (int nb : nbrs[p]) if (inComp[nb]) localAdj[p].push_back(nb);
            edgeWeight[p] = std::sqrt(plateArea[p] != 0 ? plateArea[p] : 1);
        }
        auto dijkstraFrom = [&](const std::vector<int>& start) {
            for (int p : comp) dist[p] = INF;
            std::vector<uint8_t> visited(nSeeds, 0);
            for (int s : start) dist[s] = 0;
            for (size_t iter = 0; iter < comp.size(); iter++) {
                int cur = -1; double minD = INF;
                for (int p : comp) if (!visited[p] && dist[p] < minD) { minD = dist[p]; cur = p; }
                if (cur == -1) break;
                visited[cur] = 1;
                for (int nb : localAdj[cur]) { const double nd = dist[cur] + edgeWeight[nb]; if (nd < dist[nb]) dist[nb] = nd; }
            }
        };
        std::vector<int> sds{comp[0]};
        dijkstraFrom(sds);
        for (int si = 1; si < k; si++) {
            int farthest = comp[0]; double maxDist = -1;
            for (int p : comp) if (dist[p] > maxDist) { maxDist = dist[p]; farthest = p; }
            sds.push_back(farthest);
            dijkstraFrom(sds);
        }
        std::vector<int> assignment(nSeeds, -1);
        std::vector<double> d(nSeeds, INF);
        std::vector<uint8_t> visited(nSeeds, 0);
        for (size_t si = 0; si < sds.size(); si++) { assignment[sds[si]] = nextSuper + (int)si; d[sds[si]] = 0; }
        for (size_t iter = 0; iter < comp.size(); iter++) {
            int cur = -1; double minD = INF;
            for (int p : comp) if (!visited[p] && d[p] < minD) { minD = d[p]; cur = p; }
            if (cur == -1) break;
            visited[cur] = 1;
            for (int nb : localAdj[cur]) {
                const double nd = d[cur] + edgeWeight[nb];
                if (nd < d[nb]) { d[nb] = nd; assignment[nb] = assignment[cur]; }
            }
        }
        for (int p : comp) plateToSuper[p] = assignment[p];
        nextSuper += (int)sds.size();
    }
    const int nSP = nextSuper;
    if (nSP > spCap) return -nSP;
    for (int r = 0; r < N; r++) r_superPlate[r] = plateToSuper[idx.at(r_plate[r])];
    std::vector<double> Lx(nSP, 0), Ly(nSP, 0), Lz(nSP, 0), omSum(nSP, 0), arSum(nSP, 0), largestArea(nSP, 0);
    std::vector<int> largest(nSP, -1);
    for (int p = 0; p < nSeeds; p++) {
        const int sp = plateToSuper[p];
        if (!hasVec[p]) continue;
        const double area = plateArea[p], om = omega[p];
        Lx[sp] += area * om * pole[3 * p]; Ly[sp] += area * om * pole[3 * p + 1]; Lz[sp] += area * om * pole[3 * p + 2];
        omSum[sp] += area * std::fabs(om);
        arSum[sp] += area;
        if (largest[sp] < 0 || area > largestArea[sp]) { largest[sp] = p; largestArea[sp] = area; }
    }
    for (int sp = 0; sp < nSP; sp++) {
        const double lx = Lx[sp], ly = Ly[sp], lz = Lz[sp];
        const double lLen = std::sqrt(lx * lx + ly * ly + lz * lz), totalArea = arSum[sp];
        if (lLen < 1e-8 || totalArea < 1) {
            if (largest[sp] >= 0) {                      // a plate only becomes `largest` when it has a pole
                const int p = largest[sp];
                spPole[3 * sp] = pole[3 * p]; spPole[3 * sp + 1] = pole[3 * p + 1]; spPole[3 * sp + 2] = pole[3 * p + 2];
                spOmega[sp] = omega[p];
            } else {
                spPole[3 * sp] = 0; spPole[3 * sp + 1] = 1; spPole[3 * sp + 2] = 0; spOmega[sp] = 0;
            }
            continue;
        }
        spPole[3 * sp] = lx / lLen; spPole[3 * sp + 1] = ly / lLen; spPole[3 * sp + 2] = lz / lLen;
        spOmega[sp] = omSum[sp] / totalArea;
    }
    std::vector<double> oceanArea(nSP, 0), totalArea(nSP, 0), denSum(nSP, 0), denArea(nSP, 0);
    for (int p = 0; p < nSeeds; p++) {
        const int sp = plateToSuper[p];
        totalArea[sp] += plateArea[p];
        if (isOcean[p]) oceanArea[sp] += plateArea[p];
        if (density[p] == density[p]) { denSum[sp] += plateArea[p] * density[p]; denArea[sp] += plateArea[p]; }
    }
    for (int sp = 0; sp < nSP; sp++) {
        spIsOcean[sp] = oceanArea[sp] > totalArea[sp] * 0.5;
        spDensity[sp] = denArea[sp] > 0 ? denSum[sp] / denArea[sp] : 2.7;
    }
    return nSP;
}

// js/planet-worker.js:682-727 (sampleBilinear, grayscaleToElevation, sampleHeightmap)
void sample_heightmap(int N, const float* xyz, const uint8_t* pixels, int imgW, int imgH, float* r_elevation) {
    for (int r = 0; r < N; r++) {
        const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
        const double lat = pb_asin(js::max(-1, js::min(1, y)));
        const double lon = pb_atan2(x, z);
        const double px = (lon / PB_PI + 1) * 0.5 * imgW;
        double py = (0.5 - lat / PB_PI) * imgH;
        py = js::max(0, js::min(py, imgH - 1));
        const double x0 = std::floor(px), y0 = std::floor(py);
        const double x1 = std::fmod(x0 + 1, (double)imgW);
        const double y1 = js::min(y0 + 1, imgH - 1);
        const double fx = px - x0, fy = py - y0;
        const double xw = std::fmod(std::fmod(x0, (double)imgW) + imgW, (double)imgW);
        const double v00 = pixels[(size_t)(y0 * imgW + xw)], v10 = pixels[(size_t)(y0 * imgW + x1)];
        const double v01 = pixels[(size_t)(y1 * imgW + xw)], v11 = pixels[(size_t)(y1 * imgW + x1)];
        const double gray = v00 * (1 - fx) * (1 - fy) + v10 * fx * (1 - fy) + v01 * (1 - fx) * fy + v11 * fx * fy;
        r_elevation[r] = js::f32(gray < 1 ? -0.5 : std::sqrt((gray - 1) / 254));
    }
}

// js/planet-worker.js:733-769 (deriveSyntheticPlates): one plate per connected land mass / ocean basin, id = first region met
void derive_synthetic_plates(int N, const int32_t* off, const int32_t* adj, const float* elev, int32_t* r_plate) {
    for (int r = 0; r < N; r++) r_plate[r] = -1;
    std::vector<int> queue;
    for (int r = 0; r < N; r++) {
        if (r_plate[r] >= 0) continue;
        const bool isOcean = elev[r] <= 0;
        r_plate[r] = r;
        queue.assign(1, r);
        for (size_t head = 0; head < queue.size(); head++) {
            const int cur = queue[head];
            for (int ni = off[cur], e = off[cur + 1]; ni < e; ni++) {
                const int nb = adj[ni];
                if (r_plate[nb] >= 0) continue;
                if ((elev[nb] <= 0) == isOcean) { r_plate[nb] = r; queue.push_back(nb); }
            }
        }
    }
}

// js/planet-worker.js:811-831: region classes of an imported heightmap
void classify_imported(int N, const int32_t* off, const int32_t* adj, const float* elev, uint8_t* mountain, uint8_t* coastline, uint8_t* ocean) {
    for (int r = 0; r < N; r++) {
        mountain[r] = coastline[r] = ocean[r] = 0;
        if (elev[r] <= 0) ocean[r] = 1;
        else if ((double)elev[r] > 0.5) mountain[r] = 1;
        if (elev[r] > 0)
            for (int ni = off[r], e = off[r + 1]; ni < e; ni++) if (elev[adj[ni]] <= 0) { coastline[r] = 1; break; }
    }
}

}  // namespace

extern "C" {
void orc_sample_heightmap(int N, const float* xyz, const uint8_t* pixels, int imgW, int imgH, float* r_elevation) {
    sample_heightmap(N, xyz, pixels, imgW, imgH, r_elevation);
}
void orc_derive_synthetic_plates(int N, const int32_t* off, const int32_t* adj, const float* elev, int32_t* r_plate) {
    derive_synthetic_plates(N, off, adj, elev, r_plate);
}
void orc_classify_imported(int N, const int32_t* off, const int32_t* adj, const float* elev, uint8_t* mountain, uint8_t* coastline, uint8_t* ocean) {
    classify_imported(N, off, adj, elev, mountain, coastline, ocean);
}
void orc_project_coarse_plates(int N, const float* r_xyz, int NC, const int32_t* cOff, const int32_t* cAdj, const float* coarse_xyz,
                               const int32_t* coarse_r_plate, double seed, int numPlates, int32_t* r_plate) {
    project_coarse_plates(N, r_xyz, NC, cOff, cAdj, coarse_xyz, coarse_r_plate, seed, numPlates, r_plate);
}
void orc_smooth_and_reconnect_plates(int N, const int32_t* off, const int32_t* adj, int32_t* r_plate, const int32_t* seeds, int nSeeds,
                                     int numPasses) {
    smooth_and_reconnect(N, off, adj, r_plate, seeds, nSeeds, numPasses);
}
int orc_build_super_plates(int N, const int32_t* off, const int32_t* adj, const int32_t* r_plate, int nSeeds, const int32_t* seeds,
                           const uint8_t* hasVec, const double* pole, const double* omega, const uint8_t* isOcean, const double* density,
                           int32_t* r_superPlate, double* spPole, double* spOmega, uint8_t* spIsOcean, double* spDensity, int spCap) {
    return build_super_plates(N, off, adj, r_plate, nSeeds, seeds, hasVec, pole, omega, isOcean, density, r_superPlate, spPole, spOmega,
                              spIsOcean, spDensity, spCap);
}
}
