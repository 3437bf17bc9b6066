// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).
// Single-threaded restatement of js/terrain-post.js (reference) in double arithmetic with
// explicit f32 stores, following the reference's statement order.  Each function cites
// the lines it follows.  Math.pow / Math.asin go through pb_detmath.h (see that header).
#include <algorithm>
#include <cstring>
#include <limits>
#include "js_semantics.h"
#include "noise.h"
#include "terrain_post.h"

namespace {

// js/terrain-post.js:12-47 — binary min-heap of cell ids keyed by an external f32 array.
struct MinHeap {
    const float* key;
    std::vector<int32_t> data;
    explicit MinHeap(const float* k) : key(k) {}
    size_t size() const { return data.size(); }
    void push(int32_t cell) {
        data.push_back(cell);
        size_t i = data.size() - 1;
        while (i > 0) {
            size_t parent = (i - 1) >> 1;
            if (key[data[i]] >= key[data[parent]]) break;
            std::swap(data[i], data[parent]);
            i = parent;
        }
    }
    int32_t pop() {
        int32_t top = data[0];
        int32_t last = data.back();
        data.pop_back();
        if (!data.empty()) {
            data[0] = last;
            size_t i = 0, n = data.size();
            for (;;) {
                size_t smallest = i, l = 2 * i + 1, r = 2 * i + 2;
                if (l < n && key[data[l]] < key[data[smallest]]) smallest = l;
                if (r < n && key[data[r]] < key[data[smallest]]) smallest = r;
                if (smallest == i) break;
                std::swap(data[i], data[smallest]);
                i = smallest;
            }
        }
        return top;
    }
};

inline double smoothstep_g(double x, double e0, double e1) { // js/terrain-post.js:411-414
    double t = js::max(0, js::min(1, (x - e0) / (e1 - e0)));
    return t * t * (3 - 2 * t);
}

} // namespace

// js/terrain-post.js:100-105 — JS double/ToUint32 semantics (products exceed 2^53 and round first).
double oracle_cell_noise(double r) {
    double h = (double)js::to_uint32(r * 2654435761.0);
    int32_t x1 = (int32_t)(((uint32_t)h >> 16) ^ (uint32_t)h); // ^ yields a signed int32
    h = (double)js::to_uint32((double)x1 * (double)0x45d9f3b);
    uint32_t hu = (uint32_t)h;
    hu = (hu >> 16) ^ hu;
    return ((double)hu / 4294967295.0) * 0.01;
}

// js/terrain-post.js:59-215
void oracle_priority_flood_carve(const OMesh& mesh, float* r_elevation, const uint8_t* r_isOcean,
                                 double carveStrength, FloodDebug* dbg) {
    const int N = mesh.N;
    const int32_t* adjOffset = mesh.adjOffset;
    const int32_t* adjList = mesh.adjList;
    const double EPS = 1e-7;

    // :66-94 ocean components by DFS (stack pop), largest wins, first wins ties
    std::vector<int32_t> oceanLabel(N, -1);
    std::vector<int64_t> componentSizes;
    std::vector<int32_t> queue;
    for (int r = 0; r < N; r++) {
        if (!r_isOcean[r] || oceanLabel[r] >= 0) continue;
        int32_t label = (int32_t)componentSizes.size();
        int64_t size = 0;
        queue.clear(); queue.push_back(r);
        oceanLabel[r] = label;
        while (!queue.empty()) {
            int cur = queue.back(); queue.pop_back();
            size++;
            for (int i = adjOffset[cur], iEnd = adjOffset[cur + 1]; i < iEnd; i++) {
                int nb = adjList[i];
                if (r_isOcean[nb] && oceanLabel[nb] < 0) { oceanLabel[nb] = label; queue.push_back(nb); }
            }
        }
        componentSizes.push_back(size);
    }
    size_t mainOceanLabel = 0;
    for (size_t i = 1; i < componentSizes.size(); i++)
        if (componentSizes[i] > componentSizes[mainOceanLabel]) mainOceanLabel = i;
    std::vector<uint8_t> isOpenOcean(N, 0);
    for (int r = 0; r < N; r++)
        if (r_isOcean[r] && oceanLabel[r] == (int32_t)mainOceanLabel) isOpenOcean[r] = 1;

    // :107-113
    std::vector<float> surface(r_elevation, r_elevation + N);
    std::vector<int32_t> drainTo(N, -1);
    std::vector<uint8_t> visited(N, 0);
    std::vector<float> key(N);
    for (int r = 0; r < N; r++) key[r] = js::f32((double)r_elevation[r] + oracle_cell_noise(r));

    MinHeap heap(key.data());

    // :118-128 seed
    for (int r = 0; r < N; r++) {
        if (r_isOcean[r]) { visited[r] = 1; continue; }
        for (int i = adjOffset[r], iEnd = adjOffset[r + 1]; i < iEnd; i++) {
            if (isOpenOcean[adjList[i]]) {
                visited[r] = 1;
                drainTo[r] = adjList[i];
                heap.push(r);
                break;
            }
        }
    }

    // :131-147 pass 1
    while (heap.size() > 0) {
        int r = heap.pop();
        double surfR = surface[r];
        for (int i = adjOffset[r], iEnd = adjOffset[r + 1]; i < iEnd; i++) {
            int nb = adjList[i];
            if (visited[nb]) continue;
            visited[nb] = 1;
            drainTo[nb] = r;
            if ((double)r_elevation[nb] < surfR + EPS) {
                surface[nb] = js::f32(surfR + EPS);
                key[nb] = js::f32((double)surface[nb] + oracle_cell_noise(nb));
            }
            heap.push(nb);
        }
    }
    if (dbg) {
        if (dbg->drainTo) std::memcpy(dbg->drainTo, drainTo.data(), sizeof(int32_t) * N);
        if (dbg->surface) std::memcpy(dbg->surface, surface.data(), sizeof(float) * N);
        if (dbg->isOpenOcean) std::memcpy(dbg->isOpenOcean, isOpenOcean.data(), N);
    }

    // :152-196 pass 2
    std::vector<int32_t> path;
    for (int r = 0; r < N; r++) {
        if (r_isOcean[r]) continue;
        double deficit = (double)surface[r] - (double)r_elevation[r];
        if (deficit <= EPS) continue;

        path.clear();
        int peakIdx = -1;
        double peakElev = -std::numeric_limits<double>::infinity();
        int cur = r;
        while (cur >= 0 && !r_isOcean[cur]) {
            path.push_back(cur);
            if ((double)r_elevation[cur] > peakElev) { peakElev = r_elevation[cur]; peakIdx = (int)path.size() - 1; }
            cur = drainTo[cur];
        }
        if (peakIdx < 0 || path.empty()) continue;

        double carveAmount = deficit * carveStrength;
        int plen = (int)path.size();
        int radius = (int)js::max(3, std::ceil(plen * 0.3));
        int startIdx = std::max(0, peakIdx - radius);
        int endIdx = std::min(plen - 1, peakIdx + radius);

        double kernelSum = 0;
        for (int k = startIdx; k <= endIdx; k++) {
            double dist = std::abs(k - peakIdx);
            kernelSum += 1 - dist / (radius + 1);
        }
        if (kernelSum > 0) {
            for (int k = startIdx; k <= endIdx; k++) {
                double dist = std::abs(k - peakIdx);
                double weight = (1 - dist / (radius + 1)) / kernelSum;
                r_elevation[path[k]] = js::f32((double)r_elevation[path[k]] - carveAmount * weight);
                if (r_elevation[path[k]] < 0) r_elevation[path[k]] = 0;
            }
        }
        double fillAmount = deficit * (1 - carveStrength);
        r_elevation[r] = js::f32((double)r_elevation[r] + fillAmount);
    }

    // :200-214 pass 3
    std::vector<int32_t> order;
    order.reserve(N);
    for (int r = 0; r < N; r++) if (!r_isOcean[r]) order.push_back(r);
    std::stable_sort(order.begin(), order.end(),
                     [&](int32_t a, int32_t b) { return (double)surface[a] - (double)surface[b] < 0; });
    for (size_t i = 0; i < order.size(); i++) {
        int r = order[i];
        int target = drainTo[r];
        if (target < 0) continue;
        double targetElev = r_isOcean[target] ? 0.0 : (double)r_elevation[target];
        if ((double)r_elevation[r] <= targetElev) r_elevation[r] = js::f32(targetElev + EPS);
    }
}

// js/terrain-post.js:233-309
void oracle_warp_terrain(const OMesh& mesh, float* r_elevation, const float* r_xyz, double seed,
                         double strength, const float* r_hotspot) {
    if (strength <= 0) return;
    const int N = mesh.N;
    const int32_t* adjOffset = mesh.adjOffset;
    const int32_t* adjList = mesh.adjList;
    SimplexNoise noise(seed + 9999);
    const double freq = 4;
    const int octaves = 5;
    const double maxAmp = 0.12 * strength;

    std::vector<float> out(r_elevation, r_elevation + N);
    for (int r = 0; r < N; r++) {
        double px = r_xyz[3 * r], py = r_xyz[3 * r + 1], pz = r_xyz[3 * r + 2];
        double ex = -pz, ey = 0, ez = px;
        double elen = std::sqrt(ex * ex + ez * ez);
        if (elen > 1e-10) { ex /= elen; ez /= elen; } else { ex = 1; ez = 0; }
        double nx = py * ez;
        double ny = pz * ex - px * ez;
        double nz = -py * ex;
        double nlen = js::or_default(std::sqrt(nx * nx + ny * ny + nz * nz), 1);
        double nnx = nx / nlen, nny = ny / nlen, nnz = nz / nlen;

        double d1 = noise.fbm(px * freq, py * freq, pz * freq, octaves) * maxAmp;
        double d2 = noise.fbm(px * freq + 31.7, py * freq + 47.3, pz * freq + 19.1, octaves) * maxAmp;

        double wx = px + ex * d1 + nnx * d2;
        double wy = py + ey * d1 + nny * d2;
        double wz = pz + ez * d1 + nnz * d2;
        double wlen = js::or_default(std::sqrt(wx * wx + wy * wy + wz * wz), 1);
        wx /= wlen; wy /= wlen; wz /= wlen;

        int cur = r;
        double bestDot = wx * px + wy * py + wz * pz;
        for (;;) {
            bool moved = false;
            // bounds captured once while `cur` moves inside the loop (:276-284)
            for (int i = adjOffset[cur], iEnd = adjOffset[cur + 1]; i < iEnd; i++) {
                int nb = adjList[i];
                double dot = wx * r_xyz[3 * nb] + wy * r_xyz[3 * nb + 1] + wz * r_xyz[3 * nb + 2];
                if (dot > bestDot) { bestDot = dot; cur = nb; moved = true; }
            }
            if (!moved) break;
        }
        out[r] = r_elevation[cur];
    }

    const double warpBias = 0.25 + 0.5 * strength;
    for (int r = 0; r < N; r++) {
        double orig = r_elevation[r];
        double warped = out[r];
        double bias = warpBias;
        if (r_hotspot) {
            double hotFrac = js::min(1, std::fabs((double)r_hotspot[r]) / js::or_default(std::fabs(orig), 1));
            bias *= 1 - 0.8 * hotFrac;
        }
        if (warped > orig) r_elevation[r] = js::f32(orig + (warped - orig) * bias);
        else r_elevation[r] = js::f32(warped + (orig - warped) * (1 - bias));
    }
}

// js/terrain-post.js:317-354
void oracle_smooth_elevation(const OMesh& mesh, float* r_elevation, const uint8_t* r_isOcean,
                             int iterations, double strength) {
    const int N = mesh.N;
    const int32_t* adjOffset = mesh.adjOffset;
    const int32_t* adjList = mesh.adjList;
    std::vector<float> tmp(N, 0.f);
    std::vector<uint8_t> locked(N, 0);
    for (int r = 0; r < N; r++) {
        if (r_isOcean[r]) continue;
        for (int i = adjOffset[r], iEnd = adjOffset[r + 1]; i < iEnd; i++)
            if (r_isOcean[adjList[i]]) { locked[r] = 1; break; }
    }
    for (int iter = 0; iter < iterations; iter++) {
        for (int r = 0; r < N; r++) {
            if (locked[r]) { tmp[r] = r_elevation[r]; continue; }
            double h = r_elevation[r];
            double wSum = 0, hSum = 0;
            for (int i = adjOffset[r], iEnd = adjOffset[r + 1]; i < iEnd; i++) {
                double nh = r_elevation[adjList[i]];
                double diff = std::fabs(nh - h);
                double w = 1 / (1 + diff * 8);
                wSum += w;
                hSum += nh * w;
            }
            if (wSum > 0) {
                double avg = hSum / wSum;
                tmp[r] = js::f32(h + (avg - h) * strength);
            } else tmp[r] = js::f32(h);
        }
        for (int r = 0; r < N; r++) r_elevation[r] = tmp[r];
    }
}

// js/terrain-post.js:369-707
void oracle_erode_composite(const OMesh& mesh, float* r_elevation, const float* r_xyz,
                            const uint8_t* r_isOcean, int hIters, double K, double m, double dt,
                            int tIters, double talusSlope, double kThermal, int gIters,
                            double glacialStrength, const float* neighborDist, ErodeDebug* dbg) {
    // :375-379
    if (glacialStrength != glacialStrength) glacialStrength = 0;
    int totalIters = std::max(hIters, std::max(tIters, gIters));
    if (totalIters <= 0) return;

    const int N = mesh.N;
    const int32_t* adjOffset = mesh.adjOffset;
    const int32_t* adjList = mesh.adjList;

    std::vector<int32_t> landCells;
    for (int r = 0; r < N; r++) if (!r_isOcean[r]) landCells.push_back(r);
    const int landCount = (int)landCells.size();
    if (landCount == 0) return;

    std::vector<int32_t> drainTarget(N, 0);
    std::vector<float> cellDist(N, 0.f), flow(N, 0.f), delta(N, 0.f);

    if (hIters > 0) oracle_priority_flood_carve(mesh, r_elevation, r_isOcean, 0.5, nullptr);

    // :405-433
    bool haveGlac = false;
    std::vector<float> glacIdx, iceFlow;
    std::vector<int32_t> iceTarget;
    std::vector<uint8_t> numIceUpstream;
    if (gIters > 0 && glacialStrength > 0) {
        haveGlac = true;
        glacIdx.assign(N, 0.f);
        const double thresholdLat = PB_PI / 2 - glacialStrength * PB_PI / 4.5;
        for (int r = 0; r < N; r++) {
            if (r_isOcean[r]) continue;
            double y = r_xyz[3 * r + 1];
            double polarDist = std::fabs(pb_asin(js::max(-1, js::min(1, y))));
            double latFactor = smoothstep_g(polarDist, thresholdLat, PB_PI / 2);
            double elevFactor = smoothstep_g(r_elevation[r], 0.5, 0.9);
            double latScale = smoothstep_g(polarDist, PB_PI / 8, PB_PI / 3);
            glacIdx[r] = js::f32(js::max(latFactor, elevFactor * 0.3 * (0.3 + 0.7 * latScale)) * glacialStrength);
        }
        iceTarget.assign(N, 0);
        iceFlow.assign(N, 0.f);
        numIceUpstream.assign(N, 0);
    }

    // :436-442
    const double gScale = gIters > 0 ? 1.0 / gIters : 0;
    const double gCarveRate = 0.02 * gScale;
    const double gConvergenceBonus = 0.01 * gScale;
    const double gDepositAmount = 0.005 * gScale;
    const double gFjordCarve = 0.015 * gScale;
    const double gFlowThreshold = 0.1;
    const double gFjordThreshold = 0.5;

    const int midFloodIter = (int)js::round(totalIters * 0.75);
    bool midFloodDone = false;

    int maxDeg = 0;
    for (int r = 0; r < N; r++) maxDeg = std::max(maxDeg, adjOffset[r + 1] - adjOffset[r]);
    std::vector<int32_t> excNb(maxDeg + 1);
    std::vector<float> excVal(maxDeg + 1);

    auto sortDesc = [&]() {
        std::stable_sort(landCells.begin(), landCells.end(), [&](int32_t a, int32_t b) {
            return (double)r_elevation[b] - (double)r_elevation[a] < 0;
        });
    };

    for (int iter = 0; iter < totalIters; iter++) {
        if (!midFloodDone && iter >= midFloodIter) {
            midFloodDone = true;
            oracle_priority_flood_carve(mesh, r_elevation, r_isOcean, 0.85, nullptr);
        }
        const bool glacialThisIter = iter < gIters && haveGlac;
        const bool hydraulicThisIter = iter < hIters;
        if (glacialThisIter || hydraulicThisIter) sortDesc();

        // ---- glacial :475-557
        if (glacialThisIter) {
            std::fill(iceTarget.begin(), iceTarget.end(), -1);
            std::fill(numIceUpstream.begin(), numIceUpstream.end(), 0);
            for (int i = 0; i < landCount; i++) {
                int r = landCells[i];
                if (glacIdx[r] <= 0) continue;
                double h = r_elevation[r];
                int bestNb = -1; double bestDrop = 0;
                for (int j = adjOffset[r], jEnd = adjOffset[r + 1]; j < jEnd; j++) {
                    int nb = adjList[j];
                    double drop = h - (double)r_elevation[nb];
                    if (drop > bestDrop) { bestDrop = drop; bestNb = nb; }
                }
                if (bestNb >= 0) iceTarget[r] = bestNb;
            }
            for (int r = 0; r < N; r++) iceFlow[r] = glacIdx[r];
            for (int i = 0; i < landCount; i++) {
                int r = landCells[i];
                int target = iceTarget[r];
                if (target >= 0 && iceFlow[r] > 0) {
                    iceFlow[target] = js::f32((double)iceFlow[target] + (double)iceFlow[r]);
                    numIceUpstream[target]++; // Uint8Array wraps
                }
            }
            for (int i = 0; i < landCount; i++) {
                int r = landCells[i];
                if (iceFlow[r] <= gFlowThreshold) continue;
                double deepening = gCarveRate * pb_pow(iceFlow[r], 0.6) * glacialStrength;
                r_elevation[r] = js::f32((double)r_elevation[r] - deepening);
                for (int j = adjOffset[r], jEnd = adjOffset[r + 1]; j < jEnd; j++) {
                    int nb = adjList[j];
                    if (r_isOcean[nb]) continue;
                    double d = js::or_default(neighborDist[j], 1e-6);
                    double slope = std::fabs((double)r_elevation[r] - (double)r_elevation[nb]) / d;
                    r_elevation[nb] = js::f32((double)r_elevation[nb] - deepening * 0.4 * js::max(0, 1 - slope));
                }
                if (numIceUpstream[r] >= 2)
                    r_elevation[r] = js::f32((double)r_elevation[r] - gConvergenceBonus * pb_pow(iceFlow[r], 0.4));
            }
            for (int i = 0; i < landCount; i++) {
                int r = landCells[i];
                if (iceFlow[r] <= gFlowThreshold) continue;
                int target = iceTarget[r];
                if (target < 0 || r_isOcean[target]) continue;
                if ((double)glacIdx[target] < (double)glacIdx[r] * 0.3)
                    r_elevation[target] = js::f32((double)r_elevation[target] + gDepositAmount * pb_pow(iceFlow[r], 0.3));
            }
            for (int r = 0; r < N; r++) {
                if (r_isOcean[r]) continue;
                if (glacIdx[r] <= 0.2 || iceFlow[r] <= gFjordThreshold) continue;
                bool isCoastal = false;
                for (int j = adjOffset[r], jEnd = adjOffset[r + 1]; j < jEnd; j++)
                    if (r_isOcean[adjList[j]]) { isCoastal = true; break; }
                if (isCoastal) {
                    r_elevation[r] = js::f32((double)r_elevation[r] - gFjordCarve * pb_pow(iceFlow[r], 0.5));
                    if (r_elevation[r] < 0) r_elevation[r] = 0;
                }
            }
            for (int r = 0; r < N; r++)
                if (!r_isOcean[r] && r_elevation[r] < 0) r_elevation[r] = 0;
        }

        // ---- hydraulic :560-642
        if (hydraulicThisIter) {
            if (glacialThisIter) sortDesc();
            std::fill(drainTarget.begin(), drainTarget.end(), -1);
            for (int i = 0; i < landCount; i++) {
                int r = landCells[i];
                double h = r_elevation[r];
                int bestNb = -1, bestJ = -1;
                double bestDrop = -std::numeric_limits<double>::infinity();
                for (int j = adjOffset[r], jEnd = adjOffset[r + 1]; j < jEnd; j++) {
                    int nb = adjList[j];
                    double drop = h - (double)r_elevation[nb];
                    if (drop > bestDrop) { bestDrop = drop; bestNb = nb; bestJ = j; }
                }
                if (bestDrop <= 0) {
                    double minAscent = std::numeric_limits<double>::infinity();
                    for (int j = adjOffset[r], jEnd = adjOffset[r + 1]; j < jEnd; j++) {
                        int nb = adjList[j];
                        double ascent = (double)r_elevation[nb] - h;
                        if (ascent < minAscent) { minAscent = ascent; bestNb = nb; bestJ = j; }
                    }
                }
                if (bestNb >= 0) {
                    drainTarget[r] = bestNb;
                    cellDist[r] = js::f32(js::or_default(neighborDist[bestJ], 1e-6));
                }
            }
            std::fill(flow.begin(), flow.end(), 0.f);
            for (int i = 0; i < landCount; i++) flow[landCells[i]] = 1;
            for (int i = 0; i < landCount; i++) {
                int r = landCells[i];
                int target = drainTarget[r];
                if (target >= 0) flow[target] = js::f32((double)flow[target] + (double)flow[r]);
            }
            if (dbg && iter == dbg->captureIter) {
                if (dbg->drainTarget) std::memcpy(dbg->drainTarget, drainTarget.data(), sizeof(int32_t) * N);
                if (dbg->flow) std::memcpy(dbg->flow, flow.data(), sizeof(float) * N);
                if (dbg->landOrder) std::memcpy(dbg->landOrder, landCells.data(), sizeof(int32_t) * landCount);
            }
            for (int i = landCount - 1; i >= 0; i--) {
                int r = landCells[i];
                int target = drainTarget[r];
                if (target < 0 || cellDist[r] <= 0) continue;
                double factor = K * pb_pow(flow[r], m) * dt / (double)cellDist[r];
                double h_receiver = js::max(r_elevation[target], 0);
                double h_new = ((double)r_elevation[r] + factor * h_receiver) / (1 + factor);
                if (h_new < h_receiver) h_new = h_receiver;
                if (h_new < 0) h_new = 0;
                double eroded = (double)r_elevation[r] - h_new;
                if (eroded > 0 && !r_isOcean[target]) {
                    int drainOfTarget = drainTarget[target];
                    double receiverSlope = 0;
                    if (drainOfTarget >= 0 && cellDist[target] > 0)
                        receiverSlope = std::fabs((double)r_elevation[target] - (double)r_elevation[drainOfTarget]) / (double)cellDist[target];
                    double depositFrac = 0.5 / (1 + receiverSlope * 50);
                    double deposit = eroded * depositFrac;
                    r_elevation[target] = js::f32((double)r_elevation[target] + deposit);
                    if ((double)r_elevation[target] > h_new) r_elevation[target] = js::f32(h_new);
                }
                r_elevation[r] = js::f32(h_new);
            }
        }

        // ---- thermal :645-686
        if (iter < tIters) {
            std::fill(delta.begin(), delta.end(), 0.f);
            for (int i = 0; i < landCount; i++) {
                int r = landCells[i];
                double h = r_elevation[r];
                double totalExcess = 0;
                int excCount = 0;
                for (int j = adjOffset[r], jEnd = adjOffset[r + 1]; j < jEnd; j++) {
                    int nb = adjList[j];
                    if (r_isOcean[nb]) continue;
                    double nh = r_elevation[nb];
                    if (nh >= h) continue;
                    double d = js::or_default(neighborDist[j], 1e-6);
                    double slope = (h - nh) / d;
                    if (slope > talusSlope) {
                        double excess = (slope - talusSlope) * d;
                        excNb[excCount] = nb;
                        excVal[excCount] = js::f32(excess);
                        excCount++;
                        totalExcess += excess;
                    }
                }
                if (totalExcess <= 0) continue;
                double transfer = kThermal * totalExcess * 0.5;
                for (int k = 0; k < excCount; k++) {
                    double share = ((double)excVal[k] / totalExcess) * transfer;
                    delta[r] = js::f32((double)delta[r] - share);
                    delta[excNb[k]] = js::f32((double)delta[excNb[k]] + share);
                }
            }
            for (int i = 0; i < landCount; i++)
                r_elevation[landCells[i]] = js::f32((double)r_elevation[landCells[i]] + (double)delta[landCells[i]]);
        }
    }

    // :690-706
    if (haveGlac) {
        std::vector<float> tmp(r_elevation, r_elevation + N);
        for (int r = 0; r < N; r++) {
            if (r_isOcean[r] || glacIdx[r] <= 0) continue;
            double sum = 0; int count = 0;
            for (int j = adjOffset[r], jEnd = adjOffset[r + 1]; j < jEnd; j++)
                if (!r_isOcean[adjList[j]]) { sum += r_elevation[adjList[j]]; count++; }
            if (count > 0) {
                double avg = sum / count;
                tmp[r] = js::f32((double)r_elevation[r] + (avg - (double)r_elevation[r]) * 0.3);
            }
        }
        for (int r = 0; r < N; r++)
            if (!r_isOcean[r] && glacIdx[r] > 0) r_elevation[r] = tmp[r];
    }
}

// js/terrain-post.js:713-751
void oracle_sharpen_ridges(const OMesh& mesh, float* r_elevation, const uint8_t* r_isOcean,
                           int iterations, double strength) {
    const int N = mesh.N;
    const int32_t* adjOffset = mesh.adjOffset;
    const int32_t* adjList = mesh.adjList;
    std::vector<int32_t> landCells;
    for (int r = 0; r < N; r++) if (!r_isOcean[r]) landCells.push_back(r);
    std::vector<float> tmp(N, 0.f), original(r_elevation, r_elevation + N);
    for (int iter = 0; iter < iterations; iter++) {
        for (int r : landCells) {
            double h = r_elevation[r];
            double sum = 0;
            int count = adjOffset[r + 1] - adjOffset[r];
            for (int i = adjOffset[r], iEnd = adjOffset[r + 1]; i < iEnd; i++) sum += r_elevation[adjList[i]];
            if (count == 0) { tmp[r] = js::f32(h); continue; }
            double avg = sum / count;
            if (h > avg) {
                double h_new = h + (h - avg) * strength;
                double cap = (double)original[r] * 1.5;
                if (h_new > cap) h_new = cap;
                tmp[r] = js::f32(h_new);
            } else tmp[r] = js::f32(h);
        }
        for (int r : landCells) r_elevation[r] = tmp[r];
    }
}

// js/terrain-post.js:758-794
void oracle_apply_soil_creep(const OMesh& mesh, float* r_elevation, const uint8_t* r_isOcean,
                             int iterations, double strength) {
    const int N = mesh.N;
    const int32_t* adjOffset = mesh.adjOffset;
    const int32_t* adjList = mesh.adjList;
    std::vector<int32_t> interiorLand;
    for (int r = 0; r < N; r++) {
        if (r_isOcean[r]) continue;
        bool coastal = false;
        for (int i = adjOffset[r], iEnd = adjOffset[r + 1]; i < iEnd; i++)
            if (r_isOcean[adjList[i]]) { coastal = true; break; }
        if (!coastal) interiorLand.push_back(r);
    }
    std::vector<float> tmp(N, 0.f);
    for (int iter = 0; iter < iterations; iter++) {
        for (int r : interiorLand) {
            double h = r_elevation[r];
            double sum = 0; int count = 0;
            for (int i = adjOffset[r], iEnd = adjOffset[r + 1]; i < iEnd; i++)
                if (!r_isOcean[adjList[i]]) { sum += r_elevation[adjList[i]]; count++; }
            if (count == 0) { tmp[r] = js::f32(h); continue; }
            double avg = sum / count;
            tmp[r] = js::f32(h + (avg - h) * strength);
        }
        for (int r : interiorLand) r_elevation[r] = tmp[r];
    }
}

// js/planet-worker.js:40-102 — order + slider→parameter mapping.  hItersOverride > 0 replaces
// round(20*hydraulic) while keeping K = 0.0006*hydraulic (BASELINE configs 2 and 5).
void oracle_run_post_processing(const OMesh& mesh, const float* r_xyz, float* r_elevation,
                                const PostParams& p, const float* neighborDist, double seed,
                                const float* r_hotspot, float* erosionDelta, uint8_t* isOceanOut) {
    const int N = mesh.N;
    if (p.terrainWarp > 0) oracle_warp_terrain(mesh, r_elevation, r_xyz, seed, p.terrainWarp, r_hotspot);
    std::vector<uint8_t> r_isOcean(N, 0);
    for (int r = 0; r < N; r++) if (r_elevation[r] <= 0) r_isOcean[r] = 1;
    std::vector<float> preErosion(r_elevation, r_elevation + N);
    if (p.smoothing > 0) {
        int smoothIters = (int)js::round(1 + p.smoothing * 4);
        double smoothStr = 0.2 + p.smoothing * 0.5;
        oracle_smooth_elevation(mesh, r_elevation, r_isOcean.data(), smoothIters, smoothStr);
    }
    if (p.glacialErosion > 0 || p.hydraulicErosion > 0 || p.thermalErosion > 0) {
        int gIters = (int)js::round(p.glacialErosion * 10);
        int hIters = p.hItersOverride > 0 ? p.hItersOverride : (int)js::round(p.hydraulicErosion * 20);
        double hK = p.hydraulicErosion * 0.0006;
        int tIters = (int)js::round(p.thermalErosion * 10);
        double talusSlope = 1.2 - p.thermalErosion * 0.4;
        double kThermal = p.thermalErosion * 0.15;
        oracle_erode_composite(mesh, r_elevation, r_xyz, r_isOcean.data(), hIters, hK, 0.5, 1.0, tIters,
                               talusSlope, kThermal, gIters, p.glacialErosion, neighborDist, nullptr);
    }
    if (p.ridgeSharpening > 0) {
        int rsIters = (int)js::round(1 + p.ridgeSharpening * 3);
        double rsStr = p.ridgeSharpening * 0.08;
        oracle_sharpen_ridges(mesh, r_elevation, r_isOcean.data(), rsIters, rsStr);
    }
    oracle_apply_soil_creep(mesh, r_elevation, r_isOcean.data(), 3, 0.1125);
    if (erosionDelta)
        for (int r = 0; r < N; r++) erosionDelta[r] = js::f32((double)r_elevation[r] - (double)preErosion[r]);
    if (isOceanOut) std::memcpy(isOceanOut, r_isOcean.data(), N);
}

// js/sphere-mesh.js:191-203
void oracle_compute_neighbor_dist(const OMesh& mesh, const float* r_xyz, float* neighborDist) {
    for (int r = 0; r < mesh.N; r++) {
        double x = r_xyz[3 * r], y = r_xyz[3 * r + 1], z = r_xyz[3 * r + 2];
        for (int i = mesh.adjOffset[r]; i < mesh.adjOffset[r + 1]; i++) {
            int nb = mesh.adjList[i];
            double dx = x - r_xyz[3 * nb], dy = y - r_xyz[3 * nb + 1], dz = z - r_xyz[3 * nb + 2];
            neighborDist[i] = js::f32(std::sqrt(dx * dx + dy * dy + dz * dz));
        }
    }
}
