// pb_coarse.h — generateCoarsePlates (js/coarse-plates.js:19-39): the coarse stage that feeds the plate pipeline.
//
//   buildSphere(20000, 0.75, makeRng(seed + 137))   device (pb_meshgen.h)
//   generatePlates   js/plates.js:6-232             host: seed placement and plate growth are one chain of RNG draws whose
//                                                   ranges depend on the live frontier sizes (SURVEY class R); the closing
//                                                   smoothAndReconnectPlates runs on the device (pb_plates.h)
//   assignOceanLand  js/ocean-land.js:7-238         host: per-plate scans + logic on the plate graph (tens of plates)
// The coarse mesh has 20 001 regions whatever the planet's resolution, so this stage is a fixed cost of a few ms.
#pragma once
#include "pb_meshgen.h"
#include "pb_plates.h"

namespace pb {

struct CoarsePlatesResult {
    std::vector<int> r_plate;           // per coarse region: plate seed id
    std::vector<int> seeds;             // plateSeeds in Set order
    std::vector<double> pole, omega, density;
    std::vector<uint8_t> isOcean;
};

class CoarseStage {
    // js/rng.js:3-6 — the state is an integer below 2^31, so 64-bit integer arithmetic reproduces the JS doubles exactly
    struct Lcg {
        unsigned long long s;
        explicit Lcg(double seed) { s = (unsigned long long)(fmod(fabs(floor(seed * 9301.0 + 49297.0)), 2147483646.0) + 1.0); }
        double operator()() { s = (s * 16807ull) % 2147483647ull; return (double)(s - 1) / 2147483646.0; }
        long long below(double n) { return (long long)floor((*this)() * n); }      // makeRandInt
    };
    static double or1(double v) { return (v == 0 || v != v) ? 1.0 : v; }           // `x || 1`
    static double clamp01(double v) { return v < 0 ? 0 : (v > 1 ? 1 : v); }

    // three farthest non-seed regions (js/plates.js:28-43, 53-68)
    struct Podium {
        int r[3] = {-1, -1, -1}; double d[3] = {-1, -1, -1};
        void see(int region, double dist) {
            if (!(dist > d[2])) return;
            int at = dist > d[0] ? 0 : (dist > d[1] ? 1 : 2);
            for (int k = 2; k > at; k--) { r[k] = r[k - 1]; d[k] = d[k - 1]; }
            r[at] = region; d[at] = dist;
        }
        int count() const { return (r[0] != -1) + (r[1] != -1) + (r[2] != -1); }
    };

public:
    // off/adj/xyz: host CSR and coordinates of the coarse mesh; `smooth` runs smoothAndReconnectPlates on it
    template <class Smooth>
    static void generate_plates(int N, const int* off, const int* adj, const float* xyz, int numPlates, double seed, Smooth&& smooth,
                                CoarsePlatesResult& R) {
        std::vector<int>& plate = R.r_plate;
        plate.assign(N, -1);
        Lcg rng(seed + 0.5), ints(seed);
        std::vector<char> seeded(N, 0);
        std::vector<float> nearest(N);
        auto far = [&](int r, const double c[3]) { return 1 - ((double)xyz[3 * r] * c[0] + (double)xyz[3 * r + 1] * c[1] + (double)xyz[3 * r + 2] * c[2]); };
        auto centre = [&](int r, double c[3]) { c[0] = xyz[3 * r]; c[1] = xyz[3 * r + 1]; c[2] = xyz[3 * r + 2]; };
        auto plant = [&](int r) { R.seeds.push_back(r); seeded[r] = 1; };
        auto relax = [&](const double c[3]) { for (int r = 0; r < N; r++) { const double d = far(r, c); if (d < (double)nearest[r]) nearest[r] = (float)d; } };
        double c[3];
        const int firstSeed = (int)ints.below(N);
        plant(firstSeed);
        centre(firstSeed, c);
        for (int r = 0; r < N; r++) nearest[r] = (float)far(r, c);
        nearest[firstSeed] = 0;
        while ((int)R.seeds.size() < numPlates && (int)R.seeds.size() < N) {
            Podium top;
            for (int r = 0; r < N; r++) if (!seeded[r]) top.see(r, nearest[r]);
            if (!top.count()) break;
            const int a = top.r[ints.below(top.count())];
            plant(a);
            centre(a, c);
            if ((int)R.seeds.size() < numPlates) {
                Podium next;                                   // fused: relax distances and rank in one pass (:50-69)
                for (int r = 0; r < N; r++) {
                    const double d = far(r, c);
                    if (d < (double)nearest[r]) nearest[r] = (float)d;
                    if (!seeded[r]) next.see(r, nearest[r]);
                }
                if (!next.count()) break;
                const int b = next.r[ints.below(next.count())];
                plant(b);
                centre(b, c);
            }
            relax(c);
        }
        const int P = (int)R.seeds.size();
        const double lowT = clamp01((80 - numPlates) / 60.0);
        // per-plate growth rate, preferred direction, directional strength (:99-113)
        std::vector<double> rate(P), strength(P), dir(3 * (size_t)P);
        for (int k = 0; k < P; k++) {
            const double u = rng(), v = rng();
            rate[k] = (0.7 - 0.4 * lowT) + u * v * (2.3 + 2.4 * lowT);
            centre(R.seeds[k], c);
            const double len = or1(sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]));
            const double n[3] = {c[0] / len, c[1] / len, c[2] / len};
            const double q0 = rng() - 0.5, q1 = rng() - 0.5, q2 = rng() - 0.5;
            const double along = q0 * n[0] + q1 * n[1] + q2 * n[2];
            const double t[3] = {q0 - along * n[0], q1 - along * n[1], q2 - along * n[2]};
            const double tl = or1(sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]));
            for (int i = 0; i < 3; i++) dir[3 * k + i] = t[i] / tl;
            const double s = rng() * ((0.15 + 0.25 * lowT) + (0.25 + 0.25 * lowT) / rate[k]);
            strength[k] = s < 0.85 ? s : 0.85;
        }
        // round-robin growth (:115-191)
        std::vector<std::vector<int>> front(P);
        std::vector<double> size(P, 1);
        for (int k = 0; k < P; k++) { plate[R.seeds[k]] = R.seeds[k]; front[k].push_back(R.seeds[k]); }
        long long unclaimed = (long long)N - P;
        const double compactW4 = (0.3 - 0.22 * lowT) * 4;
        double fair = (double)(N - P) / numPlates; if (fair < 1) fair = 1;
        const double governor = fair * (2.0 + 2.0 * lowT), invN = 1.0 / N;
        while (unclaimed > 0) {
            bool grew = false;
            for (int k = 0; k < P; k++) {
                std::vector<int>& f = front[k];
                if (f.empty()) continue;
                const int pid = R.seeds[k];
                const double st = strength[k];
                double steps = ceil(rate[k] * (0.5 + rng())); if (steps < 1) steps = 1;
                if (size[k] > governor) { steps = ceil(steps * 0.5); if (steps < 1) steps = 1; }
                const double reach = sqrt(or1(size[k]) * invN / PB_PI) * 2 * 1.8;
                centre(pid, c);
                for (int s = 0; s < steps && !f.empty(); s++) {
                    int pickAt = 0; double pickScore = -INFINITY;
                    double want = 3 + floor(st * 5); if ((double)f.size() < want) want = (double)f.size();
                    for (int i = 0; i < (int)want; i++) {
                        const int at = (int)ints.below((double)f.size());
                        const int cell = f[at];
                        const double dx = xyz[3 * cell] - c[0], dy = xyz[3 * cell + 1] - c[1], dz = xyz[3 * cell + 2] - c[2];
                        const double d2 = dx * dx + dy * dy + dz * dz;
                        const double align = (dx * dir[3 * k] + dy * dir[3 * k + 1] + dz * dir[3 * k + 2]) / or1(sqrt(d2));
                        double excess = d2 * 0.5 - reach; if (!(excess > 0)) excess = 0;
                        const double score = align * st + rng() * (1 - st * 0.5) - excess * compactW4;
                        if (score > pickScore) { pickScore = score; pickAt = at; }
                    }
                    const int cur = f[pickAt];
                    f[pickAt] = f.back();
                    f.pop_back();
                    for (int j = off[cur]; j < off[cur + 1]; j++) {
                        const int nb = adj[j];
                        if (plate[nb] != -1) continue;
                        plate[nb] = pid; f.push_back(nb); size[k] += 1; unclaimed--; grew = true;
                    }
                }
            }
            if (!grew) break;
        }
        // orphans take the plate of their first claimed neighbour, repeated until nothing changes (:194-208)
        for (bool again = true; again;) {
            again = false;
            for (int r = 0; r < N; r++) {
                if (plate[r] != -1) continue;
                for (int j = off[r]; j < off[r + 1]; j++) if (plate[adj[j]] != -1) { plate[r] = plate[adj[j]]; again = true; break; }
            }
        }
        smooth(plate, R.seeds, (int)floor((3 - 2 * lowT) + 0.5));
        // Euler poles (:213-229)
        R.pole.resize(3 * (size_t)P); R.omega.resize(P);
        for (int k = 0; k < P; k++) {
            const double theta = rng() * 2 * PB_PI;
            const double cosP = 2 * rng() - 1;
            const double sinP = sqrt(1 - cosP * cosP);
            R.pole[3 * k] = sinP * pb_cos(theta); R.pole[3 * k + 1] = sinP * pb_sin(theta); R.pole[3 * k + 2] = cosP;
            const double w = 0.5 + rng() * 1.5;
            R.omega[k] = rng() < 0.5 ? -w : w;
        }
    }

    // js/ocean-land.js:7-238
    static void ocean_land(int N, const int* off, const int* adj, const float* xyz, double seed, int numContinents, double variety,
                           double landCoverage, CoarsePlatesResult& R) {
        Lcg rng(seed + 42);
        const int P = (int)R.seeds.size();
        const std::vector<int>& plate = R.r_plate;
        int maxId = 0;
        for (int s : R.seeds) if (s > maxId) maxId = s;
        std::vector<int> slot((size_t)maxId + 1, -1);
        for (int k = 0; k < P; k++) slot[R.seeds[k]] = k;
        struct PlateStat { double area = 0, cx = 0, cy = 0, cz = 0, perim = 0, compact = 0; std::vector<int> nbr; int continent = -1; };
        std::vector<PlateStat> S(P);
        for (int r = 0; r < N; r++) {
            PlateStat& s = S[slot[plate[r]]];
            s.area += 1; s.cx += xyz[3 * r]; s.cy += xyz[3 * r + 1]; s.cz += xyz[3 * r + 2];
        }
        for (PlateStat& s : S) { const double a = or1(s.area); s.cx /= a; s.cy /= a; s.cz /= a; }
        for (int r = 0; r < N; r++) {
            const int me = slot[plate[r]];
            bool edge = false;
            for (int j = off[r]; j < off[r + 1]; j++) {
                const int other = plate[adj[j]];
                if (other == plate[r]) continue;
                const int o = slot[other];
                std::vector<int>& nb = S[me].nbr;            // Set: insertion order, no duplicates
                if (std::find(nb.begin(), nb.end(), o) == nb.end()) nb.push_back(o);
                edge = true;
            }
            if (edge) S[me].perim += 1;
        }
        double best = 0;
        for (PlateStat& s : S) { s.compact = sqrt(or1(s.area)) / or1(s.perim); if (s.compact > best) best = s.compact; }
        if (best > 0) for (PlateStat& s : S) s.compact /= best;

        const double landBudget = landCoverage * N;
        struct Scored { int k; double score; };
        auto pick_top3 = [&](std::vector<Scored>& v) {
            std::stable_sort(v.begin(), v.end(), [](const Scored& a, const Scored& b) { return a.score > b.score; });
            const int top = (int)v.size() < 3 ? (int)v.size() : 3;
            return v[(size_t)floor(rng() * top)].k;
        };
        // continent seeds: farthest-point sampling over plate centroids (:68-98)
        std::vector<int> roots;
        std::vector<char> taken(P, 0);
        const int wanted = numContinents < P ? numContinents : P;
        const int firstRoot = (int)floor(rng() * P);
        roots.push_back(firstRoot); taken[firstRoot] = 1;
        for (int s = 1; s < wanted; s++) {
            std::vector<Scored> cand;
            for (int k = 0; k < P; k++) {
                if (taken[k]) continue;
                double nearestRoot = INFINITY;
                for (int e : roots) {
                    const double dx = S[k].cx - S[e].cx, dy = S[k].cy - S[e].cy, dz = S[k].cz - S[e].cz;
                    const double d = dx * dx + dy * dy + dz * dz;
                    if (d < nearestRoot) nearestRoot = d;
                }
                const double raw = sqrt((double)N / P) / sqrt(or1(S[k].area));
                cand.push_back({k, nearestRoot * (1 + (raw - 1) * (1 - variety * 0.5)) * (0.3 + 0.7 * S[k].compact)});
            }
            if (cand.empty()) break;
            const int k = pick_top3(cand);
            roots.push_back(k); taken[k] = 1;
        }
        double land = 0;
        for (int k : roots) land += S[k].area;
        while (roots.size() > 1 && land > landBudget) {          // trim the largest seeds (:101-110)
            size_t big = 0;
            for (size_t i = 1; i < roots.size(); i++) if (S[roots[i]].area > S[roots[big]].area) big = i;
            land -= S[roots[big]].area;
            roots.erase(roots.begin() + (long)big);
        }
        const int C = (int)roots.size();
        for (int c = 0; c < C; c++) S[roots[c]].continent = c;
        const double goal = landBudget * 0.9;
        std::vector<double> quota(C), held(C);
        for (int c = 0; c < C; c++) held[c] = S[roots[c]].area;
        if (variety > 0 && C > 1) {
            std::vector<double> w(C);
            double sum = 0;
            for (int c = 0; c < C; c++) w[c] = pb_exp((rng() - 0.5) * variety * 2.5);
            for (int c = 0; c < C; c++) sum = sum + w[c];
            for (int c = 0; c < C; c++) quota[c] = goal * w[c] / sum;
        } else {
            for (int c = 0; c < C; c++) quota[c] = goal / (C > 1 ? C : 1);
        }
        for (bool moved = true; moved && land < goal;) {          // round-robin continent growth (:144-177)
            moved = false;
            for (int c = 0; c < C && land < goal; c++) {
                if (held[c] >= quota[c]) continue;
                std::vector<Scored> cand;
                for (int k = 0; k < P; k++) {
                    if (S[k].continent != -1) continue;
                    bool mine = false, foreign = false; int shared = 0;
                    for (int o : S[k].nbr) {
                        const int oc = S[o].continent;
                        if (oc == c) { mine = true; shared++; }
                        else if (oc != -1) { foreign = true; break; }
                    }
                    if (mine && !foreign) cand.push_back({k, shared + S[k].compact * 3 + rng() * 0.5});
                }
                if (cand.empty()) continue;
                const int k = pick_top3(cand);
                S[k].continent = c; held[c] += S[k].area; land += S[k].area; moved = true;
            }
        }
        // trapped seas (:180-228): every ocean component but the largest, if it touches exactly one continent and fits
        std::vector<std::vector<int>> seas;
        std::vector<char> seen(P, 0);
        for (int k = 0; k < P; k++) {
            if (S[k].continent != -1 || seen[k]) continue;
            std::vector<int> sea{k};
            seen[k] = 1;
            for (size_t h = 0; h < sea.size(); h++)
                for (int o : S[sea[h]].nbr) if (S[o].continent == -1 && !seen[o]) { seen[o] = 1; sea.push_back(o); }
            seas.push_back(sea);
        }
        auto extent = [&](const std::vector<int>& sea) { double a = 0; for (int k : sea) a += S[k].area; return a; };
        size_t mainSea = 0;
        for (size_t i = 1; i < seas.size(); i++) if (extent(seas[i]) > extent(seas[mainSea])) mainSea = i;
        for (size_t i = 0; i < seas.size(); i++) {
            if (i == mainSea) continue;
            std::vector<int> shores;
            for (int k : seas[i]) {
                for (int o : S[k].nbr) {
                    const int oc = S[o].continent;
                    if (oc != -1 && std::find(shores.begin(), shores.end(), oc) == shores.end()) shores.push_back(oc);
                }
                if (shores.size() > 1) break;
            }
            if (shores.size() != 1) continue;
            const double a = extent(seas[i]);
            if (land + a <= landBudget * 1.1) { for (int k : seas[i]) S[k].continent = shores[0]; land += a; }
        }
        R.isOcean.resize(P);
        for (int k = 0; k < P; k++) R.isOcean[k] = S[k].continent == -1;
        // plate densities as the worker draws them (js/planet-worker.js:196-201)
        R.density.resize(P);
        for (int k = 0; k < P; k++) {
            Lcg d((double)R.seeds[k] + 777);
            const double oceanic = 3.0 + d() * 0.5, continental = 2.4 + d() * 0.5;
            R.density[k] = R.isOcean[k] ? oceanic : continental;
        }
    }
};

}  // namespace pb
