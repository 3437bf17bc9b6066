// pb_elevation_engine.h — assignElevation (js/elevation.js:216-1391) orchestration.
//
// Stage split:
//   device  findCollisions ×1|2, the main synthesis loop, coastal roughening, island-arc and hotspot
//           uplift, peak compression                                               (class P kernels)
//   host    the order-dependent middle of the function: dual-layer set unions and stress blends,
//           propagateStress (in-place frontier order, A.7), plate representatives, the five
//           assignDistanceField fills (Park–Miller-driven random queue: class R, inherently serial —
//           run as five concurrent host threads, A.6), the six capped FIFO BFS with first-discoverer
//           payloads (A.8), the p97 stress normaliser and the ≤ 35 RNG-placed hotspot domes.
// The host part is the engine's own code (not the oracle) and is what SURVEY lists as "host-serial".
#pragma once
#include <chrono>
#include <thread>
#include <unordered_map>
#include "pb_engine.h"
#include "pb_elevation.h"

namespace pb {

struct PlateTableHost {     // rows in the caller's order; ids are arbitrary non-negative ints
    int n = 0;
    std::vector<int> ids; std::vector<uint8_t> isOcean; std::vector<double> pole, omega, density;
    std::vector<int> index;     // dense id → row
    void set(int n_, const int* ids_, const uint8_t* oc, const double* pole_, const double* om, const double* de) {
        n = n_;
        ids.assign(ids_, ids_ + n); isOcean.assign(oc, oc + n); pole.assign(pole_, pole_ + 3 * (size_t)n); omega.assign(om, om + n); density.assign(de, de + n);
        int mx = 0;
        for (int k = 0; k < n; k++) { if (ids[k] < 0) throw std::invalid_argument("negative plate id"); mx = std::max(mx, ids[k]); }
        index.assign((size_t)mx + 1, -1);
        for (int k = 0; k < n; k++) index[ids[k]] = k;
    }
    int find(int id) const { return (id >= 0 && id < (int)index.size()) ? index[id] : -1; }
    bool ocean(int id) const { const int k = find(id); return k >= 0 && isOcean[k]; }
};
struct PlateTableDev {
    DevBuf<int> index; DevBuf<uint8_t> isOcean; DevBuf<double> pole, omega, density;
    PlateTab upload(const PlateTableHost& h, cudaStream_t s) {
        dev_copy(index.ensure(h.index.size()), h.index.data(), sizeof(int) * h.index.size(), 0, s);
        dev_copy(isOcean.ensure(h.n), h.isOcean.data(), (size_t)h.n, 0, s);
        dev_copy(pole.ensure(3 * (size_t)h.n), h.pole.data(), sizeof(double) * 3 * (size_t)h.n, 0, s);
        dev_copy(omega.ensure(h.n), h.omega.data(), sizeof(double) * (size_t)h.n, 0, s);
        dev_copy(density.ensure(h.n), h.density.data(), sizeof(double) * (size_t)h.n, 0, s);
        return PlateTab{index.p, (int)h.index.size(), isOcean.p, pole.p, omega.p, density.p};
    }
};

// host copies of one findCollisions result
struct CollisionHost { std::vector<float> stress, subduct; std::vector<int8_t> btype; std::vector<uint8_t> bothOcean, hasOcean, setCode; };

// insertion-ordered Set of cells
struct OrderedCells {
    std::vector<int> items; std::vector<uint8_t> in;
    void reset(int N) { items.clear(); in.assign(N, 0); }
    void add(int r) { if (!in[r]) { in[r] = 1; items.push_back(r); } }
};

struct ElevationOutputs {   // device pointers (engine scratch or caller arrays in device mode)
    float* elev; float* stress; uint8_t* mountain; uint8_t* coastline; uint8_t* ocean; ElevDebug dbg;
};

struct Elevation {
    Mesh* m; int N;
    PlateTableDev dP, dSP;
    // device buffers
    DevBuf<float> cs[2], cf[2], dist[5], dBdry, cStress, cSub, rift, ridge, fracture, backArc, backArcS, arcD, arcS, stressD, subD, dbg[12], elevD;
    DevBuf<int8_t> cb[2], btypeD;
    DevBuf<uint8_t> cbo[2], cho[2], ccode[2], isOceanD, cConvD, noiseTabs, setM, setC, setO;
    DevBuf<int> platesD[2], maxBits;
    DevBuf<DomeDev> domesD;
    // staging for host-pointer mode
    DevBuf<float> sElevOut, sStressOut, sDbg[12];
    DevBuf<int> sPlate, sSuper;
    DevBuf<uint8_t> sM, sC, sO;

    explicit Elevation(Mesh* mesh) : m(mesh), N(mesh->N) {}
    const Exec& ex() const { return m->ex(); }

    static double jsr(double x) { return floor(x + 0.5); }

    void run_collisions(int layer, const PlateTab& P, const int* r_plate_dev, const Simplex& noise, CollisionHost& out) {
        const Exec& x = ex();
        const double dt = 1e-2 / std::max(1.0, sqrt(N / 10000.0));
        CollisionOut o{cs[layer].ensure(N), cf[layer].ensure(N), cb[layer].ensure(N), cbo[layer].ensure(N), cho[layer].ensure(N), ccode[layer].ensure(N)};
        x.for_each(N, CollisionsK{m->csr(), m->xyz.p, P, r_plate_dev, noise, dt, N > 200000 ? 2 : 3, o});
        out.stress.resize(N); out.subduct.resize(N); out.btype.resize(N); out.bothOcean.resize(N); out.hasOcean.resize(N); out.setCode.resize(N);
        dev_copy(out.stress.data(), o.stress, sizeof(float) * (size_t)N, 1, x.stream);
        dev_copy(out.subduct.data(), o.subduct, sizeof(float) * (size_t)N, 1, x.stream);
        dev_copy(out.btype.data(), o.btype, (size_t)N, 1, x.stream);
        dev_copy(out.bothOcean.data(), o.bothOcean, (size_t)N, 1, x.stream);
        dev_copy(out.hasOcean.data(), o.hasOcean, (size_t)N, 1, x.stream);
        dev_copy(out.setCode.data(), o.setCode, (size_t)N, 1, x.stream);
    }

#if defined(__GNUC__)
#define PB_PREFETCH(p) __builtin_prefetch((p), 0, 1)
#else
#define PB_PREFETCH(p) ((void)0)
#endif
    // ---- host-serial pieces ------------------------------------------------------------------------------------
    static void propagate_stress(const int* off, const int* adj, int N, std::vector<float>& stress, std::vector<float>& sub,
                                 const int* plate, const PlateTableHost& P, double decay, double subDecay, int numPasses) {   // :127-159
        const bool dbgT = getenv("PB_DEBUG") != nullptr;
        const auto tStart = std::chrono::steady_clock::now();
        std::vector<int> frontier, next;
        size_t visits = 0;
        for (int r = 0; r < N; r++) if (stress[r] > 0.01f && (double)stress[r] > 0.01) frontier.push_back(r);
        for (int pass = 0; pass < numPasses && !frontier.empty(); pass++) {
            next.clear();
            for (size_t fi = 0; fi < frontier.size(); fi++) {
                const int r = frontier[fi];
                const int pl = plate[r];
                if (P.ocean(pl)) continue;
                const float sf = sub[r];
                const double propagated = (double)stress[r] * ((double)sf > 0.5 ? subDecay : decay);
                if (propagated < 0.005) continue;
                for (int j = off[r], e = off[r + 1]; j < e; j++) {
                    const int nb = adj[j];
                    if (plate[nb] == pl && propagated > (double)stress[nb]) {
                        stress[nb] = (float)propagated; sub[nb] = sf; next.push_back(nb);
                        PB_PREFETCH(adj + off[nb]);          // read in the next pass
                    }
                }
            }
            visits += frontier.size();
            frontier.swap(next);
        }
        if (dbgT) fprintf(stderr, "[pb] propagate_stress: %zu frontier visits, %.2f ms\n", visits,
                          std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tStart).count());
    }
    struct ParkMillerInt {   // makeRandInt (js/rng.js:8-11); the state is an integer < 2^31, so the JS double
        unsigned long long s;  // arithmetic (s*16807) % 2147483647 is reproduced exactly in 64-bit integers
        explicit ParkMillerInt(double seed) { s = (unsigned long long)(fmod(fabs(floor(seed * 9301.0 + 49297.0)), 2147483646.0) + 1.0); }
        long long operator()(double n) {
            // x mod (2^31 - 1) without a division: fold the high bits twice (x < 2^46), then one conditional subtract
            unsigned long long x = s * 16807ull;
            x = (x & 2147483647ull) + (x >> 31);
            x = (x & 2147483647ull) + (x >> 31);
            s = x >= 2147483647ull ? x - 2147483647ull : x;
            return (long long)floor(((double)(s - 1) / 2147483646.0) * n);
        }
    };
    static void distance_field(const int* off, const int* adj, int N, const std::vector<int>& seeds, const uint8_t* isStop, double seed,
                               std::vector<float>& dist) {   // :164-189
        const bool dbgT = getenv("PB_DEBUG") != nullptr;
        const auto tStart = std::chrono::steady_clock::now();
        ParkMillerInt randInt(seed);
        dist.assign(N, INFINITY);
        std::vector<int> queue;
        queue.reserve(N);
        for (int r : seeds) { queue.push_back(r); dist[r] = 0; }
        for (size_t qi = 0; qi < queue.size(); qi++) {
            const size_t pos = qi + (size_t)randInt((double)(queue.size() - qi));
            const int cur = queue[pos];
            queue[pos] = queue[qi];
            const float dn = (float)((double)dist[cur] + 1);
            for (int j = off[cur], e = off[cur + 1]; j < e; j++) {
                const int nb = adj[j];
                if (dist[nb] == INFINITY && !(isStop && isStop[nb])) {
                    dist[nb] = dn; queue.push_back(nb);
                    PB_PREFETCH(adj + off[nb]);      // the row is read when nb is drawn, typically thousands of steps later
                }
            }
        }
        if (dbgT) fprintf(stderr, "[pb] distance_field(seed %.0f): %zu seeds, %zu cells, %.2f ms\n", seed, seeds.size(), queue.size(),
                          std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tStart).count());
    }

    // ---- assignElevation ----------------------------------------------------------------------------------------------
    // r_plate / r_superPlate: device pointers + host copies.  SP == nullptr → no superPlateData.
    void assign(const PlateTableHost& P, const int* r_plate_dev, const int* r_plate, const std::vector<int>& plateSeeds,
                double noiseSeed, double noiseMag, double seed, double spread, const PlateTableHost* SP, const int* r_super_dev,
                const int* r_super, const ElevationOutputs& out) {
        const Exec& x = ex();
        const int* off = m->hOffCopy.data(); const int* adj = m->hAdjCopy.data();
        const bool dbgT = getenv("PB_DEBUG") != nullptr;
        auto tNow = [] { return std::chrono::steady_clock::now(); };
        auto t0 = tNow();
        auto lap = [&](const char* what) {
            if (!dbgT) return;
            auto t1 = tNow();
            fprintf(stderr, "[pb] elevation %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
            t0 = t1;
        };
        // noise tables: main, rift(+419), fold(+557), coast(+77,+133,+211), arc(+307), hotspot(+501,+502,+503)
        const double seeds10[10] = {noiseSeed, seed + 419, seed + 557, seed + 77, seed + 133, seed + 211, seed + 307, seed + 501, seed + 502, seed + 503};
        noiseTabs.ensure(10 * 1024);
        {
            std::vector<uint8_t> all(10 * 1024);
            for (int k = 0; k < 10; k++) { SimplexTable t(seeds10[k]); memcpy(all.data() + 1024 * k, t.t, 1024); }
            dev_copy(noiseTabs.p, all.data(), all.size(), 0, x.stream);
            stream_sync(x.stream);
        }
        auto NZ = [&](int k) { return Simplex{noiseTabs.p + 1024 * k}; };

        // 1. collisions (device) → host
        const PlateTab tP = dP.upload(P, x.stream);
        CollisionHost small, super;
        run_collisions(0, tP, r_plate_dev, NZ(0), small);
        const bool dual = SP != nullptr;
        if (dual) { const PlateTab tS = dSP.upload(*SP, x.stream); run_collisions(1, tS, r_super_dev, NZ(0), super); }

        // 5a. the two coast-distance fills (:411, 426) read only r_plate and the ocean flags of the plates: their seed lists are
        // built while the collision kernels run, and the fills — the longest serial chains of the call — start before anything else
        struct Joiner { std::thread* t; int n; ~Joiner() { for (int k = 0; k < n; k++) if (t[k].joinable()) t[k].join(); } };
        std::vector<uint8_t> isOcean(N);
        for (int r = 0; r < N; r++) isOcean[r] = P.ocean(r_plate[r]) ? 1 : 0;
        OrderedCells coastSeeds; coastSeeds.reset(N);
        std::vector<int> landCoastSeeds;
        for (int r = 0; r < N; r++) {
            if (isOcean[r]) continue;
            for (int j = off[r], e = off[r + 1]; j < e; j++) if (isOcean[adj[j]]) { coastSeeds.add(adj[j]); landCoastSeeds.push_back(r); break; }
        }
        std::vector<float> hd[5];
        std::thread fill[5];
        fill[3] = std::thread([&] { distance_field(off, adj, N, coastSeeds.items, nullptr, seed + 4, hd[3]); });
        fill[4] = std::thread([&] { distance_field(off, adj, N, landCoastSeeds, isOcean.data(), seed + 5, hd[4]); });
        Joiner joinFills{fill, 5};
        lap("ocean mask + coast seeds");

        stream_sync(x.stream);
        lap("collisions (device) + d2h");

        // stress propagation (:329-362) only needs the raw collision outputs, so it starts now on its own threads and
        // overlaps the set unions / blends / seed lists below
        const double scaleFactor = sqrt(N / 10000.0);
        const double baseDecay = 0.5 + spread * 0.04;
        const double decayFactor = pb_pow(baseDecay, 1 / scaleFactor);
        const double subductDecayFactor = pb_pow(baseDecay * 0.45, 1 / scaleFactor);
        const int numPasses = (int)std::max(1.0, jsr(spread * 3 * scaleFactor));
        std::vector<float> sSt(small.stress), sSu(small.subduct), pSt, pSu;
        if (dual) { pSt = super.stress; pSu = super.subduct; }
        std::thread prop[2];
        prop[0] = std::thread([&] { propagate_stress(off, adj, N, sSt, sSu, r_plate, P, decayFactor, subductDecayFactor, numPasses); });
        if (dual) prop[1] = std::thread([&] { propagate_stress(off, adj, N, pSt, pSu, r_super, *SP, decayFactor, subductDecayFactor, numPasses); });
        Joiner joinProp{prop, 2};

        // 2. blend (:250-327)
        OrderedCells mountain, coastline, ocean;
        mountain.reset(N); coastline.reset(N); ocean.reset(N);
        std::vector<float> stress, sub;
        std::vector<int8_t> btype;
        std::vector<uint8_t> bothOcean, hasOcean;
        const double SMALL_W = 0.05, SUPER_W = 0.95;
        if (!dual) {
            for (int r = 0; r < N; r++) { const uint8_t c = small.setCode[r]; if (c == 1) mountain.add(r); else if (c == 2) coastline.add(r); else if (c == 3) ocean.add(r); }
            stress = small.stress; sub = small.subduct; btype = small.btype; bothOcean = small.bothOcean; hasOcean = small.hasOcean;
        } else {
            for (int r = 0; r < N; r++) if (super.setCode[r] == 1) mountain.add(r);
            for (int r = 0; r < N; r++) if (small.setCode[r] == 1) mountain.add(r);
            for (int r = 0; r < N; r++) if (super.setCode[r] == 3) ocean.add(r);
            for (int r = 0; r < N; r++) if (small.setCode[r] == 3) ocean.add(r);
            for (int r = 0; r < N; r++) if (super.setCode[r] == 2 && !mountain.in[r]) coastline.add(r);
            for (int r = 0; r < N; r++) if (small.setCode[r] == 2 && !mountain.in[r]) coastline.add(r);
            stress.resize(N); sub.resize(N); btype.resize(N); bothOcean.resize(N); hasOcean.resize(N);
            double maxSuperStress = 0;
            for (int r = 0; r < N; r++) if ((double)super.stress[r] > maxSuperStress) maxSuperStress = super.stress[r];
            const double invMax = maxSuperStress > 1e-6 ? 1 / maxSuperStress : 0;
            for (int r = 0; r < N; r++) {
                const double sS = small.stress[r], sP = super.stress[r];
                double proximity = sP * invMax * 3; if (proximity > 1) proximity = 1;
                const double effectiveSmallW = SMALL_W * (SMALL_W + (1 - SMALL_W) * proximity);
                stress[r] = (float)(effectiveSmallW * sS + SUPER_W * sP);
                const double wS = SMALL_W * sS, wP = SUPER_W * sP, total = wS + wP;
                if (total > 1e-6) sub[r] = (float)((wS * (double)small.subduct[r] + wP * (double)super.subduct[r]) / total);
                else sub[r] = (float)(SMALL_W * (double)small.subduct[r] + SUPER_W * (double)super.subduct[r]);
                btype[r] = wS > wP ? small.btype[r] : super.btype[r];
                bothOcean[r] = small.bothOcean[r] | super.bothOcean[r];
                hasOcean[r] = small.hasOcean[r] | super.hasOcean[r];
            }
        }

        lap("sets + blend");
        // 4a. plate representatives (need only the sets), ocean mask, coast seed lists — still overlapping the propagation
        {
            std::vector<int> plateRep(P.n, -1);          // first unclaimed cell of every plate, by table row
            int missing = P.n;
            for (int r = 0; r < N && missing > 0; r++) {
                if (mountain.in[r] || coastline.in[r] || ocean.in[r]) continue;
                const int k = P.find(r_plate[r]);
                if (k >= 0 && plateRep[k] < 0) { plateRep[k] = r; missing--; }
            }
            for (int pid : plateSeeds) {
                const int k = P.find(pid);
                if (k >= 0 && plateRep[k] >= 0) (P.ocean(pid) ? ocean : coastline).add(plateRep[k]);
            }
        }
        lap("representatives");
        // 5b. the ocean-distance fill (:393) needs the finished sets but not the propagation: it overlaps the rest of it
        fill[1] = std::thread([&] { distance_field(off, adj, N, ocean.items, coastline.in.data(), seed + 2, hd[1]); });
        // 3. join the propagation, blend its results (:343-361)
        prop[0].join();
        if (dual) prop[1].join();
        if (!dual) { stress = sSt; sub = sSu; }
        else {
            for (int r = 0; r < N; r++) {
                stress[r] = (float)(SMALL_W * (double)sSt[r] + SUPER_W * (double)pSt[r]);
                const double wS = SMALL_W * (double)sSt[r], wP = SUPER_W * (double)pSt[r], total = wS + wP;
                if (total > 1e-6) sub[r] = (float)((wS * (double)sSu[r] + wP * (double)pSu[r]) / total);
            }
        }

        lap("propagateStress");
        std::vector<int> stressMountain;
        std::vector<uint8_t> stop(N, 0);
        for (int r : mountain.items) if ((double)sub[r] < 0.55) { stressMountain.push_back(r); stop[r] = 1; }
        for (int r : coastline.items) stop[r] = 1;
        for (int r : ocean.items) stop[r] = 1;
        // 5c. the two fills that depend on the propagation (:392, 394); all five run concurrently with the capped BFS below
        fill[0] = std::thread([&] { distance_field(off, adj, N, stressMountain, ocean.in.data(), seed + 1, hd[0]); });
        fill[2] = std::thread([&] { distance_field(off, adj, N, coastline.items, stop.data(), seed + 3, hd[2]); });

        // 6. maxStress = p97 of the non-trivial stresses (:443-453)
        double maxStress = 0;
        {
            std::vector<float> vals;
            for (int r = 0; r < N; r++) { if ((double)stress[r] > 0.01) vals.push_back(stress[r]); if ((double)stress[r] > maxStress) maxStress = stress[r]; }
            if (!vals.empty()) {
                const size_t k = std::min(vals.size() - 1, (size_t)floor((double)vals.size() * 0.97));
                std::nth_element(vals.begin(), vals.begin() + k, vals.end());
                maxStress = vals[k];
            }
            if (maxStress < 0.01) maxStress = 1;
        }

        // 7. capped FIFO BFS (:464-631, :1059-1086)
        const double maxCD = std::max(8.0, jsr(8 * scaleFactor));
        std::vector<float> hBdry(N, (float)(maxCD + 1)), hCS(N, 0.f), hCSub(N, 0.f);
        std::vector<uint8_t> hConv(N, 0);
        std::thread coastBfs([&] {
            std::vector<int> q;
            for (int r = 0; r < N; r++) {
                const uint8_t rOc = isOcean[r];
                for (int j = off[r], e = off[r + 1]; j < e; j++) if (isOcean[adj[j]] != rOc) { q.push_back(r); break; }
            }
            for (int r : q) {
                hBdry[r] = 0;
                double v = (double)stress[r] / maxStress; if (v > 1) v = 1;
                hCS[r] = (float)v; hCSub[r] = sub[r]; hConv[r] = btype[r] == 1 ? 1 : 0;
            }
            for (size_t qi = 0; qi < q.size();) {
                const int r = q[qi++];
                const double nd = (double)hBdry[r] + 1;
                if (nd > maxCD) continue;
                for (int j = off[r], e = off[r + 1]; j < e; j++) {
                    const int nr = adj[j];
                    if (nd < (double)hBdry[nr]) { hBdry[nr] = (float)nd; hCS[nr] = hCS[r]; hCSub[nr] = hCSub[r]; hConv[nr] = hConv[r]; q.push_back(nr); }
                    else if (nd == (double)hBdry[nr] && hCS[r] > hCS[nr]) { hCS[nr] = hCS[r]; hCSub[nr] = hCSub[r]; hConv[nr] = hConv[r]; }
                }
            }
        });
        Joiner joinCoast{&coastBfs, 1};
        // generic capped BFS: pass(r, nr) decides whether nr may be entered from r; payload copied from the discoverer
        auto capped = [&](std::vector<float>& d, std::vector<float>* payload, std::vector<int>& q, double cap, int mode) {
            for (size_t qi = 0; qi < q.size();) {
                const int r = q[qi++];
                const double nd = (double)d[r] + 1;
                if (nd > cap) continue;
                const int pl = r_plate[r];
                for (int j = off[r], e = off[r + 1]; j < e; j++) {
                    const int nr = adj[j];
                    bool ok;
                    if (mode == 0) ok = r_plate[nr] == pl && !isOcean[nr];        // rift
                    else if (mode == 1) ok = isOcean[nr] != 0;                     // ridge, fracture
                    else if (mode == 2) ok = r_plate[nr] == pl;                    // back-arc
                    else ok = r_plate[nr] == pl && isOcean[nr];                    // island arc
                    if (nd < (double)d[nr] && ok) { d[nr] = (float)nd; if (payload) (*payload)[nr] = (*payload)[r]; q.push_back(nr); }
                }
            }
        };
        const double riftHalfWidth = std::max(2.0, jsr(4 * scaleFactor)), ridgeHalfWidth = riftHalfWidth;
        const double fractureHalfWidth = std::max(2.0, jsr(3 * scaleFactor));
        const double baStart = std::max(1.0, jsr(2 * scaleFactor)), baPeak = std::max(2.0, jsr(3 * scaleFactor)), baEnd = std::max(3.0, jsr(5 * scaleFactor));
        const double maxArcDist = std::max(5.0, jsr(5 * scaleFactor));
        std::vector<float> hRift(N, INFINITY), hRidge(N, INFINITY), hFrac(N, INFINITY), hBA(N, INFINITY), hBAS(N, 0.f), hArc(N, (float)(maxArcDist + 1)), hArcS(N, 0.f);
        {
            auto norm = [&](int r) { double v = (double)stress[r] / maxStress; return (float)(v > 1 ? 1 : v); };
            std::thread t1([&] { std::vector<int> q; for (int r = 0; r < N; r++) if (btype[r] == 2 && !hasOcean[r]) { q.push_back(r); hRift[r] = 0; } capped(hRift, nullptr, q, riftHalfWidth, 0); });
            std::thread t2([&] { std::vector<int> q; for (int r = 0; r < N; r++) if (btype[r] == 2 && bothOcean[r]) { q.push_back(r); hRidge[r] = 0; } capped(hRidge, nullptr, q, ridgeHalfWidth, 1); });
            std::thread t3([&] { std::vector<int> q; for (int r = 0; r < N; r++) if (btype[r] == 3 && bothOcean[r]) { q.push_back(r); hFrac[r] = 0; } capped(hFrac, nullptr, q, fractureHalfWidth, 1); });
            std::thread t4([&] { std::vector<int> q; for (int r = 0; r < N; r++) if (btype[r] == 1 && hasOcean[r] && (double)sub[r] < 0.50) { q.push_back(r); hBA[r] = 0; hBAS[r] = norm(r); } capped(hBA, &hBAS, q, baEnd, 2); });
            { std::vector<int> q; for (int r = 0; r < N; r++) if (btype[r] == 1 && bothOcean[r] && (double)sub[r] < 0.45) { q.push_back(r); hArc[r] = 0; hArcS[r] = norm(r); } capped(hArc, &hArcS, q, maxArcDist, 3); }
            t1.join(); t2.join(); t3.join(); t4.join();
        }
        coastBfs.join();
        for (auto& t : fill) t.join();

        lap("fills + p97 + capped BFS (threads)");
        // 8. hotspot domes (:1148-1262)
        std::vector<DomeDev> domes;
        build_domes(P, r_plate, seed, domes);

        lap("hotspot domes");
        // 9. upload, device synthesis
        auto up = [&](DevBuf<float>& b, const std::vector<float>& h) { dev_copy(b.ensure(N), h.data(), sizeof(float) * (size_t)N, 0, x.stream); return b.p; };
        auto up8 = [&](DevBuf<uint8_t>& b, const std::vector<uint8_t>& h) { dev_copy(b.ensure(N), h.data(), (size_t)N, 0, x.stream); return b.p; };
        ElevFields F;
        F.stress = up(stressD, stress); F.subduct = up(subD, sub);
        dev_copy(btypeD.ensure(N), btype.data(), (size_t)N, 0, x.stream); F.btype = btypeD.p;
        F.isOcean = up8(isOceanD, isOcean);
        F.dist_mountain = up(dist[0], hd[0]); F.dist_ocean = up(dist[1], hd[1]); F.dist_coastline = up(dist[2], hd[2]);
        F.dist_coast = up(dist[3], hd[3]); F.dist_coast_land = up(dist[4], hd[4]);
        F.riftDist = up(rift, hRift); F.ridgeDist = up(ridge, hRidge); F.fractureDist = up(fracture, hFrac);
        F.backArcDist = up(backArc, hBA); F.backArcStress = up(backArcS, hBAS);
        F.coastConvergent = up8(cConvD, hConv);
        up(dBdry, hBdry); up(cStress, hCS); up(cSub, hCSub); up(arcD, hArc); up(arcS, hArcS);
        dev_copy(domesD.ensure(PB_MAX_DOMES), domes.data(), sizeof(DomeDev) * domes.size(), 0, x.stream);
        dev_copy(out.stress, stress.data(), sizeof(float) * (size_t)N, 0, x.stream);
        if (out.mountain) dev_copy(out.mountain, mountain.in.data(), (size_t)N, 0, x.stream);
        if (out.coastline) dev_copy(out.coastline, coastline.in.data(), (size_t)N, 0, x.stream);
        if (out.ocean) dev_copy(out.ocean, ocean.in.data(), (size_t)N, 0, x.stream);

        ElevParams p;
        p.maxStress = maxStress; p.noiseMag = noiseMag; p.scaleFactor = scaleFactor;
        p.interiorBand = std::max(4.0, jsr(16 * scaleFactor)); p.tectonicReach = std::max(6.0, jsr(20 * scaleFactor));
        p.plateauStart = std::max(2.0, jsr(3 * scaleFactor)); p.riftHalfWidth = riftHalfWidth; p.ridgeHalfWidth = ridgeHalfWidth;
        p.fractureHalfWidth = fractureHalfWidth; p.baStart = baStart; p.baPeak = baPeak; p.baEnd = baEnd;
        p.warpOctaves = N > 200000 ? 2 : 3;
        x.for_each(N, ElevationMainK{m->xyz.p, r_plate_dev, tP, F, p, NZ(0), NZ(1), NZ(2), out.elev, out.dbg});
        x.for_each(N, CoastalRoughenK{m->xyz.p, dBdry.p, cStress.p, cSub.p, cConvD.p, stressD.p, isOceanD.p, maxStress, noiseMag,
                                      std::max(8.0, jsr(8 * scaleFactor)), std::max(4.0, jsr(4 * scaleFactor)), NZ(0), NZ(3), NZ(4), NZ(5),
                                      out.elev, out.dbg.coastal});
        x.for_each(N, IslandArcK{m->xyz.p, arcD.p, arcS.p, maxArcDist, scaleFactor, NZ(6), out.elev, out.dbg.coastal});
        x.for_each(N, HotspotK{m->xyz.p, domesD.p, (int)domes.size(), NZ(7), NZ(8), out.elev, out.dbg.hotspot});
        x.for_each(N, CompressPeaksK{out.elev});
        stream_sync(x.stream);      // host vectors above are the sources of the async uploads
        lap("h2d + synthesis kernels");
    }

    // hotspot dome list :1130-1262 (host: ≤ 5 plumes × chain, Park–Miller driven)
    void build_domes(const PlateTableHost& P, const int* r_plate, double seed, std::vector<DomeDev>& domes) {
        const int NUM_HOTSPOTS = 5, CHAIN_LENGTH = 6;
        const double CHAIN_DECAY = 0.75, CHAIN_SPACING = 0.06, DOME_SIGMA = 0.006, DOME_STRENGTH = 0.60, SWELL_SIGMA_MULT = 2, SWELL_STR_MULT = 0.10;
        const float* xyz = m->hXyzCopy.data();
        ParkMiller hsRng(seed + 999);
        ParkMillerInt hsRandInt(seed + 1001);
        SimplexTable t3(seed + 503);
        const Simplex hsNoise3{t3.t};
        struct Raw { double x, y, z, strength, baseStrength, sigma; int chainIndex, chainLength; double ux, uy, uz, vx, vy, vz, base; };
        std::vector<Raw> raw;
        auto frame = [](double px, double py, double pz, double dx, double dy, double dz, Raw& o) {
            const double dd = dx * px + dy * py + dz * pz;
            double ux = dx - dd * px, uy = dy - dd * py, uz = dz - dd * pz;
            double uLen = sqrt(ux * ux + uy * uy + uz * uz); if (uLen == 0 || uLen != uLen) uLen = 1;
            ux /= uLen; uy /= uLen; uz /= uLen;
            o.ux = ux; o.uy = uy; o.uz = uz; o.vx = py * uz - pz * uy; o.vy = pz * ux - px * uz; o.vz = px * uy - py * ux;
        };
        for (int h = 0; h < NUM_HOTSPOTS; h++) {
            const double hStrength = DOME_STRENGTH * (0.4 + hsRng.next() * 1.2);
            const double hSigma = DOME_SIGMA * (0.4 + hsRng.next() * 1.2);
            const double hDecay = CHAIN_DECAY + (hsRng.next() - 0.5) * 0.35;
            const int hLength = (int)std::max(3.0, CHAIN_LENGTH + jsr((hsRng.next() - 0.5) * 10));
            const int centerR = (int)hsRandInt((double)N);
            const double hx = xyz[3 * centerR], hy = xyz[3 * centerR + 1], hz = xyz[3 * centerR + 2];
            const int plate = r_plate[centerR];
            const int pk = P.find(plate);
            if (pk < 0) continue;
            const double px = P.pole[3 * pk], py = P.pole[3 * pk + 1], pz = P.pole[3 * pk + 2], om = P.omega[pk];
            double drift[3] = {om * (py * hz - pz * hy), om * (pz * hx - px * hz), om * (px * hy - py * hx)};
            const double driftLen = sqrt(drift[0] * drift[0] + drift[1] * drift[1] + drift[2] * drift[2]);
            if (driftLen < 1e-6) continue;
            drift[0] /= driftLen; drift[1] /= driftLen; drift[2] /= driftLen;
            const double oceanBoost = P.ocean(plate) ? 1.8 : 1.0;
            const double baseRiftAngle = hsNoise3.noise3D(hx * 10, hy * 10, hz * 10) * PB_PI;
            Raw d0{};
            d0.x = hx; d0.y = hy; d0.z = hz; d0.strength = hStrength * oceanBoost; d0.baseStrength = hStrength; d0.sigma = hSigma;
            d0.chainIndex = 0; d0.chainLength = hLength; d0.base = baseRiftAngle;
            frame(hx, hy, hz, drift[0], drift[1], drift[2], d0);
            raw.push_back(d0);
            double perpX = drift[1] * hz - drift[2] * hy, perpY = drift[2] * hx - drift[0] * hz, perpZ = drift[0] * hy - drift[1] * hx;
            double perpLen = sqrt(perpX * perpX + perpY * perpY + perpZ * perpZ); if (perpLen == 0 || perpLen != perpLen) perpLen = 1;
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
                const double tLen = sqrt(tx * tx + ty * ty + tz * tz);
                if (tLen < 1e-6) break;
                tx /= tLen; ty /= tLen; tz /= tLen;
                double sinA, cosA;
                pb_sincos(stepSpacing, &sinA, &cosA);
                cx = cx * cosA + tx * sinA; cy = cy * cosA + ty * sinA; cz = cz * cosA + tz * sinA;
                const double nL = sqrt(cx * cx + cy * cy + cz * cz);
                cx /= nL; cy /= nL; cz /= nL;
                Raw dc{};
                dc.x = cx; dc.y = cy; dc.z = cz; dc.strength = str; dc.baseStrength = baseStr; dc.sigma = stepSigma;
                dc.chainIndex = ci; dc.chainLength = hLength; dc.base = baseRiftAngle;
                frame(cx, cy, cz, drift[0], drift[1], drift[2], dc);
                raw.push_back(dc);
            }
        }
        if ((int)raw.size() > PB_MAX_DOMES) throw Error("too many hotspot domes");
        domes.clear();
        for (const Raw& q : raw) {
            DomeDev d{};
            d.x = q.x; d.y = q.y; d.z = q.z; d.strength = q.strength; d.ux = q.ux; d.uy = q.uy; d.uz = q.uz; d.vx = q.vx; d.vy = q.vy; d.vz = q.vz;
            d.cosThreshPeak = pb_cos(q.sigma * 5.5);
            d.invS2 = -0.5 / (q.sigma * q.sigma);
            const double swSigma = q.sigma * SWELL_SIGMA_MULT;
            d.swellStrength = q.baseStrength * SWELL_STR_MULT;
            d.cosThreshSwell = pb_cos(swSigma * 3);
            d.invS2Swell = -0.5 / (swSigma * swSigma);
            d.driftStretch = 1.0 / 1.4;
            d.hasCaldera = (q.chainIndex <= 1 && q.strength > 0.15) ? 1 : 0;
            const double calderaSigma = q.sigma * 0.25;
            d.calderaDepth = q.strength * 0.20;
            d.invS2Caldera = -0.5 / (calderaSigma * calderaSigma);
            d.ageFactor = q.chainLength > 0 ? (double)q.chainIndex / q.chainLength : 0;
            if (q.chainIndex == 0) { d.nRift = 3; d.riftAngles[0] = q.base; d.riftAngles[1] = q.base + PB_PI * 0.6; d.riftAngles[2] = q.base - PB_PI * 0.6; }
            else if (q.chainIndex == 1) { d.nRift = 2; d.riftAngles[0] = q.base; d.riftAngles[1] = q.base + PB_PI; }
            else if (q.chainIndex <= (int)floor(q.chainLength * 0.4)) { d.nRift = 1; d.riftAngles[0] = q.base; }
            else d.nRift = 0;
            domes.push_back(d);
        }
    }
};

}  // namespace pb
