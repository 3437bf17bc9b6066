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
#include <memory>
#include <unordered_map>
#include "pb_engine.h"
#include "pb_elevation.h"
#include "pb_elevation_mid.h"

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
    DevBuf<uint8_t> cbo[2], cho[2], ccode[2], isOceanD, cConvD, noiseTabs, setM, setC, setO, bothD, hasD, flagU[8];
    DevBuf<int> platesD[2], maxBits, scratchI[3], listD[8], counts, obFront[OBFS_COUNT][2], obClaim[OBFS_COUNT], obCount, obTotals;
    DevBuf<unsigned long long> obBest;
    DevBuf<uint32_t> keys;
    DevBuf<double> maxStressD;
    DevBuf<DomeDev> domesD;
    // pinned host arrays of the host-serial stages (propagation layers down and up, five distance fields up)
    PinnedBuf<float> hStress[2], hSub[2], hDist[5];
    PinnedBuf<uint8_t> hTabs, hIsOcean;
    PinnedBuf<int> hRepBuf;
    PinnedBuf<DomeDev> hDomeBuf;
    int obGrid = -1;
    // staging for host-pointer mode
    DevBuf<float> sElevOut, sStressOut, sDbg[12];
    DevBuf<int> sPlate, sSuper;
    DevBuf<uint8_t> sM, sC, sO;

    explicit Elevation(Mesh* mesh) : m(mesh), N(mesh->N) {}
    const Exec& ex() const { return m->ex(); }

    static double jsr(double x) { return floor(x + 0.5); }

    void run_collisions(int layer, const PlateTab& P, const int* r_plate_dev, const Simplex& noise) {
        const Exec& x = ex();
        const double dt = 1e-2 / std::max(1.0, sqrt(N / 10000.0));
        CollisionOut o{cs[layer].ensure(N), cf[layer].ensure(N), cb[layer].ensure(N), cbo[layer].ensure(N), cho[layer].ensure(N), ccode[layer].ensure(N)};
        x.for_each(N, CollisionsK{m->csr(), m->xyz.p, P, r_plate_dev, noise, dt, N > 200000 ? 2 : 3, o});
    }

#if defined(__GNUC__)
#define PB_PREFETCH(p) __builtin_prefetch((p), 0, 1)
#else
#define PB_PREFETCH(p) ((void)0)
#endif
    // ---- host-serial pieces ------------------------------------------------------------------------------------
    // propagateStress :127-159 for the cells of ONE plate.  Within a pass the frontier is processed in list order and
    // every update is visible to the items after it (A.7), so the pass is a serial chain — but a cell only ever reads or
    // writes cells of its own plate (`r_plate[nb] === plate` is tested before `r_stress[nb]` is read), so plates are
    // independent: the reference's single frontier list, filtered by plate, is processed plate by plate on worker
    // threads and gives the reference's values.  Cells of oceanic plates never propagate (:141) and are never written.
    static void propagate_plate(const int* off, const int* adj, float* stress, float* sub, const int* plate, int pl,
                                std::vector<int>& frontier, double decay, double subDecay, int numPasses, size_t* visits) {
        std::vector<int> next;
        for (int pass = 0; pass < numPasses && !frontier.empty(); pass++) {
            next.clear();
            for (size_t fi = 0; fi < frontier.size(); fi++) {
                const int r = frontier[fi];
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
            *visits += frontier.size();
            frontier.swap(next);
        }
    }
    // splits the initial frontier (ascending r, :133-135) by plate and runs the plates on up to `maxWorkers` threads
    void start_propagation(std::vector<std::thread>& threads, const int* off, const int* adj, float* stress, float* sub, const int* plate,
                           const PlateTableHost& P, double decay, double subDecay, int numPasses) {
        auto perPlate = std::make_shared<std::vector<std::vector<int>>>((size_t)P.n);
        for (int r = 0; r < N; r++) {
            if (!((double)stress[r] > 0.01)) continue;
            const int k = P.find(plate[r]);
            if (k >= 0 && !P.isOcean[k]) (*perPlate)[k].push_back(r);
        }
        std::vector<int> rows;
        for (int k = 0; k < P.n; k++) if (!(*perPlate)[k].empty()) rows.push_back(k);
        std::sort(rows.begin(), rows.end(), [&](int a, int b) { return (*perPlate)[a].size() > (*perPlate)[b].size(); });
        int workers = (int)std::min<size_t>(rows.size(), (size_t)host_workers());
        if (workers <= 0) return;
        std::vector<std::vector<int>> bins(workers);
        std::vector<size_t> load(workers, 0);
        for (int k : rows) {               // longest first onto the least loaded worker
            const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            bins[w].push_back(k); load[w] += (*perPlate)[k].size();
        }
        const bool dbgT = getenv("PB_DEBUG") != nullptr;
        const std::vector<int> ids = P.ids;
        for (int w = 0; w < workers; w++) {
            threads.emplace_back([=] {
                const auto tStart = std::chrono::steady_clock::now();
                size_t visits = 0;
                for (int k : bins[w]) propagate_plate(off, adj, stress, sub, plate, ids[k], (*perPlate)[k], decay, subDecay, numPasses, &visits);
                if (dbgT) fprintf(stderr, "[pb] propagate_stress worker %d: %zu plates, %zu frontier visits, %.2f ms\n", w, bins[w].size(), visits,
                                  std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tStart).count());
            });
        }
    }
    static int host_workers() {           // worker threads per propagation layer (PB_HOST_WORKERS overrides)
        if (const char* e = getenv("PB_HOST_WORKERS")) { const int v = atoi(e); if (v > 0) return v; }
        const unsigned hc = std::thread::hardware_concurrency();
        return (int)std::max(1u, std::min(6u, hc / 3));
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
        double fraction() {    // the rng() value of the next draw; randInt(n) = floor(fraction() * n)
            unsigned long long x = s * 16807ull;
            x = (x & 2147483647ull) + (x >> 31);
            x = (x & 2147483647ull) + (x >> 31);
            s = x >= 2147483647ull ? x - 2147483647ull : x;
            return (double)(s - 1) / 2147483646.0;
        }
    };
    // assignDistanceField :164-189 — class R: every step draws a uniformly random entry of the LIVE queue, so the fill is
    // one serial chain of |reached cells| steps, each a chain of dependent cache misses (queue[pos] → off[cur] → adj row
    // → dist[nb]); ≈ 65-80 ns per step on a host core.  (A software pipeline that prefetches for the draws of the next
    // steps was measured and does not pay: the drawn POSITION depends on the queue length after the previous step's
    // pushes, so a speculative position is off by one entry — a different cell — most of the time.)
    static void distance_field(const int* off, const int* adj, int N, const std::vector<int>& seeds, const uint8_t* isStop, double seed,
                               float* dist) {
        const bool dbgT = getenv("PB_DEBUG") != nullptr;
        const auto tStart = std::chrono::steady_clock::now();
        ParkMillerInt rng(seed);
        for (int r = 0; r < N; r++) dist[r] = INFINITY;
        std::vector<int> queueBuf((size_t)N + 1);
        int* q = queueBuf.data();
        size_t qn = 0;
        for (int r : seeds) { q[qn++] = r; dist[r] = 0; }
        for (size_t qi = 0; qi < qn; qi++) {
            const size_t pos = qi + (size_t)(rng.fraction() * (double)(qn - qi));   // qi + randInt(queue.length - qi); the product is >= 0: truncation = floor
            const int cur = q[pos];
            q[pos] = q[qi];
            const float dn = (float)((double)dist[cur] + 1);
            for (int j = off[cur], e = off[cur + 1]; j < e; j++) {
                const int nb = adj[j];
                if (dist[nb] == INFINITY && !(isStop && isStop[nb])) {
                    dist[nb] = dn; q[qn++] = nb;
                    PB_PREFETCH(adj + off[nb]);      // the row is read when nb is drawn, typically thousands of steps later
                }
            }
        }
        if (dbgT) fprintf(stderr, "[pb] distance_field(seed %.0f): %zu seeds, %zu cells, %.2f ms\n", seed, seeds.size(), qn,
                          std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tStart).count());
    }

    // ---- assignElevation ----------------------------------------------------------------------------------------------
    // r_plate / r_superPlate: device pointers + host copies.  SP == nullptr → no superPlateData.
    //
    // Timeline (1M cells, B200 + one host):   device                                    host
    //   collisions ×2, ocean mask, seed-set flags + ordered compactions, pre-blend  →   lists come down (pinned)
    //                                                                                    propagateStress per plate (threads), fills 4/5 start
    //                                                                                    representatives → fill 2 starts
    //   propagated layers go up, post-blend, p97, six capped BFS (one launch)       ←   fills 1/3 start as soon as the propagation is joined
    //   five distance fields go up, synthesis kernels                                ←   fills joined
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
        const size_t n = (size_t)N;
        // noise tables: main, rift(+419), fold(+557), coast(+77,+133,+211), arc(+307), hotspot(+501,+502,+503)
        const double seeds10[10] = {noiseSeed, seed + 419, seed + 557, seed + 77, seed + 133, seed + 211, seed + 307, seed + 501, seed + 502, seed + 503};
        noiseTabs.ensure(10 * 1024);
        {
            uint8_t* all = hTabs.ensure(10 * 1024);
            for (int k = 0; k < 10; k++) { SimplexTable t(seeds10[k]); memcpy(all + 1024 * k, t.t, 1024); }
            dev_copy(noiseTabs.p, all, 10 * 1024, 0, x.stream);
        }
        auto NZ = [&](int k) { return Simplex{noiseTabs.p + 1024 * k}; };
        const bool dual = SP != nullptr;
        const double scaleFactor = sqrt(N / 10000.0);

        // ---- 1. collisions (device); their raw stress / subduction factors come down for the propagation ----
        const PlateTab tP = dP.upload(P, x.stream);
        run_collisions(0, tP, r_plate_dev, NZ(0));
        if (dual) { const PlateTab tS = dSP.upload(*SP, x.stream); run_collisions(1, tS, r_super_dev, NZ(0)); }
        float* hS[2] = {hStress[0].ensure(n), dual ? hStress[1].ensure(n) : nullptr};
        float* hF[2] = {hSub[0].ensure(n), dual ? hSub[1].ensure(n) : nullptr};
        for (int l = 0; l < (dual ? 2 : 1); l++) {
            dev_copy(hS[l], cs[l].p, sizeof(float) * n, 1, x.stream);
            dev_copy(hF[l], cf[l].p, sizeof(float) * n, 1, x.stream);
        }

        // ---- 2. device: ocean mask, coast seed sets, seed sets of the layers, plate representatives, pre-blend ----
        uint8_t* isOc = isOceanD.ensure(n);
        x.for_each(N, OceanMaskK{tP, r_plate_dev, isOc});
        int* firstOc = scratchI[0].ensure(n); int* minR = scratchI[1].ensure(n);
        uint8_t* landCoastFlag = flagU[6].ensure(n); uint8_t* coastRowFlag = flagU[7].ensure(n);
        x.for_each(N, FillIntK{minR, 0x7fffffff});
        x.for_each(N, CoastFirstK{m->csr(), isOc, firstOc, landCoastFlag, minR});
        x.for_each(N, CoastSeedFlagK{firstOc, minR, coastRowFlag});
        uint8_t* part[6];
        for (int k = 0; k < 6; k++) part[k] = flagU[k].ensure(n);
        uint8_t* inM = setM.ensure(n); uint8_t* inC = setC.ensure(n); uint8_t* inO = setO.ensure(n);
        x.for_each(N, SetFlagsK{dual ? ccode[1].p : ccode[0].p, dual ? ccode[0].p : nullptr, {part[0], part[1], part[2], part[3], part[4], part[5]}, inM, inC, inO});
        counts.ensure(16);
        int* lists[8];
        for (int k = 0; k < 8; k++) lists[k] = listD[k].ensure(n);
        const uint8_t* flagsOf[8] = {part[0], part[1], part[2], part[3], part[4], part[5], landCoastFlag, coastRowFlag};
        for (int k = 0; k < 8; k++) m->prims.compact_flagged(x, flagsOf[k], N, lists[k], counts.p + k);
        int* repD = scratchI[2].ensure(std::max(1, P.n));
        x.for_each(std::max(1, P.n), FillIntK{repD, 0x7fffffff});
        x.for_each(N, PlateRepK{tP, r_plate_dev, inM, inC, inO, repD});
        float* stressDev = stressD.ensure(n); float* subDev = subD.ensure(n);
        btypeD.ensure(n); bothD.ensure(n); hasD.ensure(n);
        if (dual) {
            maxBits.ensure(1);
            dev_memset(maxBits.p, 0, sizeof(int), x.stream);
            x.for_each(N, MaxF32K{cs[1].p, maxBits.p});
            x.for_each(N, BlendPreK{cs[0].p, cs[1].p, cf[0].p, cf[1].p, cb[0].p, cb[1].p, cbo[0].p, cbo[1].p, cho[0].p, cho[1].p, maxBits.p,
                                    stressDev, subDev, btypeD.p, bothD.p, hasD.p});
        }
        const int8_t* btypeDev = dual ? btypeD.p : cb[0].p;
        const uint8_t* bothDev = dual ? bothD.p : cbo[0].p;
        const uint8_t* hasDev = dual ? hasD.p : cho[0].p;
        int hCounts[8];
        uint8_t* hOcean = hIsOcean.ensure(n);
        int* hRep = hRepBuf.ensure(std::max(1, P.n));
        dev_copy(hCounts, counts.p, sizeof hCounts, 1, x.stream);
        dev_copy(hOcean, isOc, n, 1, x.stream);
        dev_copy(hRep, repD, sizeof(int) * (size_t)std::max(1, P.n), 1, x.stream);
        stream_sync(x.stream);
        lap("collisions + sets (device) + d2h");

        // ---- 3. host: the propagation starts at once (per plate, A.7); the seed lists follow ----
        const double baseDecay = 0.5 + spread * 0.04;
        const double decayFactor = pb_pow(baseDecay, 1 / scaleFactor);
        const double subductDecayFactor = pb_pow(baseDecay * 0.45, 1 / scaleFactor);
        const int numPasses = (int)std::max(1.0, jsr(spread * 3 * scaleFactor));
        struct Joiner { std::vector<std::thread>& t; ~Joiner() { for (auto& th : t) if (th.joinable()) th.join(); } };
        std::vector<std::thread> propThreads, fillThreads(5);
        // everything the worker threads touch is declared before the joiners (unwinding joins first, then frees)
        std::vector<int> hl[8];
        std::vector<int> coastSeedList, stressMountain, oceanItems, coastItems;
        std::vector<uint8_t> oceanIn, coastIn, stop;
        std::vector<float> preSubM;
        float* hd[5];
        for (int k = 0; k < 5; k++) hd[k] = hDist[k].ensure(n);
        Joiner joinFills{fillThreads};
        Joiner joinProp{propThreads};
        for (int k = 0; k < 8; k++) { hl[k].resize((size_t)hCounts[k]); dev_copy(hl[k].data(), lists[k], sizeof(int) * (size_t)hCounts[k], 1, x.stream); }
        stream_sync(x.stream);
        // pre-propagation blend of the subduction factor at the mountain cells (:300-309): it survives the post-propagation
        // blend where both propagated stresses vanish, and the propagation overwrites the raw layers in place
        const double SMALL_W = 0.05, SUPER_W = 0.95;
        if (dual)
            for (int pI = 0; pI < 2; pI++)
                for (int r : hl[pI]) {
                    const double wS = SMALL_W * (double)hS[0][r], wP = SUPER_W * (double)hS[1][r], total = wS + wP;
                    preSubM.push_back(total > 1e-6 ? (float)((wS * (double)hF[0][r] + wP * (double)hF[1][r]) / total)
                                                    : (float)(SMALL_W * (double)hF[0][r] + SUPER_W * (double)hF[1][r]));
                }
        start_propagation(propThreads, off, adj, hS[0], hF[0], r_plate, P, decayFactor, subductDecayFactor, numPasses);
        if (dual) start_propagation(propThreads, off, adj, hS[1], hF[1], r_super, *SP, decayFactor, subductDecayFactor, numPasses);
        // coast-distance fills (:411, 426): the longest serial chains of the call, independent of the propagation
        coastSeedList.reserve(hl[7].size());
        for (int r : hl[7])
            for (int j = off[r], e = off[r + 1]; j < e; j++) if (hOcean[adj[j]]) { coastSeedList.push_back(adj[j]); break; }
        fillThreads[3] = std::thread([&] { distance_field(off, adj, N, coastSeedList, nullptr, seed + 4, hd[3]); });
        fillThreads[4] = std::thread([&] { distance_field(off, adj, N, hl[6], hOcean, seed + 5, hd[4]); });
        // seed sets in insertion order: super entries first, then the small-plate entries not present yet (:259-271),
        // plate representatives last, in plateSeeds order (:376-381)
        oceanItems = hl[2]; oceanItems.insert(oceanItems.end(), hl[3].begin(), hl[3].end());
        coastItems = hl[4]; coastItems.insert(coastItems.end(), hl[5].begin(), hl[5].end());
        oceanIn.assign(n, 0); coastIn.assign(n, 0);
        for (int r : oceanItems) oceanIn[r] = 1;
        for (int r : coastItems) coastIn[r] = 1;
        std::vector<int> repAddedO, repAddedC;
        for (int pid : plateSeeds) {
            const int k = P.find(pid);
            if (k < 0 || hRep[k] == 0x7fffffff) continue;
            const int rep = hRep[k];
            if (P.ocean(pid)) { if (!oceanIn[rep]) { oceanIn[rep] = 1; oceanItems.push_back(rep); repAddedO.push_back(rep); } }
            else if (!coastIn[rep]) { coastIn[rep] = 1; coastItems.push_back(rep); repAddedC.push_back(rep); }
        }
        if (!repAddedO.empty()) { dev_copy(scratchI[0].ensure(n), repAddedO.data(), sizeof(int) * repAddedO.size(), 0, x.stream); x.for_each((int)repAddedO.size(), ScatterByteK{scratchI[0].p, inO}); stream_sync(x.stream); }
        if (!repAddedC.empty()) { dev_copy(scratchI[0].ensure(n), repAddedC.data(), sizeof(int) * repAddedC.size(), 0, x.stream); x.for_each((int)repAddedC.size(), ScatterByteK{scratchI[0].p, inC}); stream_sync(x.stream); }
        fillThreads[1] = std::thread([&] { distance_field(off, adj, N, oceanItems, coastIn.data(), seed + 2, hd[1]); });
        lap("seed lists + representatives");

        // ---- 4. join the propagation; the two fills that need it start before anything else (:384-394) ----
        for (auto& th : propThreads) th.join();
        lap("propagateStress (host, per plate)");
        stop = coastIn;
        for (size_t r = 0; r < n; r++) stop[r] |= oceanIn[r];
        {
            // blended subduction factor of the mountain cells only (the device blends the whole field below); mountain_r in
            // insertion order = super entries, then small-plate entries not present yet
            size_t idx = 0;
            for (int pI = 0; pI < 2; pI++)
                for (int r : hl[pI]) {
                    double sb;
                    if (!dual) sb = (double)hF[0][r];
                    else {
                        const double wS = SMALL_W * (double)hS[0][r], wP = SUPER_W * (double)hS[1][r], total = wS + wP;
                        sb = total > 1e-6 ? (double)(float)((wS * (double)hF[0][r] + wP * (double)hF[1][r]) / total) : (double)preSubM[idx];
                    }
                    idx++;
                    if (sb < 0.55) { stressMountain.push_back(r); stop[r] = 1; }
                }
        }
        fillThreads[0] = std::thread([&] { distance_field(off, adj, N, stressMountain, oceanIn.data(), seed + 1, hd[0]); });
        fillThreads[2] = std::thread([&] { distance_field(off, adj, N, coastItems, stop.data(), seed + 3, hd[2]); });

        // ---- 5. device: post-blend, p97, the six capped BFS — while the fills run ----
        if (dual) {
            float* pS0 = dist[0].ensure(n); float* pF0 = dist[1].ensure(n); float* pS1 = dist[2].ensure(n); float* pF1 = dist[3].ensure(n);   // dist buffers are free until the fills end
            dev_copy(pS0, hS[0], sizeof(float) * n, 0, x.stream); dev_copy(pF0, hF[0], sizeof(float) * n, 0, x.stream);
            dev_copy(pS1, hS[1], sizeof(float) * n, 0, x.stream); dev_copy(pF1, hF[1], sizeof(float) * n, 0, x.stream);
            x.for_each(N, BlendPostK{pS0, pS1, pF0, pF1, stressDev, subDev});
        } else {
            dev_copy(stressDev, hS[0], sizeof(float) * n, 0, x.stream);
            dev_copy(subDev, hF[0], sizeof(float) * n, 0, x.stream);
        }
        keys.ensure(n); maxStressD.ensure(1);
        dev_memset(counts.p + 8, 0, sizeof(int), x.stream);
        x.for_each(N, StressKeyK{stressDev, keys.p, counts.p + 8});
        m->prims.sort_keys(x, keys.p, N);
        x.for_each(1, StressPickK{keys.p, counts.p + 8, maxStressD.p});
        const double maxCD = std::max(8.0, jsr(8 * scaleFactor));
        const double riftHalfWidth = std::max(2.0, jsr(4 * scaleFactor)), ridgeHalfWidth = riftHalfWidth;
        const double fractureHalfWidth = std::max(2.0, jsr(3 * scaleFactor));
        const double baStart = std::max(1.0, jsr(2 * scaleFactor)), baPeak = std::max(2.0, jsr(3 * scaleFactor)), baEnd = std::max(3.0, jsr(5 * scaleFactor));
        const double maxArcDist = std::max(5.0, jsr(5 * scaleFactor));
        run_capped_bfs(r_plate_dev, isOc, stressDev, subDev, btypeDev, bothDev, hasDev, maxCD, riftHalfWidth, ridgeHalfWidth, fractureHalfWidth, baEnd, maxArcDist);
        double maxStress = 1;
        dev_copy(&maxStress, maxStressD.p, sizeof(double), 1, x.stream);
        if (dual) stream_sync(x.stream);          // the propagated layers have been read: dist[0..3] may be overwritten below

        // ---- 6. hotspot domes (:1148-1262, host, ≤ 35 domes), join the fills, upload, synthesis ----
        std::vector<DomeDev> domes;
        build_domes(P, r_plate, seed, domes);
        for (auto& th : fillThreads) if (th.joinable()) th.join();
        lap("fills (host) | blend + p97 + capped BFS (device)");
        ElevFields F;
        F.stress = stressDev; F.subduct = subDev; F.btype = btypeDev; F.isOcean = isOc;
        float* dd[5];
        for (int k = 0; k < 5; k++) { dd[k] = dist[k].ensure(n); dev_copy(dd[k], hd[k], sizeof(float) * n, 0, x.stream); }
        F.dist_mountain = dd[0]; F.dist_ocean = dd[1]; F.dist_coastline = dd[2]; F.dist_coast = dd[3]; F.dist_coast_land = dd[4];
        F.riftDist = rift.p; F.ridgeDist = ridge.p; F.fractureDist = fracture.p; F.backArcDist = backArc.p; F.backArcStress = backArcS.p;
        F.coastConvergent = cConvD.p;
        DomeDev* hDomes = hDomeBuf.ensure(PB_MAX_DOMES);
        for (size_t k = 0; k < domes.size(); k++) hDomes[k] = domes[k];
        dev_copy(domesD.ensure(PB_MAX_DOMES), hDomes, sizeof(DomeDev) * domes.size(), 0, x.stream);
        dev_copy(out.stress, stressDev, sizeof(float) * n, 2, x.stream);
        if (out.mountain) dev_copy(out.mountain, inM, n, 2, x.stream);
        if (out.coastline) dev_copy(out.coastline, inC, n, 2, x.stream);
        if (out.ocean) dev_copy(out.ocean, inO, n, 2, x.stream);
        stream_sync(x.stream);                     // maxStress is on the host now

        ElevParams p;
        p.maxStress = maxStress; p.noiseMag = noiseMag; p.scaleFactor = scaleFactor;
        p.interiorBand = std::max(4.0, jsr(16 * scaleFactor)); p.tectonicReach = std::max(6.0, jsr(20 * scaleFactor));
        p.plateauStart = std::max(2.0, jsr(3 * scaleFactor)); p.riftHalfWidth = riftHalfWidth; p.ridgeHalfWidth = ridgeHalfWidth;
        p.fractureHalfWidth = fractureHalfWidth; p.baStart = baStart; p.baPeak = baPeak; p.baEnd = baEnd;
        p.warpOctaves = N > 200000 ? 2 : 3;
        x.for_each(N, ElevationMainK{m->xyz.p, r_plate_dev, tP, F, p, NZ(0), NZ(1), NZ(2), out.elev, out.dbg});
        x.for_each(N, CoastalRoughenK{m->xyz.p, dBdry.p, cStress.p, cSub.p, cConvD.p, stressDev, isOc, maxStress, noiseMag,
                                      std::max(8.0, jsr(8 * scaleFactor)), std::max(4.0, jsr(4 * scaleFactor)), NZ(0), NZ(3), NZ(4), NZ(5),
                                      out.elev, out.dbg.coastal});
        x.for_each(N, IslandArcK{m->xyz.p, arcD.p, arcS.p, maxArcDist, scaleFactor, NZ(6), out.elev, out.dbg.coastal});
        x.for_each(N, HotspotK{m->xyz.p, domesD.p, (int)domes.size(), NZ(7), NZ(8), out.elev, out.dbg.hotspot});
        x.for_each(N, CompressPeaksK{out.elev});
        stream_sync(x.stream);      // pinned host buffers above are the sources of the async uploads
        lap("h2d + synthesis kernels");
    }

    // the six capped FIFO BFS (:464-631, :1059-1086) — pb_elevation_mid.h
    void run_capped_bfs(const int* r_plate_dev, const uint8_t* isOc, const float* stressDev, const float* subDev, const int8_t* btypeDev,
                        const uint8_t* bothDev, const uint8_t* hasDev, double maxCD, double riftHW, double ridgeHW, double fracHW,
                        double baEnd, double maxArc) {
        const Exec& x = ex();
        const size_t n = (size_t)N;
        OBfsAll A{};
        A.plate = r_plate_dev; A.isOcean = isOc; A.g = m->csr();
        float* dists[OBFS_COUNT] = {dBdry.ensure(n), rift.ensure(n), ridge.ensure(n), fracture.ensure(n), backArc.ensure(n), arcD.ensure(n)};
        const int modes[OBFS_COUNT] = {4, 0, 1, 1, 2, 3};
        const double caps[OBFS_COUNT] = {maxCD, riftHW, ridgeHW, fracHW, baEnd, maxArc};
        obCount.ensure(2 * OBFS_COUNT);
        dev_memset(obCount.p, 0, sizeof(int) * 2 * OBFS_COUNT, x.stream);
        for (int k = 0; k < OBFS_COUNT; k++) {
            OBfs& b = A.b[k];
            b.dist = dists[k]; b.pay0 = nullptr; b.pay1 = nullptr; b.pay2 = nullptr; b.best = nullptr;
            b.mode = modes[k]; b.cap = (float)caps[k]; b.upgrade = 0;
            b.front[0] = obFront[k][0].ensure(n); b.front[1] = obFront[k][1].ensure(n); b.count = obCount.p + 2 * k;
            b.claim = obClaim[k].ensure(n); b.seedFlag = flagU[k].ensure(n);
        }
        A.b[OBFS_COAST].pay0 = cStress.ensure(n); A.b[OBFS_COAST].pay1 = cSub.ensure(n); A.b[OBFS_COAST].pay2 = cConvD.ensure(n);
        A.b[OBFS_COAST].upgrade = 1; A.b[OBFS_COAST].best = obBest.ensure(n);
        A.b[OBFS_BACKARC].pay0 = backArcS.ensure(n);
        A.b[OBFS_ARC].pay0 = arcS.ensure(n);
        x.for_each(N, OBfsSeedK{A, stressDev, subDev, btypeDev, bothDev, hasDev, maxStressD.p, (float)(maxCD + 1), (float)(maxArc + 1)});
        for (int k = 0; k < OBFS_COUNT; k++) m->prims.compact_flagged(x, A.b[k].seedFlag, N, A.b[k].front[0], A.b[k].count);
        int maxLevels = 0;
        for (int k = 0; k < OBFS_COUNT; k++) maxLevels = std::max(maxLevels, (int)caps[k]);
#if PB_CUDA
        if (obGrid < 0) {
            int perSm = 0, dev = 0, coop = 0;
            PB_CUDA_CHECK(cudaGetDevice(&dev));
            PB_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
            PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_obfs_persistent, PB_OBFS_THREADS, 0));
            if (!coop || perSm < 1) throw Error("cooperative launch is not available for the capped-BFS kernel");
            obGrid = x.sm_count;
        }
        A.ctaTotals = obTotals.ensure((size_t)OBFS_COUNT * obGrid);
        void* args[] = {&A, &maxLevels};
        launch_stats().launches++;
        ProfScope ps(x.prof, "pb::k_obfs_persistent", x.stream);
        PB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_obfs_persistent, dim3(obGrid), dim3(PB_OBFS_THREADS), args, 0, x.stream));
#else
        (void)maxLevels;
        x.single(OBfsSerialK{A});
#endif
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
