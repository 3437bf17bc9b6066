// pb_plates.h — the plate pipeline on the hi-res mesh.
//
//   projectCoarsePlates        js/coarse-plates.js:51-117   → ProjectPlatesK (one cell per thread)
//   smoothAndReconnectPlates   js/plates.js:241-348         → MajorityVoteK (ordered dataflow), component kernels, host orphan pass
//   buildSuperPlates           js/super-plates.js:16-273    → PlateAreaK / PlatePairFirstK scans, host plate-graph logic, SuperGatherK
//
// Order semantics kept exactly:
//  * projectCoarsePlates walks greedily over the coarse mesh from a start carried across cells; on a Delaunay mesh a
//    strict steepest ascent of the dot product ends at the nearest coarse site whatever the start (and the reference's
//    step-cap fallback is a brute-force search for that same site), so every cell walks from its own start.
//  * The majority vote is an in-place sweep in ascending id: cell r sees the NEW plate of neighbours < r and the OLD plate
//    of neighbours > r.  (plate, pass) travel in one 64-bit word per cell; a cell runs once every lower neighbour carries
//    the current pass number, and higher neighbours cannot overtake it because they wait for it in turn.
//  * Reconnection keeps the largest component per plate, lowest first id on ties (`bfs.length > best.length` is strict and
//    components are met in ascending first id).  The orphan pass (in-place scan + FIFO with first-discoverer payloads)
//    touches only the few orphaned cells and runs on the host.
//  * buildSuperPlates iterates JS Sets in insertion order: a plate's neighbour list is ordered by the first CSR slot at
//    which the pair occurs, which the device computes as a per-pair minimum.
#pragma once
#include "pb_engine.h"

namespace pb {

struct ProjectPlatesK {
    int N, NC, maxWalk; const float* xyz; const int* cOff; const int* cAdj; const float* cxyz; const int* cPlate;
    Simplex noise; double perturbAmp; int* r_plate;
    PB_DEV double dot(double px, double py, double pz, int c) const {
        return px * (double)cxyz[3 * c] + py * (double)cxyz[3 * c + 1] + pz * (double)cxyz[3 * c + 2];
    }
    PB_DEV void operator()(int r) const {
        const double ox = xyz[3 * r], oy = xyz[3 * r + 1], oz = xyz[3 * r + 2];
        double dx = 0, dy = 0, dz = 0, amp = perturbAmp, freq = 8;
        for (int oct = 0; oct < 4; oct++) {
            dx += noise.noise3D(ox * freq, oy * freq, oz * freq) * amp;
            dy += noise.noise3D(ox * freq + 100, oy * freq + 100, oz * freq + 100) * amp;
            dz += noise.noise3D(ox * freq + 200, oy * freq + 200, oz * freq + 200) * amp;
            amp *= 0.5;
            freq *= 2;
        }
        double px = ox + dx, py = oy + dy, pz = oz + dz;
        double len = sqrt(px * px + py * py + pz * pz);
        if (len == 0 || len != len) len = 1;
        px /= len; py /= len; pz /= len;
        // start at the coarse point of the same relative id (same latitude band of the Fibonacci spiral)
        int cur = (int)(((long long)r * (long long)NC) / (long long)N);
        if (cur >= NC) cur = NC - 1;
        double bestDot = dot(px, py, pz, cur);
        bool improved = true;
        int steps = 0;
        while (improved && steps < maxWalk) {
            improved = false;
            steps++;
            for (int i = cOff[cur], e = cOff[cur + 1]; i < e; i++) {
                const int nb = cAdj[i];
                const double d = dot(px, py, pz, nb);
                if (d > bestDot) { bestDot = d; cur = nb; improved = true; }
            }
        }
        if (steps >= maxWalk)
            for (int c = 0; c < NC; c++) { const double d = dot(px, py, pz, c); if (d > bestDot) { bestDot = d; cur = c; } }
        r_plate[r] = cPlate[cur];
    }
};

// ---- majority vote (in-place, ascending id) -------------------------------------------------------------------------
struct PlateWordPackK {
    const int* plate; unsigned long long* word; unsigned pass;
    PB_DEV void operator()(int r) const { word[r] = ((unsigned long long)pass << 32) | (uint32_t)plate[r]; }
};
struct PlateWordUnpackK {
    const unsigned long long* word; int* plate;
    PB_DEV void operator()(int r) const { plate[r] = (int)(uint32_t)(word[r] & 0xFFFFFFFFull); }
};
struct MajorityVoteK {
    Csr g; unsigned long long* word; const uint8_t* isSeed; unsigned pass; double threshold;
    PB_DEV bool try_run(int r) const {
        const int s = g.off[r], e = g.off[r + 1], deg = e - s;
        int plates[32], counts[32], nDistinct = 0;
        for (int j = s; j < e; j++) {
            const int nb = g.adj[j];
            const unsigned long long w = ld_word(word + nb);
            if (nb < r && (unsigned)(w >> 32) != pass) return false;       // lower neighbour not voted yet in this pass
            const int p = (int)(uint32_t)(w & 0xFFFFFFFFull);
            bool found = false;
            for (int k = 0; k < nDistinct; k++) if (plates[k] == p) { counts[k]++; found = true; break; }
            if (!found) { plates[nDistinct] = p; counts[nDistinct] = 1; nDistinct++; }
        }
        const int mine = (int)(uint32_t)(ld_word(word + r) & 0xFFFFFFFFull);
        int bestPlate = mine, bestCount = 0;
        for (int k = 0; k < nDistinct; k++) if (counts[k] > bestCount) { bestCount = counts[k]; bestPlate = plates[k]; }
        const int out = ((double)bestCount > (double)deg * threshold && !isSeed[r]) ? bestPlate : mine;
        st_word(word + r, ((unsigned long long)pass << 32) | (uint32_t)out);
        return true;
    }
};

// The same sweep as a fixed-point iteration.  The sequential result is the unique solution of the triangular system
//     new[r] = vote(new[nb] for nb < r, old[nb] for nb > r, old[r]),
// so any iteration that recomputes every cell from the current values until a whole sweep changes nothing ends at
// exactly that solution (cell 0 is right after one sweep and stays right; a cell is right one sweep after its lower
// neighbours are).  Changes are sparse and local (plate boundaries), so a handful of plain parallel sweeps replace
// the latency-bound dependency chain of MajorityVoteK, which is kept as the fallback when the iteration does not settle.
struct MajorityFixpointK {
    Csr g; const int* old; int* cur; const uint8_t* isSeed; double threshold; int* changed;
    PB_DEV void operator()(int r) const {
        const int s = g.off[r], e = g.off[r + 1], deg = e - s;
        int plates[32], counts[32], nDistinct = 0;
        for (int j = s; j < e; j++) {
            const int nb = g.adj[j];
            const int p = nb < r ? ld_volatile(cur + nb) : old[nb];
            bool found = false;
            for (int k = 0; k < nDistinct; k++) if (plates[k] == p) { counts[k]++; found = true; break; }
            if (!found) { plates[nDistinct] = p; counts[nDistinct] = 1; nDistinct++; }
        }
        const int mine = old[r];
        int bestPlate = mine, bestCount = 0;
        for (int k = 0; k < nDistinct; k++) if (counts[k] > bestCount) { bestCount = counts[k]; bestPlate = plates[k]; }
        const int out = ((double)bestCount > (double)deg * threshold && !isSeed[r]) ? bestPlate : mine;
        if (ld_volatile(cur + r) != out) { cur[r] = out; *changed = 1; }
    }
};

// ---- components of equal plate ------------------------------------------------------------------------------------------
struct PlateCcInitK {
    int* parent; int* size;
    PB_DEV void operator()(int r) const { parent[r] = r; size[r] = 0; }
};
struct PlateCcHookK {
    Csr g; const int* plate; int* parent;
    PB_DEV void operator()(int r) const {
        const int mine = plate[r];
        for (int i = g.off[r], e = g.off[r + 1]; i < e; i++) {
            const int nb = g.adj[i];
            if (nb >= r || plate[nb] != mine) continue;
            int a = r, b = nb;
            for (;;) {
                a = uf_find(parent, a);
                b = uf_find(parent, b);
                if (a == b) break;
                if (a < b) { const int t = a; a = b; b = t; }      // the root of a component is its lowest id
                if (atomic_cas(parent + a, a, b) == a) break;
            }
        }
    }
};
struct PlateCcCountK {
    int* parent; int* size;
    PB_DEV void operator()(int r) const {
        const int root = uf_find(parent, r);
        if (root != r) parent[r] = root;
        atomic_add(size + root, 1);
    }
};
// best[plate] = max over that plate's roots of (size << 32 | ~root): largest component, lowest first id on ties
struct PlateCcBestK {
    const int* plate; const int* parent; const int* size; unsigned long long* best;
    PB_DEV void operator()(int r) const {
        if (parent[r] != r) return;
        atomic_max64(best + plate[r], ((unsigned long long)(uint32_t)size[r] << 32) | (0xFFFFFFFFu - (uint32_t)r));
    }
};
struct PlateInMainK {
    const int* plate; const int* parent; const unsigned long long* best; uint8_t* inMain; int* orphans;
    PB_DEV void operator()(int r) const {
        const int mainRoot = (int)(0xFFFFFFFFu - (uint32_t)(best[plate[r]] & 0xFFFFFFFFull));
        const bool in = uf_find(parent, r) == mainRoot;
        inMain[r] = in;
        if (!in) atomic_add(orphans, 1);
    }
};
struct IntMaxK {
    const int* v; int* out;
    PB_DEV void operator()(int r) const {
        if (v[r] > ld_volatile(out)) atomic_max(out, v[r]);
        if (v[r] < 0) atomic_add(out + 2, 1);            // plate ids are region ids: never negative
    }
};

// ---- buildSuperPlates scans -------------------------------------------------------------------------------------------
struct PlateAreaK {
    const int* plate; const int* pidx; int maxId; int* area; int* bad;
    PB_DEV void operator()(int r) const {
        const int p = plate[r];
        const int k = (p >= 0 && p <= maxId) ? pidx[p] : -1;
        if (k < 0) { atomic_add(bad, 1); return; }
        atomic_add(area + k, 1);
    }
};
// first[me·P + other] = lowest CSR slot at which plate `me` sees plate `other` across an edge (Set insertion order)
struct PlatePairFirstK {
    Csr g; const int* plate; const int* pidx; int P; int* first;
    PB_DEV void operator()(int r) const {
        const int mine = plate[r];
        int me = -1;
        for (int i = g.off[r], e = g.off[r + 1]; i < e; i++) {
            const int np = plate[g.adj[i]];
            if (np == mine) continue;
            if (me < 0) me = pidx[mine];
            int* slot = first + (size_t)me * P + pidx[np];
            if (i < ld_volatile(slot)) atomic_min(slot, i);
        }
    }
};
struct SuperGatherK {
    const int* plate; const int* pidx; const int* plateToSuper; int* r_super;
    PB_DEV void operator()(int r) const { r_super[r] = plateToSuper[pidx[plate[r]]]; }
};

// ---- importHeightmap (js/planet-worker.js:682-831) ---------------------------------------------------------------------
struct SampleHeightmapK {
    const float* xyz; const uint8_t* px; int W, H; float* elev;
    PB_DEV void operator()(int r) const {
        const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
        const double lat = pb_asin(y < -1 ? -1 : (y > 1 ? 1 : y));
        const double lon = pb_atan2(x, z);
        const double fxp = (lon / PB_PI + 1) * 0.5 * W;
        double fyp = (0.5 - lat / PB_PI) * H;
        if (fyp > H - 1) fyp = H - 1;
        if (fyp < 0) fyp = 0;
        const double x0 = floor(fxp), y0 = floor(fyp);
        long long ix0 = (long long)x0 % W; if (ix0 < 0) ix0 += W;             // ((x0 % W) + W) % W
        const long long ix1 = (long long)(x0 + 1) % W;                           // px >= 0, so no negative remainder here
        const long long iy0 = (long long)y0, iy1 = iy0 + 1 > H - 1 ? H - 1 : iy0 + 1;
        const double fx = fxp - x0, fy = fyp - y0;
        const double v00 = px[iy0 * W + ix0], v10 = px[iy0 * W + ix1], v01 = px[iy1 * W + ix0], v11 = px[iy1 * W + ix1];
        const double gray = v00 * (1 - fx) * (1 - fy) + v10 * fx * (1 - fy) + v01 * (1 - fx) * fy + v11 * fx * fy;
        elev[r] = (float)(gray < 1 ? -0.5 : sqrt((gray - 1) / 254));
    }
};
// land / ocean label per region; components of equal label are the synthetic plates, named by their lowest id
struct LandOceanLabelK {
    const float* elev; int* label;
    PB_DEV void operator()(int r) const { label[r] = elev[r] <= 0 ? 1 : 0; }
};
struct RootToPlateK {
    const int* parent; int* r_plate;
    PB_DEV void operator()(int r) const { r_plate[r] = uf_find(parent, r); }
};
struct ClassifyImportedK {
    Csr g; const float* elev; uint8_t* mountain; uint8_t* coastline; uint8_t* ocean;
    PB_DEV void operator()(int r) const {
        const float e = elev[r];
        ocean[r] = e <= 0 ? 1 : 0;
        mountain[r] = (e > 0 && (double)e > 0.5) ? 1 : 0;
        uint8_t coast = 0;
        if (e > 0) for (int i = g.off[r], end = g.off[r + 1]; i < end; i++) if (elev[g.adj[i]] <= 0) { coast = 1; break; }
        coastline[r] = coast;
    }
};

struct PlateTableIn {    // parallel arrays in plateSeeds order
    int n = 0;
    const int* ids = nullptr; const uint8_t* isOcean = nullptr; const double* pole = nullptr; const double* omega = nullptr;
    const double* density = nullptr;
};
struct SuperPlatesOut {
    int n = 0;
    std::vector<double> pole, omega, density;
    std::vector<uint8_t> isOcean;
};

struct Plates {
    Mesh* m;
    DevBuf<int> cOff, cAdj, cPlate, parent, size, scratch, pidx, area, first, toSuper, seedsDev, oldPlate;
    DevBuf<float> cXyz;
    DevBuf<uint8_t> simplexTab, isSeed, inMain, image;
    DevBuf<unsigned long long> word, best;
    explicit Plates(Mesh* mesh) : m(mesh) {}

    // r_plate: device int[N] (out)
    void project(int NC, const int* hCOff, const int* hCAdj, const float* hCXyz, const int* hCPlate, double seed, int numPlates, int* r_plate) {
        if (NC < 1) throw Error("coarse mesh is empty");
        const Exec& x = m->ex();
        const cudaStream_t s = x.stream;
        const size_t cE = (size_t)hCOff[NC];
        for (int c = 0; c < NC; c++) if (hCOff[c + 1] < hCOff[c]) throw Error("coarse adjOffset must be non-decreasing");
        for (size_t i = 0; i < cE; i++) if (hCAdj[i] < 0 || hCAdj[i] >= NC) throw Error("coarse adjList entry out of range");
        dev_copy(cOff.ensure((size_t)NC + 1), hCOff, sizeof(int) * ((size_t)NC + 1), 0, s);
        dev_copy(cAdj.ensure(cE), hCAdj, sizeof(int) * cE, 0, s);
        dev_copy(cXyz.ensure(3 * (size_t)NC), hCXyz, sizeof(float) * 3 * (size_t)NC, 0, s);
        dev_copy(cPlate.ensure(NC), hCPlate, sizeof(int) * (size_t)NC, 0, s);
        SimplexTable tab(seed + 999);
        dev_copy(simplexTab.ensure(1024), tab.t, 1024, 0, s);
        const double coarseEdgeRad = PB_PI / sqrt((double)NC);
        double lowPlateT = 0;
        if (numPlates >= 0) { lowPlateT = (80 - numPlates) / 60.0; if (lowPlateT > 1) lowPlateT = 1; if (lowPlateT < 0) lowPlateT = 0; }
        const double perturbAmp = coarseEdgeRad * (1.5 + 1.0 * lowPlateT);
        // a walk from an arbitrary start may need more steps than the reference's warm-started one; the cap only selects the
        // brute-force search, which returns the same site
        x.for_each(m->N, ProjectPlatesK{m->N, NC, 4 * (int)ceil(sqrt((double)NC)) + 16, m->xyz.p, cOff.p, cAdj.p, cXyz.p, cPlate.p,
                                        Simplex{simplexTab.p}, perturbAmp, r_plate});
        stream_sync(s);      // host staging (tab) goes out of scope
    }

    // r_plate: device int[N], in place
    void smooth_and_reconnect(int* r_plate, const int* seeds, int nSeeds, int numPasses) {
        const Exec& x = m->ex();
        const cudaStream_t s = x.stream;
        const int N = m->N;
        // seed protection (js/plates.js:251-254): only seeds with r_plate[pid] === pid
        std::vector<int> hPlate(N);
        dev_memset(isSeed.ensure(N), 0, (size_t)N, s);
        if (nSeeds > 0) {
            std::vector<int> at(nSeeds, -1);
            for (int k = 0; k < nSeeds; k++)
                if (seeds[k] >= 0 && seeds[k] < N) dev_copy(&at[k], r_plate + seeds[k], sizeof(int), 1, s);
            stream_sync(s);
            const uint8_t one = 1;
            for (int k = 0; k < nSeeds; k++)
                if (seeds[k] >= 0 && seeds[k] < N && at[k] == seeds[k]) dev_copy(isSeed.p + seeds[k], &one, 1, 0, s);
            stream_sync(s);
        }
        scratch.ensure(4);
        static const bool forceOrdered = getenv("PB_PLATES_ORDERED") != nullptr;     // measurement knob
        for (int pass = 0; pass < numPasses; pass++) {
            const double threshold = pass == 0 ? 0.4 : 0.5;
            dev_copy(oldPlate.ensure(N), r_plate, sizeof(int) * (size_t)N, 2, s);
            bool settled = false;
            for (int it = 0; it < 64 && !forceOrdered; it++) {
                dev_memset(scratch.p + 3, 0, sizeof(int), s);
                x.for_each(N, MajorityFixpointK{m->csr(), oldPlate.p, r_plate, isSeed.p, threshold, scratch.p + 3});
                int changed = 0;
                dev_copy(&changed, scratch.p + 3, sizeof(int), 1, s);
                stream_sync(s);
                if (!changed) { settled = true; break; }
            }
            if (!settled) {      // long propagation chain: exact dependency-ordered sweep from the pass's start state
                word.ensure(N);
                x.for_each(N, PlateWordPackK{oldPlate.p, word.p, 0u});
                x.ordered(N, MajorityVoteK{m->csr(), word.p, isSeed.p, 1u, threshold});
                x.for_each(N, PlateWordUnpackK{word.p, r_plate});
            }
        }
        // largest component per plate
        dev_memset(scratch.p, 0, 4 * sizeof(int), s);
        x.for_each(N, IntMaxK{r_plate, scratch.p});
        int maxId = 0;
        int neg = 0;
        dev_copy(&maxId, scratch.p, sizeof(int), 1, s);
        dev_copy(&neg, scratch.p + 2, sizeof(int), 1, s);
        stream_sync(s);
        if (neg) throw Error("r_plate holds negative ids");
        parent.ensure(N); size.ensure(N); inMain.ensure(N);
        best.ensure((size_t)maxId + 1);
        dev_memset(best.p, 0, sizeof(unsigned long long) * ((size_t)maxId + 1), s);
        x.for_each(N, PlateCcInitK{parent.p, size.p});
        x.for_each(N, PlateCcHookK{m->csr(), r_plate, parent.p});
        x.for_each(N, PlateCcCountK{parent.p, size.p});
        x.for_each(N, PlateCcBestK{r_plate, parent.p, size.p, best.p});
        x.for_each(N, PlateInMainK{r_plate, parent.p, best.p, inMain.p, scratch.p + 1});
        int orphans = 0;
        dev_copy(&orphans, scratch.p + 1, sizeof(int), 1, s);
        stream_sync(s);
        if (orphans == 0) return;
        // orphan pass on the host (js/plates.js:321-346): in-place scan, then FIFO with the first discoverer's plate
        std::vector<uint8_t> hMain(N);
        dev_copy(hPlate.data(), r_plate, sizeof(int) * (size_t)N, 1, s);
        dev_copy(hMain.data(), inMain.p, (size_t)N, 1, s);
        stream_sync(s);
        const int* off = m->hOffCopy.data(); const int* adj = m->hAdjCopy.data();
        std::vector<int> queue;
        queue.reserve(orphans);
        for (int r = 0; r < N; r++) {
            if (hMain[r]) continue;
            for (int i = off[r]; i < off[r + 1]; i++)
                if (hMain[adj[i]]) { hPlate[r] = hPlate[adj[i]]; hMain[r] = 1; queue.push_back(r); break; }
        }
        for (size_t qi = 0; qi < queue.size(); qi++) {
            const int r = queue[qi];
            for (int i = off[r]; i < off[r + 1]; i++) {
                const int nb = adj[i];
                if (!hMain[nb]) { hPlate[nb] = hPlate[r]; hMain[nb] = 1; queue.push_back(nb); }
            }
        }
        dev_copy(r_plate, hPlate.data(), sizeof(int) * (size_t)N, 0, s);
        stream_sync(s);
    }

    // importHeightmap: sampleHeightmap (:715-727).  pixels: host u8[W*H]; elev: device float[N]
    void sample_heightmap(const uint8_t* pixels, int W, int H, float* elev) {
        if (W < 1 || H < 1) throw Error("image is empty");
        const Exec& x = m->ex();
        dev_copy(image.ensure((size_t)W * H), pixels, (size_t)W * H, 0, x.stream);
        x.for_each(m->N, SampleHeightmapK{m->xyz.p, image.p, W, H, elev});
        stream_sync(x.stream);
    }
    // deriveSyntheticPlates (:733-769): r_plate[r] = lowest id of r's land mass / ocean basin
    void derive_synthetic_plates(const float* elev, int* r_plate) {
        const Exec& x = m->ex();
        const int N = m->N;
        parent.ensure(N); size.ensure(N); oldPlate.ensure(N);
        x.for_each(N, LandOceanLabelK{elev, oldPlate.p});
        x.for_each(N, PlateCcInitK{parent.p, size.p});
        x.for_each(N, PlateCcHookK{m->csr(), oldPlate.p, parent.p});
        x.for_each(N, RootToPlateK{parent.p, r_plate});
    }
    void classify_imported(const float* elev, uint8_t* mountain, uint8_t* coastline, uint8_t* ocean) {
        m->ex().for_each(m->N, ClassifyImportedK{m->csr(), elev, mountain, coastline, ocean});
    }

    // r_plate: device int[N]; r_super: device int[N] (out)
    void build_super_plates(const int* r_plate, const PlateTableIn& T, int* r_super, SuperPlatesOut& out) {
        const Exec& x = m->ex();
        const cudaStream_t s = x.stream;
        const int N = m->N, P = T.n;
        if (P < 1) throw Error("plate table is empty");
        int maxId = 0;
        for (int k = 0; k < P; k++) { if (T.ids[k] < 0) throw Error("negative plate id"); if (T.ids[k] > maxId) maxId = T.ids[k]; }
        std::vector<int> hIdx((size_t)maxId + 1, -1);
        for (int k = 0; k < P; k++) hIdx[T.ids[k]] = k;
        dev_copy(pidx.ensure(hIdx.size()), hIdx.data(), sizeof(int) * hIdx.size(), 0, s);
        area.ensure((size_t)P + 1);
        dev_memset(area.p, 0, sizeof(int) * ((size_t)P + 1), s);
        first.ensure((size_t)P * P);
        x.for_each(P * P, FillIntK{first.p, 0x7FFFFFFF});
        x.for_each(N, PlateAreaK{r_plate, pidx.p, maxId, area.p, area.p + P});
        std::vector<int> hArea((size_t)P + 1);
        dev_copy(hArea.data(), area.p, sizeof(int) * ((size_t)P + 1), 1, s);
        stream_sync(s);
        if (hArea[P]) throw Error("r_plate holds ids that are not in plateSeeds");
        x.for_each(N, PlatePairFirstK{m->csr(), r_plate, pidx.p, P, first.p});
        std::vector<int> hFirst((size_t)P * P);
        dev_copy(hFirst.data(), first.p, sizeof(int) * hFirst.size(), 1, s);
        stream_sync(s);

        // ---- plate graph (≤ a few hundred nodes): host ----
        std::vector<std::vector<int>> nbrs(P);
        for (int a = 0; a < P; a++) {
            std::vector<std::pair<int, int>> seen;
            for (int b = 0; b < P; b++) if (hFirst[(size_t)a * P + b] != 0x7FFFFFFF) seen.push_back({hFirst[(size_t)a * P + b], b});
            std::sort(seen.begin(), seen.end());
            for (auto& pr : seen) nbrs[a].push_back(pr.second);
        }
        std::vector<std::vector<int>> comps;
        {
            std::vector<char> vis(P, 0);
            for (int p = 0; p < P; p++) {
                if (vis[p]) continue;
                std::vector<int> q{p};
                vis[p] = 1;
                for (size_t h = 0; h < q.size(); h++)
                    for (int nb : nbrs[q[h]]) if (!vis[nb] && (T.isOcean[nb] != 0) == (T.isOcean[p] != 0)) { vis[nb] = 1; q.push_back(nb); }
                comps.push_back(q);       // BFS pop order == push order
            }
        }
        auto jsRound = [](double v) { return floor(v + 0.5); };
        double tgt = jsRound(P / 4.0); if (tgt > 20) tgt = 20; if (tgt < 2) tgt = 2;
        std::vector<int> toSup(P, -1);
        int next = 0;
        const double INF = INFINITY;
        for (const auto& comp : comps) {
            double kd = jsRound(tgt * (double)comp.size() / (double)P); if (kd < 1) kd = 1;
            const int k = (int)kd;
            if (k <= 1) { for (int p : comp) toSup[p] = next; next++; continue; }
            std::vector<char> in(P, 0);
            for (int p : comp) in[p] = 1;
            std::vector<std::vector<int>> ladj(P);
            std::vector<double> w(P, 0), dist(P, INF);
            for (int p : comp) {
                for (int nb : nbrs[p]) if (in[nb]) ladj[p].push_back(nb);
                w[p] = sqrt(hArea[p] ? (double)hArea[p] : 1.0);
            }
            auto relax_all = [&](std::vector<double>& d, std::vector<int>* assign) {
                std::vector<char> done(P, 0);
                for (size_t it = 0; it < comp.size(); it++) {
                    int cur = -1; double mn = INF;
                    for (int p : comp) if (!done[p] && d[p] < mn) { mn = d[p]; cur = p; }
                    if (cur < 0) break;
                    done[cur] = 1;
                    for (int nb : ladj[cur]) {
                        const double nd = d[cur] + w[nb];
                        if (nd < d[nb]) { d[nb] = nd; if (assign) (*assign)[nb] = (*assign)[cur]; }
                    }
                }
            };
            std::vector<int> sd{comp[0]};
            auto from = [&](const std::vector<int>& src) { for (int p : comp) dist[p] = INF; for (int q : src) dist[q] = 0; relax_all(dist, nullptr); };
            from(sd);
            for (int si = 1; si < k; si++) {
                int far = comp[0]; double mx = -1;
                for (int p : comp) if (dist[p] > mx) { mx = dist[p]; far = p; }
                sd.push_back(far);
                from(sd);
            }
            std::vector<int> assign(P, -1);
            std::vector<double> d(P, INF);
            for (size_t si = 0; si < sd.size(); si++) { assign[sd[si]] = next + (int)si; d[sd[si]] = 0; }
            relax_all(d, &assign);
            for (int p : comp) toSup[p] = assign[p];
            next += (int)sd.size();
        }
        const int S = next;
        dev_copy(toSuper.ensure(P), toSup.data(), sizeof(int) * (size_t)P, 0, s);
        x.for_each(N, SuperGatherK{r_plate, pidx.p, toSuper.p, r_super});

        out.n = S;
        out.pole.assign(3 * (size_t)S, 0); out.omega.assign(S, 0); out.density.assign(S, 0); out.isOcean.assign(S, 0);
        std::vector<double> Lx(S, 0), Ly(S, 0), Lz(S, 0), om(S, 0), ar(S, 0), bigA(S, 0), oc(S, 0), tot(S, 0), ds(S, 0), da(S, 0);
        std::vector<int> big(S, -1);
        for (int p = 0; p < P; p++) {
            const int sp = toSup[p];
            const double a = hArea[p];
            tot[sp] += a;
            if (T.isOcean[p]) oc[sp] += a;
            if (T.density[p] == T.density[p]) { ds[sp] += a * T.density[p]; da[sp] += a; }
            const bool hasVec = T.pole[3 * p] == T.pole[3 * p];       // NaN pole ↔ no plateVec entry
            if (!hasVec) continue;
            Lx[sp] += a * T.omega[p] * T.pole[3 * p]; Ly[sp] += a * T.omega[p] * T.pole[3 * p + 1]; Lz[sp] += a * T.omega[p] * T.pole[3 * p + 2];
            om[sp] += a * fabs(T.omega[p]);
            ar[sp] += a;
            if (big[sp] < 0 || a > bigA[sp]) { big[sp] = p; bigA[sp] = a; }
        }
        for (int sp = 0; sp < S; sp++) {
            const double len = sqrt(Lx[sp] * Lx[sp] + Ly[sp] * Ly[sp] + Lz[sp] * Lz[sp]);
            if (len < 1e-8 || ar[sp] < 1) {
                if (big[sp] >= 0) { for (int c = 0; c < 3; c++) out.pole[3 * sp + c] = T.pole[3 * big[sp] + c]; out.omega[sp] = T.omega[big[sp]]; }
                else { out.pole[3 * sp + 1] = 1; out.omega[sp] = 0; }
            } else {
                out.pole[3 * sp] = Lx[sp] / len; out.pole[3 * sp + 1] = Ly[sp] / len; out.pole[3 * sp + 2] = Lz[sp] / len;
                out.omega[sp] = om[sp] / ar[sp];
            }
            out.isOcean[sp] = oc[sp] > tot[sp] * 0.5;
            out.density[sp] = da[sp] > 0 ? ds[sp] / da[sp] : 2.7;
        }
        stream_sync(s);
    }
};

}  // namespace pb
