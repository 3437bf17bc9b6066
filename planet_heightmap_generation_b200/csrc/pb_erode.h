// pb_erode.h — the per-iteration passes of erodeComposite (js/terrain-post.js:369-707): kernel
// families K9 (receivers), K11 (forest accumulation + implicit solve), K12 (thermal), K13 (glacial).
//
// Order-dependent (class S) loops of the reference are executed as sync-free dataflow in the
// reference's own order: `order[pos]` is the persistent, stably sorted landCells array and every
// item waits only for the items the sequential loop would have run before it AND whose memory it
// touches.  Per-cell completion counters make the wait O(1): the operations touching a cell x form
// a chain sorted by position; an operation waits until the counter of each cell it touches equals
// its index in that cell's chain (SURVEY.md Appendix A.2–A.4 spell out the touch sets).
#pragma once
#include "pb_platform.h"
#include "pb_stencil.h"
#include "pb_prims.h"

namespace pb {

#define PB_SPIN 4    // polls per try_run before the lane yields to the retry loop (which backs off, pb_platform.h)

struct LandFlagK { const uint8_t* isOcean; uint8_t* flag; PB_DEV void operator()(int r) const { flag[r] = isOcean[r] ? 0 : 1; } };
struct FillIntK { int* p; int v; PB_DEV void operator()(int i) const { p[i] = v; } };
struct SortKeyK {   // key of the cell currently at position i
    const int* order; const float* elev; uint32_t* keys;
    PB_DEV void operator()(int i) const { keys[i] = f32_sort_key(elev[order[i]]); }
};
struct PosK { const int* order; int* pos; PB_DEV void operator()(int i) const { pos[order[i]] = i; } };

// ---- hydraulic: receivers (:566-601) -------------------------------------------------------------
struct ReceiversK {
    Csr g; const float* elev; const uint8_t* isOcean; const float* ndist; int* drainTarget; float* cellDist;
    PB_DEV void operator()(int r) const {
        if (isOcean[r]) { drainTarget[r] = -1; return; }
        const double h = elev[r];
        int bestNb = -1, bestJ = -1;
        double bestDrop = -INFINITY;
        const int b = g.off[r], e = g.off[r + 1];
        for (int j = b; j < e; j++) {
            const int nb = g.adj[j];
            const double drop = h - (double)elev[nb];
            if (drop > bestDrop) { bestDrop = drop; bestNb = nb; bestJ = j; }
        }
        if (bestDrop <= 0) {   // pit: least-steep ascent
            double minAscent = INFINITY;
            for (int j = b; j < e; j++) {
                const int nb = g.adj[j];
                const double ascent = (double)elev[nb] - h;
                if (ascent < minAscent) { minAscent = ascent; bestNb = nb; bestJ = j; }
            }
        }
        drainTarget[r] = bestNb;
        if (bestNb >= 0) cellDist[r] = (float)or_default(ndist[bestJ], 1e-6);
    }
};

// ---- forest accumulation (:604-611 and the ice flow :495-503) ---------------------------------------
// Sequential semantics: flow[x] starts at `init[x]` and receives `+= flow[d]` from every donor d in
// position order, each donor contributing its value AT ITS OWN position.  contrib[r] is that value.
// Ordered over positions (descending elevation).
struct AccumulateK {
    Csr g; const int* order; const int* pos; const int* target; const uint8_t* isOcean;
    const float* initv;      // nullptr → 1.0
    unsigned long long* contrib;     // (value, done) words, cleared before the launch
    PB_DEV bool try_run(int i) const {
        const int r = order[i];
        const int b = g.off[r], e = g.off[r + 1];
        float acc = initv ? initv[r] : 1.0f;
        int last = -1;
        for (;;) {     // donors in position order (repeated-min: degree is tiny); each donor's word is polled directly
            int bp = 0x7fffffff, bd = -1;
            for (int j = b; j < e; j++) {
                const int d = g.adj[j];
                const int p = pos[d];
                if (target[d] == r && p >= 0 && p < i && p > last && p < bp) { bp = p; bd = d; }
            }
            if (bd < 0) break;
            unsigned long long w = 0;
            bool ok = false;
            for (int s = 0; s < PB_SPIN; s++) { w = ld_word(contrib + bd); if (word_seq(w)) { ok = true; break; } }
            if (!ok) return false;
            acc = (float)((double)acc + (double)word_value(w));
            last = bp;
        }
        st_word(contrib + r, make_word(acc, 1));
        return true;
    }
};
// Hydraulic flow accumulation without the dependency chain.  With init = 1 every contribution is an integer below 2^24
// (as long as there are fewer than 16.7M land cells), so the reference's f32 additions (:604-611) are exact and the order of
// the additions cannot matter: contrib[v] — the value v hands to its receiver when the descending sweep reaches it — is the
// number of cells in v's subtree of the forest of EARLY edges (d → target[d] with pos[d] < pos[target[d]]: a donor that sits
// later in the order adds to its receiver only after the receiver has already passed its own value on; AccumulateFinalK adds
// those late donors to the final flow).  Subtree sizes by pointer doubling: after round k, cnt[v] = descendants within
// distance < 2^k, jump[u] = 2^k-th ancestor; ⌈log2(land)⌉ rounds of one integer atomicAdd per cell instead of a chain of
// |longest river| dependent polls (0.58 ms → ≈ 0.1 ms per iteration at 1M cells).
struct SubtreeInitK {
    const int* order; const int* pos; const int* target; const uint8_t* isOcean; int* jump; int* cnt;
    PB_DEV void operator()(int i) const {
        const int r = order[i];
        const int t = target[r];
        int j = -1;
        if (t >= 0 && !isOcean[t] && pos[t] > i) j = t;
        jump[r] = j; cnt[r] = 1;
    }
};
struct SubtreeCopyK {   // next = cur for the land cells (the round then adds the far descendants)
    const int* order; const int* cur; int* next;
    PB_DEV void operator()(int i) const { const int r = order[i]; next[r] = cur[r]; }
};
struct SubtreeRoundK {
    const int* order; const int* jumpIn; int* jumpOut; const int* cntIn; int* cntOut;
    PB_DEV void operator()(int i) const {
        const int r = order[i];
        const int j = jumpIn[r];
        int jj = -1;
        if (j >= 0) { atomic_add(cntOut + j, cntIn[r]); jj = jumpIn[j]; }
        jumpOut[r] = jj;
    }
};
struct SubtreeWordsK {
    const int* order; const int* cnt; unsigned long long* contrib;
    PB_DEV void operator()(int i) const { const int r = order[i]; contrib[r] = make_word((float)cnt[r], 1); }
};

#if PB_CUDA
}  // namespace pb
#include <cooperative_groups.h>
namespace pb {
// all doubling rounds in one cooperative launch (one CTA per SM) with ONE grid barrier per round: a cell folds the additions
// it received in the previous round (dPrev) into its own count, zeroes that slot for reuse, and hands the folded count to its
// current jump target's slot of the other buffer (dCur).  The rounds stop after the first one in which nobody had a jump
// target left, i.e. after ⌈log2(deepest early-edge chain)⌉ + 1 rounds.
__global__ void __launch_bounds__(256) k_subtree_counts(const int* order, const int* pos, const int* target, const uint8_t* isOcean,
                                                        int* jA, int* jB, int* cnt, int* dPrev, int* dCur, int n, unsigned long long* contrib, int* active) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int i = gtid; i < n; i += stride) {
        const int r = order[i];
        const int t = target[r];
        jA[r] = (t >= 0 && !isOcean[t] && pos[t] > i) ? t : -1;
        cnt[r] = 1; dPrev[r] = 0; dCur[r] = 0;
    }
    if (gtid == 0) { active[0] = 0; active[1] = 0; }
    grid.sync();
    for (int k = 0; k < 40; k++) {
        int mine = 0;
        for (int i = gtid; i < n; i += stride) {
            const int r = order[i];
            const int s = cnt[r] + __ldcg(dPrev + r);
            cnt[r] = s; dPrev[r] = 0;
            const int j = __ldcg(jA + r);
            int jj = -1;
            if (j >= 0) { atomicAdd(dCur + j, s); jj = __ldcg(jA + j); mine = 1; }
            jB[r] = jj;
        }
        if (mine) atomicOr(active + (k & 1), 1);
        if (gtid == 0) active[(k + 1) & 1] = 0;
        grid.sync();
        int* t = jA; jA = jB; jB = t;
        t = dPrev; dPrev = dCur; dCur = t;
        if (!*(volatile int*)(active + (k & 1))) break;
    }
    for (int i = gtid; i < n; i += stride) { const int r = order[i]; contrib[r] = make_word((float)cnt[r], 1); }
}
#endif

// final value: init + every donor's contribution in position order; also the donor count (mod 256)
struct AccumulateFinalK {
    Csr g; const int* pos; const int* target; const uint8_t* isOcean; const float* initv;
    const unsigned long long* contrib; float* flow; uint8_t* nUp;
    PB_DEV void operator()(int r) const {
        if (isOcean[r]) { flow[r] = initv ? initv[r] : 0.0f; if (nUp) nUp[r] = 0; return; }
        const int b = g.off[r], e = g.off[r + 1];
        float acc = initv ? initv[r] : 1.0f;
        int last = -1, n = 0;
        for (;;) {
            int bp = 0x7fffffff, bd = -1;
            for (int j = b; j < e; j++) {
                const int d = g.adj[j];
                const int p = pos[d];
                if (target[d] == r && p >= 0 && p > last && p < bp) { bp = p; bd = d; }
            }
            if (bd < 0) break;
            acc = (float)((double)acc + (double)word_value(contrib[bd]));
            last = bp; n++;
        }
        flow[r] = acc;
        if (nUp) nUp[r] = (uint8_t)(n & 255);
    }
};

// ---- implicit stream-power solve with deposition (:614-641) ------------------------------------------
// op(r) touches r (RW), t = target[r] (RW when land), g = target[t] (R).  chain(x) = {x} ∪ donors(x) ∪
// grand-donors(x) sorted by descending position (= ascending elevation = execution order).
struct SolvePrepK {
    Csr g; const int* pos; const int* target; const uint8_t* isOcean; int* k0; int* k1; int* k2;
    // index of member with position p inside the chain of x = number of members with larger position
    PB_DEV int rank_in_chain(int x, int p) const {
        int k = 0;
        if (pos[x] > p) k++;
        for (int j = g.off[x], e = g.off[x + 1]; j < e; j++) {
            const int d = g.adj[j];
            if (target[d] != x || pos[d] < 0) continue;
            if (pos[d] > p) k++;
            for (int jj = g.off[d], ee = g.off[d + 1]; jj < ee; jj++) {
                const int gd = g.adj[jj];
                if (target[gd] != d || pos[gd] < 0 || gd == x) continue;
                if (pos[gd] > p) k++;
            }
        }
        return k;
    }
    PB_DEV void operator()(int x) const {
        if (isOcean[x]) return;
        // gather the chain (bounded local copy; falls back to recounting when it does not fit)
        const int CAP = 24;
        int mp[CAP];
        int n = 0;
        bool fits = true;
        mp[n++] = pos[x];
        for (int j = g.off[x], e = g.off[x + 1]; j < e; j++) {
            const int d = g.adj[j];
            if (target[d] != x || pos[d] < 0) continue;
            if (n < CAP) mp[n++] = pos[d]; else fits = false;
            for (int jj = g.off[d], ee = g.off[d + 1]; jj < ee; jj++) {
                const int gd = g.adj[jj];
                if (target[gd] != d || pos[gd] < 0 || gd == x) continue;
                if (n < CAP) mp[n++] = pos[gd]; else fits = false;
            }
        }
        auto rank_of = [&](int p) -> int {
            if (!fits) return rank_in_chain(x, p);
            int k = 0;
            for (int q = 0; q < n; q++) if (mp[q] > p) k++;
            return k;
        };
        k0[x] = rank_of(pos[x]);
        for (int j = g.off[x], e = g.off[x + 1]; j < e; j++) {
            const int d = g.adj[j];
            if (target[d] != x || pos[d] < 0) continue;
            k1[d] = rank_of(pos[d]);
            for (int jj = g.off[d], ee = g.off[d + 1]; jj < ee; jj++) {
                const int gd = g.adj[jj];
                if (target[gd] != d || pos[gd] < 0 || gd == x) continue;
                k2[gd] = rank_of(pos[gd]);
            }
        }
    }
};

// The elevation of every land cell travels in a 64-bit (elevation, sequence) word: sequence = number of
// operations of the cell's chain already applied.  An operation polls the words of the cells it touches
// until each carries its own chain index, which hands it the current elevations in the same loads — no
// fences, no separate counters (a naturally aligned 64-bit access is single-copy atomic).
struct PackElevK { const float* elev; unsigned long long* ec; PB_DEV void operator()(int r) const { ec[r] = make_word(elev[r], 0); } };
struct UnpackElevK { const unsigned long long* ec; const uint8_t* isOcean; float* elev; PB_DEV void operator()(int r) const { if (!isOcean[r]) elev[r] = word_value(ec[r]); } };
struct SolveK {
    const int* order; int landCount; const int* target; const uint8_t* isOcean;
    const float* cellDist; const float* flow; const float* elev; unsigned long long* ec; const int* k0; const int* k1; const int* k2;
    double K, m, dt;
    PB_DEV bool try_run(int a) const {
        const int r = order[landCount - 1 - a];     // ascending elevation
        const int t = target[r];
        const bool tLand = t >= 0 && !isOcean[t];
        const int gg = tLand ? target[t] : -1;
        const bool gLand = gg >= 0 && !isOcean[gg] && gg != r;
        const int n0 = k0[r], n1 = tLand ? k1[r] : 0, n2 = gLand ? k2[r] : 0;
        // everything that does not depend on other items is evaluated BEFORE the polling loop (the lane would only wait
        // otherwise): the stream-power factor with its pow and divide, the static receiver / grand-receiver elevations
        const float cd = cellDist[r];
        const bool active = t >= 0 && cd > 0;
        const double factor = active ? K * pb_pow((double)flow[r], m) * dt / (double)cd : 0.0;
        const double htStatic = (active && !tLand) ? (double)elev[t] : 0.0;            // ocean receivers never change
        const float cdt = (active && tLand && gg >= 0) ? cellDist[t] : 0.0f;
        const double hgStatic = (active && tLand && gg >= 0 && !gLand && gg != r) ? (double)elev[gg] : 0.0;
        unsigned long long wr = 0, wt = 0, wg = 0;
        bool ok = false;
        for (int s = 0; s < PB_SPIN; s++) {
            wr = ld_word(ec + r);                       // three independent loads per poll: one L2 round trip
            if (tLand) wt = ld_word(ec + t);
            if (gLand) wg = ld_word(ec + gg);
            if (word_seq(wr) != n0) continue;
            if (tLand && word_seq(wt) != n1) continue;
            if (gLand && word_seq(wg) != n2) continue;
            ok = true;
            break;
        }
        if (!ok) return false;
        float er = word_value(wr), et = tLand ? word_value(wt) : 0.0f;
        if (active) {
            const double hr0 = er;
            const double ht = tLand ? (double)et : htStatic;
            const double hrec = ht > 0 ? ht : 0.0;                 // Math.max(h_t, 0)
            double hnew = (hr0 + factor * hrec) / (1 + factor);
            if (hnew < hrec) hnew = hrec;
            if (hnew < 0) hnew = 0;
            const double eroded = hr0 - hnew;
            if (eroded > 0 && tLand) {
                double slope = 0;
                if (gg >= 0 && cdt > 0) {
                    const double hg = gLand ? (double)word_value(wg) : (gg == r ? hr0 : hgStatic);
                    slope = fabs(ht - hg) / (double)cdt;
                }
                const double deposit = eroded * (0.5 / (1 + slope * 50));
                float nt = (float)(ht + deposit);
                if ((double)nt > hnew) nt = (float)hnew;
                et = nt;
            }
            er = (float)hnew;
        }
        if (tLand) st_word(ec + t, make_word(et, n1 + 1));
        if (gLand) st_word(ec + gg, make_word(word_value(wg), n2 + 1));
        st_word(ec + r, make_word(er, n0 + 1));
        return true;
    }
};

// ---- thermal (:645-686) -------------------------------------------------------------------------------
// delta[] is an f32 accumulator in the reference; the events hitting one cell are replayed in
// position order.  Pass 1 stores each cell's totalExcess (double), pass 2 gathers.
struct ThermalExcessK {
    Csr g; const float* elev; const uint8_t* isOcean; const float* ndist; double talus; double* total;
    PB_DEV void operator()(int r) const {
        double te = 0;
        if (!isOcean[r]) {
            const double h = elev[r];
            for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
                const int nb = g.adj[j];
                if (isOcean[nb]) continue;
                const double nh = elev[nb];
                if (nh >= h) continue;
                const double d = or_default(ndist[j], 1e-6);
                const double slope = (h - nh) / d;
                if (slope > talus) te += (slope - talus) * d;
            }
        }
        total[r] = te;
    }
};
struct ThermalApplyK {
    Csr g; const float* elev; float* out; const uint8_t* isOcean; const float* ndist; const int* pos;
    const double* total; double talus, kThermal;
    PB_DEV void operator()(int r) const {
        const float hf = elev[r];
        if (isOcean[r]) { out[r] = hf; return; }
        const double h = hf;
        const int b = g.off[r], e = g.off[r + 1];
        const int myPos = pos[r];
        float delta = 0.0f;
        bool ownDone = !(total[r] > 0);
        int last = -1;
        for (;;) {
            // next incoming event: higher land neighbour whose slope towards r exceeds the talus
            int bp = 0x7fffffff, bj = -1;
            for (int j = b; j < e; j++) {
                const int nb = g.adj[j];
                if (isOcean[nb]) continue;
                const int p = pos[nb];
                if (p <= last || p >= bp) continue;
                const double nh = elev[nb];
                if (!(h < nh)) continue;                    // reference: `if (nh >= h) continue` from nb's side
                const double d = or_default(ndist[j], 1e-6);
                if ((nh - h) / d > talus) { bp = p; bj = j; }
            }
            const bool ownNext = !ownDone && myPos < bp;
            if (ownNext) {
                const double te = total[r];
                const double transfer = kThermal * te * 0.5;
                for (int j = b; j < e; j++) {
                    const int nb = g.adj[j];
                    if (isOcean[nb]) continue;
                    const double nh = elev[nb];
                    if (nh >= h) continue;
                    const double d = or_default(ndist[j], 1e-6);
                    const double slope = (h - nh) / d;
                    if (slope > talus) {
                        const float exf = (float)((slope - talus) * d);
                        const double share = ((double)exf / te) * transfer;
                        delta = (float)((double)delta - share);
                    }
                }
                ownDone = true;
                continue;
            }
            if (bj < 0) break;
            {
                const int nb = g.adj[bj];
                const double nh = elev[nb];
                const double d = or_default(ndist[bj], 1e-6);
                const double slope = (nh - h) / d;
                const float exf = (float)((slope - talus) * d);
                const double te = total[nb];
                const double share = ((double)exf / te) * (kThermal * te * 0.5);
                delta = (float)((double)delta + share);
                last = bp;
            }
        }
        out[r] = (float)(h + (double)delta);
    }
};

// ---- glacial (:404-433, :475-557) ---------------------------------------------------------------------
PB_DEV double smoothstep3(double x, double e0, double e1) {
    double t = (x - e0) / (e1 - e0);
    t = t < 1 ? t : 1.0;   // Math.min(1, t)
    t = t > 0 ? t : 0.0;   // Math.max(0, ·)
    return t * t * (3 - 2 * t);
}
struct GlacIdxK {
    const float* xyz; const float* elev; const uint8_t* isOcean; float* glacIdx; double strength;
    PB_DEV void operator()(int r) const {
        if (isOcean[r]) { glacIdx[r] = 0.0f; return; }
        double y = xyz[3 * r + 1];
        y = y < 1 ? y : 1.0; y = y > -1 ? y : -1.0;
        const double polarDist = fabs(pb_asin(y));
        const double thresholdLat = PB_PI / 2 - strength * PB_PI / 4.5;
        const double latFactor = smoothstep3(polarDist, thresholdLat, PB_PI / 2);
        const double elevFactor = smoothstep3(elev[r], 0.5, 0.9);
        const double latScale = smoothstep3(polarDist, PB_PI / 8, PB_PI / 3);
        const double b = elevFactor * 0.3 * (0.3 + 0.7 * latScale);
        glacIdx[r] = (float)((latFactor > b ? latFactor : b) * strength);
    }
};
struct IceReceiversK {
    Csr g; const float* elev; const uint8_t* isOcean; const float* glacIdx; int* iceTarget;
    PB_DEV void operator()(int r) const {
        int best = -1;
        if (!isOcean[r] && glacIdx[r] > 0) {
            const double h = elev[r];
            double bestDrop = 0;
            for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
                const int nb = g.adj[j];
                const double drop = h - (double)elev[nb];
                if (drop > bestDrop) { bestDrop = drop; best = nb; }
            }
        }
        iceTarget[r] = best;
    }
};
// carve chains: chain(x) = active cells of the closed neighbourhood of x, by ascending position
struct CarvePrepK {
    Csr g; const int* pos; const uint8_t* isOcean; const float* iceFlow; int* kSelf; uint8_t* kEdge;
    PB_DEV bool active(int c) const { return !isOcean[c] && (double)iceFlow[c] > 0.1; }
    PB_DEV void operator()(int x) const {
        if (isOcean[x]) return;
        const int b = g.off[x], e = g.off[x + 1];
        if (active(x)) {
            int k = 0;
            for (int j = b; j < e; j++) { const int o = g.adj[j]; if (active(o) && pos[o] < pos[x]) k++; }
            kSelf[x] = k;
        }
        for (int j = b; j < e; j++) {
            const int o = g.adj[j];
            if (!active(o)) continue;
            int k = 0;
            if (active(x) && pos[x] < pos[o]) k++;
            for (int jj = b; jj < e; jj++) { const int m2 = g.adj[jj]; if (m2 != o && active(m2) && pos[m2] < pos[o]) k++; }
            // store on o's directed edge o→x
            for (int q = g.off[o], qe = g.off[o + 1]; q < qe; q++) if (g.adj[q] == x) { kEdge[q] = (uint8_t)k; break; }
        }
    }
};
struct CarveK {
    Csr g; const int* order; const uint8_t* isOcean; const float* ndist; const float* iceFlow; const uint8_t* nUp;
    float* elev; int* cnt; const int* kSelf; const uint8_t* kEdge;
    double carveRate, convBonus, strength;
    PB_DEV bool try_run(int i) const {
        const int r = order[i];
        const double fl = iceFlow[r];
        if (!(fl > 0.1)) return true;
        const int b = g.off[r], e = g.off[r + 1];
        bool ok = false;
        for (int s = 0; s < PB_SPIN && !ok; s++) {
            ok = ld_volatile(cnt + r) >= kSelf[r];
            for (int j = b; ok && j < e; j++) {
                const int nb = g.adj[j];
                if (isOcean[nb]) continue;
                if (ld_volatile(cnt + nb) < (int)kEdge[j]) ok = false;
            }
        }
        if (!ok) return false;
        fence();
        const double deepening = carveRate * pb_pow(fl, 0.6) * strength;
        float hr = (float)((double)ld_cg(elev + r) - deepening);
        st_cg(elev + r, hr);
        for (int j = b; j < e; j++) {
            const int nb = g.adj[j];
            if (isOcean[nb]) continue;
            const double d = or_default(ndist[j], 1e-6);
            const double hn = ld_cg(elev + nb);
            const double slope = fabs((double)hr - hn) / d;
            double w = 1 - slope; w = w > 0 ? w : 0.0;
            st_cg(elev + nb, (float)(hn - deepening * 0.4 * w));
        }
        if (nUp[r] >= 2) { hr = (float)((double)hr - convBonus * pb_pow(fl, 0.4)); st_cg(elev + r, hr); }
        fence();
        atomic_add(cnt + r, 1);
        for (int j = b; j < e; j++) { const int nb = g.adj[j]; if (!isOcean[nb]) atomic_add(cnt + nb, 1); }
        return true;
    }
};
// moraine (:529-537): deposits onto x from its ice donors in position order; then fjord (:540-551)
// and the land clamp (:554-556) — all three only write the cell itself.
struct MoraineFjordClampK {
    Csr g; const int* pos; const uint8_t* isOcean; const float* glacIdx; const float* iceFlow; const int* iceTarget;
    float* elev; double deposit, fjordCarve;
    PB_DEV void operator()(int x) const {
        if (isOcean[x]) return;
        const int b = g.off[x], e = g.off[x + 1];
        float h = elev[x];
        int last = -1;
        for (;;) {
            int bp = 0x7fffffff, bd = -1;
            for (int j = b; j < e; j++) {
                const int d = g.adj[j];
                if (isOcean[d] || iceTarget[d] != x) continue;
                const int p = pos[d];
                if (p <= last || p >= bp) continue;
                if (!((double)iceFlow[d] > 0.1)) continue;
                if (!((double)glacIdx[x] < (double)glacIdx[d] * 0.3)) continue;
                bp = p; bd = d;
            }
            if (bd < 0) break;
            h = (float)((double)h + deposit * pb_pow((double)iceFlow[bd], 0.3));
            last = bp;
        }
        if ((double)glacIdx[x] > 0.2 && (double)iceFlow[x] > 0.5) {
            bool coastal = false;
            for (int j = b; j < e; j++) if (isOcean[g.adj[j]]) { coastal = true; break; }
            if (coastal) {
                h = (float)((double)h - fjordCarve * pb_pow((double)iceFlow[x], 0.5));
                if (h < 0) h = 0;
            }
        }
        if (h < 0) h = 0;
        elev[x] = h;
    }
};

}  // namespace pb
