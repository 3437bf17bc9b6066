// pb_elevation_mid.h — the middle of assignElevation (js/elevation.js:236-631, :1059-1086) on the device:
//   * ocean mask and the two coast seed sets of the coast-distance fills (:396-424), in the reference's insertion order
//   * dual-layer seed-set unions and stress / subduction / boundary-type blends (:250-327, :343-361)
//   * plate representatives (:368-382), the p97 stress normaliser (:443-453)
//   * the six capped FIFO BFS with first-discoverer payloads (coast boundary :464-509 incl. its equal-level payload
//     upgrade, rift :512-538, ridge :543-568, fracture :571-596, back-arc :601-631, island arc :1059-1086) as ONE
//     cooperative launch that advances all six level by level with an order-preserving frontier (kernel family K8,
//     queue position = (level, queue position of the first-discovering parent, adjacency slot)).
// Only propagateStress (in-place frontier order, A.7 — run per plate on host threads) and the five Park–Miller-driven
// assignDistanceField fills (class R) stay on the host.
#pragma once
#include "pb_platform.h"
#include "pb_stencil.h"
#include "pb_elevation.h"
#include "pb_prims.h"

namespace pb {

struct OceanMaskK {      // :397-399
    PlateTab P; const int* r_plate; uint8_t* isOcean;
    PB_DEV void operator()(int r) const { isOcean[r] = P.ocean(r_plate[r]) ? 1 : 0; }
};
// first ocean neighbour of every land cell (:402-409, :413-418).  coastSeeds is a Set filled in ascending r: the ocean
// cell c enters it when the lowest land cell whose FIRST ocean neighbour is c is visited → minR[c].
struct CoastFirstK {
    Csr g; const uint8_t* isOcean; int* firstOc; uint8_t* landCoast; int* minR;
    PB_DEV void operator()(int r) const {
        int c = -1;
        if (!isOcean[r])
            for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) if (isOcean[g.adj[j]]) { c = g.adj[j]; break; }
        firstOc[r] = c; landCoast[r] = c >= 0 ? 1 : 0;
        if (c >= 0) atomic_min(minR + c, r);
    }
};
struct CoastSeedFlagK {   // land cells that INSERT their first ocean neighbour (ascending r = insertion order)
    const int* firstOc; const int* minR; uint8_t* flag;
    PB_DEV void operator()(int r) const { const int c = firstOc[r]; flag[r] = (c >= 0 && minR[c] == r) ? 1 : 0; }
};

// Seed sets of the two collision layers (:259-271).  part[k][r] = 1 when r enters set k in phase "super" (0,2,4) or
// "small" (1,3,5); in[] = membership after both phases.  Single layer: only the "super" slots are used (with the
// small layer's codes).
struct SetFlagsK {
    const uint8_t* codeSuper; const uint8_t* codeSmall;      // codeSmall == nullptr: single layer
    uint8_t* part[6]; uint8_t* inM; uint8_t* inC; uint8_t* inO;
    PB_DEV void operator()(int r) const {
        const uint8_t cp = codeSuper[r], cs = codeSmall ? codeSmall[r] : 0;
        const bool m0 = cp == 1, m1 = !m0 && cs == 1;
        const bool o0 = cp == 3, o1 = !o0 && cs == 3;
        const bool mt = m0 || m1;
        const bool c0 = cp == 2 && !mt, c1 = !c0 && cs == 2 && !mt;
        part[0][r] = m0; part[1][r] = m1; part[2][r] = o0; part[3][r] = o1; part[4][r] = c0; part[5][r] = c1;
        inM[r] = mt; inC[r] = c0 || c1; inO[r] = o0 || o1;
    }
};
struct PlateRepK {       // :368-375 first unclaimed cell of every plate in id order
    PlateTab P; const int* r_plate; const uint8_t* inM; const uint8_t* inC; const uint8_t* inO; int* rep;
    PB_DEV void operator()(int r) const {
        if (inM[r] || inC[r] || inO[r]) return;
        const int k = P.find(r_plate[r]);
        if (k >= 0) atomic_min(rep + k, r);
    }
};
struct ScatterByteK { const int* idx; uint8_t* dst; PB_DEV void operator()(int i) const { dst[idx[i]] = 1; } };

// dual-layer blends before the propagation (:273-326); maxBits = bit pattern of max(super stress)
struct BlendPreK {
    const float* sS; const float* sP; const float* fS; const float* fP; const int8_t* bS; const int8_t* bP;
    const uint8_t* boS; const uint8_t* boP; const uint8_t* hoS; const uint8_t* hoP; const int* maxBits;
    float* stress; float* sub; int8_t* btype; uint8_t* bothOcean; uint8_t* hasOcean;
    PB_DEV void operator()(int r) const {
        const double SMALL_W = 0.05, SUPER_W = 0.95;
        float mxf;
#if PB_CUDA
        mxf = __int_as_float(*maxBits);
#else
        { const int b = *maxBits; memcpy(&mxf, &b, 4); }
#endif
        const double maxSuperStress = mxf;
        const double invMax = maxSuperStress > 1e-6 ? 1 / maxSuperStress : 0;
        const double s = sS[r], p = sP[r];
        double proximity = p * invMax * 3; if (proximity > 1) proximity = 1;
        const double effectiveSmallW = SMALL_W * (SMALL_W + (1 - SMALL_W) * proximity);
        stress[r] = (float)(effectiveSmallW * s + SUPER_W * p);
        const double wS = SMALL_W * s, wP = SUPER_W * p, total = wS + wP;
        if (total > 1e-6) sub[r] = (float)((wS * (double)fS[r] + wP * (double)fP[r]) / total);
        else sub[r] = (float)(SMALL_W * (double)fS[r] + SUPER_W * (double)fP[r]);
        btype[r] = wS > wP ? bS[r] : bP[r];
        bothOcean[r] = boS[r] | boP[r];
        hasOcean[r] = hoS[r] | hoP[r];
    }
};
// blend of the two propagated layers (:352-361); sub keeps its pre-propagation blend where total <= 1e-6
struct BlendPostK {
    const float* sS; const float* sP; const float* fS; const float* fP; float* stress; float* sub;
    PB_DEV void operator()(int r) const {
        const double SMALL_W = 0.05, SUPER_W = 0.95;
        stress[r] = (float)(SMALL_W * (double)sS[r] + SUPER_W * (double)sP[r]);
        const double wS = SMALL_W * (double)sS[r], wP = SUPER_W * (double)sP[r], total = wS + wP;
        if (total > 1e-6) sub[r] = (float)((wS * (double)fS[r] + wP * (double)fP[r]) / total);
    }
};

// p97 of the non-trivial stresses (:443-453): order-preserving keys (invalid = 0xffffffff sorts last), count, pick
struct StressKeyK {
    const float* stress; uint32_t* keys; int* count;
    PB_DEV void operator()(int r) const {
        const bool ok = (double)stress[r] > 0.01;
        keys[r] = ok ? f32_sort_key(stress[r]) : 0xffffffffu;
        if (ok) atomic_add(count, 1);
    }
};
struct StressPickK {
    const uint32_t* sortedKeys; const int* count; double* maxStress;
    PB_DEV void operator()(int) const {
        const int n = *count;
        double v = 0;           // no value above 0.01: the running maximum is below 0.01 → 1 (:453)
        if (n > 0) {
            long long k = (long long)floor((double)n * 0.97);
            if (k > n - 1) k = n - 1;
            const uint32_t u = sortedKeys[k];
            const uint32_t bits = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
            float f;
#if PB_CUDA
            f = __uint_as_float(bits);
#else
            memcpy(&f, &bits, 4);
#endif
            v = f;
        }
        if (v < 0.01) v = 1;
        *maxStress = v;
    }
};

// ---- capped FIFO BFS family ------------------------------------------------------------------------------------
enum { OBFS_COAST = 0, OBFS_RIFT, OBFS_RIDGE, OBFS_FRACTURE, OBFS_BACKARC, OBFS_ARC, OBFS_COUNT };
struct OBfs {
    float* dist; float* pay0; float* pay1; uint8_t* pay2;     // payloads follow the first discoverer (nullable)
    int mode;            // who may be entered: 0 same plate && land, 1 ocean, 2 same plate, 3 same plate && ocean, 4 anybody
    float cap;           // a popped cell with dist + 1 > cap does not expand
    int upgrade;         // coast BFS: a same-level parent with strictly larger pay0 replaces the payload (:502-506)
    int* front[2]; int* count;     // ping-pong frontiers and their sizes (count[0], count[1])
    int* claim;          // per cell: lowest (item * 32 + slot) that reached it at the current level
    unsigned long long* best;      // upgrade: per cell max of (pay0 bits << 32 | ~item)
    uint8_t* seedFlag;
};
struct OBfsAll { OBfs b[OBFS_COUNT]; const int* plate; const uint8_t* isOcean; Csr g; int* ctaTotals; };

PB_DEV bool obfs_ok(const OBfsAll& A, int mode, int pl, int nr) {
    switch (mode) {
        case 0: return A.plate[nr] == pl && !A.isOcean[nr];
        case 1: return A.isOcean[nr] != 0;
        case 2: return A.plate[nr] == pl;
        case 3: return A.plate[nr] == pl && A.isOcean[nr];
        default: return true;
    }
}

// seeds and initial values of the six BFS in one pass (:476-486, :514-521, :545-552, :573-580, :604-613, :1060-1070)
struct OBfsSeedK {
    OBfsAll A; const float* stress; const float* sub; const int8_t* btype; const uint8_t* bothOcean; const uint8_t* hasOcean;
    const double* maxStress; float coastInit, arcInit;
    PB_DEV void operator()(int r) const {
        const double ms = *maxStress;
        double v = (double)stress[r] / ms; if (v > 1) v = 1;
        const float norm = (float)v;
        const int8_t bt = btype[r];
        bool seed[OBFS_COUNT];
        const uint8_t rOc = A.isOcean[r];
        bool bd = false;
        for (int j = A.g.off[r], e = A.g.off[r + 1]; j < e; j++) if (A.isOcean[A.g.adj[j]] != rOc) { bd = true; break; }
        seed[OBFS_COAST] = bd;
        seed[OBFS_RIFT] = bt == 2 && !hasOcean[r];
        seed[OBFS_RIDGE] = bt == 2 && bothOcean[r];
        seed[OBFS_FRACTURE] = bt == 3 && bothOcean[r];
        seed[OBFS_BACKARC] = bt == 1 && hasOcean[r] && (double)sub[r] < 0.50;
        seed[OBFS_ARC] = bt == 1 && bothOcean[r] && (double)sub[r] < 0.45;
        for (int k = 0; k < OBFS_COUNT; k++) {
            const OBfs& b = A.b[k];
            const float init = k == OBFS_COAST ? coastInit : k == OBFS_ARC ? arcInit : INFINITY;
            b.dist[r] = seed[k] ? 0.f : init;
            b.seedFlag[r] = seed[k] ? 1 : 0;
            b.claim[r] = 0x7fffffff;
            if (b.best) b.best[r] = 0ull;
        }
        A.b[OBFS_COAST].pay0[r] = bd ? norm : 0.f;
        A.b[OBFS_COAST].pay1[r] = bd ? sub[r] : 0.f;
        A.b[OBFS_COAST].pay2[r] = (bd && bt == 1) ? 1 : 0;
        A.b[OBFS_BACKARC].pay0[r] = seed[OBFS_BACKARC] ? norm : 0.f;
        A.b[OBFS_ARC].pay0[r] = seed[OBFS_ARC] ? norm : 0.f;
    }
};

// Sequential form (the reference's loops; PB_EMUL build and documentation of the semantics).
struct OBfsSerialK {
    OBfsAll A;
    PB_DEV void operator()() const {
        for (int k = 0; k < OBFS_COUNT; k++) {
            const OBfs& b = A.b[k];
            int* q = b.front[0];              // capacity N: every cell enters a queue at most once
            int qn = b.count[0];
            for (int qi = 0; qi < qn; qi++) {
                const int r = q[qi];
                const double nd = (double)b.dist[r] + 1;
                if (nd > (double)b.cap) continue;
                const int pl = A.plate[r];
                for (int j = A.g.off[r], e = A.g.off[r + 1]; j < e; j++) {
                    const int nr = A.g.adj[j];
                    if (nd < (double)b.dist[nr] && obfs_ok(A, b.mode, pl, nr)) {
                        b.dist[nr] = (float)nd;
                        if (b.pay0) b.pay0[nr] = b.pay0[r];
                        if (b.pay1) b.pay1[nr] = b.pay1[r];
                        if (b.pay2) b.pay2[nr] = b.pay2[r];
                        q[qn++] = nr;
                    } else if (b.upgrade && nd == (double)b.dist[nr] && b.pay0[r] > b.pay0[nr]) {
                        b.pay0[nr] = b.pay0[r]; b.pay1[nr] = b.pay1[r]; b.pay2[nr] = b.pay2[r];
                    }
                }
            }
        }
    }
};

}  // namespace pb

#if PB_CUDA
#include <cooperative_groups.h>
namespace pb {

#define PB_OBFS_THREADS 256
// One cooperative launch, one CTA per SM.  Level L of every BFS:
//   phase 1  every frontier item (queue position i) offers key i*32+slot to each unvisited, admissible neighbour
//            (atomicMin → the first discoverer in queue order); the coast BFS also records the best payload parent
//   phase 2  the items are split into contiguous runs per CTA and per thread; every thread counts the neighbours it won
//   phase 3  exclusive scan (thread → CTA → grid) gives each winner its exact position in the next frontier; the winner
//            writes dist, payloads and the queue entry
// i.e. three grid barriers per level for all six BFS together (max cap = 80 levels at 1M cells).
__global__ void __launch_bounds__(PB_OBFS_THREADS) k_obfs_persistent(OBfsAll A, int maxLevels) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    __shared__ int sWarp[PB_OBFS_THREADS / 32];
    __shared__ int sBase;
    int myCount[OBFS_COUNT], myBegin[OBFS_COUNT], myEnd[OBFS_COUNT];
    for (int level = 0; level < maxLevels; level++) {
        const int cur = level & 1, nxt = cur ^ 1;
        const float nd = (float)(level + 1);
        bool anyActive = false;
        // ---- phase 1: claims
        for (int k = 0; k < OBFS_COUNT; k++) {
            const OBfs& b = A.b[k];
            const int n = ld_volatile(b.count + cur);
            const bool active = n > 0 && (double)nd <= (double)b.cap;
            if (!active) continue;
            anyActive = true;
            const int* F = b.front[cur];
            for (int i = cta * PB_OBFS_THREADS + tid; i < n; i += G * PB_OBFS_THREADS) {
                const int r = __ldcg(F + i);
                const int pl = A.plate[r];
                const uint32_t p0 = b.upgrade ? __float_as_uint(__ldcg(b.pay0 + r)) : 0u;
                const int o = A.g.off[r], e = A.g.off[r + 1];
                for (int j = o; j < e; j++) {
                    const int nr = A.g.adj[j];
                    if (!(nd < __ldcg(b.dist + nr)) || !obfs_ok(A, b.mode, pl, nr)) continue;
                    atomicMin(b.claim + nr, i * 32 + (j - o));
                    if (b.upgrade) atomicMax(b.best + nr, ((unsigned long long)p0 << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i));
                }
            }
        }
        if (!anyActive) break;          // uniform across the grid: counts and caps are the same for every thread
        grid.sync();
        // ---- phase 2: wins per thread over contiguous runs (order-preserving)
        for (int k = 0; k < OBFS_COUNT; k++) {
            const OBfs& b = A.b[k];
            myCount[k] = 0; myBegin[k] = myEnd[k] = 0;
            const int n = ld_volatile(b.count + cur);
            if (!(n > 0 && (double)nd <= (double)b.cap)) continue;
            const int perCta = (n + G - 1) / G;
            const int perThr = (perCta + PB_OBFS_THREADS - 1) / PB_OBFS_THREADS;
            const int cb = cta * perCta, ce = min(n, cb + perCta);
            const int tb = min(ce, cb + tid * perThr), te = min(ce, tb + perThr);
            myBegin[k] = tb; myEnd[k] = te;
            const int* F = b.front[cur];
            int c = 0;
            for (int i = tb; i < te; i++) {
                const int r = __ldcg(F + i);
                const int o = A.g.off[r], e = A.g.off[r + 1];
                for (int j = o; j < e; j++) {
                    const int nr = A.g.adj[j];
                    if (__ldcg(b.claim + nr) == i * 32 + (j - o) && nd < __ldcg(b.dist + nr)) c++;
                }
            }
            myCount[k] = c;
            // CTA total
            int v = c;
            for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) sWarp[warp] = v;
            __syncthreads();
            if (tid == 0) { int t = 0; for (int w = 0; w < PB_OBFS_THREADS / 32; w++) t += sWarp[w]; A.ctaTotals[k * G + cta] = t; }
            __syncthreads();
        }
        grid.sync();
        // ---- phase 3: ordered write
        for (int k = 0; k < OBFS_COUNT; k++) {
            const OBfs& b = A.b[k];
            const int n = ld_volatile(b.count + cur);
            if (!(n > 0 && (double)nd <= (double)b.cap)) { if (cta == 0 && tid == 0) b.count[nxt] = 0; continue; }
            // grid prefix of the CTA totals
            int part = 0, tot = 0;
            for (int c = tid; c < G; c += PB_OBFS_THREADS) { const int t = __ldcg(A.ctaTotals + k * G + c); tot += t; if (c < cta) part += t; }
            for (int d = 16; d; d >>= 1) { part += __shfl_xor_sync(0xffffffffu, part, d); tot += __shfl_xor_sync(0xffffffffu, tot, d); }
            __syncthreads();
            if (lane == 0) sWarp[warp] = part;
            __syncthreads();
            if (tid == 0) { int t = 0; for (int w = 0; w < PB_OBFS_THREADS / 32; w++) t += sWarp[w]; sBase = t; }
            __syncthreads();
            const int ctaBase = sBase;
            if (cta == 0) {
                __syncthreads();
                if (lane == 0) sWarp[warp] = tot;
                __syncthreads();
                if (tid == 0) { int t = 0; for (int w = 0; w < PB_OBFS_THREADS / 32; w++) t += sWarp[w]; b.count[nxt] = t; }
            }
            // exclusive scan of the per-thread counts inside the CTA
            int incl = myCount[k];
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
            __syncthreads();
            if (lane == 31) sWarp[warp] = incl;
            __syncthreads();
            int warpBase = 0;
            for (int w = 0; w < warp; w++) warpBase += sWarp[w];
            int pos = ctaBase + warpBase + incl - myCount[k];
            const int* F = b.front[cur];
            int* Fn = b.front[nxt];
            for (int i = myBegin[k]; i < myEnd[k]; i++) {
                const int r = __ldcg(F + i);
                const int o = A.g.off[r], e = A.g.off[r + 1];
                for (int j = o; j < e; j++) {
                    const int nr = A.g.adj[j];
                    if (!(__ldcg(b.claim + nr) == i * 32 + (j - o) && nd < __ldcg(b.dist + nr))) continue;
                    int src = r;
                    if (b.upgrade) src = __ldcg(F + (int)(0xffffffffu - (uint32_t)(__ldcg(b.best + nr) & 0xffffffffull)));
                    if (b.pay0) __stcg(b.pay0 + nr, __ldcg(b.pay0 + src));
                    if (b.pay1) __stcg(b.pay1 + nr, __ldcg(b.pay1 + src));
                    if (b.pay2) __stcg(b.pay2 + nr, __ldcg(b.pay2 + src));
                    __stcg(b.dist + nr, nd);
                    __stcg(Fn + pos, nr);
                    pos++;
                }
            }
            __syncthreads();
        }
        grid.sync();
    }
}

}  // namespace pb
#endif
