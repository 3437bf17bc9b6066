// pb_flood.h — priorityFloodCarve (js/terrain-post.js:59-215) as device passes (kernel family K14):
//   (0) open-ocean labelling: lock-free union–find over ocean cells (any CC algorithm gives the
//       reference's answer: only component sizes and membership matter, ties → lowest first id)
//   (1) heap flood: exact emulation of the reference's binary MinHeap (tie order included)
//   (2) carve: one worker per flood tree, filled cells in ascending id inside a tree
//   (3) monotone enforcement: relaxation to the unique fixed point of the ordered recursion
#pragma once
#include "pb_platform.h"
#include "pb_stencil.h"

namespace pb {

#define PB_FLOOD_EPS 1e-7

// js/terrain-post.js:100-105 — Knuth hash with JS semantics: the products exceed 2^53 and round in
// double BEFORE ToUint32, so they must be formed in double, not in uint32.
PB_DEV double cell_noise(int r) {
    const double p1 = (double)r * 2654435761.0;                    // integer-valued double < 2^58
    uint32_t h = (uint32_t)(unsigned long long)p1;                 // ToUint32
    const int32_t x1 = (int32_t)((h >> 16) ^ h);                   // ^ yields int32
    const double p2 = (double)x1 * 73244475.0;                     // 0x45d9f3b, may be negative
    h = (uint32_t)(unsigned long long)(long long)p2;               // ToUint32 (two's complement wrap)
    h = (h >> 16) ^ h;
    return ((double)h / 4294967295.0) * 0.01;
}

// ---- (0) ocean components ---------------------------------------------------------------------
PB_DEV int uf_find(const int* parent, int x) {
    int p = ld_volatile(parent + x);
    while (p != x) { x = p; p = ld_volatile(parent + x); }
    return x;
}
struct CcInitK {
    const uint8_t* isOcean; int* parent; int* size;
    PB_DEV void operator()(int r) const { parent[r] = isOcean[r] ? r : -1; size[r] = 0; }
};
struct CcHookK {
    Csr g; const uint8_t* isOcean; int* parent;
    PB_DEV void operator()(int r) const {
        if (!isOcean[r]) return;
        for (int i = g.off[r], e = g.off[r + 1]; i < e; i++) {
            const int nb = g.adj[i];
            if (nb >= r || !isOcean[nb]) continue;
            int a = r, b = nb;
            for (;;) {
                a = uf_find(parent, a);
                b = uf_find(parent, b);
                if (a == b) break;
                if (a < b) { int t = a; a = b; b = t; }      // hook the larger root under the smaller
                if (atomic_cas(parent + a, a, b) == a) break;
            }
        }
    }
};
struct CcFlattenCountK {
    const uint8_t* isOcean; int* parent; int* size;
    PB_DEV void operator()(int r) const {
        if (!isOcean[r]) return;
        const int root = uf_find(parent, r);
        if (root != r) parent[r] = root;
        atomic_add(size + root, 1);
    }
};
// best = max over roots of (size << 32 | ~root): largest component, lowest first id on ties (:87-90)
struct CcBestK {
    const uint8_t* isOcean; const int* parent; const int* size; unsigned long long* best;
    PB_DEV void operator()(int r) const {
        if (!isOcean[r] || uf_find(parent, r) != r) return;
        atomic_max64(best, ((unsigned long long)(uint32_t)size[r] << 32) | (0xFFFFFFFFu - (uint32_t)r));
    }
};
PB_DEV bool is_open_ocean(int r, const uint8_t* isOcean, const int* parent, const unsigned long long* best) {
    if (!isOcean[r]) return false;
    const unsigned long long b = *best;
    if (b == 0) return false;
    const int mainRoot = (int)(0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFull));
    // parent[] is flattened for every ocean cell except possibly chains created late; walk to be safe
    return uf_find(parent, r) == mainRoot;
}

// ---- (1) keys, seeds, heap flood ----------------------------------------------------------------
struct FloodInitK {
    Csr g; const float* elev; const uint8_t* isOcean; const int* parent; const unsigned long long* best;
    float* surface; float* key; int* drainTo; uint8_t* visited; uint8_t* seedFlag; uint8_t* openOcean;
    PB_DEV void operator()(int r) const {
        const float e = elev[r];
        surface[r] = e;
        key[r] = (float)((double)e + cell_noise(r));
        int dt = -1; uint8_t vis = 0, seed = 0;
        if (isOcean[r]) {
            vis = 1;
            if (openOcean) openOcean[r] = is_open_ocean(r, isOcean, parent, best) ? 1 : 0;
        } else {
            if (openOcean) openOcean[r] = 0;
            for (int i = g.off[r], end = g.off[r + 1]; i < end; i++) {
                const int nb = g.adj[i];
                if (is_open_ocean(nb, isOcean, parent, best)) { vis = 1; seed = 1; dt = nb; break; }
            }
        }
        drainTo[r] = dt; visited[r] = vis; seedFlag[r] = seed;
    }
};

// Serial reference form of the heap flood (one logical thread).  heap[] holds cell ids keyed by the
// external key[] array exactly like the reference's MinHeap (:12-47): sift-up stops on >=, sift-down
// prefers the left child on ties.
struct FloodSerialK {
    Csr g; const float* elev; float* surface; float* key; int* drainTo; uint8_t* visited;
    const int* seeds; const int* nSeeds; int* heap;

    PB_DEV void push(int& n, int cell) const {
        int i = n++;
        heap[i] = cell;
        const float kc = key[cell];
        while (i > 0) {
            const int parent = (i - 1) >> 1;
            const int pc = heap[parent];
            if (kc >= key[pc]) break;
            heap[i] = pc; heap[parent] = cell;
            i = parent;
        }
    }
    PB_DEV int pop(int& n) const {
        const int top = heap[0];
        const int last = heap[--n];
        if (n > 0) {
            heap[0] = last;
            const float kl = key[last];
            int i = 0;
            for (;;) {
                int smallest = i; float ks = kl;
                const int l = 2 * i + 1, r = 2 * i + 2;
                if (l < n) { const float k = key[heap[l]]; if (k < ks) { smallest = l; ks = k; } }
                if (r < n) { const float k = key[heap[r]]; if (k < ks) { smallest = r; ks = k; } }
                if (smallest == i) break;
                heap[i] = heap[smallest]; heap[smallest] = last;
                i = smallest;
            }
        }
        return top;
    }
    PB_DEV void operator()() const {
        int n = 0;
        const int ns = *nSeeds;
        for (int s = 0; s < ns; s++) push(n, seeds[s]);
        while (n > 0) {
            const int r = pop(n);
            const double surfR = surface[r];
            for (int i = g.off[r], end = g.off[r + 1]; i < end; i++) {
                const int nb = g.adj[i];
                if (visited[nb]) continue;
                visited[nb] = 1;
                drainTo[nb] = r;
                if ((double)elev[nb] < surfR + PB_FLOOD_EPS) {
                    const float s = (float)(surfR + PB_FLOOD_EPS);
                    surface[nb] = s;
                    key[nb] = (float)((double)s + cell_noise(nb));
                }
                push(n, nb);
            }
        }
    }
};

// ---- (2) carve ----------------------------------------------------------------------------------
// root of each flooded land cell = the coastal seed its drainTo chain ends at (-1: never flooded)
struct FloodRootK {
    const uint8_t* isOcean; const int* drainTo; const float* surface; const float* elev;
    int* root; uint8_t* rootActive;
    PB_DEV void operator()(int r) const {
        int rt = -1;
        if (!isOcean[r] && drainTo[r] >= 0) {
            int cur = r;
            for (;;) {
                const int nx = drainTo[cur];
                if (nx < 0 || isOcean[nx]) break;
                cur = nx;
            }
            rt = cur;
            if ((double)surface[r] - (double)elev[r] > PB_FLOOD_EPS) rootActive[rt] = 1;
        }
        root[r] = rt;
    }
};
struct FloodMemberFlagK {   // cells of trees that contain at least one filled cell
    const int* root; const uint8_t* rootActive; uint8_t* flag;
    PB_DEV void operator()(int r) const { const int rt = root[r]; flag[r] = (rt >= 0 && rootActive[rt]) ? 1 : 0; }
};
struct GatherIntK { const int* src; const int* idx; int* out; PB_DEV void operator()(int i) const { out[i] = src[idx[i]]; } };
struct SegStartFlagK {      // keys sorted: flag the first element of every run
    const int* keys; uint8_t* flag;
    PB_DEV void operator()(int i) const { flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0; }
};

// js/terrain-post.js:152-196 for the cells of ONE flood tree, ascending id.  Paths never leave the
// tree, so trees are independent; inside a tree the reference's order is kept.
struct CarveTreeK {
    const int* cells; const int* segStart; const int* nSeg; const int* nCells;
    const uint8_t* isOcean; const int* drainTo; const float* surface; float* elev; double carveStrength;
    PB_DEV void operator()(int s) const {
        if (s >= *nSeg) return;
        const int b = segStart[s];
        const int e = (s + 1 < *nSeg) ? segStart[s + 1] : *nCells;
        for (int q = b; q < e; q++) {
            const int r = cells[q];
            const double deficit = (double)surface[r] - (double)elev[r];
            if (deficit <= PB_FLOOD_EPS) continue;
            // walk 1: path length and peak (first maximum, strict >)
            int len = 0, peakIdx = -1;
            double peakElev = -INFINITY;
            for (int cur = r; cur >= 0 && !isOcean[cur]; cur = drainTo[cur]) {
                const double h = elev[cur];
                if (h > peakElev) { peakElev = h; peakIdx = len; }
                len++;
            }
            if (peakIdx < 0 || len == 0) continue;
            const double carveAmount = deficit * carveStrength;
            double rad = ceil(len * 0.3);
            if (rad < 3) rad = 3;
            const int radius = (int)rad;
            const int startIdx = peakIdx - radius > 0 ? peakIdx - radius : 0;
            const int endIdx = peakIdx + radius < len - 1 ? peakIdx + radius : len - 1;
            double kernelSum = 0;
            for (int k = startIdx; k <= endIdx; k++) {
                const double dist = k > peakIdx ? k - peakIdx : peakIdx - k;
                kernelSum += 1 - dist / (radius + 1);
            }
            if (kernelSum > 0) {
                int k = 0;
                for (int cur = r; k <= endIdx; cur = drainTo[cur], k++) {
                    if (k < startIdx) continue;
                    const double dist = k > peakIdx ? k - peakIdx : peakIdx - k;
                    const double weight = (1 - dist / (radius + 1)) / kernelSum;
                    float v = (float)((double)elev[cur] - carveAmount * weight);
                    if (v < 0) v = 0;
                    elev[cur] = v;
                }
            }
            const double fillAmount = deficit * (1 - carveStrength);
            elev[r] = (float)((double)elev[r] + fillAmount);
        }
    }
};

// ---- (3) monotone enforcement ---------------------------------------------------------------------
// Reference (:200-214): land cells in ascending (surface, id) order; a cell not above its drainTo
// target is lifted to target + EPS, reading the target's value at that moment.  Equivalent ordered
// recursion:  e'[r] = e0[r] <= T ? f32(T + EPS) : e0[r],  T = 0 for an ocean target, e'[t] when t
// precedes r in (surface, id) order, e0[t] otherwise.  The recursion is acyclic, so in-place
// relaxation converges to its unique solution; `changed` counts updates per sweep.
struct EnforceK {
    const uint8_t* isOcean; const int* drainTo; const float* surface; const float* e0; float* cur; int* changed;
    PB_DEV void operator()(int r) const {
        if (isOcean[r]) return;
        const int t = drainTo[r];
        if (t < 0) return;
        double T;
        if (isOcean[t]) T = 0.0;
        else {
            const float st = surface[t], sr = surface[r];
            const bool before = (st < sr) || (st == sr && t < r);
            T = before ? (double)ld_cg(cur + t) : (double)e0[t];
        }
        const float base = e0[r];
        const float want = ((double)base <= T) ? (float)(T + PB_FLOOD_EPS) : base;
        if (ld_cg(cur + r) != want) { st_cg(cur + r, want); atomic_add(changed, 1); }
    }
};

}  // namespace pb
