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
    float* surface; float* key; int* drainTo; uint8_t* visited; uint8_t* seedFlag; uint8_t* openOcean; double* noise;
    PB_DEV void operator()(int r) const {
        const float e = elev[r];
        surface[r] = e;
        const double cn = cell_noise(r);
        if (noise) noise[r] = cn;
        key[r] = (float)((double)e + cn);
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

// Host form of pass 1 (plain C++, both builds) — the engine's default for this pass.  The pass is ONE serial
// chain of |land| heap operations whose tie order between equal f32 keys depends on the binary-heap layout
// (SURVEY.md A.5), so it has to be replayed operation by operation; a CPU core retires that dependent chain
// ~8x faster than a single GPU warp (profiles/r01_flood_heap_ncu.md: ≈ 430 dependent instructions per pop at
// ≈ 9.5 cycles each on the SM).  Like assignDistanceField it is therefore a host-serial stage of the engine
// (class S with a data-dependent global order); "flood=device" selects the one-CTA kernel k_flood_heap instead.
// Same MinHeap semantics as js/terrain-post.js:12-47: sift-up stops on >=, sift-down prefers the left child on
// ties.  The heap stores (key, cell) pairs — keys never change once pushed — with node j in slot j+1 so the two
// children of a node share one aligned 16-byte word.
struct HostHeapEntry { float k; int c; };
struct FloodEK { float elev, key0; };              // per-cell pair read together when a cell is first visited
// device side of the host pass: visited flags as a bitmap (125 KB per million cells: the six flag tests per pop stay in
// the core's L1/L2 instead of touching six cache lines of a byte array) and (elevation, key) interleaved
struct FloodBitmapK {
    const uint8_t* visited; int N; uint32_t* bits;
    PB_DEV void operator()(int w) const {
        uint32_t v = 0;
        const int base = w * 32;
        for (int k = 0; k < 32 && base + k < N; k++) if (visited[base + k]) v |= 1u << k;
        bits[w] = v;
    }
};
struct FloodPackK {
    const float* elev; const float* key0; FloodEK* ek;
    PB_DEV void operator()(int r) const { FloodEK v; v.elev = elev[r]; v.key0 = key0[r]; ek[r] = v; }
};
// filled cells come back as a sparse (cell, surface) list; every other cell keeps surface = elevation
struct FloodFilledK {
    const int* cell; const float* value; float* surface;
    PB_DEV void operator()(int i) const { surface[cell[i]] = value[i]; }
};
inline void flood_heap_host(int N, const int* off, const int* adj, const FloodEK* ek, const float* seedSurface, int* drainTo,
                            uint32_t* visitedBits, const int* seeds, int nSeeds, std::vector<HostHeapEntry>& heapBuf,
                            std::vector<float>& surfTmp, std::vector<int>& filledCell, std::vector<float>& filledSurf) {
    (void)seedSurface;
    if ((int)heapBuf.size() < N + 4) heapBuf.resize((size_t)N + 4);
    HostHeapEntry* h = heapBuf.data() + 1;             // node j at h[j]; h[-1] unused (keeps sibling pairs 16-byte aligned)
    // surface of the cells in the heap travels with them: surf[cell] is written when a cell is visited and read when it is
    // popped; only filled cells differ from their elevation
    if ((int)surfTmp.size() < N) surfTmp.resize((size_t)N);
    float* surf = surfTmp.data();
    filledCell.clear(); filledSurf.clear();
    size_t n = 0;
    auto push = [&](float kc, int cell) {
        size_t i = n++;
        while (i > 0) {
            const size_t p = (i - 1) >> 1;
            const HostHeapEntry pe = h[p];
            if (kc >= pe.k) break;                      // MinHeap.push :22
            h[i] = pe;
            i = p;
        }
        h[i].k = kc; h[i].c = cell;
    };
    for (int s = 0; s < nSeeds; s++) { const int c = seeds[s]; surf[c] = ek[c].elev; push(ek[c].key0, c); }
    while (n > 0) {
        const int r = h[0].c;
        const HostHeapEntry last = h[--n];
        if (n > 0) {                                    // MinHeap.pop :27-46 — last → root, sift down (left child wins ties)
            h[n].k = INFINITY;                          // sentinel: a missing right child never wins (strict <), no bounds test on it
            size_t i = 0;
            for (;;) {
                const size_t l = 2 * i + 1;
                if (l >= n) break;
                const float kl = h[l].k, kr = h[l + 1].k;
                const size_t m = l + (kr < kl ? 1 : 0);          // branch-free child choice: the direction is unpredictable
                const float km = kr < kl ? kr : kl;
                if (!(km < last.k)) break;
                h[i] = h[m];
                i = m;
            }
            h[i] = last;
#if defined(__GNUC__)
            __builtin_prefetch(adj + off[h[0].c], 0, 1);   // the new root is the likely next pop
#endif
        }
        const double surfR = surf[r];
        const double lim = surfR + PB_FLOOD_EPS;
        for (int j = off[r], e = off[r + 1]; j < e; j++) {
            const int nb = adj[j];
            uint32_t& word = visitedBits[nb >> 5];
            const uint32_t bit = 1u << (nb & 31);
            if (word & bit) continue;
            word |= bit;
            drainTo[nb] = r;
            const FloodEK v = ek[nb];
            float kn = v.key0, sn = v.elev;
            if ((double)v.elev < lim) {
                sn = (float)lim;
                filledCell.push_back(nb); filledSurf.push_back(sn);
                // key0 = f32(elev + noise) is not enough here: recompute from the filled surface (cellNoise :100-105)
                const double p1 = (double)nb * 2654435761.0;
                uint32_t hh = (uint32_t)(unsigned long long)p1;
                const int32_t x1 = (int32_t)((hh >> 16) ^ hh);
                const double p2 = (double)x1 * 73244475.0;
                hh = (uint32_t)(unsigned long long)(long long)p2;
                hh = (hh >> 16) ^ hh;
                kn = (float)((double)sn + ((double)hh / 4294967295.0) * 0.01);
            }
            surf[nb] = sn;
            push(kn, nb);
        }
    }
}

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

// =================================================================================================
// CUDA-only fast forms of the two serial passes.  Same results as FloodSerialK / CarveTreeK above
// (which remain the PB_EMUL forms and document the sequential semantics).
// =================================================================================================
#if PB_CUDA
namespace pb {

// ---- (1) heap flood: one CTA --------------------------------------------------------------------
// The reference's binary MinHeap has to be replayed operation by operation (ties between equal f32
// keys are resolved by heap layout), so pass 1 is one serial chain of ~|land| pops and
// a single warp issues dependent instructions at only ~0.23 per cycle (ncu: 42 % fixed-latency
// dependency stalls, profiles/r01_flood_heap_ncu.md).  What can be done is to make every link cheap:
//   * the heap lives in shared memory as (key, cell) pairs — keys never change once pushed, so the
//     inline copy is exact and a sift step is one 16-byte LDS (both children) plus ~14 instructions
//     instead of four dependent L2 round trips; nodes beyond the shared capacity spill to global memory;
//   * the 32 lanes expand the popped cell's neighbours in parallel, and the loads of the popped cell
//     are issued before the sift-down so they overlap it;
//   * the other warps keep the CSR rows and neighbour elevations of the current heap top resident
//     in L1 (read-only prefetching of immutable arrays; the visited flags are read through L2).
// (A warp-cooperative sift that resolves four levels per round was measured 2.2× slower than this
// scalar form: ballots and shuffles cost more dependent instructions than they save.)
struct HeapEntry { uint32_t k; int c; };   // k = __float_as_uint(key); keys compare as floats

struct FloodHeapArgs {
    Csr g; const float* elev; float* surface; int* drainTo; uint8_t* visited;
    const float* key0; const double* noise;   // initial keys f32(elev + cellNoise) and cellNoise per cell (FloodInitK)
    const int* seeds; const int* nSeeds; HeapEntry* spill; int cap;
    int* status;   // [0] max heap size reached
    int prefetchWarp;   // 1: warp 3 keeps the rows of the heap top hot in L1
};

#define PB_FLOOD_THREADS 128

// node j lives in shared slot j+1 for j < cap (so the two children of a node share one aligned
// 16-byte word), else in the global tail
struct HeapStore {
    HeapEntry* sh; HeapEntry* spill; int cap;
    __device__ __forceinline__ HeapEntry ld(int j) const {
        if (j < cap) return sh[j + 1];
        const uint2 v = __ldcg((const uint2*)(spill + (j - cap)));
        HeapEntry e; e.k = v.x; e.c = (int)v.y;
        return e;
    }
    __device__ __forceinline__ void st(int j, HeapEntry e) const {
        if (j < cap) sh[j + 1] = e;
        else __stcg((uint2*)(spill + (j - cap)), make_uint2(e.k, (uint32_t)e.c));
    }
};

__global__ void __launch_bounds__(PB_FLOOD_THREADS, 1) k_flood_heap(FloodHeapArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HeapEntry* sh = (HeapEntry*)smem_raw;                                  // [cap + 2]
    __shared__ volatile int done;
    // mailboxes between the heap warp (warp 0) and the two expansion warps (warps 1 and 2; pop #t goes to warp 1 + t%2)
    __shared__ volatile int reqSeq, reqCell;          // "pop #reqSeq is cell reqCell"
    __shared__ volatile int candSeq, candCell;        // "pop #candSeq will probably be cell candCell" (root after the previous sift-down)
    __shared__ volatile int respSeq[2], respCount[2];
    __shared__ volatile uint32_t respKey[2][32];
    __shared__ volatile int respCell[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.g.N;
    const unsigned FULL = 0xffffffffu;
    if (tid == 0) { done = 0; reqSeq = 0; candSeq = 0; candCell = -1; respSeq[0] = respSeq[1] = 0; respCount[0] = respCount[1] = 0; }
    // every shared slot starts as a (+inf, 0) sentinel and slots beyond the live heap are kept that way, so
    // the sift-down needs no bounds checks while the heap occupies less than half of the shared array
    for (int w = tid; w < a.cap + 2; w += blockDim.x) { sh[w].k = 0x7f800000u; sh[w].c = 0; }
    __syncthreads();

    if (warp == 1 || warp == 2) {
        // ---- expansion warps: lane j owns neighbour j of the popped cell.  A warp reads the CSR row, elevations,
        // initial keys and noise of the cell it EXPECTS to be popped next-but-one as soon as the heap warp has
        // published that candidate (the root left by the previous sift-down), i.e. while the other expansion warp
        // and the heap warp are still busy with the pop before.  When the real request arrives it only has to
        // re-read the visited flags (the other warp may just have claimed a shared neighbour), decide fill / no
        // fill, write surface / drainTo / visited and hand the new (key, cell) pairs back in adjacency order.
        const int me = warp - 1;
        for (int t = 1 + me;; t += 2) {
            // phase A: static data of the candidate (or of the real cell if the request is already there)
            int c = -1;
            for (;;) {
                if (reqSeq >= t) { c = reqCell; break; }
                if (candSeq >= t) { c = candCell; break; }
                if (done) return;
            }
            int b = 0, e = 0, nb = -1; float surfRf = 0.f, el = 0.f, k0 = 0.f; double noise = 0;
            auto load_static = [&](int cell) {
                b = __ldg(a.g.off + cell); e = __ldg(a.g.off + cell + 1);
                surfRf = __ldcg(a.surface + cell);
                nb = -1;
                if (b + lane < e) {
                    nb = __ldg(a.g.adj + b + lane);
                    el = __ldg(a.elev + nb); k0 = __ldg(a.key0 + nb); noise = __ldg(a.noise + nb);
                }
            };
            if (c >= 0) load_static(c);
            // phase B: the real request
            while (reqSeq < t) { if (done) return; }
            __threadfence_block();
            const int r = reqCell;
            if (r != c) load_static(r);
            const double surfR = (double)surfRf;
            uint32_t kbits = 0; bool fresh = false;
            if (nb >= 0 && __ldcg(a.visited + nb) == 0) {
                fresh = true;
                float kf = k0;
                if ((double)el < surfR + PB_FLOOD_EPS) {
                    const float s = (float)(surfR + PB_FLOOD_EPS);
                    __stcg(a.surface + nb, s);
                    kf = (float)((double)s + noise);
                }
                kbits = __float_as_uint(kf);
                __stcg(a.drainTo + nb, r);
                __stcg(a.visited + nb, (uint8_t)1);
            }
            const unsigned m = __ballot_sync(FULL, fresh);
            if (fresh) { const int slot = __popc(m & ((1u << lane) - 1)); respKey[me][slot] = kbits; respCell[me][slot] = nb; }
            if (lane == 0) respCount[me] = __popc(m);
            __syncwarp();
            __threadfence_block();
            if (lane == 0) respSeq[me] = t;
        }
    }
    if (warp > 2) {
        // ---- prefetch helper: keep the rows of the heap's top entries hot in L1 -------------------
        if (!a.prefetchWarp) return;
        while (!done) {
            for (int idx = 0; idx < 15; idx++) {
                int c = ((volatile HeapEntry*)sh)[idx + 1].c;
                if (c < 0 || c >= N) continue;
                const int b = __ldg(a.g.off + c), e = __ldg(a.g.off + c + 1);
                if (lane < e - b && e - b <= 32) {
                    const int nb = __ldg(a.g.adj + b + lane);
                    if (nb >= 0 && nb < N) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.elev + nb));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.key0 + nb));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.noise + nb));
                    }
                }
            }
            __nanosleep(100);
        }
        return;
    }

    // ---- warp 0: the heap engine.  All 32 lanes run the heap operations redundantly (same addresses,
    // same values), which keeps the warp converged.
    const HeapStore H{sh, a.spill, a.cap};
    int n = 0, maxN = 0;
    // The shared-memory paths are kept free of any global-tail code: mixing both in one loop costs the
    // fast path a factor ~2.8 (measured), although the tail is hardly ever touched.
    const int cap = a.cap;
    auto push = [&](uint32_t kbits, int cell) {
        int i = n++;
        const float kc = __uint_as_float(kbits);
        bool placed = false;
        while (i >= cap) {                                       // slow: node in the global tail
            const int p = (i - 1) >> 1;
            const HeapEntry pe = H.ld(p);
            if (kc >= __uint_as_float(pe.k)) { placed = true; break; }
            H.st(i, pe);
            i = p;
        }
        if (!placed)
            while (i > 0) {                                      // fast: shared memory only
                const int p = (i - 1) >> 1;
                const HeapEntry pe = sh[p + 1];
                if (kc >= __uint_as_float(pe.k)) break;          // MinHeap.push stops on >=   (:22)
                sh[i + 1] = pe;
                i = p;
            }
        HeapEntry me; me.k = kbits; me.c = cell;
        if (i < cap) sh[i + 1] = me; else H.st(i, me);
    };
    const int ns = *a.nSeeds;
    for (int s = 0; s < ns; s++) {
        const int c = a.seeds[s];
        push(__float_as_uint(__ldg(a.key0 + c)), c);
    }
    __syncwarp();
    for (int t = 1; n > 0; t++) {
        if (n > maxN) maxN = n;
        // hand the popped cell to the expansion warp, then restore the heap while it works
        if (lane == 0) { reqCell = sh[1].c; reqSeq = t; }      // volatile shared stores stay in program order
        // pop: MinHeap.pop (:27-46) — last → root, sift down along the min-child path (left child on ties)
        --n;
        const HeapEntry last = n < cap ? sh[n + 1] : H.ld(n);
        if (n < cap) { HeapEntry inf; inf.k = 0x7f800000u; inf.c = 0; sh[n + 1] = inf; }     // keep the sentinel invariant
        if (n > 0 && 4 * n + 12 < cap) {
            // small heap: every grandchild index is inside the (sentinel-padded) shared array, so the pairs below BOTH
            // children are loaded one level ahead — the dependent chain per level is compare → select → compare
            // instead of compare → address → shared-memory load (≈ 30 instead of ≈ 100 cycles per level).
            const float kl = __uint_as_float(last.k);
            int i = 0, l = 1;
            uint4 P = *(const uint4*)(sh + l + 1);
            uint4 QL = *(const uint4*)(sh + (2 * l + 1) + 1);
            uint4 QR = *(const uint4*)(sh + (2 * l + 3) + 1);
            for (;;) {
                const bool right = __uint_as_float(P.z) < __uint_as_float(P.x);      // left child wins ties
                const uint32_t mk = right ? P.z : P.x;
                if (!(__uint_as_float(mk) < kl)) break;
                HeapEntry m; m.k = mk; m.c = (int)(right ? P.w : P.y);
                sh[i + 1] = m;
                i = l + (right ? 1 : 0);
                P = right ? QR : QL;
                l = 2 * i + 1;
                QL = *(const uint4*)(sh + (2 * l + 1) + 1);
                QR = *(const uint4*)(sh + (2 * l + 3) + 1);
            }
            sh[i + 1] = last;
        } else if (n > 0) {
            // any heap size: children inside the shared array need no bounds check (slots beyond the heap hold +inf,
            // and when the heap spills every shared slot is live); only the global tail takes the general loop
            const float kl = __uint_as_float(last.k);
            int i = 0;
            bool placed = false;
            for (;;) {
                const int l = 2 * i + 1;
                if (l + 1 >= cap) break;
                const uint4 v = *(const uint4*)(sh + l + 1);
                const bool right = __uint_as_float(v.z) < __uint_as_float(v.x);      // left child wins ties
                const uint32_t mk = right ? v.z : v.x;
                if (!(__uint_as_float(mk) < kl)) { placed = true; break; }
                HeapEntry m; m.k = mk; m.c = (int)(right ? v.w : v.y);
                sh[i + 1] = m;
                i = l + (right ? 1 : 0);
            }
            if (!placed)
                for (;;) {                                       // children in the global tail
                    const int l = 2 * i + 1;
                    if (l >= n) break;
                    const HeapEntry le = H.ld(l);
                    HeapEntry re = le;
                    if (l + 1 < n) re = H.ld(l + 1);
                    const bool right = (l + 1 < n) && (__uint_as_float(re.k) < __uint_as_float(le.k));
                    const HeapEntry m = right ? re : le;
                    if (!(__uint_as_float(m.k) < kl)) break;
                    H.st(i, m);
                    i = l + (right ? 1 : 0);
                }
            if (i < cap) sh[i + 1] = last; else H.st(i, last);
        }
        __syncwarp();
        // the root now is the likely next pop: let the idle expansion warp start on its static data
        if (lane == 0) { candCell = n > 0 ? sh[1].c : -1; candSeq = t + 1; }
        // push the cells the expansion warp discovered, in adjacency order
        const int w = (t + 1) & 1;              // pop #t was handled by expansion warp (t-1)%2 ... see `me` above
        while (respSeq[w] != t) {}
        const int cnt = respCount[w];
        for (int k = 0; k < cnt; k++) push(respKey[w][k], respCell[w][k]);
        __syncwarp();
    }
    if (lane == 0) { done = 1; a.status[0] = maxN; }
}

// ---- (2) carve: binary lifting over the flood forest, one CTA per flood tree ---------------------------
// up[0][c] = drainTo[c] when both c and its target are flooded land, else -1; up[k] = up[k-1]∘up[k-1];
// depth[c] = hops to the tree root (the coastal seed).  Built by pointer doubling.
struct LiftInitK {
    const uint8_t* isOcean; const int* drainTo; int* up0; int* depth;
    PB_DEV void operator()(int c) const {
        int p = -1;
        if (!isOcean[c]) { const int t = drainTo[c]; if (t >= 0 && !isOcean[t]) p = t; }
        up0[c] = p; depth[c] = p >= 0 ? 1 : 0;
    }
};
struct LiftStepK {   // level k from level k-1; depthIn/depthOut ping-pong; *any set when some 2^k-th ancestor exists
    const int* upPrev; int* upNext; const int* depthIn; int* depthOut; int* any;
    PB_DEV void operator()(int c) const {
        const int p = upPrev[c];
        int q = -1, d = depthIn[c];
        if (p >= 0) { q = upPrev[p]; d += depthIn[p]; if (q >= 0) *any = 1; }
        upNext[c] = q; depthOut[c] = d;
    }
};

struct CarveLiftArgs {
    const int* cells; const int* segStart; const int* nSeg; const int* nCells;
    const uint8_t* isOcean; const float* surface; float* elev;
    const int* up; int levels; int N; const int* depth; double carveStrength;
};
#define PB_CARVE_THREADS 256
#define PB_CARVE_PATH_CAP 2048

__device__ __forceinline__ int lift_ancestor(const int* up, int N, int c, int j) {
    for (int k = 0; j; k++, j >>= 1) if (j & 1) c = __ldg(up + (size_t)k * N + c);
    return c;
}

// One CTA per flood tree: the cells of a tree are processed in ascending id, one after the other, and at a million cells a
// drainTo path to the coast is hundreds of cells long, so each step spreads its path over 256 lanes.  (A warp-per-tree form
// without block barriers was measured 2x SLOWER at 1M cells — 53 ms instead of 26 ms per launch: its 32 lanes need 8x more
// dependent binary-lifting rounds per path, which costs more than the barriers it saves.)
__global__ void __launch_bounds__(PB_CARVE_THREADS) k_carve_lift(CarveLiftArgs a) {
    const int s = blockIdx.x;
    if (s >= *a.nSeg) return;
    const int segB = a.segStart[s];
    const int segE = (s + 1 < *a.nSeg) ? a.segStart[s + 1] : *a.nCells;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int sFirst;
    __shared__ double sRedH[PB_CARVE_THREADS / 32];
    __shared__ int sRedJ[PB_CARVE_THREADS / 32];
    __shared__ double sTerm[PB_CARVE_THREADS];
    __shared__ double sKernelSum;
    __shared__ int sPath[PB_CARVE_PATH_CAP];      // ancestors of the current cell (the window pass reuses them)
    int q0 = segB;
    while (q0 < segE) {
        // find the next cell (ascending id) whose current deficit exceeds EPS
        if (tid == 0) sFirst = 0x7fffffff;
        __syncthreads();
        {
            const int q = q0 + tid;
            if (q < segE) {
                const int r = a.cells[q];
                const double deficit = (double)__ldg(a.surface + r) - (double)__ldcg(a.elev + r);
                if (deficit > PB_FLOOD_EPS) atomicMin(&sFirst, q);
            }
        }
        __syncthreads();
        const int qf = sFirst;
        if (qf == 0x7fffffff) { q0 += PB_CARVE_THREADS; __syncthreads(); continue; }
        const int r = a.cells[qf];
        const double deficit = (double)__ldg(a.surface + r) - (double)__ldcg(a.elev + r);
        const int len = a.depth[r] + 1;
        // peak = first maximum along the path (strict >)
        double bh = -INFINITY; int bj = 0x7fffffff;
        for (int j = tid; j < len; j += PB_CARVE_THREADS) {
            const int c = lift_ancestor(a.up, a.N, r, j);
            if (j < PB_CARVE_PATH_CAP) sPath[j] = c;
            const double h = (double)__ldcg(a.elev + c);
            if (h > bh) { bh = h; bj = j; }
        }
        for (int o = 16; o; o >>= 1) {
            const double oh = __shfl_down_sync(0xffffffffu, bh, o);
            const int oj = __shfl_down_sync(0xffffffffu, bj, o);
            if (oh > bh || (oh == bh && oj < bj)) { bh = oh; bj = oj; }
        }
        if (lane == 0) { sRedH[warp] = bh; sRedJ[warp] = bj; }
        __syncthreads();
        int peakIdx;
        {
            double h = sRedH[0]; int j = sRedJ[0];
            for (int w = 1; w < PB_CARVE_THREADS / 32; w++)
                if (sRedH[w] > h || (sRedH[w] == h && sRedJ[w] < j)) { h = sRedH[w]; j = sRedJ[w]; }
            peakIdx = j;
        }
        if (peakIdx == 0x7fffffff) { q0 = qf + 1; __syncthreads(); continue; }   // all-NaN path: reference skips the cell
        const double carveAmount = deficit * a.carveStrength;
        double rad = ceil(len * 0.3);
        if (rad < 3) rad = 3;
        const int radius = (int)rad;
        const int startIdx = peakIdx - radius > 0 ? peakIdx - radius : 0;
        const int endIdx = peakIdx + radius < len - 1 ? peakIdx + radius : len - 1;
        // kernelSum is a sequential double sum in the reference: terms in parallel, sum by one thread
        double ksum = 0;
        for (int base = startIdx; base <= endIdx; base += PB_CARVE_THREADS) {
            const int k = base + tid;
            if (k <= endIdx) {
                const double dist = k > peakIdx ? k - peakIdx : peakIdx - k;
                sTerm[tid] = 1 - dist / (radius + 1);
            }
            __syncthreads();
            if (tid == 0) {
                const int cnt = endIdx - base + 1 < PB_CARVE_THREADS ? endIdx - base + 1 : PB_CARVE_THREADS;
                int t = 0;
                for (; t + 8 <= cnt; t += 8) {       // loads issued together; the additions stay in sequence
                    const double v0 = sTerm[t], v1 = sTerm[t + 1], v2 = sTerm[t + 2], v3 = sTerm[t + 3];
                    const double v4 = sTerm[t + 4], v5 = sTerm[t + 5], v6 = sTerm[t + 6], v7 = sTerm[t + 7];
                    ksum += v0; ksum += v1; ksum += v2; ksum += v3; ksum += v4; ksum += v5; ksum += v6; ksum += v7;
                }
                for (; t < cnt; t++) ksum += sTerm[t];
                sKernelSum = ksum;
            }
            __syncthreads();
        }
        const double kernelSum = sKernelSum;
        if (kernelSum > 0) {
            for (int k = startIdx + tid; k <= endIdx; k += PB_CARVE_THREADS) {
                const int c = k < PB_CARVE_PATH_CAP ? sPath[k] : lift_ancestor(a.up, a.N, r, k);
                const double dist = k > peakIdx ? k - peakIdx : peakIdx - k;
                const double weight = (1 - dist / (radius + 1)) / kernelSum;
                float v = (float)((double)__ldcg(a.elev + c) - carveAmount * weight);
                if (v < 0) v = 0;
                __stcg(a.elev + c, v);
            }
        }
        __syncthreads();
        if (tid == 0) {
            const double fillAmount = deficit * (1 - a.carveStrength);
            __stcg(a.elev + r, (float)((double)__ldcg(a.elev + r) + fillAmount));
        }
        __syncthreads();
        q0 = qf + 1;
    }
}

}  // namespace pb
#endif  // PB_CUDA
