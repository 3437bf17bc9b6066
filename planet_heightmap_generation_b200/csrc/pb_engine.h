// pb_engine.h — host orchestration: context, device-resident mesh, and the stage drivers that
// string the per-cell kernels together in the reference's order (js/planet-worker.js:40-102,
// js/terrain-post.js).  All state that the reference allocates per call as typed arrays lives in
// mesh-owned device buffers that are reused between calls.
#pragma once
#include <chrono>
#include "pb_platform.h"
#include "pb_noise.h"
#include "pb_stencil.h"
#include "pb_prims.h"
#include "pb_flood.h"
#include "pb_erode.h"
#include "../../include/planet_b200.h"

namespace pb {

inline double js_round(double x) { return floor(x + 0.5); }   // Math.round

// contexts alive in this process: several planets in flight on one GPU (one context + stream + host thread each) share the SMs
inline std::atomic<int>& live_contexts() { static std::atomic<int> n{0}; return n; }

struct Context {
    int device = 0;
    int pointerMode = PB_POINTER_HOST;
    int flowMode = 0;                   // option "flow": 0 auto, 1 doubling (integer-exact subtree sizes), 2 ordered (dataflow)
    bool meshOrderDelaunator = false;   // option "mesh_order": "canonical" (device builder) | "delaunator" (reference's own row starts, host)
    bool floodOnHost = true;       // option "flood": "host" (default) = the serial heap pass of priorityFloodCarve on a host core, "device" = k_flood_heap
    Exec ex;
    DevBuf<int> ticket;
    Profiler profiler;
    explicit Context(int dev) : device(dev) {
#if PB_CUDA
        PB_CUDA_CHECK(cudaSetDevice(dev));
        cudaDeviceProp prop;
        PB_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
        ex.sm_count = prop.multiProcessorCount;
#endif
        ex.ticket = ticket.ensure(4);
        ex.prof = &profiler;
        live_contexts()++;
    }
    ~Context() { live_contexts()--; }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    void bind() const {
#if PB_CUDA
        PB_CUDA_CHECK(cudaSetDevice(device));
#endif
    }
};

struct ErodeTaps {
    int captureIter = -1;
    int* drainTarget = nullptr;   // device or host per pointer mode (handled by caller): here DEVICE
    float* flow = nullptr;
    int* landOrder = nullptr;
};
struct FloodTaps {
    int* drainTo = nullptr; float* surface = nullptr; uint8_t* openOcean = nullptr;   // DEVICE
};

struct StageTimer {
#if PB_CUDA
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
#endif
    double ms[5] = {0, 0, 0, 0, 0};
    bool pending = false;
    ~StageTimer() {
#if PB_CUDA
        for (auto e : ev) if (e) cudaEventDestroy(e);
#endif
    }
    void mark(int i, cudaStream_t s) {
#if PB_CUDA
        if (!ev[i]) PB_CUDA_CHECK(cudaEventCreate(&ev[i]));
        PB_CUDA_CHECK(cudaEventRecord(ev[i], s));
#endif
    }
    void resolve() {
#if PB_CUDA
        if (!pending) return;
        PB_CUDA_CHECK(cudaEventSynchronize(ev[5]));
        for (int i = 0; i < 5; i++) { float f = 0; PB_CUDA_CHECK(cudaEventElapsedTime(&f, ev[i], ev[i + 1])); ms[i] = f; }
        pending = false;
#endif
    }
};

struct SweepShards;
struct Mesh;
// `passes` double-buffered sweeps field ← F(field): on this GPU alone, or — when a shard group is attached to the mesh and
// the planet is large enough — over this rank's cell-id range with a peer-memory halo exchange per sweep (pb_shardsweep.h)
void smooth_field_impl(Mesh& m, float* field, int passes);

struct Mesh {
    Context* ctx;
    SweepShards* shards = nullptr;      // set by pb_mesh_attach_shards (not owned)
    int N = 0;
    long long E = 0;
    DevBuf<int> off, adj;
    DevBuf<PackedRow> pack;            // 16-byte packed rows for the sweep family (pb_stencil.h)
    DevBuf<float> xyz, ndist;
    std::vector<int> hOffCopy, hAdjCopy;       // host CSR for the host-serial stages
    std::vector<float> hXyzCopy;
    Prims prims;
    StageTimer timer;

    // staging for host-pointer mode
    DevBuf<float> sElev, sHot, sDelta, sField, sEdge;
    DevBuf<uint8_t> sOcean, sU8;
    DevBuf<int> sI0, sI1;

    // scratch (names follow the reference's typed arrays)
    DevBuf<float> tmp, original, preErosion, warped;
    DevBuf<uint8_t> isOceanBuf, cls, flag8, flag8b, simplexTab;
    // flood
    DevBuf<int> parent, ccSize, drainTo, seeds, heap, root, cells, segStart, counters;
    DevBuf<uint32_t> keys32;
    DevBuf<float> surface, key, e0;
    DevBuf<uint8_t> visited, seedFlag, rootActive, openOcean;
    DevBuf<unsigned long long> best;
    DevBuf<double> cellNoiseBuf;
    // erosion
    DevBuf<int> order, pos, drainTarget, cnt, k0, k1, k2, iceTarget, kSelf;
    DevBuf<float> cellDist, flow, contrib, glacIdx, iceFlow;
    DevBuf<double> total;
    DevBuf<unsigned long long> words;
    DevBuf<uint8_t> nUp, kEdge;

    Mesh(Context* c, int n, const int* hOff, const int* hAdj, const float* hXyz) : ctx(c), N(n) {
        if (n <= 0) throw Error("numRegions must be positive");
        if (hOff[0] != 0) throw Error("adjOffset[0] must be 0");
        for (int r = 0; r < n; r++) if (hOff[r + 1] < hOff[r]) throw Error("adjOffset must be non-decreasing");
        for (int r = 0; r < n; r++) if (hOff[r + 1] - hOff[r] > 32) throw Error("cell degree above 32 is not supported");
        E = hOff[n];
        for (long long i = 0; i < E; i++) if (hAdj[i] < 0 || hAdj[i] >= n) throw Error("adjList entry out of range");
        hOffCopy.assign(hOff, hOff + n + 1); hAdjCopy.assign(hAdj, hAdj + E); hXyzCopy.assign(hXyz, hXyz + 3 * (size_t)n);
        const Exec& ex = ctx->ex;
        dev_copy(off.ensure(n + 1), hOff, sizeof(int) * (size_t)(n + 1), 0, ex.stream);
        dev_copy(adj.ensure(E), hAdj, sizeof(int) * (size_t)E, 0, ex.stream);
        dev_copy(xyz.ensure(3 * (size_t)n), hXyz, sizeof(float) * 3 * (size_t)n, 0, ex.stream);
        ndist.ensure(E);
        ex.for_each(N, NeighborDistK{csr(), xyz.p, ndist.p});
        if (!getenv("PB_NO_PACKED_ROWS")) {
            PackedRow* pk = (PackedRow*)dev_alloc(sizeof(PackedRow) * (size_t)n);      // csr() must not see it before it is filled
            ex.for_each(N, PackRowsK{csr(), pk});
            pack.p = pk; pack.cap = (size_t)n;
        }
        stream_sync(ex.stream);
    }

    Csr csr() const { return Csr{N, off.p, adj.p, pack.p}; }
    const Exec& ex() const { return ctx->ex; }
    bool hostMode() const { return ctx->pointerMode == PB_POINTER_HOST; }

    // ---- pointer-mode plumbing -------------------------------------------------------------------
    template <class T>
    T* arg_in(const T* user, size_t n, DevBuf<T>& stage) {
        if (!user) return nullptr;
        if (!hostMode()) return const_cast<T*>(user);
        dev_copy(stage.ensure(n), user, n * sizeof(T), 0, ex().stream);
        return stage.p;
    }
    template <class T>
    T* arg_out(T* user, size_t n, DevBuf<T>& stage) {   // output-only: no copy in
        if (!user) return nullptr;
        if (!hostMode()) return user;
        return stage.ensure(n);
    }
    template <class T>
    void arg_back(T* user, const T* dev, size_t n) {
        if (!user || !hostMode()) return;
        dev_copy(user, dev, n * sizeof(T), 1, ex().stream);
    }
    void finish() { if (hostMode()) stream_sync(ex().stream); }

    int read_int(const int* d) {
        int v = 0;
        dev_copy(&v, d, sizeof(int), 1, ex().stream);
        stream_sync(ex().stream);
        return v;
    }

    // ---- class-P stages ----------------------------------------------------------------------------
    void smooth_field(float* field, int passes) { smooth_field_impl(*this, field, passes); }   // js/climate-util.js:5-25

    void warp_terrain(float* elev, double seed, double strength, const float* hotspot) {   // :233-309
        if (!(strength > 0)) return;
        SimplexTable tab(seed + 9999);
        dev_copy(simplexTab.ensure(1024), tab.t, 1024, 0, ex().stream);
        warped.ensure(N);
        ex().for_each(N, WarpWalkK{csr(), xyz.p, elev, warped.p, Simplex{simplexTab.p}, 0.12 * strength});
        ex().for_each(N, WarpBlendK{elev, warped.p, hotspot, 0.25 + 0.5 * strength});
    }

    void smooth_elevation(float* elev, const uint8_t* isOcean, int iterations, double strength) {   // :317-354
        if (iterations <= 0) return;
        cls.ensure(N);
        ex().for_each(N, CellClassK{csr(), isOcean, cls.p});
        float* src = elev; float* dst = tmp.ensure(N);
        for (int it = 0; it < iterations; it++) {
            ex().for_each(N, BilateralK{csr(), src, dst, cls.p, strength});
            std::swap(src, dst);
        }
        if (src != elev) dev_copy(elev, src, sizeof(float) * (size_t)N, 2, ex().stream);
    }

    void sharpen_ridges(float* elev, const uint8_t* isOcean, int iterations, double strength) {     // :713-751
        if (iterations <= 0) return;
        dev_copy(original.ensure(N), elev, sizeof(float) * (size_t)N, 2, ex().stream);
        float* src = elev; float* dst = tmp.ensure(N);
        for (int it = 0; it < iterations; it++) {
            ex().for_each(N, SharpenK{csr(), src, dst, original.p, isOcean, strength});
            std::swap(src, dst);
        }
        if (src != elev) dev_copy(elev, src, sizeof(float) * (size_t)N, 2, ex().stream);
    }

    void apply_soil_creep(float* elev, const uint8_t* isOcean, int iterations, double strength) {   // :758-794
        if (iterations <= 0) return;
        cls.ensure(N);
        ex().for_each(N, CellClassK{csr(), isOcean, cls.p});
        float* src = elev; float* dst = tmp.ensure(N);
        for (int it = 0; it < iterations; it++) {
            ex().for_each(N, CreepK{csr(), src, dst, cls.p, strength});
            std::swap(src, dst);
        }
        if (src != elev) dev_copy(elev, src, sizeof(float) * (size_t)N, 2, ex().stream);
    }

    // pass 1 on one host core (the default, pb_flood.h): (elevation, key) pairs, the visited bitmap, drainTo and the
    // seed list come down over PCIe into pinned buffers (≈ 12 B/cell), drainTo and the sparse list of filled cells go back up
    PinnedBuf<FloodEK> hfEK;
    PinnedBuf<int> hfDrain, hfSeeds, hfFilledCell;
    PinnedBuf<float> hfFilledSurf;
    PinnedBuf<uint32_t> hfBits;
    DevBuf<FloodEK> dEK;
    DevBuf<uint32_t> dBits;
    DevBuf<int> dFilledCell;
    DevBuf<float> dFilledSurf;
    std::vector<HostHeapEntry> hfHeap;
    std::vector<float> hfSurf, hfFS;
    std::vector<int> hfFC;
    double lastFloodHostMs = 0;
    void flood_heap_on_host(const float* elev) {
        const Exec& x = ex();
        const int words = (N + 31) / 32;
        x.for_each(words, FloodBitmapK{visited.p, N, dBits.ensure(words)});
        x.for_each(N, FloodPackK{elev, key.p, dEK.ensure(N)});
        hfEK.ensure(N); hfDrain.ensure(N); hfSeeds.ensure(N); hfBits.ensure(words);
        int nSeeds = 0;
        dev_copy(&nSeeds, counters.p + 0, sizeof(int), 1, x.stream);
        dev_copy(hfEK.data(), dEK.p, sizeof(FloodEK) * (size_t)N, 1, x.stream);
        dev_copy(hfDrain.data(), drainTo.p, sizeof(int) * (size_t)N, 1, x.stream);
        dev_copy(hfBits.data(), dBits.p, sizeof(uint32_t) * (size_t)words, 1, x.stream);
        dev_copy(hfSeeds.data(), seeds.p, sizeof(int) * (size_t)N, 1, x.stream);   // nSeeds is not known yet: the list is ≤ N ints
        stream_sync(x.stream);
        const auto t0 = std::chrono::steady_clock::now();
        flood_heap_host(N, hOffCopy.data(), hAdjCopy.data(), hfEK.data(), nullptr, hfDrain.data(), hfBits.data(), hfSeeds.data(), nSeeds,
                        hfHeap, hfSurf, hfFC, hfFS);
        lastFloodHostMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (getenv("PB_DEBUG")) fprintf(stderr, "[pb] flood pass 1 on the host: %d seeds, %zu filled cells, %.2f ms\n", nSeeds, hfFC.size(), lastFloodHostMs);
        const size_t nf = hfFC.size();
        dev_copy(drainTo.p, hfDrain.data(), sizeof(int) * (size_t)N, 0, x.stream);
        if (nf) {
            memcpy(hfFilledCell.ensure(nf), hfFC.data(), sizeof(int) * nf);
            memcpy(hfFilledSurf.ensure(nf), hfFS.data(), sizeof(float) * nf);
            dev_copy(dFilledCell.ensure(nf), hfFilledCell.data(), sizeof(int) * nf, 0, x.stream);
            dev_copy(dFilledSurf.ensure(nf), hfFilledSurf.data(), sizeof(float) * nf, 0, x.stream);
            x.for_each((int)nf, FloodFilledK{dFilledCell.p, dFilledSurf.p, surface.p});     // surface was initialised to the elevation by FloodInitK
        }
        stream_sync(x.stream);
    }
#if PB_CUDA
    // pass 1 on one CTA: shared-memory heap + visited bitmap (pb_flood.h)
    DevBuf<HeapEntry> heapSpill;
    DevBuf<int> liftUp, liftDepthA, liftDepthB;
    int floodSmemMax = -1;
    int subtreeGrid = -1;
    DevBuf<int> subtreeExtra;
    int lastMaxHeap = 0;
    void flood_heap_cuda(const float* elev) {
        const Exec& x = ex();
        if (floodSmemMax < 0) {
            int dev = 0, v = 0;
            PB_CUDA_CHECK(cudaGetDevice(&dev));
            PB_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            floodSmemMax = v - 2048;    // static shared + reserve
            PB_CUDA_CHECK(cudaFuncSetAttribute(k_flood_heap, cudaFuncAttributeMaxDynamicSharedMemorySize, floodSmemMax));
        }
        const int cap = ((int)(floodSmemMax / sizeof(HeapEntry)) - 4) & ~1;
        FloodHeapArgs a{csr(), elev, surface.p, drainTo.p, visited.p, key.p, cellNoiseBuf.p, seeds.p, counters.p + 0,
                        heapSpill.ensure(N), cap, counters.p + 12, getenv("PB_FLOOD_NO_PREFETCH") ? 0 : 1};
        launch_stats().launches++;
        ProfScope ps(x.prof, "pb::k_flood_heap", x.stream);
        k_flood_heap<<<1, PB_FLOOD_THREADS, floodSmemMax, x.stream>>>(a);
        PB_CUDA_CHECK(cudaGetLastError());
        if (getenv("PB_DEBUG")) {
            lastMaxHeap = read_int(counters.p + 12);
            fprintf(stderr, "[pb] flood: max heap %d (%d entries fit in shared memory)\n", lastMaxHeap, cap);
        }
    }
    // pass 2: binary lifting over the flood forest, one CTA per flood tree (pb_flood.h)
    void carve_lift_cuda(float* elev, const uint8_t* isOcean, double carveStrength) {
        const Exec& x = ex();
        int maxLevels = 1; while ((1ll << maxLevels) < (long long)N && maxLevels < 16) maxLevels++;
        liftUp.ensure((size_t)maxLevels * N); liftDepthA.ensure(N); liftDepthB.ensure(N);
        x.for_each(N, LiftInitK{isOcean, drainTo.p, liftUp.p, liftDepthA.p});
        int* dIn = liftDepthA.p; int* dOut = liftDepthB.p;
        int levels = 1;
        for (;;) {
            if (levels >= maxLevels) throw Error("flood forest deeper than 2^16 hops");
            dev_memset(counters.p + 13, 0, sizeof(int), x.stream);
            x.for_each(N, LiftStepK{liftUp.p + (size_t)(levels - 1) * N, liftUp.p + (size_t)levels * N, dIn, dOut, counters.p + 13});
            std::swap(dIn, dOut);
            const int any = read_int(counters.p + 13);
            if (!any) break;
            levels++;
        }
        const int nSeg = read_int(counters.p + 2);
        if (nSeg <= 0) return;
        CarveLiftArgs a{cells.p, segStart.p, counters.p + 2, counters.p + 1, isOcean, surface.p, elev,
                        liftUp.p, levels, N, dIn, carveStrength};
        launch_stats().launches++;
        ProfScope ps(x.prof, "pb::k_carve_lift", x.stream);
        k_carve_lift<<<nSeg, PB_CARVE_THREADS, 0, x.stream>>>(a);
        PB_CUDA_CHECK(cudaGetLastError());
    }
#endif

    // ---- priorityFloodCarve :59-215 -------------------------------------------------------------------
    void priority_flood_carve(float* elev, const uint8_t* isOcean, double carveStrength, const FloodTaps* taps) {
        const Exec& x = ex();
        const Csr g = csr();
        parent.ensure(N); ccSize.ensure(N); best.ensure(1);
        x.for_each(N, CcInitK{isOcean, parent.p, ccSize.p});
        x.for_each(N, CcHookK{g, isOcean, parent.p});
        x.for_each(N, CcFlattenCountK{isOcean, parent.p, ccSize.p});
        dev_memset(best.p, 0, sizeof(unsigned long long), x.stream);
        x.for_each(N, CcBestK{isOcean, parent.p, ccSize.p, best.p});

        surface.ensure(N); key.ensure(N); drainTo.ensure(N); visited.ensure(N); seedFlag.ensure(N);
        uint8_t* oo = (taps && taps->openOcean) ? taps->openOcean : nullptr;
        x.for_each(N, FloodInitK{g, elev, isOcean, parent.p, best.p, surface.p, key.p, drainTo.p, visited.p, seedFlag.p, oo, cellNoiseBuf.ensure(N)});
        counters.ensure(16);
        seeds.ensure(N); heap.ensure(N);
        prims.compact_flagged(x, seedFlag.p, N, seeds.p, counters.p + 0);
        if (ctx->floodOnHost) flood_heap_on_host(elev);
        else {
#if PB_CUDA
            flood_heap_cuda(elev);
#else
            x.single(FloodSerialK{g, elev, surface.p, key.p, drainTo.p, visited.p, seeds.p, counters.p + 0, heap.p});
#endif
        }
        if (taps) {
            if (taps->drainTo) dev_copy(taps->drainTo, drainTo.p, sizeof(int) * (size_t)N, 2, x.stream);
            if (taps->surface) dev_copy(taps->surface, surface.p, sizeof(float) * (size_t)N, 2, x.stream);
        }

        // pass 2
        root.ensure(N); rootActive.ensure(N); flag8.ensure(N); cells.ensure(N);
        dev_memset(rootActive.p, 0, (size_t)N, x.stream);
        x.for_each(N, FloodRootK{isOcean, drainTo.p, surface.p, elev, root.p, rootActive.p});
        x.for_each(N, FloodMemberFlagK{root.p, rootActive.p, flag8.p});
        prims.compact_flagged(x, flag8.p, N, cells.p, counters.p + 1);
        const int nCells = read_int(counters.p + 1);
        if (nCells > 0) {
            keys32.ensure(nCells); flag8b.ensure(nCells); segStart.ensure(nCells);
            x.for_each(nCells, GatherIntK{root.p, cells.p, (int*)keys32.p});
            int bits = 1; while (bits < 32 && (1ll << bits) < (long long)N) bits++;
            prims.sort_pairs(x, keys32.p, cells.p, nCells, false, bits);
            x.for_each(nCells, SegStartFlagK{(const int*)keys32.p, flag8b.p});
            prims.compact_flagged(x, flag8b.p, nCells, segStart.p, counters.p + 2);
#if PB_CUDA
            carve_lift_cuda(elev, isOcean, carveStrength);
#else
            x.for_each(nCells, CarveTreeK{cells.p, segStart.p, counters.p + 2, counters.p + 1, isOcean, drainTo.p, surface.p, elev, carveStrength});
#endif
        }

        // pass 3
        e0.ensure(N);
        dev_copy(e0.p, elev, sizeof(float) * (size_t)N, 2, x.stream);
        const int BATCH = 8;
        int* ch = counters.p + 4;
        for (long long sweeps = 0; sweeps <= (long long)N + BATCH; sweeps += BATCH) {
            dev_memset(ch, 0, sizeof(int) * BATCH, x.stream);
            for (int b = 0; b < BATCH; b++) x.for_each(N, EnforceK{isOcean, drainTo.p, surface.p, e0.p, elev, ch + b});
            int h[BATCH];
            dev_copy(h, ch, sizeof(int) * BATCH, 1, x.stream);
            stream_sync(x.stream);
            bool done = false;
            for (int b = 0; b < BATCH; b++) if (h[b] == 0) done = true;
            if (done) break;
        }
    }

    // ---- erodeComposite :369-707 ------------------------------------------------------------------------
    void sort_land_desc(const float* elev, int landCount) {
        const Exec& x = ex();
        keys32.ensure(landCount);
        x.for_each(landCount, SortKeyK{order.p, elev, keys32.p});
        prims.sort_pairs(x, keys32.p, order.p, landCount, true);
        x.for_each(landCount, PosK{order.p, pos.p});
    }

    void erode_composite(float* elev, const uint8_t* isOcean, int hIters, double K, double m, double dt,
                         int tIters, double talus, double kThermal, int gIters, double glacialStrength,
                         const ErodeTaps* taps) {
        if (glacialStrength != glacialStrength) glacialStrength = 0;
        const int totalIters = std::max(hIters, std::max(tIters, gIters));
        if (totalIters <= 0) return;
        const Exec& x = ex();
        const Csr g = csr();

        flag8.ensure(N); order.ensure(N); pos.ensure(N); counters.ensure(16);
        x.for_each(N, LandFlagK{isOcean, flag8.p});
        prims.compact_flagged(x, flag8.p, N, order.p, counters.p + 3);
        const int landCount = read_int(counters.p + 3);
        if (landCount == 0) return;
        x.for_each(N, FillIntK{pos.p, -1});
        x.for_each(landCount, PosK{order.p, pos.p});

        drainTarget.ensure(N); cellDist.ensure(N); flow.ensure(N); contrib.ensure(N); cnt.ensure(N);
        k0.ensure(N); k1.ensure(N); k2.ensure(N); tmp.ensure(N); total.ensure(N); words.ensure(N);

        if (hIters > 0) priority_flood_carve(elev, isOcean, 0.5, nullptr);

        const bool haveGlac = gIters > 0 && glacialStrength > 0;
        if (haveGlac) {
            glacIdx.ensure(N); iceTarget.ensure(N); iceFlow.ensure(N); nUp.ensure(N); kSelf.ensure(N); kEdge.ensure(E);
            x.for_each(N, GlacIdxK{xyz.p, elev, isOcean, glacIdx.p, glacialStrength});
        }
        const double gScale = gIters > 0 ? 1.0 / gIters : 0;
        const double gCarveRate = 0.02 * gScale, gConvergenceBonus = 0.01 * gScale;
        const double gDepositAmount = 0.005 * gScale, gFjordCarve = 0.015 * gScale;

        const int midFloodIter = (int)js_round(totalIters * 0.75);
        bool midFloodDone = false;

        for (int iter = 0; iter < totalIters; iter++) {
            if (!midFloodDone && iter >= midFloodIter) {
                midFloodDone = true;
                priority_flood_carve(elev, isOcean, 0.85, nullptr);
            }
            const bool glacialThisIter = iter < gIters && haveGlac;
            const bool hydraulicThisIter = iter < hIters;
            if (glacialThisIter || hydraulicThisIter) sort_land_desc(elev, landCount);

            if (glacialThisIter) {
                x.for_each(N, IceReceiversK{g, elev, isOcean, glacIdx.p, iceTarget.p});
                dev_memset(words.p, 0, sizeof(unsigned long long) * (size_t)N, x.stream);
                x.ordered(landCount, AccumulateK{g, order.p, pos.p, iceTarget.p, isOcean, glacIdx.p, words.p});
                x.for_each(N, AccumulateFinalK{g, pos.p, iceTarget.p, isOcean, glacIdx.p, words.p, iceFlow.p, nUp.p});
                dev_memset(cnt.p, 0, sizeof(int) * (size_t)N, x.stream);
                x.for_each(N, CarvePrepK{g, pos.p, isOcean, iceFlow.p, kSelf.p, kEdge.p});
                x.ordered(landCount, CarveK{g, order.p, isOcean, ndist.p, iceFlow.p, nUp.p, elev, cnt.p, kSelf.p, kEdge.p,
                                            gCarveRate, gConvergenceBonus, glacialStrength});
                x.for_each(N, MoraineFjordClampK{g, pos.p, isOcean, glacIdx.p, iceFlow.p, iceTarget.p, elev,
                                                 gDepositAmount, gFjordCarve});
            }

            if (hydraulicThisIter) {
                if (glacialThisIter) sort_land_desc(elev, landCount);
                x.for_each(N, ReceiversK{g, elev, isOcean, ndist.p, drainTarget.p, cellDist.p});
                // Planets in flight on other streams: the cooperative doubling kernel of one planet next to the backing-off
                // dataflow kernels of another was observed to hang (grids that need all their CTAs resident next to grids whose
                // lanes sleep); with more than one context alive the accumulation therefore takes the ordered dataflow form.
                const bool doubling = ctx->flowMode == 1 || (ctx->flowMode == 0 && live_contexts() <= 1 && !getenv("PB_ORDERED_FLOW"));
                if (landCount < (1 << 24) && doubling) {
                    // integer-valued, exact in f32: subtree sizes by pointer doubling (pb_erode.h)
                    int* jA = k0.p; int* jB = k1.p; int* cA = k2.p; int* cB = cnt.p;       // free until SolvePrepK
#if PB_CUDA
                    if (subtreeGrid < 0) {
                        int perSm = 0, dev = 0, coop = 0;
                        PB_CUDA_CHECK(cudaGetDevice(&dev));
                        PB_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
                        PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_subtree_counts, 256, 0));
                        subtreeGrid = (coop && perSm > 0) ? x.sm_count : 0;     // one CTA per SM: several planets in flight must be able to hold their cooperative grids side by side
                    }
                    if (subtreeGrid > 0) {
                        const int* po = order.p; const int* pp = pos.p; const int* pt = drainTarget.p; const uint8_t* pi = isOcean;
                        int nn = landCount; unsigned long long* pw = words.p; int* pa = counters.p + 14;
                        int* dC = subtreeExtra.ensure(N);
                        void* args[] = {&po, &pp, &pt, &pi, &jA, &jB, &cA, &cB, &dC, &nn, &pw, &pa};
                        launch_stats().launches++;
                        ProfScope ps(x.prof, "pb::k_subtree_counts", x.stream);
                        PB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_subtree_counts, dim3(subtreeGrid), dim3(256), args, 0, x.stream));
                    } else
#endif
                    {
                        x.for_each(landCount, SubtreeInitK{order.p, pos.p, drainTarget.p, isOcean, jA, cA});
                        int rounds = 1; while ((1ll << rounds) < (long long)landCount + 1) rounds++;
                        for (int k = 0; k < rounds; k++) {
                            x.for_each(landCount, SubtreeCopyK{order.p, cA, cB});
                            x.for_each(landCount, SubtreeRoundK{order.p, jA, jB, cA, cB});
                            std::swap(jA, jB); std::swap(cA, cB);
                        }
                        x.for_each(landCount, SubtreeWordsK{order.p, cA, words.p});
                    }
                } else {
                    dev_memset(words.p, 0, sizeof(unsigned long long) * (size_t)N, x.stream);
                    x.ordered(landCount, AccumulateK{g, order.p, pos.p, drainTarget.p, isOcean, nullptr, words.p});
                }
                x.for_each(N, AccumulateFinalK{g, pos.p, drainTarget.p, isOcean, nullptr, words.p, flow.p, nullptr});
                if (taps && iter == taps->captureIter) {
                    if (taps->drainTarget) dev_copy(taps->drainTarget, drainTarget.p, sizeof(int) * (size_t)N, 2, x.stream);
                    if (taps->flow) dev_copy(taps->flow, flow.p, sizeof(float) * (size_t)N, 2, x.stream);
                    if (taps->landOrder) {
                        x.for_each(N, FillIntK{taps->landOrder, -1});
                        dev_copy(taps->landOrder, order.p, sizeof(int) * (size_t)landCount, 2, x.stream);
                    }
                }
                x.for_each(landCount, OverRowsK<SolvePrepK>{order.p, SolvePrepK{g, pos.p, drainTarget.p, isOcean, k0.p, k1.p, k2.p}});   // land rows only
                x.for_each(N, PackElevK{elev, words.p});
                x.ordered(landCount, SolveK{order.p, landCount, drainTarget.p, isOcean, cellDist.p, flow.p, elev, words.p,
                                            k0.p, k1.p, k2.p, K, m, dt});
                x.for_each(N, UnpackElevK{words.p, isOcean, elev});
            }

            if (iter < tIters) {
                x.for_each(N, ThermalExcessK{g, elev, isOcean, ndist.p, talus, total.p});
                x.for_each(N, ThermalApplyK{g, elev, tmp.p, isOcean, ndist.p, pos.p, total.p, talus, kThermal});
                dev_copy(elev, tmp.p, sizeof(float) * (size_t)N, 2, x.stream);
            }
        }

        if (haveGlac) {
            x.for_each(N, GlacialBlendK{g, elev, tmp.p, isOcean, glacIdx.p});
            dev_copy(elev, tmp.p, sizeof(float) * (size_t)N, 2, x.stream);
        }
    }

    // ---- runPostProcessing js/planet-worker.js:40-102 -------------------------------------------------------
    void run_post_processing(float* elev, const pb_post_params& p, double seed, const float* hotspot,
                             float* erosionDelta, uint8_t* isOceanOut) {
        const Exec& x = ex();
        timer.mark(0, x.stream);
        if (p.terrainWarp > 0) warp_terrain(elev, seed, p.terrainWarp, hotspot);
        timer.mark(1, x.stream);
        uint8_t* isOcean = isOceanOut ? isOceanOut : isOceanBuf.ensure(N);
        x.for_each(N, IsOceanK{elev, isOcean});
        dev_copy(preErosion.ensure(N), elev, sizeof(float) * (size_t)N, 2, x.stream);
        if (p.smoothing > 0)
            smooth_elevation(elev, isOcean, (int)js_round(1 + p.smoothing * 4), 0.2 + p.smoothing * 0.5);
        timer.mark(2, x.stream);
        if (p.glacialErosion > 0 || p.hydraulicErosion > 0 || p.thermalErosion > 0) {
            const int gIters = (int)js_round(p.glacialErosion * 10);
            const int hIters = p.hItersOverride > 0 ? p.hItersOverride : (int)js_round(p.hydraulicErosion * 20);
            erode_composite(elev, isOcean, hIters, p.hydraulicErosion * 0.0006, 0.5, 1.0,
                            (int)js_round(p.thermalErosion * 10), 1.2 - p.thermalErosion * 0.4,
                            p.thermalErosion * 0.15, gIters, p.glacialErosion, nullptr);
        }
        timer.mark(3, x.stream);
        if (p.ridgeSharpening > 0)
            sharpen_ridges(elev, isOcean, (int)js_round(1 + p.ridgeSharpening * 3), p.ridgeSharpening * 0.08);
        timer.mark(4, x.stream);
        apply_soil_creep(elev, isOcean, 3, 0.1125);
        if (erosionDelta) x.for_each(N, SubF32K{elev, preErosion.p, erosionDelta});
        timer.mark(5, x.stream);
        timer.pending = true;
    }
};

}  // namespace pb
