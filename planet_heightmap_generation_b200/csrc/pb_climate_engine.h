// pb_climate_engine.h — host orchestration of the climate stack in the reference's order
// (js/planet-worker.js:229-268 / handleComputeClimate :579-672): computeWind → computeOceanCurrents →
// computePrecipitation → computeTemperature → classifyKoppen.  The result objects of the reference
// (windResult, oceanResult, precipResult, tempResult — the worker's W.cachedWind / W.cachedOcean) stay
// resident in HBM under their reference key names; callers read individual fields back by name.
#pragma once
#include "pb_engine.h"
#include "pb_shardsweep.h"
#include "pb_climate.h"

namespace pb {

struct Climate {
    Mesh* m;
    int N;
    std::map<std::string, DevBuf<float>> f;
    std::map<std::string, DevBuf<int>> i;
    std::map<std::string, DevBuf<uint8_t>> u;
    std::map<std::string, size_t> count;     // elements of every published field
    std::map<std::string, int> kind;         // 0 f32, 1 i32, 2 u8
    bool haveWind = false, haveOcean = false, havePrecip = false, haveTemp = false;
    int bfsGrid = -1;

    // scratch
    DevBuf<float> sElev, a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, upWt, dnWt;
    DevBuf<int> sPlate, plateIds, frontA, frontB, binCell, binOffset, cnt;
    DevBuf<uint32_t> keys;
    DevBuf<uint8_t> plateTable, contPlate, flagA, flagB, flagC, noiseTab;
    DevBuf<double> samples, scalars;
    DevBuf<SplineDev> splines;
    DevBuf<int> landRow, landIndex; DevBuf<PackedRow> landPack; DevBuf<float> landWt; int nLand = 0, landLo = 0, landHi = 0;

    explicit Climate(Mesh* mesh) : m(mesh), N(mesh->N) {}

    const Exec& ex() const { return m->ex(); }
    Csr csr() const { return m->csr(); }
    float* F(const std::string& name, size_t n = 0) { n = n ? n : (size_t)N; count[name] = n; kind[name] = 0; return f[name].ensure(n); }
    int* I(const std::string& name) { count[name] = N; kind[name] = 1; return i[name].ensure(N); }
    uint8_t* U(const std::string& name) { count[name] = N; kind[name] = 2; return u[name].ensure(N); }
    const float* cF(const std::string& name) { auto it = f.find(name); if (it == f.end() || !it->second.p) throw std::invalid_argument("climate field not computed yet: " + name); return it->second.p; }
    const int* cI(const std::string& name) { auto it = i.find(name); if (it == i.end() || !it->second.p) throw std::invalid_argument("climate field not computed yet: " + name); return it->second.p; }
    const uint8_t* cU(const std::string& name) { auto it = u.find(name); if (it == u.end() || !it->second.p) throw std::invalid_argument("climate field not computed yet: " + name); return it->second.p; }

    static int js_round_i(double x) { return (int)floor(x + 0.5); }
    double avgEdgeKm() const { return (PB_PI * 6371) / sqrt((double)N); }

    // ---- shared building blocks -------------------------------------------------------------------------
    void smooth_masked(float* field, const uint8_t* mask, int passes, bool zeroOutside) {
        const Csr g = csr(); const int z = zeroOutside ? 1 : 0;
        sweep_loop(*m, field, passes, m->tmp.ensure(N), [=](const float* src, float* dst) { return SmoothMaskedK{g, mask, src, dst, z}; });
    }

    // hop counts from the cells flagged in seedFlag (their dist is already 0, everything else -1) for up to three independent
    // BFS at once
    DevBuf<int> bfsFront[3][2], bfsCnt;
    void bfs_many(int k, int* const* dist, const uint8_t* const* passable, const uint8_t* const* seedFlag) {
        const Exec& x = ex();
        bfsCnt.ensure(12);
        dev_memset(bfsCnt.p, 0, 12 * sizeof(int), x.stream);
        for (int b = 0; b < k; b++) {
            bfsFront[b][0].ensure(N); bfsFront[b][1].ensure(N);
            m->prims.compact_flagged(x, seedFlag[b], N, bfsFront[b][0].p, bfsCnt.p + 4 * b);
        }
#if PB_CUDA
        if (bfsGrid < 0) {
            int perSm = 0, dev = 0, coop = 0;
            PB_CUDA_CHECK(cudaGetDevice(&dev));
            PB_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
            PB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_bfs_persistent, 256, 0));
            bfsGrid = (coop && perSm > 0) ? x.sm_count : 0;      // one CTA per SM: frontiers are small, the grid barrier is the cost
        }
        if (bfsGrid > 0) {
            Csr g = csr();
            BfsMulti B{};
            B.k = k;
            for (int b = 0; b < k; b++) { B.passable[b] = passable[b]; B.dist[b] = dist[b]; B.fa[b] = bfsFront[b][0].p; B.fb[b] = bfsFront[b][1].p; B.cnt[b] = bfsCnt.p + 4 * b; }
            void* args[] = {&g, &B};
            launch_stats().launches++;
            ProfScope ps(x.prof, "pb::k_bfs_persistent", x.stream);
            PB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_bfs_persistent, dim3(bfsGrid), dim3(256), args, 0, x.stream));
            return;
        }
#endif
        for (int b = 0; b < k; b++) {
            int* cur = bfsFront[b][0].p; int* nxt = bfsFront[b][1].p; int* c = bfsCnt.p + 4 * b;
            const int CHECK = 16;
            for (int level = 0;; level++) {
                x.for_each_dev(c + level % 3, BfsLevelK{csr(), passable[b], dist[b], cur, nxt, c, level});
                std::swap(cur, nxt);
                if ((level + 1) % CHECK == 0) {
                    int h[3];
                    dev_copy(h, c, sizeof h, 1, x.stream);
                    stream_sync(x.stream);
                    if (h[(level + 1) % 3] == 0) break;
                }
                if (level > N) throw Error("bfs did not terminate");
            }
        }
    }
    void bfs(int* dist, const uint8_t* passable, const uint8_t* seedFlag) { bfs_many(1, &dist, &passable, &seedFlag); }

    // *out = percentile of the values whose order-preserving keys are in keys[0..n) (keys are consumed)
    void percentile(uint32_t* k, int n, const int* nDev, double p, double* out) {
        m->prims.sort_keys(ex(), k, n);
        ex().for_each(1, PercentilePickK{k, n, nDev, p, out});
    }

    // ---- computeWind  js/wind.js:394-687 ----------------------------------------------------------------------
    void compute_wind(const float* elev, const int* plateIsOceanIdsHost, int nIds, const int* r_plate, double noiseSeed, double axialTilt) {
        (void)axialTilt;   // the reference computes tiltRad and never uses it
        const Exec& x = ex();
        const Csr g = csr();
        const double aek = avgEdgeKm();
        float *lat = F("r_lat"), *lon = F("r_lon"), *sinLat = F("r_sinLat"), *cosLat = F("r_cosLat");
        uint8_t *isLand = U("r_isLand"), *isOcean = U("r_isOcean");
        float *eX = F("r_eastX"), *eY = F("r_eastY"), *eZ = F("r_eastZ"), *nX = F("r_northX"), *nY = F("r_northY"), *nZ = F("r_northZ");
        x.for_each(N, WindPrecomputeK{m->xyz.p, elev, lat, lon, sinLat, cosLat, isLand, isOcean, eX, eY, eZ, nX, nY, nZ});

        // geo index + ITCZ (:88-232)
        keys.ensure(N); binCell.ensure(N);
        const int numBins = PB_LAT_BINS * PB_LON_BINS;
        binOffset.ensure(numBins + 1);
        x.for_each(N, GeoBinK{lat, lon, keys.p, binCell.p});
        m->prims.sort_pairs(x, keys.p, binCell.p, N, false, 12);
        x.for_each(numBins + 1, BinOffsetK{keys.p, N, binOffset.p});
        const int nSamples = 2 * PB_ITCZ_NLON * 4;
        samples.ensure(2 * nSamples); splines.ensure(2);
        ItczSampleArgs sa{binOffset.p, binCell.p, lon, sinLat, cosLat, elev, isLand, samples.p};
#if PB_CUDA
        {
            launch_stats().launches++;
            ProfScope ps(x.prof, "pb::k_itcz_sample", x.stream);
            k_itcz_sample<<<nSamples, PB_ITCZ_THREADS, 0, x.stream>>>(sa);
            PB_CUDA_CHECK(cudaGetLastError());
        }
#else
        x.for_each(nSamples, ItczSampleSerialK{sa});
#endif
        x.for_each(2, ItczFinishK{samples.p, splines.p});
        float *itczLons = F("itczLons", PB_ITCZ_SAMPLES), *ls = F("itczLatsSummer", PB_ITCZ_SAMPLES), *lw = F("itczLatsWinter", PB_ITCZ_SAMPLES);
        x.for_each(PB_ITCZ_SAMPLES, ItczTableK{splines.p, itczLons, ls, lw});

        // main ocean = largest component of non-land cells (:481-508); coast distance through land (:510-538)
        m->parent.ensure(N); m->ccSize.ensure(N); m->best.ensure(1);
        x.for_each(N, CcInitK{isOcean, m->parent.p, m->ccSize.p});
        x.for_each(N, CcHookK{g, isOcean, m->parent.p});
        x.for_each(N, CcFlattenCountK{isOcean, m->parent.p, m->ccSize.p});
        dev_memset(m->best.p, 0, sizeof(unsigned long long), x.stream);
        x.for_each(N, CcBestK{isOcean, m->parent.p, m->ccSize.p, m->best.p});
        int* coastDist = I("r_coastDistLand");
        flagA.ensure(N); flagB.ensure(N);
        x.for_each(N, LandCoastSeedK{g, isLand, isOcean, m->parent.p, m->best.p, coastDist, flagA.p});

        // plate-based continentality (:556-593): its BFS shares the launch (and the grid barriers) with the coast BFS
        int maxId = 0;
        for (int k = 0; k < nIds; k++) { if (plateIsOceanIdsHost[k] < 0) throw std::invalid_argument("negative plate id"); maxId = std::max(maxId, plateIsOceanIdsHost[k]); }
        const int tableSize = maxId + 1;
        plateTable.ensure(tableSize);
        dev_memset(plateTable.p, 0, (size_t)tableSize, x.stream);
        if (nIds > 0) {
            dev_copy(plateIds.ensure(nIds), plateIsOceanIdsHost, sizeof(int) * (size_t)nIds, 0, x.stream);
            x.for_each(nIds, PlateTableK{plateIds.p, plateTable.p});
        }
        contPlate.ensure(N);
        x.for_each(N, ContPlateK{r_plate, plateTable.p, tableSize, contPlate.p});
        int* plateDist = I("r_plateDist");
        x.for_each(N, MaskBoundarySeedK{g, contPlate.p, plateDist, flagB.p});
        {
            int* dists[2] = {coastDist, plateDist};
            const uint8_t* pass[2] = {isLand, contPlate.p};
            const uint8_t* seedsF[2] = {flagA.p, flagB.p};
            bfs_many(2, dists, pass, seedsF);
        }
        float* cont = F("r_continentality");
        x.for_each(N, ContinentalityK{isLand, coastDist, cont, aek});
        const int contSmoothPasses = std::max(1, js_round_i(100 / aek));
        m->smooth_field(cont, contSmoothPasses);
        float* pcont = F("r_plateContinentality");
        x.for_each(N, ContinentalityK{contPlate.p, plateDist, pcont, aek});
        m->smooth_field(pcont, contSmoothPasses);

        // seasons (:600-653)
        SimplexTable tab(noiseSeed);
        dev_copy(noiseTab.ensure(1024), tab.t, 1024, 0, x.stream);
        float *pressure = a0.ensure(N), *gradE = a1.ensure(N), *gradN = a2.ensure(N);
        scalars.ensure(8);
        const int pressSmoothPasses = std::max(1, js_round_i(75 / aek));
        for (int s = 0; s < 2; s++) {
            const std::string name = s == 0 ? "summer" : "winter";
            x.for_each(N, PressureK{lat, lon, splines.p + s, s == 0 ? 1 : 0, cont, elev, Simplex{noiseTab.p}, m->xyz.p, pressure});
            m->smooth_field(pressure, pressSmoothPasses);
            x.for_each(N, GradientsK{g, m->xyz.p, pressure, eX, eY, eZ, nX, nY, nZ, gradE, gradN});
            float *windE = F("r_wind_east_" + name), *windN = F("r_wind_north_" + name), *speed = F("r_wind_speed_" + name);
            x.for_each(N, PressureToWindK{gradE, gradN, sinLat, windE, windN, speed, keys.p});
            percentile(keys.p, N, nullptr, 0.95, scalars.p + s);
            x.for_each(N, NormalizeMin1K{speed, scalars.p + s});
            x.for_each(N, PressureDevK{pressure, F("r_pressure_" + name)});
        }
        haveWind = true; haveOcean = havePrecip = haveTemp = false;
    }

    // ---- computeOceanCurrents  js/ocean.js:204-382 ------------------------------------------------------------
    void compute_ocean_currents(const float* elev) {
        (void)elev;
        if (!haveWind) throw std::invalid_argument("computeOceanCurrents needs computeWind's result");
        const Exec& x = ex();
        const Csr g = csr();
        const double aek = avgEdgeKm();
        const float *lat = cF("r_lat"), *lon = cF("r_lon");
        const uint8_t* isOcean = cU("r_isOcean");
        int *coast = I("r_oceanCoastDist"), *west = I("r_westCoastDist"), *east = I("r_eastCoastDist");
        flagA.ensure(N); flagB.ensure(N); flagC.ensure(N);
        x.for_each(N, OceanCoastSeedK{g, m->xyz.p, isOcean, cF("r_eastX"), cF("r_eastY"), cF("r_eastZ"), coast, west, east, flagA.p, flagB.p, flagC.p});
        {
            int* dists[3] = {coast, west, east};
            const uint8_t* pass[3] = {isOcean, isOcean, isOcean};
            const uint8_t* seedsF[3] = {flagA.p, flagB.p, flagC.p};
            bfs_many(3, dists, pass, seedsF);
        }
        cnt.ensure(160);
        int* bins = cnt.p + 4; int* circ = cnt.p + 148; int* oceanCount = cnt.p + 150;
        dev_memset(bins, 0, sizeof(int) * 148, x.stream);
        x.for_each(N, CircumpolarBinsK{lat, lon, isOcean, bins});
        x.for_each(2, CircumpolarFlagK{bins, circ});
        const double coastThreshold = std::max(5.0, floor(sqrt((double)N) * 0.035 + 0.5));
        const double warmthRange = coastThreshold * 2;
        const int oceanSmoothPasses = std::max(2, js_round_i(125 / aek));
        const int warmthSmoothPasses = std::max(3, js_round_i(900 / aek));
        keys.ensure(N); scalars.ensure(8);
        for (int s = 0; s < 2; s++) {
            const std::string name = s == 0 ? "summer" : "winter";
            const double shift = s == 0 ? 5 : -5;
            const float* itczLats = cF(s == 0 ? "itczLatsSummer" : "itczLatsWinter");
            float *curE = F("r_ocean_current_east_" + name), *curN = F("r_ocean_current_north_" + name);
            x.for_each(N, OceanCurrentsK{lat, lon, isOcean, itczLats, west, east, circ, shift, coastThreshold, curE, curN});
            smooth_masked(curE, isOcean, oceanSmoothPasses, false);
            smooth_masked(curN, isOcean, oceanSmoothPasses, false);
            x.for_each(N, ZeroOutsideK{isOcean, curE, curN});
            float* warmth = F("r_ocean_warmth_" + name);
            x.for_each(N, WarmthK{isOcean, lat, west, east, warmthRange, shift, warmth});
            smooth_masked(warmth, isOcean, warmthSmoothPasses, false);
            float* speed = F("r_ocean_speed_" + name);
            dev_memset(oceanCount + s, 0, sizeof(int), x.stream);
            x.for_each(N, OceanSpeedK{curE, curN, isOcean, speed, keys.p, oceanCount + s});
            percentile(keys.p, N, oceanCount + s, 0.95, scalars.p + 2 + s);
            x.for_each(N, NormalizeMin1K{speed, scalars.p + 2 + s});
        }
        haveOcean = true; havePrecip = haveTemp = false;
    }

    // ---- computePrecipitation  js/precipitation.js:196-684 (+ js/heuristic-precip.js) --------------------------------
    void compute_precipitation(const float* elev, double precipitationOffset, double landCoverage) {
        if (!haveWind || !haveOcean) throw std::invalid_argument("computePrecipitation needs the wind and ocean results");
        const Exec& x = ex();
        const Csr g = csr();
        const double aek = avgEdgeKm();
        const double avgEdgeRad = PB_PI / sqrt((double)N);
        const int maxHops = (int)std::max(8.0, std::min(20.0, floor(2000 / aek + 0.5)));
        const float *lat = cF("r_lat"), *lon = cF("r_lon"), *cont = cF("r_continentality");
        const uint8_t* isLand = cU("r_isLand");
        const float *eX = cF("r_eastX"), *eY = cF("r_eastY"), *eZ = cF("r_eastZ"), *nX = cF("r_northX"), *nY = cF("r_northY"), *nZ = cF("r_northZ");
        const int* coastDistLand = cI("r_coastDistLand");

        const int elevSmoothPasses = std::max(2, js_round_i(200 / aek));
        float* elevSmoothed = a0.ensure(N);
        dev_copy(elevSmoothed, elev, sizeof(float) * (size_t)N, 2, x.stream);
        m->smooth_field(elevSmoothed, elevSmoothPasses);
        x.for_each(N, BlendElevK{elevSmoothed, elev});
        float *gradE = F("r_elevGradE"), *gradN = F("r_elevGradN");
        x.for_each(N, GradientsK{g, m->xyz.p, elevSmoothed, eX, eY, eZ, nX, nY, nZ, gradE, gradN});
        float* heightKm = a1.ensure(N);
        x.for_each(N, HeightKmK{elev, heightKm});

        float *windE = a2.ensure(N), *windN = a3.ensure(N), *wX = a4.ensure(N), *wY = a5.ensure(N), *wZ = a6.ensure(N);
        float *conv = a7.ensure(N), *bufA = a8.ensure(N), *bufB = a9.ensure(N);
        upWt.ensure(m->E); dnWt.ensure(m->E);
        keys.ensure(N); scalars.ensure(8);
        const int convSmoothPasses = std::max(3, js_round_i(400 / aek));
        const int shadowHops = std::max(8, js_round_i(2500 / aek));
        const int windwardHops = std::max(6, js_round_i(1500 / aek));
        const int rsSmoothPasses = std::max(2, js_round_i(150 / aek));
        const int precipSmoothPasses = std::max(1, js_round_i(100 / aek));
        const double depletionBase = 1 - pb_pow(0.78, 1.0 / maxHops);
        const double shadowDecay = 1 - pb_pow(0.15, 1.0 / shadowHops);
        const double windwardDecay = 1 - pb_pow(0.25, 1.0 / windwardHops);

        // compacted land rows for the rain-shadow / windward propagation (one 16-byte packed word + eight weights per land row)
        const bool landOk = g.pack != nullptr && !getenv("PB_NO_LAND_COMPACT");
        if (landOk) {
            cnt.ensure(160);
            m->prims.compact_flagged(x, isLand, N, landRow.ensure(N), cnt.p + 152);
            nLand = m->read_int(cnt.p + 152);
            if (nLand > 0) {
                x.for_each(nLand, LandPackK{landRow.p, g.pack, landPack.ensure(nLand)});
                x.for_each(N, FillIntK{landIndex.ensure(N), -1});
                x.for_each(nLand, LandIndexK{landRow.p, landIndex.p});
                landLo = 0; landHi = nLand;
                if (sweeps_sharded(*m)) {       // the items of this rank's cell-id range
                    x.for_each(1, LandRangeK{landRow.p, nLand, m->shards->lo, m->shards->hi, cnt.p + 153});
                    landLo = m->read_int(cnt.p + 153); landHi = m->read_int(cnt.p + 154);
                }
            }
        }
        for (int s = 0; s < 2; s++) {
            const std::string name = s == 0 ? "summer" : "winter";
            const float* itczLats = cF(s == 0 ? "itczLatsSummer" : "itczLatsWinter");
            x.for_each(N, WindBlendK{lat, lon, itczLats, cF("r_wind_east_" + name), cF("r_wind_north_" + name), eX, eY, eZ, nX, nY, nZ,
                                     windE, windN, wX, wY, wZ});
            x.for_each(N, ConvergenceK{g, m->xyz.p, wX, wY, wZ, conv});
            m->smooth_field(conv, convSmoothPasses);
            // advectMoisture (:59-182)
            float* src = bufA;
            x.for_each(N, MoistureInitK{g, m->xyz.p, isLand, coastDistLand, cF("r_ocean_warmth_" + name), wX, wY, wZ, src});
            {
                const float* xyzp = m->xyz.p;
                auto make = [=](const float* in, float* out) { return AdvectK{g, xyzp, isLand, windE, windN, wX, wY, wZ, heightKm, in, out, depletionBase, maxHops}; };
                if (landOk && nLand > 0) {
                    // ocean cells keep their initial moisture (:122): only the land rows are swept, 32 working lanes per warp
                    dev_copy(bufB, src, sizeof(float) * (size_t)N, 2, x.stream);
                    const int* lr = landRow.p;
                    sweep_loop_items(*m, nLand, landLo, landHi, src, maxHops, bufB, [=](const float* in, float* out) { return OverRowsK<AdvectK>{lr, make(in, out)}; });
                } else
                    sweep_loop(*m, src, maxHops, bufB, make);
            }
            float* precip = F("r_precip_complex_" + name);
            float* rainShadow = F("r_rainshadow_" + name);
            PrecipParams P{s == 0 ? 1 : 0, maxHops, aek, avgEdgeRad, precipitationOffset, landCoverage};
            x.for_each(N, PrecipMechanismsK{lat, lon, elev, isLand, cont, itczLats, src, conv, windE, windN, gradE, gradN,
                                            cF("r_pressure_" + name), coastDistLand, heightKm, P, precip, rainShadow});
            // rain shadow: propagate downwind / upwind over the wind-aligned edge lists (:515-606)
            x.for_each(N, EdgeWeightsK{g, m->xyz.p, isLand, wX, wY, wZ, upWt.p, dnWt.p});
            float* shadowField = bufA; float* windwardField = bufB;      // moisture is no longer needed
            float* ping = conv; float* pong = m->tmp.ensure(N);            // nor is the convergence field
            dev_copy(shadowField, rainShadow, sizeof(float) * (size_t)N, 2, x.stream);
            dev_copy(windwardField, rainShadow, sizeof(float) * (size_t)N, 2, x.stream);
            const bool compact = landOk && nLand > 0;      // land-compacted records (pb_climate.h: ShadowLandK)
            for (int dir = 0; dir < 2; dir++) {
                const float* wt = dir == 0 ? upWt.p : dnWt.p;
                const double keep = dir == 0 ? 1 - shadowDecay : 1 - windwardDecay;
                const int sign = dir == 0 ? -1 : +1, hops = dir == 0 ? shadowHops : windwardHops;
                dev_copy(ping, rainShadow, sizeof(float) * (size_t)N, 2, x.stream);
                if (compact) {
                    dev_copy(pong, rainShadow, sizeof(float) * (size_t)N, 2, x.stream);     // ocean rows are never written: both buffers hold them
                    float* lw = landWt.ensure((size_t)PB_ROW_FAST * nLand);
                    x.for_each(nLand, LandWeightsK{g, landRow.p, wt, lw});
                    const int* lr = landRow.p; const PackedRow* lp = landPack.p; const int* li = landIndex.p;
                    sweep_loop_items(*m, nLand, landLo, landHi, ping, hops, pong, [=](const float* in, float* out) { return ShadowLandK{g, lr, lp, lw, wt, isLand, in, out, keep, sign, li}; });
                } else
                    sweep_loop(*m, ping, hops, pong, [=](const float* in, float* out) { return ShadowSweepK{g, isLand, wt, in, out, keep, sign}; });
                x.for_each(N, KeepExtremeK{ping, dir == 0 ? shadowField : windwardField, sign});
            }
            x.for_each(N, MergeShadowK{shadowField, windwardField, rainShadow});
            m->smooth_field(rainShadow, rsSmoothPasses);
            x.for_each(N, ApplyShadowK{isLand, rainShadow, precip});
            m->smooth_field(precip, precipSmoothPasses);
        }

        // heuristic model (heuristic-precip.js:119-269), blend, p95 normalise, continental cap (:644-679)
        float* westCoast = F("r_westCoast");
        x.for_each(N, WestCoastSeedK{g, m->xyz.p, isLand, coastDistLand, eX, eY, eZ, westCoast});
        smooth_masked(westCoast, isLand, std::max(2, js_round_i(300 / aek)), true);
        for (int s = 0; s < 2; s++) {
            const std::string name = s == 0 ? "summer" : "winter";
            const float* itczLats = cF(s == 0 ? "itczLatsSummer" : "itczLatsWinter");
            float* heur = F("r_precip_heuristic_" + name);
            x.for_each(N, HeuristicPrecipK{lat, lon, isLand, cont, elev, itczLats, westCoast, gradE, gradN, coastDistLand, s == 0 ? 1 : 0, aek, heur});
            m->smooth_field(heur, precipSmoothPasses);
            float* blended = F("r_precip_" + name);
            x.for_each(N, BlendPrecipK{cF("r_precip_complex_" + name), heur, blended, keys.p});
            percentile(keys.p, N, nullptr, 0.95, scalars.p + 4 + s);
            x.for_each(N, NormalizeCapK{blended, scalars.p + 4 + s, isLand, cont});
        }
        havePrecip = true; haveTemp = false;
    }

    // ---- computeTemperature  js/temperature.js:69-237 --------------------------------------------------------------------
    void compute_temperature(const float* elev, double temperatureOffset) {
        if (!havePrecip) throw std::invalid_argument("computeTemperature needs the precipitation result");
        const Exec& x = ex();
        const Csr g = csr();
        const int passes = std::max(4, js_round_i(1400 / avgEdgeKm()));
        const float* pcont = cF("r_plateContinentality");
        for (int s = 0; s < 2; s++) {
            const std::string name = s == 0 ? "summer" : "winter";
            const float* warmth = cF("r_ocean_warmth_" + name);
            float* src = a0.ensure(N);
            x.for_each(N, CoastalSeedK{cU("r_isLand"), warmth, src});
            sweep_loop(*m, src, passes, a1.ensure(N), [=](const float* in, float* out) { return DiffuseWarmthK{g, pcont, in, out}; });
            float* temp = F("r_temperature_" + name);
            x.for_each(N, TemperatureK{cF("r_lat"), cF("r_lon"), cU("r_isLand"), elev, cF("r_continentality"), pcont,
                                       cF(s == 0 ? "itczLatsSummer" : "itczLatsWinter"), warmth, cF("r_ocean_speed_" + name),
                                       cF("r_precip_" + name), src, s == 0 ? 1 : 0, temperatureOffset, temp});
            m->smooth_field(temp, 1);
            x.for_each(N, TempNormalizeK{temp});
        }
        haveTemp = true;
    }

    // ---- classifyKoppen  js/koppen.js:67-288 -----------------------------------------------------------------------------------
    void classify_koppen(const float* elev) {
        if (!haveTemp) throw std::invalid_argument("classifyKoppen needs the temperature result");
        ex().for_each(N, KoppenK{elev, cF("r_temperature_summer"), cF("r_temperature_winter"), cF("r_precip_summer"), cF("r_precip_winter"), U("r_koppen")});
    }
};

}  // namespace pb
