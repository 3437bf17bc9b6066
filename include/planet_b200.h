/*
 * planet_b200.h — C ABI of the B200-native planet-heightmap engine.
 *
 * This is the drop-in boundary for the per-Voronoi-cell hot path of
 * raguilar011095/planet_heightmap_generation.  The reference has no native layer: its hot path is
 * a set of plain exported JavaScript functions over typed arrays, called from the Web-Worker
 * command layer (js/planet-worker.js).  Each entry point below replaces one of those functions
 * and cites it.  A Node N-API addon / ctypes stub binds them 1:1.
 *
 * Conventions
 *  - Every function returns pb_status (0 = PB_OK).  On failure pb_last_error() returns a message
 *    for the calling thread (the reference's handlers turn exceptions into
 *    {type:'error', message}; js/planet-worker.js:336-338).
 *  - A pb_context owns one CUDA device + stream.  Calls on one context are synchronous with respect
 *    to the caller in host-pointer mode and are NOT re-entrant (the reference runs one worker,
 *    one command at a time; js/planet-worker.js:944-954).
 *  - Per-cell array arguments follow the context's pointer mode (like cuBLAS pointer mode):
 *      PB_POINTER_HOST   (default) host arrays; the call copies in, computes on the GPU, copies out
 *                        and returns when the result is in the caller's buffer.
 *      PB_POINTER_DEVICE device arrays already resident in HBM; the call only enqueues work on the
 *                        context's stream (results are ordered on that stream).
 *    Mesh construction always takes host arrays.
 *  - `r_elevation` is mutated in place, exactly like the reference functions do.
 *  - `seed` is a double: the reference passes non-integer seeds (js/plates.js:9).
 *  - Layouts are the reference's: adjOffset int32[N+1], adjList int32[E], r_xyz float32[3N] (AoS),
 *    per-cell fields float32[N] / int32[N] / uint8[N], per-edge fields float32[E].
 */
#ifndef PLANET_B200_H
#define PLANET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int pb_status;
#define PB_OK 0
#define PB_ERR_INVALID 1
#define PB_ERR_CUDA 2
#define PB_ERR_INTERNAL 3

#define PB_POINTER_HOST 0
#define PB_POINTER_DEVICE 1

typedef struct pb_context pb_context;
typedef struct pb_mesh pb_mesh;

/* ---- context ------------------------------------------------------------------------------- */
pb_status pb_context_create(int device, pb_context** out);
void pb_context_destroy(pb_context* ctx);
const char* pb_last_error(void);
const char* pb_version(void);
/* cuda_stream is a cudaStream_t (NULL = default stream). */
pb_status pb_set_stream(pb_context* ctx, void* cuda_stream);
pb_status pb_set_pointer_mode(pb_context* ctx, int mode);
pb_status pb_synchronize(pb_context* ctx);
/* Engine options (name, value).  "flood": where pass 1 of priorityFloodCarve (js/terrain-post.js:131-147) runs —
 * one serial chain of heap pops whose tie order depends on the heap layout.  "host" (default): on a host core, like
 * the other host-serial stages (assignDistanceField); "device": the one-CTA CUDA kernel k_flood_heap (≈ 8x slower:
 * a single warp retires the dependent chain at ~0.1 instructions per cycle).  Everything else of the function stays
 * on the GPU either way and the results are identical.
 * "flow": how the hydraulic flow accumulation (js/terrain-post.js:604-611) runs — "doubling": subtree sizes by pointer
 * doubling (the additions are integer-valued below 2^24 land cells, hence exact in any order), "ordered": the sync-free
 * dataflow in the reference's order (always used from 2^24 land cells up), "auto" (default): doubling, except while
 * several contexts are alive in the process (planets in flight on one GPU), where the ordered form is used.
 * "mesh_order": "canonical" (default) — pb_triangulate_sphere / pb_mesh_create_from_points / the coarse mesh of
 * pb_generate_coarse_plates use the device mesh builder, whose neighbour rows start at a canonical triangle; "delaunator" —
 * they use the reference's own row starts (delaunator@5.0.1's triangle numbering, serial host algorithm, see
 * pb_mesh_create_delaunator): same seed, same planet as the web app. */
pb_status pb_set_option(pb_context* ctx, const char* name, const char* value);
/* kernels launched by this library on any context since process start (bench: gpu_launches) */
int64_t pb_launch_count(void);

/* Per-kernel device timing (the reference's counterpart is its performance.now() stage timers,
 * js/planet-worker.js:42-101).  Between start and stop every launch of this library whose kernel
 * name contains `filter` (NULL = all) is bracketed by CUDA events on the context's stream; stop
 * synchronises and writes a JSON array [{"name":..,"launches":..,"ms":..}] into out[cap]. */
pb_status pb_profile_start(pb_context* ctx, const char* filter);
pb_status pb_profile_stop(pb_context* ctx, char* out, int64_t cap);

/* ---- mesh ------------------------------------------------------------------------------------
 * Replaces the `mesh` object the hot-path functions receive: {numRegions, adjOffset, adjList}
 * (js/sphere-mesh.js:94-146) plus r_xyz (js/sphere-mesh.js:174-186).  neighborDist
 * (computeNeighborDist, js/sphere-mesh.js:191-203) is computed on the device at creation and kept
 * with the mesh; the reference passes it to erodeComposite as a separate argument. */
pb_status pb_mesh_create(pb_context* ctx, int32_t numRegions, const int32_t* adjOffset,
                         const int32_t* adjList, const float* r_xyz, pb_mesh** out);
void pb_mesh_destroy(pb_mesh* mesh);
int32_t pb_mesh_num_regions(const pb_mesh* mesh);
int64_t pb_mesh_num_edges(const pb_mesh* mesh);
/* computeNeighborDist(mesh, r_xyz) → Float32Array[E]            js/sphere-mesh.js:191-203 */
pb_status pb_compute_neighbor_dist(pb_mesh* mesh, float* neighborDist_out);

/* ---- terrain-post (js/terrain-post.js) -------------------------------------------------------- */
/* warpTerrain(mesh, r_elevation, r_xyz, seed, strength, r_hotspot)      js/terrain-post.js:233-309
 * r_hotspot may be NULL (handleReapply passes none, js/planet-worker.js:358). */
pb_status pb_warp_terrain(pb_mesh* mesh, float* r_elevation, double seed, double strength,
                          const float* r_hotspot);
/* smoothElevation(mesh, r_elevation, r_isOcean, iterations, strength)  js/terrain-post.js:317-354 */
pb_status pb_smooth_elevation(pb_mesh* mesh, float* r_elevation, const uint8_t* r_isOcean,
                              int32_t iterations, double strength);
/* priorityFloodCarve(mesh, r_elevation, r_isOcean, carveStrength)      js/terrain-post.js:59-215
 * (module-private in the reference; exported here because erodeComposite calls it twice and the
 * parity tests tap it).  Optional outputs (NULL to skip): drainTo int32[N], surface float32[N],
 * isOpenOcean uint8[N] — the state after pass 1. */
pb_status pb_priority_flood_carve(pb_mesh* mesh, float* r_elevation, const uint8_t* r_isOcean,
                                  double carveStrength, int32_t* drainTo_out, float* surface_out,
                                  uint8_t* isOpenOcean_out);
/* erodeComposite(mesh, r_elevation, r_xyz, r_isOcean, hIters, K, m, dt, tIters, talusSlope,
 *                kThermal, gIters, glacialStrength, neighborDist)       js/terrain-post.js:369-707 */
pb_status pb_erode_composite(pb_mesh* mesh, float* r_elevation, const uint8_t* r_isOcean,
                             int32_t hIters, double K, double m, double dt, int32_t tIters,
                             double talusSlope, double kThermal, int32_t gIters,
                             double glacialStrength);
/* Same, plus taps on hydraulic iteration `captureIter` taken right before the implicit solve:
 * drainTarget int32[N] (the drainage-receiver field), flow float32[N], landOrder int32[N] (the
 * sorted landCells, -1 padded).  Any tap may be NULL. */
pb_status pb_erode_composite_debug(pb_mesh* mesh, float* r_elevation, const uint8_t* r_isOcean,
                                   int32_t hIters, double K, double m, double dt, int32_t tIters,
                                   double talusSlope, double kThermal, int32_t gIters,
                                   double glacialStrength, int32_t captureIter,
                                   int32_t* drainTarget_out, float* flow_out, int32_t* landOrder_out);
/* sharpenRidges(mesh, r_elevation, r_isOcean, iterations, strength)    js/terrain-post.js:713-751 */
pb_status pb_sharpen_ridges(pb_mesh* mesh, float* r_elevation, const uint8_t* r_isOcean,
                            int32_t iterations, double strength);
/* applySoilCreep(mesh, r_elevation, r_isOcean, iterations, strength)   js/terrain-post.js:758-794 */
pb_status pb_apply_soil_creep(pb_mesh* mesh, float* r_elevation, const uint8_t* r_isOcean,
                              int32_t iterations, double strength);

/* runPostProcessing(mesh, r_xyz, r_elevation, params, neighborDist, seed, r_hotspot)
 *                                                                     js/planet-worker.js:40-102
 * Slider values as the worker receives them.  hItersOverride > 0 replaces round(20*hydraulic)
 * while K stays 0.0006*hydraulic (BASELINE.json configs 2 and 5 ask for 50 / 200 iterations, more
 * than the UI slider can request); 0 or a negative value — so also a zero-initialised struct —
 * means "no override" (to switch the hydraulic stage off set hydraulicErosion = 0 like the UI).  Outputs: erosionDelta float32[N] (dl_erosionDelta) and, when
 * not NULL, the r_isOcean mask the pipeline derived after the warp (uint8[N]). */
typedef struct pb_post_params {
    double smoothing, glacialErosion, hydraulicErosion, thermalErosion, ridgeSharpening, terrainWarp;
    int32_t hItersOverride;
} pb_post_params;
pb_status pb_run_post_processing(pb_mesh* mesh, float* r_elevation, const pb_post_params* params,
                                 double seed, const float* r_hotspot, float* erosionDelta_out,
                                 uint8_t* r_isOcean_out);

/* per-stage device time of the last pb_run_post_processing (CUDA events), same stage order as the
 * reference's postTiming array: warp, smoothing, erosion, ridge, creep.  ms_out: double[5]. */
pb_status pb_last_post_timing(pb_mesh* mesh, double* ms_out);

/* ---- climate-util (js/climate-util.js) --------------------------------------------------------- */
/* smoothField(mesh, field, passes)                                    js/climate-util.js:5-25 */
pb_status pb_smooth_field(pb_mesh* mesh, float* field, int32_t passes);

/* ---- elevation (js/elevation.js) --------------------------------------------------------------------------
 * assignElevation(mesh, r_xyz, plateIsOcean, r_plate, plateVec, plateSeeds, noise, noiseMag, seed, spread,
 *                 plateDensity, superPlateData) → {r_elevation, mountain_r, coastline_r, ocean_r, r_stress,
 *                 debugLayers}                                                       js/elevation.js:216-1391
 * The JS objects keyed by plate id (plateVec {pid:{pole,omega}}, plateDensity {pid:d}) and the plateIsOcean Set
 * arrive as one table with a row per plate (host arrays, any order).  plateSeeds is the Set's iteration order.
 * superPlates == NULL means superPlateData == null (js/planet-worker.js:207-211).  The three region Sets come
 * back as membership masks.  debug[] order: base, tectonic, noise, interior, coastal, ocean, hotspot,
 * tecActivity, margins, backArc, foldRidge, orogenicPower (js/elevation.js:1386); any entry may be NULL.
 * `noiseSeed` seeds the caller's SimplexNoise instance (new SimplexNoise(seed), js/planet-worker.js:203).
 * Per-cell arrays follow the context's pointer mode; the plate tables are always host arrays.
 * The order-dependent middle of the function (stress propagation, the five randomized distance fills, the
 * capped FIFO BFS with payloads, the RNG-placed hotspot domes — SURVEY.md classes R/F/S) runs on host threads
 * inside this call; collisions and all per-cell synthesis run as CUDA kernels. */
typedef struct pb_plate_table {
    int32_t n; const int32_t* ids; const uint8_t* isOcean; const double* pole /* 3n */; const double* omega;
    const double* density;
} pb_plate_table;
typedef struct pb_elevation_result {
    float* r_elevation; float* r_stress; uint8_t* mountain_r; uint8_t* coastline_r; uint8_t* ocean_r; float* debug[12];
} pb_elevation_result;
pb_status pb_assign_elevation(pb_mesh* mesh, const pb_plate_table* plates, const int32_t* r_plate,
                              const int32_t* plateSeeds, int32_t numPlateSeeds, double noiseSeed, double noiseMag,
                              double seed, double spread, const pb_plate_table* superPlates,
                              const int32_t* r_superPlate, const pb_elevation_result* out);

/* ---- climate (js/wind.js, js/ocean.js, js/precipitation.js, js/heuristic-precip.js, js/temperature.js,
 * js/koppen.js) ------------------------------------------------------------------------------------------
 * A pb_climate holds the reference's result objects (windResult, oceanResult, precipResult, tempResult —
 * what the worker retains as W.cachedWind / W.cachedOcean, js/planet-worker.js:291, 588-606) resident in
 * HBM.  Stages must run in the reference's order (js/planet-worker.js:229-268); a stage called before its
 * inputs exist fails with PB_ERR_INVALID, like the reference throws on an undefined result object.
 * Fields are read back by the reference's own key names with pb_climate_get:
 *   wind   r_lat r_lon r_sinLat r_cosLat r_eastX/Y/Z r_northX/Y/Z r_continentality r_plateContinentality
 *          r_pressure_<s> r_wind_east_<s> r_wind_north_<s> r_wind_speed_<s> itczLons itczLatsSummer
 *          itczLatsWinter (f32; the ITCZ tables have 360 entries), r_isLand (u8), r_coastDistLand (i32)
 *   ocean  r_ocean_current_east_<s> r_ocean_current_north_<s> r_ocean_speed_<s> r_ocean_warmth_<s> (f32),
 *          r_oceanCoastDist r_westCoastDist r_eastCoastDist (i32; locals of computeCoastFields)
 *   precip r_precip_<s> r_rainshadow_<s> (f32)        temperature r_temperature_<s> (f32)
 *   koppen r_koppen (u8)                                <s> = summer | winter
 * r_elevation / r_plate follow the context's pointer mode; plateIsOcean is always a host array of plate ids
 * (the reference passes a Set, js/wind.js:394). */
typedef struct pb_climate pb_climate;
pb_status pb_climate_create(pb_mesh* mesh, pb_climate** out);
void pb_climate_destroy(pb_climate* climate);
/* computeWind(mesh, r_xyz, r_elevation, plateIsOcean, r_plate, noise, axialTilt)        js/wind.js:394-687
 * `noiseSeed` seeds the SimplexNoise instance the worker passes (new SimplexNoise(seed), planet-worker.js:203). */
pb_status pb_compute_wind(pb_climate* climate, const float* r_elevation, const int32_t* plateIsOcean,
                          int32_t numPlateIsOcean, const int32_t* r_plate, double noiseSeed, double axialTilt);
/* computeOceanCurrents(mesh, r_xyz, r_elevation, windResult)                            js/ocean.js:204-382 */
pb_status pb_compute_ocean_currents(pb_climate* climate, const float* r_elevation);
/* computePrecipitation(mesh, r_xyz, r_elevation, windResult, oceanResult, precipitationOffset, landCoverage)
 *                                                                              js/precipitation.js:196-684 */
pb_status pb_compute_precipitation(pb_climate* climate, const float* r_elevation, double precipitationOffset,
                                   double landCoverage);
/* computeTemperature(mesh, r_xyz, r_elevation, windResult, oceanResult, precipResult, temperatureOffset)
 *                                                                                js/temperature.js:69-237 */
pb_status pb_compute_temperature(pb_climate* climate, const float* r_elevation, double temperatureOffset);
/* classifyKoppen(mesh, r_elevation, tempResult, precipResult) → Uint8Array             js/koppen.js:67-288 */
pb_status pb_classify_koppen(pb_climate* climate, const float* r_elevation, uint8_t* r_koppen_out);
/* handleComputeClimate without the cache: all five stages                     js/planet-worker.js:579-672 */
pb_status pb_compute_climate(pb_climate* climate, const float* r_elevation, const int32_t* plateIsOcean,
                             int32_t numPlateIsOcean, const int32_t* r_plate, double noiseSeed,
                             double temperatureOffset, double precipitationOffset, double landCoverage,
                             uint8_t* r_koppen_out);
/* kind: 0 float32, 1 int32, 2 uint8.  PB_ERR_INVALID when the field does not exist (yet). */
pb_status pb_climate_field_info(pb_climate* climate, const char* name, int32_t* kind_out, int64_t* count_out);
pb_status pb_climate_get(pb_climate* climate, const char* name, void* out);

/* ---- plate pipeline on the hi-res mesh --------------------------------------------------
 * pb_project_coarse_plates replaces projectCoarsePlates(mesh, r_xyz, coarseMesh, coarse_xyz, coarse_r_plate, seed,
 * numPlates) (js/coarse-plates.js:51-117): the coarse mesh tables are host arrays (≈ 20 000 regions), numPlates < 0
 * stands for `null`; r_plate[numRegions] (pointer mode) receives the plate seed id of every region.
 * pb_smooth_and_reconnect_plates replaces smoothAndReconnectPlates(mesh, r_plate, plateSeeds, numPasses)
 * (js/plates.js:241-348): r_plate is mutated in place; plateSeeds is a host array in Set order.
 * pb_build_super_plates replaces buildSuperPlates(mesh, r_plate, plateSeeds, plateVec, plateIsOcean, plateDensity)
 * (js/super-plates.js:16-273): `plates` lists the plates in plateSeeds order (pole[3k] = NaN: the plate has no
 * plateVec entry; density = NaN: undefined); r_superPlate[numRegions] (pointer mode) and the caller-allocated
 * super-plate table (capacity >= plates->n is always enough) receive the result; super plate ids are 0 … n-1. */
typedef struct pb_super_plate_table {
    int32_t capacity;        /* in: entries the arrays below can hold */
    int32_t numSuperPlates;  /* out */
    double* pole;            /* out: 3 per super plate */
    double* omega;           /* out */
    uint8_t* isOcean;        /* out */
    double* density;         /* out */
} pb_super_plate_table;
pb_status pb_project_coarse_plates(pb_mesh* mesh, int32_t numCoarse, const int32_t* coarseAdjOffset, const int32_t* coarseAdjList,
                                   const float* coarse_xyz, const int32_t* coarse_r_plate, double seed, int32_t numPlates,
                                   int32_t* r_plate);
pb_status pb_smooth_and_reconnect_plates(pb_mesh* mesh, int32_t* r_plate, const int32_t* plateSeeds, int32_t numSeeds,
                                         int32_t numPasses);
pb_status pb_build_super_plates(pb_mesh* mesh, const int32_t* r_plate, const pb_plate_table* plates, int32_t* r_superPlate,
                                pb_super_plate_table* superOut);

/* pb_generate_coarse_plates replaces generateCoarsePlates(seed, numPlates, numContinents, continentSizeVariety,
 * landCoverage) (js/coarse-plates.js:19-39 = buildSphere(N_COARSE, 0.75, makeRng(seed + 137)) + generatePlates,
 * js/plates.js:6-232 + assignOceanLand, js/ocean-land.js:7-238).  numCoarse is the reference's N_COARSE (20000).
 * Host arrays only: coarse_xyz[3*(numCoarse+1)], coarse_r_plate[numCoarse+1]; the plate table (caller-allocated,
 * capacity >= numPlates) lists the plates in plateSeeds order with pole/omega (coarsePlateVec), isOcean
 * (coarsePlateIsOcean) and the density the worker draws for each plate (js/planet-worker.js:196-201).  The coarse mesh
 * comes back as an ordinary pb_mesh (pb_mesh_get_adjacency gives its CSR arrays; destroy it with pb_mesh_destroy). */
typedef struct pb_plate_table_out {
    int32_t capacity; int32_t n; int32_t* ids; uint8_t* isOcean; double* pole /* 3 per plate */; double* omega; double* density;
} pb_plate_table_out;
pb_status pb_generate_coarse_plates(pb_context* ctx, double seed, int32_t numPlates, int32_t numContinents,
                                    double continentSizeVariety, double landCoverage, int32_t numCoarse, pb_mesh** coarseMesh,
                                    float* coarse_xyz, int32_t* coarse_r_plate, pb_plate_table_out* plates);

/* ---- mesh construction ------------------------------------------------------------------
 * pb_triangulate_sphere replaces the triangulation inside buildSphere and the SphereMesh constructor's adjacency
 * (js/sphere-mesh.js:94-146, 174-186; Delaunator 5.0.1 + addPoleToMesh): r_xyz holds numRegions unit vectors (the
 * pole vertex (0,0,1) included as the reference does, :179-183); adjOffset[numRegions+1] and adjList[6*numRegions-12]
 * receive the CSR neighbour lists of the spherical Delaunay triangulation (= convex hull), each row in the
 * reference's circulation order `s = next(halfedges[s])` for the canonical triangle numbering described in
 * csrc/pb_meshgen.h.  Arrays follow the context's pointer mode.  PB_ERR_INVALID when the points do not give a closed
 * triangulated sphere (duplicates, fewer than 4 points) or when a region has more than 32 neighbours.  The candidate
 * search is sized for the mean spacing; point sets whose density varies by orders of magnitude are still triangulated
 * (the search block of the affected regions grows up to the whole sphere), only slowly.
 * pb_mesh_create_from_points = pb_triangulate_sphere + pb_mesh_create without the round trip through the host;
 * pb_mesh_get_adjacency returns the CSR arrays of a mesh (host pointers).
 * pb_generate_fibonacci_sphere replaces generateFibonacciSphere (js/sphere-mesh.js:9-37, jitter draws from
 * makeRng(seed), js/rng.js:3-6) and appends the pole vertex as buildSphere does (:179-183): xyz receives
 * 3*(numPoints+1) floats. */
pb_status pb_generate_fibonacci_sphere(pb_context* ctx, int32_t numPoints, double jitter, double seed, float* r_xyz);
pb_status pb_triangulate_sphere(pb_context* ctx, int32_t numRegions, const float* r_xyz, int32_t* adjOffset, int32_t* adjList);
pb_status pb_mesh_create_from_points(pb_context* ctx, int32_t numRegions, const float* r_xyz, pb_mesh** out);
/* The same mesh in the reference's OWN neighbour order: buildSphere with the published algorithm of its external
 * dependency delaunator@5.0.1 (js/sphere-mesh.js:174-186, js/planet-worker.js:17) — stereographic projection :41-53,
 * sweep-hull Delaunay with Delaunator's triangle numbering, addPoleToMesh :56-91, SphereMesh constructor :95-146
 * (a row starts at the first side seen for the region, :102-106).  That numbering fixes every order-dependent stage
 * downstream, so this is the mesh to use when the same seed must give the same planet as the web app.  Serial host
 * algorithm (≈ 1 s per million cells); pb_mesh_create_from_points is the 4 ms device builder with a canonical row
 * start.  r_xyz: numRegions points, the pole (0,0,1) last.  pb_mesh_get_triangles etc. return Delaunator's numbering. */
pb_status pb_mesh_create_delaunator(pb_context* ctx, int32_t numRegions, const float* r_xyz, pb_mesh** out);
pb_status pb_mesh_get_adjacency(const pb_mesh* mesh, int32_t* adjOffset, int32_t* adjList);

/* ---- importHeightmap pieces (js/planet-worker.js:682-831) -------------------------------------------------------------
 * pb_sample_heightmap replaces sampleHeightmap(mesh, r_xyz, imageData, imgW, imgH) (:715-727, with sampleBilinear :682 and
 * grayscaleToElevation :705): grayscale is a host array of width*height bytes (equirectangular, 0 = ocean).
 * pb_derive_synthetic_plates replaces deriveSyntheticPlates (:733-769): r_plate[r] = lowest region id of r's connected land
 * mass / ocean basin; plateSeeds are the ids with r_plate[r] == r in ascending order, a plate is oceanic iff
 * r_elevation[seed] <= 0, plateVec is all-zero.
 * pb_classify_imported_regions gives the mountain_r / coastline_r / ocean_r masks of :811-831.
 * Per-region arrays follow the context's pointer mode. */
pb_status pb_sample_heightmap(pb_mesh* mesh, const uint8_t* grayscale, int32_t width, int32_t height, float* r_elevation);
pb_status pb_derive_synthetic_plates(pb_mesh* mesh, const float* r_elevation, int32_t* r_plate);
pb_status pb_classify_imported_regions(pb_mesh* mesh, const float* r_elevation, uint8_t* mountain_r, uint8_t* coastline_r,
                                       uint8_t* ocean_r);

/* ---- per-region colour ramps -------------------------------------------------------------------
 * pb_region_colors fills rgb[3*numRegions] (Float32 colour buffer order r,g,b) for one of the renderer's colour modes:
 *   PB_COLOR_TERRAIN         elevationToColor(r_elevation[r])                     js/color-map.js:116-125
 *   PB_COLOR_BIOME           smoothBiomeColors(mesh, koppen, r_elevation)         js/planet-mesh.js:30-62 (biomeColor, js/color-map.js:73-114)
 *   PB_COLOR_HEIGHTMAP / PB_COLOR_LAND_HEIGHTMAP / PB_COLOR_LAND_MASK             js/planet-mesh.js:64-80
 *   PB_COLOR_BIOME_RAW       biomeColor(koppen[r], r_elevation[r]) without the neighbour blend
 *   PB_COLOR_KOPPEN          koppenColor(koppen[r]) = KOPPEN_CLASSES[id].color       js/planet-mesh.js:175-178, js/koppen.js:19-51
 * r_koppen is only read by the biome and koppen modes.  Arrays follow the context's pointer mode. */
enum { PB_COLOR_TERRAIN = 0, PB_COLOR_BIOME = 1, PB_COLOR_HEIGHTMAP = 2, PB_COLOR_LAND_HEIGHTMAP = 3, PB_COLOR_LAND_MASK = 4,
       PB_COLOR_BIOME_RAW = 5, PB_COLOR_KOPPEN = 6 };
pb_status pb_region_colors(pb_mesh* mesh, int32_t mode, const float* r_elevation, const uint8_t* r_koppen, float* rgb);

/* ---- equirectangular map export ------------------------------------------------------------------------------------
 * pb_export_map replaces exportMap(type, width) (js/planet-mesh.js:1752-1950) up to the ImageData the reference puts on its
 * canvas; the PNG container (canvas.toBlob there) is the caller's.  height = width / 2; rgba = width*height*4 bytes, top
 * row (north) first, alpha 255.  Export type → colorMode: 'colormap' PB_COLOR_TERRAIN, 'biome' (Satellite) PB_COLOR_BIOME,
 * 'koppen' PB_COLOR_KOPPEN, 'heightmap' / 'landheightmap' / 'landmask' the three grey modes (black background; the others
 * clear to THREE.Color(0x1a1a2e)).  Per side s of the mesh one map triangle (centroid of the inner triangle, centroid of
 * the outer triangle, the begin region) in lon/lat space, drawn twice when it straddles the date line (:1773-1846), flat
 * colour of the begin region; a pixel takes the LAST triangle in draw order whose closed area contains the pixel centre
 * (WebGL draws in order with depth LEQUAL at equal z); the render target's unorm8 levels go through the sRGB curve
 * (:1897-1915).  The rule is evaluated on the pixel grid of the whole image, so the reference's ≤ 2048-pixel tiles (a
 * browser memory workaround) do not appear.  pixelSide (optional, width*height ints): the side whose triangle owns the
 * pixel, -1 = background.  Arrays follow the context's pointer mode. */
pb_status pb_export_map(pb_mesh* mesh, int32_t colorMode, int32_t width, const float* r_elevation, const uint8_t* r_koppen,
                        uint8_t* rgba, int32_t* pixelSide);

/* ---- triangles: what the worker's `done` / `reapplyDone` / `editDone` replies carry for the renderer ----------------
 * pb_mesh_get_triangles: SphereMesh.triangles / .halfedges (js/sphere-mesh.js:94-100) of the mesh, 3*numTriangles ints
 * each, in the canonical numbering (csrc/pb_meshgen.h); numTriangles = 2*numRegions - 4.
 * pb_generate_triangle_centers replaces generateTriangleCenters(mesh, r_xyz) (js/sphere-mesh.js:206-219);
 * pb_compute_triangle_elevations replaces computeTriangleElevations(mesh, r_elevation) (js/planet-worker.js:29-37).
 * Arrays follow the context's pointer mode. */
int32_t pb_mesh_num_triangles(const pb_mesh* mesh);
pb_status pb_mesh_get_triangles(pb_mesh* mesh, int32_t* triangles, int32_t* halfedges);
/* SphereMesh._adjTriList (js/sphere-mesh.js:128-143, read by r_circulate_t): for every adjacency slot the triangle on the
 * inner side of that half-edge; numEdges ints. */
pb_status pb_mesh_get_adj_triangles(pb_mesh* mesh, int32_t* adjTriList);
pb_status pb_generate_triangle_centers(pb_mesh* mesh, float* t_xyz);
pb_status pb_compute_triangle_elevations(pb_mesh* mesh, const float* r_elevation, float* t_elevation);

/* ---- cell-range shards of the sweep loops (no reference counterpart: the reference is one thread) ----
 * One process per GPU; every rank creates the SAME mesh (the whole planet) and attaches a shard group to it.  From then
 * on every Jacobi / propagation sweep loop run on that mesh — pb_smooth_field (js/climate-util.js:5-25) and the loops
 * inside pb_compute_ocean_currents / pb_compute_precipitation / pb_compute_temperature / pb_compute_wind (smoothOcean
 * js/ocean.js:168-189, advectMoisture js/precipitation.js:118-179, rain-shadow / windward propagation :555-598,
 * diffuseOceanWarmth js/temperature.js:33-51) — computes only the rows of this rank's contiguous cell-id range
 * [N*rank/world, N*(rank+1)/world) and exchanges the one-cell halo with the adjacent ranges by storing the boundary
 * values straight into the peers' buffers over NVLink (CUDA-IPC mapped memory, flags in peer memory; no NCCL call and
 * no host round trip per sweep); at the end of a loop the ranges are all-gathered by peer stores so that the field is
 * whole on every rank again.  Everything else of the pipeline runs replicated.  Results are bit-identical to the
 * unsharded calls.  Planets below `minCells` (default 2 000 000) keep running unsharded: one sweep is then shorter than
 * the flag round trip.  Handles: 3 x 64 bytes (two sweep buffers, one control block), exchanged once by the caller
 * (torch.distributed all_gather in the Python mirror).  All ranks must issue the same sequence of calls. */
typedef struct pb_sweep_shards pb_sweep_shards;
pb_status pb_sweep_shards_create(pb_mesh* mesh, int32_t rank, int32_t worldSize, pb_sweep_shards** out);
void pb_sweep_shards_destroy(pb_sweep_shards* shards);
pb_status pb_sweep_shards_export(pb_sweep_shards* shards, unsigned char* handles192);
pb_status pb_sweep_shards_connect(pb_sweep_shards* shards, int32_t peerRank, const unsigned char* handles192);
pb_status pb_sweep_shards_set_min_cells(pb_sweep_shards* shards, int64_t minCells);
/* out8: first row, end row, adjacent ranks, halo bytes sent per sweep, sweeps run sharded, loops run sharded, active, minCells */
pb_status pb_sweep_shards_info(pb_sweep_shards* shards, int64_t* out8);

#ifdef __cplusplus
}
#endif
#endif /* PLANET_B200_H */
