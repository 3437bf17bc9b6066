// planet_b200_addon.cc — Node-API addon over include/planet_b200.h.
//
// Exposes the reference's stage functions (same names, same argument order, typed arrays in place) so that
// js/planet-worker.js can import them from this addon instead of the ES modules js/terrain-post.js,
// js/elevation.js, js/wind.js, js/ocean.js, js/precipitation.js, js/temperature.js, js/koppen.js
// (see bindings/node/planet_worker_shim.mjs and INTEGRATION.md).  The worker keeps exactly one mesh at a time
// (its retained state W, js/planet-worker.js:277-292); so does this addon.
//
// Not buildable in this image (no Node headers); type-checked against bindings/node/stub/node_api.h.
// Build with node-gyp:  sources: [planet_b200_addon.cc], include_dirs: [../../include],
//                        libraries: [-lplanet_b200], cflags_cc: [-std=c++17].
#include <node_api.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "planet_b200.h"

namespace {

pb_context* g_ctx = nullptr;
pb_mesh* g_mesh = nullptr;
pb_climate* g_clim = nullptr;

#define PB_TRY(env, expr)                                                     \
    do { if ((expr) != PB_OK) { napi_throw_error(env, nullptr, pb_last_error()); return nullptr; } } while (0)
#define NEED_MESH(env)                                                        \
    do { if (!g_mesh) { napi_throw_error(env, nullptr, "setMesh() has not been called"); return nullptr; } } while (0)

template <class T> struct napi_type_of;
template <> struct napi_type_of<float> { static const napi_typedarray_type value = napi_float32_array; };
template <> struct napi_type_of<double> { static const napi_typedarray_type value = napi_float64_array; };
template <> struct napi_type_of<int32_t> { static const napi_typedarray_type value = napi_int32_array; };
template <> struct napi_type_of<uint8_t> { static const napi_typedarray_type value = napi_uint8_array; };
template <> struct napi_type_of<const float> : napi_type_of<float> {};
template <> struct napi_type_of<const int32_t> : napi_type_of<int32_t> {};
template <> struct napi_type_of<const uint8_t> : napi_type_of<uint8_t> {};
template <> struct napi_type_of<const double> : napi_type_of<double> {};

struct Args {
    // napi_get_cb_info fills at most `argc` (in) slots and then stores the ACTUAL argument count: clamp it to the capacity,
    // otherwise present(i) would read past v[] for a call with more arguments (assignElevationFlat takes 17)
    static const size_t kMaxArgs = 20;
    napi_env env; size_t argc = kMaxArgs; napi_value v[kMaxArgs];
    Args(napi_env e, napi_callback_info info) : env(e) {
        napi_get_cb_info(e, info, &argc, v, nullptr, nullptr);
        if (argc > kMaxArgs) argc = kMaxArgs;
    }
    bool present(size_t i) const {
        if (i >= argc) return false;
        napi_valuetype t; napi_typeof(env, v[i], &t);
        return t != napi_undefined && t != napi_null;
    }
    double num(size_t i, double dflt = 0) const { double d = dflt; if (present(i)) napi_get_value_double(env, v[i], &d); return d; }
    int32_t i32(size_t i, int32_t dflt = 0) const { int32_t d = dflt; if (present(i)) napi_get_value_int32(env, v[i], &d); return d; }
    // Typed-array argument i as T*: the element type must be T's (a Float64Array where a Float32Array is expected would be read
    // as garbage), and — for per-cell arrays, cells() — the length must be the mesh's numRegions: the C ABI reads exactly that
    // many elements.  A mismatch records an error; the caller raises it as a TypeError before touching the library.
    mutable const char* error = nullptr;
    template <class T> T* typed(size_t i, size_t* len = nullptr, size_t want = (size_t)-1) const {
        if (!present(i)) return nullptr;
        napi_typedarray_type t; size_t n; void* data; napi_value buf; size_t off;
        if (napi_get_typedarray_info(env, v[i], &t, &n, &data, &buf, &off) != napi_ok) { error = "argument is not a typed array"; return nullptr; }
        if (t != napi_type_of<T>::value) { error = "typed array has the wrong element type"; return nullptr; }
        if (want != (size_t)-1 && n != want) { error = "typed array has the wrong length"; return nullptr; }
        if (len) *len = n;
        return static_cast<T*>(data);
    }
    template <class T> T* cells(size_t i) const { return typed<T>(i, nullptr, g_mesh ? (size_t)pb_mesh_num_regions(g_mesh) : 0); }
    bool failed() const { return error != nullptr; }
    napi_value raise() const { napi_throw_type_error(env, nullptr, error); return nullptr; }
};
#define ARGS_OK(a) do { if ((a).failed()) return (a).raise(); } while (0)

double prop_num(napi_env env, napi_value obj, const char* name, double dflt = 0) {
    bool has = false; napi_has_named_property(env, obj, name, &has);
    if (!has) return dflt;
    napi_value v; napi_get_named_property(env, obj, name, &v);
    double d = dflt; napi_get_value_double(env, v, &d);
    return d;
}

napi_value new_typed(napi_env env, napi_typedarray_type type, size_t n, size_t elem, void** data) {
    napi_value ab, ta;
    napi_create_arraybuffer(env, n * elem, data, &ab);
    napi_create_typedarray(env, type, n, ab, 0, &ta);
    return ta;
}
napi_value undefined(napi_env env) { napi_value u; napi_get_undefined(env, &u); return u; }

// setMesh(numRegions, adjOffset:Int32Array, adjList:Int32Array, r_xyz:Float32Array)   after buildSphere (:149)
napi_value SetMesh(napi_env env, napi_callback_info info) {
    Args a(env, info);
    if (!g_ctx) PB_TRY(env, pb_context_create(0, &g_ctx));
    if (g_clim) { pb_climate_destroy(g_clim); g_clim = nullptr; }
    if (g_mesh) { pb_mesh_destroy(g_mesh); g_mesh = nullptr; }
    const int32_t n = a.i32(0);
    if (n <= 0) { napi_throw_type_error(env, nullptr, "numRegions must be positive"); return nullptr; }
    size_t nAdj = 0;
    const int32_t* off = a.typed<int32_t>(1, nullptr, (size_t)n + 1);
    const int32_t* adj = a.typed<int32_t>(2, &nAdj);
    const float* xyz = a.typed<float>(3, nullptr, 3 * (size_t)n);
    ARGS_OK(a);
    if (!off || !adj || !xyz || (size_t)off[n] != nAdj) { napi_throw_type_error(env, nullptr, "adjList length must equal adjOffset[numRegions]"); return nullptr; }
    PB_TRY(env, pb_mesh_create(g_ctx, n, off, adj, xyz, &g_mesh));
    PB_TRY(env, pb_climate_create(g_mesh, &g_clim));
    return undefined(env);
}
// setOption(name, value)
napi_value SetOption(napi_env env, napi_callback_info info) {
    Args a(env, info);
    char name[64] = {0}, value[64] = {0}; size_t n;
    napi_get_value_string_utf8(env, a.v[0], name, sizeof name, &n);
    napi_get_value_string_utf8(env, a.v[1], value, sizeof value, &n);
    if (!g_ctx) PB_TRY(env, pb_context_create(0, &g_ctx));
    PB_TRY(env, pb_set_option(g_ctx, name, value));
    return undefined(env);
}
// computeNeighborDist(mesh, r_xyz) → Float32Array                                   js/sphere-mesh.js:191
napi_value ComputeNeighborDist(napi_env env, napi_callback_info info) {
    NEED_MESH(env);
    void* data; napi_value out = new_typed(env, napi_float32_array, (size_t)pb_mesh_num_edges(g_mesh), 4, &data);
    PB_TRY(env, pb_compute_neighbor_dist(g_mesh, static_cast<float*>(data)));
    return out;
}
// warpTerrain(mesh, r_elevation, r_xyz, seed, strength, r_hotspot)                  js/terrain-post.js:233
napi_value WarpTerrain(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    (void)(a.cells<float>(1), a.cells<float>(5)); ARGS_OK(a);
    PB_TRY(env, pb_warp_terrain(g_mesh, a.cells<float>(1), a.num(3), a.num(4), a.cells<float>(5)));
    return undefined(env);
}
// smoothElevation(mesh, r_elevation, r_isOcean, iterations, strength)               :317
napi_value SmoothElevation(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    (void)(a.cells<float>(1), a.cells<uint8_t>(2)); ARGS_OK(a);
    PB_TRY(env, pb_smooth_elevation(g_mesh, a.cells<float>(1), a.cells<uint8_t>(2), a.i32(3), a.num(4)));
    return undefined(env);
}
// erodeComposite(mesh, r_elevation, r_xyz, r_isOcean, hIters, K, m, dt, tIters, talusSlope, kThermal, gIters,
//                glacialStrength, neighborDist)                                      :369
napi_value ErodeComposite(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    (void)(a.cells<float>(1), a.cells<uint8_t>(3)); ARGS_OK(a);
    PB_TRY(env, pb_erode_composite(g_mesh, a.cells<float>(1), a.cells<uint8_t>(3), a.i32(4), a.num(5), a.num(6), a.num(7), a.i32(8),
                                   a.num(9), a.num(10), a.i32(11, 0), a.num(12, 0)));
    return undefined(env);
}
// sharpenRidges / applySoilCreep(mesh, r_elevation, r_isOcean, iterations, strength) :713 / :758
napi_value SharpenRidges(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    (void)(a.cells<float>(1), a.cells<uint8_t>(2)); ARGS_OK(a);
    PB_TRY(env, pb_sharpen_ridges(g_mesh, a.cells<float>(1), a.cells<uint8_t>(2), a.i32(3), a.num(4)));
    return undefined(env);
}
napi_value ApplySoilCreep(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    (void)(a.cells<float>(1), a.cells<uint8_t>(2)); ARGS_OK(a);
    PB_TRY(env, pb_apply_soil_creep(g_mesh, a.cells<float>(1), a.cells<uint8_t>(2), a.i32(3), a.num(4)));
    return undefined(env);
}
// runPostProcessing(mesh, r_xyz, r_elevation, params, neighborDist, seed, r_hotspot) → {dl_erosionDelta, postTiming}
//                                                                                    js/planet-worker.js:40-102
napi_value RunPostProcessing(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    pb_post_params p;
    p.smoothing = prop_num(env, a.v[3], "smoothing"); p.glacialErosion = prop_num(env, a.v[3], "glacialErosion");
    p.hydraulicErosion = prop_num(env, a.v[3], "hydraulicErosion"); p.thermalErosion = prop_num(env, a.v[3], "thermalErosion");
    p.ridgeSharpening = prop_num(env, a.v[3], "ridgeSharpening"); p.terrainWarp = prop_num(env, a.v[3], "terrainWarp");
    p.hItersOverride = (int32_t)prop_num(env, a.v[3], "hItersOverride", -1);
    const size_t n = (size_t)pb_mesh_num_regions(g_mesh);
    void* d; napi_value delta = new_typed(env, napi_float32_array, n, 4, &d);
    (void)(a.cells<float>(2), a.cells<float>(6)); ARGS_OK(a);
    PB_TRY(env, pb_run_post_processing(g_mesh, a.cells<float>(2), &p, a.num(5), a.cells<float>(6), static_cast<float*>(d), nullptr));
    napi_value out; napi_create_object(env, &out);
    napi_set_named_property(env, out, "dl_erosionDelta", delta);
    double ms[5]; static const char* stage[5] = {"Terrain warp", "Smoothing", "Erosion composite", "Ridge sharpening", "Soil creep"};
    if (pb_last_post_timing(g_mesh, ms) == PB_OK) {
        napi_value t; napi_create_object(env, &t);
        for (int k = 0; k < 5; k++) { napi_value v; napi_create_double(env, ms[k], &v); napi_set_named_property(env, t, stage[k], v); }
        napi_set_named_property(env, out, "postTimingMs", t);
    }
    return out;
}

// plate table from parallel arrays: (ids:Int32Array, isOcean:Uint8Array, pole:Float64Array[3n], omega:Float64Array, density:Float64Array)
bool plate_table(const Args& a, size_t first, pb_plate_table* t) {
    size_t n = 0;
    t->ids = a.typed<int32_t>(first, &n); t->n = (int32_t)n;
    t->isOcean = a.typed<uint8_t>(first + 1, nullptr, n); t->pole = a.typed<double>(first + 2, nullptr, 3 * n);
    t->omega = a.typed<double>(first + 3, nullptr, n); t->density = a.typed<double>(first + 4, nullptr, n);
    return !a.failed() && t->ids && t->isOcean && t->pole && t->omega && t->density;
}
// assignElevation(r_plate, plateIds, plateIsOcean, platePole, plateOmega, plateDensity, plateSeeds:Int32Array, noiseSeed,
//                 noiseMag, seed, spread, [r_superPlate, sIds, sIsOcean, sPole, sOmega, sDensity])
//   → {r_elevation, r_stress, mountain_r, coastline_r, ocean_r, debugLayers}         js/elevation.js:216
// (the shim flattens plateVec / plateDensity / plateIsOcean / superPlateData into these arrays)
napi_value AssignElevation(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    pb_plate_table P, SP;
    if (!plate_table(a, 1, &P)) { napi_throw_error(env, nullptr, "bad plate table"); return nullptr; }
    size_t nSeeds = 0; const int32_t* seeds = a.typed<int32_t>(6, &nSeeds);
    const bool dual = a.present(11);
    if (dual && !plate_table(a, 12, &SP)) { napi_throw_error(env, nullptr, "bad super-plate table"); return nullptr; }
    const size_t n = (size_t)pb_mesh_num_regions(g_mesh);
    pb_elevation_result r;
    napi_value out, dbg; napi_create_object(env, &out); napi_create_object(env, &dbg);
    void* d;
    napi_set_named_property(env, out, "r_elevation", new_typed(env, napi_float32_array, n, 4, &d)); r.r_elevation = static_cast<float*>(d);
    napi_set_named_property(env, out, "r_stress", new_typed(env, napi_float32_array, n, 4, &d)); r.r_stress = static_cast<float*>(d);
    napi_set_named_property(env, out, "mountain_r", new_typed(env, napi_uint8_array, n, 1, &d)); r.mountain_r = static_cast<uint8_t*>(d);
    napi_set_named_property(env, out, "coastline_r", new_typed(env, napi_uint8_array, n, 1, &d)); r.coastline_r = static_cast<uint8_t*>(d);
    napi_set_named_property(env, out, "ocean_r", new_typed(env, napi_uint8_array, n, 1, &d)); r.ocean_r = static_cast<uint8_t*>(d);
    static const char* layers[12] = {"base", "tectonic", "noise", "interior", "coastal", "ocean", "hotspot", "tecActivity", "margins",
                                     "backArc", "foldRidge", "orogenicPower"};
    for (int k = 0; k < 12; k++) { napi_set_named_property(env, dbg, layers[k], new_typed(env, napi_float32_array, n, 4, &d)); r.debug[k] = static_cast<float*>(d); }
    napi_set_named_property(env, out, "debugLayers", dbg);
    (void)(a.cells<int32_t>(0)); ARGS_OK(a);
    PB_TRY(env, pb_assign_elevation(g_mesh, &P, a.cells<int32_t>(0), seeds, (int32_t)nSeeds, a.num(7), a.num(8), a.num(9), a.num(10),
                                    dual ? &SP : nullptr, dual ? a.typed<int32_t>(11) : nullptr, &r));
    return out;
}

// climate stages: results stay on the device; getClimateField(name) fetches one array by the reference's key
napi_value ComputeWind(napi_env env, napi_callback_info info) {          // (r_elevation, plateIsOceanIds:Int32Array, r_plate, noiseSeed, axialTilt)
    NEED_MESH(env); Args a(env, info);
    size_t nIds = 0; const int32_t* ids = a.typed<int32_t>(1, &nIds);
    (void)(a.cells<float>(0), a.cells<int32_t>(2)); ARGS_OK(a);
    PB_TRY(env, pb_compute_wind(g_clim, a.cells<float>(0), ids, (int32_t)nIds, a.cells<int32_t>(2), a.num(3), a.num(4, 23.5)));
    return undefined(env);
}
napi_value ComputeOceanCurrents(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    (void)(a.cells<float>(0)); ARGS_OK(a);
    PB_TRY(env, pb_compute_ocean_currents(g_clim, a.cells<float>(0)));
    return undefined(env);
}
napi_value ComputePrecipitation(napi_env env, napi_callback_info info) {  // (r_elevation, precipitationOffset, landCoverage)
    NEED_MESH(env); Args a(env, info);
    (void)(a.cells<float>(0)); ARGS_OK(a);
    PB_TRY(env, pb_compute_precipitation(g_clim, a.cells<float>(0), a.num(1, 0), a.num(2, 0.3)));
    return undefined(env);
}
napi_value ComputeTemperature(napi_env env, napi_callback_info info) {    // (r_elevation, temperatureOffset)
    NEED_MESH(env); Args a(env, info);
    (void)(a.cells<float>(0)); ARGS_OK(a);
    PB_TRY(env, pb_compute_temperature(g_clim, a.cells<float>(0), a.num(1, 0)));
    return undefined(env);
}
napi_value ClassifyKoppen(napi_env env, napi_callback_info info) {        // (r_elevation) → Uint8Array
    NEED_MESH(env); Args a(env, info);
    void* d; napi_value out = new_typed(env, napi_uint8_array, (size_t)pb_mesh_num_regions(g_mesh), 1, &d);
    (void)(a.cells<float>(0)); ARGS_OK(a);
    PB_TRY(env, pb_classify_koppen(g_clim, a.cells<float>(0), static_cast<uint8_t*>(d)));
    return out;
}
napi_value GetClimateField(napi_env env, napi_callback_info info) {       // (name) → typed array
    NEED_MESH(env); Args a(env, info);
    char name[96] = {0}; size_t len;
    napi_get_value_string_utf8(env, a.v[0], name, sizeof name, &len);
    int32_t kind; int64_t count;
    PB_TRY(env, pb_climate_field_info(g_clim, name, &kind, &count));
    void* d;
    napi_value out = new_typed(env, kind == 0 ? napi_float32_array : kind == 1 ? napi_int32_array : napi_uint8_array, (size_t)count,
                               kind == 2 ? 1 : 4, &d);
    PB_TRY(env, pb_climate_get(g_clim, name, d));
    return out;
}
// smoothField(mesh, field, passes)                                                  js/climate-util.js:5
napi_value SmoothField(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    (void)(a.cells<float>(1)); ARGS_OK(a);
    PB_TRY(env, pb_smooth_field(g_mesh, a.cells<float>(1), a.i32(2)));
    return undefined(env);
}

// getMeshTriangles() → {triangles:Int32Array, halfedges:Int32Array}: SphereMesh.triangles / .halfedges of the retained mesh
// (js/sphere-mesh.js:94-100; the worker posts both in its `done` reply and reads s_begin_r(s) = triangles[s])
napi_value GetMeshTriangles(napi_env env, napi_callback_info info) {
    NEED_MESH(env);
    const size_t n3 = 3 * (size_t)pb_mesh_num_triangles(g_mesh);
    void *t, *h;
    napi_value tt = new_typed(env, napi_int32_array, n3, 4, &t), th = new_typed(env, napi_int32_array, n3, 4, &h);
    PB_TRY(env, pb_mesh_get_triangles(g_mesh, static_cast<int32_t*>(t), static_cast<int32_t*>(h)));
    napi_value out; napi_create_object(env, &out);
    napi_set_named_property(env, out, "triangles", tt); napi_set_named_property(env, out, "halfedges", th);
    return out;
}
// generateTriangleCenters(mesh, r_xyz) → Float32Array                              js/sphere-mesh.js:206
napi_value GenerateTriangleCenters(napi_env env, napi_callback_info info) {
    NEED_MESH(env);
    void* d; napi_value out = new_typed(env, napi_float32_array, 3 * (size_t)pb_mesh_num_triangles(g_mesh), 4, &d);
    PB_TRY(env, pb_generate_triangle_centers(g_mesh, static_cast<float*>(d)));
    return out;
}

// exportMapPixels(type, width, r_elevation, r_koppen | null) → Uint8ClampedArray(width * width/2 * 4): the ImageData of exportMap
// (js/planet-mesh.js:1752-1950) — `new ImageData(pixels, width)` + putImageData + canvas.toBlob stay on the JS side
napi_value ExportMapPixels(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    char type[32] = {0}; size_t len = 0;
    if (!a.present(0) || napi_get_value_string_utf8(env, a.v[0], type, sizeof type, &len) != napi_ok) { napi_throw_type_error(env, nullptr, "export type must be a string"); return nullptr; }
    const std::string t = type;
    const uint8_t* koppen = a.cells<uint8_t>(3);
    const float* elev = a.cells<float>(2);
    ARGS_OK(a);
    int32_t mode = t == "landmask" ? PB_COLOR_LAND_MASK : t == "landheightmap" ? PB_COLOR_LAND_HEIGHTMAP : t == "heightmap" ? PB_COLOR_HEIGHTMAP
                 : t == "biome" ? PB_COLOR_BIOME : t == "koppen" ? PB_COLOR_KOPPEN : PB_COLOR_TERRAIN;
    if ((mode == PB_COLOR_BIOME || mode == PB_COLOR_KOPPEN) && !koppen) mode = PB_COLOR_TERRAIN;     // no climate yet (:1764-1765, 1788-1793)
    const int32_t width = a.i32(1);
    if (width < 2 || width > 65536 || width % 2) { napi_throw_type_error(env, nullptr, "width must be even, 2 … 65536"); return nullptr; }
    void* px; napi_value out = new_typed(env, napi_uint8_clamped_array, (size_t)width * (size_t)(width / 2) * 4, 1, &px);
    PB_TRY(env, pb_export_map(g_mesh, mode, width, elev, koppen, static_cast<uint8_t*>(px), nullptr));
    return out;
}

// buildSphereFlat(N, jitter, seed) → {numRegions, r_xyz, adjOffset, adjList}; the mesh becomes the retained mesh   js/sphere-mesh.js:174
napi_value BuildSphere(napi_env env, napi_callback_info info) {
    Args a(env, info);
    if (!g_ctx) PB_TRY(env, pb_context_create(0, &g_ctx));
    const int32_t n = a.i32(0) + 1;
    void *xyz, *off, *adj;
    napi_value txyz = new_typed(env, napi_float32_array, 3 * (size_t)n, 4, &xyz);
    PB_TRY(env, pb_generate_fibonacci_sphere(g_ctx, n - 1, a.num(1), a.num(2), static_cast<float*>(xyz)));
    if (g_clim) { pb_climate_destroy(g_clim); g_clim = nullptr; }
    if (g_mesh) { pb_mesh_destroy(g_mesh); g_mesh = nullptr; }
    PB_TRY(env, pb_mesh_create_from_points(g_ctx, n, static_cast<float*>(xyz), &g_mesh));
    PB_TRY(env, pb_climate_create(g_mesh, &g_clim));
    napi_value toff = new_typed(env, napi_int32_array, (size_t)n + 1, 4, &off);
    napi_value tadj = new_typed(env, napi_int32_array, (size_t)pb_mesh_num_edges(g_mesh), 4, &adj);
    PB_TRY(env, pb_mesh_get_adjacency(g_mesh, static_cast<int32_t*>(off), static_cast<int32_t*>(adj)));
    napi_value out, nv; napi_create_object(env, &out); napi_create_double(env, n, &nv);
    napi_set_named_property(env, out, "numRegions", nv); napi_set_named_property(env, out, "r_xyz", txyz);
    napi_set_named_property(env, out, "adjOffset", toff); napi_set_named_property(env, out, "adjList", tadj);
    return out;
}
// generateCoarsePlatesFlat(seed, numPlates, numContinents, continentSizeVariety, landCoverage)                       js/coarse-plates.js:19
//   → {numRegions, adjOffset, adjList, coarse_xyz, coarse_r_plate, seeds, isOcean, pole, omega, density}
napi_value GenerateCoarsePlates(napi_env env, napi_callback_info info) {
    Args a(env, info);
    if (!g_ctx) PB_TRY(env, pb_context_create(0, &g_ctx));
    const int32_t P = a.i32(1), NC = 20000, n = NC + 1;
    void *xyz, *rp, *ids, *oc, *pole, *om, *de, *off, *adj;
    napi_value txyz = new_typed(env, napi_float32_array, 3 * (size_t)n, 4, &xyz), trp = new_typed(env, napi_int32_array, n, 4, &rp);
    napi_value tids = new_typed(env, napi_int32_array, P, 4, &ids), toc = new_typed(env, napi_uint8_array, P, 1, &oc);
    napi_value tpole = new_typed(env, napi_float64_array, 3 * (size_t)P, 8, &pole), tom = new_typed(env, napi_float64_array, P, 8, &om);
    napi_value tde = new_typed(env, napi_float64_array, P, 8, &de);
    pb_plate_table_out t{P, 0, static_cast<int32_t*>(ids), static_cast<uint8_t*>(oc), static_cast<double*>(pole), static_cast<double*>(om),
                         static_cast<double*>(de)};
    pb_mesh* coarse = nullptr;
    PB_TRY(env, pb_generate_coarse_plates(g_ctx, a.num(0), P, a.i32(2), a.num(3, 0.0), a.num(4, 0.3), NC, &coarse, static_cast<float*>(xyz),
                                          static_cast<int32_t*>(rp), &t));
    napi_value toff = new_typed(env, napi_int32_array, (size_t)n + 1, 4, &off);
    napi_value tadj = new_typed(env, napi_int32_array, (size_t)pb_mesh_num_edges(coarse), 4, &adj);
    const pb_status st = pb_mesh_get_adjacency(coarse, static_cast<int32_t*>(off), static_cast<int32_t*>(adj));
    pb_mesh_destroy(coarse);
    PB_TRY(env, st);
    napi_value out, nv, np; napi_create_object(env, &out); napi_create_double(env, n, &nv); napi_create_double(env, t.n, &np);
    const char* names[] = {"adjOffset", "adjList", "coarse_xyz", "coarse_r_plate", "seeds", "isOcean", "pole", "omega", "density"};
    napi_value vals[] = {toff, tadj, txyz, trp, tids, toc, tpole, tom, tde};
    for (int k = 0; k < 9; k++) napi_set_named_property(env, out, names[k], vals[k]);
    napi_set_named_property(env, out, "numRegions", nv); napi_set_named_property(env, out, "numPlates", np);
    return out;
}
// projectCoarsePlatesFlat(numCoarse, cAdjOffset, cAdjList, coarse_xyz, coarse_r_plate, seed, numPlates|null) → Int32Array   js/coarse-plates.js:51
napi_value ProjectCoarsePlates(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    void* d;
    napi_value out = new_typed(env, napi_int32_array, (size_t)pb_mesh_num_regions(g_mesh), 4, &d);
    // the coarse mesh has its own size: offsets numCoarse + 1, xyz 3 numCoarse, r_plate numCoarse, adjList adjOffset[numCoarse]
    const int32_t nc = a.i32(0);
    if (nc <= 0) { napi_throw_type_error(env, nullptr, "numCoarse must be positive"); return nullptr; }
    size_t nAdj = 0;
    const int32_t* cOff = a.typed<int32_t>(1, nullptr, (size_t)nc + 1);
    const int32_t* cAdj = a.typed<int32_t>(2, &nAdj);
    const float* cXyz = a.typed<float>(3, nullptr, 3 * (size_t)nc);
    const int32_t* cPlate = a.typed<int32_t>(4, nullptr, (size_t)nc);
    ARGS_OK(a);
    if (!cOff || !cAdj || !cXyz || !cPlate || (size_t)cOff[nc] != nAdj) { napi_throw_type_error(env, nullptr, "coarse adjList length must equal adjOffset[numCoarse]"); return nullptr; }
    PB_TRY(env, pb_project_coarse_plates(g_mesh, nc, cOff, cAdj, cXyz, cPlate, a.num(5), a.present(6) ? a.i32(6) : -1, static_cast<int32_t*>(d)));
    return out;
}
// smoothAndReconnectPlatesFlat(r_plate, plateSeeds:Int32Array, numPasses)  in place                                 js/plates.js:241
napi_value SmoothAndReconnectPlates(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    size_t ns = 0;
    const int32_t* seeds = a.typed<int32_t>(1, &ns);
    (void)(a.cells<int32_t>(0)); ARGS_OK(a);
    PB_TRY(env, pb_smooth_and_reconnect_plates(g_mesh, a.cells<int32_t>(0), seeds, (int32_t)ns, a.i32(2)));
    return undefined(env);
}
// buildSuperPlatesFlat(r_plate, ids, isOcean, pole, omega, density) → {r_superPlate, numSuperPlates, pole, omega, isOcean, density}   js/super-plates.js:16
napi_value BuildSuperPlates(napi_env env, napi_callback_info info) {
    NEED_MESH(env); Args a(env, info);
    size_t P = 0;
    const int32_t* ids = a.typed<int32_t>(1, &P);
    pb_plate_table t{(int32_t)P, ids, a.typed<uint8_t>(2, nullptr, P), a.typed<double>(3, nullptr, 3 * P), a.typed<double>(4, nullptr, P), a.typed<double>(5, nullptr, P)};
    const size_t cap = P < 2 ? 2 : P;
    void *rs, *pole, *om, *oc, *de;
    napi_value trs = new_typed(env, napi_int32_array, (size_t)pb_mesh_num_regions(g_mesh), 4, &rs);
    napi_value tpole = new_typed(env, napi_float64_array, 3 * cap, 8, &pole), tom = new_typed(env, napi_float64_array, cap, 8, &om);
    napi_value toc = new_typed(env, napi_uint8_array, cap, 1, &oc), tde = new_typed(env, napi_float64_array, cap, 8, &de);
    pb_super_plate_table sp{(int32_t)cap, 0, static_cast<double*>(pole), static_cast<double*>(om), static_cast<uint8_t*>(oc), static_cast<double*>(de)};
    (void)(a.cells<int32_t>(0)); ARGS_OK(a);
    PB_TRY(env, pb_build_super_plates(g_mesh, a.cells<int32_t>(0), &t, static_cast<int32_t*>(rs), &sp));
    napi_value out, nv; napi_create_object(env, &out); napi_create_double(env, sp.numSuperPlates, &nv);
    napi_set_named_property(env, out, "r_superPlate", trs); napi_set_named_property(env, out, "numSuperPlates", nv);
    napi_set_named_property(env, out, "pole", tpole); napi_set_named_property(env, out, "omega", tom);
    napi_set_named_property(env, out, "isOcean", toc); napi_set_named_property(env, out, "density", tde);
    return out;
}

}  // namespace

NAPI_MODULE_INIT() {
    const napi_property_descriptor props[] = {
        {"setMesh", nullptr, SetMesh, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"setOption", nullptr, SetOption, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"computeNeighborDist", nullptr, ComputeNeighborDist, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"warpTerrain", nullptr, WarpTerrain, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"smoothElevation", nullptr, SmoothElevation, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"erodeComposite", nullptr, ErodeComposite, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"sharpenRidges", nullptr, SharpenRidges, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"applySoilCreep", nullptr, ApplySoilCreep, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"runPostProcessing", nullptr, RunPostProcessing, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"assignElevationFlat", nullptr, AssignElevation, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"computeWindFlat", nullptr, ComputeWind, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"computeOceanCurrentsFlat", nullptr, ComputeOceanCurrents, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"computePrecipitationFlat", nullptr, ComputePrecipitation, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"computeTemperatureFlat", nullptr, ComputeTemperature, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"classifyKoppenFlat", nullptr, ClassifyKoppen, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"getClimateField", nullptr, GetClimateField, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"smoothField", nullptr, SmoothField, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"getMeshTriangles", nullptr, GetMeshTriangles, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"generateTriangleCenters", nullptr, GenerateTriangleCenters, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"exportMapPixels", nullptr, ExportMapPixels, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"buildSphereFlat", nullptr, BuildSphere, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"generateCoarsePlatesFlat", nullptr, GenerateCoarsePlates, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"projectCoarsePlatesFlat", nullptr, ProjectCoarsePlates, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"smoothAndReconnectPlatesFlat", nullptr, SmoothAndReconnectPlates, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"buildSuperPlatesFlat", nullptr, BuildSuperPlates, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
    return exports;
}
