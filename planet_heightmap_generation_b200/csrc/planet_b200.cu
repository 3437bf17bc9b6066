// planet_b200.cu — extern "C" surface declared in include/planet_b200.h.
// Compiled by nvcc for sm_100a into libplanet_b200.so (product), or by g++ with -DPB_EMUL into the
// test-only host emulation used by the CPU test-suite (tests/emul/).
#include "pb_engine.h"
#include "pb_climate_engine.h"
#include "pb_elevation_engine.h"
#include "pb_shardsweep.h"
#include "pb_meshgen.h"
#include "pb_delaunator.h"
#include "pb_plates.h"
#include "pb_coarse.h"
#include "pb_climate.h"
#include "pb_colors.h"
#include "pb_export.h"
#include <memory>
#include <cxxabi.h>

namespace {
thread_local std::string g_err;

template <class F>
pb_status guard(F&& f) {
    try {
        f();
        return PB_OK;
    } catch (const pb::Error& e) {
        g_err = e.what();
        return PB_ERR_CUDA;
    } catch (const std::invalid_argument& e) {
        g_err = e.what();
        return PB_ERR_INVALID;
    } catch (const std::exception& e) {
        g_err = e.what();
        return PB_ERR_INTERNAL;
    }
}
void need(bool c, const char* what) {
    if (!c) throw std::invalid_argument(what);
}
}  // namespace

struct pb_context {
    pb::Context c;
    std::unique_ptr<pb::SphereTriangulator> tri;   // scratch of pb_triangulate_sphere, kept between calls
    pb::DevBuf<float> triXyz; pb::DevBuf<int> triOff, triAdj;
    std::unique_ptr<pb::FibonacciSphere> fib;
    explicit pb_context(int d) : c(d) {}
};
struct pb_mesh {
    pb::Mesh m;
    std::unique_ptr<pb::Elevation> elevation;
    std::unique_ptr<pb::Plates> plates;
    pb::MeshTriangles triangles;
    pb::DevBuf<float> sTriOut;
    pb::DevBuf<uint8_t> sMaskA, sMaskB, sMaskC;
    pb::DevBuf<float> sColorRaw;
    std::unique_ptr<pb::MapExport> mapExport;
    pb::DevBuf<int> sPlateIO, sSuperIO;
    pb_mesh(pb::Context* c, int n, const int* o, const int* a, const float* x) : m(c, n, o, a, x) {}
};

struct pb_climate { pb::Climate c; explicit pb_climate(pb::Mesh* m) : c(m) {} };
struct pb_sweep_shards { pb::SweepShards s; pb_sweep_shards(pb::Mesh* m, int rank, int world) : s(m, rank, world) {} };

extern "C" {

const char* pb_last_error(void) { return g_err.c_str(); }
const char* pb_version(void) {
#if PB_CUDA
    return "planet_b200 0.1 (cuda sm_100a)";
#else
    return "planet_b200 0.1 (HOST EMULATION — test only)";
#endif
}
int64_t pb_launch_count(void) { return pb::launch_stats().launches; }

pb_status pb_context_create(int device, pb_context** out) {
    return guard([&] {
        need(out != nullptr, "out is NULL");
#if PB_CUDA
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) throw pb::Error("no CUDA device available: this library has no CPU path");
        need(device >= 0 && device < n, "device index out of range");
#endif
        *out = new pb_context(device);
    });
}
void pb_context_destroy(pb_context* ctx) { delete ctx; }
pb_status pb_set_stream(pb_context* ctx, void* s) {
    return guard([&] { need(ctx, "ctx is NULL"); ctx->c.ex.stream = (cudaStream_t)s; });
}
pb_status pb_set_pointer_mode(pb_context* ctx, int mode) {
    return guard([&] {
        need(ctx, "ctx is NULL");
        need(mode == PB_POINTER_HOST || mode == PB_POINTER_DEVICE, "unknown pointer mode");
        ctx->c.pointerMode = mode;
    });
}
pb_status pb_synchronize(pb_context* ctx) {
    return guard([&] { need(ctx, "ctx is NULL"); ctx->c.bind(); pb::stream_sync(ctx->c.ex.stream); });
}

pb_status pb_set_option(pb_context* ctx, const char* name, const char* value) {
    return guard([&] {
        need(ctx && name && value, "NULL argument");
        const std::string n = name, v = value;
        if (n == "flood") {
            need(v == "device" || v == "host", "flood must be 'device' or 'host'");
            ctx->c.floodOnHost = v == "host";
        } else if (n == "flow") {
            need(v == "auto" || v == "doubling" || v == "ordered", "flow must be 'auto', 'doubling' or 'ordered'");
            ctx->c.flowMode = v == "doubling" ? 1 : v == "ordered" ? 2 : 0;
        } else if (n == "mesh_order") {
            need(v == "canonical" || v == "delaunator", "mesh_order must be 'canonical' or 'delaunator'");
            ctx->c.meshOrderDelaunator = v == "delaunator";
        } else throw std::invalid_argument("unknown option: " + n);
    });
}

pb_status pb_profile_start(pb_context* ctx, const char* filter) {
    return guard([&] { need(ctx, "ctx is NULL"); ctx->c.bind(); ctx->c.profiler.start(filter); });
}
pb_status pb_profile_stop(pb_context* ctx, char* out, int64_t cap) {
    return guard([&] {
        need(ctx && out && cap > 2, "bad argument");
        ctx->c.bind();
        pb::stream_sync(ctx->c.ex.stream);
        std::map<std::string, std::pair<long long, double>> agg;
        ctx->c.profiler.collect(&agg);
        ctx->c.profiler.on = false;
        std::string js = "[";
        bool first = true;
        for (auto& kv : agg) {
            int st = 0;
            char* dm = abi::__cxa_demangle(kv.first.c_str(), nullptr, nullptr, &st);
            std::string nm = (st == 0 && dm) ? dm : kv.first;
            free(dm);
            char buf[512];
            snprintf(buf, sizeof buf, "%s{\"name\":\"%s\",\"launches\":%lld,\"ms\":%.6f}", first ? "" : ",", nm.c_str(),
                     kv.second.first, kv.second.second);
            js += buf;
            first = false;
        }
        js += "]";
        if ((int64_t)js.size() + 1 > cap) throw std::invalid_argument("profile buffer too small");
        memcpy(out, js.c_str(), js.size() + 1);
    });
}

pb_status pb_mesh_create(pb_context* ctx, int32_t n, const int32_t* off, const int32_t* adj, const float* xyz, pb_mesh** out) {
    return guard([&] {
        need(ctx && off && adj && xyz && out, "NULL argument");
        ctx->c.bind();
        *out = new pb_mesh(&ctx->c, n, off, adj, xyz);
    });
}
void pb_mesh_destroy(pb_mesh* mesh) { delete mesh; }
int32_t pb_mesh_num_regions(const pb_mesh* mesh) { return mesh ? mesh->m.N : 0; }
int64_t pb_mesh_num_edges(const pb_mesh* mesh) { return mesh ? mesh->m.E : 0; }

pb_status pb_compute_neighbor_dist(pb_mesh* mesh, float* out) {
    return guard([&] {
        need(mesh && out, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        if (m.hostMode()) pb::dev_copy(out, m.ndist.p, sizeof(float) * (size_t)m.E, 1, m.ex().stream);
        else pb::dev_copy(out, m.ndist.p, sizeof(float) * (size_t)m.E, 2, m.ex().stream);
        m.finish();
    });
}

pb_status pb_smooth_field(pb_mesh* mesh, float* field, int32_t passes) {
    return guard([&] {
        need(mesh && field, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        float* d = m.arg_in(field, m.N, m.sField);
        m.smooth_field(d, passes);
        m.arg_back(field, d, m.N);
        m.finish();
    });
}

pb_status pb_warp_terrain(pb_mesh* mesh, float* elev, double seed, double strength, const float* hotspot) {
    return guard([&] {
        need(mesh && elev, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        float* d = m.arg_in(elev, m.N, m.sElev);
        const float* h = m.arg_in(hotspot, m.N, m.sHot);
        m.warp_terrain(d, seed, strength, h);
        m.arg_back(elev, d, m.N);
        m.finish();
    });
}

#define PB_ELEV_OCEAN_PROLOGUE                                   \
    need(mesh && elev && isOcean, "NULL argument");              \
    pb::Mesh& m = mesh->m; m.ctx->bind();                        \
    float* d = m.arg_in(elev, m.N, m.sElev);                     \
    const uint8_t* o = m.arg_in(isOcean, m.N, m.sOcean);

pb_status pb_smooth_elevation(pb_mesh* mesh, float* elev, const uint8_t* isOcean, int32_t iterations, double strength) {
    return guard([&] {
        PB_ELEV_OCEAN_PROLOGUE
        m.smooth_elevation(d, o, iterations, strength);
        m.arg_back(elev, d, m.N);
        m.finish();
    });
}
pb_status pb_sharpen_ridges(pb_mesh* mesh, float* elev, const uint8_t* isOcean, int32_t iterations, double strength) {
    return guard([&] {
        PB_ELEV_OCEAN_PROLOGUE
        m.sharpen_ridges(d, o, iterations, strength);
        m.arg_back(elev, d, m.N);
        m.finish();
    });
}
pb_status pb_apply_soil_creep(pb_mesh* mesh, float* elev, const uint8_t* isOcean, int32_t iterations, double strength) {
    return guard([&] {
        PB_ELEV_OCEAN_PROLOGUE
        m.apply_soil_creep(d, o, iterations, strength);
        m.arg_back(elev, d, m.N);
        m.finish();
    });
}

pb_status pb_priority_flood_carve(pb_mesh* mesh, float* elev, const uint8_t* isOcean, double carveStrength,
                                  int32_t* drainTo, float* surface, uint8_t* openOcean) {
    return guard([&] {
        PB_ELEV_OCEAN_PROLOGUE
        pb::FloodTaps t;
        t.drainTo = m.arg_out(drainTo, m.N, m.sI0);
        t.surface = m.arg_out(surface, m.N, m.sField);
        t.openOcean = m.arg_out(openOcean, m.N, m.sU8);
        m.priority_flood_carve(d, o, carveStrength, &t);
        m.arg_back(elev, d, m.N);
        m.arg_back(drainTo, t.drainTo, m.N);
        m.arg_back(surface, t.surface, m.N);
        m.arg_back(openOcean, t.openOcean, m.N);
        m.finish();
    });
}

pb_status pb_erode_composite_debug(pb_mesh* mesh, float* elev, const uint8_t* isOcean, int32_t hIters, double K,
                                   double mm, double dt, int32_t tIters, double talus, double kThermal, int32_t gIters,
                                   double glacialStrength, int32_t captureIter, int32_t* drainTarget, float* flow,
                                   int32_t* landOrder) {
    return guard([&] {
        PB_ELEV_OCEAN_PROLOGUE
        pb::ErodeTaps t;
        t.captureIter = captureIter;
        t.drainTarget = m.arg_out(drainTarget, m.N, m.sI0);
        t.flow = m.arg_out(flow, m.N, m.sField);
        t.landOrder = m.arg_out(landOrder, m.N, m.sI1);
        m.erode_composite(d, o, hIters, K, mm, dt, tIters, talus, kThermal, gIters, glacialStrength, &t);
        m.arg_back(elev, d, m.N);
        m.arg_back(drainTarget, t.drainTarget, m.N);
        m.arg_back(flow, t.flow, m.N);
        m.arg_back(landOrder, t.landOrder, m.N);
        m.finish();
    });
}
pb_status pb_erode_composite(pb_mesh* mesh, float* elev, const uint8_t* isOcean, int32_t hIters, double K, double mm,
                             double dt, int32_t tIters, double talus, double kThermal, int32_t gIters,
                             double glacialStrength) {
    return pb_erode_composite_debug(mesh, elev, isOcean, hIters, K, mm, dt, tIters, talus, kThermal, gIters,
                                    glacialStrength, -1, nullptr, nullptr, nullptr);
}

pb_status pb_run_post_processing(pb_mesh* mesh, float* elev, const pb_post_params* p, double seed, const float* hotspot,
                                 float* erosionDelta, uint8_t* isOceanOut) {
    return guard([&] {
        need(mesh && elev && p, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        float* d = m.arg_in(elev, m.N, m.sElev);
        const float* h = m.arg_in(hotspot, m.N, m.sHot);
        float* dd = m.arg_out(erosionDelta, m.N, m.sDelta);
        uint8_t* oo = m.arg_out(isOceanOut, m.N, m.sOcean);
        m.run_post_processing(d, *p, seed, h, dd, oo);
        m.arg_back(elev, d, m.N);
        m.arg_back(erosionDelta, dd, m.N);
        m.arg_back(isOceanOut, oo, m.N);
        m.finish();
    });
}

pb_status pb_last_post_timing(pb_mesh* mesh, double* ms) {
    return guard([&] {
        need(mesh && ms, "NULL argument");
        mesh->m.ctx->bind();
        mesh->m.timer.resolve();
        for (int i = 0; i < 5; i++) ms[i] = mesh->m.timer.ms[i];
    });
}

// ---- elevation ---------------------------------------------------------------------------------------------
pb_status pb_assign_elevation(pb_mesh* mesh, const pb_plate_table* plates, const int32_t* r_plate, const int32_t* plateSeeds,
                              int32_t nSeeds, double noiseSeed, double noiseMag, double seed, double spread,
                              const pb_plate_table* superPlates, const int32_t* r_superPlate, const pb_elevation_result* out) {
    return guard([&] {
        need(mesh && plates && r_plate && out && out->r_elevation && (plateSeeds || nSeeds == 0) && nSeeds >= 0, "NULL argument");
        need(!superPlates || r_superPlate, "superPlates given without r_superPlate");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        if (!mesh->elevation) mesh->elevation.reset(new pb::Elevation(&m));
        pb::Elevation& E = *mesh->elevation;
        const size_t N = (size_t)m.N;
        pb::PlateTableHost P, SP;
        P.set(plates->n, plates->ids, plates->isOcean, plates->pole, plates->omega, plates->density);
        if (superPlates) SP.set(superPlates->n, superPlates->ids, superPlates->isOcean, superPlates->pole, superPlates->omega, superPlates->density);
        // r_plate / r_superPlate: one device copy for the kernels, one host copy for the host-serial stages
        std::vector<int> hPlate(N), hSuper;
        const int* dPlate; const int* dSuper = nullptr;
        if (m.hostMode()) {
            memcpy(hPlate.data(), r_plate, sizeof(int) * N);
            dPlate = m.arg_in(r_plate, N, E.sPlate);
            if (superPlates) { hSuper.assign(r_superPlate, r_superPlate + N); dSuper = m.arg_in(r_superPlate, N, E.sSuper); }
        } else {
            dPlate = r_plate;
            pb::dev_copy(hPlate.data(), r_plate, sizeof(int) * N, 1, m.ex().stream);
            if (superPlates) { hSuper.resize(N); dSuper = r_superPlate; pb::dev_copy(hSuper.data(), r_superPlate, sizeof(int) * N, 1, m.ex().stream); }
            pb::stream_sync(m.ex().stream);
        }
        for (size_t r = 0; r < N; r++) if (P.find(hPlate[r]) < 0) throw std::invalid_argument("r_plate holds an id that is not in the plate table");
        if (superPlates) for (size_t r = 0; r < N; r++) if (SP.find(hSuper[r]) < 0) throw std::invalid_argument("r_superPlate holds an id that is not in the super-plate table");
        pb::ElevationOutputs o;
        o.elev = m.arg_out(out->r_elevation, N, E.sElevOut);
        o.stress = out->r_stress ? m.arg_out(out->r_stress, N, E.sStressOut) : E.sStressOut.ensure(N);
        o.mountain = m.arg_out(out->mountain_r, N, E.sM);
        o.coastline = m.arg_out(out->coastline_r, N, E.sC);
        o.ocean = m.arg_out(out->ocean_r, N, E.sO);
        float** dbg[12] = {&o.dbg.base, &o.dbg.tectonic, &o.dbg.noise, &o.dbg.interior, &o.dbg.coastal, &o.dbg.ocean, &o.dbg.hotspot,
                           &o.dbg.tecActivity, &o.dbg.margins, &o.dbg.backArc, &o.dbg.foldRidge, &o.dbg.orogenicPower};
        for (int k = 0; k < 12; k++)
            *dbg[k] = (out->debug[k] && !m.hostMode()) ? out->debug[k] : E.sDbg[k].ensure(N);
        std::vector<int> seeds(plateSeeds, plateSeeds + nSeeds);
        E.assign(P, dPlate, hPlate.data(), seeds, noiseSeed, noiseMag, seed, spread, superPlates ? &SP : nullptr, dSuper,
                 superPlates ? hSuper.data() : nullptr, o);
        m.arg_back(out->r_elevation, o.elev, N);
        m.arg_back(out->r_stress, o.stress, N);
        m.arg_back(out->mountain_r, o.mountain, N);
        m.arg_back(out->coastline_r, o.coastline, N);
        m.arg_back(out->ocean_r, o.ocean, N);
        for (int k = 0; k < 12; k++) if (out->debug[k]) m.arg_back(out->debug[k], *dbg[k], N);
        m.finish();
    });
}

// ---- climate ---------------------------------------------------------------------------------------------
pb_status pb_climate_create(pb_mesh* mesh, pb_climate** out) {
    return guard([&] { need(mesh && out, "NULL argument"); mesh->m.ctx->bind(); *out = new pb_climate(&mesh->m); });
}
void pb_climate_destroy(pb_climate* c) { delete c; }

#define PB_CLIMATE_PROLOGUE                                          \
    need(climate && elev, "NULL argument");                          \
    pb::Climate& c = climate->c; pb::Mesh& m = *c.m; m.ctx->bind();  \
    const float* d = m.arg_in(elev, m.N, c.sElev);

pb_status pb_compute_wind(pb_climate* climate, const float* elev, const int32_t* ids, int32_t nIds, const int32_t* r_plate,
                          double noiseSeed, double axialTilt) {
    return guard([&] {
        PB_CLIMATE_PROLOGUE
        need(r_plate != nullptr && (ids != nullptr || nIds == 0) && nIds >= 0, "bad plate arguments");
        const int* p = m.arg_in(r_plate, m.N, c.sPlate);
        c.compute_wind(d, ids, nIds, p, noiseSeed, axialTilt);
        m.finish();
    });
}
pb_status pb_compute_ocean_currents(pb_climate* climate, const float* elev) {
    return guard([&] { PB_CLIMATE_PROLOGUE c.compute_ocean_currents(d); m.finish(); });
}
pb_status pb_compute_precipitation(pb_climate* climate, const float* elev, double precipitationOffset, double landCoverage) {
    return guard([&] { PB_CLIMATE_PROLOGUE c.compute_precipitation(d, precipitationOffset, landCoverage); m.finish(); });
}
pb_status pb_compute_temperature(pb_climate* climate, const float* elev, double temperatureOffset) {
    return guard([&] { PB_CLIMATE_PROLOGUE c.compute_temperature(d, temperatureOffset); m.finish(); });
}
pb_status pb_classify_koppen(pb_climate* climate, const float* elev, uint8_t* out) {
    return guard([&] {
        PB_CLIMATE_PROLOGUE
        c.classify_koppen(d);
        if (out) pb::dev_copy(out, c.cU("r_koppen"), (size_t)m.N, m.hostMode() ? 1 : 2, m.ex().stream);
        m.finish();
    });
}
pb_status pb_compute_climate(pb_climate* climate, const float* elev, const int32_t* ids, int32_t nIds, const int32_t* r_plate,
                             double noiseSeed, double temperatureOffset, double precipitationOffset, double landCoverage,
                             uint8_t* koppenOut) {
    return guard([&] {
        PB_CLIMATE_PROLOGUE
        need(r_plate != nullptr && (ids != nullptr || nIds == 0) && nIds >= 0, "bad plate arguments");
        const int* p = m.arg_in(r_plate, m.N, c.sPlate);
        c.compute_wind(d, ids, nIds, p, noiseSeed, 23.5);
        c.compute_ocean_currents(d);
        c.compute_precipitation(d, precipitationOffset, landCoverage);
        c.compute_temperature(d, temperatureOffset);
        c.classify_koppen(d);
        if (koppenOut) pb::dev_copy(koppenOut, c.cU("r_koppen"), (size_t)m.N, m.hostMode() ? 1 : 2, m.ex().stream);
        m.finish();
    });
}
pb_status pb_climate_field_info(pb_climate* climate, const char* name, int32_t* kind, int64_t* count) {
    return guard([&] {
        need(climate && name, "NULL argument");
        auto it = climate->c.count.find(name);
        if (it == climate->c.count.end()) throw std::invalid_argument(std::string("no such climate field: ") + name);
        if (kind) *kind = climate->c.kind[name];
        if (count) *count = (int64_t)it->second;
    });
}
pb_status pb_climate_get(pb_climate* climate, const char* name, void* out) {
    return guard([&] {
        need(climate && name && out, "NULL argument");
        pb::Climate& c = climate->c; pb::Mesh& m = *c.m; m.ctx->bind();
        auto it = c.count.find(name);
        if (it == c.count.end()) throw std::invalid_argument(std::string("no such climate field: ") + name);
        const int k = c.kind[name];
        const void* src = k == 0 ? (const void*)c.cF(name) : k == 1 ? (const void*)c.cI(name) : (const void*)c.cU(name);
        pb::dev_copy(out, src, it->second * (k == 2 ? 1 : 4), m.hostMode() ? 1 : 2, m.ex().stream);
        m.finish();
    });
}

static void triangulate(pb_context* ctx, int n, const float* xyz, bool hostPtrs, std::vector<int>* hOff, std::vector<int>* hAdj,
                        int32_t* outOff, int32_t* outAdj);

// ---- plate pipeline on the hi-res mesh -------------------------------------------------------------------------------
static pb::Plates& plates_of(pb_mesh* mesh) {
    if (!mesh->plates) mesh->plates.reset(new pb::Plates(&mesh->m));
    return *mesh->plates;
}
pb_status pb_project_coarse_plates(pb_mesh* mesh, int32_t numCoarse, const int32_t* cOff, const int32_t* cAdj, const float* coarse_xyz,
                                   const int32_t* coarse_r_plate, double seed, int32_t numPlates, int32_t* r_plate) {
    return guard([&] {
        need(mesh && cOff && cAdj && coarse_xyz && coarse_r_plate && r_plate, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        int* d = m.arg_out(r_plate, (size_t)m.N, mesh->sPlateIO);
        plates_of(mesh).project(numCoarse, cOff, cAdj, coarse_xyz, coarse_r_plate, seed, numPlates, d);
        m.arg_back(r_plate, d, (size_t)m.N);
        m.finish();
    });
}
pb_status pb_smooth_and_reconnect_plates(pb_mesh* mesh, int32_t* r_plate, const int32_t* plateSeeds, int32_t numSeeds, int32_t numPasses) {
    return guard([&] {
        need(mesh && r_plate && (plateSeeds || numSeeds == 0) && numSeeds >= 0 && numPasses >= 0, "bad argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        int* d = m.arg_in(r_plate, (size_t)m.N, mesh->sPlateIO);
        plates_of(mesh).smooth_and_reconnect(d, plateSeeds, numSeeds, numPasses);
        m.arg_back(r_plate, d, (size_t)m.N);
        m.finish();
    });
}
pb_status pb_build_super_plates(pb_mesh* mesh, const int32_t* r_plate, const pb_plate_table* plates, int32_t* r_superPlate,
                                pb_super_plate_table* superOut) {
    return guard([&] {
        need(mesh && r_plate && plates && r_superPlate && superOut, "NULL argument");
        need(plates->n > 0 && plates->ids && plates->isOcean && plates->pole && plates->omega && plates->density, "incomplete plate table");
        need(superOut->capacity >= 0 && superOut->pole && superOut->omega && superOut->isOcean && superOut->density, "incomplete output table");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        const int* dPlate = m.arg_in(r_plate, (size_t)m.N, mesh->sPlateIO);
        int* dSuper = m.arg_out(r_superPlate, (size_t)m.N, mesh->sSuperIO);
        pb::PlateTableIn T;
        T.n = plates->n; T.ids = plates->ids; T.isOcean = plates->isOcean; T.pole = plates->pole; T.omega = plates->omega; T.density = plates->density;
        pb::SuperPlatesOut o;
        plates_of(mesh).build_super_plates(dPlate, T, dSuper, o);
        if (o.n > superOut->capacity) throw std::invalid_argument("super-plate table capacity too small");
        superOut->numSuperPlates = o.n;
        for (int sp = 0; sp < o.n; sp++) {
            for (int c = 0; c < 3; c++) superOut->pole[3 * sp + c] = o.pole[3 * sp + c];
            superOut->omega[sp] = o.omega[sp]; superOut->isOcean[sp] = o.isOcean[sp]; superOut->density[sp] = o.density[sp];
        }
        m.arg_back(r_superPlate, dSuper, (size_t)m.N);
        m.finish();
    });
}

// ---- importHeightmap pieces ---------------------------------------------------------------------------------------------
pb_status pb_sample_heightmap(pb_mesh* mesh, const uint8_t* grayscale, int32_t width, int32_t height, float* r_elevation) {
    return guard([&] {
        need(mesh && grayscale && r_elevation, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        float* d = m.arg_out(r_elevation, (size_t)m.N, m.sElev);
        plates_of(mesh).sample_heightmap(grayscale, width, height, d);
        m.arg_back(r_elevation, d, (size_t)m.N);
        m.finish();
    });
}
pb_status pb_derive_synthetic_plates(pb_mesh* mesh, const float* r_elevation, int32_t* r_plate) {
    return guard([&] {
        need(mesh && r_elevation && r_plate, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        const float* e = m.arg_in(r_elevation, (size_t)m.N, m.sElev);
        int* d = m.arg_out(r_plate, (size_t)m.N, mesh->sPlateIO);
        plates_of(mesh).derive_synthetic_plates(e, d);
        m.arg_back(r_plate, d, (size_t)m.N);
        m.finish();
    });
}
pb_status pb_classify_imported_regions(pb_mesh* mesh, const float* r_elevation, uint8_t* mountain_r, uint8_t* coastline_r, uint8_t* ocean_r) {
    return guard([&] {
        need(mesh && r_elevation && mountain_r && coastline_r && ocean_r, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        const size_t N = (size_t)m.N;
        const float* e = m.arg_in(r_elevation, N, m.sElev);
        uint8_t* a = m.arg_out(mountain_r, N, mesh->sMaskA); uint8_t* b = m.arg_out(coastline_r, N, mesh->sMaskB); uint8_t* c = m.arg_out(ocean_r, N, mesh->sMaskC);
        plates_of(mesh).classify_imported(e, a, b, c);
        m.arg_back(mountain_r, a, N); m.arg_back(coastline_r, b, N); m.arg_back(ocean_r, c, N);
        m.finish();
    });
}

// ---- colour ramps ---------------------------------------------------------------------------------------------------------
pb_status pb_region_colors(pb_mesh* mesh, int32_t mode, const float* r_elevation, const uint8_t* r_koppen, float* rgb) {
    return guard([&] {
        need(mesh && r_elevation && rgb, "NULL argument");
        need(mode >= 0 && mode <= 6, "unknown colour mode");
        const bool biome = mode == pb::COLOR_BIOME || mode == pb::COLOR_BIOME_RAW || mode == pb::COLOR_KOPPEN;
        need(!biome || r_koppen, "biome / koppen colours need r_koppen");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        const size_t N = (size_t)m.N;
        const float* e = m.arg_in(r_elevation, N, m.sElev);
        const uint8_t* k = biome ? m.arg_in(r_koppen, N, mesh->sMaskA) : nullptr;
        float* out = m.arg_out(rgb, 3 * N, mesh->sTriOut);
        if (mode == pb::COLOR_BIOME) {
            float* raw = mesh->sColorRaw.ensure(3 * N);
            m.ex().for_each(m.N, pb::RegionColorK{mode, e, k, raw});
            m.ex().for_each(m.N, pb::BiomeBlendK{m.csr(), raw, out});
        } else {
            m.ex().for_each(m.N, pb::RegionColorK{mode, e, k, out});
        }
        m.arg_back(rgb, out, 3 * N);
        m.finish();
    });
}

// exportMap(type, width) js/planet-mesh.js:1752-1950 up to the ImageData: width × width/2 RGBA8 pixels, top row first
pb_status pb_export_map(pb_mesh* mesh, int32_t colorMode, int32_t width, const float* r_elevation, const uint8_t* r_koppen,
                        uint8_t* rgba, int32_t* pixelSide) {
    return guard([&] {
        need(mesh && r_elevation && rgba, "NULL argument");
        need(colorMode >= 0 && colorMode <= 6 && colorMode != pb::COLOR_BIOME_RAW, "unknown export colour mode");
        need(width >= 2 && width <= 65536 && width % 2 == 0, "width must be even, 2 … 65536");
        const bool biome = colorMode == pb::COLOR_BIOME || colorMode == pb::COLOR_KOPPEN;
        need(!biome || r_koppen, "biome / koppen exports need r_koppen");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        if (!mesh->mapExport) mesh->mapExport.reset(new pb::MapExport());
        pb::MapExport& X = *mesh->mapExport;
        const size_t N = (size_t)m.N, px = (size_t)width * (size_t)(width / 2);
        const float* e = m.arg_in(r_elevation, N, m.sElev);
        const uint8_t* k = biome ? m.arg_in(r_koppen, N, mesh->sMaskA) : nullptr;
        float* rgb = X.rgb.ensure(3 * N);
        if (colorMode == pb::COLOR_BIOME) {
            m.ex().for_each(m.N, pb::RegionColorK{colorMode, e, k, X.rgbRaw.ensure(3 * N)});
            m.ex().for_each(m.N, pb::BiomeBlendK{m.csr(), X.rgbRaw.p, rgb});
        } else {
            m.ex().for_each(m.N, pb::RegionColorK{colorMode, e, k, rgb});
        }
        uint8_t* out = m.arg_out(rgba, 4 * px, X.sRgba);
        int* side = pixelSide ? m.arg_out(pixelSide, px, X.sSide) : nullptr;
        const bool bw = colorMode == pb::COLOR_HEIGHTMAP || colorMode == pb::COLOR_LAND_HEIGHTMAP || colorMode == pb::COLOR_LAND_MASK;
        X.run(m, mesh->triangles, rgb, bw, width, out, side);
        m.arg_back(rgba, out, 4 * px);
        if (pixelSide) m.arg_back(pixelSide, side, px);
        m.finish();
    });
}

// ---- triangles (render-side consumers of the mesh) ------------------------------------------------------------------
int32_t pb_mesh_num_triangles(const pb_mesh* mesh) { return mesh ? 2 * mesh->m.N - 4 : 0; }
pb_status pb_mesh_get_triangles(pb_mesh* mesh, int32_t* triangles, int32_t* halfedges) {
    return guard([&] {
        need(mesh && triangles && halfedges, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        mesh->triangles.build(m.ex(), m.csr(), m.N);
        const size_t n3 = 3 * (size_t)mesh->triangles.T;
        const int kind = m.hostMode() ? 1 : 2;
        pb::dev_copy(triangles, mesh->triangles.tri.p, sizeof(int) * n3, kind, m.ex().stream);
        pb::dev_copy(halfedges, mesh->triangles.half.p, sizeof(int) * n3, kind, m.ex().stream);
        m.finish();
    });
}
pb_status pb_mesh_get_adj_triangles(pb_mesh* mesh, int32_t* adjTriList) {
    return guard([&] {
        need(mesh && adjTriList, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        mesh->triangles.build(m.ex(), m.csr(), m.N);
        pb::dev_copy(adjTriList, mesh->triangles.adjTri.p, sizeof(int) * (size_t)m.E, m.hostMode() ? 1 : 2, m.ex().stream);
        m.finish();
    });
}
pb_status pb_generate_triangle_centers(pb_mesh* mesh, float* t_xyz) {
    return guard([&] {
        need(mesh && t_xyz, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        mesh->triangles.build(m.ex(), m.csr(), m.N);
        const size_t T = (size_t)mesh->triangles.T;
        float* d = m.arg_out(t_xyz, 3 * T, mesh->sTriOut);
        m.ex().for_each((int)T, pb::TriCentersK{mesh->triangles.tri.p, m.xyz.p, d});
        m.arg_back(t_xyz, d, 3 * T);
        m.finish();
    });
}
pb_status pb_compute_triangle_elevations(pb_mesh* mesh, const float* r_elevation, float* t_elevation) {
    return guard([&] {
        need(mesh && r_elevation && t_elevation, "NULL argument");
        pb::Mesh& m = mesh->m; m.ctx->bind();
        mesh->triangles.build(m.ex(), m.csr(), m.N);
        const size_t T = (size_t)mesh->triangles.T;
        const float* e = m.arg_in(r_elevation, (size_t)m.N, m.sElev);
        float* d = m.arg_out(t_elevation, T, mesh->sTriOut);
        m.ex().for_each((int)T, pb::TriElevationK{mesh->triangles.tri.p, e, d});
        m.arg_back(t_elevation, d, T);
        m.finish();
    });
}

// generateCoarsePlates: coarse mesh on the device, growth / ocean-land logic on the host, smoothing on the device
pb_status pb_generate_coarse_plates(pb_context* ctx, double seed, int32_t numPlates, int32_t numContinents, double continentSizeVariety,
                                    double landCoverage, int32_t numCoarse, pb_mesh** coarseMesh, float* coarse_xyz,
                                    int32_t* coarse_r_plate, pb_plate_table_out* plates) {
    return guard([&] {
        need(ctx && coarseMesh && coarse_xyz && coarse_r_plate && plates, "NULL argument");
        need(numPlates >= 1 && numCoarse >= 4, "numPlates >= 1 and numCoarse >= 4 required");
        need(plates->capacity >= numPlates && plates->ids && plates->isOcean && plates->pole && plates->omega && plates->density,
             "plate table capacity below numPlates");
        ctx->c.bind();
        const pb::Exec& ex = ctx->c.ex;
        const int n = numCoarse + 1;
        if (!ctx->fib) ctx->fib.reset(new pb::FibonacciSphere());
        pb::DevBuf<float> dXyz;
        ctx->fib->generate(ex, numCoarse, 0.75, seed + 137, dXyz.ensure(3 * (size_t)n));     // COARSE_JITTER, isolated RNG (:12, 20)
        pb::dev_copy(coarse_xyz, dXyz.p, sizeof(float) * 3 * (size_t)n, 1, ex.stream);
        pb::stream_sync(ex.stream);
        std::vector<int> hOff, hAdj;
        const int savedMode = ctx->c.pointerMode;
        triangulate(ctx, n, coarse_xyz, true, &hOff, &hAdj, nullptr, nullptr);
        std::unique_ptr<pb_mesh> cm(new pb_mesh(&ctx->c, n, hOff.data(), hAdj.data(), coarse_xyz));
        pb::CoarsePlatesResult R;
        pb::Plates& P = plates_of(cm.get());
        pb::DevBuf<int> dPlate;
        auto smooth = [&](std::vector<int>& plate, const std::vector<int>& seeds, int passes) {
            pb::dev_copy(dPlate.ensure(n), plate.data(), sizeof(int) * (size_t)n, 0, ex.stream);
            P.smooth_and_reconnect(dPlate.p, seeds.data(), (int)seeds.size(), passes);
            pb::dev_copy(plate.data(), dPlate.p, sizeof(int) * (size_t)n, 1, ex.stream);
            pb::stream_sync(ex.stream);
        };
        pb::CoarseStage::generate_plates(n, hOff.data(), hAdj.data(), coarse_xyz, numPlates, seed, smooth, R);
        pb::CoarseStage::ocean_land(n, hOff.data(), hAdj.data(), coarse_xyz, seed, numContinents, continentSizeVariety, landCoverage, R);
        (void)savedMode;
        memcpy(coarse_r_plate, R.r_plate.data(), sizeof(int) * (size_t)n);
        plates->n = (int32_t)R.seeds.size();
        for (size_t k = 0; k < R.seeds.size(); k++) {
            plates->ids[k] = R.seeds[k]; plates->isOcean[k] = R.isOcean[k]; plates->omega[k] = R.omega[k]; plates->density[k] = R.density[k];
            for (int c = 0; c < 3; c++) plates->pole[3 * k + c] = R.pole[3 * k + c];
        }
        *coarseMesh = cm.release();
    });
}

// ---- mesh construction ------------------------------------------------------------------------------------------
static void triangulate(pb_context* ctx, int n, const float* xyz, bool hostPtrs, std::vector<int>* hOff, std::vector<int>* hAdj,
                        int32_t* outOff, int32_t* outAdj) {
    if (n < 4) throw std::invalid_argument("a sphere mesh needs at least 4 points");
    const pb::Exec& ex = ctx->c.ex;
    const size_t E = 6 * (size_t)n - 12;
    if (ctx->c.meshOrderDelaunator) {
        // option mesh_order = delaunator: the reference's own neighbour order (host algorithm, pb_delaunator.h)
        std::vector<float> hx;
        const float* h = xyz;
        if (!hostPtrs) { hx.resize(3 * (size_t)n); pb::dev_copy(hx.data(), xyz, sizeof(float) * hx.size(), 1, ex.stream); pb::stream_sync(ex.stream); h = hx.data(); }
        pb::delaunator::SphereMeshHost sm;
        try { pb::delaunator::build_sphere(h, n, sm); }
        catch (const std::invalid_argument&) { throw; }
        catch (const std::exception& e) { throw pb::Error(e.what()); }
        if (sm.adjList.size() != E) throw pb::Error("Delaunator mesh is not a closed triangulated sphere");
        if (outOff) { pb::dev_copy(outOff, sm.adjOffset.data(), sizeof(int) * ((size_t)n + 1), hostPtrs ? 3 : 0, ex.stream);
                      pb::dev_copy(outAdj, sm.adjList.data(), sizeof(int) * E, hostPtrs ? 3 : 0, ex.stream); pb::stream_sync(ex.stream); }
        if (hOff) { *hOff = sm.adjOffset; *hAdj = sm.adjList; }
        return;
    }
    pb::DevBuf<float>& dXyz = ctx->triXyz; pb::DevBuf<int>& dOff = ctx->triOff; pb::DevBuf<int>& dAdj = ctx->triAdj;
    const float* px = xyz;
    if (hostPtrs) { pb::dev_copy(dXyz.ensure(3 * (size_t)n), xyz, sizeof(float) * 3 * (size_t)n, 0, ex.stream); px = dXyz.p; }
    int* po = (!hostPtrs && outOff) ? outOff : dOff.ensure((size_t)n + 1);
    int* pa = (!hostPtrs && outAdj) ? outAdj : dAdj.ensure(E);
    if (!ctx->tri) ctx->tri.reset(new pb::SphereTriangulator());
    ctx->tri->build(ex, n, px, po, pa);
    if (hostPtrs && outOff) { pb::dev_copy(outOff, po, sizeof(int) * ((size_t)n + 1), 1, ex.stream); pb::dev_copy(outAdj, pa, sizeof(int) * E, 1, ex.stream); }
    if (hOff) { hOff->resize((size_t)n + 1); hAdj->resize(E);
                pb::dev_copy(hOff->data(), po, sizeof(int) * ((size_t)n + 1), 1, ex.stream); pb::dev_copy(hAdj->data(), pa, sizeof(int) * E, 1, ex.stream); }
    pb::stream_sync(ex.stream);
}
pb_status pb_generate_fibonacci_sphere(pb_context* ctx, int32_t numPoints, double jitter, double seed, float* xyz) {
    return guard([&] {
        need(ctx && xyz, "NULL argument");
        need(numPoints >= 1, "numPoints must be positive");
        ctx->c.bind();
        if (!ctx->fib) ctx->fib.reset(new pb::FibonacciSphere());
        const pb::Exec& ex = ctx->c.ex;
        const size_t n3 = 3 * ((size_t)numPoints + 1);
        if (ctx->c.pointerMode == PB_POINTER_HOST) {
            ctx->fib->generate(ex, numPoints, jitter, seed, ctx->triXyz.ensure(n3));
            pb::dev_copy(xyz, ctx->triXyz.p, sizeof(float) * n3, 1, ex.stream);
            pb::stream_sync(ex.stream);
        } else {
            ctx->fib->generate(ex, numPoints, jitter, seed, xyz);
        }
    });
}
pb_status pb_triangulate_sphere(pb_context* ctx, int32_t n, const float* xyz, int32_t* off, int32_t* adj) {
    return guard([&] {
        need(ctx && xyz && off && adj, "NULL argument");
        ctx->c.bind();
        triangulate(ctx, n, xyz, ctx->c.pointerMode == PB_POINTER_HOST, nullptr, nullptr, off, adj);
    });
}
pb_status pb_mesh_create_delaunator(pb_context* ctx, int32_t n, const float* xyz, pb_mesh** out);
pb_status pb_mesh_create_from_points(pb_context* ctx, int32_t n, const float* xyz, pb_mesh** out) {
    if (ctx && ctx->c.meshOrderDelaunator) return pb_mesh_create_delaunator(ctx, n, xyz, out);
    return guard([&] {
        need(ctx && xyz && out, "NULL argument");
        ctx->c.bind();
        std::vector<int> hOff, hAdj;
        triangulate(ctx, n, xyz, true, &hOff, &hAdj, nullptr, nullptr);
        *out = new pb_mesh(&ctx->c, n, hOff.data(), hAdj.data(), xyz);
    });
}
pb_status pb_mesh_create_delaunator(pb_context* ctx, int32_t n, const float* xyz, pb_mesh** out) {
    return guard([&] {
        need(ctx && xyz && out && n >= 5, "bad argument");
        ctx->c.bind();
        std::vector<float> hx;
        const float* h = xyz;
        if (ctx->c.pointerMode == PB_POINTER_DEVICE) {
            hx.resize(3 * (size_t)n);
            pb::dev_copy(hx.data(), xyz, sizeof(float) * hx.size(), 1, ctx->c.ex.stream);
            pb::stream_sync(ctx->c.ex.stream);
            h = hx.data();
        }
        pb::delaunator::SphereMeshHost sm;
        try { pb::delaunator::build_sphere(h, n, sm); }
        catch (const std::invalid_argument&) { throw; }
        catch (const std::exception& e) { throw pb::Error(e.what()); }
        std::unique_ptr<pb_mesh> mesh(new pb_mesh(&ctx->c, n, sm.adjOffset.data(), sm.adjList.data(), h));
        // the triangle arrays keep Delaunator's numbering too (worker replies: triangles, halfedges, t_xyz, t_elevation)
        pb::MeshTriangles& t = mesh->triangles;
        const size_t S = sm.triangles.size();
        if ((long long)S != 3ll * (2ll * n - 4)) throw pb::Error("Delaunator mesh is not a closed triangulated sphere");
        const cudaStream_t st = ctx->c.ex.stream;
        pb::dev_copy(t.tri.ensure(S), sm.triangles.data(), sizeof(int) * S, 0, st);
        pb::dev_copy(t.half.ensure(S), sm.halfedges.data(), sizeof(int) * S, 0, st);
        pb::dev_copy(t.adjTri.ensure(sm.adjTri.size()), sm.adjTri.data(), sizeof(int) * sm.adjTri.size(), 0, st);
        pb::stream_sync(st);
        t.T = (int)(S / 3);
        *out = mesh.release();
    });
}
pb_status pb_mesh_get_adjacency(const pb_mesh* mesh, int32_t* off, int32_t* adj) {
    return guard([&] {
        need(mesh && off && adj, "NULL argument");
        memcpy(off, mesh->m.hOffCopy.data(), sizeof(int) * mesh->m.hOffCopy.size());
        memcpy(adj, mesh->m.hAdjCopy.data(), sizeof(int) * mesh->m.hAdjCopy.size());
    });
}

// ---- cell-range shards of the sweep loops (pb_shardsweep.h) ---------------------------------------------------------
pb_status pb_sweep_shards_create(pb_mesh* mesh, int32_t rank, int32_t world, pb_sweep_shards** out) {
    return guard([&] {
        need(mesh && out, "NULL argument");
        need(mesh->m.shards == nullptr, "the mesh already has a shard group");
        mesh->m.ctx->bind();
        *out = new pb_sweep_shards(&mesh->m, rank, world);
        mesh->m.shards = &(*out)->s;
    });
}
void pb_sweep_shards_destroy(pb_sweep_shards* g) {
    if (!g) return;
    if (g->s.m->shards == &g->s) g->s.m->shards = nullptr;
    delete g;
}
pb_status pb_sweep_shards_export(pb_sweep_shards* g, unsigned char* handles192) {
    return guard([&] { need(g && handles192, "NULL argument"); g->s.m->ctx->bind(); g->s.export_handles(handles192); });
}
pb_status pb_sweep_shards_connect(pb_sweep_shards* g, int32_t peerRank, const unsigned char* handles192) {
    return guard([&] { need(g && handles192, "NULL argument"); g->s.m->ctx->bind(); g->s.connect(peerRank, handles192); });
}
pb_status pb_sweep_shards_set_min_cells(pb_sweep_shards* g, int64_t minCells) {
    return guard([&] { need(g && minCells >= 0, "bad argument"); g->s.minCells = minCells; });
}
pb_status pb_sweep_shards_info(pb_sweep_shards* g, int64_t* out8) {
    return guard([&] {
        need(g && out8, "NULL argument");
        out8[0] = g->s.lo; out8[1] = g->s.hi; out8[2] = (int64_t)g->s.adj.size(); out8[3] = g->s.haloBytesPerSweep;
        out8[4] = g->s.sweepsRun; out8[5] = g->s.loopsRun; out8[6] = g->s.active() ? 1 : 0; out8[7] = g->s.minCells;
    });
}

}  // extern "C"
