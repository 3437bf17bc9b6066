// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).
// extern "C" surface for ctypes (oracle/binding.py).  No product code links this.
#include <cstring>
#include "noise.h"
#include "terrain_post.h"
#include "climate.h"
#include "mesh.h"

extern "C" {

void orc_rng(double seed, int n, double* out) {
    Rng r(seed);
    for (int i = 0; i < n; i++) out[i] = r.next();
}

void orc_simplex_perm(double seed, uint8_t* perm512, uint8_t* pm12_512) {
    SimplexNoise s(seed);
    std::memcpy(perm512, s.perm, 512);
    if (pm12_512) std::memcpy(pm12_512, s.pm12, 512);
}

// kind: 0 noise3D, 1 fbm(octaves, persistence), 2 ridgedFbm(octaves) defaults otherwise
void orc_noise(double seed, int kind, int n, const double* xyz, int octaves, double persistence, double* out) {
    SimplexNoise s(seed);
    for (int i = 0; i < n; i++) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        if (kind == 0) out[i] = s.noise3D(x, y, z);
        else if (kind == 1) out[i] = s.fbm(x, y, z, octaves, persistence);
        else out[i] = s.ridgedFbm(x, y, z, octaves);
    }
}

double orc_cell_noise(double r) { return oracle_cell_noise(r); }

// kind: 0 exp 1 log 2 pow(x,y) 3 atan 4 asin 5 atan2(x=y_arg,y=x_arg) 6 sin 7 cos 8 tanh
void orc_detmath(int kind, int n, const double* x, const double* y, double* out) {
    for (int i = 0; i < n; i++) {
        switch (kind) {
            case 0: out[i] = pb_exp(x[i]); break;
            case 1: out[i] = pb_log(x[i]); break;
            case 2: out[i] = pb_pow(x[i], y[i]); break;
            case 3: out[i] = pb_atan(x[i]); break;
            case 4: out[i] = pb_asin(x[i]); break;
            case 5: out[i] = pb_atan2(x[i], y[i]); break;
            case 6: out[i] = pb_sin(x[i]); break;
            case 7: out[i] = pb_cos(x[i]); break;
            case 8: out[i] = pb_tanh(x[i]); break;
            default: out[i] = 0;
        }
    }
}

void orc_fibonacci_sphere(int n, double jitter, double seed, float* xyz_out) {
    oracle_fibonacci_sphere(n, jitter, seed, xyz_out);
}

void orc_neighbor_dist(int N, const int32_t* off, const int32_t* adj, const float* xyz, float* out) {
    OMesh m{N, off, adj};
    oracle_compute_neighbor_dist(m, xyz, out);
}

void orc_warp_terrain(int N, const int32_t* off, const int32_t* adj, float* elev, const float* xyz,
                      double seed, double strength, const float* hotspot) {
    OMesh m{N, off, adj};
    oracle_warp_terrain(m, elev, xyz, seed, strength, hotspot);
}

void orc_smooth_elevation(int N, const int32_t* off, const int32_t* adj, float* elev,
                          const uint8_t* isOcean, int iterations, double strength) {
    OMesh m{N, off, adj};
    oracle_smooth_elevation(m, elev, isOcean, iterations, strength);
}

void orc_priority_flood_carve(int N, const int32_t* off, const int32_t* adj, float* elev,
                              const uint8_t* isOcean, double carveStrength, int32_t* drainTo,
                              float* surface, uint8_t* isOpenOcean) {
    OMesh m{N, off, adj};
    FloodDebug d{drainTo, surface, isOpenOcean};
    oracle_priority_flood_carve(m, elev, isOcean, carveStrength, &d);
}

void orc_erode_composite(int N, const int32_t* off, const int32_t* adj, float* elev, const float* xyz,
                         const uint8_t* isOcean, int hIters, double K, double m_, double dt, int tIters,
                         double talus, double kThermal, int gIters, double glacialStrength,
                         const float* neighborDist, int captureIter, int32_t* drainTarget, float* flow,
                         int32_t* landOrder) {
    OMesh m{N, off, adj};
    ErodeDebug d{captureIter, drainTarget, flow, landOrder};
    oracle_erode_composite(m, elev, xyz, isOcean, hIters, K, m_, dt, tIters, talus, kThermal, gIters,
                           glacialStrength, neighborDist, captureIter >= 0 ? &d : nullptr);
}

void orc_sharpen_ridges(int N, const int32_t* off, const int32_t* adj, float* elev,
                        const uint8_t* isOcean, int iterations, double strength) {
    OMesh m{N, off, adj};
    oracle_sharpen_ridges(m, elev, isOcean, iterations, strength);
}

void orc_apply_soil_creep(int N, const int32_t* off, const int32_t* adj, float* elev,
                          const uint8_t* isOcean, int iterations, double strength) {
    OMesh m{N, off, adj};
    oracle_apply_soil_creep(m, elev, isOcean, iterations, strength);
}

void orc_run_post_processing(int N, const int32_t* off, const int32_t* adj, const float* xyz, float* elev,
                             double smoothing, double glacial, double hydraulic, double thermal,
                             double ridge, double warp, int hItersOverride, const float* neighborDist,
                             double seed, const float* hotspot, float* erosionDelta, uint8_t* isOceanOut) {
    OMesh m{N, off, adj};
    PostParams p{smoothing, glacial, hydraulic, thermal, ridge, warp, hItersOverride};
    oracle_run_post_processing(m, xyz, elev, p, neighborDist, seed, hotspot, erosionDelta, isOceanOut);
}

void orc_smooth_field(int N, const int32_t* off, const int32_t* adj, float* field, int passes) {
    OMesh m{N, off, adj};
    oracle_smooth_field(m, field, passes);
}

double orc_percentile(const float* arr, int n, double p) { return oracle_percentile(arr, n, p); }

} // extern "C"
