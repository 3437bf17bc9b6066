// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).
#pragma once
#include "js_semantics.h"

struct FloodDebug {        // optional taps on priorityFloodCarve internals (after pass 1)
    int32_t* drainTo;
    float* surface;
    uint8_t* isOpenOcean;
};

struct ErodeDebug {        // optional taps on one hydraulic iteration (before the implicit solve)
    int32_t captureIter;
    int32_t* drainTarget;
    float* flow;
    int32_t* landOrder;    // landCells after the sort used by that iteration
};

struct PostParams {        // slider values, js/planet-worker.js:41
    double smoothing, glacialErosion, hydraulicErosion, thermalErosion, ridgeSharpening, terrainWarp;
    int32_t hItersOverride; // < 0: use round(20*hydraulicErosion)
};

double oracle_cell_noise(double r);
void oracle_priority_flood_carve(const OMesh& mesh, float* r_elevation, const uint8_t* r_isOcean,
                                 double carveStrength, FloodDebug* dbg);
void oracle_warp_terrain(const OMesh& mesh, float* r_elevation, const float* r_xyz, double seed,
                         double strength, const float* r_hotspot);
void oracle_smooth_elevation(const OMesh& mesh, float* r_elevation, const uint8_t* r_isOcean,
                             int iterations, double strength);
void oracle_erode_composite(const OMesh& mesh, float* r_elevation, const float* r_xyz,
                            const uint8_t* r_isOcean, int hIters, double K, double m, double dt,
                            int tIters, double talusSlope, double kThermal, int gIters,
                            double glacialStrength, const float* neighborDist, ErodeDebug* dbg);
void oracle_sharpen_ridges(const OMesh& mesh, float* r_elevation, const uint8_t* r_isOcean,
                           int iterations, double strength);
void oracle_apply_soil_creep(const OMesh& mesh, float* r_elevation, const uint8_t* r_isOcean,
                             int iterations, double strength);
void oracle_run_post_processing(const OMesh& mesh, const float* r_xyz, float* r_elevation,
                                const PostParams& p, const float* neighborDist, double seed,
                                const float* r_hotspot, float* erosionDelta, uint8_t* isOceanOut);
void oracle_compute_neighbor_dist(const OMesh& mesh, const float* r_xyz, float* neighborDist);
