// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).
// Restates js/sphere-mesh.js:9-37 (generateFibonacciSphere).  asin/sin/cos go through
// pb_detmath.h so the product's host-side generator yields bit-identical f32 coordinates.
#pragma once
#include "js_semantics.h"
#include "noise.h"

inline void oracle_fibonacci_sphere(int N, double jitter, double seed, float* r_xyz) {
    Rng rng(seed);
    const double s = 3.6 / std::sqrt((double)N);
    const double dlong = PB_PI * (3 - std::sqrt(5.0));
    const double dz = 2.0 / N;
    double lng = 0, z = 1 - dz / 2;
    for (int k = 0; k < N; k++, z -= dz) {
        double r = std::sqrt(1 - z * z);
        double latDeg = pb_asin(z) * 180 / PB_PI;
        double lonDeg = lng * 180 / PB_PI;
        if (jitter > 0) {
            double a = rng.next(); double b = rng.next();
            double jLat = a - b;
            double c = rng.next(); double d = rng.next();
            double jLon = c - d;
            double nextZ = js::max(-1, z - dz * 2 * PB_PI * r / s);
            latDeg += jitter * jLat * (latDeg - pb_asin(nextZ) * 180 / PB_PI);
            lonDeg += jitter * jLon * (s / r * 180 / PB_PI);
        }
        double latR = latDeg * PB_PI / 180;
        double lonR = lonDeg * PB_PI / 180;
        r_xyz[3 * k]     = js::f32(pb_cos(latR) * pb_cos(lonR));
        r_xyz[3 * k + 1] = js::f32(pb_cos(latR) * pb_sin(lonR));
        r_xyz[3 * k + 2] = js::f32(pb_sin(latR));
        lng += dlong;
    }
}
