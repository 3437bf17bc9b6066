// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).
// Restates js/rng.js:3-11 and js/simplex-noise.js:5-53.
#pragma once
#include "js_semantics.h"

// js/rng.js:3-6 — Park–Miller LCG; all products < 2^53 so double arithmetic is exact.
struct Rng {
    double s;
    explicit Rng(double seed) {
        s = std::fmod(std::fabs(std::floor(seed * 9301 + 49297)), 2147483646.0) + 1;
    }
    double next() {
        s = std::fmod(s * 16807, 2147483647.0);
        return (s - 1) / 2147483646.0;
    }
};

// js/rng.js:8-11
struct RandInt {
    Rng r;
    explicit RandInt(double seed) : r(seed) {}
    int64_t operator()(double n) { return (int64_t)std::floor(r.next() * n); }
};

// js/simplex-noise.js:5-53
struct SimplexNoise {
    uint8_t perm[512];
    uint8_t pm12[512];
    explicit SimplexNoise(double seed = 0) {
        Rng rng(seed);
        uint8_t p[256];
        for (int i = 0; i < 256; i++) p[i] = (uint8_t)i;
        for (int i = 255; i > 0; i--) {
            int j = (int)std::floor(rng.next() * (i + 1));
            uint8_t t = p[i]; p[i] = p[j]; p[j] = t;
        }
        for (int i = 0; i < 512; i++) { perm[i] = p[i & 255]; pm12[i] = perm[i] % 12; }
    }
    static const int8_t* grad(int g) {
        static const int8_t G[12][3] = {{1,1,0},{-1,1,0},{1,-1,0},{-1,-1,0},{1,0,1},{-1,0,1},
                                        {1,0,-1},{-1,0,-1},{0,1,1},{0,-1,1},{0,1,-1},{0,-1,-1}};
        return G[g];
    }
    // js/simplex-noise.js:17-32
    double noise3D(double x, double y, double z) const {
        const double F = 1.0 / 3, H = 1.0 / 6;
        double s = (x + y + z) * F;
        double i = std::floor(x + s), j = std::floor(y + s), k = std::floor(z + s);
        double t = (i + j + k) * H, x0 = x - i + t, y0 = y - j + t, z0 = z - k + t;
        int i1, j1, k1, i2, j2, k2;
        if (x0 >= y0) {
            if (y0 >= z0) { i1=1;j1=0;k1=0;i2=1;j2=1;k2=0; }
            else if (x0 >= z0) { i1=1;j1=0;k1=0;i2=1;j2=0;k2=1; }
            else { i1=0;j1=0;k1=1;i2=1;j2=0;k2=1; }
        } else {
            if (y0 < z0) { i1=0;j1=0;k1=1;i2=0;j2=1;k2=1; }
            else if (x0 < z0) { i1=0;j1=1;k1=0;i2=0;j2=1;k2=1; }
            else { i1=0;j1=1;k1=0;i2=1;j2=1;k2=0; }
        }
        double x1 = x0 - i1 + H, y1 = y0 - j1 + H, z1 = z0 - k1 + H;
        double x2 = x0 - i2 + 2 * H, y2 = y0 - j2 + 2 * H, z2 = z0 - k2 + 2 * H;
        double x3 = x0 - 1 + 3 * H, y3 = y0 - 1 + 3 * H, z3 = z0 - 1 + 3 * H;
        // i & 255 on a double: ToInt32 then mask (|i| is tiny here, but keep the semantics)
        int ii = js::to_int32(i) & 255, jj = js::to_int32(j) & 255, kk = js::to_int32(k) & 255;
        double n0 = 0, n1 = 0, n2 = 0, n3 = 0;
        double a = 0.6 - x0 * x0 - y0 * y0 - z0 * z0;
        if (a > 0) { a *= a; const int8_t* v = grad(pm12[ii + perm[jj + perm[kk]]]); n0 = a * a * (v[0] * x0 + v[1] * y0 + v[2] * z0); }
        double b = 0.6 - x1 * x1 - y1 * y1 - z1 * z1;
        if (b > 0) { b *= b; const int8_t* v = grad(pm12[ii + i1 + perm[jj + j1 + perm[kk + k1]]]); n1 = b * b * (v[0] * x1 + v[1] * y1 + v[2] * z1); }
        double c = 0.6 - x2 * x2 - y2 * y2 - z2 * z2;
        if (c > 0) { c *= c; const int8_t* v = grad(pm12[ii + i2 + perm[jj + j2 + perm[kk + k2]]]); n2 = c * c * (v[0] * x2 + v[1] * y2 + v[2] * z2); }
        double d = 0.6 - x3 * x3 - y3 * y3 - z3 * z3;
        if (d > 0) { d *= d; const int8_t* v = grad(pm12[ii + 1 + perm[jj + 1 + perm[kk + 1]]]); n3 = d * d * (v[0] * x3 + v[1] * y3 + v[2] * z3); }
        return 32 * (n0 + n1 + n2 + n3);
    }
    // js/simplex-noise.js:34-38
    double fbm(double x, double y, double z, int octaves = 5, double persistence = 2.0 / 3) const {
        double sum = 0, max = 0, amp = 1;
        for (int o = 0; o < octaves; o++) {
            double f = (double)(1 << o);
            sum += amp * noise3D(x * f, y * f, z * f);
            max += amp;
            amp *= persistence;
        }
        return sum / max;
    }
    // js/simplex-noise.js:40-53
    double ridgedFbm(double x, double y, double z, int octaves = 6, double lacunarity = 2.0,
                     double gain = 0.5, double offset = 1.0) const {
        double sum = 0, freq = 1, amp = 1, prev = 1, maxVal = 0;
        for (int o = 0; o < octaves; o++) {
            double n = noise3D(x * freq, y * freq, z * freq);
            n = offset - std::fabs(n);
            n = n * n;
            sum += n * amp * prev;
            maxVal += amp;
            prev = js::min(n, 1);
            freq *= lacunarity;
            amp *= gain;
        }
        return sum / maxVal;
    }
};
