// pb_noise.h — seeded 3-D simplex noise, fBm and ridged fBm as device functions (kernel family K4).
// Mirrors js/simplex-noise.js:5-53 and js/rng.js:3-11 of the reference: FP64 arithmetic using only
// + - * floor, so results are bit-identical to the JS doubles.
#pragma once
#include "pb_platform.h"

namespace pb {

// js/rng.js:3-6 — Park–Miller; every product stays below 2^53, so double arithmetic is exact.
struct ParkMiller {
    double s;
    explicit ParkMiller(double seed) { s = fmod(fabs(floor(seed * 9301.0 + 49297.0)), 2147483646.0) + 1.0; }
    double next() {
        s = fmod(s * 16807.0, 2147483647.0);
        return (s - 1.0) / 2147483646.0;
    }
};

// Host-side table construction (js/simplex-noise.js:6-15).  Layout: perm[0..511] then pm12[0..511].
struct SimplexTable {
    uint8_t t[1024];
    explicit SimplexTable(double seed) {
        ParkMiller rng(seed);
        uint8_t p[256];
        for (int i = 0; i < 256; i++) p[i] = (uint8_t)i;
        for (int i = 255; i > 0; i--) {
            int j = (int)floor(rng.next() * (double)(i + 1));
            uint8_t tmp = p[i];
            p[i] = p[j];
            p[j] = tmp;
        }
        for (int i = 0; i < 512; i++) {
            t[i] = p[i & 255];
            t[512 + i] = (uint8_t)(t[i] % 12);
        }
    }
};

// Device view of one table (1 KiB in global memory; L1-resident after the first few cells).
struct Simplex {
    const uint8_t* tab;

    PB_HDEV int P(int i) const { return tab[i]; }
    PB_HDEV int M(int i) const { return tab[512 + i]; }

    // gradient g·(x,y,z) with the reference's table order (js/simplex-noise.js:7); the products by
    // 0/±1 are kept so that signed zeros come out as in JS.
    static PB_HDEV double gdot(int g, double x, double y, double z) {
        const double s1 = (g & 1) ? -1.0 : 1.0;
        const double s2 = (g & 2) ? -1.0 : 1.0;
        const int grp = g >> 2;
        double v0, v1, v2;
        if (grp == 0) { v0 = s1; v1 = s2; v2 = 0.0; }
        else if (grp == 1) { v0 = s1; v1 = 0.0; v2 = s2; }
        else { v0 = 0.0; v1 = s1; v2 = s2; }
        return v0 * x + v1 * y + v2 * z;
    }

    static PB_HDEV double corner(double x, double y, double z, int g) {
        double a = 0.6 - x * x - y * y - z * z;
        if (a > 0) {
            a *= a;
            return a * a * gdot(g, x, y, z);
        }
        return 0.0;
    }

    // js/simplex-noise.js:17-32
    PB_HDEV double noise3D(double x, double y, double z) const {
        const double F = 1.0 / 3.0, H = 1.0 / 6.0;
        const double s = (x + y + z) * F;
        const double i = floor(x + s), j = floor(y + s), k = floor(z + s);
        const double t = (i + j + k) * H;
        const double x0 = x - i + t, y0 = y - j + t, z0 = z - k + t;
        // simplex ordering → offsets of the 2nd and 3rd corner
        int i1, j1, k1, i2, j2, k2;
        if (x0 >= y0) {
            if (y0 >= z0)      { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
            else if (x0 >= z0) { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 0; k2 = 1; }
            else               { i1 = 0; j1 = 0; k1 = 1; i2 = 1; j2 = 0; k2 = 1; }
        } else {
            if (y0 < z0)       { i1 = 0; j1 = 0; k1 = 1; i2 = 0; j2 = 1; k2 = 1; }
            else if (x0 < z0)  { i1 = 0; j1 = 1; k1 = 0; i2 = 0; j2 = 1; k2 = 1; }
            else               { i1 = 0; j1 = 1; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
        }
        const double H2 = 2.0 * H, H3 = 3.0 * H;
        const double x1 = x0 - i1 + H, y1 = y0 - j1 + H, z1 = z0 - k1 + H;
        const double x2 = x0 - i2 + H2, y2 = y0 - j2 + H2, z2 = z0 - k2 + H2;
        const double x3 = x0 - 1.0 + H3, y3 = y0 - 1.0 + H3, z3 = z0 - 1.0 + H3;
        // i & 255 in JS is ToInt32 then mask; lattice indices here are far below 2^31
        const int ii = ((int)(long long)i) & 255, jj = ((int)(long long)j) & 255, kk = ((int)(long long)k) & 255;
        const double n0 = corner(x0, y0, z0, M(ii + P(jj + P(kk))));
        const double n1 = corner(x1, y1, z1, M(ii + i1 + P(jj + j1 + P(kk + k1))));
        const double n2 = corner(x2, y2, z2, M(ii + i2 + P(jj + j2 + P(kk + k2))));
        const double n3 = corner(x3, y3, z3, M(ii + 1 + P(jj + 1 + P(kk + 1))));
        return 32.0 * (n0 + n1 + n2 + n3);
    }

    // js/simplex-noise.js:34-38
    PB_HDEV double fbm(double x, double y, double z, int octaves, double persistence) const {
        double sum = 0, mx = 0, amp = 1;
        for (int o = 0; o < octaves; o++) {
            const double f = (double)(1 << o);
            sum += amp * noise3D(x * f, y * f, z * f);
            mx += amp;
            amp *= persistence;
        }
        return sum / mx;
    }
    PB_HDEV double fbm(double x, double y, double z, int octaves) const { return fbm(x, y, z, octaves, 2.0 / 3.0); }

    // js/simplex-noise.js:40-53
    PB_HDEV double ridgedFbm(double x, double y, double z, int octaves, double lacunarity, double gain,
                            double offset) const {
        double sum = 0, freq = 1, amp = 1, prev = 1, maxVal = 0;
        for (int o = 0; o < octaves; o++) {
            double n = noise3D(x * freq, y * freq, z * freq);
            n = offset - fabs(n);
            n = n * n;
            sum += n * amp * prev;
            maxVal += amp;
            prev = n < 1.0 ? n : 1.0;
            freq *= lacunarity;
            amp *= gain;
        }
        return sum / maxVal;
    }
    PB_HDEV double ridgedFbm(double x, double y, double z) const { return ridgedFbm(x, y, z, 6, 2.0, 0.5, 1.0); }
};

}  // namespace pb
