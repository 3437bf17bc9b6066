// pb_export.h — the equirectangular map export of the render side: exportMap(type, width),
// js/planet-mesh.js:1752-1950, up to the ImageData the reference hands to the canvas (`ctx.putImageData`); the PNG container
// itself is the platform's `canvas.toBlob` there and the host mirror's encoder here (planet_mesh.py).
//
//   reference                                                         here
//   per side s: map triangle (t_xyz[inner t], t_xyz[outer t],         MapRasterK, one side per thread: the three lon/lat pairs in
//     r_xyz[begin r]) in lon/lat space as Float32 positions,            f64 (deterministic atan2 / asin), rounded to f32 like posArr,
//     duplicated −2π when it straddles the date line (:1773-1846)       then both copies are rasterised at once — positions never
//   WebGL: orthographic tiles of ≤ 2048², flat vertex colours,          touch memory.  A pixel belongs to the LAST triangle in draw
//     depth LEQUAL at equal z → the last triangle drawn wins           order whose closed area contains the pixel centre:
//     (:1848-1895)                                                      atomicMax of the draw index (2·s + copy) per pixel
//   rows flipped, linear → sRGB, `* 255 + 0.5 | 0` (:1897-1915)       MapShadeK, one pixel per thread: owner side → region colour →
//                                                                       unorm8 like the render target → sRGB table → RGBA8
// The coverage rule is tile-independent (pixel centres of the whole image, edges inclusive), so the tiling of the reference —
// a memory workaround of the browser — has no counterpart.  All arithmetic is f64 in a fixed order with no contraction, the
// oracle (oracle/export.cpp) states the same rule sequentially; integer `pixelSide` and the RGBA bytes are compared bit for bit.
#pragma once
#include "pb_engine.h"
#include "pb_meshgen.h"
#include "pb_colors.h"

namespace pb {

// equirectangular position of a unit-sphere point, y-up like the renderer (:1794-1796)
struct LonLat { double lon, lat; };
PB_DEV LonLat map_lon_lat(const float* p) {
    const double x = p[0], y = p[1], z = p[2];
    double c = y; c = c < 1 ? c : 1; c = c > -1 ? c : -1;
    return {pb_atan2(x, z), pb_asin(c)};
}
PB_DEV double map_clx(double v) { v = v < 2 ? v : 2; return v > -2 ? v : -2; }
PB_DEV double map_cly(double v) { v = v < 1 ? v : 1; return v > -1 ? v : -1; }

struct MapRasterK {
    const int* tri; const int* half; const float* t_xyz; const float* r_xyz;
    int W, H; int* owner;

    // pixel centres (i + ½, j + ½), j counted from the top row (y = +1); u, v are exact in f64 (f32 position × integer size)
    PB_DEV void raster(const float* X, const float* Y, int key) const {
        double u[3], v[3];
        for (int k = 0; k < 3; k++) { u[k] = ((double)X[k] + 2) * W / 4; v[k] = (1 - (double)Y[k]) * H / 2; }
        const double area = (u[1] - u[0]) * (v[2] - v[0]) - (v[1] - v[0]) * (u[2] - u[0]);
        if (!(area > 0 || area < 0)) return;                      // degenerate: no fragments
        double lo = u[0] < u[1] ? u[0] : u[1]; lo = lo < u[2] ? lo : u[2];
        double hi = u[0] > u[1] ? u[0] : u[1]; hi = hi > u[2] ? hi : u[2];
        int i0 = (int)ceil(lo - 0.5), i1 = (int)floor(hi - 0.5);
        lo = v[0] < v[1] ? v[0] : v[1]; lo = lo < v[2] ? lo : v[2];
        hi = v[0] > v[1] ? v[0] : v[1]; hi = hi > v[2] ? hi : v[2];
        int j0 = (int)ceil(lo - 0.5), j1 = (int)floor(hi - 0.5);
        if (i0 < 0) i0 = 0;
        if (j0 < 0) j0 = 0;
        if (i1 > W - 1) i1 = W - 1;
        if (j1 > H - 1) j1 = H - 1;
        const double ax = u[1] - u[0], ay = v[1] - v[0], bx = u[2] - u[1], by = v[2] - v[1], cx = u[0] - u[2], cy = v[0] - v[2];
        for (int j = j0; j <= j1; j++) {
            const double py = j + 0.5;
            for (int i = i0; i <= i1; i++) {
                const double px = i + 0.5;
                const double e0 = ax * (py - v[0]) - ay * (px - u[0]);
                const double e1 = bx * (py - v[1]) - by * (px - u[1]);
                const double e2 = cx * (py - v[2]) - cy * (px - u[2]);
                const bool in = area > 0 ? (e0 >= 0 && e1 >= 0 && e2 >= 0) : (e0 <= 0 && e1 <= 0 && e2 <= 0);
                if (in) atomic_max(owner + ((size_t)j * (size_t)W + (size_t)i), key);
            }
        }
    }

    PB_DEV void operator()(int s) const {
        const int it = s / 3, ot = half[s] / 3, br = tri[s];
        LonLat p[3] = {map_lon_lat(t_xyz + 3 * (size_t)it), map_lon_lat(t_xyz + 3 * (size_t)ot), map_lon_lat(r_xyz + 3 * (size_t)br)};
        const double sx = 2 / PB_PI;
        double mx = p[0].lon > p[1].lon ? p[0].lon : p[1].lon; mx = mx > p[2].lon ? mx : p[2].lon;
        double mn = p[0].lon < p[1].lon ? p[0].lon : p[1].lon; mn = mn < p[2].lon ? mn : p[2].lon;
        float X[3], Y[3];
        for (int k = 0; k < 3; k++) Y[k] = (float)map_cly(p[k].lat * sx);
        if (mx - mn > PB_PI) {                                     // straddles the date line: drawn twice (:1811-1834)
            for (int k = 0; k < 3; k++) if (p[k].lon < 0) p[k].lon += 2 * PB_PI;
            for (int k = 0; k < 3; k++) X[k] = (float)map_clx(p[k].lon * sx);
            raster(X, Y, 2 * s + 1);
            for (int k = 0; k < 3; k++) X[k] = (float)map_clx((p[k].lon - 2 * PB_PI) * sx);
            raster(X, Y, 2 * s + 2);
        } else {
            for (int k = 0; k < 3; k++) X[k] = (float)map_clx(p[k].lon * sx);
            raster(X, Y, 2 * s + 1);
        }
    }
};

// owner key → RGBA8: the render target holds unorm8 of the (linear) vertex colour, the export applies the sRGB curve to it
struct MapShadeK {
    const int* owner; const int* tri; const float* rgb; uint8_t* rgba; int* pixelSide;
    size_t first;                                              // first pixel of this band
    uint8_t bg[3];
    uint8_t lut[256];
    PB_DEV void operator()(int k) const {
        const size_t px = first + (size_t)k;
        const int key = owner[px];
        uint8_t* o = rgba + 4 * px;
        if (key <= 0) {
            o[0] = bg[0]; o[1] = bg[1]; o[2] = bg[2]; o[3] = 255;
            if (pixelSide) pixelSide[px] = -1;
            return;
        }
        const int s = (key - 1) >> 1, r = tri[s];
        for (int c = 0; c < 3; c++) {
            double v = rgb[3 * (size_t)r + c];
            v = v < 1 ? v : 1; v = v > 0 ? v : 0;
            o[c] = lut[(int)floor(v * 255 + 0.5)];
        }
        o[3] = 255;
        if (pixelSide) pixelSide[px] = s;
    }
};

// linear → sRGB of the 256 render-target levels (:1904-1908), and THREE.Color(0x1a1a2e) in the linear working space
inline void map_srgb_table(uint8_t lut[256]) {
    for (int q = 0; q < 256; q++) {
        const double v = q / 255.0;
        lut[q] = (uint8_t)(int)((v <= 0.0031308 ? v * 12.92 : 1.055 * pb_pow(v, 1 / 2.4) - 0.055) * 255 + 0.5);
    }
}
inline void map_background(bool blackAndWhite, const uint8_t lut[256], uint8_t bg[3]) {
    const int hex[3] = {0x1a, 0x1a, 0x2e};
    for (int k = 0; k < 3; k++) {
        if (blackAndWhite) { bg[k] = lut[0]; continue; }
        const double c = hex[k] / 255.0;
        const double lin = c < 0.04045 ? c * 0.0773993808 : pb_pow(c * 0.9478672986 + 0.0521327014, 2.4);
        bg[k] = lut[(int)floor((double)(float)lin * 255 + 0.5)];
    }
}

struct MapExport {
    DevBuf<int> owner;
    DevBuf<float> tXyz, rgb, rgbRaw;
    DevBuf<uint8_t> sRgba; DevBuf<int> sSide;

    // rgbaDev / sideDev: device pointers (sideDev may be null); regionRgb: 3N floats on the device
    void run(Mesh& m, MeshTriangles& t, const float* regionRgb, bool blackAndWhite, int W, uint8_t* rgbaDev, int* sideDev) {
        const Exec& x = m.ex();
        const int H = W / 2;
        const size_t px = (size_t)W * (size_t)H;
        t.build(x, m.csr(), m.N);
        const size_t T = (size_t)t.T;
        x.for_each((int)T, TriCentersK{t.tri.p, m.xyz.p, tXyz.ensure(3 * T)});
        dev_memset(owner.ensure(px), 0, sizeof(int) * px, x.stream);
        x.for_each((int)(3 * T), MapRasterK{t.tri.p, t.half.p, tXyz.p, m.xyz.p, W, H, owner.p});
        MapShadeK sh{owner.p, t.tri.p, regionRgb, rgbaDev, sideDev, 0, {0, 0, 0}, {0}};
        map_srgb_table(sh.lut);
        map_background(blackAndWhite, sh.lut, sh.bg);
        const size_t band = (size_t)1 << 28;
        for (size_t first = 0; first < px; first += band) {
            sh.first = first;
            x.for_each((int)std::min(band, px - first), sh);
        }
    }
};

}  // namespace pb
