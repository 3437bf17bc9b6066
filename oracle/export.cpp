// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).  Triangle stage pinned against the reference's own loop run under tests/golden/minijs.py (scenario F_render_600); the framebuffer rule is this repository's.
// exportMap(type, width), js/planet-mesh.js:1752-1950, restated up to the ImageData that is put on the canvas:
//   :1773-1846  one map triangle per side (two when it straddles the date line) into Float32 posArr / colArr
//   :1848-1895  THREE.Mesh with MeshBasicMaterial({vertexColors, DoubleSide}) rendered through orthographic tiles
//   :1897-1915  rows flipped, linear → sRGB, `* 255 + 0.5 | 0`
// The WebGL rasteriser is the platform, not the reference's code; what this file states for it is the rule written in
// include/planet_b200.h (pb_export_map): triangles are drawn in array order, a fragment is produced for every pixel whose
// centre lies in the closed triangle, fragments at equal depth replace earlier ones (depthFunc LessEqualDepth), the render
// target stores unorm8 = floor(c·255 + ½) of the interpolated (here flat) colour, and the tiling does not change which
// pixel centres are covered.  generateTriangleCenters: js/sphere-mesh.js:206-219.
#include <cmath>
#include <cstdint>
#include <vector>

#include "js_semantics.h"

extern "C" void orc_region_colors(int N, const int32_t* off, const int32_t* adj, int mode, const float* elev, const uint8_t* koppen, float* rgb);

namespace {

struct Framebuffer {
    int W, H;
    std::vector<uint8_t> px;          // unorm8 r,g,b per pixel, row 0 = top
    std::vector<int32_t> side;
    void clear(double r, double g, double b) {
        const uint8_t q[3] = {(uint8_t)std::floor((double)(float)r * 255 + 0.5), (uint8_t)std::floor((double)(float)g * 255 + 0.5),
                              (uint8_t)std::floor((double)(float)b * 255 + 0.5)};
        for (size_t i = 0; i < (size_t)W * H; i++) { px[3 * i] = q[0]; px[3 * i + 1] = q[1]; px[3 * i + 2] = q[2]; side[i] = -1; }
    }
    // one triangle of the position / colour buffers (9 floats each); the flat colour is that of its first vertex
    void draw(const float* pos, const float* col, int s) {
        double u[3], v[3];
        for (int k = 0; k < 3; k++) {
            u[k] = ((double)pos[3 * k] + 2) * W / 4;          // map x ∈ [-2, 2] → pixel columns
            v[k] = (1 - (double)pos[3 * k + 1]) * H / 2;      // map y ∈ [-1, 1] → pixel rows, top row first
        }
        const double area = (u[1] - u[0]) * (v[2] - v[0]) - (v[1] - v[0]) * (u[2] - u[0]);
        if (area == 0 || area != area) return;
        const double uMin = std::fmin(u[0], std::fmin(u[1], u[2])), uMax = std::fmax(u[0], std::fmax(u[1], u[2]));
        const double vMin = std::fmin(v[0], std::fmin(v[1], v[2])), vMax = std::fmax(v[0], std::fmax(v[1], v[2]));
        const int i0 = std::max(0, (int)std::ceil(uMin - 0.5)), i1 = std::min(W - 1, (int)std::floor(uMax - 0.5));
        const int j0 = std::max(0, (int)std::ceil(vMin - 0.5)), j1 = std::min(H - 1, (int)std::floor(vMax - 0.5));
        uint8_t q[3];
        for (int c = 0; c < 3; c++) {
            const double lin = js::min(1, js::max(0, (double)col[c]));
            q[c] = (uint8_t)std::floor(lin * 255 + 0.5);
        }
        for (int j = j0; j <= j1; j++)
            for (int i = i0; i <= i1; i++) {
                const double x = i + 0.5, y = j + 0.5;
                double e[3];
                for (int k = 0; k < 3; k++) {
                    const int n = (k + 1) % 3;
                    e[k] = (u[n] - u[k]) * (y - v[k]) - (v[n] - v[k]) * (x - u[k]);
                }
                const bool inside = area > 0 ? (e[0] >= 0 && e[1] >= 0 && e[2] >= 0) : (e[0] <= 0 && e[1] <= 0 && e[2] <= 0);
                if (!inside) continue;
                const size_t p = (size_t)j * W + i;
                px[3 * p] = q[0]; px[3 * p + 1] = q[1]; px[3 * p + 2] = q[2]; side[p] = s;
            }
    }
};

double srgb_to_linear(double c) {       // THREE.Color.setHex under ColorManagement: sRGB hex → linear working space
    return c < 0.04045 ? c * 0.0773993808 : pb_pow(c * 0.9478672986 + 0.0521327014, 2.4);
}

}  // namespace

extern "C" {

// generateTriangleCenters(mesh, r_xyz) js/sphere-mesh.js:206-219
void orc_triangle_centers(int numTriangles, const int32_t* triangles, const float* r_xyz, float* t_xyz) {
    for (int t = 0; t < numTriangles; t++) {
        const int a = triangles[3 * t], b = triangles[3 * t + 1], c = triangles[3 * t + 2];
        for (int k = 0; k < 3; k++)
            t_xyz[3 * t + k] = js::f32(((double)r_xyz[3 * a + k] + (double)r_xyz[3 * b + k] + (double)r_xyz[3 * c + k]) / 3);
    }
}

// the triangle stage of exportMap (:1766-1846): Float32 position / colour buffers (9 floats per triangle, capacity 2 * numSides
// triangles) and, per triangle, the side it came from; returns triCount
int orc_export_map_triangles(int N, const int32_t* off, const int32_t* adj, int numSides, const int32_t* triangles, const int32_t* halfedges,
                             const float* r_xyz, int colorMode, const float* r_elevation, const uint8_t* r_koppen,
                             float* posArr, float* colArr, int32_t* triSide) {
    const int numTriangles = numSides / 3;
    std::vector<float> t_xyz(3 * (size_t)numTriangles), regionColor(3 * (size_t)N);
    orc_triangle_centers(numTriangles, triangles, r_xyz, t_xyz.data());
    orc_region_colors(N, off, adj, colorMode, r_elevation, r_koppen, regionColor.data());

    const double PI = PB_PI, sx = 2 / PI;
    size_t triCount = 0;
    auto clx = [](double v) { return js::max(-2, js::min(2, v)); };
    auto cly = [](double v) { return js::max(-1, js::min(1, v)); };
    for (int s = 0; s < numSides; s++) {
        const int it = s / 3, ot = halfedges[s] / 3, br = triangles[s];
        const float cr = regionColor[3 * br], cg = regionColor[3 * br + 1], cb = regionColor[3 * br + 2];
        const double x0 = t_xyz[3 * it], y0 = t_xyz[3 * it + 1], z0 = t_xyz[3 * it + 2];
        const double x1 = t_xyz[3 * ot], y1 = t_xyz[3 * ot + 1], z1 = t_xyz[3 * ot + 2];
        const double x2 = r_xyz[3 * br], y2 = r_xyz[3 * br + 1], z2 = r_xyz[3 * br + 2];
        double lon0 = pb_atan2(x0, z0), lat0 = pb_asin(js::max(-1, js::min(1, y0)));
        double lon1 = pb_atan2(x1, z1), lat1 = pb_asin(js::max(-1, js::min(1, y1)));
        double lon2 = pb_atan2(x2, z2), lat2 = pb_asin(js::max(-1, js::min(1, y2)));
        const double maxLon = js::max(lon0, js::max(lon1, lon2)), minLon = js::min(lon0, js::min(lon1, lon2));
        const bool wraps = (maxLon - minLon) > PI;
        auto emit = [&](double a0, double a1, double a2) {
            const size_t o = triCount * 9;
            posArr[o] = js::f32(clx(a0 * sx)); posArr[o + 1] = js::f32(cly(lat0 * sx)); posArr[o + 2] = 0;
            posArr[o + 3] = js::f32(clx(a1 * sx)); posArr[o + 4] = js::f32(cly(lat1 * sx)); posArr[o + 5] = 0;
            posArr[o + 6] = js::f32(clx(a2 * sx)); posArr[o + 7] = js::f32(cly(lat2 * sx)); posArr[o + 8] = 0;
            for (int k = 0; k < 3; k++) { colArr[o + 3 * k] = cr; colArr[o + 3 * k + 1] = cg; colArr[o + 3 * k + 2] = cb; }
            triSide[triCount] = s;
            triCount++;
        };
        if (wraps) {
            if (lon0 < 0) lon0 += 2 * PI;
            if (lon1 < 0) lon1 += 2 * PI;
            if (lon2 < 0) lon2 += 2 * PI;
            emit(lon0, lon1, lon2);
            emit(lon0 - 2 * PI, lon1 - 2 * PI, lon2 - 2 * PI);
        } else {
            emit(lon0, lon1, lon2);
        }
    }
    return (int)triCount;
}

// colorMode as orc_region_colors (0 colormap, 1 biome = Satellite, 2 heightmap, 3 landheightmap, 4 landmask, 6 koppen)
void orc_export_map(int N, const int32_t* off, const int32_t* adj, int numSides, const int32_t* triangles, const int32_t* halfedges,
                    const float* r_xyz, int colorMode, int width, const float* r_elevation, const uint8_t* r_koppen,
                    uint8_t* rgba, int32_t* pixelSide) {
    const int height = width / 2;
    const bool isBW = colorMode == 2 || colorMode == 3 || colorMode == 4;
    std::vector<float> posArr((size_t)numSides * 18), colArr((size_t)numSides * 18);
    std::vector<int32_t> triSide((size_t)numSides * 2);
    const size_t triCount = (size_t)orc_export_map_triangles(N, off, adj, numSides, triangles, halfedges, r_xyz, colorMode, r_elevation, r_koppen,
                                                             posArr.data(), colArr.data(), triSide.data());

    Framebuffer fb{width, height, std::vector<uint8_t>(3 * (size_t)width * height), std::vector<int32_t>((size_t)width * height)};
    if (isBW) fb.clear(0, 0, 0);
    else fb.clear(srgb_to_linear(0x1a / 255.0), srgb_to_linear(0x1a / 255.0), srgb_to_linear(0x2e / 255.0));
    for (size_t t = 0; t < triCount; t++) {
        // positions pass through Float32 with z = 0; a triangle's three vertices carry the same colour
        float p2[9];
        for (int k = 0; k < 3; k++) { p2[3 * k] = posArr[t * 9 + 3 * k]; p2[3 * k + 1] = posArr[t * 9 + 3 * k + 1]; p2[3 * k + 2] = 0; }
        fb.draw(p2, &colArr[t * 9], triSide[t]);
    }
    // :1897-1915 (the row flip is already in the top-first row order)
    for (size_t p = 0; p < (size_t)width * height; p++) {
        for (int c = 0; c < 3; c++) {
            const double v = fb.px[3 * p + c] / 255.0;
            rgba[4 * p + c] = (uint8_t)js::to_int32(std::floor((v <= 0.0031308 ? v * 12.92 : 1.055 * pb_pow(v, 1 / 2.4) - 0.055) * 255 + 0.5));
        }
        rgba[4 * p + 3] = 255;
        if (pixelSide) pixelSide[p] = fb.side[p];
    }
}

}
