// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).  Pinned against the reference's own source run under tests/golden/minijs.py (scenario F_render_600).
// Per-region colour ramps of the render / export side:
//   elevToHeightKm, biomeColor, elevationToColor   js/color-map.js:7-12, 73-125
//   smoothBiomeColors, heightmapColor, landHeightmapColor, landMaskColor   js/planet-mesh.js:30-80
#include <cmath>
#include <cstdint>
#include <vector>

#include "js_semantics.h"

namespace {

double height_km(double elev) {                         // js/color-map.js:7-12
    if (elev <= 0) return elev * 10;
    const double t = js::min(elev, 1), t2 = t * t;
    return 6 * t2 * t2 * (5 - 4 * t);
}
void elevation_to_color(double e, double c[3]) {        // js/color-map.js:116-125
    auto set = [&](double r, double g, double b) { c[0] = r; c[1] = g; c[2] = b; };
    if (e < -0.50) return set(0.04, 0.06, 0.30);
    if (e < -0.10) { const double t = (e + 0.50) / 0.40; return set(0.04 + t * 0.07, 0.06 + t * 0.14, 0.30 + t * 0.18); }
    if (e < 0.00) { const double t = (e + 0.10) / 0.10; return set(0.11 + t * 0.19, 0.20 + t * 0.22, 0.48 + t * 0.12); }
    if (e < 0.03) { const double t = e / 0.03; return set(0.72 + t * 0.08, 0.68 - t * 0.02, 0.46 - t * 0.10); }
    if (e < 0.25) { const double t = (e - 0.03) / 0.22; return set(0.20 - t * 0.06, 0.54 - t * 0.12, 0.12 + t * 0.08); }
    if (e < 0.50) { const double t = (e - 0.25) / 0.25; return set(0.14 + t * 0.30, 0.42 - t * 0.14, 0.20 - t * 0.06); }
    if (e < 0.75) { const double t = (e - 0.50) / 0.25; return set(0.44 + t * 0.16, 0.28 + t * 0.12, 0.14 + t * 0.18); }
    const double t = js::min(1, (e - 0.75) / 0.20);
    set(0.60 + t * 0.35, 0.40 + t * 0.50, 0.32 + t * 0.60);
}
const double BIOME[31][3] = {
    {0, 0, 0}, {0.05, 0.30, 0.05}, {0.08, 0.33, 0.07}, {0.42, 0.50, 0.18}, {0.82, 0.72, 0.50}, {0.60, 0.55, 0.48}, {0.72, 0.62, 0.30},
    {0.55, 0.52, 0.32}, {0.18, 0.42, 0.12}, {0.12, 0.38, 0.10}, {0.10, 0.28, 0.10}, {0.45, 0.48, 0.22}, {0.40, 0.45, 0.20},
    {0.35, 0.40, 0.20}, {0.20, 0.44, 0.14}, {0.15, 0.40, 0.12}, {0.12, 0.32, 0.10}, {0.12, 0.36, 0.08}, {0.10, 0.32, 0.08},
    {0.06, 0.22, 0.08}, {0.05, 0.18, 0.07}, {0.38, 0.38, 0.18}, {0.35, 0.35, 0.17}, {0.08, 0.22, 0.08}, {0.06, 0.18, 0.07},
    {0.14, 0.36, 0.10}, {0.12, 0.32, 0.09}, {0.07, 0.22, 0.08}, {0.05, 0.18, 0.07}, {0.35, 0.32, 0.22}, {0.78, 0.80, 0.84}};
void thresholds(int id, double& alpine, double& snow) {  // js/color-map.js:57-68
    if (id <= 0) { alpine = 0; snow = 0; }
    else if (id <= 3) { alpine = 3.5; snow = 5.5; }
    else if (id <= 7) { alpine = 3.0; snow = 5.0; }
    else if (id <= 16) { alpine = 2.0; snow = 3.5; }
    else if (id <= 18 || id == 21 || id == 22 || id == 25 || id == 26) { alpine = 1.5; snow = 3.0; }
    else if (id <= 28) { alpine = 0.8; snow = 2.0; }
    else if (id == 29) { alpine = 0.4; snow = 1.5; }
    else { alpine = 0; snow = 0.5; }
}
void biome_color(int koppen, double elevation, double c[3]) {   // js/color-map.js:73-114
    if (koppen == 0 || elevation <= 0) return elevation_to_color(elevation, c);
    double r, g, b;
    if (koppen >= 1 && koppen <= 30) { r = BIOME[koppen][0]; g = BIOME[koppen][1]; b = BIOME[koppen][2]; }
    else { r = 0.30; g = 0.50; b = 0.20; }
    const double hKm = height_km(elevation);
    double alpine, snow;
    thresholds(koppen, alpine, snow);
    if (hKm < 0.2) { const double dark = 0.93 + 0.07 * (hKm / 0.2); r *= dark; g *= dark; b *= dark; }
    if (alpine > 0 && hKm > 0.2 && hKm < alpine) {
        const double t = (hKm - 0.2) / (alpine - 0.2), darken = 1.0 - t * 0.15;
        r *= darken; g *= darken; b *= darken;
    }
    if (alpine > 0 && hKm > alpine) {
        const double rockZone = snow > alpine ? snow - alpine : 2.0;
        const double rockT = js::min(1, (hKm - alpine) / rockZone), s = rockT * rockT;
        r = r + (0.42 - r) * s; g = g + (0.38 - g) * s; b = b + (0.32 - b) * s;
    }
    if (snow > 0 && hKm > snow) {
        const double snowT = js::min(1, (hKm - snow) / 2.5), s = snowT * snowT;
        r = r + (0.92 - r) * s; g = g + (0.93 - g) * s; b = b + (0.96 - b) * s;
    }
    c[0] = r; c[1] = g; c[2] = b;
}

const double KOPPEN[31][3] = {                           // KOPPEN_CLASSES[i].color, js/koppen.js:19-51
    {0.29, 0.44, 0.65}, {0.00, 0.00, 1.00}, {0.00, 0.47, 1.00}, {0.27, 0.67, 0.98}, {1.00, 0.00, 0.00}, {1.00, 0.59, 0.59},
    {0.96, 0.65, 0.00}, {1.00, 0.86, 0.39}, {0.78, 1.00, 0.31}, {0.39, 1.00, 0.31}, {0.20, 0.78, 0.00}, {1.00, 1.00, 0.00},
    {0.78, 0.78, 0.00}, {0.59, 0.59, 0.00}, {0.59, 1.00, 0.59}, {0.39, 0.78, 0.39}, {0.20, 0.59, 0.20}, {0.00, 1.00, 1.00},
    {0.22, 0.78, 1.00}, {0.00, 0.49, 0.49}, {0.00, 0.27, 0.37}, {0.90, 0.50, 1.00}, {0.70, 0.35, 0.85}, {0.50, 0.20, 0.65},
    {0.35, 0.10, 0.45}, {0.67, 0.69, 1.00}, {0.43, 0.47, 0.78}, {0.29, 0.31, 0.78}, {0.20, 0.00, 0.53}, {0.70, 0.70, 0.70},
    {0.41, 0.41, 0.41}};

}  // namespace

extern "C" {
// mode: 0 terrain (elevationToColor), 1 biome smoothed (smoothBiomeColors), 2 heightmap, 3 land heightmap, 4 land mask,
//       5 biome unsmoothed (biomeColor), 6 koppenColor (js/planet-mesh.js:175-178: `KOPPEN_CLASSES[classId] || KOPPEN_CLASSES[0]`)
void orc_region_colors(int N, const int32_t* off, const int32_t* adj, int mode, const float* elev, const uint8_t* koppen, float* rgb) {
    std::vector<float> raw;
    if (mode == 1) raw.resize(3 * (size_t)N);
    for (int r = 0; r < N; r++) {
        double c[3] = {0, 0, 0};
        const double e = elev[r];
        if (mode == 0) elevation_to_color(e, c);
        else if (mode == 1 || mode == 5) biome_color(koppen[r], e, c);
        else if (mode == 2) { const double t = js::max(0, js::min(1, (height_km(e) + 5) / 11)); c[0] = c[1] = c[2] = t; }
        else if (mode == 3) { if (e > 0) { const double t = js::max(0, js::min(1, height_km(e) / 6)); c[0] = c[1] = c[2] = t; } }
        else if (mode == 4) { if (e > 0) c[0] = c[1] = c[2] = 1; }
        else if (mode == 6) { const int id = koppen[r] <= 30 ? koppen[r] : 0; for (int k = 0; k < 3; k++) c[k] = KOPPEN[id][k]; }
        float* dst = mode == 1 ? raw.data() : rgb;
        for (int k = 0; k < 3; k++) dst[3 * r + k] = js::f32(c[k]);
    }
    if (mode != 1) return;
    const double alpha = 0.35;                           // js/planet-mesh.js:30-62
    for (int r = 0; r < N; r++) {
        const int start = off[r], end = off[r + 1], count = end - start;
        if (count == 0) { for (int k = 0; k < 3; k++) rgb[3 * r + k] = raw[3 * r + k]; continue; }
        double avg[3] = {0, 0, 0};
        for (int i = start; i < end; i++) for (int k = 0; k < 3; k++) avg[k] += raw[3 * adj[i] + k];
        for (int k = 0; k < 3; k++) { avg[k] /= count; rgb[3 * r + k] = js::f32(raw[3 * r + k] * (1 - alpha) + avg[k] * alpha); }
    }
}
}
