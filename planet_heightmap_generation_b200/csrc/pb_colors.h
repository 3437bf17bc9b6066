// pb_colors.h — per-region colour ramps of the render / export side.
//   elevationToColor, biomeColor                          js/color-map.js:73-125   (elevToHeightKm: pb_climate.h)
//   smoothBiomeColors, heightmapColor, landHeightmapColor, landMaskColor   js/planet-mesh.js:30-80
// The JS functions return doubles that end up in Float32 colour buffers: the kernels compute in f64 and round once.
#pragma once
#include "pb_engine.h"

namespace pb {

struct Rgb { double r, g, b; };

PB_DEV Rgb terrain_color(double e) {                     // elevationToColor
    if (e < -0.50) return {0.04, 0.06, 0.30};
    if (e < -0.10) { const double t = (e + 0.50) / 0.40; return {0.04 + t * 0.07, 0.06 + t * 0.14, 0.30 + t * 0.18}; }
    if (e < 0.00) { const double t = (e + 0.10) / 0.10; return {0.11 + t * 0.19, 0.20 + t * 0.22, 0.48 + t * 0.12}; }
    if (e < 0.03) { const double t = e / 0.03; return {0.72 + t * 0.08, 0.68 - t * 0.02, 0.46 - t * 0.10}; }
    if (e < 0.25) { const double t = (e - 0.03) / 0.22; return {0.20 - t * 0.06, 0.54 - t * 0.12, 0.12 + t * 0.08}; }
    if (e < 0.50) { const double t = (e - 0.25) / 0.25; return {0.14 + t * 0.30, 0.42 - t * 0.14, 0.20 - t * 0.06}; }
    if (e < 0.75) { const double t = (e - 0.50) / 0.25; return {0.44 + t * 0.16, 0.28 + t * 0.12, 0.14 + t * 0.18}; }
    double t = (e - 0.75) / 0.20; if (t > 1) t = 1;
    return {0.60 + t * 0.35, 0.40 + t * 0.50, 0.32 + t * 0.60};
}

// base colour of Köppen classes 1..30 (satellite-view palette), packed as 3 doubles per class
PB_DEV Rgb biome_base(int id) {
    const double T[31][3] = {
        {0, 0, 0}, {0.05, 0.30, 0.05}, {0.08, 0.33, 0.07}, {0.42, 0.50, 0.18}, {0.82, 0.72, 0.50}, {0.60, 0.55, 0.48}, {0.72, 0.62, 0.30},
        {0.55, 0.52, 0.32}, {0.18, 0.42, 0.12}, {0.12, 0.38, 0.10}, {0.10, 0.28, 0.10}, {0.45, 0.48, 0.22}, {0.40, 0.45, 0.20},
        {0.35, 0.40, 0.20}, {0.20, 0.44, 0.14}, {0.15, 0.40, 0.12}, {0.12, 0.32, 0.10}, {0.12, 0.36, 0.08}, {0.10, 0.32, 0.08},
        {0.06, 0.22, 0.08}, {0.05, 0.18, 0.07}, {0.38, 0.38, 0.18}, {0.35, 0.35, 0.17}, {0.08, 0.22, 0.08}, {0.06, 0.18, 0.07},
        {0.14, 0.36, 0.10}, {0.12, 0.32, 0.09}, {0.07, 0.22, 0.08}, {0.05, 0.18, 0.07}, {0.35, 0.32, 0.22}, {0.78, 0.80, 0.84}};
    if (id < 1 || id > 30) return {0.30, 0.50, 0.20};
    return {T[id][0], T[id][1], T[id][2]};
}

PB_DEV Rgb biome_color(int id, double elevation) {       // biomeColor
    if (id == 0 || elevation <= 0) return terrain_color(elevation);
    Rgb c = biome_base(id);
    const double hKm = elev_to_height_km(elevation);
    double alpine, snow;                                   // altitudeThresholds
    if (id <= 3) { alpine = 3.5; snow = 5.5; }
    else if (id <= 7) { alpine = 3.0; snow = 5.0; }
    else if (id <= 16) { alpine = 2.0; snow = 3.5; }
    else if (id <= 18 || id == 21 || id == 22 || id == 25 || id == 26) { alpine = 1.5; snow = 3.0; }
    else if (id <= 28) { alpine = 0.8; snow = 2.0; }
    else if (id == 29) { alpine = 0.4; snow = 1.5; }
    else { alpine = 0; snow = 0.5; }
    if (hKm < 0.2) { const double k = 0.93 + 0.07 * (hKm / 0.2); c.r *= k; c.g *= k; c.b *= k; }
    if (alpine > 0 && hKm > 0.2 && hKm < alpine) {
        const double t = (hKm - 0.2) / (alpine - 0.2), k = 1.0 - t * 0.15;
        c.r *= k; c.g *= k; c.b *= k;
    }
    if (alpine > 0 && hKm > alpine) {
        const double zone = snow > alpine ? snow - alpine : 2.0;
        double t = (hKm - alpine) / zone; if (t > 1) t = 1;
        const double s = t * t;
        c.r = c.r + (0.42 - c.r) * s; c.g = c.g + (0.38 - c.g) * s; c.b = c.b + (0.32 - c.b) * s;
    }
    if (snow > 0 && hKm > snow) {
        double t = (hKm - snow) / 2.5; if (t > 1) t = 1;
        const double s = t * t;
        c.r = c.r + (0.92 - c.r) * s; c.g = c.g + (0.93 - c.g) * s; c.b = c.b + (0.96 - c.b) * s;
    }
    return c;
}

// koppenColor: KOPPEN_CLASSES[classId].color, ids outside the table take class 0 (js/planet-mesh.js:175-178, js/koppen.js:19-51)
PB_DEV Rgb koppen_class_color(int id) {
    const double T[31][3] = {
        {0.29, 0.44, 0.65}, {0.00, 0.00, 1.00}, {0.00, 0.47, 1.00}, {0.27, 0.67, 0.98}, {1.00, 0.00, 0.00}, {1.00, 0.59, 0.59},
        {0.96, 0.65, 0.00}, {1.00, 0.86, 0.39}, {0.78, 1.00, 0.31}, {0.39, 1.00, 0.31}, {0.20, 0.78, 0.00}, {1.00, 1.00, 0.00},
        {0.78, 0.78, 0.00}, {0.59, 0.59, 0.00}, {0.59, 1.00, 0.59}, {0.39, 0.78, 0.39}, {0.20, 0.59, 0.20}, {0.00, 1.00, 1.00},
        {0.22, 0.78, 1.00}, {0.00, 0.49, 0.49}, {0.00, 0.27, 0.37}, {0.90, 0.50, 1.00}, {0.70, 0.35, 0.85}, {0.50, 0.20, 0.65},
        {0.35, 0.10, 0.45}, {0.67, 0.69, 1.00}, {0.43, 0.47, 0.78}, {0.29, 0.31, 0.78}, {0.20, 0.00, 0.53}, {0.70, 0.70, 0.70},
        {0.41, 0.41, 0.41}};
    if (id < 0 || id > 30) id = 0;
    return {T[id][0], T[id][1], T[id][2]};
}

enum ColorMode { COLOR_TERRAIN = 0, COLOR_BIOME = 1, COLOR_HEIGHTMAP = 2, COLOR_LAND_HEIGHTMAP = 3, COLOR_LAND_MASK = 4, COLOR_BIOME_RAW = 5,
                 COLOR_KOPPEN = 6 };

struct RegionColorK {
    int mode; const float* elev; const uint8_t* koppen; float* rgb;
    PB_DEV void operator()(int r) const {
        const double e = elev[r];
        Rgb c{0, 0, 0};
        if (mode == COLOR_TERRAIN) c = terrain_color(e);
        else if (mode == COLOR_BIOME || mode == COLOR_BIOME_RAW) c = biome_color(koppen[r], e);
        else if (mode == COLOR_KOPPEN) c = koppen_class_color(koppen[r]);
        else if (mode == COLOR_HEIGHTMAP) { double t = (elev_to_height_km(e) + 5) / 11; t = t > 1 ? 1 : t; t = t < 0 ? 0 : t; c = {t, t, t}; }
        else if (mode == COLOR_LAND_HEIGHTMAP) { if (e > 0) { double t = elev_to_height_km(e) / 6; t = t > 1 ? 1 : t; t = t < 0 ? 0 : t; c = {t, t, t}; } }
        else if (e > 0) c = {1, 1, 1};
        rgb[3 * r] = (float)c.r; rgb[3 * r + 1] = (float)c.g; rgb[3 * r + 2] = (float)c.b;
    }
};
// smoothBiomeColors: 65 % own colour + 35 % mean of the neighbours' raw colours (js/planet-mesh.js:30-62)
struct BiomeBlendK {
    Csr g; const float* raw; float* out;
    PB_DEV void operator()(int r) const {
        const int s = g.off[r], e = g.off[r + 1], count = e - s;
        if (count == 0) { for (int k = 0; k < 3; k++) out[3 * r + k] = raw[3 * r + k]; return; }
        double a0 = 0, a1 = 0, a2 = 0;
        for (int i = s; i < e; i++) { const int nr = g.adj[i]; a0 += raw[3 * nr]; a1 += raw[3 * nr + 1]; a2 += raw[3 * nr + 2]; }
        a0 /= count; a1 /= count; a2 /= count;
        const double alpha = 0.35;
        out[3 * r] = (float)((double)raw[3 * r] * (1 - alpha) + a0 * alpha);
        out[3 * r + 1] = (float)((double)raw[3 * r + 1] * (1 - alpha) + a1 * alpha);
        out[3 * r + 2] = (float)((double)raw[3 * r + 2] * (1 - alpha) + a2 * alpha);
    }
};

}  // namespace pb
