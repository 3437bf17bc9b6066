// pb_elevation.h — per-cell kernels of assignElevation (js/elevation.js:216-1391), kernel family K5:
// collision detection (:27-122), dual-layer blends (:250-327), the main harmonic-mean / noise
// synthesis loop (:638-973), coastal roughening (:978-1050), island arcs (:1088-1106), hotspot dome
// uplift (:1264-1372) and peak compression (:1378-1382).
// Float32Array read-modify-writes of the reference (`r_elevation[r] += x`) round to f32 at every
// step; the kernels keep the running elevation in a float and do the same.
#pragma once
#include "pb_platform.h"
#include "pb_stencil.h"
#include "pb_noise.h"
#include "pb_climate.h"   // jmin / jmax

namespace pb {

// device view of a JS object keyed by plate id ({pid: {pole, omega}}, {pid: density}, Set of oceanic ids)
struct PlateTab {
    const int* index;      // [tableSize] id → row or -1
    int tableSize;
    const uint8_t* isOcean; const double* pole; const double* omega; const double* density;
    PB_DEV int find(int id) const { return (id >= 0 && id < tableSize) ? index[id] : -1; }
    PB_DEV bool ocean(int id) const { const int k = find(id); return k >= 0 && isOcean[k]; }
};

PB_DEV double pair_intensity(int a, int b) {   // :44-53, JS ToInt32 / ToUint32 semantics on doubles
    const double lo = a < b ? a : b, hi = a < b ? b : a;
    const int32_t x1 = (int32_t)(uint32_t)(unsigned long long)(long long)(lo * 16807.0);
    const int32_t x2 = (int32_t)(uint32_t)(unsigned long long)(long long)(hi * 48271.0);
    const int32_t h1 = x1 ^ x2;                                   // (>>> 0 then ToInt32 again: same bits)
    const int32_t x = (h1 >> 16) ^ h1;
    const double p = (double)x * 73244475.0;                      // may exceed 2^53: rounds in double first
    const uint32_t h = (uint32_t)(unsigned long long)(long long)p;
    return 0.5 + (double)(h % 10001u) / 10000;
}

struct CollisionOut { float* stress; float* subduct; int8_t* btype; uint8_t* bothOcean; uint8_t* hasOcean; uint8_t* setCode; };

// findCollisions :27-122.  setCode: 0 none, 1 mountain_r, 2 coastline_r, 3 ocean_r
struct CollisionsK {
    Csr g; const float* xyz; PlateTab P; const int* r_plate; Simplex noise; double dt; int undulOctaves; CollisionOut o;
    PB_DEV void vel(int k, double x, double y, double z, double* v) const {
        if (k < 0) { v[0] = v[1] = v[2] = NAN; return; }      // unknown plate id (the reference would throw)
        const double px = P.pole[3 * k], py = P.pole[3 * k + 1], pz = P.pole[3 * k + 2], om = P.omega[k];
        v[0] = om * (py * z - pz * y); v[1] = om * (pz * x - px * z); v[2] = om * (px * y - py * x);
    }
    PB_DEV void operator()(int r) const {
        const int myPlate = r_plate[r];
        const int kMy = P.find(myPlate);
        const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
        double bestComp = -INFINITY, bestNormalComp = 0;
        int best = -1;
        for (int j = g.off[r], e = g.off[r + 1]; j < e; j++) {
            const int nb = g.adj[j];
            const int nbPlate = r_plate[nb];
            if (myPlate == nbPlate) continue;
            const double nx = xyz[3 * nb], ny = xyz[3 * nb + 1], nz = xyz[3 * nb + 2];
            const double dx = x - nx, dy = y - ny, dz = z - nz;
            const double dBefore = sqrt(dx * dx + dy * dy + dz * dz);
            double v1[3], v2[3];
            vel(kMy, x, y, z, v1);
            vel(P.find(nbPlate), nx, ny, nz, v2);
            const double ax = x + v1[0] * dt, ay = y + v1[1] * dt, az = z + v1[2] * dt;
            const double bx = nx + v2[0] * dt, by = ny + v2[1] * dt, bz = nz + v2[2] * dt;
            const double adx = ax - bx, ady = ay - by, adz = az - bz;
            const double dAfter = sqrt(adx * adx + ady * ady + adz * adz);
            const double comp = dBefore - dAfter;
            if (comp > bestComp) {
                bestComp = comp; best = nb;
                const double rvx = v1[0] - v2[0], rvy = v1[1] - v2[1], rvz = v1[2] - v2[2];
                bestNormalComp = -(rvx * dx + rvy * dy + rvz * dz) / or_default(dBefore, 1.0);
            }
        }
        float stress = 0.0f, sub = 0.5f;
        int8_t bt = 0; uint8_t both = 0, has = 0, code = 0;
        if (best != -1) {
            const int bestPlate = r_plate[best];
            const bool collided = bestComp > 0.75 * dt;
            const bool rOcean = P.ocean(myPlate), nOcean = P.ocean(bestPlate);
            both = (rOcean && nOcean) ? 1 : 0;
            has = (rOcean || nOcean) ? 1 : 0;
            const double thresh = 0.3 * dt;
            bt = bestNormalComp > thresh ? 1 : (bestNormalComp < -thresh ? 2 : 3);
            if (collided) stress = (float)((bestComp / dt) * pair_intensity(myPlate, bestPlate));
            const int kB = P.find(bestPlate);
            const double densityDiff = (kMy >= 0 ? P.density[kMy] : NAN) - (kB >= 0 ? P.density[kB] : NAN);
            const double baseFactor = 0.5 + 0.5 * pb_tanh(densityDiff * 8);
            const double undulationStrength = pb_exp(-fabs(densityDiff) * 12);
            const double undulation = noise.fbm(x * 6, y * 6, z * 6, undulOctaves) * 0.4 * undulationStrength;
            sub = (float)jmax(0, jmin(1, baseFactor + undulation));
            if (rOcean && nOcean) code = collided ? 2 : 3;
            else if (!rOcean && !nOcean) { if (collided) code = ((double)sub < 0.55) ? 1 : 2; }
            else code = collided ? 1 : 2;
        }
        o.stress[r] = stress; o.subduct[r] = sub; o.btype[r] = bt; o.bothOcean[r] = both; o.hasOcean[r] = has; o.setCode[r] = code;
    }
};

// max of a non-negative f32 field (bit pattern order == value order)
struct MaxF32K { const float* v; int* out; PB_DEV void operator()(int r) const { const float f = v[r]; if (f > 0) {
#if PB_CUDA
    atomic_max(out, __float_as_int(f));
#else
    int b; memcpy(&b, &f, 4); atomic_max(out, b);
#endif
} } };

// ---- main loop inputs ------------------------------------------------------------------------------------
struct ElevFields {
    const float* stress; const float* subduct; const int8_t* btype; const uint8_t* isOcean;
    const float *dist_mountain, *dist_ocean, *dist_coastline, *dist_coast, *dist_coast_land;
    const float *riftDist, *ridgeDist, *fractureDist, *backArcDist, *backArcStress;
    const uint8_t* coastConvergent;
};
struct ElevParams {
    double maxStress, noiseMag, scaleFactor, interiorBand, tectonicReach, plateauStart, riftHalfWidth, ridgeHalfWidth,
        fractureHalfWidth, baStart, baPeak, baEnd;
    int warpOctaves;
};
struct ElevDebug { float *base, *tectonic, *noise, *interior, *coastal, *ocean, *hotspot, *tecActivity, *margins, *backArc, *foldRidge, *orogenicPower; };

PB_DEV double js_round_d(double x) { return floor(x + 0.5); }
PB_DEV float addf(float e, double v) { return (float)((double)e + v); }

PB_DEV double back_arc_effect(double bad, double dMtn, double stressB, const ElevParams& p) {
    const double orogenyFactor = (dMtn != INFINITY && dMtn < bad) ? jmax(0, dMtn / bad) : 1.0;
    double baEffect = 0;
    if (bad <= p.baPeak) { const double t = (bad - p.baStart) / jmax(1, p.baPeak - p.baStart); const double s = t * t * (3 - 2 * t); baEffect = -0.10 * stressB * s * orogenyFactor; }
    else if (bad <= p.baEnd) { const double t = (bad - p.baPeak) / jmax(1, p.baEnd - p.baPeak); const double s = t * t * (3 - 2 * t); baEffect = -0.10 * stressB * (1 - s) * orogenyFactor; }
    return baEffect;
}

// main per-cell loop :638-973
struct ElevationMainK {
    const float* xyz; const int* r_plate; PlateTab P; ElevFields F; ElevParams p; Simplex noise, riftNoise, foldNoise; float* elev; ElevDebug d;
    PB_DEV void operator()(int r) const {
        const bool isOceanPlate = F.isOcean[r] != 0;
        const double eps = 1e-3, warpScale = 0.4;
        const double sfAsym = F.subduct[r];
        const double asymmetry = 1.0 + (sfAsym - 0.5) * 0.8;
        const double a = (double)F.dist_mountain[r] * asymmetry + eps;
        const double b = (double)F.dist_ocean[r] + eps;
        const double c = (double)F.dist_coastline[r] + eps;
        float e;
        if (a == INFINITY && b == INFINITY) e = (float)(0.1 * 0.6);
        else e = (float)((1 / a - 1 / b) / (1 / a + 1 / b + 1 / c) * 0.6);
        d.base[r] = e;
        const double stressNorm = jmin(1, (double)F.stress[r] / p.maxStress);
        const int btype = F.btype[r];
        const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
        const double wx = x + warpScale * noise.fbm(x + 5.3, y + 1.7, z + 3.1, p.warpOctaves);
        const double wy = y + warpScale * noise.fbm(x + 8.1, y + 2.9, z + 7.3, p.warpOctaves);
        const double wz = z + warpScale * noise.fbm(x + 1.4, y + 6.2, z + 4.8, p.warpOctaves);
        const double rawOro = noise.noise3D(x * 1.5 + 33.7, y * 1.5 + 11.2, z * 1.5 + 22.9);
        const double shaped = rawOro >= 0 ? sqrt(rawOro) : -sqrt(-rawOro);
        const double orogenicPower = jmax(0, jmin(1, 0.5 + 0.5 * shaped));
        d.orogenicPower[r] = (float)(orogenicPower - 0.5);
        float dlTect = 0, dlNoise = 0, dlInterior = 0, dlBackArc = 0, dlFold = 0, dlTecAct = 0, dlOcean = 0, dlMargins = 0;

        if (!isOceanPlate) {
            const double sf = F.subduct[r];
            const float elevBefore = e;
            if (sf > 0.5 && e > 0) { const double suppression = (sf - 0.5) * 2; e = (float)((double)e * (1 - suppression * 0.42)); }
            if (stressNorm > 0.01) {
                const double stressMag = stressNorm * stressNorm * 0.55 * orogenicPower;
                const double uplift = stressMag * (1 - sf), depress = stressMag * 0.4 * sf;
                const double heightVar = 0.60 + 0.8 * noise.fbm(x * 8 + 13.7, y * 8 + 9.2, z * 8 + 4.5, 3);
                e = addf(e, (uplift - depress) * heightVar);
            }
            if (stressNorm > 0 && stressNorm < 0.10) { const double forelandT = stressNorm / 0.10; e = addf(e, -(0.06 * (1 - forelandT))); }
            {
                const double rd = F.riftDist[r];
                if (rd != INFINITY) {
                    const double floorEnd = jmax(1, js_round_d(1.5 * p.scaleFactor)), shoulderEnd = jmax(2, js_round_d(2.5 * p.scaleFactor));
                    double riftEffect = 0;
                    if (rd <= 0.5) { riftEffect = -0.15; riftEffect += riftNoise.ridgedFbm(x * 8, y * 8, z * 8, 3, 2.0, 0.5, 1.0) * 0.04; }
                    else if (rd <= floorEnd) { const double t = rd / floorEnd; riftEffect = -0.12 * (1 - t * 0.3); riftEffect += riftNoise.ridgedFbm(x * 8, y * 8, z * 8, 3, 2.0, 0.5, 1.0) * 0.03 * (1 - t); }
                    else if (rd <= shoulderEnd) { const double t = (rd - floorEnd) / (shoulderEnd - floorEnd); riftEffect = 0.03 * (1 - t); }
                    else if (p.riftHalfWidth > shoulderEnd) {
                        const double t = (rd - shoulderEnd) / (p.riftHalfWidth - shoulderEnd);
                        const double fadeT = jmin(1, t);
                        const double fade = fadeT * fadeT * (3 - 2 * fadeT);
                        riftEffect = 0.03 * (1 - fade) * 0.2;
                    }
                    e = addf(e, riftEffect);
                }
            }
            {
                const double bad = F.backArcDist[r];
                if (bad != INFINITY && bad >= p.baStart) {
                    const double baEffect = back_arc_effect(bad, F.dist_mountain[r], F.backArcStress[r], p);
                    e = addf(e, baEffect);
                    dlBackArc = (float)baEffect;
                }
            }
            dlTect = (float)((double)e - (double)elevBefore);
            const double dMtn = F.dist_mountain[r];
            const double rawProximity = (dMtn == INFINITY || dMtn >= p.tectonicReach) ? 0 : (1 - dMtn / p.tectonicReach);
            const double tectonicActivity = jmax(stressNorm, rawProximity * rawProximity);
            dlTecAct = (float)tectonicActivity;
            {
                const int k = P.find(r_plate[r]);
                const double foldActivity = tectonicActivity * tectonicActivity;
                if (k >= 0 && foldActivity > 0.01) {
                    const double uu = x * P.pole[3 * k] + y * P.pole[3 * k + 1] + z * P.pole[3 * k + 2];
                    const double phaseWarp = foldNoise.fbm(x * 3 + 55.3, y * 3 + 33.7, z * 3 + 17.2, 2) * 0.08;
                    const double phase = (uu + phaseWarp) * 30 * PB_PI;
                    const double ridge = 1 - fabs(pb_sin(phase));
                    const double foldCentered = ridge - 0.36;
                    const double ampMod = 0.6 + 0.4 * foldNoise.fbm(x * 4 + 88.1, y * 4 + 62.3, z * 4 + 41.7, 2);
                    const double elevBoost = 1 + 4 * jmax(0, (double)e);
                    const double foldAmp = foldActivity * jmax(0, 1 - sf * 1.5) * p.noiseMag * 0.8 * elevBoost;
                    const double foldContrib = foldCentered * foldAmp * ampMod;
                    e = addf(e, foldContrib);
                    dlFold = (float)foldContrib;
                }
            }
            const bool isPlateauZone = sf < 0.45 && dMtn != INFINITY && dMtn > p.plateauStart;
            const double blend = jmin(1, stressNorm * 3);
            const double smoothNoise = noise.fbm(wx, wy, wz, 5) * p.noiseMag;
            const double ridgedNoise = noise.ridgedFbm(wx, wy, wz) * p.noiseMag * 1.5;
            const double noiseVal = smoothNoise * (1 - blend) + ridgedNoise * blend;
            const double detailNoise = noise.fbm(wx * 4 + 22.1, wy * 4 + 6.8, wz * 4 + 15.4, 4, 0.5) * p.noiseMag * 0.5;
            const double noiseActivity = jmin(1, stressNorm * 4);
            const double plateauSuppress = isPlateauZone ? jmax(0.30, 1 - tectonicActivity * 0.60) : 1.0;
            const double noiseScale = (0.25 + 0.75 * noiseActivity) * plateauSuppress;
            const double fineNoise = noise.fbm(wx * 8 + 41.7, wy * 8 + 13.2, wz * 8 + 27.9, 3, 0.5) * p.noiseMag * 0.25;
            const double fineScale = sqrt(noiseScale);
            const double totalNoise = (noiseVal + detailNoise) * noiseScale + fineNoise * fineScale;
            e = addf(e, totalNoise);
            dlNoise = (float)totalNoise;
            {
                const double currentElev = e;
                if (currentElev > 0.12) {
                    const double elevExcess = currentElev - 0.12;
                    const double dissectVal = noise.fbm(wx * 16 + 71.3, wy * 16 + 44.8, wz * 16 + 29.1, 3, 0.5);
                    const double dissectContrib = dissectVal * (sqrt(elevExcess) * stressNorm * p.noiseMag * 0.4);
                    e = addf(e, dissectContrib);
                    dlNoise = addf(dlNoise, dissectContrib);
                }
            }
            {
                const double currentElev = e;
                if (currentElev > 0.65 && stressNorm > 0.2) {
                    const double excess = currentElev - 0.65;
                    const double peakNoise = noise.ridgedFbm(wx * 24 + 91.3, wy * 24 + 55.7, wz * 24 + 38.2, 3, 0.5, 0.5, 1.0);
                    const double spike = jmax(0, peakNoise - 0.45);
                    const double peakContrib = spike * excess * stressNorm * 1.2;
                    e = addf(e, peakContrib);
                    dlNoise = addf(dlNoise, peakContrib);
                }
            }
            const double lcd = F.dist_coast_land[r];
            if (lcd < INFINITY) {
                const double tDown = jmin(lcd / p.interiorBand, 1);
                const double sDown = tDown * tDown * (3 - 2 * tDown);
                const double tUp = jmin(lcd / (p.interiorBand * 0.4), 1);
                const double sUp = tUp * tUp * (3 - 2 * tUp);
                const double interiorUplift = 0.06 + tectonicActivity * 0.16;
                const double baseBias = -0.08 * (1 - sDown) + interiorUplift * sUp;
                const double mod = 1.0 + 0.2 * noise.fbm(x * 2 + 19.3, y * 2 + 7.6, z * 2 + 13.1, 2);
                const double bias = baseBias * mod;
                e = addf(e, bias);
                dlInterior = (float)bias;
            }
            if (isPlateauZone && tectonicActivity > 0.1) {
                const double plateauBoost = 0.025 * tectonicActivity * (1 - sf);
                e = addf(e, plateauBoost);
                dlInterior = addf(dlInterior, plateauBoost);
            }
        } else {
            const double dc = F.dist_coast[r];
            double oceanBase;
            if (dc < 5) oceanBase = -0.04 - 0.06 * (dc / 5);
            else if (dc < 12) oceanBase = -0.10 - 0.25 * ((dc - 5) / 7);
            else oceanBase = -0.35 + noise.fbm(x * 2, y * 2, z * 2, 3) * 0.03;
            e = (float)jmin((double)e, oceanBase);
            dlOcean = e;
            dlMargins = F.coastConvergent[r] == 1 ? 0.8f : 0.2f;
            const double rd = F.ridgeDist[r], fd = F.fractureDist[r];
            if (rd != INFINITY && rd <= p.ridgeHalfWidth) dlMargins = 1.0f;
            if (fd != INFINITY && fd <= p.fractureHalfWidth) dlMargins = -0.5f;
            const float elevBeforeOcTec = e;
            if (rd != INFINITY && rd <= p.ridgeHalfWidth) {
                const double t = rd / p.ridgeHalfWidth;
                const double ridgeFade = (1 - t) * (1 - t);
                const double ridgeNoise = noise.ridgedFbm(x * 3, y * 3, z * 3, 4, 2.0, 0.5, 1.0);
                e = addf(e, (0.12 * ridgeNoise + 0.06) * ridgeFade);
            }
            if (fd != INFINITY && fd <= p.fractureHalfWidth) { const double ft = fd / p.fractureHalfWidth; e = addf(e, -(0.03 * (1 - ft))); }
            if (btype == 1) e = addf(e, -(0.15 + 0.15 * stressNorm));
            {
                const double bad = F.backArcDist[r];
                if (bad != INFINITY && bad >= p.baStart) {
                    const double baEffect = back_arc_effect(bad, F.dist_mountain[r], F.backArcStress[r], p);
                    e = addf(e, baEffect);
                    dlBackArc = (float)baEffect;
                }
            }
            dlTect = (float)((double)e - (double)elevBeforeOcTec);
            const double oceanNoise = noise.fbm(wx, wy, wz, 5) * p.noiseMag * 0.3;
            e = addf(e, oceanNoise);
            dlNoise = (float)oceanNoise;
        }
        elev[r] = e;
        d.tectonic[r] = dlTect; d.noise[r] = dlNoise; d.interior[r] = dlInterior; d.backArc[r] = dlBackArc; d.foldRidge[r] = dlFold;
        d.tecActivity[r] = dlTecAct; d.ocean[r] = dlOcean; d.margins[r] = dlMargins;
    }
};

// coastal roughening :984-1049
struct CoastalRoughenK {
    const float* xyz; const float* dBdry; const float* coastStressMax; const float* coastSubductMax; const uint8_t* coastConvergent;
    const float* stress; const uint8_t* isOcean; double maxStress, noiseMag, coastRoughenDist, islandReach;
    Simplex noise, cNoise, cNoise2, cNoise3; float* elev; float* dlCoastal;
    PB_DEV void operator()(int r) const {
        const double db = dBdry[r];
        if (db > coastRoughenDist) { dlCoastal[r] = 0; return; }
        const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
        const double t = db / coastRoughenDist;
        const double sn = jmin(1, jmax((double)coastStressMax[r], (double)stress[r] / maxStress));
        const bool conv = coastConvergent[r] != 0;
        const bool isSubductingOcean = isOcean[r] && conv && (double)coastSubductMax[r] > 0.45;
        const double subSup = isSubductingOcean ? jmin(1, ((double)coastSubductMax[r] - 0.45) / 0.55) : 0;
        float e = elev[r];
        const float elevBeforeCoast = e;
        const bool isPassiveCoast = !conv;
        const double falloff1 = (1 - t) * (1 - t);
        const double stressAmp1 = 1 + sn * 5;
        const double coastFreq = isPassiveCoast ? 12 : 18;
        const double coastAmp = isPassiveCoast ? 0.08 : 0.12;
        const double n1 = cNoise.fbm(x * coastFreq + 3.7, y * coastFreq + 7.1, z * coastFreq + 2.3, 5, 0.55);
        double coastNoise1 = n1 * coastAmp * falloff1 * stressAmp1;
        if (subSup > 0 && coastNoise1 > 0) coastNoise1 *= (1 - subSup);
        e = addf(e, coastNoise1);
        const double warpReach = isPassiveCoast ? 1.2 : 1.5;
        const double falloffW = jmax(0, 1 - t * warpReach);
        if (falloffW > 0) {
            const double warpAmt = 0.35 * falloffW * (1 + sn * 2);
            const double dwx = cNoise3.fbm(x * 6 + 11.3, y * 6 + 4.7, z * 6 + 8.2, 3, 0.6) * warpAmt;
            const double dwy = cNoise3.fbm(x * 6 + 2.9, y * 6 + 9.4, z * 6 + 1.6, 3, 0.6) * warpAmt;
            const double dwz = cNoise3.fbm(x * 6 + 7.5, y * 6 + 0.3, z * 6 + 5.9, 3, 0.6) * warpAmt;
            const double origN = noise.fbm(x, y, z, 5) * noiseMag;
            const double warpN = noise.fbm(x + dwx, y + dwy, z + dwz, 5) * noiseMag;
            double warpDelta = (warpN - origN) * falloffW;
            if (subSup > 0 && warpDelta > 0) warpDelta *= (1 - subSup);
            e = addf(e, warpDelta);
        }
        if (isOcean[r] && db > 0 && db <= islandReach && subSup < 0.3) {
            const double islandN = cNoise2.fbm(x * 35 + 5.1, y * 35 + 9.3, z * 35 + 2.7, 4, 0.5);
            const double threshold = 0.25 - sn * 0.2;
            if (islandN > threshold) {
                const double excess = (islandN - threshold) / (1 - threshold);
                const double distFade = 1 - (db / islandReach);
                double bump = excess * excess * 0.18 * (1 + sn * 2) * distFade;
                bump *= (1 - subSup / 0.3);
                e = addf(e, bump);
            }
        }
        elev[r] = e;
        dlCoastal[r] = (float)(0.0 + ((double)e - (double)elevBeforeCoast));     // dl_coastal starts at 0 (+=)
    }
};

// island-arc uplift :1088-1106
struct IslandArcK {
    const float* xyz; const float* arcDist; const float* arcStress; double maxArcDist, scaleFactor; Simplex arcNoise; float* elev; float* dlCoastal;
    PB_DEV void operator()(int r) const {
        const double dd = arcDist[r];
        if (dd < 1 || dd > maxArcDist) return;
        const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
        const double peakDist = jmax(1.5, 1.5 * scaleFactor), sigma = jmax(1.5, 1.5 * scaleFactor);
        const double q = (dd - peakDist) / sigma;
        const double distWeight = pb_exp(-0.5 * (q * q));
        const double n = arcNoise.ridgedFbm(x * 4, y * 4, z * 4, 4, 2.0, 0.5, 1.0);
        if (n > 0.30) {
            const double excess = (n - 0.30) / (1 - 0.30);
            const double uplift = excess * excess * 0.55 * distWeight * (0.5 + (double)arcStress[r]);
            elev[r] = addf(elev[r], uplift);
            dlCoastal[r] = addf(dlCoastal[r], uplift);
        }
    }
};

// hotspot domes :1239-1372.  The dome list (RNG-placed: class R) is built on the host: 5 hotspots, each an active dome plus
// a trail of max(3, 6 + round((u - 0.5)·10)) ≤ 11 domes (:1116-1117, 1153), i.e. at most 60 entries.
#define PB_MAX_DOMES 64
struct DomeDev {
    double x, y, z, strength, ux, uy, uz, vx, vy, vz, cosThreshPeak, invS2, swellStrength, cosThreshSwell, invS2Swell,
        driftStretch, calderaDepth, invS2Caldera, ageFactor, riftAngles[3];
    int nRift, hasCaldera;
};
struct HotspotK {
    const float* xyz; const DomeDev* domes; int nDomes; Simplex hsNoise, hsNoise2; float* elev; float* dlHotspot;
    PB_DEV void operator()(int r) const {
        const double rx = xyz[3 * r], ry = xyz[3 * r + 1], rz = xyz[3 * r + 2];
        bool nearSwell = false, nearPeak = false;
        for (int k = 0; k < nDomes; k++) {
            const DomeDev& dm = domes[k];
            const double cdot = dm.x * rx + dm.y * ry + dm.z * rz;
            if (cdot > dm.cosThreshSwell) { nearSwell = true; if (cdot > dm.cosThreshPeak) { nearPeak = true; break; } }
        }
        if (!nearSwell) { dlHotspot[r] = 0; return; }
        double shapeWarpSq = 1.0;
        if (nearPeak) {
            const double ws = 8;
            const double wx = hsNoise2.fbm(rx * ws + 5.1, ry * ws + 3.7, rz * ws + 9.2, 2, 0.5) * 0.4;
            const double wy = hsNoise2.fbm(rx * ws + 11.3, ry * ws + 7.1, rz * ws + 2.9, 2, 0.5) * 0.4;
            const double wz = hsNoise2.fbm(rx * ws + 1.7, ry * ws + 13.5, rz * ws + 6.4, 2, 0.5) * 0.4;
            const double shapeWarp = 1.0 + 0.40 * hsNoise.fbm((rx + wx) * 20 + 3.2, (ry + wy) * 20 + 7.8, (rz + wz) * 20 + 1.5, 4, 0.5);
            shapeWarpSq = shapeWarp * shapeWarp;
        }
        double totalUplift = 0, totalSwellUplift = 0, weightedAge = 0, ageWeightSum = 0;
        for (int k = 0; k < nDomes; k++) {
            const DomeDev& dm = domes[k];
            const double dot = dm.x * rx + dm.y * ry + dm.z * rz;
            if (dot > dm.cosThreshSwell) { const double swAngleSq = 2 * (1 - dot); totalSwellUplift += dm.swellStrength * pb_exp(swAngleSq * dm.invS2Swell); }
            if (dot < dm.cosThreshPeak) continue;
            const double offX = rx - dot * dm.x, offY = ry - dot * dm.y, offZ = rz - dot * dm.z;
            const double parComp = offX * dm.ux + offY * dm.uy + offZ * dm.uz;
            const double perpComp = offX * dm.vx + offY * dm.vy + offZ * dm.vz;
            const double stretchedParSq = (parComp * dm.driftStretch) * (parComp * dm.driftStretch);
            const double angleSq = stretchedParSq + perpComp * perpComp;
            double gauss = pb_exp(angleSq * shapeWarpSq * dm.invS2);
            if (dm.nRift > 0 && gauss > 0.01) {
                const double angle = pb_atan2(perpComp, parComp);
                double maxRift = 0;
                for (int ri = 0; ri < dm.nRift; ri++) {
                    double da = angle - dm.riftAngles[ri];
                    da = da - js_round_d(da / (2 * PB_PI)) * 2 * PB_PI;
                    const double c2 = pb_cos(da);
                    const double riftFactor = c2 * c2 * c2 * c2;
                    if (riftFactor > maxRift) maxRift = riftFactor;
                }
                gauss *= (1.0 + 0.5 * maxRift);
            }
            const double peakUplift = dm.strength * gauss;
            totalUplift += peakUplift;
            weightedAge += dm.ageFactor * peakUplift;
            ageWeightSum += peakUplift;
            if (dm.hasCaldera) totalUplift -= dm.calderaDepth * pb_exp(angleSq * dm.invS2Caldera);
        }
        float out = 0;
        const double combinedUplift = totalSwellUplift + totalUplift;
        if (combinedUplift > 0.001) {
            const double age = ageWeightSum > 0 ? weightedAge / ageWeightSum : 0;
            const double texBase = 0.7 * hsNoise.ridgedFbm(rx * 12, ry * 12, rz * 12, 4, 2.0, 0.5, 1.0);
            const double texDetail = 0.3 * hsNoise.ridgedFbm(rx * 30, ry * 30, rz * 30, 3, 2.0, 0.5, 1.0);
            const double texRaw = texBase + texDetail;
            const double texMin = 0.4 + age * 0.3, texMax = 1.2 - age * 0.2;
            const double volc = texMin + (texMax - texMin) * texRaw;
            const double uplift = totalSwellUplift + jmax(0, totalUplift) * volc;
            elev[r] = addf(elev[r], uplift);
            out = (float)uplift;
        }
        dlHotspot[r] = out;
    }
};

struct CompressPeaksK { float* elev; PB_DEV void operator()(int r) const { const float e = elev[r]; if (e > 0) elev[r] = (float)pb_pow((double)e, 0.92); } };

}  // namespace pb
