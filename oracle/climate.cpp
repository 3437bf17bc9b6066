// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).  The reference has no tests or golden
// vectors and no JS runtime exists in this image; this file restates its source line by line and is pinned against that source
// executed under tests/golden/minijs.py (tests/test_zz_reference_vectors.py: every climate array of every reply, bit for bit).
//
// CPU restatement of the reference's climate stack, single thread, double arithmetic with f32
// typed-array stores, same loop order:
//   js/wind.js            computeWind            :394-687   (+ spline :11-72, geo index :88-165,
//                                                              ITCZ :174-232, pressure :239-301,
//                                                              gradients :306-339, wind :343-378)
//   js/ocean.js           computeOceanCurrents   :204-382
//   js/heuristic-precip.js                       :16-269
//   js/precipitation.js   computePrecipitation   :196-684
//   js/temperature.js     computeTemperature     :69-237
//   js/koppen.js          classifyKoppen         :67-288
//   js/climate-util.js    smoothField/makeItczLookup/percentile
//   js/color-map.js       elevToHeightKm         :7-12
// Transcendentals go through include/pb_detmath.h (see its header for why).
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <unordered_set>
#include <vector>
#include "climate.h"
#include "noise.h"

namespace {

typedef std::vector<float> F32;
typedef std::vector<int32_t> I32;
typedef std::vector<uint8_t> U8;
typedef std::vector<double> F64;

const double DEG = PB_PI / 180;
const double RAD = 180 / PB_PI;

// js/color-map.js:7-12
double elevToHeightKm(double elev) {
    if (elev <= 0) return elev * 10;
    const double t = js::min(elev, 1);
    const double t2 = t * t;
    return 6 * t2 * t2 * (5 - 4 * t);
}

// js/wind.js:76-80
double smoothstep(double edge0, double edge1, double x) {
    if (edge0 == edge1) return x >= edge1 ? 1 : 0;
    const double t = js::max(0, js::min(1, (x - edge0) / (edge1 - edge0)));
    return t * t * (3 - 2 * t);
}

// js/wind.js:11-55
struct Spline {
    F64 xs, ys, b, c, d, h;
    int n = 0;
    double period = 0;
};
Spline buildPeriodicSpline(const F64& xs, const F64& ys) {
    Spline s;
    const int n = (int)xs.size();
    const double period = 2 * PB_PI;
    F64 h(n), alpha(n);
    for (int i = 0; i < n; i++) {
        const int next = (i + 1) % n;
        h[i] = std::fmod(xs[next] - xs[i] + period, period);
        if (h[i] == 0) h[i] = period / n;
    }
    for (int i = 0; i < n; i++) {
        const int prev = (i - 1 + n) % n, next = (i + 1) % n;
        alpha[i] = (3 / h[i]) * (ys[next] - ys[i]) - (3 / h[prev]) * (ys[i] - ys[prev]);
    }
    F64 c(n, 0.0);
    for (int iter = 0; iter < 20; iter++)
        for (int i = 0; i < n; i++) {
            const int prev = (i - 1 + n) % n, next = (i + 1) % n;
            c[i] = (alpha[i] - h[prev] * c[prev] - h[i] * c[next]) / (2 * (h[prev] + h[i]));
        }
    F64 b(n), d(n);
    for (int i = 0; i < n; i++) {
        const int next = (i + 1) % n;
        b[i] = (ys[next] - ys[i]) / h[i] - h[i] * (c[next] + 2 * c[i]) / 3;
        d[i] = (c[next] - c[i]) / (3 * h[i]);
    }
    s.xs = xs; s.ys = ys; s.b = b; s.c = c; s.d = d; s.h = h; s.n = n; s.period = period;
    return s;
}
// js/wind.js:57-72
double evaluateSpline(const Spline& sp, double lon) {
    const int n = sp.n;
    const double period = sp.period;
    const double t = std::fmod(std::fmod(lon - sp.xs[0], period) + period, period) + sp.xs[0];
    int seg = 0;
    for (int i = 0; i < n; i++) {
        const int next = (i + 1) % n;
        const double lo = sp.xs[i];
        const double hi = i < n - 1 ? sp.xs[next] : sp.xs[0] + period;
        if (t >= lo && t < hi) { seg = i; break; }
    }
    const double dx = t - sp.xs[seg];
    return sp.ys[seg] + sp.b[seg] * dx + sp.c[seg] * dx * dx + sp.d[seg] * dx * dx * dx;
}

// js/climate-util.js:29-43
struct ItczLookup {
    const float* lats;
    int n;
    double step, lonStart;
    ItczLookup(const F32& lons, const F32& l) : lats(l.data()), n((int)lons.size()) {
        step = (2 * PB_PI) / n;
        lonStart = -PB_PI + step * 0.5;
    }
    double operator()(double lon) const {
        double fi = (lon - lonStart) / step;
        fi = std::fmod(std::fmod(fi, (double)n) + n, (double)n);
        const double i0 = std::floor(fi);
        const int i1 = ((int)i0 + 1) % n;
        const double frac = fi - i0;
        return lats[(int)i0] * (1 - frac) + lats[i1] * frac;
    }
};

// js/wind.js:88-165
struct GeoIndex {
    static const int LAT_BINS = 36, LON_BINS = 72;
    std::vector<uint32_t> binOffset, indices;
    const F32 &lat, &lon, &sinLat, &cosLat;
    const float* elev;
    const U8& isLand;
    GeoIndex(const F32& la, const F32& lo, const F32& sl, const F32& cl, const float* el, const U8& il, int N)
        : lat(la), lon(lo), sinLat(sl), cosLat(cl), elev(el), isLand(il) {
        const int numBins = LAT_BINS * LON_BINS;
        std::vector<uint32_t> binCount(numBins, 0), fillPos(numBins, 0);
        auto binOf = [&](int r) {
            const int latBin = (int)js::max(0, js::min(LAT_BINS - 1, std::floor((lat[r] + PB_PI / 2) / PB_PI * LAT_BINS)));
            const int lonBin = (int)js::max(0, js::min(LON_BINS - 1, std::floor((lon[r] + PB_PI) / (2 * PB_PI) * LON_BINS)));
            return latBin * LON_BINS + lonBin;
        };
        for (int r = 0; r < N; r++) binCount[binOf(r)]++;
        binOffset.assign(numBins + 1, 0);
        for (int i = 0; i < numBins; i++) binOffset[i + 1] = binOffset[i] + binCount[i];
        indices.assign(N, 0);
        for (int r = 0; r < N; r++) {
            const int bin = binOf(r);
            indices[binOffset[bin] + fillPos[bin]] = r;
            fillPos[bin]++;
        }
    }
    void sample(double la, double lo, double radius, double* landFrac, double* avgElev) const {
        const double latMin = la - radius, latMax = la + radius;
        const int bMin = (int)js::max(0, std::floor((latMin + PB_PI / 2) / PB_PI * LAT_BINS));
        const int bMax = (int)js::min(LAT_BINS - 1, std::floor((latMax + PB_PI / 2) / PB_PI * LAT_BINS));
        const double cosLat_ = js::or_default(pb_cos(la), 0.01);
        const double lonSpan = radius / cosLat_;
        const int lMin = (int)std::floor((lo - lonSpan + PB_PI) / (2 * PB_PI) * LON_BINS);
        const int lMax = (int)std::floor((lo + lonSpan + PB_PI) / (2 * PB_PI) * LON_BINS);
        double landCount = 0, totalCount = 0, elevSum = 0;
        const double cosRadius = pb_cos(radius);
        const double sinLat0 = pb_sin(la), cosLat0 = pb_cos(la);
        for (int bi = bMin; bi <= bMax; bi++)
            for (int li = lMin; li <= lMax; li++) {
                const int lj = ((li % LON_BINS) + LON_BINS) % LON_BINS;
                const int bin = bi * LON_BINS + lj;
                for (uint32_t k = binOffset[bin]; k < binOffset[bin + 1]; k++) {
                    const int r = (int)indices[k];
                    const double dlon = lon[r] - lo;
                    const double cosDist = sinLat0 * sinLat[r] + cosLat0 * cosLat[r] * pb_cos(dlon);
                    if (cosDist >= cosRadius) {
                        totalCount++;
                        if (isLand[r]) landCount++;
                        elevSum += js::max(0, elev[r]);
                    }
                }
            }
        if (totalCount == 0) { *landFrac = 0; *avgElev = 0; return; }
        *landFrac = landCount / totalCount;
        *avgElev = elevSum / totalCount;
    }
};

// js/wind.js:174-232
struct Itcz { Spline spline; F64 lons, lats; };
Itcz computeITCZ(const GeoIndex& geo, bool summer) {
    const int NUM_LON = 72;
    const double sampleRadius = 20 * DEG;
    const double sign = summer ? 1 : -1;
    F64 lons(NUM_LON), rawLats(NUM_LON);
    for (int i = 0; i < NUM_LON; i++) {
        const double lon = -PB_PI + (i + 0.5) * (2 * PB_PI / NUM_LON);
        lons[i] = lon;
        double landSum = 0, elevSum = 0, samples = 0;
        for (int deg = 5; deg <= 20; deg += 5) {
            const double lat = deg * sign * DEG;
            double landFrac, avgElev;
            geo.sample(lat, lon, sampleRadius, &landFrac, &avgElev);
            landSum += landFrac;
            elevSum += avgElev;
            samples++;
        }
        const double avgLand = landSum / samples;
        const double avgElev = elevSum / samples;
        const double landPull = js::min(1, avgLand * 2);
        const double itczDeg = 5 + landPull * 15 - elevToHeightKm(avgElev) * 1.5;
        const double clampedDeg = js::max(5, js::min(20, itczDeg));
        rawLats[i] = clampedDeg * sign * DEG;
    }
    F64 lats(rawLats), tmp(NUM_LON);
    for (int pass = 0; pass < 3; pass++) {
        for (int i = 0; i < NUM_LON; i++) {
            const int p = (i - 1 + NUM_LON) % NUM_LON, n = (i + 1) % NUM_LON;
            tmp[i] = 0.25 * lats[p] + 0.5 * lats[i] + 0.25 * lats[n];
        }
        lats = tmp;
    }
    const double clampMin = (sign > 0 ? 5 : -20) * DEG;
    const double clampMax = (sign > 0 ? 20 : -5) * DEG;
    for (int i = 0; i < NUM_LON; i++) lats[i] = js::max(clampMin, js::min(clampMax, lats[i]));
    Itcz out;
    out.spline = buildPeriodicSpline(lons, lats);
    out.lons = lons; out.lats = lats;
    return out;
}

// js/wind.js:239-301
double regionPressure(double lat, double lon, const Spline& itczSpline, bool summer, double landFrac,
                      double elevation, const SimplexNoise* noiseFn, double px, double py, double pz) {
    const double itczLat = evaluateSpline(itczSpline, lon);
    const double latDeg = lat * RAD;
    const double seasonSign = summer ? 1 : -1;
    double p = 1013;
    const double dItcz = (lat - itczLat) * RAD;
    { const double q = dItcz / 8; p -= 15 * pb_exp(-0.5 * (q * q)); }
    const double shiftDeg = seasonSign * 5;
    const double nhSubHigh = 30 + shiftDeg;
    const double shSubHigh = -(30 - shiftDeg);
    const double highIntensity = 12 * (1 - 0.3 * landFrac);
    { const double q = (latDeg - nhSubHigh) / 10; p += highIntensity * pb_exp(-0.5 * (q * q)); }
    { const double q = (latDeg - shSubHigh) / 10; p += highIntensity * pb_exp(-0.5 * (q * q)); }
    { const double q = (latDeg - 60) / 10; p -= 10 * pb_exp(-0.5 * (q * q)); }
    { const double q = (latDeg + 60) / 10; p -= 10 * pb_exp(-0.5 * (q * q)); }
    { const double q = (latDeg - 85) / 8; p += 8 * pb_exp(-0.5 * (q * q)); }
    { const double q = (latDeg + 85) / 8; p += 8 * pb_exp(-0.5 * (q * q)); }
    const double continentalScale = smoothstep(0.2, 0.5, landFrac);
    if (continentalScale > 0.001) {
        const double absLatDeg = std::fabs(lat) * RAD;
        const double latFactor = absLatDeg < 15 ? 0
            : absLatDeg < 30 ? 0.75 * smoothstep(15, 30, absLatDeg)
            : absLatDeg < 45 ? 0.75 + 0.25 * smoothstep(30, 45, absLatDeg)
            : absLatDeg < 60 ? 1
            : absLatDeg < 90 ? smoothstep(90, 60, absLatDeg)
            : 0;
        const bool isSummerHemisphere = (seasonSign > 0 && lat > 0) || (seasonSign < 0 && lat < 0);
        if (isSummerHemisphere) p -= 10 * latFactor * continentalScale;
        else p += 14 * latFactor * continentalScale;
    }
    p -= 3 * elevToHeightKm(js::max(0, elevation));
    if (noiseFn) p += noiseFn->fbm(px * 2, py * 2, pz * 2, 3) * 2;
    return p;
}

// js/ocean.js:168-189 (and the same shape with other masks)
void smoothOcean(const OMesh& mesh, F32& field, const U8& isOcean, int passes) {
    const int N = mesh.N;
    F32 tmp(N, 0.f);
    for (int pass = 0; pass < passes; pass++) {
        for (int r = 0; r < N; r++) {
            if (!isOcean[r]) { tmp[r] = field[r]; continue; }
            double sum = field[r];
            int count = 1;
            for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                const int nb = mesh.adjList[ni];
                if (isOcean[nb]) { sum += field[nb]; count++; }
            }
            tmp[r] = js::f32(sum / count);
        }
        field = tmp;
    }
}

// js/heuristic-precip.js:16-38
double zonalBase(double distDeg) {
    if (distDeg < 5) return 1.0;
    else if (distDeg < 10) return 1.0 - 0.65 * smoothstep(5, 10, distDeg);
    else if (distDeg < 33) return 0.35 - 0.33 * smoothstep(10, 28, distDeg);
    else if (distDeg < 55) return 0.02 + 0.48 * smoothstep(33, 55, distDeg);
    else if (distDeg < 70) return 0.5 - 0.2 * smoothstep(55, 70, distDeg);
    else return 0.3 - 0.2 * smoothstep(70, 90, distDeg);
}
// js/heuristic-precip.js:52-86
void heuristicWind(double distFromItczDeg, bool isNorthOfItcz, double* we, double* wn) {
    const double hemiSign = isNorthOfItcz ? 1 : -1;
    if (distFromItczDeg < 5) {
        *we = 0;
        *wn = -hemiSign * 0.1;
    } else if (distFromItczDeg < 30) {
        const double tradeStrength = smoothstep(5, 15, distFromItczDeg) * (1 - smoothstep(25, 32, distFromItczDeg));
        *we = -tradeStrength * 0.8;
        *wn = -hemiSign * tradeStrength * 0.3;
    } else if (distFromItczDeg < 60) {
        const double westStrength = smoothstep(30, 40, distFromItczDeg) * (1 - smoothstep(55, 65, distFromItczDeg));
        *we = westStrength * 0.9;
        *wn = hemiSign * westStrength * 0.25;
    } else {
        const double polarStrength = smoothstep(60, 70, distFromItczDeg);
        *we = -polarStrength * 0.4;
        *wn = -hemiSign * polarStrength * 0.15;
    }
}

}  // namespace

struct OracleClimate {
    OMesh mesh;
    int N;
    const float* xyz;
    std::map<std::string, F32> f;    // float32 result fields under the reference's key names
    std::map<std::string, I32> i;    // int32 fields
    std::map<std::string, U8> u;     // uint8 fields
    bool haveWind = false, haveOcean = false, havePrecip = false, haveTemp = false;

    OracleClimate(const OMesh& m, const float* x) : mesh(m), N(m.N), xyz(x) {}

    // js/wind.js:306-339
    void computeGradients(const F32& P, F32& gradE, F32& gradN) {
        const F32 &eX = f["r_eastX"], &eY = f["r_eastY"], &eZ = f["r_eastZ"];
        const F32 &nX = f["r_northX"], &nY = f["r_northY"], &nZ = f["r_northZ"];
        for (int r = 0; r < N; r++) {
            const double px = xyz[3 * r], py = xyz[3 * r + 1], pz = xyz[3 * r + 2];
            const double ex = eX[r], ey = eY[r], ez = eZ[r];
            const double nx = nX[r], ny = nY[r], nz = nZ[r];
            const double pHere = P[r];
            double sumEP = 0, sumEE = 0, sumNP = 0, sumNN = 0;
            for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                const int nb = mesh.adjList[ni];
                const double dx = xyz[3 * nb] - px, dy = xyz[3 * nb + 1] - py, dz = xyz[3 * nb + 2] - pz;
                const double de = dx * ex + dy * ey + dz * ez;
                const double dn = dx * nx + dy * ny + dz * nz;
                const double dp = P[nb] - pHere;
                sumEP += de * dp; sumEE += de * de; sumNP += dn * dp; sumNN += dn * dn;
            }
            gradE[r] = js::f32(sumEE > 1e-12 ? sumEP / sumEE : 0);
            gradN[r] = js::f32(sumNN > 1e-12 ? sumNP / sumNN : 0);
        }
    }

    // FIFO BFS with hop counts (wind.js:525-538, 575-586; ocean.js:58-80)
    void bfs(I32& dist, std::vector<int>& queue, const U8& passable) {
        size_t head = 0;
        while (head < queue.size()) {
            const int r = queue[head++];
            const int d = dist[r] + 1;
            for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                const int nb = mesh.adjList[ni];
                if (passable[nb] && dist[nb] == -1) { dist[nb] = d; queue.push_back(nb); }
            }
        }
    }

    // js/wind.js:394-687
    void computeWind(const float* elev, const int32_t* plateIsOceanIds, int nIds, const int32_t* r_plate,
                     double noiseSeed, double axialTilt) {
        (void)axialTilt;
        const double avgEdgeKm = (PB_PI * 6371) / std::sqrt((double)N);
        SimplexNoise noise(noiseSeed);
        std::unordered_set<int32_t> plateIsOcean(plateIsOceanIds, plateIsOceanIds + nIds);

        F32 &r_lat = f["r_lat"], &r_lon = f["r_lon"], &r_sinLat = f["r_sinLat"], &r_cosLat = f["r_cosLat"];
        U8& r_isLand = u["r_isLand"];
        F32 &eX = f["r_eastX"], &eY = f["r_eastY"], &eZ = f["r_eastZ"], &nX = f["r_northX"], &nY = f["r_northY"], &nZ = f["r_northZ"];
        for (F32* a : {&r_lat, &r_lon, &r_sinLat, &r_cosLat, &eX, &eY, &eZ, &nX, &nY, &nZ}) a->assign(N, 0.f);
        r_isLand.assign(N, 0);
        for (int r = 0; r < N; r++) {
            const double x = xyz[3 * r], y = xyz[3 * r + 1], z = xyz[3 * r + 2];
            r_lat[r] = js::f32(pb_asin(js::max(-1, js::min(1, y))));
            r_lon[r] = js::f32(pb_atan2(x, z));
            r_sinLat[r] = js::f32(y);
            r_cosLat[r] = js::f32(js::or_default(std::sqrt(1 - y * y), 0.01));
            r_isLand[r] = elev[r] > 0 ? 1 : 0;
            double ex = z, ey = 0, ez = -x;
            double elen = std::sqrt(ex * ex + ez * ez);
            if (elen < 1e-10) { ex = 1; ez = 0; elen = 1; }
            ex /= elen; ez /= elen;
            double nx = y * ez - z * ey;
            double ny = z * ex - x * ez;
            double nz = x * ey - y * ex;
            const double nlen = js::or_default(std::sqrt(nx * nx + ny * ny + nz * nz), 1);
            nx /= nlen; ny /= nlen; nz /= nlen;
            eX[r] = js::f32(ex); eY[r] = js::f32(ey); eZ[r] = js::f32(ez);
            nX[r] = js::f32(nx); nY[r] = js::f32(ny); nZ[r] = js::f32(nz);
        }

        GeoIndex geo(r_lat, r_lon, r_sinLat, r_cosLat, elev, r_isLand, N);
        Itcz itczSummer = computeITCZ(geo, true);
        Itcz itczWinter = computeITCZ(geo, false);

        // main ocean = largest connected component of non-land cells (:481-508)
        I32& r_oceanLabel = i["r_oceanLabel"];
        r_oceanLabel.assign(N, -1);
        int mainOceanLabel = -1, mainOceanSize = 0, nextLabel = 0;
        for (int r = 0; r < N; r++) {
            if (r_isLand[r] || r_oceanLabel[r] >= 0) continue;
            const int label = nextLabel++;
            int size = 0;
            std::vector<int> q{r};
            r_oceanLabel[r] = label;
            size_t head = 0;
            while (head < q.size()) {
                const int cur = q[head++];
                size++;
                for (int ni = mesh.adjOffset[cur]; ni < mesh.adjOffset[cur + 1]; ni++) {
                    const int nb = mesh.adjList[ni];
                    if (!r_isLand[nb] && r_oceanLabel[nb] == -1) { r_oceanLabel[nb] = label; q.push_back(nb); }
                }
            }
            if (size > mainOceanSize) { mainOceanSize = size; mainOceanLabel = label; }
        }

        // coast distance through land from the main-ocean coastline (:510-538)
        I32& r_coastDist = i["r_coastDistLand"];
        r_coastDist.assign(N, -1);
        {
            std::vector<int> q;
            for (int r = 0; r < N; r++) {
                if (!r_isLand[r]) continue;
                for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                    const int nb = mesh.adjList[ni];
                    if (!r_isLand[nb] && r_oceanLabel[nb] == mainOceanLabel) { r_coastDist[r] = 0; q.push_back(r); break; }
                }
            }
            bfs(r_coastDist, q, r_isLand);
        }
        const double CONT_RANGE_KM = 2000;
        F32& r_cont = f["r_continentality"];
        r_cont.assign(N, 0.f);
        for (int r = 0; r < N; r++)
            if (r_isLand[r] && r_coastDist[r] >= 0) r_cont[r] = js::f32(smoothstep(0, CONT_RANGE_KM, r_coastDist[r] * avgEdgeKm));
        const int contSmoothPasses = (int)js::max(1, js::round(100 / avgEdgeKm));
        oracle_smooth_field(mesh, r_cont.data(), contSmoothPasses);

        // plate-based continentality (:556-593)
        F32& r_pcont = f["r_plateContinentality"];
        r_pcont.assign(N, 0.f);
        I32& r_plateDist = i["r_plateDist"];
        r_plateDist.assign(N, -1);
        U8 contPlate(N);
        for (int r = 0; r < N; r++) contPlate[r] = plateIsOcean.count(r_plate[r]) ? 0 : 1;
        {
            std::vector<int> q;
            for (int r = 0; r < N; r++) {
                if (!contPlate[r]) continue;
                for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++)
                    if (!contPlate[mesh.adjList[ni]]) { r_plateDist[r] = 0; q.push_back(r); break; }
            }
            bfs(r_plateDist, q, contPlate);
        }
        for (int r = 0; r < N; r++)
            if (contPlate[r] && r_plateDist[r] >= 0) r_pcont[r] = js::f32(smoothstep(0, CONT_RANGE_KM, r_plateDist[r] * avgEdgeKm));
        oracle_smooth_field(mesh, r_pcont.data(), contSmoothPasses);

        F32 r_gradE(N), r_gradN(N);
        for (int s = 0; s < 2; s++) {
            const bool summer = s == 0;
            const std::string name = summer ? "summer" : "winter";
            const Itcz& itcz = summer ? itczSummer : itczWinter;
            F32 r_pressure(N);
            for (int r = 0; r < N; r++)
                r_pressure[r] = js::f32(regionPressure(r_lat[r], r_lon[r], itcz.spline, summer, r_cont[r], elev[r], &noise,
                                                       xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]));
            const int pressSmoothPasses = (int)js::max(1, js::round(75 / avgEdgeKm));
            oracle_smooth_field(mesh, r_pressure.data(), pressSmoothPasses);
            computeGradients(r_pressure, r_gradE, r_gradN);

            // pressureToWind (:343-378)
            F32 &r_windE = f["r_wind_east_" + name], &r_windN = f["r_wind_north_" + name], &r_windSpeed = f["r_wind_speed_" + name];
            r_windE.assign(N, 0.f); r_windN.assign(N, 0.f); r_windSpeed.assign(N, 0.f);
            const double sin5 = pb_sin(5 * DEG);
            for (int r = 0; r < N; r++) {
                const double pgfE = -(double)r_gradE[r], pgfN = -(double)r_gradN[r];
                const double sinLat = r_sinLat[r];
                const double absSinLat = std::fabs(sinLat);
                const double geoAngle = 70 * DEG * smoothstep(0, sin5, absSinLat);
                const double frictionAngle = 20 * DEG;
                const double sign = sinLat >= 0 ? -1 : 1;
                const double totalAngle = sign * (geoAngle - frictionAngle);
                const double cosA = pb_cos(totalAngle), sinA = pb_sin(totalAngle);
                const double we = (pgfE * cosA - pgfN * sinA) * 0.6;
                const double wn = (pgfE * sinA + pgfN * cosA) * 0.6;
                r_windE[r] = js::f32(we);
                r_windN[r] = js::f32(wn);
                r_windSpeed[r] = js::f32(std::sqrt(we * we + wn * wn));
            }
            const double maxSpeed = oracle_percentile(r_windSpeed.data(), N, 0.95);
            for (int r = 0; r < N; r++) r_windSpeed[r] = js::f32(js::min(1, r_windSpeed[r] / maxSpeed));
            F32& dev = f["r_pressure_" + name];
            dev.assign(N, 0.f);
            for (int r = 0; r < N; r++) dev[r] = js::f32((double)r_pressure[r] - 1013);
        }

        const int ITCZ_SAMPLES = 360;
        F32 &itczLons = f["itczLons"], &ls = f["itczLatsSummer"], &lw = f["itczLatsWinter"];
        itczLons.assign(ITCZ_SAMPLES, 0.f); ls.assign(ITCZ_SAMPLES, 0.f); lw.assign(ITCZ_SAMPLES, 0.f);
        for (int k = 0; k < ITCZ_SAMPLES; k++) {
            const double lon = -PB_PI + (k + 0.5) * (2 * PB_PI / ITCZ_SAMPLES);
            itczLons[k] = js::f32(lon);
            ls[k] = js::f32(evaluateSpline(itczSummer.spline, lon));
            lw[k] = js::f32(evaluateSpline(itczWinter.spline, lon));
        }
        haveWind = true;
    }

    // js/ocean.js:204-382
    void computeOceanCurrents(const float* elev) {
        (void)elev;
        const double avgEdgeKm = (PB_PI * 6371) / std::sqrt((double)N);
        const F32 &r_lat = f["r_lat"], &r_lon = f["r_lon"];
        const U8& r_isLand = u["r_isLand"];
        const F32 &eX = f["r_eastX"], &eY = f["r_eastY"], &eZ = f["r_eastZ"];
        U8& r_isOcean = u["r_isOcean"];
        r_isOcean.assign(N, 0);
        for (int r = 0; r < N; r++) r_isOcean[r] = r_isLand[r] ? 0 : 1;

        // computeCoastFields (:13-84)
        std::vector<int> westSeeds, eastSeeds, allCoastSeeds;
        for (int r = 0; r < N; r++) {
            if (!r_isOcean[r]) continue;
            double landDirX = 0, landDirY = 0, landDirZ = 0;
            bool hasLandNeighbor = false;
            for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                const int nb = mesh.adjList[ni];
                if (!r_isOcean[nb]) {
                    hasLandNeighbor = true;
                    landDirX += (double)xyz[3 * nb] - (double)xyz[3 * r];
                    landDirY += (double)xyz[3 * nb + 1] - (double)xyz[3 * r + 1];
                    landDirZ += (double)xyz[3 * nb + 2] - (double)xyz[3 * r + 2];
                }
            }
            if (!hasLandNeighbor) continue;
            allCoastSeeds.push_back(r);
            const double normalE = landDirX * eX[r] + landDirY * eY[r] + landDirZ * eZ[r];
            if (normalE < -0.2) westSeeds.push_back(r);
            else if (normalE > 0.2) eastSeeds.push_back(r);
            else { if (normalE <= 0) westSeeds.push_back(r); else eastSeeds.push_back(r); }
        }
        auto bfsDistance = [&](std::vector<int>& seeds, I32& dist) {
            dist.assign(N, -1);
            for (int s : seeds) dist[s] = 0;
            std::vector<int> q(seeds);
            bfs(dist, q, r_isOcean);
        };
        I32 &r_coastDist = i["r_oceanCoastDist"], &r_westCoastDist = i["r_westCoastDist"], &r_eastCoastDist = i["r_eastCoastDist"];
        bfsDistance(allCoastSeeds, r_coastDist);
        bfsDistance(westSeeds, r_westCoastDist);
        bfsDistance(eastSeeds, r_eastCoastDist);

        // hasCircumpolarChannel (:88-111)
        auto circumpolar = [&](double targetLat, double bandWidth) {
            const int NUM_BINS = 72;
            uint8_t has[72] = {0};
            const double latMin = targetLat - bandWidth, latMax = targetLat + bandWidth;
            for (int r = 0; r < N; r++) {
                if (!r_isOcean[r]) continue;
                const double lat = r_lat[r];
                if (lat < latMin || lat > latMax) continue;
                double bin = std::floor(((r_lon[r] + PB_PI) / (2 * PB_PI)) * NUM_BINS);
                bin = std::fmod(std::fmod(bin, NUM_BINS) + NUM_BINS, NUM_BINS);
                has[(int)bin] = 1;
            }
            for (int k = 0; k < NUM_BINS; k++) if (!has[k]) return false;
            return true;
        };
        const bool circumpolarNH = circumpolar(60 * DEG, 5 * DEG);
        const bool circumpolarSH = circumpolar(-60 * DEG, 5 * DEG);
        const double coastThreshold = js::max(5, js::round(std::sqrt((double)N) * 0.035));
        const double warmthRange = coastThreshold * 2;

        for (int s = 0; s < 2; s++) {
            const bool summer = s == 0;
            const std::string name = summer ? "summer" : "winter";
            ItczLookup itczLookup(f["itczLons"], summer ? f["itczLatsSummer"] : f["itczLatsWinter"]);
            const double seasonalShiftDeg = summer ? 5 : -5;
            F32 &currentE = f["r_ocean_current_east_" + name], &currentN = f["r_ocean_current_north_" + name];
            currentE.assign(N, 0.f); currentN.assign(N, 0.f);
            for (int r = 0; r < N; r++) {
                if (!r_isOcean[r]) continue;
                const double lat = r_lat[r];
                const double absLatDeg = std::fabs(lat) / DEG;
                const double lon = r_lon[r];
                const double hemisphereSign = lat >= 0 ? 1 : -1;
                const double bandLatDeg = std::fabs(lat / DEG - seasonalShiftDeg);
                const double itczLat = itczLookup(lon);
                const double distFromItcz = std::fabs(lat - itczLat) / DEG;
                double baseE;
                if (distFromItcz < 3) baseE = 1 - 2 * smoothstep(0, 3, distFromItcz);
                else if (bandLatDeg < 30) baseE = -1;
                else if (bandLatDeg < 35) baseE = -1 + 2 * smoothstep(30, 35, bandLatDeg);
                else if (bandLatDeg < 58) baseE = 1;
                else if (bandLatDeg < 65) baseE = 1 - 1.5 * smoothstep(58, 65, bandLatDeg);
                else baseE = -0.5;
                currentE[r] = js::f32(baseE);
                currentN[r] = 0;
                const double wDist = r_westCoastDist[r], eDist = r_eastCoastDist[r];
                if (wDist >= 0 && wDist < coastThreshold) {
                    const double t = 1 - wDist / coastThreshold;
                    const double strength = t * t * 2.0;
                    currentN[r] = js::f32(currentN[r] + hemisphereSign * strength);
                    currentE[r] = js::f32(currentE[r] * (1 - t * t * 0.7));
                }
                if (eDist >= 0 && eDist < coastThreshold) {
                    const double t = 1 - eDist / coastThreshold;
                    const double strength = t * t * 0.8;
                    currentN[r] = js::f32(currentN[r] - hemisphereSign * strength);
                    currentE[r] = js::f32(currentE[r] * (1 - t * t * 0.5));
                }
                const bool isCircumpolar = (lat > 0 && circumpolarNH) || (lat < 0 && circumpolarSH);
                if (isCircumpolar && absLatDeg >= 55 && absLatDeg <= 75) {
                    const double cStrength = 1 - std::fabs(absLatDeg - 65) / 10;
                    currentE[r] = js::f32(currentE[r] * (1 - cStrength) + 1.5 * cStrength);
                    currentN[r] = js::f32(currentN[r] * (1 - cStrength * 0.8));
                }
            }
            const int oceanSmoothPasses = (int)js::max(2, js::round(125 / avgEdgeKm));
            smoothOcean(mesh, currentE, r_isOcean, oceanSmoothPasses);
            smoothOcean(mesh, currentN, r_isOcean, oceanSmoothPasses);
            for (int r = 0; r < N; r++) if (!r_isOcean[r]) { currentE[r] = 0; currentN[r] = 0; }

            // classifyWarmth (:120-164)
            F32& r_warmth = f["r_ocean_warmth_" + name];
            r_warmth.assign(N, 0.f);
            for (int r = 0; r < N; r++) {
                if (!r_isOcean[r]) continue;
                const double bandLatDeg = std::fabs(r_lat[r] / DEG - seasonalShiftDeg);
                double cellSign;
                if (bandLatDeg < 28) cellSign = 1;
                else if (bandLatDeg < 35) cellSign = 1 - 2 * smoothstep(28, 35, bandLatDeg);
                else if (bandLatDeg < 55) cellSign = -1;
                else if (bandLatDeg < 65) cellSign = -1 + 2 * smoothstep(55, 65, bandLatDeg);
                else cellSign = 1;
                const double wDist = r_westCoastDist[r], eDist = r_eastCoastDist[r];
                double warm = 0;
                if (wDist >= 0 && wDist < warmthRange) { const double t = 1 - wDist / warmthRange; warm += cellSign * t * t; }
                if (eDist >= 0 && eDist < warmthRange) { const double t = 1 - eDist / warmthRange; warm -= cellSign * t * t; }
                r_warmth[r] = js::f32(js::max(-1, js::min(1, warm)));
            }
            const int warmthSmoothPasses = (int)js::max(3, js::round(900 / avgEdgeKm));
            smoothOcean(mesh, r_warmth, r_isOcean, warmthSmoothPasses);

            F32& r_speed = f["r_ocean_speed_" + name];
            r_speed.assign(N, 0.f);
            F32 oceanSpeeds(N, 0.f);
            int oceanCount = 0;
            for (int r = 0; r < N; r++) {
                const double spd = std::sqrt((double)currentE[r] * currentE[r] + (double)currentN[r] * currentN[r]);
                r_speed[r] = js::f32(spd);
                if (r_isOcean[r] && spd > 0) oceanSpeeds[oceanCount++] = js::f32(spd);
            }
            const double p95 = oracle_percentile(oceanSpeeds.data(), oceanCount, 0.95);
            for (int r = 0; r < N; r++) r_speed[r] = js::f32(js::min(1, r_speed[r] / p95));
        }
        haveOcean = true;
    }

    // js/heuristic-precip.js:90-108
    void computeHeuristicWindField(const ItczLookup& itczLookup, F32& hWindE, F32& hWindN) {
        const F32 &r_lat = f["r_lat"], &r_lon = f["r_lon"];
        hWindE.assign(N, 0.f); hWindN.assign(N, 0.f);
        for (int r = 0; r < N; r++) {
            const double lat = r_lat[r];
            const double itczLat = itczLookup(r_lon[r]) * 0.3;
            const double signedDist = lat - itczLat;
            const double distDeg = std::fabs(signedDist) / DEG;
            double we, wn;
            heuristicWind(distDeg, signedDist > 0, &we, &wn);
            hWindE[r] = js::f32(we); hWindN[r] = js::f32(wn);
        }
    }

    // js/heuristic-precip.js:119-269
    void computeHeuristicPrecipitation(const float* elev, const F32& r_elevGradE, const F32& r_elevGradN, F32 out[2]) {
        const F32 &r_lat = f["r_lat"], &r_lon = f["r_lon"], &r_cont = f["r_continentality"];
        const U8& r_isLand = u["r_isLand"];
        const I32& r_coastDistLand = i["r_coastDistLand"];
        const F32 &eX = f["r_eastX"], &eY = f["r_eastY"], &eZ = f["r_eastZ"];
        const double avgEdgeKm = (PB_PI * 6371) / std::sqrt((double)N);
        F32& r_westCoast = f["r_westCoast"];
        r_westCoast.assign(N, 0.f);
        for (int r = 0; r < N; r++) {
            if (!r_isLand[r] || r_coastDistLand[r] != 0) continue;
            double oceanDotEast = 0;
            int count = 0;
            for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                const int nb = mesh.adjList[ni];
                if (!r_isLand[nb]) {
                    const double dx = (double)xyz[3 * nb] - (double)xyz[3 * r];
                    const double dy = (double)xyz[3 * nb + 1] - (double)xyz[3 * r + 1];
                    const double dz = (double)xyz[3 * nb + 2] - (double)xyz[3 * r + 2];
                    oceanDotEast += dx * eX[r] + dy * eY[r] + dz * eZ[r];
                    count++;
                }
            }
            if (count > 0) r_westCoast[r] = oceanDotEast < 0 ? 1 : -1;
        }
        const int wcPasses = (int)js::max(2, js::round(300 / avgEdgeKm));
        F32 wcTmp(N, 0.f);
        for (int pass = 0; pass < wcPasses; pass++) {
            for (int r = 0; r < N; r++) {
                if (!r_isLand[r]) { wcTmp[r] = 0; continue; }
                double sum = r_westCoast[r];
                int count = 1;
                for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                    const int nb = mesh.adjList[ni];
                    if (r_isLand[nb]) { sum += r_westCoast[nb]; count++; }
                }
                wcTmp[r] = js::f32(sum / count);
            }
            r_westCoast = wcTmp;
        }
        for (int s = 0; s < 2; s++) {
            const bool isSummer = s == 0;
            ItczLookup itczLookup(f["itczLons"], isSummer ? f["itczLatsSummer"] : f["itczLatsWinter"]);
            F32& precip = out[s];
            precip.assign(N, 0.f);
            for (int r = 0; r < N; r++) {
                const double lat = r_lat[r], lon = r_lon[r];
                const double itczLat = itczLookup(lon) * 0.3;
                const double signedDist = lat - itczLat;
                const double distFromItczDeg = std::fabs(signedDist) / DEG;
                const bool isNorthOfItcz = signedDist > 0;
                const double zonal = zonalBase(distFromItczDeg);
                const double absLatDeg = std::fabs(lat) / DEG;
                const bool inSummerHemi = isSummer ? (lat >= 0) : (lat < 0);
                double seasonMod = inSummerHemi ? 1.1 : 0.9;
                if (inSummerHemi && absLatDeg > 22 && absLatDeg < 45) {
                    const double medSuppress = smoothstep(22, 30, absLatDeg) * (1 - smoothstep(38, 45, absLatDeg));
                    const double wc = r_westCoast[r];
                    const double strength = 0.15 + wc * 0.20;
                    seasonMod *= (1 - medSuppress * js::max(0, strength));
                }
                double contMod = 1.0;
                const double cont = r_isLand[r] ? (double)r_cont[r] : 0;
                if (cont > 0) contMod = 1.0 - cont * cont * 0.65;
                double oroMod = 1.0;
                if (r_isLand[r] && elev[r] > 0) {
                    double we, wn;
                    heuristicWind(distFromItczDeg, isNorthOfItcz, &we, &wn);
                    const double windDotGrad = we * r_elevGradE[r] + wn * r_elevGradN[r];
                    if (windDotGrad > 0) {
                        const double uplift = js::min(1, windDotGrad * 15);
                        oroMod = 1.0 + uplift * 0.6;
                    } else {
                        const double heightKm = elevToHeightKm(js::max(0, elev[r]));
                        const double heightScale = js::min(1, heightKm / 3);
                        const double shadow = js::min(1, -windDotGrad * 18);
                        oroMod = js::max(0.3, 1.0 - shadow * 0.7 * heightScale);
                    }
                }
                double distMod = 1.0;
                if (r_isLand[r] && r_coastDistLand[r] > 0) {
                    const double distKm = r_coastDistLand[r] * avgEdgeKm;
                    if (distKm > 2000) distMod = js::max(0.03, 1 - smoothstep(2000, 3000, distKm));
                }
                precip[r] = js::f32(js::max(0.05, zonal * seasonMod * contMod * oroMod * distMod));
            }
            const int smoothPasses = (int)js::max(1, js::round(100 / avgEdgeKm));
            oracle_smooth_field(mesh, precip.data(), smoothPasses);
        }
    }

    // js/precipitation.js:19-52
    void computeWindConvergence(const F32& wX, const F32& wY, const F32& wZ, F32& convergence) {
        convergence.assign(N, 0.f);
        for (int r = 0; r < N; r++) {
            const double wdx = wX[r], wdy = wY[r], wdz = wZ[r];
            double conv = 0;
            int count = 0;
            for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                const int nb = mesh.adjList[ni];
                const double dx = (double)xyz[3 * nb] - (double)xyz[3 * r];
                const double dy = (double)xyz[3 * nb + 1] - (double)xyz[3 * r + 1];
                const double dz = (double)xyz[3 * nb + 2] - (double)xyz[3 * r + 2];
                conv -= (wX[nb] + wdx) * dx + (wY[nb] + wdy) * dy + (wZ[nb] + wdz) * dz;
                count++;
            }
            convergence[r] = js::f32(count > 0 ? conv / count : 0);
        }
    }

    // js/precipitation.js:59-182
    F32 advectMoisture(const F32& r_heightKm, const F32& r_windE, const F32& r_windN, const F32& wX, const F32& wY,
                       const F32& wZ, const F32& r_oceanWarmth, int maxHops) {
        const U8& r_isLand = u["r_isLand"];
        const I32& r_coastDistLand = i["r_coastDistLand"];
        F32 moisture(N, 0.f);
        for (int r = 0; r < N; r++) {
            if (!r_isLand[r]) {
                const double warmth = r_oceanWarmth[r];
                moisture[r] = js::f32(0.4 + 0.35 * js::max(0, warmth));
                continue;
            }
            if (r_coastDistLand[r] != 0) continue;
            double warmthSum = 0;
            int oceanCount = 0;
            double oceanDirX = 0, oceanDirY = 0, oceanDirZ = 0;
            for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                const int nb = mesh.adjList[ni];
                if (!r_isLand[nb]) {
                    oceanCount++;
                    warmthSum += r_oceanWarmth[nb];
                    oceanDirX += (double)xyz[3 * nb] - (double)xyz[3 * r];
                    oceanDirY += (double)xyz[3 * nb + 1] - (double)xyz[3 * r + 1];
                    oceanDirZ += (double)xyz[3 * nb + 2] - (double)xyz[3 * r + 2];
                }
            }
            if (oceanCount == 0) continue;
            const double avgWarmth = warmthSum / oceanCount;
            const double wdx = wX[r], wdy = wY[r], wdz = wZ[r];
            const double windDotOcean = wdx * oceanDirX + wdy * oceanDirY + wdz * oceanDirZ;
            const double onshore = windDotOcean < 0 ? 1.0 : 0.25;
            const double warmthFactor = 0.5 + 0.5 * js::max(-0.8, js::min(1, avgWarmth));
            moisture[r] = js::f32(onshore * warmthFactor);
        }
        const double depletionBase = 1 - pb_pow(0.78, 1.0 / maxHops);
        F32 a = moisture, b(N, 0.f);
        F32 *src = &a, *dst = &b;
        for (int iter = 0; iter < maxHops; iter++) {
            for (int r = 0; r < N; r++) {
                if (!r_isLand[r]) { (*dst)[r] = (*src)[r]; continue; }
                const double we = r_windE[r], wn = r_windN[r];
                if (we * we + wn * wn < 1e-6) { (*dst)[r] = (*src)[r]; continue; }
                double upwindMoisture = 0, upwindWeight = 0, upwindHeightSum = 0;
                const double heightHere = r_heightKm[r];
                for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                    const int nb = mesh.adjList[ni];
                    const double dx = (double)xyz[3 * r] - (double)xyz[3 * nb];
                    const double dy = (double)xyz[3 * r + 1] - (double)xyz[3 * nb + 1];
                    const double dz = (double)xyz[3 * r + 2] - (double)xyz[3 * nb + 2];
                    const double dot = wX[nb] * dx + wY[nb] * dy + wZ[nb] * dz;
                    if (dot > 0) {
                        upwindMoisture += (*src)[nb] * dot;
                        upwindHeightSum += r_heightKm[nb] * dot;
                        upwindWeight += dot;
                    }
                }
                if (upwindWeight > 0) {
                    const double incoming = upwindMoisture / upwindWeight;
                    const double upwindHeight = upwindHeightSum / upwindWeight;
                    const double heightGain = js::max(0, heightHere - upwindHeight);
                    const double normalizedGain = heightGain * maxHops;
                    const double elevDepletion = js::min(0.8, normalizedGain * 0.55);
                    const double depletion = depletionBase + elevDepletion;
                    const double carried = incoming * js::max(0, 1 - depletion);
                    (*dst)[r] = js::f32(js::max((*src)[r], carried));
                } else (*dst)[r] = (*src)[r];
            }
            std::swap(src, dst);
        }
        return *src;
    }

    // js/precipitation.js:196-684
    void computePrecipitation(const float* elev, double precipitationOffset, double landCoverage) {
        const F32 &r_lat = f["r_lat"], &r_lon = f["r_lon"], &r_cont = f["r_continentality"];
        const U8& r_isLand = u["r_isLand"];
        const F32 &eX = f["r_eastX"], &eY = f["r_eastY"], &eZ = f["r_eastZ"], &nX = f["r_northX"], &nY = f["r_northY"], &nZ = f["r_northZ"];
        const I32& r_coastDistLand = i["r_coastDistLand"];
        const double avgEdgeKm = (PB_PI * 6371) / std::sqrt((double)N);
        const double avgEdgeRad = PB_PI / std::sqrt((double)N);
        const int maxHops = (int)js::max(8, js::min(20, js::round(2000 / avgEdgeKm)));

        const int elevSmoothPasses = (int)js::max(2, js::round(200 / avgEdgeKm));
        F32 r_elevSmoothed(elev, elev + N);
        oracle_smooth_field(mesh, r_elevSmoothed.data(), elevSmoothPasses);
        for (int r = 0; r < N; r++) r_elevSmoothed[r] = js::f32(r_elevSmoothed[r] * 0.6 + elev[r] * 0.4);
        F32 &r_elevGradE = f["r_elevGradE"], &r_elevGradN = f["r_elevGradN"];
        r_elevGradE.assign(N, 0.f); r_elevGradN.assign(N, 0.f);
        computeGradients(r_elevSmoothed, r_elevGradE, r_elevGradN);
        F32 r_heightKm(N);
        for (int r = 0; r < N; r++) r_heightKm[r] = js::f32(elevToHeightKm(js::max(0, elev[r])));

        for (int s = 0; s < 2; s++) {
            const bool summer = s == 0;
            const std::string name = summer ? "summer" : "winter";
            const F32 &r_windE_raw = f["r_wind_east_" + name], &r_windN_raw = f["r_wind_north_" + name];
            const F32& r_pressure = f["r_pressure_" + name];
            const F32& r_oceanWarmth = f["r_ocean_warmth_" + name];
            ItczLookup itczLookup(f["itczLons"], summer ? f["itczLatsSummer"] : f["itczLatsWinter"]);
            F32 hWindE, hWindN;
            computeHeuristicWindField(itczLookup, hWindE, hWindN);
            F32 r_windE(N), r_windN(N), wX(N), wY(N), wZ(N);
            for (int r = 0; r < N; r++) {
                r_windE[r] = js::f32(0.5 * r_windE_raw[r] + 0.5 * hWindE[r]);
                r_windN[r] = js::f32(0.5 * r_windN_raw[r] + 0.5 * hWindN[r]);
            }
            for (int r = 0; r < N; r++) {
                const double we = r_windE[r], wn = r_windN[r];
                wX[r] = js::f32(we * eX[r] + wn * nX[r]);
                wY[r] = js::f32(we * eY[r] + wn * nY[r]);
                wZ[r] = js::f32(we * eZ[r] + wn * nZ[r]);
            }
            F32 r_convergence;
            computeWindConvergence(wX, wY, wZ, r_convergence);
            const int convSmoothPasses = (int)js::max(3, js::round(400 / avgEdgeKm));
            oracle_smooth_field(mesh, r_convergence.data(), convSmoothPasses);
            F32 moisture = advectMoisture(r_heightKm, r_windE, r_windN, wX, wY, wZ, r_oceanWarmth, maxHops);

            F32 precip(N, 0.f);
            F32& rainShadow = f["r_rainshadow_" + name];
            rainShadow.assign(N, 0.f);
            for (int r = 0; r < N; r++) {
                const double lat = r_lat[r], lon = r_lon[r];
                const double absLatDeg = std::fabs(lat) / DEG;
                const double el = elev[r];
                const bool isLand = r_isLand[r] != 0;
                double p = moisture[r];
                const double itczLat = itczLookup(lon);
                const double distFromItcz = std::fabs(lat - itczLat) / DEG;
                const double cont = isLand ? (double)r_cont[r] : 0;
                if (distFromItcz < 15) {
                    const double itczStrength = smoothstep(15, 0, distFromItcz);
                    const double coreBoost = distFromItcz < 5 ? 1.5 : 1.0;
                    p = p * (1 + itczStrength * coreBoost) + itczStrength * 0.3;
                }
                const double conv = r_convergence[r];
                if (conv > 0) {
                    const double convStrength = js::min(1, (conv / avgEdgeRad) * 0.055);
                    p = p * (1 + convStrength * 1.2) + convStrength * moisture[r] * 0.4;
                }
                if (isLand && el > 0) {
                    const double we = r_windE[r], wn = r_windN[r];
                    const double windDotGrad = we * r_elevGradE[r] + wn * r_elevGradN[r];
                    if (windDotGrad > 0) {
                        const double uplift = js::min(1, windDotGrad * 15);
                        p += uplift * 1.0;
                    } else {
                        const double shadow = js::min(1, -windDotGrad * 18);
                        p *= js::max(0.02, 1 - shadow * 0.95);
                    }
                }
                const double pDev = r_pressure[r];
                const bool inLocalSummer = summer ? (lat >= 0) : (lat < 0);
                const double subtropCenter = inLocalSummer ? 30 : 24;
                const double subtropWidth = inLocalSummer ? 16 : 12;
                double subtropPeak = inLocalSummer ? 0.50 : 0.30;
                if (isLand && inLocalSummer) {
                    const double polewardWind = lat >= 0 ? (double)r_windN[r] : -(double)r_windN[r];
                    if (polewardWind > 0) {
                        const double coastDist = r_coastDistLand[r] >= 0 ? r_coastDistLand[r] : maxHops;
                        const double coastProximity = 1 - smoothstep(0, maxHops * 0.4, coastDist);
                        const double monsoonRelief = smoothstep(0, 0.15, polewardWind) * coastProximity;
                        subtropPeak *= (1 - monsoonRelief * 0.7);
                    }
                }
                const double subtropDist = std::fabs(absLatDeg - subtropCenter);
                const double latBandSuppression = subtropDist < subtropWidth ? smoothstep(subtropWidth, 0, subtropDist) * subtropPeak : 0;
                double pressureMod = 0;
                if (pDev > 0) pressureMod = smoothstep(0, 12, pDev) * 0.25;
                else pressureMod = -smoothstep(0, 15, -pDev) * 0.2;
                const double totalSuppression = js::max(0, latBandSuppression + pressureMod);
                if (totalSuppression > 0) p *= js::max(0.05, 1 - totalSuppression);
                else p *= (1 - totalSuppression);
                if (absLatDeg > 40) {
                    const double polarStrength = smoothstep(40, 70, absLatDeg);
                    const double coastDist = r_coastDistLand[r] < 0 ? maxHops : r_coastDistLand[r];
                    const double inlandFade = 1 - smoothstep(0, maxHops, coastDist);
                    const double polarBase = polarStrength * 0.10;
                    const double polarCoastal = polarStrength * 0.20 * inlandFade;
                    p += polarBase + polarCoastal;
                    p *= (1 + polarStrength * 0.15);
                }
                if (isLand && cont > 0) {
                    const double dryness = cont * cont * 0.55;
                    p *= js::max(0.03, 1 - dryness);
                }
                const double heightKm = r_heightKm[r];
                if (isLand && heightKm > 1.5) {
                    const double we = r_windE[r], wn = r_windN[r];
                    const double windDotGrad = we * r_elevGradE[r] + wn * r_elevGradN[r];
                    const double leeCoastHops = js::max(2, js::round(200 / avgEdgeKm));
                    if (windDotGrad < -0.01 && r_coastDistLand[r] >= 0 && r_coastDistLand[r] < leeCoastHops)
                        p += 0.15 * js::min(1, heightKm / 5);
                }
                if (!isLand) {
                    const double highPressureFade = pDev > 0 ? smoothstep(0, 12, pDev) : 0;
                    const double oceanBase = 0.15 * (1 - highPressureFade);
                    p = js::max(p, oceanBase);
                }
                if (isLand && r_coastDistLand[r] > 0) {
                    const double distKm = r_coastDistLand[r] * avgEdgeKm;
                    if (distKm > 2000) {
                        const double fade = 1 - smoothstep(2000, 3000, distKm);
                        p *= js::max(0.03, fade);
                    }
                }
                const double precipMult = 1 + precipitationOffset * 0.5;
                double finalPrecip = p * precipMult;
                if (landCoverage > 0.4) {
                    const double t = (landCoverage - 0.4) / 0.6;
                    finalPrecip *= 1 - t * t * 0.98;
                }
                precip[r] = js::f32(js::max(0, finalPrecip));
            }

            // rain shadow seed (:501-513)
            for (int r = 0; r < N; r++) {
                if (!r_isLand[r] || elev[r] <= 0) continue;
                const double we = r_windE[r], wn = r_windN[r];
                const double windDotGrad = we * r_elevGradE[r] + wn * r_elevGradN[r];
                const double heightKm = r_heightKm[r];
                if (heightKm < 0.8) continue;
                const double heightScale = js::min(1, (heightKm - 0.5) / 2.5);
                if (windDotGrad > 0) rainShadow[r] = js::f32(js::min(1, windDotGrad * 20) * heightScale);
                else if (windDotGrad < 0) rainShadow[r] = js::f32(-js::min(1, -windDotGrad * 18) * heightScale);
            }
            // wind-aligned neighbour lists (:520-547)
            const int E = mesh.adjOffset[N];
            I32 upNb(E), dnNb(E), upOff(N + 1), dnOff(N + 1);
            F32 upWt(E), dnWt(E);
            int upCount = 0, dnCount = 0;
            for (int r = 0; r < N; r++) {
                upOff[r] = upCount; dnOff[r] = dnCount;
                if (!r_isLand[r]) continue;
                for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) {
                    const int nb = mesh.adjList[ni];
                    const double dx = (double)xyz[3 * r] - (double)xyz[3 * nb];
                    const double dy = (double)xyz[3 * r + 1] - (double)xyz[3 * nb + 1];
                    const double dz = (double)xyz[3 * r + 2] - (double)xyz[3 * nb + 2];
                    const double upDot = wX[nb] * dx + wY[nb] * dy + wZ[nb] * dz;
                    if (upDot > 0) { upNb[upCount] = nb; upWt[upCount] = js::f32(upDot); upCount++; }
                    const double dnDot = -(wX[r] * dx + wY[r] * dy + wZ[r] * dz);
                    if (dnDot > 0) { dnNb[dnCount] = nb; dnWt[dnCount] = js::f32(dnDot); dnCount++; }
                }
            }
            upOff[N] = upCount; dnOff[N] = dnCount;

            const int shadowHops = (int)js::max(8, js::round(2500 / avgEdgeKm));
            const double shadowDecay = 1 - pb_pow(0.15, 1.0 / shadowHops);
            F32 shadowField(rainShadow);
            {
                F32 a(shadowField), b(N, 0.f);
                F32 *src = &a, *dst = &b;
                for (int iter = 0; iter < shadowHops; iter++) {
                    for (int r = 0; r < N; r++) {
                        double upVal = 0, upW = 0;
                        for (int ui = upOff[r]; ui < upOff[r + 1]; ui++) {
                            const double val = (*src)[upNb[ui]];
                            if (val < 0) { upVal += val * upWt[ui]; upW += upWt[ui]; }
                        }
                        if (upW > 0) {
                            const double carried = (upVal / upW) * (1 - shadowDecay);
                            (*dst)[r] = js::f32(js::min((*src)[r], carried));
                        } else (*dst)[r] = (*src)[r];
                    }
                    std::swap(src, dst);
                }
                for (int r = 0; r < N; r++) if ((*src)[r] < shadowField[r]) shadowField[r] = (*src)[r];
            }
            const int windwardHops = (int)js::max(6, js::round(1500 / avgEdgeKm));
            const double windwardDecay = 1 - pb_pow(0.25, 1.0 / windwardHops);
            F32 windwardField(rainShadow);
            {
                F32 a(windwardField), b(N, 0.f);
                F32 *src = &a, *dst = &b;
                for (int iter = 0; iter < windwardHops; iter++) {
                    for (int r = 0; r < N; r++) {
                        double dnVal = 0, dnW = 0;
                        for (int di = dnOff[r]; di < dnOff[r + 1]; di++) {
                            const double val = (*src)[dnNb[di]];
                            if (val > 0) { dnVal += val * dnWt[di]; dnW += dnWt[di]; }
                        }
                        if (dnW > 0) {
                            const double carried = (dnVal / dnW) * (1 - windwardDecay);
                            (*dst)[r] = js::f32(js::max((*src)[r], carried));
                        } else (*dst)[r] = (*src)[r];
                    }
                    std::swap(src, dst);
                }
                for (int r = 0; r < N; r++) if ((*src)[r] > windwardField[r]) windwardField[r] = (*src)[r];
            }
            for (int r = 0; r < N; r++) rainShadow[r] = shadowField[r] < 0 ? shadowField[r] : windwardField[r];
            const int rsSmoothPasses = (int)js::max(2, js::round(150 / avgEdgeKm));
            oracle_smooth_field(mesh, rainShadow.data(), rsSmoothPasses);

            for (int r = 0; r < N; r++) {
                if (!r_isLand[r]) continue;
                const double rs = rainShadow[r];
                if (rs < -0.01) {
                    const double strength = js::min(1, -rs * 2.25);
                    precip[r] = js::f32(precip[r] * js::max(0.02, 1 - strength * 0.92));
                } else if (rs > 0.01) precip[r] = js::f32(precip[r] + rs * 1.2);
            }
            const int precipSmoothPasses = (int)js::max(1, js::round(100 / avgEdgeKm));
            oracle_smooth_field(mesh, precip.data(), precipSmoothPasses);
            f["r_precip_complex_" + name] = precip;
        }

        F32 heur[2];
        computeHeuristicPrecipitation(elev, r_elevGradE, r_elevGradN, heur);
        for (int s = 0; s < 2; s++) {
            const std::string name = s == 0 ? "summer" : "winter";
            const F32& complex = f["r_precip_complex_" + name];
            F32& blended = f["r_precip_" + name];
            blended.assign(N, 0.f);
            for (int r = 0; r < N; r++) blended[r] = js::f32(0.5 * complex[r] + 0.5 * heur[s][r]);
            const double maxPrecip = oracle_percentile(blended.data(), N, 0.95);
            for (int r = 0; r < N; r++) blended[r] = js::f32(js::min(1, blended[r] / maxPrecip));
            for (int r = 0; r < N; r++)
                if (r_isLand[r] && r_cont[r] > 0.5) {
                    const double t = smoothstep(0.5, 1.0, r_cont[r]);
                    const double cap = 1.0 - t * 0.80;
                    blended[r] = js::f32(js::min(blended[r], cap));
                }
            f["r_precip_heuristic_" + name] = heur[s];
        }
        havePrecip = true;
    }

    // js/temperature.js:19-55
    F32 diffuseOceanWarmth(const F32& r_oceanWarmth, const U8& r_isLand, const F32& plateCont, int passes) {
        F32 coastal(N, 0.f);
        for (int r = 0; r < N; r++) if (!r_isLand[r]) coastal[r] = r_oceanWarmth[r];
        F32 tmp(N, 0.f);
        for (int pass = 0; pass < passes; pass++) {
            tmp = coastal;
            for (int r = 0; r < N; r++) {
                if (plateCont[r] >= 0.95) continue;
                double sum = coastal[r];
                int count = 1;
                for (int ni = mesh.adjOffset[r]; ni < mesh.adjOffset[r + 1]; ni++) { sum += coastal[mesh.adjList[ni]]; count++; }
                tmp[r] = js::f32(sum / count);
            }
            coastal = tmp;
        }
        return coastal;
    }

    // js/temperature.js:69-237
    void computeTemperature(const float* elev, double temperatureOffset) {
        const F32 &r_lat = f["r_lat"], &r_lon = f["r_lon"], &r_cont = f["r_continentality"], &r_pcont = f["r_plateContinentality"];
        const U8& r_isLand = u["r_isLand"];
        const double T_MIN = -45, T_MAX = 45, T_RANGE = T_MAX - T_MIN;
        for (int s = 0; s < 2; s++) {
            const bool summer = s == 0;
            const std::string name = summer ? "summer" : "winter";
            const F32 &r_oceanWarmth = f["r_ocean_warmth_" + name], &r_oceanSpeed = f["r_ocean_speed_" + name];
            const F32& r_precip = f["r_precip_" + name];
            ItczLookup itczLookup(f["itczLons"], summer ? f["itczLatsSummer"] : f["itczLatsWinter"]);
            const double avgEdgeKm = (PB_PI * 6371) / std::sqrt((double)N);
            const int oceanWarmthPasses = (int)js::max(4, js::round(1400 / avgEdgeKm));
            F32 coastalWarmth = diffuseOceanWarmth(r_oceanWarmth, r_isLand, r_pcont, oceanWarmthPasses);
            F32& temp = f["r_temperature_" + name];
            temp.assign(N, 0.f);
            for (int r = 0; r < N; r++) {
                const double lat = r_lat[r], lon = r_lon[r];
                const bool isLand = r_isLand[r] != 0;
                const double el = elev[r];
                const double cont = r_cont[r];
                const double pCont = r_pcont[r];
                const double tropicalHW = 13;
                const double maxDist = 90 - tropicalHW;
                const double itczLat = itczLookup(lon);
                const double distItcz = std::fabs(lat - itczLat) / DEG;
                const double tItcz = js::max(0, distItcz - tropicalHW) / maxDist;
                const double T_itcz = 28 - 47 * pb_pow(tItcz, 1.4);
                const double flatItczLat = (summer ? 5 : -5) * DEG;
                const double distFlat = std::fabs(lat - flatItczLat) / DEG;
                const double tFlat = js::max(0, distFlat - tropicalHW) / maxDist;
                const double T_flat = 28 - 47 * pb_pow(tFlat, 1.4);
                const double absLatDeg = std::fabs(lat) / DEG;
                const double blend = smoothstep(45, 90, absLatDeg);
                double T = T_itcz * (1 - blend) + T_flat * blend;
                const double moisture = r_precip[r];
                const double lapse = 4.5 + 4.8 * (1 - moisture);
                if (isLand && el > 0) T -= lapse * elevToHeightKm(el);
                if (!isLand) {
                    const double warmth = r_oceanWarmth[r], speed = r_oceanSpeed[r];
                    T += warmth * js::min(1, speed * 2) * 16;
                } else {
                    const double cw = coastalWarmth[r];
                    if (std::fabs(cw) > 0.001) T += cw * (1 - smoothstep(0, 0.95, pCont)) * 20;
                }
                {
                    const double p = r_precip[r];
                    if (p > 0.5) { const double mod = smoothstep(0.5, 1.0, p) * 0.15; T *= (1 - mod); }
                    else if (p < 0.3) { const double amp = smoothstep(0.3, 0.0, p) * 0.15; T *= (1 + amp); }
                }
                {
                    const double distAnn = std::fabs(lat) / DEG;
                    const double tAnn = js::max(0, distAnn - tropicalHW) / maxDist;
                    const double T_annual = 28 - 47 * pb_pow(tAnn, 1.4);
                    const double T_ann_adj = isLand && el > 0 ? T_annual - lapse * elevToHeightKm(el) : T_annual;
                    const double deviation = T - T_ann_adj;
                    const double seasonalBoost = 12 * smoothstep(10, 55, distAnn) * (1 - smoothstep(75, 90, distAnn));
                    const bool isLocalSummer = summer ? (lat >= 0) : (lat < 0);
                    const double seasonSign = isLocalSummer ? 1 : -1;
                    const double boostedDeviation = deviation + seasonSign * seasonalBoost;
                    const double maritimeFactor = 0.50 + cont * 0.70;
                    T = T_ann_adj + boostedDeviation * maritimeFactor;
                }
                T += temperatureOffset;
                temp[r] = js::f32(T);
            }
            oracle_smooth_field(mesh, temp.data(), 1);
            for (int r = 0; r < N; r++) temp[r] = js::f32(js::max(0, js::min(1, (temp[r] - T_MIN) / T_RANGE)));
        }
        haveTemp = true;
    }

    // js/koppen.js:67-288.  Class ids follow KOPPEN_CLASSES (:19-51).
    void classifyKoppen(const float* elev) {
        enum { Ocean, Af, Am, Aw, BWh, BWk, BSh, BSk, Cfa, Cfb, Cfc, Csa, Csb, Csc, Cwa, Cwb, Cwc, Dfa, Dfb, Dfc, Dfd,
               Dsa, Dsb, Dsc, Dsd, Dwa, Dwb, Dwc, Dwd, ET, EF };
        const F32 &tSummer = f["r_temperature_summer"], &tWinter = f["r_temperature_winter"];
        const F32 &pSummer = f["r_precip_summer"], &pWinter = f["r_precip_winter"];
        U8& k = u["r_koppen"];
        k.assign(N, 0);
        for (int r = 0; r < N; r++) {
            if (elev[r] <= 0) { k[r] = Ocean; continue; }
            const double Ts = -45 + js::max(0, js::min(1, tSummer[r])) * 90;
            const double Tw = -45 + js::max(0, js::min(1, tWinter[r])) * 90;
            const double Thot = js::max(Ts, Tw), Tcold = js::min(Ts, Tw);
            const double Tann = (Ts + Tw) / 2;
            const double Tshoulder = Thot - (Thot - Tcold) * (2.0 / 6);
            const bool localSummerIsSim = Ts >= Tw;
            const double Ps = js::max(0, pSummer[r]) * 1000, Pw = js::max(0, pWinter[r]) * 1000;
            const double Pann = Ps + Pw;
            const double PsummerLocal = localSummerIsSim ? Ps : Pw;
            const double PwinterLocal = localSummerIsSim ? Pw : Ps;
            const double PsMonthLocal = PsummerLocal / 6, PwMonthLocal = PwinterLocal / 6;
            const double Pdry = js::min(PsMonthLocal, PwMonthLocal);
            char band;
            if (Thot < 0) { k[r] = EF; continue; }
            else if (Thot < 10) { k[r] = ET; continue; }
            else if (Tcold >= 18) band = 'A';
            else if (Tcold >= 0) band = 'C';
            else band = 'D';
            double Pthresh;
            const double summerFrac = Pann > 0 ? PsummerLocal / Pann : 0.5;
            if (summerFrac >= 0.7) Pthresh = 20 * Tann + 280;
            else if (summerFrac <= 0.3) Pthresh = 20 * Tann;
            else Pthresh = 20 * Tann + 140;
            Pthresh = js::max(0, Pthresh);
            if (Pann < Pthresh) {
                const bool isHot = Tann >= 18;
                if (Pann < Pthresh * 0.5) k[r] = isHot ? BWh : BWk;
                else k[r] = isHot ? BSh : BSk;
                continue;
            }
            int pat;   // 0 f, 1 s, 2 w
            const bool localSummerDrier = PsummerLocal < PwinterLocal;
            if (localSummerDrier && PsMonthLocal < 50 && PsMonthLocal < PwMonthLocal / 2) pat = 1;
            else if (!localSummerDrier && PwMonthLocal < PsMonthLocal / 10) pat = 2;
            else pat = 0;
            int tl;    // 0 a, 1 b, 2 c, 3 d
            if (Thot >= 22) tl = 0;
            else if (Tshoulder >= 10) tl = 1;
            else if (Tcold >= -38) tl = 2;
            else tl = 3;
            if (band == 'A') {
                if (Pdry >= 60) k[r] = Af;
                else if (Pann >= 25 * (100 - Pdry)) k[r] = Am;
                else k[r] = Aw;
                continue;
            }
            if (band == 'C') {
                // codes 'C'+pattern+letter that exist: Cf[abc], Cs[abc], Cw[abc]; anything with 'd' → Cfb
                static const int C[3][3] = {{Cfa, Cfb, Cfc}, {Csa, Csb, Csc}, {Cwa, Cwb, Cwc}};
                k[r] = tl < 3 ? C[pat][tl] : Cfb;
                continue;
            }
            static const int D[3][4] = {{Dfa, Dfb, Dfc, Dfd}, {Dsa, Dsb, Dsc, Dsd}, {Dwa, Dwb, Dwc, Dwd}};
            k[r] = D[pat][tl];
        }
    }
};

extern "C" {

void* orc_climate_create(int N, const int32_t* off, const int32_t* adj, const float* xyz) {
    return new OracleClimate(OMesh{N, off, adj}, xyz);
}
void orc_climate_destroy(void* h) { delete (OracleClimate*)h; }
void orc_climate_wind(void* h, const float* elev, const int32_t* plateIsOcean, int nIds, const int32_t* r_plate,
                      double noiseSeed, double axialTilt) {
    ((OracleClimate*)h)->computeWind(elev, plateIsOcean, nIds, r_plate, noiseSeed, axialTilt);
}
void orc_climate_ocean(void* h, const float* elev) { ((OracleClimate*)h)->computeOceanCurrents(elev); }
void orc_climate_precip(void* h, const float* elev, double precipitationOffset, double landCoverage) {
    ((OracleClimate*)h)->computePrecipitation(elev, precipitationOffset, landCoverage);
}
void orc_climate_temperature(void* h, const float* elev, double temperatureOffset) {
    ((OracleClimate*)h)->computeTemperature(elev, temperatureOffset);
}
void orc_climate_koppen(void* h, const float* elev) { ((OracleClimate*)h)->classifyKoppen(elev); }
// kind: 0 f32, 1 i32, 2 u8.  Returns the element count, or -1 when the field does not exist.
int64_t orc_climate_get(void* h, const char* name, int kind, void* out, int64_t cap) {
    OracleClimate* c = (OracleClimate*)h;
    const void* p = nullptr;
    int64_t n = 0, es = 4;
    if (kind == 0) { auto it = c->f.find(name); if (it == c->f.end()) return -1; p = it->second.data(); n = (int64_t)it->second.size(); }
    else if (kind == 1) { auto it = c->i.find(name); if (it == c->i.end()) return -1; p = it->second.data(); n = (int64_t)it->second.size(); }
    else { auto it = c->u.find(name); if (it == c->u.end()) return -1; p = it->second.data(); n = (int64_t)it->second.size(); es = 1; }
    if (out && cap >= n) std::memcpy(out, p, (size_t)(n * es));
    return n;
}
// stand-alone passes for per-pass parity tests
void orc_compute_gradients(int N, const int32_t* off, const int32_t* adj, const float* xyz, const float* field,
                           const float* frames6, float* gradE, float* gradN) {
    OracleClimate c(OMesh{N, off, adj}, xyz);
    const char* names[6] = {"r_eastX", "r_eastY", "r_eastZ", "r_northX", "r_northY", "r_northZ"};
    for (int k = 0; k < 6; k++) c.f[names[k]].assign(frames6 + (size_t)k * N, frames6 + (size_t)(k + 1) * N);
    F32 P(field, field + N), ge(N), gn(N);
    c.computeGradients(P, ge, gn);
    std::memcpy(gradE, ge.data(), sizeof(float) * N);
    std::memcpy(gradN, gn.data(), sizeof(float) * N);
}
void orc_smooth_masked(int N, const int32_t* off, const int32_t* adj, float* field, const uint8_t* mask, int passes) {
    F32 fld(field, field + N);
    U8 m(mask, mask + N);
    smoothOcean(OMesh{N, off, adj}, fld, m, passes);
    std::memcpy(field, fld.data(), sizeof(float) * N);
}

}  // extern "C"
