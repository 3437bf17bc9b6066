/*
 * pb_detmath.h — deterministic transcendental functions for host and device.
 *
 * ECMAScript defines Math.pow/exp/log/sin/cos/asin/atan2/tanh as
 * "implementation-approximated"; the reference (V8) uses an fdlibm-derived
 * libm.  glibc and CUDA libdevice each round a little differently in the last
 * ulp, which is enough to flip an f32 store and — through the erosion stack's
 * sort/receiver decisions — diverge chaotically.  These routines use ONLY
 * + - * / sqrt floor and bit casts (all exactly rounded in IEEE-754), so the
 * CPU oracle (gcc -ffp-contract=off) and the CUDA kernels (nvcc -fmad=false)
 * produce bit-identical doubles.  Accuracy is a few ulp(double) of the true
 * value, i.e. indistinguishable from V8's result after an f32 store except with
 * probability ~1e-8 per evaluation.
 *
 * Special case kept from fdlibm (which V8's pow follows): pow(x, 0.5) == sqrt(x)
 * for x >= 0, so the stream-power exponent m = 0.5 is exactly rounded.
 *
 * Compile with FMA contraction disabled on both sides.
 */
#ifndef PB_DETMATH_H
#define PB_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD static inline
#endif

#define PB_PI 3.141592653589793
#define PB_LN2_HI 6.93147180369123816490e-01 /* 0x3fe62e42fee00000 */
#define PB_LN2_LO 1.90821492927058770002e-10 /* 0x3dea39ef35793c76 */
#define PB_INV_LN2 1.44269504088896338700e+00

PB_HD double pb_bits2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
PB_HD uint64_t pb_d2bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}

/* 2^k for integer k in the normal range, exact. */
PB_HD double pb_exp2i(int k) {
    if (k > 1023) return pb_bits2d(0x7ff0000000000000ULL);
    if (k < -1022) { /* subnormal or zero: two-step scaling */
        if (k < -1074) return 0.0;
        return pb_bits2d((uint64_t)(k + 1023 + 200) << 52) * pb_bits2d((uint64_t)(1023 - 200) << 52);
    }
    return pb_bits2d((uint64_t)(k + 1023) << 52);
}

/* exp(x): k = round(x/ln2); r = x - k ln2 (hi/lo); Taylor to r^17; scale. */
PB_HD double pb_exp_hl(double xh, double xl) {
    if (xh != xh) return xh;
    if (xh > 709.782712893384) return pb_bits2d(0x7ff0000000000000ULL);
    if (xh < -745.2) return 0.0;
    double kf = floor(xh * PB_INV_LN2 + 0.5);
    int k = (int)kf;
    double r = (xh - kf * PB_LN2_HI) - kf * PB_LN2_LO + xl;
    /* Horner on exp(r) - 1 - r, |r| <= 0.3466 */
    double p = 1.0 / 355687428096000.0; /* 1/17! */
    p = p * r + 1.0 / 20922789888000.0;
    p = p * r + 1.0 / 1307674368000.0;
    p = p * r + 1.0 / 87178291200.0;
    p = p * r + 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    double e = 1.0 + (r + r * r * p);
    if (k > 1000) return e * pb_exp2i(k - 500) * pb_exp2i(500);
    if (k < -1000) return e * pb_exp2i(k + 500) * pb_exp2i(-500);
    return e * pb_exp2i(k);
}
PB_HD double pb_exp(double x) { return pb_exp_hl(x, 0.0); }

/* log(x) as hi + lo (lo is a small correction), x > 0 finite normal/subnormal. */
PB_HD void pb_log_hl(double x, double* hi, double* lo) {
    int eadj = 0;
    uint64_t u = pb_d2bits(x);
    if ((u >> 52) == 0) { x = x * 18014398509481984.0; u = pb_d2bits(x); eadj = -54; } /* subnormal */
    int e = (int)((u >> 52) & 0x7ff) - 1023 + eadj;
    double m = pb_bits2d((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL); /* [1,2) */
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    /* ln m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716 */
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double q = 1.0 / 27.0;
    q = q * z + 1.0 / 25.0;
    q = q * z + 1.0 / 23.0;
    q = q * z + 1.0 / 21.0;
    q = q * z + 1.0 / 19.0;
    q = q * z + 1.0 / 17.0;
    q = q * z + 1.0 / 15.0;
    q = q * z + 1.0 / 13.0;
    q = q * z + 1.0 / 11.0;
    q = q * z + 1.0 / 9.0;
    q = q * z + 1.0 / 7.0;
    q = q * z + 1.0 / 5.0;
    q = q * z + 1.0 / 3.0;
    double lm = 2.0 * (s + s * z * q);
    double ef = (double)e;
    *hi = ef * PB_LN2_HI; /* exact: PB_LN2_HI has 32 significant bits, |e| < 2^11 */
    *lo = ef * PB_LN2_LO + lm;
}
PB_HD double pb_log(double x) {
    if (x != x || x < 0) return pb_bits2d(0x7ff8000000000000ULL);
    if (x == 0) return -pb_bits2d(0x7ff0000000000000ULL);
    if (x > 1.7976931348623157e308) return x;
    double h, l; pb_log_hl(x, &h, &l); return h + l;
}

/* Veltkamp/Dekker exact product a*b = p + e without FMA. */
PB_HD void pb_two_prod(double a, double b, double* p, double* e) {
    const double split = 134217729.0; /* 2^27 + 1 */
    double ca = split * a, ah = ca - (ca - a), al = a - ah;
    double cb = split * b, bh = cb - (cb - b), bl = b - bh;
    *p = a * b;
    *e = ((ah * bh - *p) + ah * bl + al * bh) + al * bl;
}

/* pow for x >= 0, finite y.  Matches fdlibm special cases that the hot path can hit. */
PB_HD double pb_pow(double x, double y) {
    if (y == 0.0) return 1.0;
    if (x != x || y != y) return pb_bits2d(0x7ff8000000000000ULL);
    if (y == 1.0) return x;
    if (y == 0.5 && x >= 0.0) return sqrt(x);
    if (x == 0.0) return y > 0 ? 0.0 : pb_bits2d(0x7ff0000000000000ULL);
    if (x < 0.0) return pb_bits2d(0x7ff8000000000000ULL); /* hot path never raises negatives to fractions */
    if (x == 1.0) return 1.0;
    if (x > 1.7976931348623157e308) return y > 0 ? x : 0.0;
    double lh, ll; pb_log_hl(x, &lh, &ll);
    /* y * (lh + ll) as hi + lo */
    double ph, pe; pb_two_prod(y, lh, &ph, &pe);
    double pl = pe + y * ll;
    double vh = ph + pl;
    double vl = pl - (vh - ph);
    return pb_exp_hl(vh, vl);
}

/* atan for any finite t: two half-angle reductions then odd Taylor series. */
PB_HD double pb_atan(double t) {
    if (t != t) return t;
    double sign = 1.0;
    if (t < 0) { t = -t; sign = -1.0; }
    double add = 0.0;
    if (t > 1.0) { t = -1.0 / t; add = PB_PI * 0.5; } /* atan t = pi/2 - atan(1/t) */
    /* atan t = 2 atan( t / (1 + sqrt(1+t^2)) ), applied twice → |u| <= tan(pi/16) */
    double u = t / (1.0 + sqrt(1.0 + t * t));
    u = u / (1.0 + sqrt(1.0 + u * u));
    double z = u * u;
    double p = -1.0 / 39.0;
    p = p * z + 1.0 / 37.0;
    p = p * z - 1.0 / 35.0;
    p = p * z + 1.0 / 33.0;
    p = p * z - 1.0 / 31.0;
    p = p * z + 1.0 / 29.0;
    p = p * z - 1.0 / 27.0;
    p = p * z + 1.0 / 25.0;
    p = p * z - 1.0 / 23.0;
    p = p * z + 1.0 / 21.0;
    p = p * z - 1.0 / 19.0;
    p = p * z + 1.0 / 17.0;
    p = p * z - 1.0 / 15.0;
    p = p * z + 1.0 / 13.0;
    p = p * z - 1.0 / 11.0;
    p = p * z + 1.0 / 9.0;
    p = p * z - 1.0 / 7.0;
    p = p * z + 1.0 / 5.0;
    p = p * z - 1.0 / 3.0;
    double a = 4.0 * (u + u * z * p);
    return sign * (add + a);
}

PB_HD double pb_asin(double x) {
    if (x != x || x > 1.0 || x < -1.0) return pb_bits2d(0x7ff8000000000000ULL);
    if (x == 1.0) return PB_PI * 0.5;
    if (x == -1.0) return -PB_PI * 0.5;
    /* asin x = 2 atan( x / (1 + sqrt(1 - x^2)) ); (1-x)(1+x) avoids cancellation */
    return 2.0 * pb_atan(x / (1.0 + sqrt((1.0 - x) * (1.0 + x))));
}

PB_HD double pb_atan2(double y, double x) {
    if (x != x || y != y) return pb_bits2d(0x7ff8000000000000ULL);
    if (x == 0.0 && y == 0.0) {
        /* signed-zero cases of IEEE atan2 */
        int xneg = (int)(pb_d2bits(x) >> 63), yneg = (int)(pb_d2bits(y) >> 63);
        double r = xneg ? PB_PI : 0.0;
        return yneg ? -r : r;
    }
    if (x == 0.0) return y > 0 ? PB_PI * 0.5 : -PB_PI * 0.5;
    double a = pb_atan(y / x);
    if (x > 0) return a;
    return (y >= 0 && !(pb_d2bits(y) >> 63)) ? a + PB_PI : a - PB_PI;
}

/* sin/cos: reduce by pi/2 with a 3-part constant (|x| up to ~1e5 keeps full accuracy). */
PB_HD void pb_sincos(double x, double* s, double* c) {
    if (x != x || x > 1.7976931348623157e308 || x < -1.7976931348623157e308) {
        *s = *c = pb_bits2d(0x7ff8000000000000ULL); return;
    }
    const double P1 = 1.57079632673412561417e+00; /* first 33 bits of pi/2 */
    const double P2 = 6.07710050630396597660e-11; /* next 33 bits */
    const double P3 = 2.02226624871116645580e-21; /* remainder */
    double kf = floor(x * 0.63661977236758134308 + 0.5);
    double r = ((x - kf * P1) - kf * P2) - kf * P3;
    double z = r * r;
    /* sin r, |r| <= pi/4 */
    double ps = 1.0 / 121645100408832000.0; /* 1/19! */
    ps = ps * z - 1.0 / 355687428096000.0;
    ps = ps * z + 1.0 / 1307674368000.0;
    ps = ps * z - 1.0 / 6227020800.0;
    ps = ps * z + 1.0 / 39916800.0;
    ps = ps * z - 1.0 / 362880.0;
    ps = ps * z + 1.0 / 5040.0;
    ps = ps * z - 1.0 / 120.0;
    ps = ps * z + 1.0 / 6.0;
    double sr = r - r * z * ps;
    double pc = 1.0 / 2432902008176640000.0; /* 1/20! */
    pc = pc * z - 1.0 / 6402373705728000.0;
    pc = pc * z + 1.0 / 20922789888000.0;
    pc = pc * z - 1.0 / 87178291200.0;
    pc = pc * z + 1.0 / 479001600.0;
    pc = pc * z - 1.0 / 3628800.0;
    pc = pc * z + 1.0 / 40320.0;
    pc = pc * z - 1.0 / 720.0;
    pc = pc * z + 1.0 / 24.0;
    double cr = 1.0 - (0.5 * z - z * z * pc);
    /* quadrant: kf mod 4 */
    double q = kf - 4.0 * floor(kf * 0.25);
    int qi = (int)q;
    if (qi == 0) { *s = sr; *c = cr; }
    else if (qi == 1) { *s = cr; *c = -sr; }
    else if (qi == 2) { *s = -sr; *c = -cr; }
    else { *s = -cr; *c = sr; }
}
PB_HD double pb_sin(double x) { double s, c; pb_sincos(x, &s, &c); return s; }
PB_HD double pb_cos(double x) { double s, c; pb_sincos(x, &s, &c); return c; }

/* exp(r) - 1 for |r| <= 0.35 without cancellation (same Horner polynomial as pb_exp_hl). */
PB_HD double pb_expm1_small(double r) {
    double p = 1.0 / 355687428096000.0;
    p = p * r + 1.0 / 20922789888000.0;
    p = p * r + 1.0 / 1307674368000.0;
    p = p * r + 1.0 / 87178291200.0;
    p = p * r + 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    return r + r * r * p;
}

PB_HD double pb_tanh(double x) {
    if (x != x) return x;
    double ax = x < 0 ? -x : x;
    double r;
    if (ax > 22.0) r = 1.0;
    else if (ax < 0.17) { double t = pb_expm1_small(2.0 * ax); r = t / (t + 2.0); }
    else { double e = pb_exp(2.0 * ax); r = 1.0 - 2.0 / (e + 1.0); }
    return x < 0 ? -r : r;
}

PB_HD double pb_acos(double x) { return PB_PI * 0.5 - pb_asin(x); }

#endif /* PB_DETMATH_H */
