// ORACLE — TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load anything under oracle/.
//
// JS-semantics helpers used by the CPU restatement of the reference
// (raguilar011095/planet_heightmap_generation, js/*.js).  Parity status: the
// reference ships no tests or golden vectors and no JS runtime exists in this
// image.  The oracle is pinned (a) against the derived known-answer vectors in
// SURVEY.md §8(c) (tests/test_oracle_kat.py) and (b) against the reference's own
// source executed under the minimal evaluator tests/golden/minijs.py: whole worker
// replies committed under tests/golden/reference_*.npz and reproduced bit for bit
// (tests/test_zz_reference_vectors.py; DESIGN.md §3 lists what that pin inherits).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include "../include/pb_detmath.h"

namespace js {

// Math.round: floor(x + 0.5)  (not C round(): JS rounds .5 toward +inf)
inline double round(double x) { return std::floor(x + 0.5); }

// ToUint32 on a finite double (values here are exact integers, possibly > 2^53 after rounding)
inline uint32_t to_uint32(double d) {
    double m = std::fmod(std::trunc(d), 4294967296.0);
    if (m < 0) m += 4294967296.0;
    return (uint32_t)m;
}
inline int32_t to_int32(double d) { return (int32_t)to_uint32(d); }

// a || b on numbers: b when a is 0, -0 or NaN
inline double or_default(double a, double b) { return (a == 0.0 || a != a) ? b : a; }

inline float f32(double x) { return (float)x; }

inline double max(double a, double b) { // Math.max semantics incl. NaN and +0 > -0
    if (a != a || b != b) return NAN;
    if (a == 0.0 && b == 0.0) return std::signbit(a) ? b : a;
    return a > b ? a : b;
}
inline double min(double a, double b) {
    if (a != a || b != b) return NAN;
    if (a == 0.0 && b == 0.0) return std::signbit(a) ? a : b;
    return a < b ? a : b;
}

} // namespace js

struct OMesh {
    int32_t N;                 // numRegions (cells + pole)
    const int32_t* adjOffset;  // [N+1]
    const int32_t* adjList;    // [E]
};
