// ORACLE — TEST INFRASTRUCTURE ONLY (see js_semantics.h).
// Restates js/climate-util.js.
#pragma once
#include <algorithm>
#include "js_semantics.h"

// js/climate-util.js:5-25
inline void oracle_smooth_field(const OMesh& mesh, float* field, int passes) {
    const int N = mesh.N;
    std::vector<float> tmp(N, 0.f);
    float* src = field; float* dst = tmp.data();
    for (int pass = 0; pass < passes; pass++) {
        for (int r = 0; r < N; r++) {
            double sum = src[r];
            int count = 1;
            const int end = mesh.adjOffset[r + 1];
            for (int ni = mesh.adjOffset[r]; ni < end; ni++) { sum += src[mesh.adjList[ni]]; count++; }
            dst[r] = js::f32(sum / count);
        }
        std::swap(src, dst);
    }
    if (src != field) std::copy(src, src + N, field);
}

// js/climate-util.js:103-110 — Floyd–Rivest selects the value at sorted index floor(n*p);
// any exact selection returns the same value (NaN-free inputs).
inline double oracle_percentile(const float* arr, int n, double p) {
    if (n == 0) return 1;
    std::vector<float> work(arr, arr + n);
    int k = (int)std::floor(n * p);
    std::nth_element(work.begin(), work.begin() + k, work.end());
    return js::or_default(work[k], 1);
}
