// pb_prims.h — global primitives used between the per-cell passes: ordered stream compaction and
// stable radix sort of (key, cell id) pairs (kernel family K10).  These are plumbing, not the hot
// path: the CUDA build calls CUB's device-wide primitives (both are stable, which the reference's
// persistent `landCells` array + stable Array.prototype.sort semantics require — SURVEY.md A.1).
#pragma once
#include "pb_platform.h"
#if PB_CUDA
#include <cub/cub.cuh>
#endif

namespace pb {

struct Prims {
    DevBuf<uint8_t> temp;
    DevBuf<uint32_t> kAlt;
    DevBuf<int> vAlt;
    DevBuf<int> iota;
    size_t iotaFilled = 0;

    struct IotaK { int* p; PB_DEV void operator()(int i) const { p[i] = i; } };

    const int* ensure_iota(const Exec& ex, int n) {
        if ((size_t)n > iotaFilled) {
            iota.ensure(n);
            ex.for_each(n, IotaK{iota.p});
            iotaFilled = n;
        }
        return iota.p;
    }

    // out[0..count) = ascending indices i in [0,n) with flag[i] != 0; *dCount = count (device int)
    void compact_flagged(const Exec& ex, const uint8_t* flag, int n, int* out, int* dCount) {
        if (n <= 0) { dev_memset(dCount, 0, sizeof(int), ex.stream); return; }
        const int* idx = nullptr;
#if PB_CUDA
        idx = ensure_iota(ex, n);
#endif
        (void)idx;
        launch_stats().launches++;
        ProfScope ps(ex.prof, "cub::DeviceSelect::Flagged", ex.stream);
#if PB_CUDA
        size_t bytes = 0;
        PB_CUDA_CHECK(cub::DeviceSelect::Flagged(nullptr, bytes, idx, flag, out, dCount, n, ex.stream));
        temp.ensure(bytes);
        PB_CUDA_CHECK(cub::DeviceSelect::Flagged(temp.p, bytes, idx, flag, out, dCount, n, ex.stream));
#else
        int c = 0;
        for (int i = 0; i < n; i++) if (flag[i]) out[c++] = i;
        *dCount = c;
#endif
    }

    // ascending sort of unsigned 32-bit keys in place (selection / percentile)
    void sort_keys(const Exec& ex, uint32_t* keys, int n) {
        if (n <= 1) return;
        launch_stats().launches++;
        ProfScope ps(ex.prof, "cub::DeviceRadixSort::SortKeys", ex.stream);
#if PB_CUDA
        kAlt.ensure(n);
        cub::DoubleBuffer<uint32_t> dk(keys, kAlt.p);
        size_t bytes = 0;
        PB_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, bytes, dk, n, 0, 32, ex.stream));
        temp.ensure(bytes);
        PB_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(temp.p, bytes, dk, n, 0, 32, ex.stream));
        if (dk.Current() != keys) dev_copy(keys, dk.Current(), (size_t)n * sizeof(uint32_t), 2, ex.stream);
#else
        std::sort(keys, keys + n);
#endif
    }

    // stable sort of (key, val) pairs in place; descending or ascending on unsigned 32-bit keys
    void sort_pairs(const Exec& ex, uint32_t* keys, int* vals, int n, bool descending, int endBit = 32) {
        if (n <= 1) return;
        launch_stats().launches++;
        ProfScope ps(ex.prof, "cub::DeviceRadixSort::SortPairs", ex.stream);
#if PB_CUDA
        kAlt.ensure(n);
        vAlt.ensure(n);
        cub::DoubleBuffer<uint32_t> dk(keys, kAlt.p);
        cub::DoubleBuffer<int> dv(vals, vAlt.p);
        size_t bytes = 0;
        if (descending) PB_CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, dk, dv, n, 0, endBit, ex.stream));
        else PB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, n, 0, endBit, ex.stream));
        temp.ensure(bytes);
        if (descending) PB_CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(temp.p, bytes, dk, dv, n, 0, endBit, ex.stream));
        else PB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(temp.p, bytes, dk, dv, n, 0, endBit, ex.stream));
        if (dk.Current() != keys) dev_copy(keys, dk.Current(), (size_t)n * sizeof(uint32_t), 2, ex.stream);
        if (dv.Current() != vals) dev_copy(vals, dv.Current(), (size_t)n * sizeof(int), 2, ex.stream);
#else
        std::vector<int> perm(n);
        for (int i = 0; i < n; i++) perm[i] = i;
        if (descending) std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return keys[a] > keys[b]; });
        else std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return keys[a] < keys[b]; });
        std::vector<uint32_t> k2(n);
        std::vector<int> v2(n);
        for (int i = 0; i < n; i++) { k2[i] = keys[perm[i]]; v2[i] = vals[perm[i]]; }
        memcpy(keys, k2.data(), (size_t)n * sizeof(uint32_t));
        memcpy(vals, v2.data(), (size_t)n * sizeof(int));
#endif
    }
};

// order-preserving image of an f32 in uint32 (−0 folded onto +0: the reference's comparator
// `b − a` treats them as equal)
PB_DEV uint32_t f32_sort_key(float f) {
    if (f == 0.0f) f = 0.0f;
    uint32_t u;
#if PB_CUDA
    u = __float_as_uint(f);
#else
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

}  // namespace pb
