// pb_platform.h — execution layer shared by every kernel of the engine.
//
// Product build:  nvcc -gencode arch=compute_100a,code=sm_100a  (PB_CUDA == 1).  Every per-cell
// functor below becomes a __global__ grid-stride kernel; ordered (class S) functors become
// ticketed sync-free dataflow kernels.
//
// PB_EMUL build (tests/emul only, g++): the same functors are run by a sequential loop so the
// host orchestration and the per-cell logic can be checked against the oracle on a box without a
// GPU.  It is test infrastructure — the python package never loads it (see _lib.py).
#pragma once
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <stdexcept>
#include <algorithm>
#include <atomic>
#include <typeinfo>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <functional>
#include <cxxabi.h>

#if defined(__CUDACC__) && !defined(PB_EMUL)
#define PB_CUDA 1
#include <cuda_runtime.h>
#define PB_DEV __device__ __forceinline__
#define PB_HDEV __host__ __device__ __forceinline__
#define PB_GLOBAL __global__
#else
#define PB_CUDA 0
#define PB_DEV inline
#define PB_HDEV inline
typedef void* cudaStream_t;
#endif

#include "../../include/pb_detmath.h"

namespace pb {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#if PB_CUDA
#define PB_CUDA_CHECK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            throw pb::Error(std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
    } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// memory
// ---------------------------------------------------------------------------------------------
// Device allocations go through a small per-process cache: freed blocks are kept (by device and size) and handed out again.
// Every generate call builds and drops a coarse mesh, staging arrays and scratch buffers of the same sizes; without the cache
// that is ≈ 60 cudaMalloc / cudaFree calls per planet, each a device-wide synchronisation inside the driver.  Blocks above
// the cache budget (PB_ALLOC_CACHE_MB, default 8 GiB per process) are returned to the driver.  A block is only handed back to
// the host thread that freed it: each planet in flight has its own thread and stream, and a block dropped while that stream
// still has kernels in flight must not be written from another stream (cudaFree used to provide that by synchronising).
struct AllocCache {
    std::mutex mu;
    typedef std::tuple<int, size_t, size_t> Key;                  // (device, bytes, freeing thread)
    std::multimap<Key, void*> freeBlocks;
    std::map<void*, std::pair<int, size_t>> live;
    static size_t me() { return std::hash<std::thread::id>()(std::this_thread::get_id()); }
    size_t cached = 0, budget;
    AllocCache() { const char* e = getenv("PB_ALLOC_CACHE_MB"); budget = (size_t)(e ? atoll(e) : 8192) << 20; }
    static AllocCache& get() { static AllocCache* c = new AllocCache(); return *c; }     // never destroyed: outlives every context
};
inline void* dev_alloc(size_t bytes) {
    if (bytes == 0) bytes = 16;
    bytes = (bytes + 255) & ~(size_t)255;
    AllocCache& c = AllocCache::get();
    int dev = 0;
#if PB_CUDA
    PB_CUDA_CHECK(cudaGetDevice(&dev));
#endif
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.freeBlocks.find(AllocCache::Key{dev, bytes, AllocCache::me()});
        if (it != c.freeBlocks.end()) {
            void* p = it->second;
            c.freeBlocks.erase(it);
            c.cached -= bytes;
            c.live[p] = {dev, bytes};
            return p;
        }
    }
    void* p = nullptr;
#if PB_CUDA
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {          // out of memory: give the cached blocks back and retry once
        cudaGetLastError();
        std::vector<void*> drop;
        { std::lock_guard<std::mutex> lk(c.mu); for (auto& kv : c.freeBlocks) drop.push_back(kv.second); c.freeBlocks.clear(); c.cached = 0; }
        cudaDeviceSynchronize();
        for (void* q : drop) cudaFree(q);
        PB_CUDA_CHECK(cudaMalloc(&p, bytes));
    }
#else
    p = malloc(bytes);
    if (!p) throw Error("malloc failed");
#endif
    std::lock_guard<std::mutex> lk(c.mu);
    c.live[p] = {dev, bytes};
    return p;
}
inline void dev_free(void* p) {
    if (!p) return;
    AllocCache& c = AllocCache::get();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.live.find(p);
        if (it != c.live.end()) {
            const std::pair<int, size_t> key = it->second;
            c.live.erase(it);
            if (c.cached + key.second <= c.budget) { c.freeBlocks.insert({AllocCache::Key{key.first, key.second, AllocCache::me()}, p}); c.cached += key.second; return; }
        }
    }
#if PB_CUDA
    cudaFree(p);
#else
    free(p);
#endif
}
inline void dev_memset(void* p, int v, size_t bytes, cudaStream_t s) {
#if PB_CUDA
    PB_CUDA_CHECK(cudaMemsetAsync(p, v, bytes, s));
#else
    memset(p, v, bytes);
#endif
}
// kind: 0 h2d, 1 d2h, 2 d2d, 3 h2h
inline void dev_copy(void* dst, const void* src, size_t bytes, int kind, cudaStream_t s) {
    if (bytes == 0) return;
    if (kind == 3) { memmove(dst, src, bytes); return; }
#if PB_CUDA
    cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    PB_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, k, s));
#else
    memmove(dst, src, bytes);
#endif
}
inline void stream_sync(cudaStream_t s) {
#if PB_CUDA
    PB_CUDA_CHECK(cudaStreamSynchronize(s));
#endif
}

// Typed device buffer that grows on demand and is reused between calls (no per-call cudaMalloc
// once a context is warm).
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { dev_free(p); }
    T* ensure(size_t n) {
        if (n > cap) {
            dev_free(p);
            p = nullptr;
            cap = 0;
            p = (T*)dev_alloc(n * sizeof(T));
            cap = n;
        }
        return p;
    }
    operator T*() const { return p; }
};

// Page-locked host buffer for the arrays of the host-serial stages (flood pass 1, randomized fills): they
// cross PCIe at full rate and asynchronously to the host thread.
template <class T>
struct PinnedBuf {
    T* p = nullptr;
    size_t cap = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() { release(); }
    void release() {
        if (!p) return;
#if PB_CUDA
        cudaFreeHost(p);
#else
        free(p);
#endif
        p = nullptr; cap = 0;
    }
    T* ensure(size_t n) {
        if (n > cap) {
            release();
            if (n == 0) n = 1;
#if PB_CUDA
            void* q = nullptr;
            PB_CUDA_CHECK(cudaMallocHost(&q, n * sizeof(T)));
            p = (T*)q;
#else
            p = (T*)malloc(n * sizeof(T));
            if (!p) throw Error("malloc failed");
#endif
            cap = n;
        }
        return p;
    }
    T* data() const { return p; }
    T& operator[](size_t i) const { return p[i]; }
};

// ---------------------------------------------------------------------------------------------
// device-side primitives
// ---------------------------------------------------------------------------------------------
#if PB_CUDA
PB_DEV int atomic_add(int* p, int v) { return atomicAdd(p, v); }
PB_DEV unsigned long long atomic_max64(unsigned long long* p, unsigned long long v) { return atomicMax(p, v); }
PB_DEV int atomic_min(int* p, int v) { return atomicMin(p, v); }
PB_DEV int atomic_max(int* p, int v) { return atomicMax(p, v); }
PB_DEV int atomic_cas(int* p, int cmp, int v) { return atomicCAS(p, cmp, v); }
PB_DEV void fence() { __threadfence(); }
PB_DEV int ld_volatile(const int* p) { return *(const volatile int*)p; }
// L2-coherent loads/stores for data exchanged between CTAs inside one ordered kernel
PB_DEV float ld_cg(const float* p) { return __ldcg(p); }
PB_DEV int ld_cg(const int* p) { return __ldcg(p); }
PB_DEV void st_cg(float* p, float v) { __stcg(p, v); }
PB_DEV void st_cg(int* p, int v) { __stcg(p, v); }
// 64-bit (value, sequence) words: one naturally aligned access carries both, so a consumer that polls
// the word needs no fence to pair the value with its sequence number
PB_DEV unsigned long long ld_word(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
PB_DEV void st_word(unsigned long long* p, unsigned long long v) { __stcg(p, v); }
PB_DEV float word_value(unsigned long long w) { return __uint_as_float((unsigned)(w & 0xffffffffull)); }
PB_DEV unsigned long long make_word(float v, int seq) { return ((unsigned long long)(unsigned)seq << 32) | __float_as_uint(v); }
#else
PB_DEV int atomic_add(int* p, int v) { int o = *p; *p = o + v; return o; }
PB_DEV unsigned long long atomic_max64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
PB_DEV int atomic_min(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
PB_DEV int atomic_max(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
PB_DEV int atomic_cas(int* p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
PB_DEV void fence() {}
PB_DEV int ld_volatile(const int* p) { return *p; }
PB_DEV float ld_cg(const float* p) { return *p; }
PB_DEV int ld_cg(const int* p) { return *p; }
PB_DEV void st_cg(float* p, float v) { *p = v; }
PB_DEV void st_cg(int* p, int v) { *p = v; }
PB_DEV unsigned long long ld_word(const unsigned long long* p) { return *p; }
PB_DEV void st_word(unsigned long long* p, unsigned long long v) { *p = v; }
PB_DEV float word_value(unsigned long long w) { unsigned u = (unsigned)(w & 0xffffffffull); float f; memcpy(&f, &u, 4); return f; }
PB_DEV unsigned long long make_word(float v, int seq) { unsigned u; memcpy(&u, &v, 4); return ((unsigned long long)(unsigned)seq << 32) | u; }
#endif
PB_DEV int word_seq(unsigned long long w) { return (int)(w >> 32); }

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
struct LaunchStats {
    std::atomic<long long> launches{0};  // kernels launched through this layer (contexts may run on several host threads)
};
inline LaunchStats& launch_stats() {
    static LaunchStats s;
    return s;
}

// Optional per-kernel device timing (bench / DESIGN evidence): when enabled, every launch whose
// name contains `filter` is bracketed by two CUDA events on the launching stream.
struct Profiler {
    bool on = false;
    std::string filter;
#if PB_CUDA
    struct Rec { const char* name; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        PB_CUDA_CHECK(cudaEventCreate(&e));
        return e;
    }
    ~Profiler() {
        for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        for (auto e : pool) cudaEventDestroy(e);
    }
#endif
    // launch names are typeid names (mangled) or plain strings; the filter is matched on the readable form
    std::map<const char*, std::string> readable;
    const std::string& pretty(const char* name) {
        auto it = readable.find(name);
        if (it != readable.end()) return it->second;
        int st = 0;
        char* dm = abi::__cxa_demangle(name, nullptr, nullptr, &st);
        std::string out = (st == 0 && dm) ? dm : name;
        free(dm);
        return readable[name] = out;
    }
    // filter: empty = everything, else '|'-separated substrings
    bool wants(const char* name) {
        if (!on) return false;
        if (filter.empty()) return true;
        const std::string& p = pretty(name);
        size_t b = 0;
        while (b <= filter.size()) {
            size_t e = filter.find('|', b);
            if (e == std::string::npos) e = filter.size();
            if (e > b && p.find(filter.substr(b, e - b)) != std::string::npos) return true;
            b = e + 1;
        }
        return false;
    }
    void start(const char* f) {
        collect(nullptr);
        filter = f ? f : "";
        on = true;
    }
    // name -> (launches, total ms); empties the record list
    void collect(std::map<std::string, std::pair<long long, double>>* out) {
#if PB_CUDA
        for (auto& r : recs) {
            if (out) {
                PB_CUDA_CHECK(cudaEventSynchronize(r.b));
                float ms = 0;
                PB_CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
                auto& e = (*out)[r.name];
                e.first++;
                e.second += ms;
            }
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        recs.clear();
#endif
    }
};
struct ProfScope {
#if PB_CUDA
    Profiler* p = nullptr;
    cudaStream_t s;
    cudaEvent_t a, b;
    const char* name;
    bool trace = false;
    static bool tracing() { static const bool t = getenv("PB_TRACE") != nullptr; return t; }   // diagnostic: name + sync per launch
    ProfScope(Profiler* prof, const char* nm, cudaStream_t st) : s(st), name(nm) {
        if (tracing()) { trace = true; fprintf(stderr, "[pb trace] %s ...", nm); fflush(stderr); }
        if (prof && prof->wants(nm)) {
            p = prof;
            a = p->get(); b = p->get();
            cudaEventRecord(a, s);
        }
    }
    ~ProfScope() {
        if (p) { cudaEventRecord(b, s); p->recs.push_back({name, a, b}); }
        if (trace) { const cudaError_t e = cudaStreamSynchronize(s); fprintf(stderr, " %s\n", e == cudaSuccess ? "ok" : cudaGetErrorString(e)); fflush(stderr); }
    }
#else
    ProfScope(Profiler*, const char*, cudaStream_t) {}
#endif
};

#if PB_CUDA
template <class F>
PB_GLOBAL void __launch_bounds__(256) k_for(F f, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) f(i);
}

// Sync-free ordered dataflow: item i may only depend on items j < i.  CTAs take a ticket so the
// logical CTA order equals the order in which CTAs became resident — every dependency of a
// resident item is resident or finished, hence the polling loop cannot deadlock.  Lanes never
// block on one another: a lane whose dependencies are not yet met simply retries.
#ifndef PB_ORDERED_BACKOFF_NS
#define PB_ORDERED_BACKOFF_NS 40
#endif
template <class F>
PB_GLOBAL void __launch_bounds__(128) k_ordered(F f, int n, int* ticket, unsigned backoffNs) {
    __shared__ int base;
    if (threadIdx.x == 0) base = atomicAdd(ticket, 1) * blockDim.x;
    __syncthreads();
    const int i = base + threadIdx.x;
    bool pending = i < n;
    // ncu (profiles/r02_ordered_ncu.md): with every waiting lane polling back to back the kernel runs at 86 % of the L1/LSU
    // throughput and 67 % of the L2 throughput while issuing 0.1 instructions per cycle — the pollers slow down the few lanes
    // on the critical path.  Lanes whose dependencies are not ready therefore back off between attempts.
    int fails = 0;
    while (pending) {
        if (f.try_run(i)) pending = false;
        else {
            fails++;
            const unsigned ns = backoffNs;
            if (ns) __nanosleep(fails < 4 ? ns : fails < 16 ? 4 * ns : 16 * ns);
        }
    }
}

// grid-stride loop whose bound lives in device memory (frontier sizes); global thread 0 first runs
// the functor's block0() hook.
template <class F>
PB_GLOBAL void __launch_bounds__(256) k_for_dev(F f, const int* nDev) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) f.block0();
    const int n = *nDev;
    for (int i = t; i < n; i += gridDim.x * blockDim.x) f(i);
}
#endif

struct Exec {
    cudaStream_t stream = 0;
    int sm_count = 148;
    int* ticket = nullptr;  // device int used by ordered launches
    Profiler* prof = nullptr;

    template <class F>
    void for_each(int n, const F& f) const {
        if (n <= 0) return;
        launch_stats().launches++;
        ProfScope ps(prof, typeid(F).name(), stream);
#if PB_CUDA
        const int block = 256;
        long long want = ((long long)n + block - 1) / block;
        int grid = (int)std::min<long long>(want, (long long)sm_count * 16);
        k_for<F><<<grid, block, 0, stream>>>(f, n);
        PB_CUDA_CHECK(cudaGetLastError());
#else
        for (int i = 0; i < n; i++) f(i);
#endif
    }

    template <class F>
    void ordered(int n, const F& f) const {
        if (n <= 0) return;
        launch_stats().launches++;
        ProfScope ps(prof, typeid(F).name(), stream);
#if PB_CUDA
        const int block = 128;
        PB_CUDA_CHECK(cudaMemsetAsync(ticket, 0, sizeof(int), stream));
        int grid = (n + block - 1) / block;
        static const unsigned backoff = getenv("PB_BACKOFF_NS") ? (unsigned)atoi(getenv("PB_BACKOFF_NS")) : PB_ORDERED_BACKOFF_NS;
        k_ordered<F><<<grid, block, 0, stream>>>(f, n, ticket, backoff);
        PB_CUDA_CHECK(cudaGetLastError());
#else
        for (int i = 0; i < n; i++) {
            if (!f.try_run(i)) throw Error("ordered dataflow: dependency of a later item (emulation)");
        }
#endif
    }

    // f(i) for i < *nDev (device-resident bound); f.block0() runs once first
    template <class F>
    void for_each_dev(const int* nDev, const F& f) const {
        launch_stats().launches++;
        ProfScope ps(prof, typeid(F).name(), stream);
#if PB_CUDA
        k_for_dev<F><<<sm_count * 4, 256, 0, stream>>>(f, nDev);
        PB_CUDA_CHECK(cudaGetLastError());
#else
        f.block0();
        const int n = *nDev;
        for (int i = 0; i < n; i++) f(i);
#endif
    }

#if !PB_CUDA
    // sequential reference forms of the serial passes (heap flood, capped BFS): host emulation only — the CUDA library has no
    // single-thread kernel
    template <class F>
    void single(const F& f) const {
        launch_stats().launches++;
        f();
    }
#endif
};

}  // namespace pb
