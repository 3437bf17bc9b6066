// pb_shardsweep.h — cell-range sharding of the Jacobi / propagation sweep LOOPS (js/climate-util.js:5-25,
// js/ocean.js:168-189, js/temperature.js:33-51, js/precipitation.js:118-179, :555-598) across the GPUs of one box.
//
// One process per GPU, every rank holds the whole planet (mesh + fields, replicated) and runs the class S/R/F stages and
// the pointwise kernels redundantly.  What is sharded is the part that dominates a large planet: the sweep loops
// (2 492 graph sweeps per climate pass at 10M cells).  Inside a loop a rank only computes the rows of its contiguous
// cell-id range [lo, hi) — Fibonacci ids advance in z, so a range is a latitude band that touches the two adjacent ranges
// (and, through the pole vertex N-1, the first one).  Global cell ids are kept, so a halo value lives at the same index
// on every rank:
//   sweep k     k_shard_sweep<F>: the first few CTAs wait for the adjacent ranks' flags ("my sweep k-1 boundary values are in
//               your buffer"), compute the rows other ranks read (send lists), store each new value straight into the peers'
//               buffers over NVLink (CUDA-IPC mapped cudaMalloc memory), fence, and the last of them raises this rank's flag at
//               the peers; all other CTAs sweep the interior rows without waiting → the exchange hides behind the interior.
//   loop end    every rank stores its range of the result into every peer's buffer (all-gather by peer stores), a
//               device-side barrier over the ranks follows, and the replicated field is whole again on every GPU.
// No NCCL call and no host round trip inside a loop; torch.distributed only carries the IPC handles once.
// Results are bit-identical to the single-GPU pass: every row is computed by the same functor from the same neighbour values.
#pragma once
#include <thread>
#include "pb_engine.h"

namespace pb {

constexpr int kShardMaxRanks = 16;
constexpr int kShardMaxAdj = 8;

struct ShardSweepArgs {
    int lo, hi;                                     // owned rows
    volatile int* flags; int waitValue;             // my flag array (written by the peers); 0 = nothing to wait for
    int setValue; unsigned* ticket;
    int nAdj, sendTotal, nBoundaryCtas;
    const unsigned* boundaryMask;                   // bit (row - lo) set: the row is on a send list
    int adjRank[kShardMaxAdj]; const int* sendIdx[kShardMaxAdj]; int sendCount[kShardMaxAdj];
    float* peerOut[kShardMaxAdj]; volatile int* peerFlag[kShardMaxAdj];
    int itemLo, itemHi;                             // item functors: the items whose rows lie in [lo, hi)
};

template <class F, class = void> struct ShardIsItems { static constexpr bool value = false; };
template <class F> struct ShardIsItems<F, decltype((void)F::kItems)> { static constexpr bool value = true; };
template <class F> PB_DEV void shard_do_row(const F& f, int row) { if constexpr (ShardIsItems<F>::value) f.by_row(row); else f(row); }

#if PB_CUDA
template <class F>
__global__ void __launch_bounds__(256) k_shard_sweep(F f, const float* out, const ShardSweepArgs a) {
    if ((int)blockIdx.x < a.nBoundaryCtas) {
        __shared__ int sLast;
        if (threadIdx.x == 0 && a.waitValue > 0) {
            for (int p = 0; p < a.nAdj; p++)
                while (a.flags[a.adjRank[p]] < a.waitValue) __nanosleep(40);
            __threadfence_system();
        }
        __syncthreads();
        bool pushed = false;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.sendTotal; j += a.nBoundaryCtas * blockDim.x) {
            int p = 0, base = 0;
            while (j >= base + a.sendCount[p]) base += a.sendCount[p++];
            const int row = a.sendIdx[p][j - base];
            shard_do_row(f, row);                      // a row on two lists is computed twice: same value
            a.peerOut[p][row] = out[row];
            pushed = true;
        }
        if (pushed) __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            sLast = (atomicAdd(a.ticket, 1u) == (unsigned)a.nBoundaryCtas - 1u);
        }
        __syncthreads();
        if (!sLast || threadIdx.x != 0) return;
        __threadfence_system();
        for (int p = 0; p < a.nAdj; p++) *a.peerFlag[p] = a.setValue;
        *a.ticket = 0u;
        __threadfence_system();
        return;
    }
    const int nCtas = gridDim.x - a.nBoundaryCtas;
    if constexpr (ShardIsItems<F>::value) {
        for (int i = a.itemLo + (blockIdx.x - a.nBoundaryCtas) * blockDim.x + threadIdx.x; i < a.itemHi; i += nCtas * blockDim.x) {
            const int k = f.row_of(i) - a.lo;
            const bool boundary = a.boundaryMask && ((a.boundaryMask[k >> 5] >> (k & 31)) & 1u);
            if (!boundary) f(i);
        }
    } else {
        for (int i = a.lo + (blockIdx.x - a.nBoundaryCtas) * blockDim.x + threadIdx.x; i < a.hi; i += nCtas * blockDim.x) {
            const int k = i - a.lo;
            const bool boundary = a.boundaryMask && ((a.boundaryMask[k >> 5] >> (k & 31)) & 1u);
            if (!boundary) f(i);
        }
    }
}
// my range of `src` → the same range of every peer's buffer
__global__ void __launch_bounds__(256) k_shard_allgather(const float* src, int lo, int hi, int nPeers, float* const* peerDst) {
    const int n = hi - lo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float v = src[lo + i];
        for (int p = 0; p < nPeers; p++) peerDst[p][lo + i] = v;
    }
}
// device-side barrier over the ranks: write my epoch into slot[me] of every rank, wait for every slot of mine
__global__ void k_rank_barrier(volatile int* mine, int* const* peerBar, int nPeers, int me, int world, int value) {
    __threadfence_system();
    for (int p = threadIdx.x; p < nPeers; p += blockDim.x) ((volatile int*)peerBar[p])[me] = value;
    __threadfence_system();
    for (int r = threadIdx.x; r < world; r += blockDim.x)
        if (r != me) while (mine[r] < value) __nanosleep(100);
    __threadfence_system();
}
#endif

struct SweepShards {
    Mesh* m; int rank, world, N;
    std::vector<int> bounds;
    int lo = 0, hi = 0;
    float* buf[2] = {nullptr, nullptr};
    int* ctl = nullptr;                         // [0, world): halo flags written by the peers; [world, 2·world): barrier slots
    struct Peer { float* buf[2] = {nullptr, nullptr}; int* ctl = nullptr; void* mapped[3] = {nullptr, nullptr, nullptr}; bool local = false; };
    Peer peers[kShardMaxRanks];
    std::vector<int> adj;                       // adjacent ranks (halo partners), ascending
    std::vector<int> sendBase, sendCount;       // per adjacent rank: slice of sendIdx
    DevBuf<int> sendIdx;
    DevBuf<unsigned> boundaryMask, ticket;
    DevBuf<float*> peerBufTab[2];               // device tables for the all-gather kernel
    DevBuf<int*> peerBarTab;
    int sendTotal = 0;
    int epoch = 0, barEpoch = 0;
    long long minCells = 2000000;               // loops of smaller planets run unsharded (a sweep is shorter than the flag chain)
    bool connected = false;
    long long haloBytesPerSweep = 0, sweepsRun = 0, loopsRun = 0;

    SweepShards(Mesh* mesh, int rank_, int world_) : m(mesh), rank(rank_), world(world_), N(mesh->N) {
        if (world < 1 || world > kShardMaxRanks || rank < 0 || rank >= world) throw std::invalid_argument("bad rank / world size");
        bounds.resize(world + 1);
        for (int k = 0; k <= world; k++) bounds[k] = (int)(((long long)N * k) / world);
        lo = bounds[rank]; hi = bounds[rank + 1];
        // send lists: my rows that have a neighbour owned by rank p (the graph is undirected, so these are also exactly my
        // rows that read a value of p)
        const int* off = m->hOffCopy.data(); const int* ad = m->hAdjCopy.data();
        std::vector<std::vector<int>> lists(world);
        std::vector<unsigned> mask(((size_t)(hi - lo) + 31) / 32, 0u);
        for (int r = lo; r < hi; r++) {
            int seen[kShardMaxRanks]; int ns = 0;
            for (int j = off[r]; j < off[r + 1]; j++) {
                const int nb = ad[j];
                if (nb >= lo && nb < hi) continue;
                const int p = (int)(std::upper_bound(bounds.begin(), bounds.end(), nb) - bounds.begin()) - 1;
                bool dup = false;
                for (int q = 0; q < ns; q++) if (seen[q] == p) dup = true;
                if (!dup) { seen[ns++] = p; lists[p].push_back(r); }
            }
            if (ns) mask[(size_t)(r - lo) >> 5] |= 1u << ((r - lo) & 31);
        }
        std::vector<int> flat;
        for (int p = 0; p < world; p++) {
            if (lists[p].empty()) continue;
            adj.push_back(p); sendBase.push_back((int)flat.size()); sendCount.push_back((int)lists[p].size());
            flat.insert(flat.end(), lists[p].begin(), lists[p].end());
        }
        if ((int)adj.size() > kShardMaxAdj) throw std::invalid_argument("a shard touches more than 8 other shards");
        sendTotal = (int)flat.size();
        haloBytesPerSweep = 4ll * sendTotal;
        const cudaStream_t s = m->ex().stream;
        dev_copy(sendIdx.ensure(std::max<size_t>(1, flat.size())), flat.data(), sizeof(int) * flat.size(), 0, s);
        dev_copy(boundaryMask.ensure(std::max<size_t>(1, mask.size())), mask.data(), sizeof(unsigned) * mask.size(), 0, s);
        dev_memset(ticket.ensure(1), 0, sizeof(unsigned), s);
        // peer-visible allocations: one cudaMalloc each (exportable through CUDA IPC)
        buf[0] = (float*)dev_alloc(sizeof(float) * (size_t)N);
        buf[1] = (float*)dev_alloc(sizeof(float) * (size_t)N);
        ctl = (int*)dev_alloc(sizeof(int) * 2 * (size_t)world);
        dev_memset(ctl, 0, sizeof(int) * 2 * (size_t)world, s);
        stream_sync(s);
    }
    ~SweepShards() {
#if PB_CUDA
        for (auto& p : peers) for (void* q : p.mapped) if (q) cudaIpcCloseMemHandle(q);
#endif
        dev_free(buf[0]); dev_free(buf[1]); dev_free(ctl);
    }
    bool active() const { return connected && world > 1 && (long long)N >= minCells; }

    void export_handles(unsigned char* out) {   // 3 × 64 bytes: buf0, buf1, ctl
#if PB_CUDA
        void* ptrs[3] = {buf[0], buf[1], ctl};
        for (int k = 0; k < 3; k++) {
            cudaIpcMemHandle_t h;
            PB_CUDA_CHECK(cudaIpcGetMemHandle(&h, ptrs[k]));
            static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
            memcpy(out + 64 * k, &h, 64);
        }
#else
        memset(out, 0, 192);
        void* ptrs[3] = {buf[0], buf[1], ctl};     // emulation: ranks are threads of one process, the "handles" are the pointers
        memcpy(out, ptrs, sizeof ptrs);
#endif
    }
    void connect(int peerRank, const unsigned char* handles) {
        if (peerRank < 0 || peerRank >= world || peerRank == rank) throw std::invalid_argument("peer rank out of range");
        Peer& p = peers[peerRank];
#if PB_CUDA
        for (int k = 0; k < 3; k++) {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + 64 * k, 64);
            PB_CUDA_CHECK(cudaIpcOpenMemHandle(&p.mapped[k], h, cudaIpcMemLazyEnablePeerAccess));
        }
        p.buf[0] = (float*)p.mapped[0]; p.buf[1] = (float*)p.mapped[1]; p.ctl = (int*)p.mapped[2];
#else
        void* ptrs[3];
        memcpy(ptrs, handles, sizeof ptrs);
        p.buf[0] = (float*)ptrs[0]; p.buf[1] = (float*)ptrs[1]; p.ctl = (int*)ptrs[2];
#endif
        p.local = true;
        bool all = true;
        for (int r = 0; r < world; r++) if (r != rank && !peers[r].local) all = false;
        if (all) finish_connect();
    }
    void finish_connect() {
        std::vector<float*> t0, t1; std::vector<int*> tb;
        for (int r = 0; r < world; r++) if (r != rank) { t0.push_back(peers[r].buf[0]); t1.push_back(peers[r].buf[1]); tb.push_back(peers[r].ctl + world); }
        const cudaStream_t s = m->ex().stream;
        dev_copy(peerBufTab[0].ensure(t0.size()), t0.data(), sizeof(float*) * t0.size(), 0, s);
        dev_copy(peerBufTab[1].ensure(t1.size()), t1.data(), sizeof(float*) * t1.size(), 0, s);
        dev_copy(peerBarTab.ensure(tb.size()), tb.data(), sizeof(int*) * tb.size(), 0, s);
        stream_sync(s);
        connected = true;
    }

    void barrier() {
        const Exec& x = m->ex();
        barEpoch++;
        launch_stats().launches++;
#if PB_CUDA
        ProfScope ps(x.prof, "pb::k_rank_barrier", x.stream);
        k_rank_barrier<<<1, 32, 0, x.stream>>>(ctl + world, peerBarTab.p, world - 1, rank, world, barEpoch);
        PB_CUDA_CHECK(cudaGetLastError());
#else
        std::atomic_thread_fence(std::memory_order_seq_cst);
        for (int r = 0; r < world; r++) if (r != rank) ((volatile int*)(peers[r].ctl + world))[rank] = barEpoch;
        std::atomic_thread_fence(std::memory_order_seq_cst);
        for (int r = 0; r < world; r++) if (r != rank) while (((volatile int*)(ctl + world))[r] < barEpoch) std::this_thread::yield();
        std::atomic_thread_fence(std::memory_order_seq_cst);
#endif
    }

    // `passes` sweeps dst = F(src) over the sharded rows; field is the replicated array (whole on entry and on exit)
    // itemLo/itemHi >= 0: `make` builds an item functor (kItems) over a compacted list; only those items are swept, every other
    // cell keeps the value it has in `field` (both buffers start as the field)
    template <class Make>
    void run(float* field, int passes, const Make& make, int itemLo = -1, int itemHi = -1) {
        const Exec& x = m->ex();
        const cudaStream_t s = x.stream;
        const int nA = (int)adj.size();
        dev_copy(buf[0], field, sizeof(float) * (size_t)N, 2, s);
        if (itemLo >= 0) dev_copy(buf[1], field, sizeof(float) * (size_t)N, 2, s);
        barrier();                                  // every rank has left the previous loop: its buffers may be written
        ShardSweepArgs a{};
        a.lo = lo; a.hi = hi; a.flags = ctl; a.ticket = ticket.p; a.nAdj = nA; a.sendTotal = sendTotal;
        a.boundaryMask = nA ? boundaryMask.p : nullptr;
        for (int p = 0; p < nA; p++) {
            a.adjRank[p] = adj[p]; a.sendIdx[p] = sendIdx.p + sendBase[p]; a.sendCount[p] = sendCount[p];
            a.peerFlag[p] = peers[adj[p]].ctl + rank;
        }
        a.nBoundaryCtas = nA ? std::max(1, std::min((sendTotal + 255) / 256, 32)) : 0;
        a.itemLo = itemLo; a.itemHi = itemHi;
        const int rows = hi - lo;
        const int grid = a.nBoundaryCtas + std::max(1, std::min((rows + 255) / 256, x.sm_count * 8 - a.nBoundaryCtas));
        for (int k = 1; k <= passes; k++) {
            const float* in = buf[(k - 1) & 1]; float* out = buf[k & 1];
            ++epoch;
            a.waitValue = k > 1 ? epoch - 1 : 0; a.setValue = epoch;
            for (int p = 0; p < nA; p++) a.peerOut[p] = peers[adj[p]].buf[k & 1];
            auto f = make(in, out);
            launch_stats().launches++;
#if PB_CUDA
            ProfScope ps(x.prof, typeid(f).name(), s);
            k_shard_sweep<<<grid, 256, 0, s>>>(f, out, a);
#else
            (void)grid;
            if (a.waitValue > 0)
                for (int p = 0; p < nA; p++) while (((volatile int*)a.flags)[a.adjRank[p]] < a.waitValue) std::this_thread::yield();
            std::atomic_thread_fence(std::memory_order_seq_cst);
            for (int p = 0; p < nA; p++)
                for (int j = 0; j < a.sendCount[p]; j++) { const int row = a.sendIdx[p][j]; shard_do_row(f, row); a.peerOut[p][row] = out[row]; }
            std::atomic_thread_fence(std::memory_order_seq_cst);
            for (int p = 0; p < nA; p++) *a.peerFlag[p] = a.setValue;
            if constexpr (ShardIsItems<decltype(f)>::value) {
                for (int i = itemLo; i < itemHi; i++) { const int q = f.row_of(i) - lo; if (!((a.boundaryMask[q >> 5] >> (q & 31)) & 1u)) f(i); }
            } else {
                for (int i = lo; i < hi; i++) { const int q = i - lo; if (!((a.boundaryMask[q >> 5] >> (q & 31)) & 1u)) f(i); }
            }
#endif
        }
#if PB_CUDA
        PB_CUDA_CHECK(cudaGetLastError());
#endif
        // all-gather: my range of the result into every peer's buffer, then the rank barrier
        const int fin = passes & 1;
        launch_stats().launches++;
#if PB_CUDA
        {
            ProfScope ps(x.prof, "pb::k_shard_allgather", s);
            const int g = std::max(1, std::min((rows + 255) / 256, x.sm_count * 4));
            k_shard_allgather<<<g, 256, 0, s>>>(buf[fin], lo, hi, world - 1, peerBufTab[fin].p);
            PB_CUDA_CHECK(cudaGetLastError());
        }
#else
        for (int r = 0; r < world; r++) if (r != rank) memcpy(peers[r].buf[fin] + lo, buf[fin] + lo, sizeof(float) * (size_t)rows);
#endif
        barrier();
        dev_copy(field, buf[fin], sizeof(float) * (size_t)N, 2, s);
        sweepsRun += passes; loopsRun++;
    }
};

// field ← F^passes(field) with make(src, dst) building the sweep functor of one pass (dst[r] = F(src)[r])
template <class Make>
void sweep_loop(Mesh& m, float* field, int passes, float* scratch, const Make& make) {
    if (passes <= 0) return;
    if (m.shards && m.shards->active()) { m.shards->run(field, passes, make); return; }
    float* src = field; float* dst = scratch;
    for (int p = 0; p < passes; p++) {
        m.ex().for_each(m.N, make((const float*)src, dst));
        std::swap(src, dst);
    }
    if (src != field) dev_copy(field, src, sizeof(float) * (size_t)m.N, 2, m.ex().stream);
}
inline bool sweeps_sharded(const Mesh& m) { return m.shards && m.shards->active(); }
// loop over a compacted item list (`count` items, rows ascending; in a sharded run this rank sweeps the items [itemLo, itemHi) of
// its cell-id range); scratch must already hold the values of the cells the items never write
template <class Make>
void sweep_loop_items(Mesh& m, int count, int itemLo, int itemHi, float* field, int passes, float* scratch, const Make& make) {
    if (passes <= 0) return;
    if (m.shards && m.shards->active()) { m.shards->run(field, passes, make, itemLo, itemHi); return; }
    float* src = field; float* dst = scratch;
    for (int p = 0; p < passes; p++) {
        m.ex().for_each(count, make((const float*)src, dst));
        std::swap(src, dst);
    }
    if (src != field) dev_copy(field, src, sizeof(float) * (size_t)m.N, 2, m.ex().stream);
}
inline void smooth_field_impl(Mesh& m, float* field, int passes) {
    const Csr g = m.csr();
    sweep_loop(m, field, passes, m.tmp.ensure(m.N), [g](const float* src, float* dst) { return SmoothFieldK{g, src, dst}; });
}

}  // namespace pb
