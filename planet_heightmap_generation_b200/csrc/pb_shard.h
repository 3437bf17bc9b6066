// pb_shard.h — device-side halo exchange for the cell-range-sharded Jacobi sweeps.
//
// One process per GPU.  Every rank owns a contiguous cell-id range plus the halo cells its rows read; the two
// ping-pong field buffers and two small flag arrays of a shard are ordinary cudaMalloc memory exported with CUDA
// IPC, so a peer's kernels store straight into them over NVLink / NVSwitch — no NCCL call and no host round trip
// per sweep.  Pass k of a rank is ONE kernel launch (k_shard_sweep):
//     the first few CTAs   wait until every peer's flag says "my pass k-1 values are in your halo", compute the rows on
//                          the send lists (exactly the rows that read halo slots), store each new value straight into
//                          the peer's halo slot, fence system-wide, and the last of them (ticket) writes
//                          peer.flags[me] = epoch + k;
//     all other CTAs       sweep the remaining owned rows, buf[k%2] = smoothField(buf[(k-1)%2]), without waiting.
// A peer cannot overwrite a halo slot that is still being read: its pass k stores follow its own wait for this rank's
// pass k-1 flag, which is raised only after this rank's boundary rows — the only readers of halo slots — are done.
// k_wait_flags / k_flag_set (one thread each) bracket a call: the previous call of every peer must be over before
// halo slots are reused, and the final halo values must have arrived before the result is copied out.
#pragma once
#include "pb_engine.h"

#if PB_CUDA
namespace pb {

__global__ void k_flag_set(volatile int* peerFlag, int value) {
    __threadfence_system();
    *peerFlag = value;
    __threadfence_system();
}
__global__ void k_wait_flags(volatile int* flags, const int* peerRanks, int nPeers, int value) {
    for (int p = threadIdx.x; p < nPeers; p += blockDim.x)
        while (flags[peerRanks[p]] < value) __nanosleep(200);
    __threadfence_system();
}

constexpr int kMaxPeers = 8;
struct ShardSweepArgs {
    Csr g; const float* in; float* out; int nOwn;
    volatile int* flags; int waitValue;              // 0 = nothing to wait for
    int nPeers; int setValue; unsigned* ticket;
    int sendTotal, nBoundaryCtas;
    const unsigned* boundaryMask;                    // bit i set: row i is on a send list (nullptr: none)
    int peerRank[kMaxPeers]; const int* sendIdx[kMaxPeers]; int sendCount[kMaxPeers];
    float* peerDst[kMaxPeers]; volatile int* peerFlag[kMaxPeers];
};
// wait → sweep → push → flag in one launch, with the exchange hidden behind the interior rows.
//   * The rows that read halo slots are exactly the rows on the send lists (the graph is undirected), so only the
//     first `nBoundaryCtas` CTAs wait for the peers' flags; they compute the send-list rows, store each new value
//     straight into the peer's halo slot, fence, and the last of them (ticket) raises this rank's flag at the peers.
//   * All other CTAs sweep the remaining rows without waiting for anybody; they skip send-list rows (one bit per row), so once the flag is up nothing on this GPU reads a halo slot of the
//     source buffer any more and the peer may overwrite it for its next pass.
__global__ void __launch_bounds__(256) k_shard_sweep(const ShardSweepArgs a) {
    SmoothFieldK f{a.g, a.in, a.out};
    if ((int)blockIdx.x < a.nBoundaryCtas) {
        __shared__ int sLast;
        if (threadIdx.x == 0 && a.waitValue > 0) {
            for (int p = 0; p < a.nPeers; p++)
                while (a.flags[a.peerRank[p]] < a.waitValue) __nanosleep(64);
            __threadfence_system();
        }
        __syncthreads();
        bool pushed = false;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.sendTotal; j += a.nBoundaryCtas * blockDim.x) {
            int p = 0, base = 0;
            while (j >= base + a.sendCount[p]) base += a.sendCount[p++];
            const int row = a.sendIdx[p][j - base];
            f(row);                                    // a row on two lists is computed twice: same value
            a.peerDst[p][j - base] = a.out[row];
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
        for (int p = 0; p < a.nPeers; p++) *a.peerFlag[p] = a.setValue;
        *a.ticket = 0u;
        __threadfence_system();
        return;
    }
    const int nCtas = gridDim.x - a.nBoundaryCtas;
    for (int i = (blockIdx.x - a.nBoundaryCtas) * blockDim.x + threadIdx.x; i < a.nOwn; i += nCtas * blockDim.x) {
        const bool boundary = a.boundaryMask && ((a.boundaryMask[i >> 5] >> (i & 31)) & 1u);
        if (!boundary) f(i);
    }
}

struct ShardPeer {
    int rank = -1;
    int sendCount = 0, sendBase = 0;      // slice of sendIdx (ascending local ids)
    int recvOffset = 0;                   // where my block starts in the peer's local field
    float* buf[2] = {nullptr, nullptr};   // peer's ping-pong buffers (IPC-mapped)
    int* flags = nullptr;                 // peer's "values delivered" flags   [world]
    int* done = nullptr;                  // peer's "call finished" flags      [world]
    void* mapped[4] = {nullptr, nullptr, nullptr, nullptr};
};

struct Shard {
    Mesh* m; int nOwn, nLocal, myRank, world;
    std::vector<ShardPeer> peers;
    DevBuf<int> sendIdx, peerRanksDev;
    DevBuf<unsigned> ticket, boundaryMask;
    float* buf[2] = {nullptr, nullptr};
    int* flags = nullptr; int* done = nullptr;
    int epoch = 0, sendTotal = 0;

    Shard(Mesh* mesh, int own, int rank, int worldSize, int nPeers, const int* peerRanks, const int* sendCounts,
          const int* sendIdxHost, const int* peerRecvOffset)
        : m(mesh), nOwn(own), nLocal(mesh->N), myRank(rank), world(worldSize) {
        if (own <= 0 || own > mesh->N) throw std::invalid_argument("nOwn out of range");
        if (nPeers > kMaxPeers) throw std::invalid_argument("more than 8 peers per shard");
        int total = 0;
        peers.resize(nPeers);
        for (int p = 0; p < nPeers; p++) {
            peers[p].rank = peerRanks[p]; peers[p].sendCount = sendCounts[p]; peers[p].sendBase = total;
            peers[p].recvOffset = peerRecvOffset[p];
            total += sendCounts[p];
        }
        for (int i = 0; i < total; i++) if (sendIdxHost[i] < 0 || sendIdxHost[i] >= own) throw std::invalid_argument("send index is not an owned cell");
        std::vector<unsigned> mask(((size_t)own + 31) / 32, 0u);
        for (int i = 0; i < total; i++) mask[sendIdxHost[i] >> 5] |= 1u << (sendIdxHost[i] & 31);
        sendTotal = total;
        const cudaStream_t s = m->ex().stream;
        dev_copy(sendIdx.ensure(total), sendIdxHost, sizeof(int) * (size_t)total, 0, s);
        dev_copy(boundaryMask.ensure(mask.size()), mask.data(), sizeof(unsigned) * mask.size(), 0, s);
        dev_copy(peerRanksDev.ensure(nPeers), peerRanks, sizeof(int) * (size_t)nPeers, 0, s);
        // IPC-exportable allocations (plain cudaMalloc, one allocation each)
        PB_CUDA_CHECK(cudaMalloc(&buf[0], sizeof(float) * (size_t)nLocal));
        PB_CUDA_CHECK(cudaMalloc(&buf[1], sizeof(float) * (size_t)nLocal));
        PB_CUDA_CHECK(cudaMalloc(&flags, sizeof(int) * (size_t)world));
        PB_CUDA_CHECK(cudaMalloc(&done, sizeof(int) * (size_t)world));
        PB_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)world, s));
        PB_CUDA_CHECK(cudaMemsetAsync(done, 0, sizeof(int) * (size_t)world, s));
        PB_CUDA_CHECK(cudaMemsetAsync(ticket.ensure(1), 0, sizeof(unsigned), s));
        stream_sync(s);
    }
    ~Shard() {
        for (auto& p : peers) for (void* q : p.mapped) if (q) cudaIpcCloseMemHandle(q);
        cudaFree(buf[0]); cudaFree(buf[1]); cudaFree(flags); cudaFree(done);
    }
    void export_handles(unsigned char* out) {   // 4 × 64 bytes: buf0, buf1, flags, done
        void* ptrs[4] = {buf[0], buf[1], flags, done};
        for (int k = 0; k < 4; k++) {
            cudaIpcMemHandle_t h;
            PB_CUDA_CHECK(cudaIpcGetMemHandle(&h, ptrs[k]));
            static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
            memcpy(out + 64 * k, &h, 64);
        }
    }
    void connect(int peerIndex, const unsigned char* handles) {
        if (peerIndex < 0 || peerIndex >= (int)peers.size()) throw std::invalid_argument("peer index out of range");
        ShardPeer& p = peers[peerIndex];
        for (int k = 0; k < 4; k++) {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + 64 * k, 64);
            PB_CUDA_CHECK(cudaIpcOpenMemHandle(&p.mapped[k], h, cudaIpcMemLazyEnablePeerAccess));
        }
        p.buf[0] = (float*)p.mapped[0]; p.buf[1] = (float*)p.mapped[1]; p.flags = (int*)p.mapped[2]; p.done = (int*)p.mapped[3];
    }

    // `passes` smoothField sweeps (js/climate-util.js:5-25) over the sharded mesh; field = owned + halo, halo current.
    void smooth_field(float* field, int passes) {
        if (passes <= 0) return;
        for (auto& p : peers) if (!p.buf[0]) throw std::invalid_argument("pb_shard_connect has not been called for every peer");
        const Exec& x = m->ex();
        const cudaStream_t s = x.stream;
        const int nP = (int)peers.size();
        // the previous call of every peer must be over before its halo slots in my buffers are reused
        if (nP) { launch_stats().launches++; k_wait_flags<<<1, 32, 0, s>>>(done, peerRanksDev.p, nP, epoch); }
        dev_copy(buf[0], field, sizeof(float) * (size_t)nLocal, 2, s);
        ShardSweepArgs a{};
        a.g = m->csr(); a.nOwn = nOwn; a.flags = flags; a.nPeers = nP; a.ticket = ticket.p;
        for (int p = 0; p < nP; p++) {
            a.peerRank[p] = peers[p].rank; a.sendIdx[p] = sendIdx.p + peers[p].sendBase; a.sendCount[p] = peers[p].sendCount;
            a.peerFlag[p] = peers[p].flags + myRank;
        }
        a.sendTotal = sendTotal; a.nBoundaryCtas = nP ? std::max(1, std::min((a.sendTotal + 255) / 256, 32)) : 0;
        const int grid = a.nBoundaryCtas + std::max(1, std::min((nOwn + 255) / 256, x.sm_count * 8 - a.nBoundaryCtas));
        static const bool noSync = getenv("PB_SHARD_NOSYNC") != nullptr;   // timing diagnostic only: results are wrong
        a.boundaryMask = nP ? boundaryMask.p : nullptr;
        if (noSync) { a.nPeers = 0; a.nBoundaryCtas = 0; a.boundaryMask = nullptr; }
        for (int k = 1; k <= passes; k++) {
            a.in = buf[(k - 1) & 1]; a.out = buf[k & 1];
            a.waitValue = (k > 1 && !noSync) ? epoch + k - 1 : 0; a.setValue = epoch + k;
            for (int p = 0; p < nP; p++) a.peerDst[p] = peers[p].buf[k & 1] + peers[p].recvOffset;
            launch_stats().launches++;
            k_shard_sweep<<<grid, 256, 0, s>>>(a);
        }
        PB_CUDA_CHECK(cudaGetLastError());
        if (nP && !noSync) { launch_stats().launches++; k_wait_flags<<<1, 32, 0, s>>>(flags, peerRanksDev.p, nP, epoch + passes); }
        dev_copy(field, buf[passes & 1], sizeof(float) * (size_t)nLocal, 2, s);
        epoch += passes;
        for (auto& p : peers) { launch_stats().launches++; k_flag_set<<<1, 1, 0, s>>>(p.done + myRank, epoch); }
        PB_CUDA_CHECK(cudaGetLastError());
    }
};

}  // namespace pb
#endif
