// pb_tiled.h — TMA-staged shared-memory tiles for the Jacobi sweeps (kernel family K1).
//
// Fibonacci ids advance monotonically in z, so the neighbours of a run of consecutive cells lie in a
// narrow id window (|nb - r| <= W with W ≈ 5·√N; the host measures W per mesh and treats the few
// outliers — the pole vertex and its ring — through the global path).  One CTA owns TILE consecutive
// cells: an elected thread issues two 1-D bulk-tensor copies (cp.async.bulk, SASS: UBLKCP) that land
// the tile's slice of adjList and the source-field window [r0-W, r0+TILE+W) in shared memory and
// signals an mbarrier; every gather of the sweep then reads shared memory instead of issuing 32
// scattered L1 sectors per warp instruction.  Same arithmetic as SmoothFieldK (f64, reference order).
#pragma once
#include "pb_platform.h"
#include "pb_stencil.h"

#if PB_CUDA
namespace pb {

#define PB_TILE 1024
#define PB_TILE_THREADS 256

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global → shared, completion counted in bytes on the mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void tma_load_1d(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct TileGeom { int W; int adjCap; };   // window half-width (multiple of 4) and shared adjList capacity per tile (ints)

// dst = smoothField sweep of src (js/climate-util.js:11-20); optional mask semantics:
//   MODE 0 plain; MODE 1 smoothOcean-style mask (cells outside keep src / become 0, only masked neighbours count);
//   MODE 2 diffuseOceanWarmth (cells with pcont >= 0.95 keep src, all neighbours count)
template <int MODE>
__global__ void __launch_bounds__(PB_TILE_THREADS) k_sweep_tiled(Csr g, const float* __restrict__ src, float* __restrict__ dst,
                                                                   const uint8_t* __restrict__ mask, const float* __restrict__ pcont,
                                                                   int zeroOutside, TileGeom tg) {
    extern __shared__ __align__(128) unsigned char smraw[];
    uint64_t* bar = (uint64_t*)smraw;                        // 16 bytes reserved
    float* win = (float*)(smraw + 16);                       // [PB_TILE + 2W + 8]
    const int winCap = PB_TILE + 2 * tg.W + 8;
    int* sOff = (int*)(win + winCap);                        // [PB_TILE + 1]
    int* sAdj = sOff + PB_TILE + 4;                          // [adjCap + 8]   (16-byte aligned: winCap, PB_TILE+4 are multiples of 4)
    const int N = g.N;
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * PB_TILE;
    const int r1 = min(N, r0 + PB_TILE);
    const int lo = max(0, r0 - tg.W);                        // multiple of 4 (r0 and W are)
    const int hi = min(N, r1 + tg.W);
    const int a0 = g.off[r0], a1 = g.off[r1];
    const int aLo = a0 & ~3;
    const bool staged = (a1 - aLo) <= tg.adjCap;
    const int nWin = hi - lo, nWinBulk = nWin & ~3;
    const int nAdj = a1 - aLo, nAdjBulk = staged ? (nAdj & ~3) : 0;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)(4 * (nWinBulk + nAdjBulk)));
        if (nWinBulk) tma_load_1d(win, src + lo, 4u * nWinBulk, bar);
        if (nAdjBulk) tma_load_1d(sAdj, g.adj + aLo, 4u * nAdjBulk, bar);
    }
    // the (< 4 element) tails that a 16-byte granular bulk copy cannot cover, and the row offsets
    if (tid < nWin - nWinBulk) win[nWinBulk + tid] = src[lo + nWinBulk + tid];
    if (staged && tid < nAdj - nAdjBulk) sAdj[nAdjBulk + tid] = g.adj[aLo + nAdjBulk + tid];
    for (int k = tid; k <= r1 - r0; k += PB_TILE_THREADS) sOff[k] = g.off[r0 + k];
    __syncthreads();                                         // mbarrier init + plain stores visible
    mbar_wait(bar, 0);
    for (int k = tid; k < r1 - r0; k += PB_TILE_THREADS) {
        const int r = r0 + k;
        const float self = win[r - lo];
        if (MODE == 1 && !mask[r]) { dst[r] = zeroOutside ? 0.0f : self; continue; }
        if (MODE == 2 && (double)pcont[r] >= 0.95) { dst[r] = self; continue; }
        double sum = self;
        int count = 1;
        const int b = sOff[k], e = sOff[k + 1];
        for (int j = b; j < e; j++) {
            const int nb = staged ? sAdj[j - aLo] : g.adj[j];
            if (MODE == 1 && !mask[nb]) continue;
            const float v = (nb >= lo && nb < hi) ? win[nb - lo] : src[nb];     // outliers (pole ring) through L2
            sum += v;
            count++;
        }
        dst[r] = (float)(sum / count);
    }
}

}  // namespace pb
#endif
