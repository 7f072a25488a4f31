// pipeline.cuh -- persistent, warp-specialised SELL tile pipeline for sm_100a.
//
// The adjacency (SELL-32 column blocks) is the one large, perfectly sequential stream of every
// gather kernel on this path (24 B/cell on a hex mesh against 8-16 B/cell of field data), and it is
// what the per-cell gathers depend on.  Loading it with ordinary LDGs puts two DRAM latencies in
// series in every thread (labels -> values).  Here a producer warp streams whole tiles
// (8 slices = 256 rows, contiguous in memory) into a shared-memory ring with the TMA bulk-copy
// engine (cp.async.bulk, SASS UBLKCP) several tiles ahead, signalled through mbarriers; the 8
// consumer warps read their labels from shared memory (column-major: conflict-free) and keep only
// the value gathers in flight.  One CTA per resident slot (persistent grid), tiles round-robin.
#pragma once
#include "common.cuh"

constexpr int kTileSlices = 8;                       // slices per tile
constexpr int kTileRows = kTileSlices * kSlice;      // 256 rows per tile
constexpr int kPipeThreads = kTileRows + 32;         // 8 consumer warps + 1 producer warp
constexpr int kStages = 3;                           // ring depth at most; SellView::stages (2 or 3) is what a launch uses
constexpr int kHdrInts = 16;                         // slice offsets of a tile: 9 used, fetched from the 16-byte aligned address below the
                                                     // tile's first slice (a TMA bulk copy needs 16-byte alignment; tiles start anywhere)

__device__ __forceinline__ unsigned smemAddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smemAddr(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.  The adjacency is a pure stream
// (every tile is read once per kernel): it goes through L2 with an evict-first policy so that it does not push out the
// field values / flags that the gathers of neighbouring tiles are about to re-use.
__device__ __forceinline__ uint64_t evictFirstPolicy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulkLoad(void *dstSmem, const void *srcGlobal, unsigned bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smemAddr(dstSmem)),
                 "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)), "l"(policy)
                 : "memory");
}

struct SellView {
    const int32_t *sliceOff;   // [slices + 16], tail padded
    const int32_t *col;
    int64_t rows, slices;
    const int32_t *tileFirst;  // [tiles + 1] first slice of every tile (<= 8 slices, columns capped: sell.cuh)
    int64_t tiles;
    int stageInts;             // capacity of one ring stage in int32 (>= largest tile)
    int stages;                // ring depth of this launch: 3, or 2 when three stages of this mesh's largest tile would cost resident CTAs
    const u8 *sliceGhost;      // [slices] slice holds a cell that owns a coupled face, or nullptr
};

// dynamic shared memory layout: [kStages][stageInts] labels | [kStages][kHdrInts] headers | barriers
__host__ __device__ inline size_t pipeSmemBytes(int stageInts, int stages = kStages) {
    return (size_t)stages * stageInts * 4 + (size_t)kStages * kHdrInts * 4 + 2 * kStages * 8 + 64;
}

// Body signature: void body(int64_t row, bool valid, const int32_t* rowLabels /*smem, stride 32*/, int width, bool ghostSlice)
// called by every consumer thread once per tile (valid=false for rows >= S.rows; all 32 lanes of a
// warp call it together so that warp collectives inside the body are legal).
template <class Body>
__device__ __forceinline__ void sellPipeline(const SellView &S, Body body) {
    extern __shared__ __align__(128) unsigned char smemRaw[];
    int32_t *ring = reinterpret_cast<int32_t *>(smemRaw);
    const int nStages = S.stages;
    int32_t *hdr = ring + (size_t)nStages * S.stageInts;
    uint64_t *full = reinterpret_cast<uint64_t *>(hdr + kStages * kHdrInts);
    uint64_t *empty = full + kStages;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < nStages; s++) { mbarInit(&full[s], 1); mbarInit(&empty[s], kTileSlices); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t ntiles = S.tiles;
    if (warp == kTileSlices) {
        // ---------------- producer warp (one elected lane drives the TMA engine)
        if (lane == 0) {
            const uint64_t pol = evictFirstPolicy();
            int stage = 0;
            unsigned phase = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                mbarWait(&empty[stage], phase ^ 1u);
                const int64_t s0 = S.tileFirst[t];
                const int64_t s1 = S.tileFirst[t + 1];
                const int32_t b0 = S.sliceOff[s0], b1 = S.sliceOff[s1];
                const unsigned bytes = (unsigned)(b1 - b0) * 128u;
                mbarExpectTx(&full[stage], bytes + kHdrInts * 4);
                bulkLoad(hdr + stage * kHdrInts, S.sliceOff + (s0 & ~(int64_t)3), kHdrInts * 4, &full[stage], pol);
                if (bytes) bulkLoad(ring + (size_t)stage * S.stageInts, S.col + ((size_t)b0 << 5), bytes, &full[stage], pol);
                if (++stage == nStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ---------------- consumer warps
        int stage = 0;
        unsigned phase = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            // tile bounds (and the per-slice flag) fetched BEFORE the wait so that their latency hides behind the tile's arrival
            const int64_t t0s = S.tileFirst[t], t1s = S.tileFirst[t + 1];
            const int64_t s = t0s + warp;
            const bool sliceValid = s < t1s;
            const unsigned ghostSlice = (S.sliceGhost && sliceValid) ? S.sliceGhost[s] : 0u;
            mbarWait(&full[stage], phase);
            const int32_t *h = hdr + stage * kHdrInts + (int)(t0s & 3);
            const int64_t row = s * 32 + lane;
            int32_t base = 0, w = 0;
            if (sliceValid) { base = h[warp] - h[0]; w = h[warp + 1] - h[warp]; }
            body(row, sliceValid && row < S.rows, ring + (size_t)stage * S.stageInts + ((size_t)base << 5) + lane, w, ghostSlice != 0);
            __syncwarp();
            if (lane == 0) mbarArrive(&empty[stage]);
            if (++stage == nStages) { stage = 0; phase ^= 1u; }
        }
    }
}
