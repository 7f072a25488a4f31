// scan.cuh -- exclusive scan and flag -> ascending label list compaction.
//
// The reference returns every decision as an ascending labelList built by scanning a boolList
// (hexRef.C:1440-1460, hexRef3D.C:1930-1950); this is the device equivalent:
//   pass 1  per-tile popcount of the byte flags (16 flags per 128-bit load),
//   pass 2  exclusive scan of the tile counts,
//   pass 3  per-tile rescan + ordered scatter of the set indices.
// Order inside the output equals label order, so the result is bit-identical to the host scan.
#pragma once
#include "common.cuh"

constexpr int kFlagsPerThread = 16;
constexpr int kTile = kThreads * kFlagsPerThread;   // 4096 flags per CTA

__device__ __forceinline__ int warpInclusiveScan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// block-wide exclusive scan of one int per thread (kThreads threads); returns exclusive prefix,
// *total = block sum.  smem: int[kThreads/32]
__device__ __forceinline__ int blockExclusiveScan(int v, int *smem, int *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = warpInclusiveScan(v);
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < (kThreads / 32) ? smem[lane] : 0;
        int winc = warpInclusiveScan(w);
        if (lane < (kThreads / 32)) smem[lane] = winc - w;   // exclusive warp offsets
        if (lane == (kThreads / 32) - 1) smem[kThreads / 32] = winc;
    }
    __syncthreads();
    int res = inc - v + smem[wid];
    *total = smem[kThreads / 32];
    __syncthreads();
    return res;
}

// 16 flags of thread `t` of the tile as 4 words with one bit per set byte (bit 0 of each byte)
__device__ __forceinline__ uint4 loadFlags16(const u8 *flags, int64_t base, int64_t N, bool aligned) {
    uint4 w = make_uint4(0, 0, 0, 0);
    if (aligned && base + 16 <= N) {
        w = *reinterpret_cast<const uint4 *>(flags + base);
    } else {
        unsigned v[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (base + k < N) v[k >> 2] |= (unsigned)flags[base + k] << (8 * (k & 3));
        w = make_uint4(v[0], v[1], v[2], v[3]);
    }
    // normalise any non-zero byte to 0x01
    auto norm = [](unsigned x) {
        x |= x >> 4; x |= x >> 2; x |= x >> 1;
        return x & 0x01010101u;
    };
    w.x = norm(w.x); w.y = norm(w.y); w.z = norm(w.z); w.w = norm(w.w);
    return w;
}

__global__ void __launch_bounds__(kThreads) k_flag_count(const u8 *__restrict__ flags, int64_t N, bool aligned,
                                                        int32_t *__restrict__ tileCnt) {
    __shared__ int smem[kThreads / 32 + 1];
    int64_t base = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kFlagsPerThread;
    int c = 0;
    if (base < N) {
        uint4 w = loadFlags16(flags, base, N, aligned);
        c = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    }
    // block reduce
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < kThreads / 32; i++) s += smem[i];
        tileCnt[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(kThreads) k_flag_scatter(const u8 *__restrict__ flags, int64_t N, bool aligned,
                                                          const int32_t *__restrict__ tileOff,
                                                          int32_t *__restrict__ out) {
    __shared__ int smem[kThreads / 32 + 1];
    int64_t base = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kFlagsPerThread;
    uint4 w = make_uint4(0, 0, 0, 0);
    if (base < N) w = loadFlags16(flags, base, N, aligned);
    int c = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    int total;
    int pos = blockExclusiveScan(c, smem, &total) + tileOff[blockIdx.x];
    if (c) {
        unsigned v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 16; k++)
            if ((v[k >> 2] >> (8 * (k & 3))) & 1u) out[pos++] = (int32_t)(base + k);
    }
}

// ---- exclusive scan of int32 (multi-level, 1024 items per CTA) ---------------------------
constexpr int kScanItems = 4;
constexpr int kScanTile = kThreads * kScanItems;

__global__ void __launch_bounds__(kThreads) k_scan_tile(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                       int64_t N, int32_t *__restrict__ tileSum) {
    __shared__ int smem[kThreads / 32 + 1];
    int64_t base = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { v[k] = (base + k < N) ? in[base + k] : 0; s += v[k]; }
    int total;
    int ex = blockExclusiveScan(s, smem, &total);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { if (base + k < N) out[base + k] = ex; ex += v[k]; }
    if (threadIdx.x == 0 && tileSum) tileSum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kThreads) k_scan_add(int32_t *__restrict__ out, int64_t N,
                                                      const int32_t *__restrict__ tileOff) {
    int64_t base = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kScanItems;
    int add = tileOff[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) if (base + k < N) out[base + k] += add;
}

// out[i] = sum_{j<i} in[j] for i in [0, N]; out has N+1 entries (out[N] = total). in/out may alias.
static int deviceExclusiveScan(amr_ctx *ctx, const int32_t *in, int32_t *out, int64_t N, int32_t *auxBase,
                               int64_t auxCap) {
    // treat the array as N+1 items with a trailing 0 so that out[N] = total
    int64_t M = N + 1;
    int64_t tiles = (M + kScanTile - 1) / kScanTile;
    if (tiles == 1) {
        LAUNCH(k_scan_tile, 1, kThreads, in, out, M, (int32_t *)nullptr);
        // the virtual element in[N] is read as in[N]; callers guarantee in has N+1 readable
        // entries whose last one is 0 (see widthBuf / tileCnt allocation)
        return AMR_OK;
    }
    if (tiles + 1 > auxCap) FAIL(AMR_ERR_STATE, "scan aux buffer too small");
    CK(cudaMemsetAsync(auxBase + tiles, 0, sizeof(int32_t), ctx->stream));
    LAUNCH(k_scan_tile, (unsigned)tiles, kThreads, in, out, M, auxBase);
    TRY(deviceExclusiveScan(ctx, auxBase, auxBase, tiles, auxBase + tiles + 1, auxCap - tiles - 1));
    LAUNCH(k_scan_add, (unsigned)tiles, kThreads, out, M, auxBase);
    return AMR_OK;
}

// flags[N] (bytes) -> ascending indices in `out`; count stored to ctx->dCount[slot] (device) via tile offsets.
// After the call tileCnt[tiles] holds the total.
static int compactFlags(amr_ctx *ctx, const u8 *flags, int64_t N, int32_t *out, int32_t **totalDev) {
    int64_t tiles = (N + kTile - 1) / kTile;
    if (tiles < 1) tiles = 1;
    if (tiles + 2 > ctx->tileCap) FAIL(AMR_ERR_STATE, "compaction tile buffer too small");
    bool aligned = (((uintptr_t)flags) & 15) == 0;
    CK(cudaMemsetAsync(ctx->tileCnt + tiles, 0, sizeof(int32_t), ctx->stream));
    LAUNCH(k_flag_count, (unsigned)tiles, kThreads, flags, N, aligned, ctx->tileCnt);
    TRY(deviceExclusiveScan(ctx, ctx->tileCnt, ctx->tileCnt, tiles, ctx->scanAux, ctx->scanAuxCap));
    if (out) LAUNCH(k_flag_scatter, (unsigned)tiles, kThreads, flags, N, aligned, ctx->tileCnt, out);
    *totalDev = ctx->tileCnt + tiles;
    return AMR_OK;
}
