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

// ---- single-pass compaction (decoupled look-back) ---------------------------------------------------------
// One launch turns byte flags into the ascending list: every CTA takes the next tile (dynamic ticket, so that all
// predecessors of a tile are already running), counts its set flags, publishes the count, looks back over the
// predecessors' published counts / inclusive prefixes -- 32 at a time, one per lane of warp 0 -- until it knows its
// own offset, and scatters.  Status words carry a per-call tag, so the status array is never cleared.
// flag = 1: aggregate, 2: inclusive prefix.
//   src (optional)  : entry i stands for element src[i]: its flag is flags[src[i]] (compaction of a list by a flag array)
//   map (optional)  : the output is map[element] instead of the element
//   nDev (optional) : N is read from device memory (*nDev, <= N)
constexpr int kLbGroups = 4;                                   // 16-flag groups per thread
constexpr int kLbTile = kThreads * 16 * kLbGroups;             // 16384 flags per CTA
__device__ __forceinline__ unsigned long long packStatus(unsigned tag, unsigned flag, unsigned value) {
    return ((unsigned long long)((tag << 2) | flag) << 32) | value;
}
// mail (optional): the tile that learns the total also posts the results of the speculative loop that ends with this
// compaction -- queue sizes, list size, by-product counts, peer error word -- to the pinned mailbox the host polls
// (Mailbox, common.cuh): no separate mailbox launch.  The host reads nothing but the mailbox before the stream has
// drained, and everything that follows is stream-ordered, so posting before the other tiles have scattered is safe.
struct MailArgs {
    Mailbox *mb;
    const int *gq;
    int nq;
    const unsigned long long *marked;
    const int *peerError;
    unsigned seq;
};
__global__ void __launch_bounds__(kThreads)
k_compact_lb(const u8 *__restrict__ flags, int64_t N, const int32_t *__restrict__ nDev, bool aligned, const label *__restrict__ src,
             const label *__restrict__ map, unsigned long long *status, unsigned *ticket, unsigned tag, int32_t *__restrict__ out,
             int32_t *__restrict__ total, const int *__restrict__ gate, const MailArgs mail) {
    if (gate && *gate == 0) return;
    __shared__ int smem[kThreads / 32 + 1];
    __shared__ unsigned sTile;
    __shared__ int sPrefix;
    if (nDev) N = min(N, (int64_t)*nDev);
    const unsigned tiles = gridDim.x;
    if (threadIdx.x == 0) sTile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = sTile;
    // thread t owns groups t, t + 256, ... of the tile (coalesced 128-bit loads); output order = group order
    unsigned bits[kLbGroups];
    int c = 0;
#pragma unroll
    for (int g = 0; g < kLbGroups; g++) {
        const int64_t base = (int64_t)tile * kLbTile + ((int64_t)g * kThreads + threadIdx.x) * 16;
        unsigned b = 0;
        if (base < N) {
            if (!src) {
                const uint4 w = loadFlags16(flags, base, N, aligned);
                const unsigned v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int k = 0; k < 16; k++) b |= ((v[k >> 2] >> (8 * (k & 3))) & 1u) << k;
            } else if (base + 16 <= N) {
                int4 s4[4];                                    // 16 elements: 4 vector loads, then 16 independent flag gathers
#pragma unroll
                for (int k = 0; k < 4; k++) s4[k] = reinterpret_cast<const int4 *>(src + base)[k];
                const int e[16] = {s4[0].x, s4[0].y, s4[0].z, s4[0].w, s4[1].x, s4[1].y, s4[1].z, s4[1].w,
                                   s4[2].x, s4[2].y, s4[2].z, s4[2].w, s4[3].x, s4[3].y, s4[3].z, s4[3].w};
                u8 f[16];
#pragma unroll
                for (int k = 0; k < 16; k++) f[k] = flags[e[k]];
#pragma unroll
                for (int k = 0; k < 16; k++) b |= (f[k] ? 1u : 0u) << k;
            } else {
                for (int k = 0; k < 16 && base + k < N; k++) b |= (flags[src[base + k]] ? 1u : 0u) << k;
            }
        }
        bits[g] = b;
        c += __popc(b);
    }
    // per-group block scans: group g of all threads precedes group g + 1
    int pos[kLbGroups];
    int tileTotal = 0;
#pragma unroll
    for (int g = 0; g < kLbGroups; g++) {
        int gt;
        pos[g] = blockExclusiveScan(__popc(bits[g]), smem, &gt) + tileTotal;
        tileTotal += gt;
    }
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int prefix = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&status[0], packStatus(tag, 2u, (unsigned)tileTotal));
        } else {
            if (lane == 0) atomicExch(&status[tile], packStatus(tag, 1u, (unsigned)tileTotal));
            int j = (int)tile - 1 - lane;
            while (true) {
                unsigned long long sv = packStatus(tag, 2u, 0u);             // before tile 0: an inclusive prefix of zero
                if (j >= 0) {
                    do { sv = *reinterpret_cast<volatile unsigned long long *>(&status[j]); } while ((unsigned)(sv >> 34) != tag || ((sv >> 32) & 3u) == 0u);
                }
                const unsigned inclMask = __ballot_sync(0xffffffffu, ((sv >> 32) & 3u) == 2u);
                int v = (int)(unsigned)sv;
                if (inclMask) {
                    const int first = __ffs(inclMask) - 1;                   // nearest predecessor that knows its inclusive prefix
                    if (lane > first) v = 0;
                }
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                prefix += v;
                if (inclMask) break;
                j -= 32;
            }
            if (lane == 0) atomicExch(&status[tile], packStatus(tag, 2u, (unsigned)(prefix + tileTotal)));
        }
        if (lane == 0) {
            sPrefix = prefix;
            if (tile == tiles - 1) {                   // every ticket has been drawn: reset for the next call
                *ticket = 0;
                if (total) *total = prefix + tileTotal;
            }
        }
        if (tile == tiles - 1 && mail.mb) {
            Mailbox *mb = mail.mb;
            for (int t = lane; t < mail.nq; t += 32) mb->q[t] = mail.gq[t];
            if (lane < 16) mb->marked[lane] = (long long)mail.marked[lane];
            if (lane == 0) { mb->total = prefix + tileTotal; mb->peerError = mail.peerError ? *mail.peerError : 0; }
            __threadfence_system();
            __syncwarp();
            if (lane == 0) mb->seq = mail.seq;
        }
    }
    __syncthreads();
    if (out) {
#pragma unroll
        for (int g = 0; g < kLbGroups; g++) {
            const int64_t base = (int64_t)tile * kLbTile + ((int64_t)g * kThreads + threadIdx.x) * 16;
            unsigned b = bits[g];
            int p = pos[g] + sPrefix;
            if (src && b && base + 16 <= N) {
                // list form: the 16 elements again (4 vector loads, L1/L2 hits), then all map gathers of the set entries
                // at once -- one round of independent loads instead of a dependent src -> map chain per set entry
                int4 s4[4];
#pragma unroll
                for (int k = 0; k < 4; k++) s4[k] = reinterpret_cast<const int4 *>(src + base)[k];
                int e[16] = {s4[0].x, s4[0].y, s4[0].z, s4[0].w, s4[1].x, s4[1].y, s4[1].z, s4[1].w,
                             s4[2].x, s4[2].y, s4[2].z, s4[2].w, s4[3].x, s4[3].y, s4[3].z, s4[3].w};
                if (map) {
#pragma unroll
                    for (int k = 0; k < 16; k++) if ((b >> k) & 1u) e[k] = map[e[k]];
                }
#pragma unroll
                for (int k = 0; k < 16; k++) if ((b >> k) & 1u) out[p++] = e[k];
                continue;
            }
            while (b) {
                const int k = __ffs(b) - 1;
                b &= b - 1;
                const int64_t i = src ? (int64_t)src[base + k] : base + k;
                out[p++] = map ? map[i] : (int32_t)i;
            }
        }
    }
}

// flags[N] (bytes, any non-zero value counts) -> ascending indices in `out`; the count is stored to *totalDev (device).
static int compactFlagsEx(amr_ctx *ctx, const u8 *flags, int64_t N, const int32_t *nDev, const label *src, const label *map,
                          int32_t *out, int32_t **totalDev, int totalSlot, const int *gate, const MailArgs *mail = nullptr) {
    int64_t tiles = (N + kLbTile - 1) / kLbTile;
    if (tiles < 1) tiles = 1;
    if (tiles + 2 > ctx->tileCap) FAIL(AMR_ERR_STATE, "compaction tile buffer too small");
    const bool aligned = (((uintptr_t)flags) & 15) == 0;
    ctx->scanTag = (ctx->scanTag + 1u) & 0x3fffffffu;
    if (ctx->scanTag == 0) ctx->scanTag = 1;
    int32_t *total = ctx->dTotals + totalSlot;
    MailArgs noMail;
    memset(&noMail, 0, sizeof(noMail));
    k_compact_lb<<<(unsigned)tiles, kThreads, 0, ctx->stream>>>(flags, N, nDev, aligned, src, map, ctx->lbStatus, ctx->lbTicket, ctx->scanTag,
                                                                 out, total, gate, mail ? *mail : noMail);
    ctx->launches++;
    traceMark(ctx, "k_compact_lb");
    *totalDev = total;
    return AMR_OK;
}
static int compactFlags(amr_ctx *ctx, const u8 *flags, int64_t N, int32_t *out, int32_t **totalDev) {
    return compactFlagsEx(ctx, flags, N, nullptr, nullptr, nullptr, out, totalDev, 0, nullptr);
}
