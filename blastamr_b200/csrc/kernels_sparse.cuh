// kernels_sparse.cuh -- the marking layers in place, and the two closure loops as work-queue (frontier) iterations.
//
// (1) Buffer layers in place.  The marking is a byte per cell; a cell first marked by layer k of a call sequence
//     holds the generation k+1 (the candidates hold 1).  Layer k treats 0 < v <= gen as sources and writes gen+1,
//     so one kernel can read and write the same array: a mark written by this layer is never taken for a source of
//     this layer.  No ping-pong buffer, no copy; every reader of a marking tests "non-zero".  Two forms with identical
//     results: a dense pull (every unmarked cell gathers over its row; TMA-staged) and a sparse push (the rows of the
//     source cells only), chosen by the host from the source count the same call site reported last time.
// (2) hexRef::consistentRefinement (hexRef.C:1402-1468), maxSet = true, on a 2:1-balanced mesh: only a MARKED cell can
//     stand two levels above a face neighbour, and once a pair (marked c, neighbour q) has been looked at it never
//     needs another look (levels are fixed, marks only grow).  So: sweep 0 walks the rows of all marked cells and marks
//     the neighbours the rule forces; every newly marked cell goes to a queue; iteration j walks the rows of queue j
//     only.  The closure is reached when the queue stays empty -- no confirming sweep over the mesh.
// (3) hexRef3D::consistentUnrefinement (hexRef3D.C:1719-1953) on a mesh-constant split-point table (the 8 cells of
//     every split point, each cell belonging to at most one): sweep 0 examines every cell marked for unrefinement; a
//     cell that would end up more than one level below a neighbour knocks out its split point at once, which unmarks
//     the point's 8 cells; the knocked points go to a queue, and iteration j looks only at the marked neighbours of the
//     cells that queue j has just reverted.  Flags only drop and every drop is forced, so this chaotic order reaches
//     the reference's (greatest) fixed point.
// All queue appends are made exactly once per element (atomicOr / atomicAnd on the word that holds the flag byte).
#pragma once
#include "kernels.cuh"
#include "kernels_pipe.cuh"
#include "p2p.cuh"

// ---- byte flags inside 32-bit words: exact "I was first" detection without byte atomics
__device__ __forceinline__ bool setByteOnce(u8 *flags, int64_t i) {          // 0 -> 1; true if this thread did it
    unsigned *w = reinterpret_cast<unsigned *>(flags + (i & ~(int64_t)3));
    const unsigned sh = 8u * (unsigned)(i & 3);
    const unsigned old = atomicOr(w, 1u << sh);
    return ((old >> sh) & 0xffu) == 0u;
}
__device__ __forceinline__ bool clearByteOnce(u8 *flags, int64_t i) {        // non-zero -> 0; true if this thread did it
    unsigned *w = reinterpret_cast<unsigned *>(flags + (i & ~(int64_t)3));
    const unsigned sh = 8u * (unsigned)(i & 3);
    const unsigned old = atomicAnd(w, ~(0xffu << sh));
    return ((old >> sh) & 0xffu) != 0u;
}

// block-level count -> one atomic per CTA
__device__ __forceinline__ void blockAddCount(int cnt, unsigned long long *total) {
    __shared__ int part[kThreads / 32 + 2];
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += part[i];
        if (t) atomicAdd(total, (unsigned long long)t);
    }
}

// ------------------------------------------------------------------------------------------
// (1) buffer layers in place
// ------------------------------------------------------------------------------------------

// source test of a layer: separate source array (the forced first layer) or generation range of the marking itself
__device__ __forceinline__ bool isSource(unsigned v, int gen, bool separate) { return separate ? v != 0u : (v - 1u) < (unsigned)gen; }

// sparse push form: rows of the source cells; marks written are gen+1.  nSrc receives the exact number of sources.
__global__ void __launch_bounds__(kThreads)
k_layer_push(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col, const u8 *srcArr, bool separate,
             int gen, bool aligned, u8 *cur, unsigned long long *__restrict__ nSrc) {
    __shared__ unsigned short setIdx[kThreads / 32][512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warpBase = ((int64_t)blockIdx.x * kThreads + (threadIdx.x & ~31)) * 16;
    int mine = 0;
    if (warpBase < n) {
        const int64_t base = warpBase + (int64_t)lane * 16;
        unsigned mask = 0;
        if (base < n) {
            if (aligned && base + 16 <= n) {
                const uint4 w = *reinterpret_cast<const uint4 *>(srcArr + base);
                const unsigned v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int k = 0; k < 16; k++) mask |= (isSource((v[k >> 2] >> (8 * (k & 3))) & 0xffu, gen, separate) ? 1u : 0u) << k;
            } else {
                for (int k = 0; k < 16 && base + k < n; k++) mask |= (isSource(srcArr[base + k], gen, separate) ? 1u : 0u) << k;
            }
        }
        const int cnt = __popc(mask);
        mine = cnt;
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
            int o = incl - cnt;
            while (mask) {
                const int k = __ffs(mask) - 1;
                mask &= mask - 1;
                setIdx[warp][o++] = (unsigned short)(lane * 16 + k);
            }
            __syncwarp();
            const u8 tag = (u8)(gen + 1);
            for (int i = lane; i < total; i += 32) {
                const int64_t c = warpBase + setIdx[warp][i];
                const int32_t b0 = sliceOff[c >> 5];
                const int w = sliceOff[(c >> 5) + 1] - b0;
                const int32_t *row = col + ((size_t)b0 << 5) + (c & 31);
                for (int j0 = 0; j0 < w; j0 += 8) {
                    int32_t q[8];
                    u8 v[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) q[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : (int32_t)c;
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = cur[q[k]];
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        if (v[k] == 0 && q[k] != (int32_t)c) cur[q[k]] = tag;      // all writers store the same tag: benign race
                }
            }
            __syncwarp();
        }
    }
    if (nSrc) blockAddCount(mine, nSrc);
}

// sparse pull form: when MOST cells are marked already (the unrefinement layer on a graded mesh: the marking covers the
// whole refined band), only the rows of the UNMARKED cells matter -- a 16-flags-per-load scan finds them, the warp deals
// them to its lanes, each gathers the source values of its row.  Same result as the other two forms; reads the
// adjacency of the unmarked cells only instead of streaming all of it.
// The gathers of that walk read ONE flag per neighbour: as bytes they cost a 32-byte DRAM sector each (the marking of a
// 72 M-cell mesh does not stay in L2 next to the adjacency stream); as bits the whole source mask is n/8 bytes (9 MB) and
// lives in L2.  k_pack_source_bits builds the mask in one streaming pass (and counts the sources on the way).
__global__ void __launch_bounds__(kThreads)
k_pack_source_bits(int64_t n, const u8 *__restrict__ srcArr, bool separate, int gen, bool aligned, unsigned *__restrict__ bits,
                   unsigned long long *__restrict__ nSrc) {
    const int64_t word = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t base = word * 32;
    unsigned m = 0;
    if (base < n) {
        if (aligned && base + 32 <= n) {
            const uint4 a = *reinterpret_cast<const uint4 *>(srcArr + base), b = *reinterpret_cast<const uint4 *>(srcArr + base + 16);
            const unsigned v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 32; k++) m |= (isSource((v[k >> 2] >> (8 * (k & 3))) & 0xffu, gen, separate) ? 1u : 0u) << k;
        } else {
            for (int k = 0; k < 32 && base + k < n; k++) m |= (isSource(srcArr[base + k], gen, separate) ? 1u : 0u) << k;
        }
        bits[word] = m;
    }
    if (nSrc) blockAddCount(__popc(m), nSrc);
}

__global__ void __launch_bounds__(kThreads)
k_layer_pull_sparse(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col, const unsigned *__restrict__ srcBits,
                    int gen, bool aligned, u8 *cur) {
    __shared__ unsigned short unsetIdx[kThreads / 32][512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warpBase = ((int64_t)blockIdx.x * kThreads + (threadIdx.x & ~31)) * 16;
    if (warpBase >= n) return;
    const int64_t base = warpBase + (int64_t)lane * 16;
    unsigned zmask = 0;                                   // unmarked cells of this thread's 16
    if (base < n) {
        if (aligned && base + 16 <= n) {
            const uint4 w = *reinterpret_cast<const uint4 *>(cur + base);
            const unsigned v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned z = ~__vcmpne4(v[k], 0u);         // 0xff per zero byte
                zmask |= (((z >> 0) & 1u) | ((z >> 7) & 2u) | ((z >> 14) & 4u) | ((z >> 21) & 8u)) << (4 * k);
            }
        } else {
            for (int k = 0; k < 16 && base + k < n; k++) zmask |= (cur[base + k] == 0 ? 1u : 0u) << k;
        }
    }
    const int cnt = __popc(zmask);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    int o = incl - cnt;
    while (zmask) {
        const int k = __ffs(zmask) - 1;
        zmask &= zmask - 1;
        unsetIdx[warp][o++] = (unsigned short)(lane * 16 + k);
    }
    __syncwarp();
    const u8 tag = (u8)(gen + 1);
    for (int i = lane; i < total; i += 32) {
        const int64_t c = warpBase + unsetIdx[warp][i];
        const int32_t b0 = sliceOff[c >> 5];
        const int w = sliceOff[(c >> 5) + 1] - b0;
        const int32_t *row = col + ((size_t)b0 << 5) + (c & 31);
        unsigned hit = 0;
        for (int j0 = 0; j0 < w; j0 += 8) {
            int32_t q[8];
            unsigned v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) q[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : (int32_t)c;
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = srcBits[q[k] >> 5];
#pragma unroll
            // padding reads the cell's own bit: an unmarked cell is no source (a source is always marked: the candidates
            // kernel writes both arrays)
            for (int k = 0; k < 8; k++) hit |= (v[k] >> (q[k] & 31)) & 1u;
        }
        if (hit) cur[c] = tag;
    }
    __syncwarp();
}

// dense pull form (TMA-staged): every unmarked cell gathers the source values of its row
__global__ void __launch_bounds__(kPipeThreads, 7)
kp_layer_pull(SellView S, const u8 *srcArr, bool separate, int gen, u8 *cur, unsigned long long *__restrict__ nSrc) {
    const int32_t n = (int32_t)S.rows;
    const u8 tag = (u8)(gen + 1);
    int mineCnt = 0;
    sellPipeline(S, [&](int64_t c64, bool valid, const int32_t *row, int w, bool) {
        const int32_t c = valid ? (int32_t)c64 : n - 1;
        const unsigned own = cur[c];
        const unsigned ownSrc = separate ? (unsigned)srcArr[c] : own;
        if (valid && isSource(ownSrc, gen, separate)) mineCnt++;
        unsigned hit = 0;
        // a warp whose 32 cells are all marked has nothing to gather
        if (!__all_sync(0xffffffffu, own != 0)) {
            sellRow(row, w, c, [&](auto W, const int32_t *q) {
                constexpr int NW = decltype(W)::value;
                unsigned v[NW];
#pragma unroll
                for (int k = 0; k < NW; k++) v[k] = srcArr[q[k]];
#pragma unroll
                // padding reads the cell's own value: a marked cell is never written, an unmarked one is no source
                for (int k = 0; k < NW; k++) hit |= isSource(v[k], gen, separate) ? 1u : 0u;
            });
        }
        if (valid && own == 0 && hit) cur[c] = tag;
    });
    if (nSrc && threadIdx.x < kTileRows) {
        for (int d = 16; d > 0; d >>= 1) mineCnt += __shfl_xor_sync(0xffffffffu, mineCnt, d);
        if ((threadIdx.x & 31) == 0 && mineCnt) atomicAdd(nSrc, (unsigned long long)mineCnt);
    }
}
// direct variant (tiles too large for the ring / AMR_B200_NO_PIPELINE)
__global__ void __launch_bounds__(kThreads)
k_layer_pull(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col, const u8 *srcArr, bool separate,
             int gen, u8 *cur, unsigned long long *__restrict__ nSrc) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int mineCnt = 0;
    if (c < n) {
        const unsigned own = cur[c];
        if (isSource(separate ? (unsigned)srcArr[c] : own, gen, separate)) mineCnt = 1;
        if (own == 0) {
            const int32_t b0 = sliceOff[c >> 5];
            const int w = sliceOff[(c >> 5) + 1] - b0;
            const int32_t *row = col + ((size_t)b0 << 5) + (c & 31);
            unsigned hit = 0;
            for (int j = 0; j < w; j++) {
                const int32_t q = row[(size_t)j << 5];
                if (q != (int32_t)c && isSource(srcArr[q], gen, separate)) hit = 1;
            }
            if (hit) cur[c] = (u8)(gen + 1);
        }
    }
    if (nSrc) blockAddCount(mineCnt, nSrc);
}
// syncFaceList(orEqOp) part of a layer: the source value of the cell on the other side of a processor face, fused with
// the receive side of the swap (GhostLink, p2p.cuh): no unpack launch
__global__ void __launch_bounds__(kThreads)
k_fix_layer_link(const __grid_constant__ GhostLink G, int64_t nC, const label *__restrict__ cplFace, const label *__restrict__ bOwner,
                 bool separate, int gen, u8 *cur) {
    const u8 *ghostSrc = reinterpret_cast<const u8 *>(ghostAcquire(G));
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ghostSrc && i < nC) {
        const label bf = cplFace[i];
        if (isSource(ghostSrc[bf], gen, separate)) {
            const label o = bOwner[bf];
            if (cur[o] == 0) cur[o] = (u8)(gen + 1);
        }
    }
    ghostRelease(G);
}
// marking -> boolList (0/1) for the caller; also the import direction
__global__ void __launch_bounds__(kThreads)
k_flags_normalize(int64_t n, const u8 *in, u8 *out, bool aligned) {               // used in place as well: no __restrict__
    const int64_t base = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 16;
    if (base >= n) return;
    if (aligned && base + 16 <= n) {
        uint4 w = *reinterpret_cast<const uint4 *>(in + base);
        w.x = __vcmpne4(w.x, 0u) & 0x01010101u; w.y = __vcmpne4(w.y, 0u) & 0x01010101u;
        w.z = __vcmpne4(w.z, 0u) & 0x01010101u; w.w = __vcmpne4(w.w, 0u) & 0x01010101u;
        *reinterpret_cast<uint4 *>(out + base) = w;
    } else {
        for (int k = 0; k < 16 && base + k < n; k++) out[base + k] = in[base + k] ? 1 : 0;
    }
}

// the results of a speculative loop go straight to pinned host memory (Mailbox, common.cuh); the host polls seq
__global__ void k_mailbox(const int *__restrict__ gq, int nq, const int32_t *__restrict__ total, const unsigned long long *__restrict__ marked,
                          const int *__restrict__ peerError, Mailbox *mb, unsigned seq) {
    const int t = threadIdx.x;
    if (t < nq) mb->q[t] = gq[t];
    if (t < 16) mb->marked[t] = (long long)marked[t];
    if (t == 0) { mb->total = total ? *total : 0; mb->peerError = peerError ? *peerError : 0; }
    __threadfence_system();
    __syncthreads();
    if (t == 0) mb->seq = seq;
}

// ------------------------------------------------------------------------------------------
// (2) consistentRefinement as queue iterations
// ------------------------------------------------------------------------------------------

// one marked cell c: mark every unmarked face neighbour that would end up more than one level coarser
__device__ __forceinline__ void refinePushRow(int64_t c, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                                              const u8 *__restrict__ lvl8, u8 *lf, u8 *set, label *qOut, int *qOutCount, int64_t qCap) {
    const int32_t b0 = sliceOff[c >> 5];
    const int w = sliceOff[(c >> 5) + 1] - b0;
    const int32_t *row = col + ((size_t)b0 << 5) + (c & 31);
    const int mine = (int)lvl8[c] + 1;
    for (int j0 = 0; j0 < w; j0 += 8) {
        int32_t q[8];
        int lq[8];
#pragma unroll
        for (int k = 0; k < 8; k++) q[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : (int32_t)c;
#pragma unroll
        for (int k = 0; k < 8; k++) lq[k] = lvl8[q[k]];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (mine > lq[k] + 1) {                                 // padding (q == c) fails this test
                if (setByteOnce(set, q[k])) {
                    lf[q[k]] = (u8)(lq[k] + 1);
                    const int pos = atomicAdd(qOutCount, 1);
                    if (pos < qCap) qOut[pos] = q[k];
                }
            }
        }
    }
}
// sweep 0: all marked cells (16 flags per 128-bit load, set cells dealt to the lanes of the warp)
__global__ void __launch_bounds__(kThreads)
k_refine_push0(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col, const u8 *__restrict__ lvl8,
               u8 *lf, u8 *set, bool aligned, label *qOut, int *qOutCount, int64_t qCap) {
    forEachSetWarp(set, n, aligned, [&](int64_t c) { refinePushRow(c, sliceOff, col, lvl8, lf, set, qOut, qOutCount, qCap); });
}
// iteration j >= 1: the cells queue j holds.  gate: all-reduced size of queue j (multi-rank) or qInCount itself.
__global__ void __launch_bounds__(kThreads)
k_refine_frontier(const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col, const u8 *__restrict__ lvl8, u8 *lf, u8 *set,
                  const label *__restrict__ qIn, const int *__restrict__ qInCount, label *qOut, int *qOutCount, int64_t qCap,
                  const int *__restrict__ gate) {
    if (*gate == 0) return;
    const int cnt = min(*qInCount, (int)min(qCap, (int64_t)0x7fffffff));
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < cnt; i += gridDim.x * kThreads)
        refinePushRow(qIn[i], sliceOff, col, lvl8, lf, set, qOut, qOutCount, qCap);
}
// coupled faces of an iteration (hexRef.C:758-780, maxSet = true): ghostLf = level + flag of the cell across the face;
// fused with the receive side of the swap and with the all-reduce of the queue size (sum over ranks of
// *qOutCount -> *gqOut): one launch instead of unpack + fix-up + reduction
__global__ void __launch_bounds__(kThreads)
k_fix_refine_queue_link(const __grid_constant__ GhostLink G, int64_t nC, const label *__restrict__ cplFace, const label *__restrict__ bOwner,
                        const u8 *__restrict__ lvl8, u8 *lf, u8 *set, label *qOut, int *qOutCount, int64_t qCap, int *gqOut,
                        const int *__restrict__ gate) {
    if (gate && *gate == 0) return;
    const u8 *ghostLf = reinterpret_cast<const u8 *>(ghostAcquire(G));
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ghostLf && i < nC) {
        const label bf = cplFace[i];
        const label o = bOwner[bf];
        if ((int)ghostLf[bf] > (int)lvl8[o] + 1 && !set[o]) {
            if (setByteOnce(set, o)) {
                lf[o] = (u8)(lvl8[o] + 1);
                const int pos = atomicAdd(qOutCount, 1);
                if (pos < qCap) qOut[pos] = o;
            }
        }
    }
    if (ghostRelease(G) && G.peer) ctaAllReduceSum(G, qOutCount, gqOut);
}

// ------------------------------------------------------------------------------------------
// (3) consistentUnrefinement on the split-point table
// ------------------------------------------------------------------------------------------

// table from the split-point list: spCells[k * nSp + i] = k-th cell of split point i; cellSp[c] = i
__global__ void __launch_bounds__(kThreads)
k_sp_table(int64_t nSp, const label *__restrict__ spPoint, const int32_t *__restrict__ pcOff, const int32_t *__restrict__ pcCol,
           label *__restrict__ spCells, label *__restrict__ cellSp) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= nSp) return;
    const int64_t pt = spPoint[i];
    const int32_t *row = pcCol + ((size_t)pcOff[pt >> 5] << 5) + (pt & 31);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const label c = row[(size_t)k << 5];
        spCells[(size_t)k * nSp + i] = c;
        cellSp[c] = (label)i;
    }
}
// hexRef3D::selectUnrefineElems filter (hexRef3D.C:2336-2361) + maxCellField (fvMeshRefiner.C:110-125), cell-centric:
// a split point is dropped when one of its cells is marked or has error >= unrefineLevel (max < level <=> all < level).
// spFlag pre-set to 1.
__global__ void __launch_bounds__(kThreads)
k_sp_filter(int64_t n, const label *__restrict__ cellSp, const double *__restrict__ err, double unrefineLevel,
            const u8 *__restrict__ marked, u8 *__restrict__ spFlag) {
    const int64_t c0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    if (c0 >= n) return;
    if (c0 + 4 <= n) {
        const int4 sp = *reinterpret_cast<const int4 *>(cellSp + c0);
        const uchar4 mk = *reinterpret_cast<const uchar4 *>(marked + c0);
        const int s[4] = {sp.x, sp.y, sp.z, sp.w};
        const u8 m[4] = {mk.x, mk.y, mk.z, mk.w};
        if ((s[0] & s[1] & s[2] & s[3]) == -1) return;
        const double2 e01 = *reinterpret_cast<const double2 *>(err + c0);
        const double2 e23 = *reinterpret_cast<const double2 *>(err + c0 + 2);
        const double e[4] = {e01.x, e01.y, e23.x, e23.y};
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (s[k] >= 0 && (m[k] || !(e[k] < unrefineLevel))) spFlag[s[k]] = 0;
    } else {
        for (int64_t c = c0; c < n; c++) {
            const label s = cellSp[c];
            if (s >= 0 && (marked[c] || !(err[c] < unrefineLevel))) spFlag[s] = 0;
        }
    }
}
// unrefineCell = union of the cells of the listed split points (hexRef3D.C:1763-1776); uc pre-zeroed
__global__ void __launch_bounds__(kThreads)
k_sp_to_cells(const label *__restrict__ spList, const int32_t *__restrict__ nList, int64_t nSp, const label *__restrict__ spCells,
              u8 *__restrict__ uc) {
    const int cnt = *nList;
    for (int t = blockIdx.x * kThreads + threadIdx.x; t < cnt * 8; t += gridDim.x * kThreads) {
        const int i = spList[t >> 3];
        uc[spCells[(size_t)(t & 7) * nSp + i]] = 1;
    }
}
// knock split point i out: flag, its 8 cells, queue
__device__ __forceinline__ void knockPoint(label i, int64_t nSp, const label *__restrict__ spCells, u8 *spFlag, u8 *uc,
                                           label *qOut, int *qOutCount, int64_t qCap) {
    if (clearByteOnce(spFlag, i)) {
#pragma unroll
        for (int k = 0; k < 8; k++) uc[spCells[(size_t)k * nSp + i]] = 0;
        const int pos = atomicAdd(qOutCount, 1);
        if (pos < qCap) qOut[pos] = i;
    }
}
// sweep 0: every cell of every listed split point (one thread per cell): marked cell c (final level lvl-1) against the
// final levels lvl[q] - uc[q] of its face neighbours (hexRef3D.C:1785-1839)
__global__ void __launch_bounds__(kThreads)
k_unref_sweep0(const label *__restrict__ spList, const int32_t *__restrict__ nList, int64_t nSp, const label *__restrict__ spCells,
               const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col, const u8 *__restrict__ lvl8, u8 *spFlag, u8 *uc,
               label *qOut, int *qOutCount, int64_t qCap) {
    const int cnt = *nList;
    for (int t = blockIdx.x * kThreads + threadIdx.x; t < cnt * 8; t += gridDim.x * kThreads) {
        const label i = spList[t >> 3];
        const label c = spCells[(size_t)(t & 7) * nSp + i];
        const int32_t b0 = sliceOff[c >> 5];
        const int w = sliceOff[(c >> 5) + 1] - b0;
        const int32_t *row = col + ((size_t)b0 << 5) + (c & 31);
        const int mine = (int)lvl8[c] - 1;
        int mx = 0;
        for (int j0 = 0; j0 < w; j0 += 8) {
            int32_t q[8];
            int lq[8], uq[8];
#pragma unroll
            for (int k = 0; k < 8; k++) q[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : c;
#pragma unroll
            for (int k = 0; k < 8; k++) { lq[k] = lvl8[q[k]]; uq[k] = uc[q[k]]; }
#pragma unroll
            for (int k = 0; k < 8; k++) mx = max(mx, lq[k] - uq[k]);
        }
        if (mine < mx - 1) knockPoint(i, nSp, spCells, spFlag, uc, qOut, qOutCount, qCap);
    }
}
// iteration j >= 1: the cells that the knocked points of queue j reverted (uc = 0 now, final level lvl[q]) against their
// marked neighbours: a marked neighbour c with lvl[c] - 1 < lvl[q] - 1 cannot be unrefined any more
__global__ void __launch_bounds__(kThreads)
k_unref_frontier(const label *__restrict__ qIn, const int *__restrict__ qInCount, int64_t nSp, const label *__restrict__ spCells,
                 const label *__restrict__ cellSp, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                 const u8 *__restrict__ lvl8, u8 *spFlag, u8 *uc, label *qOut, int *qOutCount, int64_t qCap,
                 const int *__restrict__ gate) {
    if (*gate == 0) return;
    const int cnt = min(*qInCount, (int)min(qCap, (int64_t)0x7fffffff));
    for (int t = blockIdx.x * kThreads + threadIdx.x; t < cnt * 8; t += gridDim.x * kThreads) {
        const label i = qIn[t >> 3];
        const label q = spCells[(size_t)(t & 7) * nSp + i];
        const int32_t b0 = sliceOff[q >> 5];
        const int w = sliceOff[(q >> 5) + 1] - b0;
        const int32_t *row = col + ((size_t)b0 << 5) + (q & 31);
        const int lq = lvl8[q];
        // 8 neighbours per round, loads batched (labels, then flags + levels, then the split points of the hits): the walk
        // is a chain of dependent gathers and the queues are short, so this kernel runs at the latency of one chain
        for (int j0 = 0; j0 < w; j0 += 8) {
            int32_t c[8];
            unsigned uq[8], lc[8];
#pragma unroll
            for (int k = 0; k < 8; k++) c[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : q;
#pragma unroll
            for (int k = 0; k < 8; k++) { uq[k] = uc[c[k]]; lc[k] = lvl8[c[k]]; }
            label sp[8];
#pragma unroll
            for (int k = 0; k < 8; k++) sp[k] = (c[k] != q && uq[k] && (int)lc[k] < lq) ? cellSp[c[k]] : -1;
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (sp[k] >= 0) knockPoint(sp[k], nSp, spCells, spFlag, uc, qOut, qOutCount, qCap);
        }
    }
}
// coupled faces of an iteration (hexRef3D.C:1855-1889): ghost = lvl - uc of the cell across the face; fused with the
// receive side of the swap and the all-reduce of the queue size
__global__ void __launch_bounds__(kThreads)
k_fix_unref_queue_link(const __grid_constant__ GhostLink G, int64_t nC, const label *__restrict__ cplFace, const label *__restrict__ bOwner,
                       const u8 *__restrict__ lvl8, int64_t nSp, const label *__restrict__ spCells, const label *__restrict__ cellSp,
                       u8 *spFlag, u8 *uc, label *qOut, int *qOutCount, int64_t qCap, int *gqOut, const int *__restrict__ gate) {
    if (gate && *gate == 0) return;
    const u8 *ghostLfm = reinterpret_cast<const u8 *>(ghostAcquire(G));
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ghostLfm && i < nC) {
        const label bf = cplFace[i];
        const label o = bOwner[bf];
        if (uc[o] && ((int)lvl8[o] - 1) < (int)ghostLfm[bf] - 1) {
            const label sp = cellSp[o];
            if (sp >= 0) knockPoint(sp, nSp, spCells, spFlag, uc, qOut, qOutCount, qCap);
        }
    }
    if (ghostRelease(G) && G.peer) ctaAllReduceSum(G, qOutCount, gqOut);
}
// "problem" of hexRef3D.C:1808-1833: an UNMARKED cell more than one level below the final level of a face neighbour.
// Impossible on a 2:1-balanced mesh; evaluated on the final state when the mesh was found unbalanced.
__global__ void __launch_bounds__(kThreads)
k_unref_problem(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col, const u8 *__restrict__ lvl8,
                const u8 *__restrict__ uc, int *problem) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n || uc[c]) return;
    const int32_t b0 = sliceOff[c >> 5];
    const int w = sliceOff[(c >> 5) + 1] - b0;
    const int32_t *row = col + ((size_t)b0 << 5) + (c & 31);
    const int mine = lvl8[c];
    for (int j = 0; j < w; j++) {
        const int32_t q = row[(size_t)j << 5];
        if ((int)lvl8[q] - (int)uc[q] - 1 > mine) { *problem = 1; return; }
    }
}
__global__ void k_fix_unref_problem(int64_t nC, const label *__restrict__ cplFace, const label *__restrict__ bOwner,
                                    const u8 *__restrict__ lvl8, const u8 *__restrict__ ghostLfm, const u8 *__restrict__ uc, int *problem) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nC) return;
    const label bf = cplFace[i];
    const label o = bOwner[bf];
    if (!uc[o] && (int)lvl8[o] < (int)ghostLfm[bf] - 1) *problem = 1;
}

// ------------------------------------------------------------------------------------------
// STAGE 4  all registered fields through the map in ONE launch
// ------------------------------------------------------------------------------------------
// fvMesh::mapFields maps every registered field through the same mapPolyMesh::cellMap() (MapGeometricFields over the
// object registry); the refiner then carries refineCell through the same map (fvMeshHexRefiner.C:1562-1586).  One
// launch covers all of them: CTA b works on chunk b / nStreams of the new cells for stream b % nStreams (a field, or
// the marking), so the CTAs that need the same piece of cellMap are adjacent in launch order and run at the same time
// -- the piece comes from HBM once and from L2 for the other streams.  Each CTA runs the single-field code path
// (pairs of consecutive elements: one 128-bit store, one 128-bit load where the two sources are consecutive), which
// keeps the register count of the standalone kernels.
constexpr int kMapChunk = 2048;          // new cells per CTA
constexpr int kMaxMapFields = 8;
struct MapFieldSet {
    int nFields;
    int nc[kMaxMapFields];
    const double *oldF[kMaxMapFields];
    double *newF[kMaxMapFields];
};

template <int NC>
__device__ __forceinline__ void mapChunkField(const label *__restrict__ cellMap, int64_t c0, int cnt, const double *__restrict__ oldF,
                                              double *__restrict__ newF) {
    const int64_t e0 = c0 * NC;                              // even: c0 is a multiple of kMapChunk
    const int total = cnt * NC;
    if ((((uintptr_t)oldF) | ((uintptr_t)newF)) & 15) {      // unaligned field: element form
        for (int e = threadIdx.x; e < total; e += kThreads) {
            const int i = e / NC;
            newF[e0 + e] = oldF[(int64_t)cellMap[c0 + i] * NC + (e - i * NC)];
        }
        return;
    }
    const int pairs = total >> 1;
    for (int p0 = threadIdx.x; p0 < pairs; p0 += 4 * kThreads) {
        int64_t s0[4], s1[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int p = p0 + t * kThreads;
            s0[t] = -1; s1[t] = -1;
            if (p < pairs) {
                const int e = 2 * p;
                const int i = e / NC, k = e - i * NC;
                s0[t] = (int64_t)cellMap[c0 + i] * NC + k;
                s1[t] = (k + 1 < NC) ? s0[t] + 1 : (int64_t)cellMap[c0 + i + 1] * NC;
            }
        }
        double2 v[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            if (s0[t] < 0) continue;
            if (s1[t] == s0[t] + 1 && !(s0[t] & 1)) v[t] = *reinterpret_cast<const double2 *>(oldF + s0[t]);
            else { v[t].x = oldF[s0[t]]; v[t].y = oldF[s1[t]]; }
        }
#pragma unroll
        for (int t = 0; t < 4; t++)
            if (s0[t] >= 0) *reinterpret_cast<double2 *>(newF + e0 + 2 * (p0 + t * kThreads)) = v[t];
    }
    if ((total & 1) && threadIdx.x == 0) {                   // odd tail (last chunk only)
        const int e = total - 1, i = e / NC;
        newF[e0 + e] = oldF[(int64_t)cellMap[c0 + i] * NC + (e - i * NC)];
    }
}
__device__ __noinline__ void mapChunkFieldN(const label *__restrict__ cellMap, int64_t c0, int cnt, int nc, const double *__restrict__ oldF,
                                            double *__restrict__ newF) {
    const int total = cnt * nc;
    for (int e = threadIdx.x; e < total; e += kThreads) {
        const int i = e / nc;
        newF[c0 * nc + e] = oldF[(int64_t)cellMap[c0 + i] * nc + (e - i * nc)];
    }
}

template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_map_fields_fused(int64_t nNew, const label *__restrict__ cellMap, const __grid_constant__ MapFieldSet F, int nStreams,
                   const label *__restrict__ reverseCellMap, const u8 *__restrict__ oldFlag, u8 *__restrict__ newFlag) {
    const int64_t chunk = blockIdx.x / nStreams;
    const int stream = blockIdx.x - (int)chunk * nStreams;
    const int64_t c0 = chunk * kMapChunk;
    const int cnt = (int)min((int64_t)kMapChunk, nNew - c0);
    if (stream < F.nFields) {
        const double *oldF = F.oldF[stream];
        double *newF = F.newF[stream];
        switch (F.nc[stream]) {
        case 1: mapChunkField<1>(cellMap, c0, cnt, oldF, newF); break;
        case 3: mapChunkField<3>(cellMap, c0, cnt, oldF, newF); break;
        case 6: mapChunkField<6>(cellMap, c0, cnt, oldF, newF); break;
        case 9: mapChunkField<9>(cellMap, c0, cnt, oldF, newF); break;
        default: mapChunkFieldN(cellMap, c0, cnt, F.nc[stream], oldF, newF); break;
        }
    } else {
        // refineCell through the map, fvMeshHexRefiner.C:1562-1586: new cells (no old cell, or not the cell the old one
        // became) are marked, the others keep their flag.  The three dependent loads batched over 8 cells per thread.
        for (int h = 0; h < kMapChunk; h += 4 * kThreads) {
            label o[4], r[4];
            u8 v[4];
#pragma unroll
            for (int t = 0; t < 4; t++) { const int i = h + threadIdx.x + t * kThreads; o[t] = (i < cnt) ? cellMap[c0 + i] : -1; }
#pragma unroll
            for (int t = 0; t < 4; t++) r[t] = (o[t] >= 0) ? reverseCellMap[o[t]] : -1;
#pragma unroll
            for (int t = 0; t < 4; t++) v[t] = (o[t] >= 0) ? oldFlag[o[t]] : (u8)1;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int i = h + threadIdx.x + t * kThreads;
                if (i < cnt) newFlag[c0 + i] = (o[t] < 0 || (int64_t)r[t] != c0 + i) ? (u8)1 : (v[t] ? (u8)1 : (u8)0);
            }
        }
    }
}
