// kernels_linked.cuh -- the two closure loops across ranks as ONE launch per iteration over NVLink peer memory.
//
// hexRef::consistentRefinement (hexRef.C:1402-1468) and hexRef3D::consistentUnrefinement (hexRef3D.C:1719-1953) swap
// the level +/- flag of the cell behind every processor face (syncTools::swapBoundaryFaceList, hexRef.C:756,
// hexRef3D.C:1843-1853) and reduce nChanged over the ranks (hexRef.C:1424) in every sweep.  kernels_sparse.cuh (2)/(3)
// run these loops as queue iterations; with a communicator an iteration used to be three launches (push the halo, the
// queue kernel, the coupled-face fix-up with the receive side and the all-reduce).  Here it is one:
//   * the swap happens ONCE per closure (k_halo_push_cells, the state before sweep 0) into the neighbours' ghost
//     arrays; after that a cell that owns processor faces sends its new value the moment it changes -- the thread that
//     flips it stores the byte straight into the neighbours' ghost arrays (NVLink store, "push on change").  Values
//     only grow in both loops (marks are only set / unrefinement flags only dropped), every flip is forced, so a
//     neighbour that reads the old or the new byte during an iteration reaches the same fixed point; what it missed
//     it sees in the next iteration, which runs because the flipped cell sits in a queue and the all-reduced queue
//     size is therefore non-zero on every rank.
//   * the coupled-face rule runs in extra CTAs of the same grid (blockIdx.x < nFixBlocks), concurrently with the queue
//     CTAs: it reads the ghost array in place (the parity buffer of the initial swap IS the ghost array).
//   * the last CTA to finish all-reduces the size of the queue it produced (ctaAllReduceSum); every CTA that stored to a
//     peer fences at system scope before it counts itself finished, so the reduction doubles as the arrival signal of
//     the iteration's remote stores: a rank leaves iteration j only after every rank has finished it.
// Results are identical to the three-launch form (and to the reference): same rules on the same monotone closure.
#pragma once
#include "kernels_sparse.cuh"

// where the value of a cell that owns processor faces goes: for every such face the byte offset in the neighbour's
// ghost array (boundary-face layout, HaloPlan::offTheirs + k) and the neighbour's rank, packed rank << 26 | offset
// (offsets stay below kHaloCap = 2^26, ranks below kMaxRanks = 2^6)
struct LinkOut {
    const label *cplCell;      // [nCpl] ascending owners of coupled faces
    const int32_t *off;        // [nCpl + 1]
    const unsigned *dst;       // [nCoupled]
    const u8 *cellGhost;       // [n] cell owns a coupled face
    int nCpl;
};
static_assert(kHaloCap == ((int64_t)1 << 26) && kMaxRanks <= 64, "LinkOut::dst packing");

__device__ __forceinline__ void linkPush(const LinkOut &L, char *const *peer, int parity, label c, u8 v) {
    int lo = 0, hi = L.nCpl - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (L.cplCell[mid] < c) lo = mid + 1; else hi = mid;
    }
    for (int r = L.off[lo]; r < L.off[lo + 1]; r++) {
        const unsigned d = L.dst[r];
        peer[d >> 26][sizeof(P2PRegionHeader) + (size_t)parity * kHaloCap + (size_t)(d & 0x3ffffffu)] = (char)v;
    }
}

// end of a linked iteration: true in the last CTA to finish.  A CTA that stored to a peer fences at system scope.
__device__ __forceinline__ bool linkedRelease(const GhostLink &G, const int *sRemote, bool first, unsigned seq0) {
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (*sRemote) __threadfence_system(); else __threadfence();
        last = (atomicAdd(G.counter, 1u) == gridDim.x - 1);
        if (last) { *G.counter = 0; if (first && !G.ghost) G.seq[0] = seq0 + 1u; }   // the initial swap is consumed by sweep 0 (a rank
                                                                                      // without processor faces made none)
    }
    __syncthreads();
    return last != 0;
}

// ---- consistentRefinement, maxSet = true ------------------------------------------------------------------------
// FIRST: sweep 0 (rows of all marked cells; waits for the initial swap).  Otherwise iteration j >= 1 on queue j, gated by
// the all-reduced size of queue j.
template <bool FIRST>
__global__ void __launch_bounds__(kThreads)
k_refine_linked(const __grid_constant__ GhostLink G, const LinkOut L, unsigned nFixBlocks, int64_t nC, const label *__restrict__ cplFace,
                const label *__restrict__ bOwner, int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                const u8 *__restrict__ lvl8, u8 *lf, u8 *set, bool aligned, const label *__restrict__ qIn, const int *__restrict__ qInCount,
                label *qOut, int *qOutCount, int64_t qCap, int *gqOut, const int *__restrict__ gate) {
    if (!FIRST && *gate == 0) return;
    __shared__ int sRemote;
    const unsigned seq0 = G.seq[0];
    const int parity = FIRST ? (int)((seq0 + 1u) & 1u) : (int)(seq0 & 1u);
    if (threadIdx.x == 0) sRemote = 0;
    __syncthreads();
    auto onSet = [&](label c, u8 v) {
        if (L.cellGhost && L.cellGhost[c]) { linkPush(L, G.peer, parity, c, v); sRemote = 1; }
    };
    if (blockIdx.x < nFixBlocks) {
        // coupled faces (hexRef.C:758-780): ghostLf = level + flag of the cell across the face
        const u8 *ghostLf = FIRST ? reinterpret_cast<const u8 *>(ghostAcquire(G))
                                  : reinterpret_cast<const u8 *>(G.region + sizeof(P2PRegionHeader) + (size_t)parity * kHaloCap);
        const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (ghostLf && i < nC) {
            const label bf = cplFace[i];
            const label o = bOwner[bf];
            const int lo = lvl8[o];
            if ((int)ghostLf[bf] > lo + 1 && !set[o]) {
                if (setByteOnce(set, o)) {
                    lf[o] = (u8)(lo + 1);
                    const int pos = atomicAdd(qOutCount, 1);
                    if (pos < qCap) qOut[pos] = o;
                    onSet(o, (u8)(lo + 1));
                }
            }
        }
    } else {
        const unsigned blk = blockIdx.x - nFixBlocks, nBlk = gridDim.x - nFixBlocks;
        auto rowOf = [&](int64_t c) {
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
                            onSet(q[k], (u8)(lq[k] + 1));
                        }
                    }
                }
            }
        };
        if (FIRST) {
            forEachSetWarpAt(blk, set, n, aligned, rowOf);
        } else {
            const int cnt = min(*qInCount, (int)min(qCap, (int64_t)0x7fffffff));
            for (int i = blk * kThreads + threadIdx.x; i < cnt; i += nBlk * kThreads) rowOf(qIn[i]);
        }
    }
    if (linkedRelease(G, &sRemote, FIRST, seq0)) ctaAllReduceSum(G, qOutCount, gqOut);
}

// ---- consistentUnrefinement on the split-point table ----------------------------------------------------------------
// ghost = cellLevel - unrefineCell of the cell across the face (hexRef3D.C:1843-1853); a knocked split point reverts its
// 8 cells, whose ghost value becomes their level again.
template <bool FIRST>
__global__ void __launch_bounds__(kThreads)
k_unref_linked(const __grid_constant__ GhostLink G, const LinkOut L, unsigned nFixBlocks, int64_t nC, const label *__restrict__ cplFace,
               const label *__restrict__ bOwner, const u8 *__restrict__ lvl8, int64_t nSp, const label *__restrict__ spCells,
               const label *__restrict__ cellSp, u8 *spFlag, u8 *uc, const label *__restrict__ listIn, const int *__restrict__ listCount,
               const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col, label *qOut, int *qOutCount, int64_t qCap, int *gqOut,
               const int *__restrict__ gate) {
    if (!FIRST && *gate == 0) return;
    __shared__ int sRemote;
    const unsigned seq0 = G.seq[0];
    const int parity = FIRST ? (int)((seq0 + 1u) & 1u) : (int)(seq0 & 1u);
    if (threadIdx.x == 0) sRemote = 0;
    __syncthreads();
    auto knock = [&](label i) {
        if (clearByteOnce(spFlag, i)) {
            label cs[8];
#pragma unroll
            for (int k = 0; k < 8; k++) cs[k] = spCells[(size_t)k * nSp + i];
#pragma unroll
            for (int k = 0; k < 8; k++) uc[cs[k]] = 0;
            const int pos = atomicAdd(qOutCount, 1);
            if (pos < qCap) qOut[pos] = i;
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (L.cellGhost && L.cellGhost[cs[k]]) { linkPush(L, G.peer, parity, cs[k], lvl8[cs[k]]); sRemote = 1; }
        }
    };
    if (blockIdx.x < nFixBlocks) {
        const u8 *ghostLfm = FIRST ? reinterpret_cast<const u8 *>(ghostAcquire(G))
                                   : reinterpret_cast<const u8 *>(G.region + sizeof(P2PRegionHeader) + (size_t)parity * kHaloCap);
        const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (ghostLfm && i < nC) {
            const label bf = cplFace[i];
            const label o = bOwner[bf];
            if (uc[o] && ((int)lvl8[o] - 1) < (int)ghostLfm[bf] - 1) {
                const label sp = cellSp[o];
                if (sp >= 0) knock(sp);
            }
        }
    } else {
        const unsigned blk = blockIdx.x - nFixBlocks, nBlk = gridDim.x - nFixBlocks;
        const int cnt = FIRST ? *listCount : min(*listCount, (int)min(qCap, (int64_t)0x7fffffff));
        for (int t = blk * kThreads + threadIdx.x; t < cnt * 8; t += nBlk * kThreads) {
            const label i = listIn[t >> 3];
            const label c = spCells[(size_t)(t & 7) * nSp + i];
            const int32_t b0 = sliceOff[c >> 5];
            const int w = sliceOff[(c >> 5) + 1] - b0;
            const int32_t *row = col + ((size_t)b0 << 5) + (c & 31);
            if (FIRST) {
                // sweep 0 (hexRef3D.C:1785-1839): marked cell c (final level lvl - 1) against the final levels of its neighbours
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
                if (mine < mx - 1) knock(i);
            } else {
                // iteration j >= 1: c was reverted by queue j (uc = 0, final level lvl[c]); its marked neighbours that would end
                // up more than one level below it cannot be unrefined any more
                const int lc = lvl8[c];
                for (int j0 = 0; j0 < w; j0 += 8) {
                    int32_t q[8];
                    unsigned uq[8], lq[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) q[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : c;
#pragma unroll
                    for (int k = 0; k < 8; k++) { uq[k] = uc[q[k]]; lq[k] = lvl8[q[k]]; }
                    label sp[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) sp[k] = (q[k] != c && uq[k] && (int)lq[k] < lc) ? cellSp[q[k]] : -1;
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        if (sp[k] >= 0) knock(sp[k]);
                }
            }
        }
    }
    if (linkedRelease(G, &sRemote, FIRST, seq0)) ctaAllReduceSum(G, qOutCount, gqOut);
}
