// sell.cuh -- device-side construction of the SELL-32 adjacencies (setup, not on the timed path).
//
// The reference walks mesh.cells()[c] -> faces -> owner/neighbour (fvMeshRefiner.C:807-862,
// hexRef.C:709-741).  Every decision on this path is a max / OR / integer compare over a cell's
// face neighbours, so the order inside a row never changes a result; the
// device layout is therefore free: one warp = one slice of 32 consecutive rows, column-major,
// padded with the row's own index, which is a no-op for every operator used here.  Rows hold internal
// neighbours only; processor-patch faces are applied by the boundary fix-up kernels.
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include <algorithm>
#include <vector>

__global__ void k_deg_internal(const label *__restrict__ owner, const label *__restrict__ neighbour, int64_t f,
                               int32_t *__restrict__ deg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f) return;
    atomicAdd(&deg[owner[i]], 1);
    atomicAdd(&deg[neighbour[i]], 1);
}
// one warp per slice: width = max row length
__global__ void k_slice_width(const int32_t *__restrict__ deg, int64_t rows, int64_t slices,
                              int32_t *__restrict__ width) {
    int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (s >= slices) return;
    int64_t r = s * 32 + lane;
    int d = r < rows ? deg[r] : 0;
    for (int k = 16; k > 0; k >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, k));
    if (lane == 0) width[s] = d;
}
__global__ void k_sell_fill_pad(int32_t *__restrict__ col, const int32_t *__restrict__ sliceOff, int64_t rows,
                                int selfPad) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t s = r >> 5;
    if (s * 32 >= ((rows + 31) & ~(int64_t)31)) return;
    int lane = (int)(r & 31);
    int32_t base = sliceOff[s], w = sliceOff[s + 1] - base;
    int32_t v = selfPad ? (int32_t)(r < rows ? r : rows - 1) : -1;
    for (int j = 0; j < w; j++) col[((size_t)(base + j) << 5) + lane] = v;
}
__device__ __forceinline__ void sellPut(int32_t *col, const int32_t *sliceOff, int32_t *cursor, int32_t row,
                                        int32_t v) {
    int j = atomicAdd(&cursor[row], 1);
    col[((size_t)(sliceOff[row >> 5] + j) << 5) + (row & 31)] = v;
}
__global__ void k_sell_fill_internal(const label *__restrict__ owner, const label *__restrict__ neighbour, int64_t f,
                                     const int32_t *__restrict__ sliceOff, int32_t *__restrict__ cursor,
                                     int32_t *__restrict__ col) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f) return;
    label o = owner[i], q = neighbour[i];
    sellPut(col, sliceOff, cursor, o, q);
    sellPut(col, sliceOff, cursor, q, o);
}
__global__ void k_sell_fill_csr(const int32_t *__restrict__ off, const int32_t *__restrict__ idx, int64_t rows,
                                const int32_t *__restrict__ sliceOff, int32_t *__restrict__ col) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int32_t s = off[r], e = off[r + 1];
    int32_t base = sliceOff[r >> 5];
    int lane = (int)(r & 31);
    for (int32_t j = 0; j < e - s; j++) col[((size_t)(base + j) << 5) + lane] = idx[s + j];
}
__global__ void k_csr_deg(const int32_t *__restrict__ off, int64_t rows, int32_t *__restrict__ deg) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) deg[r] = off[r + 1] - off[r];
}
// ascending insertion sort of each row (short rows; makes the layout deterministic and the gathers monotone)
__global__ void k_sell_sort_rows(int32_t *__restrict__ col, const int32_t *__restrict__ sliceOff,
                                 const int32_t *__restrict__ deg, int64_t rows) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int32_t *row = col + ((size_t)sliceOff[r >> 5] << 5) + (r & 31);
    int d = deg[r];
    for (int i = 1; i < d; i++) {
        int32_t v = row[(size_t)i << 5];
        int j = i - 1;
        while (j >= 0 && row[(size_t)j << 5] > v) { row[(size_t)(j + 1) << 5] = row[(size_t)j << 5]; j--; }
        row[(size_t)(j + 1) << 5] = v;
    }
}
__global__ void k_deg_to_u8(const int32_t *__restrict__ deg, int64_t rows, u8 *__restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) out[r] = (u8)min(deg[r], 255);
}

// largest tile (kTileSlices consecutive slices) in slice-width units
__global__ void k_max_tile(const int32_t *__restrict__ sliceOff, int64_t slices, int tileSlices, int *__restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t s0 = t * tileSlices;
    if (s0 >= slices) return;
    int64_t s1 = s0 + tileSlices < slices ? s0 + tileSlices : slices;
    atomicMax(out, sliceOff[s1] - sliceOff[s0]);
}
__global__ void k_fill_tail(int32_t *__restrict__ sliceOff, int64_t slices, int extra) {
    int i = threadIdx.x;
    if (i >= 1 && i <= extra) sliceOff[slices + i] = sliceOff[slices];
}

constexpr int kSliceOffPad = 16;
constexpr int kTileCapCols = 64;        // columns (128-byte lines) of a pipeline tile: 8 KB per ring stage

__global__ void k_slice_ghost_cells(const u8 *__restrict__ cellGhost, int64_t rows, u8 *__restrict__ sliceGhost) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows && cellGhost[r]) sliceGhost[r >> 5] = 1;
}

static void sellFree(amr_ctx *ctx, Sell &S) {
    if (S.sliceGhost) { cudaFree(S.sliceGhost); }
    if (S.tileFirst) { cudaFree(S.tileFirst); }
    if (S.sliceOff) { cudaFree(S.sliceOff); ctx->devBytes -= (int64_t)((S.slices + kSliceOffPad) * 4); }
    if (S.col) { cudaFree(S.col); ctx->devBytes -= (int64_t)(S.entries * 4 > 4 ? S.entries * 4 : 4); }
    if (S.count) { cudaFree(S.count); ctx->devBytes -= (int64_t)(S.rows > 0 ? S.rows : 1); }
    S = Sell();
}

// shared tail of both builders: deg[rows] (device, int32) is known
static int sellAllocFromDeg(amr_ctx *ctx, Sell &S, int64_t rows, int32_t *deg, int selfPad) {
    S.rows = rows;
    S.slices = (rows + 31) / 32;
    TRY(devAlloc(ctx, &S.sliceOff, S.slices + kSliceOffPad));
    CK(cudaMemsetAsync(S.sliceOff, 0, (size_t)(S.slices + kSliceOffPad) * 4, ctx->stream));
    if (S.slices > 0) LAUNCH(k_slice_width, gridFor(S.slices * 32), kThreads, deg, rows, S.slices, S.sliceOff);
    // scan aux
    int64_t auxNeed = (S.slices + 1) / kScanTile + 1024;
    int32_t *aux = nullptr;
    CK(cudaMalloc((void **)&aux, (size_t)auxNeed * 4 * 2));
    int rc = deviceExclusiveScan(ctx, S.sliceOff, S.sliceOff, S.slices, aux, auxNeed * 2);
    if (rc != AMR_OK) { cudaFree(aux); return rc; }
    int32_t total = 0;
    LAUNCH(k_fill_tail, 1, 32, S.sliceOff, S.slices, kSliceOffPad - 1);
    // Tiles of the TMA pipeline (pipeline.cuh): up to 8 consecutive slices whose columns fit kTileCapCols, so that the ring
    // stage stays small on graded meshes too (a few slices next to a level jump are 9-24 columns wide; sizing the stage for
    // the widest run of 8 slices costs resident CTAs: profiles/r2_ncu_full.csv).  Greedy and sequential: done on the host
    // from the slice offsets (4 B per 32 rows), once per mesh.
    {
        std::vector<int32_t> off((size_t)S.slices + 1);
        CK(cudaMemcpyAsync(off.data(), S.sliceOff, (size_t)(S.slices + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        total = off[(size_t)S.slices];
        std::vector<int32_t> first;
        first.reserve((size_t)S.slices / 8 + 2);
        int maxCols = 0;
        for (int64_t s0 = 0; s0 < S.slices;) {
            int64_t s1 = s0 + 1;
            while (s1 < S.slices && s1 - s0 < 8 && off[(size_t)s1 + 1] - off[(size_t)s0] <= kTileCapCols) s1++;
            first.push_back((int32_t)s0);
            maxCols = std::max(maxCols, off[(size_t)s1] - off[(size_t)s0]);
            s0 = s1;
        }
        S.tiles = (int64_t)first.size();
        first.push_back((int32_t)S.slices);
        S.maxTileInts = maxCols * 32;
        TRY(devAlloc(ctx, &S.tileFirst, (int64_t)first.size()));
        CK(cudaMemcpyAsync(S.tileFirst, first.data(), first.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(aux);
    S.entries = (int64_t)total * 32;
    TRY(devAlloc(ctx, &S.col, S.entries));
    TRY(devAlloc(ctx, &S.count, rows));
    if (rows > 0) {
        LAUNCH(k_sell_fill_pad, gridFor(S.slices * 32), kThreads, S.col, S.sliceOff, rows, selfPad);
        LAUNCH(k_deg_to_u8, gridFor(rows), kThreads, deg, rows, S.count);
    }
    return AMR_OK;
}

// cell -> face-neighbour cells from owner/neighbour (internal faces only; coupled faces: k_fix_* kernels)
static int sellBuildCellCells(amr_ctx *ctx) {
    sellFree(ctx, ctx->cc);
    int64_t n = ctx->n, f = ctx->f;
    int32_t *deg = nullptr, *cursor = nullptr;
    CK(cudaMalloc((void **)&deg, (size_t)(n + 1) * 4));
    CK(cudaMalloc((void **)&cursor, (size_t)(n + 1) * 4));
    CK(cudaMemsetAsync(deg, 0, (size_t)(n + 1) * 4, ctx->stream));
    CK(cudaMemsetAsync(cursor, 0, (size_t)(n + 1) * 4, ctx->stream));
    if (f > 0) LAUNCH(k_deg_internal, gridFor(f), kThreads, ctx->owner, ctx->neighbour, f, deg);
    int rc = sellAllocFromDeg(ctx, ctx->cc, n, deg, 1);
    if (rc == AMR_OK) {
        if (f > 0) LAUNCH(k_sell_fill_internal, gridFor(f), kThreads, ctx->owner, ctx->neighbour, f, ctx->cc.sliceOff, cursor, ctx->cc.col);
        if (n > 0) LAUNCH(k_sell_sort_rows, gridFor(n), kThreads, ctx->cc.col, ctx->cc.sliceOff, deg, n);
        if (cudaMalloc((void **)&ctx->cc.sliceGhost, (size_t)ctx->cc.slices + 16) == cudaSuccess) {
            cudaMemsetAsync(ctx->cc.sliceGhost, 0, (size_t)ctx->cc.slices + 16, ctx->stream);
            // slices containing a cell that owns a coupled face (cellGhost is built by amr_mesh_set)
            if (n > 0 && ctx->nCoupled > 0 && ctx->cellGhost) LAUNCH(k_slice_ghost_cells, gridFor(n), kThreads, (const u8 *)ctx->cellGhost, n, ctx->cc.sliceGhost);
        }
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("sellBuildCellCells: ") + cudaGetErrorString(e); rc = AMR_ERR_CUDA; }
    }
    cudaFree(deg); cudaFree(cursor);
    return rc;
}

// point -> cells from the host's pointCells CSR (device copies off/idx)
static int sellBuildFromCsr(amr_ctx *ctx, Sell &S, int64_t rows, const int32_t *off, const int32_t *idx) {
    sellFree(ctx, S);
    int32_t *deg = nullptr;
    CK(cudaMalloc((void **)&deg, (size_t)(rows + 1) * 4));
    if (rows > 0) LAUNCH(k_csr_deg, gridFor(rows), kThreads, off, rows, deg);
    int rc = sellAllocFromDeg(ctx, S, rows, deg, 0);
    if (rc == AMR_OK && rows > 0) {
        LAUNCH(k_sell_fill_csr, gridFor(rows), kThreads, off, idx, rows, S.sliceOff, S.col);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("sellBuildFromCsr: ") + cudaGetErrorString(e); rc = AMR_ERR_CUDA; }
    }
    cudaFree(deg);
    return rc;
}

// ---- N1: adjacency carried through a topology change (hexRef::updateMesh, hexRef.C:2422-2608) -------------------------
// After a cut or a merge most rows of an adjacency are what they were, up to the renumbering of the labels: row r of the
// new mesh is row map[r] of the old one with every entry e replaced by rev[e] (mapPolyMesh::cellMap / reverseCellMap;
// rev == nullptr: labels kept, elements only appended).  Only the rows the host lists as dirty (ascending, CSR) are
// taken from the host.  The same kernels serve cell->cell (map = cellMap) and point->cell (map = pointMap, entries
// are cells: rev = reverseCellMap).
__global__ void k_upd_dirty_index(int64_t nDirty, const label *__restrict__ dirty, int64_t rowsNew, int32_t *__restrict__ dirtyIdx,
                                  int *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nDirty) return;
    const label r = dirty[i];
    if (r < 0 || r >= rowsNew || (i > 0 && dirty[i - 1] >= r)) { *bad = 1; return; }
    dirtyIdx[r] = (int32_t)i;
}
__global__ void k_upd_deg(int64_t rowsNew, int64_t rowsOld, const label *__restrict__ map, const int32_t *__restrict__ dirtyIdx,
                          const int32_t *__restrict__ dirtyOff, const u8 *__restrict__ oldCount, int32_t *__restrict__ deg,
                          int *__restrict__ bad) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rowsNew) return;
    const int32_t d = dirtyIdx[r];
    if (d >= 0) {
        const int32_t w = dirtyOff[d + 1] - dirtyOff[d];
        if (w < 0) { *bad = 1; deg[r] = 0; return; }
        deg[r] = w;
        return;
    }
    const label ro = map[r];
    if (ro < 0 || ro >= rowsOld) { *bad = 2; deg[r] = 0; return; }       // a clean row needs the row it came from
    const int c = oldCount[ro];
    if (c == 255) *bad = 3;                                              // saturated count: the old length is not known
    deg[r] = c;
}
__global__ void k_upd_fill(int64_t rowsNew, const label *__restrict__ map, const int32_t *__restrict__ dirtyIdx,
                           const int32_t *__restrict__ dirtyOff, const label *__restrict__ dirtyEntries, int64_t entryLimit,
                           const int32_t *__restrict__ oldSliceOff, const int32_t *__restrict__ oldCol, const label *__restrict__ rev,
                           const int32_t *__restrict__ deg, const int32_t *__restrict__ newSliceOff, int32_t *__restrict__ newCol,
                           int *__restrict__ bad) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rowsNew) return;
    int32_t *dst = newCol + ((size_t)newSliceOff[r >> 5] << 5) + (r & 31);
    const int w = deg[r];
    const int32_t d = dirtyIdx[r];
    if (d >= 0) {
        const label *src = dirtyEntries + dirtyOff[d];
        for (int j = 0; j < w; j++) {
            const label e = src[j];
            if (e < 0 || e >= entryLimit) { *bad = 4; continue; }
            dst[(size_t)j << 5] = e;
        }
        return;
    }
    const label ro = map[r];
    const int32_t *src = oldCol + ((size_t)oldSliceOff[ro >> 5] << 5) + (ro & 31);
    for (int j = 0; j < w; j++) {
        label e = src[(size_t)j << 5];
        if (rev) e = rev[e];
        if (e < 0 || e >= entryLimit) { *bad = 5; continue; }            // a clean row may not point at a removed element
        dst[(size_t)j << 5] = e;
    }
}
// per-row values through the change: new[r] = dirty ? given[d] : old[map[r]]
template <typename T>
__global__ void k_upd_values(int64_t rowsNew, const label *__restrict__ map, const int32_t *__restrict__ dirtyIdx,
                             const T *__restrict__ given, const T *__restrict__ oldV, T *__restrict__ newV) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rowsNew) return;
    const int32_t d = dirtyIdx[r];
    newV[r] = d >= 0 ? given[d] : oldV[map[r]];
}

// new adjacency from the old one + the dirty rows.  dirtyIdx [rowsNew] is scratch filled here and left for the caller.
static int sellCarryOver(amr_ctx *ctx, Sell &out, const Sell &old, int64_t rowsNew, const label *map, const label *rev,
                         int64_t nDirty, const label *dirty, const int32_t *dirtyOff, const label *dirtyEntries, int64_t entryLimit,
                         int selfPad, bool sortRows, int32_t *dirtyIdx) {
    int32_t *deg = nullptr;
    CK(cudaMalloc((void **)&deg, (size_t)(rowsNew + 1) * 4));
    int *bad = ctx->dFlags + 5;
    CK(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(dirtyIdx, 0xFF, (size_t)(rowsNew > 0 ? rowsNew : 1) * 4, ctx->stream));
    if (nDirty > 0) LAUNCH(k_upd_dirty_index, gridFor(nDirty), kThreads, nDirty, dirty, rowsNew, dirtyIdx, bad);
    if (rowsNew > 0) LAUNCH(k_upd_deg, gridFor(rowsNew), kThreads, rowsNew, old.rows, map, (const int32_t *)dirtyIdx, dirtyOff, (const u8 *)old.count, deg, bad);
    int rc = sellAllocFromDeg(ctx, out, rowsNew, deg, selfPad);
    if (rc == AMR_OK && rowsNew > 0) {
        LAUNCH(k_upd_fill, gridFor(rowsNew), kThreads, rowsNew, map, (const int32_t *)dirtyIdx, dirtyOff, dirtyEntries, entryLimit,
               (const int32_t *)old.sliceOff, (const int32_t *)old.col, rev, (const int32_t *)deg, (const int32_t *)out.sliceOff, out.col, bad);
        if (sortRows) LAUNCH(k_sell_sort_rows, gridFor(rowsNew), kThreads, out.col, out.sliceOff, deg, rowsNew);
    }
    int hBad = 0;
    if (rc == AMR_OK) {
        cudaError_t e = cudaMemcpyAsync(&hBad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("sellCarryOver: ") + cudaGetErrorString(e); rc = AMR_ERR_CUDA; }
    }
    cudaFree(deg);
    if (rc == AMR_OK && hBad) {
        static const char *why[] = {"", "dirty list not ascending / out of range", "a clean row has no old row (map out of range)",
                                    "an old row is longer than 254 entries", "a dirty row holds a label out of range",
                                    "a clean row points at a removed element: it must be listed as dirty"};
        ctx->err = std::string("incremental adjacency update: ") + why[hBad < 6 ? hBad : 0];
        rc = AMR_ERR_ARG;
    }
    return rc;
}

// ---- cell -> cells sharing an EDGE (polyRefiner) ------------------------------------------------------------------------
// refinement::edgeConsistentRefinement (refinement.C:305-387) applies the 2:1 rule to every unordered pair of an edge's
// cells.  Two cells that share a face share its edges, so on the internal faces the face rule is contained in the edge
// rule, and the closure of both is the closure of ONE rule on the graph "cells that share an edge".  That graph in
// SELL form lets the polyRefiner's refinement closure run on the hexRefiner's queue machinery (kernels_sparse.cuh (2)).
// Built once per mesh from edgeCells: ordered pairs per edge -> rows with duplicates -> sorted, deduplicated rows.
__global__ void k_ce_count(int64_t nEdges, const int32_t *__restrict__ ecOff, const int32_t *__restrict__ ecCol, const u8 *__restrict__ ecCount,
                           int32_t *__restrict__ degRaw) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nEdges) return;
    const int k = ecCount[e];
    if (k < 2) return;
    const int32_t *row = ecCol + ((size_t)ecOff[e >> 5] << 5) + (e & 31);
    for (int i = 0; i < k; i++) atomicAdd(&degRaw[row[(size_t)i << 5]], k - 1);
}
__global__ void k_ce_fill(int64_t nEdges, const int32_t *__restrict__ ecOff, const int32_t *__restrict__ ecCol, const u8 *__restrict__ ecCount,
                          const int64_t *__restrict__ rawOff, int32_t *__restrict__ cursor, int32_t *__restrict__ raw) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nEdges) return;
    const int k = ecCount[e];
    if (k < 2) return;
    const int32_t *row = ecCol + ((size_t)ecOff[e >> 5] << 5) + (e & 31);
    for (int i = 0; i < k; i++) {
        const int32_t a = row[(size_t)i << 5];
        const int pos = atomicAdd(&cursor[a], k - 1);
        int o = 0;
        for (int j = 0; j < k; j++) if (j != i) raw[rawOff[a] + pos + o++] = row[(size_t)j << 5];
    }
}
// exclusive scan of int32 counts into int64 offsets: one block-sum pass + a serial pass over the block sums (host-free,
// set-up path only)
__global__ void k_ce_block_sums(const int32_t *__restrict__ deg, int64_t n, int64_t *__restrict__ blockSum) {
    __shared__ long long part[kThreads / 32];
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    long long v = i < n ? deg[i] : 0;
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { long long t = 0; for (int k = 0; k < kThreads / 32; k++) t += part[k]; blockSum[blockIdx.x] = t; }
}
__global__ void k_ce_scan_blocks(int64_t *blockSum, int64_t nBlocks) {          // <<<1,1>>>: a few hundred thousand adds, once per mesh
    long long run = 0;
    for (int64_t i = 0; i < nBlocks; i++) { const long long v = blockSum[i]; blockSum[i] = run; run += v; }
    blockSum[nBlocks] = run;
}
__global__ void k_ce_offsets(const int32_t *__restrict__ deg, int64_t n, const int64_t *__restrict__ blockSum, int64_t *__restrict__ off) {
    __shared__ int smem[kThreads / 32 + 1];
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int v = i < n ? deg[i] : 0;
    int total;
    const int ex = blockExclusiveScan(v, smem, &total);
    if (i < n) off[i] = blockSum[blockIdx.x] + ex;
}
// sort + unique each raw row in place (short rows); deg receives the number of distinct neighbours
__global__ void k_ce_unique(int64_t n, const int64_t *__restrict__ rawOff, const int32_t *__restrict__ degRaw, int32_t *__restrict__ raw,
                            int32_t *__restrict__ deg) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    int32_t *r = raw + rawOff[c];
    const int d = degRaw[c];
    for (int i = 1; i < d; i++) {
        const int32_t v = r[i];
        int j = i - 1;
        while (j >= 0 && r[j] > v) { r[j + 1] = r[j]; j--; }
        r[j + 1] = v;
    }
    int u = 0;
    for (int i = 0; i < d; i++)
        if (r[i] != (int32_t)c && (u == 0 || r[u - 1] != r[i])) r[u++] = r[i];
    deg[c] = u;
}
__global__ void k_ce_to_sell(int64_t n, const int64_t *__restrict__ rawOff, const int32_t *__restrict__ raw, const int32_t *__restrict__ deg,
                             const int32_t *__restrict__ sliceOff, int32_t *__restrict__ col) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int32_t *r = raw + rawOff[c];
    int32_t *dst = col + ((size_t)sliceOff[c >> 5] << 5) + (c & 31);
    const int d = deg[c];
    for (int j = 0; j < d; j++) dst[(size_t)j << 5] = r[j];
}
// returns AMR_OK with out.col == nullptr when the transient pair list would not fit (the caller keeps the dense sweeps)
static int sellBuildEdgeNeighbours(amr_ctx *ctx, Sell &out, const Sell &ec, int64_t nEdges, int64_t n) {
    sellFree(ctx, out);
    if (n <= 0 || nEdges <= 0 || !ec.col) return AMR_OK;
    int32_t *degRaw = nullptr, *cursor = nullptr, *deg = nullptr, *raw = nullptr;
    int64_t *rawOff = nullptr, *blockSum = nullptr;
    const int64_t nBlocks = (n + kThreads - 1) / kThreads;
    auto release = [&]() { cudaFree(degRaw); cudaFree(cursor); cudaFree(deg); cudaFree(raw); cudaFree(rawOff); cudaFree(blockSum); };
    if (cudaMalloc((void **)&degRaw, (size_t)(n + 1) * 4) != cudaSuccess || cudaMalloc((void **)&cursor, (size_t)(n + 1) * 4) != cudaSuccess ||
        cudaMalloc((void **)&deg, (size_t)(n + 1) * 4) != cudaSuccess || cudaMalloc((void **)&rawOff, (size_t)(n + 1) * 8) != cudaSuccess ||
        cudaMalloc((void **)&blockSum, (size_t)(nBlocks + 1) * 8) != cudaSuccess) { cudaGetLastError(); release(); return AMR_OK; }
    cudaMemsetAsync(degRaw, 0, (size_t)(n + 1) * 4, ctx->stream);
    cudaMemsetAsync(cursor, 0, (size_t)(n + 1) * 4, ctx->stream);
    LAUNCH(k_ce_count, gridFor(nEdges), kThreads, nEdges, (const int32_t *)ec.sliceOff, (const int32_t *)ec.col, (const u8 *)ec.count, degRaw);
    LAUNCH(k_ce_block_sums, (unsigned)nBlocks, kThreads, (const int32_t *)degRaw, n, blockSum);
    LAUNCH(k_ce_scan_blocks, 1, 1, blockSum, nBlocks);
    LAUNCH(k_ce_offsets, (unsigned)nBlocks, kThreads, (const int32_t *)degRaw, n, (const int64_t *)blockSum, rawOff);
    long long total = 0;
    if (cudaMemcpyAsync(&total, blockSum + nBlocks, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        ctx->err = "sellBuildEdgeNeighbours: count pass failed"; cudaGetLastError(); release(); return AMR_ERR_CUDA;
    }
    size_t freeB = 0, totalB = 0;
    cudaMemGetInfo(&freeB, &totalB);
    if (total <= 0 || (size_t)total * 4 > freeB / 2 || cudaMalloc((void **)&raw, (size_t)total * 4) != cudaSuccess) { cudaGetLastError(); release(); return AMR_OK; }
    LAUNCH(k_ce_fill, gridFor(nEdges), kThreads, nEdges, (const int32_t *)ec.sliceOff, (const int32_t *)ec.col, (const u8 *)ec.count, (const int64_t *)rawOff, cursor, raw);
    LAUNCH(k_ce_unique, gridFor(n), kThreads, n, (const int64_t *)rawOff, (const int32_t *)degRaw, raw, deg);
    int rc = sellAllocFromDeg(ctx, out, n, deg, 1);
    if (rc == AMR_OK) {
        LAUNCH(k_ce_to_sell, gridFor(n), kThreads, n, (const int64_t *)rawOff, (const int32_t *)raw, (const int32_t *)deg, (const int32_t *)out.sliceOff, out.col);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "sellBuildEdgeNeighbours: fill failed"; cudaGetLastError(); rc = AMR_ERR_CUDA; }
    }
    release();
    return rc;
}
