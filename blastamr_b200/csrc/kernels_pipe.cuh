// kernels_pipe.cuh -- TMA-pipelined variants of the dense SELL gather kernels (see pipeline.cuh).
// Same arithmetic as the direct kernels in kernels.cuh (which remain the path for meshes whose
// tiles do not fit the shared-memory ring and for the sparse sweeps); results are bit-identical.
//
// Instruction discipline (the first version was issue-bound, not memory-bound): the slice width is
// warp-uniform and rows are padded with the row's own label, so a row is walked with a compile-time
// width (switch on w, fully unrolled), no per-entry predicate, no branch between the gathers -- all
// W loads of a row are issued back to back.  Labels are compared as 32-bit.
#pragma once
#include <type_traits>

#include "kernels.cuh"
#include "pipeline.cuh"

template <int W> using WTag = std::integral_constant<int, W>;

// acc(WTag<W>, const int32_t (&q)[W]) is called once (w <= 8) or once per chunk of 4 (w > 8)
template <class Acc>
__device__ __forceinline__ void sellRow(const int32_t *row, int w, int32_t self, Acc acc) {
#define SELL_CASE(W)                                                     \
    case W: {                                                            \
        int32_t q[W];                                                    \
        _Pragma("unroll") for (int k = 0; k < W; k++) q[k] = row[k << 5]; \
        acc(WTag<W>{}, q);                                               \
    } break;
    switch (w) {
    case 0: break;
        SELL_CASE(1) SELL_CASE(2) SELL_CASE(3) SELL_CASE(4) SELL_CASE(5) SELL_CASE(6) SELL_CASE(7) SELL_CASE(8)
    default:
        for (int j0 = 0; j0 < w; j0 += 4) {
            int32_t q[4];
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = (j0 + k < w) ? row[(j0 + k) << 5] : self;
            acc(WTag<4>{}, q);
        }
    }
#undef SELL_CASE
}

// The adjacency holds internal neighbours only.  Coupled (processor-patch) faces are applied afterwards by the
// k_fix_* kernels (kernels.cuh), one thread per coupled face -- the reference's own second loop over boundary
// faces -- so a decomposed mesh runs exactly the single-rank kernels.  (Earlier versions kept ghost slots in the
// rows: the ghost-aware walk cost +50 % on kp_estimate even when only 8 % of the slices had a ghost.)
template <int KIND>
__device__ __forceinline__ double estimateRow(const int32_t *row, int w, int32_t c, const double *__restrict__ x,
                                              double p0, double p1, bool useOffset) {
    const double xc = fieldXform<KIND>(x[c], p1, useOffset);
    double eL = 0.0;
    sellRow(row, w, c, [&](auto W, const int32_t *q) {
        constexpr int NW = decltype(W)::value;
        double v[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) v[k] = x[q[k]];
#pragma unroll
        for (int k = 0; k < NW; k++) {
            const double xv = fieldXform<KIND>(v[k], p1, useOffset);
            double eT = fabs(xc - xv);
            if (KIND == 1) {
                eT = eT / fmaxFoam(fminFoam(xc, xv), p0);     // scaledDelta.C:103-104
                eT = (q[k] == c) ? 0.0 : eT;                  // padding (0/0 guard)
            }
            eL = fmaxFoam(eL, eT);
        }
    });
    // delta clamps internal faces (delta.C:103); min(max_j d_j, .5) == max_j min(d_j, .5)
    return (KIND == 0) ? fminFoam(eL, 0.5) : eL;
}

// cellGhost (nullable): cells that own a coupled face keep the RAW value; k_fix_estimate / k_fix_normalize finish them
template <int KIND, bool SCALE>
__global__ void __launch_bounds__(kPipeThreads, (KIND == 1) ? 4 : 7)
kp_estimate(SellView S, const double *__restrict__ x, double p0, double p1, Thresholds thr, const u8 *__restrict__ cellGhost,
            double *__restrict__ err) {
    const int32_t n = (int32_t)S.rows;
    const bool useOffset = fabs(p1) > kSMALL;
    sellPipeline(S, [&](int64_t c64, bool valid, const int32_t *row, int w, bool ghostSlice) {
        const int32_t c = valid ? (int32_t)c64 : n - 1;
        const double e = estimateRow<KIND>(row, w, c, x, p0, p1, useOffset);
        if (valid) {
            const bool raw = !SCALE || (ghostSlice && cellGhost[c]);
            err[c] = raw ? e : normalizeValue(e, thr);
        }
    });
}

template <bool MAXSET>
__global__ void __launch_bounds__(kPipeThreads, 7)
kp_refine_sweep(SellView S, u8 *lf, u8 *set, int *changed, const int *__restrict__ gate) {
    if (gate && *gate == 1) return;                 // the sparse (push) form runs instead (k_refine_push)
    const int32_t n = (int32_t)S.rows;
    sellPipeline(S, [&](int64_t c64, bool valid, const int32_t *row, int w, bool) {
        const int32_t c = valid ? (int32_t)c64 : n - 1;
        const int isSet = set[c];
        int flip = 0;
        const bool idle = MAXSET ? __all_sync(0xffffffffu, isSet != 0) : !__any_sync(0xffffffffu, isSet != 0);
        if (!idle) {
            const int mine = lf[c];
            int ext = mine;      // padding contributes the cell's own value: neutral for max and min
            const u8 *lfr = lf;
            sellRow(row, w, c, [&](auto W, const int32_t *q) {
                constexpr int NW = decltype(W)::value;
                int v[NW];
#pragma unroll
                for (int k = 0; k < NW; k++) v[k] = lfr[q[k]];
#pragma unroll
                for (int k = 0; k < NW; k++) ext = MAXSET ? max(ext, v[k]) : min(ext, v[k]);
            });
            if (valid) {
                if (MAXSET) { if (!isSet && ext > mine + 1) { set[c] = 1; lf[c] = (u8)(mine + 1); flip = 1; } }
                else { if (isSet && mine > ext + 1) { set[c] = 0; lf[c] = (u8)(mine - 1); flip = 1; } }
            }
        }
        if (__any_sync(0xffffffffu, flip) && (threadIdx.x & 31) == 0) *changed = 1;
    });
}

__global__ void __launch_bounds__(kPipeThreads, 7)
kp_select_unrefine_points(SellView S, const u8 *__restrict__ pcClass,
                          const label *__restrict__ parentIdx, const u8 *__restrict__ lvl8,
                          const double *__restrict__ err, double unrefineLevel, const u8 *__restrict__ marked,
                          u8 *__restrict__ isSplit, u8 *__restrict__ keep) {
    sellPipeline(S, [&](int64_t pt, bool valid, const int32_t *row, int w, bool) {
        if (!valid) return;
        bool split = pcClass[pt] != 0;
        bool kp = false;
        if (split) {
            int32_t q[8];
#pragma unroll
            for (int k = 0; k < 8; k++) q[k] = row[k << 5];
            label par[8];
            int l[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { par[k] = parentIdx[q[k]]; l[k] = lvl8[q[k]]; }   // 16 gathers in flight
            bool same = par[0] >= 0;
#pragma unroll
            for (int k = 1; k < 8; k++) same = same && (par[k] == par[0]) && (l[k] == l[0]);
            split = same;
            if (split && keep) {
                double e[8];
                unsigned mk[8];
#pragma unroll
                for (int k = 0; k < 8; k++) { e[k] = err[q[k]]; mk[k] = marked[q[k]]; }
                double m = -kGREAT;
                unsigned anyMarked = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) { m = fmaxFoam(m, e[k]); anyMarked |= mk[k]; }
                kp = (m < unrefineLevel) && !anyMarked;
            }
        }
        if (isSplit) isSplit[pt] = split;
        if (keep) keep[pt] = kp;
    });
}
