// kernels_pipe.cuh -- TMA-pipelined variants of the dense SELL gather kernels (see pipeline.cuh).
// Same arithmetic as the direct kernels in kernels.cuh (which remain the path for meshes whose
// tiles do not fit the shared-memory ring and for the sparse sweeps); results are bit-identical.
#pragma once
#include "kernels.cuh"
#include "pipeline.cuh"

template <int KIND, bool SCALE>
__global__ void __launch_bounds__(kPipeThreads)
kp_estimate(SellView S, const double *__restrict__ x, const double *__restrict__ ghost, double p0, double p1,
            Thresholds thr, double *__restrict__ err) {
    const int64_t n = S.rows;
    const bool useOffset = fabs(p1) > kSMALL;
    sellPipeline(S, [&](int64_t c, bool valid, const int32_t *row, int w) {
        if (!valid) return;
        const double xc = fieldXform<KIND>(x[c], p1, useOffset);
        double e = 0.0;
        for (int j0 = 0; j0 < w; j0 += kChunk) {
            int32_t q[kChunk];
#pragma unroll
            for (int k = 0; k < kChunk; k++) q[k] = (j0 + k < w) ? row[(j0 + k) << 5] : (int32_t)c;
            double xq[kChunk];
#pragma unroll
            for (int k = 0; k < kChunk; k++) {
                if (q[k] < n) xq[k] = x[q[k]];
                else xq[k] = ghost[q[k] - n];
            }
#pragma unroll
            for (int k = 0; k < kChunk; k++) {
                if (q[k] == (int32_t)c) continue;
                double v = fieldXform<KIND>(xq[k], p1, useOffset);
                double eT;
                if (KIND == 0) {
                    eT = fabs(xc - v);
                    if (q[k] < n) eT = fminFoam(eT, 0.5);
                } else if (KIND == 1) {
                    eT = fabs(xc - v) / fmaxFoam(fminFoam(xc, v), p0);
                } else {
                    eT = fabs(xc - v);
                }
                e = fmaxFoam(e, eT);
            }
        }
        err[c] = SCALE ? normalizeValue(e, thr) : e;
    });
}

__global__ void __launch_bounds__(kPipeThreads)
kp_extend(SellView S, const u8 *__restrict__ in, const u8 *__restrict__ src, const u8 *__restrict__ ghost,
          u8 *__restrict__ out) {
    const int64_t n = S.rows;
    sellPipeline(S, [&](int64_t c, bool valid, const int32_t *row, int w) {
        unsigned m = valid ? in[c] : 1u;
        if (__all_sync(0xffffffffu, m != 0)) { if (valid) out[c] = 1; return; }
        if (!valid) return;
        for (int j0 = 0; j0 < w; j0 += kChunk) {
            int32_t q[kChunk];
#pragma unroll
            for (int k = 0; k < kChunk; k++) q[k] = (j0 + k < w) ? row[(j0 + k) << 5] : (int32_t)c;
#pragma unroll
            for (int k = 0; k < kChunk; k++) {
                if (q[k] == (int32_t)c) continue;
                m |= (q[k] < n) ? src[q[k]] : ghost[q[k] - n];
            }
        }
        out[c] = m ? 1 : 0;
    });
}

template <bool MAXSET>
__global__ void __launch_bounds__(kPipeThreads)
kp_refine_sweep(SellView S, u8 *__restrict__ lf, u8 *__restrict__ set, const u8 *__restrict__ ghostLf, int *changed) {
    const int64_t n = S.rows;
    sellPipeline(S, [&](int64_t c, bool valid, const int32_t *row, int w) {
        int flip = 0;
        const int isSet = valid ? set[c] : (MAXSET ? 1 : 0);
        const bool warpIdle = MAXSET ? __all_sync(0xffffffffu, isSet != 0) : !__any_sync(0xffffffffu, isSet != 0);
        if (!warpIdle && valid) {
            const int mine = lf[c];
            int ext = MAXSET ? 0 : 255;
            for (int j0 = 0; j0 < w; j0 += kChunk) {
                int32_t q[kChunk];
#pragma unroll
                for (int k = 0; k < kChunk; k++) q[k] = (j0 + k < w) ? row[(j0 + k) << 5] : (int32_t)c;
#pragma unroll
                for (int k = 0; k < kChunk; k++) {
                    int v = (q[k] < n) ? lf[q[k]] : ghostLf[q[k] - n];
                    ext = MAXSET ? max(ext, v) : min(ext, v);
                }
            }
            if (MAXSET) { if (!isSet && ext > mine + 1) { set[c] = 1; lf[c] = (u8)(mine + 1); flip = 1; } }
            else { if (isSet && mine > ext + 1) { set[c] = 0; lf[c] = (u8)(mine - 1); flip = 1; } }
        }
        if (__any_sync(0xffffffffu, flip) && (threadIdx.x & 31) == 0) *changed = 1;
    });
}

__global__ void __launch_bounds__(kPipeThreads)
kp_select_unrefine_points(SellView S, const u8 *__restrict__ pcCount, const u8 *__restrict__ bndPoint,
                          const label *__restrict__ parentIdx, const u8 *__restrict__ lvl8,
                          const double *__restrict__ err, double unrefineLevel, const u8 *__restrict__ marked,
                          u8 *__restrict__ isSplit, u8 *__restrict__ keep) {
    sellPipeline(S, [&](int64_t pt, bool valid, const int32_t *row, int w) {
        if (!valid) return;
        bool split = (pcCount[pt] == 8) && !bndPoint[pt];
        bool kp = false;
        if (split) {
            int32_t q[8];
#pragma unroll
            for (int k = 0; k < 8; k++) q[k] = row[k << 5];
            label par[8];
#pragma unroll
            for (int k = 0; k < 8; k++) par[k] = parentIdx[q[k]];
            bool same = par[0] >= 0;
#pragma unroll
            for (int k = 1; k < 8; k++) same = same && (par[k] == par[0]);
            if (same) {
                int l0 = lvl8[q[0]];
#pragma unroll
                for (int k = 1; k < 8; k++) same = same && (lvl8[q[k]] == l0);
            }
            split = same;
            if (split && keep) {
                double m = -kGREAT;
                unsigned anyMarked = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) { m = fmaxFoam(m, err[q[k]]); anyMarked |= marked[q[k]]; }
                kp = (m < unrefineLevel) && !anyMarked;
            }
        }
        if (isSplit) isSplit[pt] = split;
        if (keep) keep[pt] = kp;
    });
}
