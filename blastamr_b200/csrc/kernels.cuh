// kernels.cuh -- the hot kernels of the AMR decision path (sm_100a, HBM-bound gather work).
//
// Common shape: one thread per cell (or point), one warp per SELL-32 slice.  Column j of a
// slice is one coalesced 128-byte load of neighbour labels; the per-neighbour gathers
// (8-byte field values, 1-byte level/flag bytes) hit L1/L2 for any mesh numbering with
// locality.  Up to 8 columns are loaded before the dependent gathers are issued so that each
// thread keeps 8 independent requests in flight.  No floating-point atomics anywhere: every
// cell reduces its own row (max / OR), which is order-independent and therefore bit-identical
// to the reference's face loops.
#pragma once
#include "common.cuh"
#include "p2p.cuh"

constexpr int kChunk = 8;   // SELL columns fetched per batch (memory-level parallelism per thread)

struct Thresholds {
    double lowerRefine, upperRefine, lowerUnrefine, upperUnrefine;
};

// errorEstimator::normalize, errorEstimator.C:204-228
__device__ __forceinline__ double normalizeValue(double e, const Thresholds &t) {
    if (e < t.lowerUnrefine || e > t.upperUnrefine) return -1.0;
    if (e > t.lowerRefine && e < t.upperRefine) return 1.0;
    return 0.0;
}

// Foam::max / Foam::min (a > b ? a : b) -- spelled out so NaN handling matches the host code
__device__ __forceinline__ double fmaxFoam(double a, double b) { return (a > b) ? a : b; }
__device__ __forceinline__ double fminFoam(double a, double b) { return (a < b) ? a : b; }

// ------------------------------------------------------------------------------------------
// STAGE 1  delta / scaledDelta / coded delta (+ fused normalize)
//   delta.C:93-131, scaledDelta.C:78-136, errorEstimator.C:204-228
// KIND: 0 delta, 1 scaledDelta, 4 coded delta.  Ghost entries (q >= n) are coupled-patch faces:
// value = patchNeighbourField()[bface]; delta does not clamp there (delta.C:126).
// ------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ double fieldXform(double v, double offset, bool useOffset) {
    if (KIND == 1) {
        v = fabs(v);                 // getFieldValue -> mag(field), errorEstimatorTemplates.C:30-46
        if (useOffset) v -= offset;  // scaledDelta.C:89-93
    }
    return v;
}

template <int KIND, bool SCALE>
__global__ void __launch_bounds__(kThreads)
k_estimate(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
           const double *__restrict__ x, const u8 *__restrict__ cellGhost, double p0, double p1, Thresholds thr,
           double *__restrict__ err) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const int64_t s = c >> 5;
    const int lane = (int)(c & 31);
    const int32_t base = sliceOff[s];
    const int w = sliceOff[s + 1] - base;
    const int32_t *row = col + ((size_t)base << 5) + lane;
    const bool useOffset = fabs(p1) > kSMALL;
    const double xc = fieldXform<KIND>(x[c], p1, useOffset);
    double e = 0.0;
    for (int j0 = 0; j0 < w; j0 += kChunk) {
        int32_t q[kChunk];
#pragma unroll
        for (int k = 0; k < kChunk; k++) q[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : (int32_t)c;
        double xq[kChunk];
#pragma unroll
        for (int k = 0; k < kChunk; k++) xq[k] = x[q[k]];   // internal neighbours only; coupled faces: k_fix_estimate
#pragma unroll
        for (int k = 0; k < kChunk; k++) {
            if (q[k] == (int32_t)c) continue;   // padding
            double v = fieldXform<KIND>(xq[k], p1, useOffset);
            double eT;
            if (KIND == 0) {
                eT = fabs(xc - v);
                if (q[k] < n) eT = fminFoam(eT, 0.5);          // delta.C:103 vs :126
            } else if (KIND == 1) {
                eT = fabs(xc - v) / fmaxFoam(fminFoam(xc, v), p0);   // scaledDelta.C:103-104,129-131
            } else {
                eT = fabs(xc - v);
            }
            e = fmaxFoam(e, eT);
        }
    }
    err[c] = (SCALE && !(cellGhost && cellGhost[c])) ? normalizeValue(e, thr) : e;
}

// fieldValue::update, fieldValue.C:68-82 (+ normalize); also the plain normalize pass (ABS=false)
template <bool ABS, bool SCALE>
__global__ void __launch_bounds__(kThreads)
k_pointwise(int64_t n, const double *in, Thresholds thr, double *out) {       // in == out for the plain normalize pass: no __restrict__
    int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    double v = in[c];
    if (ABS) v = fabs(v);
    out[c] = SCALE ? normalizeValue(v, thr) : v;
}

// gMin / gMax of a field (gradientRange.C:102-103): per-CTA partials, exact (min/max are order-free)
__global__ void __launch_bounds__(kThreads)
k_minmax_partial(int64_t n, const double *__restrict__ v, double *__restrict__ part) {
    __shared__ double smn[kThreads / 32], smx[kThreads / 32];
    double mn = kGREAT * 1e285, mx = -kGREAT * 1e285;   // +-1e300: outside any field value
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const double x = v[i];
        mn = fminFoam(mn, x);
        mx = fmaxFoam(mx, x);
    }
    for (int d = 16; d > 0; d >>= 1) {
        mn = fminFoam(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = fmaxFoam(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < kThreads / 32; i++) { mn = fminFoam(mn, smn[i]); mx = fmaxFoam(mx, smx[i]); }
        part[2 * blockIdx.x] = mn; part[2 * blockIdx.x + 1] = mx;
    }
}
__global__ void k_minmax_final(int nPart, const double *__restrict__ part, double *__restrict__ out) {
    double mn = part[0], mx = part[1];
    for (int i = 1; i < nPart; i++) { mn = fminFoam(mn, part[2 * i]); mx = fmaxFoam(mx, part[2 * i + 1]); }
    out[0] = mn; out[1] = mx;
}

// patchInternalField of the field for every boundary face: send buffer of the swap
__global__ void k_pack_f64(int64_t b, const label *__restrict__ bOwner, const double *__restrict__ x,
                           double *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b) out[i] = x[bOwner[i]];
}

// ------------------------------------------------------------------------------------------
// STAGE 2  marking
// ------------------------------------------------------------------------------------------

// fvMeshRefiner::selectRefineCandidates, fvMeshRefiner.C:170-187 (inclusive compares), fused with
// the "can still spread" test of the first forced buffer layer (fvMeshRefiner.C:809-813):
//   mark[c] = lo <= e <= hi ;  src[c] = mark[c] && maxLevel(c) > cellLevel[c]
__global__ void __launch_bounds__(kThreads)
k_candidates(int64_t n, const double *__restrict__ err, double lo, double hi, const u8 *__restrict__ lvl,
             const label *__restrict__ maxLvl, label maxLvlUniform, u8 *__restrict__ mark, u8 *__restrict__ src,
             bool aligned) {
    // 4 consecutive cells per thread: 2 x 16 B of error, 16 B of levels in flight, 4 B stores
    const int64_t c0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    if (c0 >= n) return;
    if (aligned && c0 + 4 <= n) {
        const double2 e01 = *reinterpret_cast<const double2 *>(err + c0);
        const double2 e23 = *reinterpret_cast<const double2 *>(err + c0 + 2);
        const double e[4] = {e01.x, e01.y, e23.x, e23.y};
        uchar4 m;
        m.x = (e[0] >= lo) && (e[0] <= hi); m.y = (e[1] >= lo) && (e[1] <= hi);
        m.z = (e[2] >= lo) && (e[2] <= hi); m.w = (e[3] >= lo) && (e[3] <= hi);
        *reinterpret_cast<uchar4 *>(mark + c0) = m;
        if (src) {
            const uchar4 l = *reinterpret_cast<const uchar4 *>(lvl + c0);
            int4 mx = make_int4(maxLvlUniform, maxLvlUniform, maxLvlUniform, maxLvlUniform);
            if (maxLvl) mx = *reinterpret_cast<const int4 *>(maxLvl + c0);
            uchar4 sr;
            sr.x = m.x && (mx.x > (int)l.x); sr.y = m.y && (mx.y > (int)l.y); sr.z = m.z && (mx.z > (int)l.z); sr.w = m.w && (mx.w > (int)l.w);
            *reinterpret_cast<uchar4 *>(src + c0) = sr;
        }
        return;
    }
    for (int64_t c = c0; c < n && c < c0 + 4; c++) {
        double e = err[c];
        u8 m = (e >= lo) && (e <= hi);
        mark[c] = m;
        if (src) {
            label mx = maxLvl ? maxLvl[c] : maxLvlUniform;
            src[c] = m && (mx > (label)lvl[c]);
        }
    }
}

// selectRefineCells filter for a uniform maxRefinement (the case of every config): 16 cells per
// thread on packed bytes (SIMD-in-word compares), lf = level + set.
__global__ void __launch_bounds__(kThreads)
k_select_filter16(int64_t n, const u8 *__restrict__ cand, const u8 *__restrict__ lvl8, int maxLvlUniform,
                  const u8 *__restrict__ unrefineable, u8 *__restrict__ set, u8 *__restrict__ lf) {
    const int64_t c0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 16;
    if (c0 >= n) return;
    const unsigned mxb = (unsigned)min(max(maxLvlUniform, 0), 255) * 0x01010101u;
    if (c0 + 16 <= n) {
        const uint4 c4 = *reinterpret_cast<const uint4 *>(cand + c0);
        const uint4 l4 = *reinterpret_cast<const uint4 *>(lvl8 + c0);
        uint4 u4 = make_uint4(0, 0, 0, 0);
        if (unrefineable) u4 = *reinterpret_cast<const uint4 *>(unrefineable + c0);
        auto one = [&](unsigned c, unsigned l, unsigned u, unsigned &sel, unsigned &lfv) {
            const unsigned lt = __vcmpltu4(l, mxb) & 0x01010101u;       // level < maxRefinement
            const unsigned nz = __vcmpne4(c, 0u) & 0x01010101u;         // candidate
            const unsigned nu = __vcmpeq4(u, 0u) & 0x01010101u;         // !unrefineable
            sel = lt & nz & nu;
            lfv = __vadd4(l, sel);
        };
        uint4 s4, f4;
        one(c4.x, l4.x, u4.x, s4.x, f4.x); one(c4.y, l4.y, u4.y, s4.y, f4.y);
        one(c4.z, l4.z, u4.z, s4.z, f4.z); one(c4.w, l4.w, u4.w, s4.w, f4.w);
        *reinterpret_cast<uint4 *>(set + c0) = s4;
        *reinterpret_cast<uint4 *>(lf + c0) = f4;
        return;
    }
    for (int64_t c = c0; c < n; c++) {
        const u8 sel = cand[c] && ((int)lvl8[c] < maxLvlUniform) && !(unrefineable && unrefineable[c]);
        set[c] = sel;
        lf[c] = (u8)(lvl8[c] + sel);
    }
}

// values of a byte array at the owners of all boundary faces (send buffer of syncFaceList /
// swapBoundaryFaceList for byte payloads)
__global__ void k_pack_u8(int64_t b, const label *__restrict__ bOwner, const u8 *__restrict__ v,
                          u8 *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b) out[i] = v[bOwner[i]];
}

// clear / set flags of the cells on the faces of a patch (fvMeshHexRefiner.C:1514-1523, 1612-1621,
// errorEstimator.C:336-344)
__global__ void k_patch_cells_set(const label *__restrict__ owner, int64_t start, int64_t size, u8 *__restrict__ flag,
                                  u8 value) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < size) flag[owner[start + i]] = value;
}

// flag[c] = value where mask[c]  (protectedCell_ handling, fvMeshHexRefiner.C:1524-1533, 1603-1611)
__global__ void k_masked_set(int64_t n, const u8 *__restrict__ mask, u8 *__restrict__ flag, u8 value) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && mask[c]) flag[c] = value;
}

// error[c] = 0 where prot[c]; counts protected cells (errorEstimator.C:352-361)
__global__ void __launch_bounds__(kThreads)
k_zero_protected(int64_t n, const u8 *__restrict__ prot, double *__restrict__ err, unsigned long long *count) {
    int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int hit = (c < n) && prot[c];
    if (hit) err[c] = 0.0;
    int cnt = __syncthreads_count(hit);
    if (threadIdx.x == 0 && cnt) atomicAdd(count, (unsigned long long)cnt);
}

// fvMeshHexRefiner::selectRefineCells filter (fvMeshHexRefiner.C:328-345):
//   set[c] = cellLevel < maxRefinement && candidate && !unrefineable ;  lf[c] = cellLevel + set
// LEVEL >= 0 restricts to cellLevel == LEVEL and ORs into set (the maxCells truncation loop :346-374).
__global__ void __launch_bounds__(kThreads)
k_select_filter(int64_t n, const u8 *cand, const label *__restrict__ lvl, const label *__restrict__ maxLvl,
                label maxLvlUniform, const u8 *__restrict__ unrefineable, int onlyLevel, u8 *set,
                u8 *__restrict__ lf, bool aligned) {                               // cand may be set (amr_consistent_refinement): no __restrict__ on the pair
    const int64_t c0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    if (c0 >= n) return;
    const bool vec = aligned && c0 + 4 <= n;
    label l[4], mx[4];
    u8 cd[4], ur[4] = {0, 0, 0, 0}, st[4] = {0, 0, 0, 0};
    const int cnt = vec ? 4 : (int)min((int64_t)4, n - c0);
    if (vec) {
        const int4 lv = *reinterpret_cast<const int4 *>(lvl + c0);
        l[0] = lv.x; l[1] = lv.y; l[2] = lv.z; l[3] = lv.w;
        const uchar4 c4 = *reinterpret_cast<const uchar4 *>(cand + c0);
        cd[0] = c4.x; cd[1] = c4.y; cd[2] = c4.z; cd[3] = c4.w;
        if (maxLvl) { const int4 m4 = *reinterpret_cast<const int4 *>(maxLvl + c0); mx[0] = m4.x; mx[1] = m4.y; mx[2] = m4.z; mx[3] = m4.w; }
        if (unrefineable) { const uchar4 u4 = *reinterpret_cast<const uchar4 *>(unrefineable + c0); ur[0] = u4.x; ur[1] = u4.y; ur[2] = u4.z; ur[3] = u4.w; }
        if (onlyLevel >= 0) { const uchar4 s4 = *reinterpret_cast<const uchar4 *>(set + c0); st[0] = s4.x; st[1] = s4.y; st[2] = s4.z; st[3] = s4.w; }
    } else {
        for (int k = 0; k < cnt; k++) {
            l[k] = lvl[c0 + k]; cd[k] = cand[c0 + k];
            if (maxLvl) mx[k] = maxLvl[c0 + k];
            if (unrefineable) ur[k] = unrefineable[c0 + k];
            if (onlyLevel >= 0) st[k] = set[c0 + k];
        }
    }
    u8 sel[4], lfv[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const label m = maxLvl ? mx[k] : maxLvlUniform;
        if (onlyLevel < 0) sel[k] = cd[k] && (l[k] < m) && !ur[k];
        else sel[k] = st[k] || (cd[k] && l[k] == onlyLevel && !ur[k]);
        lfv[k] = (u8)(l[k] + sel[k]);
    }
    if (vec) {
        *reinterpret_cast<uchar4 *>(set + c0) = make_uchar4(sel[0], sel[1], sel[2], sel[3]);
        if (lf) *reinterpret_cast<uchar4 *>(lf + c0) = make_uchar4(lfv[0], lfv[1], lfv[2], lfv[3]);
    } else {
        for (int k = 0; k < cnt; k++) { set[c0 + k] = sel[k]; if (lf) lf[c0 + k] = lfv[k]; }
    }
}

// sum of a byte-flag array (returnReduce(count(...)) sites: fvMeshHexRefiner.C:323-324,369)
__global__ void __launch_bounds__(kThreads)
k_count_flags(int64_t n, const u8 *__restrict__ flag, unsigned long long *count) {
    int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int cnt = __syncthreads_count((c < n) && flag[c]);
    if (threadIdx.x == 0 && cnt) atomicAdd(count, (unsigned long long)cnt);
}

// ------------------------------------------------------------------------------------------
// STAGE 3  2:1 level-consistency sweeps
// ------------------------------------------------------------------------------------------

// lf[c] = cellLevel[c] + set[c] at the owners of boundary faces: payload of
// syncTools::swapBoundaryFaceList(neiLevel), hexRef.C:745-756 -- uses k_pack_u8 on lf.

// hexRef::faceConsistentRefinement (hexRef.C:700-783), maxSet = true, as a chaotic in-place
// iteration of the same monotone rule: a cell is added when a face neighbour would end up more
// than one level finer.  Flags only grow and every addition is forced, so any sweep order
// converges to the reference's (unique, least) closure; only the sweep COUNT differs from the
// reference's Gauss-Seidel pass.  `changed` receives 1 if any cell of the grid flipped
// (__syncthreads_or + one store per CTA: the reduce(nChanged) of hexRef.C:1424).
__global__ void __launch_bounds__(kThreads)
k_refine_sweep_max(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                   u8 *__restrict__ lf, u8 *__restrict__ set, int *changed, const int *__restrict__ gate) {
    if (gate && *gate == 1) return;                 // the sparse (push) form runs instead (k_refine_push)
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const unsigned live = __ballot_sync(0xffffffffu, c < n);
    int flip = 0;
    if (c < n) {
        const int64_t s = c >> 5;
        const int lane = (int)(c & 31);
        const int32_t base = sliceOff[s];
        const int w = sliceOff[s + 1] - base;
        const int32_t *row = col + ((size_t)base << 5) + lane;
        const int mine = lf[c];
        const int isSet = set[c];
        // marked cells cannot change; a warp of 32 marked cells skips its gathers
        if (!__all_sync(live, isSet != 0)) {
            int mx = 0;
            for (int j0 = 0; j0 < w; j0 += kChunk) {
                int32_t q[kChunk];
#pragma unroll
                for (int k = 0; k < kChunk; k++) q[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : (int32_t)c;
#pragma unroll
                for (int k = 0; k < kChunk; k++) {
                    mx = max(mx, (int)lf[q[k]]);
                }
            }
            if (!isSet && mx > mine + 1) { set[c] = 1; lf[c] = (u8)(mine + 1); flip = 1; }
        }
    }
    if (__syncthreads_or(flip) && threadIdx.x == 0) *changed = 1;
}

// maxSet = false variant (hexRef.C:723-741 else-branches): a marked cell is removed when it would
// end up more than one level finer than a face neighbour.  Monotone decreasing -> greatest fixed point.
__global__ void __launch_bounds__(kThreads)
k_refine_sweep_min(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                   u8 *__restrict__ lf, u8 *__restrict__ set, int *changed) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const unsigned live = __ballot_sync(0xffffffffu, c < n);
    int flip = 0;
    if (c < n) {
        const int64_t s = c >> 5;
        const int lane = (int)(c & 31);
        const int32_t base = sliceOff[s];
        const int w = sliceOff[s + 1] - base;
        const int32_t *row = col + ((size_t)base << 5) + lane;
        const int mine = lf[c];
        const int isSet = set[c];
        if (__any_sync(live, isSet != 0)) {
            int mn = 255;
            for (int j0 = 0; j0 < w; j0 += kChunk) {
                int32_t q[kChunk];
#pragma unroll
                for (int k = 0; k < kChunk; k++) q[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : (int32_t)c;
#pragma unroll
                for (int k = 0; k < kChunk; k++) {
                    mn = min(mn, (int)lf[q[k]]);
                }
            }
            if (isSet && mine > mn + 1) { set[c] = 0; lf[c] = (u8)(mine - 1); flip = 1; }
        }
    }
    if (__syncthreads_or(flip) && threadIdx.x == 0) *changed = 1;
}

// hexRef::checkRefinementLevels 2:1 part, hexRef.C:2962-3025: count faces with |dL| > 1
// (each internal face is seen from both cells -> counted once from the lower label; coupled
// faces once from their owner, like the reference's boundary loop)
__global__ void __launch_bounds__(kThreads)
k_check_levels(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
               const label *__restrict__ lvl, unsigned long long *bad) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int cnt = 0;
    if (c < n) {
        const int64_t s = c >> 5;
        const int lane = (int)(c & 31);
        const int32_t base = sliceOff[s];
        const int w = sliceOff[s + 1] - base;
        const int32_t *row = col + ((size_t)base << 5) + lane;
        const int mine = lvl[c];
        for (int j = 0; j < w; j++) {
            int32_t q = row[(size_t)j << 5];
            if (q == (int32_t)c) continue;
            if (q > c && abs(mine - lvl[q]) > 1) cnt++;     // each internal face once; coupled faces: k_fix_check_levels
        }
    }
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(bad, (unsigned long long)cnt);
}
__global__ void k_pack_i32(int64_t b, const label *__restrict__ bOwner, const label *__restrict__ v,
                           label *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b) out[i] = v[bOwner[i]];
}

// ------------------------------------------------------------------------------------------
// STAGE 3b  unrefinement selection (3-D hex): point-centric
// ------------------------------------------------------------------------------------------

// history_.parentIndex(c) (hexRefRefinementHistory.H:299-310) or -1 when the cell is not
// visible / has no parent: precomputed once per levels_set so the hot kernel gathers 4 bytes.
__global__ void k_parent_index(int64_t n, const label *__restrict__ visible, const label *__restrict__ splitParent,
                               int64_t nSplit, label *__restrict__ parentIdx) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    label v = visible[c];
    label pi = -1;
    if (v >= 0 && v < nSplit) { pi = splitParent[v]; if (pi < 0) pi = -1; }
    parentIdx[c] = pi;
}
__global__ void k_lvl_to_u8(int64_t n, const label *__restrict__ lvl, u8 *__restrict__ out) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) out[c] = (u8)lvl[c];
}

// fvMeshRefiner::maxCellField, fvMeshRefiner.C:110-125
__global__ void __launch_bounds__(kThreads)
k_max_cell_field(int64_t p, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                 const double *__restrict__ v, double *__restrict__ out) {
    const int64_t pt = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (pt >= p) return;
    const int32_t base = sliceOff[pt >> 5];
    const int w = sliceOff[(pt >> 5) + 1] - base;
    const int32_t *row = col + ((size_t)base << 5) + (pt & 31);
    double m = -kGREAT;
    for (int j = 0; j < w; j++) {
        int32_t q = row[(size_t)j << 5];
        if (q >= 0) m = fmaxFoam(m, v[q]);
    }
    out[pt] = m;
}

// Fused hexRef3D::getSplitElems (hexRef3D.C:2183-2321) + selectUnrefineElems filter
// (hexRef3D.C:2336-2361) + maxCellField (fvMeshRefiner.C:110-125), one thread per point:
//   split   <=> |pointCells| == 8, no boundary face uses the point, every cell is visible with
//               parentIndex >= 0, same parentIndex and same cellLevel for all 8
//   keep    <=> split && max_{cells} error < unrefineLevel && no cell marked
// isSplit / keep may be NULL.  Rejects as early as possible: most points of a mesh are not
// split points and cost one SELL column only.
__global__ void __launch_bounds__(kThreads)
k_select_unrefine_points(int64_t p, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                         const u8 *__restrict__ pcClass, const label *__restrict__ parentIdx, const u8 *__restrict__ lvl8,
                         const double *__restrict__ err, double unrefineLevel, const u8 *__restrict__ marked,
                         u8 *__restrict__ isSplit, u8 *__restrict__ keep) {
    const int64_t pt = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (pt >= p) return;
    bool split = pcClass[pt] != 0;      // |pointCells| == 8, no boundary face, (pointLevel >= 1 when known)
    bool kp = false;
    if (split) {
        const int32_t base = sliceOff[pt >> 5];
        const int32_t *row = col + ((size_t)base << 5) + (pt & 31);
        int32_t q[8];
#pragma unroll
        for (int k = 0; k < 8; k++) q[k] = row[(size_t)k << 5];
        label par[8];
        int l[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { par[k] = parentIdx[q[k]]; l[k] = lvl8[q[k]]; }
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
}

// same, over a compact ascending list of the candidate points (sparse case: most points of a mesh cannot be
// split points, so neither their rows nor empty CTAs are touched); non-candidates keep the pre-zeroed flag
__global__ void __launch_bounds__(kThreads)
k_select_unrefine_cand(int64_t nCand, const label *__restrict__ cand, const int32_t *__restrict__ sliceOff,
                       const int32_t *__restrict__ col, const label *__restrict__ parentIdx, const u8 *__restrict__ lvl8,
                       const double *__restrict__ err, double unrefineLevel, const u8 *__restrict__ marked,
                       u8 *__restrict__ isSplit, u8 *__restrict__ keep) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= nCand) return;
    const int64_t pt = cand[i];
    const int32_t *row = col + ((size_t)sliceOff[pt >> 5] << 5) + (pt & 31);
    int32_t q[8];
#pragma unroll
    for (int k = 0; k < 8; k++) q[k] = row[(size_t)k << 5];
    label par[8];
    int l[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { par[k] = parentIdx[q[k]]; l[k] = lvl8[q[k]]; }
    bool split = par[0] >= 0;
#pragma unroll
    for (int k = 1; k < 8; k++) split = split && (par[k] == par[0]) && (l[k] == l[0]);
    bool kp = false;
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
    if (isSplit && split) isSplit[pt] = 1;
    if (keep && kp) keep[pt] = 1;
}

// per-point pre-filter of getSplitElems, mesh-constant: exactly 8 pointCells (hexRef3D.C:2214-2222), not on a
// boundary face (:2279-2294) and -- when hexRef's pointLevel_ is supplied -- pointLevel >= 1: a point whose 8
// cells share one parent is that parent's centre point, which setRefinement created with
// pointLevel = cellLevel + 1 (hexRef3D.C:1238-1260), so level-0 points can never pass the history test.
__global__ void __launch_bounds__(kThreads)
k_point_class(int64_t p, const u8 *__restrict__ pcCount, const u8 *__restrict__ bndPoint, const label *__restrict__ pointLevel,
              u8 *__restrict__ pcClass, unsigned long long *nCand) {
    const int64_t pt = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int c = 0;
    if (pt < p) {
        c = (pcCount[pt] == 8) && !bndPoint[pt] && (!pointLevel || pointLevel[pt] >= 1);
        pcClass[pt] = (u8)c;
    }
    const int cnt = __syncthreads_count(c);
    if (threadIdx.x == 0 && cnt) atomicAdd(nCand, (unsigned long long)cnt);
}

// hexRef2D::getSplitElems (hexRef2D.C:2202-2342) + selectUnrefineElems filter (hexRef2D.C:1860-1906), one
// thread per edge: split <=> |edgeCells| == 4, not an edge of a boundary face, the 4 cells visible with the
// same parentIndex >= 0 and level; kept when EITHER end point has max error < unrefineLevel and no marked
// pointCell (the 2-D quirk).  2-D meshes are small: plain gathers, no staging.
__global__ void __launch_bounds__(kThreads)
k_select_unrefine_edges(int64_t nEdges, const int32_t *__restrict__ ecOff, const int32_t *__restrict__ ecCol,
                        const u8 *__restrict__ ecCount, const u8 *__restrict__ bndEdge, const label *__restrict__ edgePts,
                        const int32_t *__restrict__ pcOff, const int32_t *__restrict__ pcCol,
                        const label *__restrict__ parentIdx, const u8 *__restrict__ lvl8, const double *__restrict__ err,
                        double unrefineLevel, const u8 *__restrict__ marked, u8 *__restrict__ isSplit, u8 *__restrict__ keep) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= nEdges) return;
    bool split = (ecCount[e] == 4) && !bndEdge[e];
    bool kp = false;
    if (split) {
        const int32_t *row = ecCol + ((size_t)ecOff[e >> 5] << 5) + (e & 31);
        int32_t q[4];
#pragma unroll
        for (int k = 0; k < 4; k++) q[k] = row[(size_t)k << 5];
        const label p0 = parentIdx[q[0]];
        const int l0 = lvl8[q[0]];
        bool same = p0 >= 0;
#pragma unroll
        for (int k = 1; k < 4; k++) same = same && (parentIdx[q[k]] == p0) && (lvl8[q[k]] == l0);
        split = same;
        if (split && keep) {
            for (int k = 0; k < 2 && !kp; k++) {
                const label pt = edgePts[2 * e + k];
                const int32_t base = pcOff[pt >> 5];
                const int w = pcOff[(pt >> 5) + 1] - base;
                const int32_t *prow = pcCol + ((size_t)base << 5) + (pt & 31);
                double m = -kGREAT;
                unsigned anyMarked = 0;
                for (int j = 0; j < w; j++) {
                    const int32_t c = prow[(size_t)j << 5];
                    if (c < 0) continue;
                    m = fmaxFoam(m, err[c]);
                    anyMarked |= marked[c];
                }
                if (m < unrefineLevel && !anyMarked) kp = true;
            }
        }
    }
    if (isSplit) isSplit[e] = split;
    if (keep) keep[e] = kp;
}

// ---- polyRefiner -------------------------------------------------------------------------------------

// refinement::edgeConsistentRefinement (refinement.C:305-387), one thread per mesh edge: every unordered pair
// of the edge's cells must satisfy the 2:1 rule on level+flag.  maxSet: a cell more than one level below the
// edge's maximum is added; !maxSet: a marked cell more than one level above the edge's minimum is removed.
// Same monotone closure as the face rule, so in-place chaotic iteration reaches the reference's fixed point.
// up to 8 labels of a SELL row starting at entry j0, loaded together (independent loads); entries past the row are -1
__device__ __forceinline__ void rowLabels8(const int32_t *row, int w, int j0, int32_t (&c)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : -1;
}

template <bool MAXSET>
__global__ void __launch_bounds__(kThreads)
k_edge_refine_sweep(int64_t nEdges, const int32_t *__restrict__ ecOff, const int32_t *__restrict__ ecCol,
                    u8 *lf, u8 *set, int *changed) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= nEdges) return;
    const int32_t base = ecOff[e >> 5];
    const int w = ecOff[(e >> 5) + 1] - base;
    const int32_t *row = ecCol + ((size_t)base << 5) + (e & 31);
    // labels and values are gathered 8 at a time (two dependent rounds per batch instead of two per cell); an edge of a
    // hex mesh has at most 4 cells plus hanging-node partners, so one batch is the usual case
    int mx = 0, mn = 255;
    int32_t c[8];
    int v[8];
    for (int j0 = 0; j0 < w; j0 += 8) {
        rowLabels8(row, w, j0, c);
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = c[k] >= 0 ? (int)lf[c[k]] : -1;
#pragma unroll
        for (int k = 0; k < 8; k++) if (v[k] >= 0) { mx = max(mx, v[k]); mn = min(mn, v[k]); }
    }
    if (mx - mn <= 1) return;
    for (int j0 = 0; j0 < w; j0 += 8) {
        if (w > 8) {                      // (a single batch is still in registers)
            rowLabels8(row, w, j0, c);
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = c[k] >= 0 ? (int)lf[c[k]] : -1;
        }
        unsigned st[8];
#pragma unroll
        for (int k = 0; k < 8; k++) st[k] = c[k] >= 0 ? (unsigned)set[c[k]] : 0u;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (c[k] < 0) continue;
            if (MAXSET) { if (!st[k] && v[k] + 1 < mx) { set[c[k]] = 1; lf[c[k]] = (u8)(v[k] + 1); *changed = 1; } }
            else { if (st[k] && v[k] > mn + 1) { set[c[k]] = 0; lf[c[k]] = (u8)(v[k] - 1); *changed = 1; } }
        }
    }
}

// refinement::edgeConsistentUnrefinement (refinement.C:511-604), maxSet = false: a cell marked for
// unrefinement is unmarked when it would end up more than one level below another cell of the edge.
__global__ void __launch_bounds__(kThreads)
k_edge_unrefine_sweep(int64_t nEdges, const int32_t *__restrict__ ecOff, const int32_t *__restrict__ ecCol,
                      const u8 *__restrict__ lvl8, u8 *uc, int *changed) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= nEdges) return;
    const int32_t base = ecOff[e >> 5];
    const int w = ecOff[(e >> 5) + 1] - base;
    const int32_t *row = ecCol + ((size_t)base << 5) + (e & 31);
    int mx = 0;
    unsigned any = 0;
    int32_t c[8];
    int u[8], l[8];
    for (int j0 = 0; j0 < w; j0 += 8) {
        rowLabels8(row, w, j0, c);
#pragma unroll
        for (int k = 0; k < 8; k++) { u[k] = c[k] >= 0 ? (int)uc[c[k]] : 0; l[k] = c[k] >= 0 ? (int)lvl8[c[k]] : 0; }
#pragma unroll
        for (int k = 0; k < 8; k++) if (c[k] >= 0) { any |= (unsigned)u[k]; mx = max(mx, l[k] - u[k]); }
    }
    if (!any) return;
    for (int j0 = 0; j0 < w; j0 += 8) {
        if (w > 8) {
            rowLabels8(row, w, j0, c);
#pragma unroll
            for (int k = 0; k < 8; k++) { u[k] = c[k] >= 0 ? (int)uc[c[k]] : 0; l[k] = c[k] >= 0 ? (int)lvl8[c[k]] : 0; }
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (c[k] >= 0 && u[k] && l[k] - 1 < mx - 1) { uc[c[k]] = 0; *changed = 1; }
    }
}

// syncTools::syncPointList(markedPoint, orEqOp<bool>()) (fvMeshRefiner.C:952-958), receive side: OR the neighbour's
// values into the points of my processor patches (same point order on both sides)
__global__ void __launch_bounds__(kThreads)
k_or_scatter(int64_t count, const label *__restrict__ idx, const u8 *__restrict__ recv, u8 *__restrict__ pflag, int *changed) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= count) return;
    if (recv[i] && !pflag[idx[i]]) { pflag[idx[i]] = 1; *changed = 1; }
}

// first half of fvMeshRefiner::extendMarkedCellsAcrossPoints (fvMeshRefiner.C:924-950): a point is marked
// when any of its cells is (the second half, points -> cells, is k_points_to_cells)
__global__ void __launch_bounds__(kThreads)
k_cells_to_points_or(int64_t p, const int32_t *__restrict__ pcOff, const int32_t *__restrict__ pcCol,
                     const u8 *__restrict__ marked, u8 *__restrict__ pflag, u8 *__restrict__ pneed) {
    const int64_t pt = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (pt >= p) return;
    const int32_t base = pcOff[pt >> 5];
    const int w = pcOff[(pt >> 5) + 1] - base;
    const int32_t *row = pcCol + ((size_t)base << 5) + (pt & 31);
    unsigned m = 0, un = 0;                   // a marked cell / an unmarked cell among the point's cells
    for (int j0 = 0; j0 < w; j0 += 8) {
        int32_t c[8];
        rowLabels8(row, w, j0, c);
        unsigned v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = c[k] >= 0 ? (unsigned)marked[c[k]] : 0xffffu;
#pragma unroll
        for (int k = 0; k < 8; k++) { if (v[k] == 0u) un = 1; else if (v[k] != 0xffffu) m = 1; }
    }
    pflag[pt] = m ? 1 : 0;
    if (pneed) pneed[pt] = un ? 1 : 0;
}

// polyhedralRefinement split-point markup, face-centric (polyhedralRefinement.C:2110-2184): every point of a
// face whose points do not share one pointLevel, and every point of a boundary face, cannot be a split point
__global__ void __launch_bounds__(kThreads)
k_face_point_levels(int64_t nFaces, int64_t f, const label *__restrict__ faceOff, const label *__restrict__ facePts,
                    const label *__restrict__ pointLevel, u8 *__restrict__ bad) {
    const int64_t face = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (face >= nFaces) return;
    const label s = faceOff[face], e = faceOff[face + 1];
    bool reject = face >= f;
    const label l0 = pointLevel[facePts[s]];
    for (label j = s + 1; j < e && !reject; j++) reject = pointLevel[facePts[j]] != l0;
    if (reject) for (label j = s; j < e; j++) bad[facePts[j]] = 1;
}

// prismatic2DRefinement split-point markup, edge-centric (prismatic2DRefinement.C:2756-2817): a point whose
// pointEdges join two different pointLevels cannot be a split point
__global__ void __launch_bounds__(kThreads)
k_edge_point_levels(int64_t nEdges, const label *__restrict__ edgePts, const label *__restrict__ pointLevel, u8 *__restrict__ bad) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= nEdges) return;
    const int2 ab = reinterpret_cast<const int2 *>(edgePts)[e];
    if (pointLevel[ab.x] != pointLevel[ab.y]) { bad[ab.x] = 1; bad[ab.y] = 1; }
}
// ... and neither can a point of a processor-patch face (prismatic2DRefinement.C:2829-2849)
__global__ void __launch_bounds__(kThreads)
k_coupled_face_points(int64_t nCoupled, const label *__restrict__ cplFace, int64_t f, const label *__restrict__ faceOff,
                      const label *__restrict__ facePts, u8 *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= nCoupled) return;
    const int64_t face = f + cplFace[i];
    for (label j = faceOff[face]; j < faceOff[face + 1]; j++) bad[facePts[j]] = 1;
}

// fvMeshPolyRefiner::selectUnrefinePoints (fvMeshPolyRefiner.C:48-79) intersected with the split points
// (polyhedralRefinement.C:2196-2205)
__global__ void __launch_bounds__(kThreads)
k_poly_select_points(int64_t p, const int32_t *__restrict__ pcOff, const int32_t *__restrict__ pcCol,
                     const label *__restrict__ pointLevel, const u8 *__restrict__ bad, const double *__restrict__ err,
                     double unrefineLevel, const u8 *__restrict__ marked, u8 *__restrict__ isSplit, u8 *__restrict__ keep) {
    const int64_t pt = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (pt >= p) return;
    const bool split = pointLevel[pt] >= 1 && !bad[pt];
    bool kp = false;
    if (split && keep) {
        const int32_t base = pcOff[pt >> 5];
        const int w = pcOff[(pt >> 5) + 1] - base;
        const int32_t *row = pcCol + ((size_t)base << 5) + (pt & 31);
        double m = -kGREAT;
        unsigned anyMarked = 0;
        for (int j = 0; j < w; j++) {
            const int32_t c = row[(size_t)j << 5];
            if (c < 0) continue;
            m = fmaxFoam(m, err[c]);
            anyMarked |= marked[c];
        }
        kp = (m < unrefineLevel) && !anyMarked;
    }
    if (isSplit) isSplit[pt] = split;
    if (keep) keep[pt] = kp;
}

// ---- sparse passes of consistentUnrefinement --------------------------------------------------
// Candidate split points / cells marked for unrefinement are a small fraction of the mesh, so these
// passes scan 16 flags per thread with one 128-bit load and only walk the rows of set entries.
template <class Fn>
__device__ __forceinline__ void forEachSet16(const u8 *flags, int64_t N, bool aligned, Fn fn) {
    const int64_t base = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 16;
    if (base >= N) return;
    if (aligned && base + 16 <= N) {
        const uint4 w = *reinterpret_cast<const uint4 *>(flags + base);
        if ((w.x | w.y | w.z | w.w) == 0) return;
        const unsigned v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 16; k++)
            if ((v[k >> 2] >> (8 * (k & 3))) & 0xffu) fn(base + k);
    } else {
        for (int k = 0; k < 16 && base + k < N; k++)
            if (flags[base + k]) fn(base + k);
    }
}

// Warp-balanced form: the 512 flags of a warp are scanned as above, the indices of the set ones are gathered into a
// per-warp shared buffer and then dealt round-robin to the lanes.  Runs of consecutive set entries (the usual shape
// of an AMR marking) would otherwise be walked serially by ONE thread with its 31 neighbours idle; dealt out, their
// rows are also read coalesced.  Needs blockDim.x == kThreads.
template <class Fn>
__device__ __forceinline__ void forEachSetWarpAt(unsigned block, const u8 *flags, int64_t N, bool aligned, Fn fn) {
    __shared__ unsigned short setIdx[kThreads / 32][512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warpBase = ((int64_t)block * kThreads + (threadIdx.x & ~31)) * 16;
    if (warpBase >= N) return;                                     // whole warp out of range
    const int64_t base = warpBase + (int64_t)lane * 16;
    unsigned mask = 0;
    if (base < N) {
        if (aligned && base + 16 <= N) {
            const uint4 w = *reinterpret_cast<const uint4 *>(flags + base);
            const unsigned v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned nz = __vcmpne4(v[k], 0u);           // 0xff per non-zero byte
                mask |= (((nz >> 0) & 1u) | ((nz >> 7) & 2u) | ((nz >> 14) & 4u) | ((nz >> 21) & 8u)) << (4 * k);
            }
        } else {
            for (int k = 0; k < 16 && base + k < N; k++) mask |= (flags[base + k] ? 1u : 0u) << k;
        }
    }
    const int cnt = __popc(mask);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    int o = incl - cnt;
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        setIdx[warp][o++] = (unsigned short)(lane * 16 + k);
    }
    __syncwarp();
    for (int i = lane; i < total; i += 32) fn(warpBase + setIdx[warp][i]);
    __syncwarp();
}
template <class Fn>
__device__ __forceinline__ void forEachSetWarp(const u8 *flags, int64_t N, bool aligned, Fn fn) {
    forEachSetWarpAt(blockIdx.x, flags, N, aligned, fn);
}

// unrefineCell = union of pointCells over flagged points (hexRef3D.C:1763-1776); cells pre-zeroed
__global__ void __launch_bounds__(kThreads)
k_points_to_cells(int64_t p, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                  const u8 *__restrict__ pflag, bool aligned, u8 *__restrict__ cflag, const u8 *__restrict__ need) {
    forEachSet16(pflag, p, aligned, [&](int64_t pt) {
        if (need && !need[pt]) return;          // all cells of the point are set already
        const int32_t base = sliceOff[pt >> 5];
        const int w = sliceOff[(pt >> 5) + 1] - base;
        const int32_t *row = col + ((size_t)base << 5) + (pt & 31);
        for (int j0 = 0; j0 < w; j0 += 8) {
            int32_t q[8];
            rowLabels8(row, w, j0, q);
#pragma unroll
            for (int k = 0; k < 8; k++) if (q[k] >= 0) cflag[q[k]] = 1;
        }
    });
}

// cellLevel - unrefineCell at the owners of the boundary faces (payload of the swap at hexRef3D.C:1843-1853)
__global__ void k_pack_level_minus(int64_t b, const label *__restrict__ bOwner, const u8 *__restrict__ lvl8,
                                   const u8 *__restrict__ uc, u8 *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b) { label o = bOwner[i]; out[i] = (u8)(lvl8[o] - uc[o]); }
}

// Face sweep of hexRef3D::consistentUnrefinement (hexRef3D.C:1785-1889, maxSet=false): a cell marked
// for unrefinement is unmarked when a face neighbour would end up more than one level finer.
// In-place on uc (monotone: flags only drop).  Only marked cells can change, so only they are walked.
__global__ void __launch_bounds__(kThreads)
k_unrefine_sweep(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                 const u8 *__restrict__ lvl8, u8 *uc, bool aligned, int *changed) {
    forEachSet16(uc, n, aligned, [&](int64_t c) {
        const int32_t base = sliceOff[c >> 5];
        const int w = sliceOff[(c >> 5) + 1] - base;
        const int32_t *row = col + ((size_t)base << 5) + (c & 31);
        const int mine = (int)lvl8[c] - 1;
        int mx = 0;
        for (int j0 = 0; j0 < w; j0 += 8) {                  // 8 neighbours per round, gathers batched (rows of the edge graph are long)
            int32_t q[8];
            int lq[8], uq[8];
#pragma unroll
            for (int k = 0; k < 8; k++) q[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : (int32_t)c;
#pragma unroll
            for (int k = 0; k < 8; k++) { lq[k] = lvl8[q[k]]; uq[k] = uc[q[k]]; }     // coupled faces: k_fix_unrefine_sweep
#pragma unroll
            for (int k = 0; k < 8; k++) mx = max(mx, lq[k] - uq[k]);
        }
        if (mine < mx - 1) { uc[c] = 0; *changed = 1; }
    });
}

// knock out every flagged point that has a cell which cannot be unrefined (hexRef3D.C:1908-1926)
__global__ void __launch_bounds__(kThreads)
k_knock_points(int64_t p, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
               u8 *pflag, bool aligned, const u8 *__restrict__ uc) {
    forEachSet16(pflag, p, aligned, [&](int64_t pt) {
        const int32_t base = sliceOff[pt >> 5];
        const int w = sliceOff[(pt >> 5) + 1] - base;
        const int32_t *row = col + ((size_t)base << 5) + (pt & 31);
        bool ok = true;
        for (int j0 = 0; j0 < w && ok; j0 += 8) {
            int32_t q[8];
            rowLabels8(row, w, j0, q);
            unsigned u[8];
#pragma unroll
            for (int k = 0; k < 8; k++) u[k] = q[k] >= 0 ? (unsigned)uc[q[k]] : 1u;
#pragma unroll
            for (int k = 0; k < 8; k++) if (!u[k]) ok = false;
        }
        if (!ok) pflag[pt] = 0;
    });
}

// ------------------------------------------------------------------------------------------
// Coupled (processor-patch) faces: boundary fix-ups
// ------------------------------------------------------------------------------------------
// The main kernels walk internal neighbours only; the few coupled faces are applied afterwards by
// one thread per coupled face, exactly like the reference's second loop over boundary faces
// (hexRef.C:758-780, delta.C:110-131, fvMeshRefiner.C:853-862).  cplFace = boundary-face indices
// (face - f) of the coupled faces; ghost arrays are indexed by boundary face.

// hexRef::faceConsistentRefinement boundary loop, hexRef.C:758-780 (lf = level + flag)
template <bool MAXSET>
__global__ void k_fix_refine_sweep(int64_t nC, const label *__restrict__ cplFace, const label *__restrict__ bOwner,
                                   const u8 *__restrict__ lvl8, const u8 *__restrict__ ghostLf, u8 *lf, u8 *set, int *changed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nC) return;
    const label bf = cplFace[i];
    const label o = bOwner[bf];
    const int ownLevel = (int)lvl8[o] + (int)set[o];
    const int nei = ghostLf[bf];
    if (MAXSET) { if (!set[o] && nei > ownLevel + 1) { set[o] = 1; lf[o] = (u8)(lvl8[o] + 1); *changed = 1; } }
    else { if (set[o] && ownLevel > nei + 1) { set[o] = 0; lf[o] = lvl8[o]; *changed = 1; } }
}
// hexRef3D::consistentUnrefinement boundary loop, hexRef3D.C:1855-1889
__global__ void k_fix_unrefine_sweep(int64_t nC, const label *__restrict__ cplFace, const label *__restrict__ bOwner,
                                     const u8 *__restrict__ lvl8, const u8 *__restrict__ ghostLfm, u8 *uc, int *changed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nC) return;
    const label bf = cplFace[i];
    const label o = bOwner[bf];
    if (uc[o] && ((int)lvl8[o] - 1) < (int)ghostLfm[bf] - 1) { uc[o] = 0; *changed = 1; }
}
// calculateProtectedCells boundary part, fvMeshHexRefiner.C:113-128,157-170
__global__ void k_fix_protected(int64_t nC, const label *__restrict__ cplFace, const label *__restrict__ bOwner,
                                const u8 *__restrict__ lvl8, const u8 *__restrict__ ghostU, const u8 *__restrict__ ghostLvl,
                                u8 *u, int *changed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nC) return;
    const label bf = cplFace[i];
    const label o = bOwner[bf];
    if (!u[o] && ghostU[bf] && (int)lvl8[o] > (int)ghostLvl[bf]) { u[o] = 1; *changed = 1; }
}
// checkRefinementLevels coupled part, hexRef.C:2988-3025
__global__ void k_fix_check_levels(int64_t nC, const label *__restrict__ cplFace, const label *__restrict__ bOwner,
                                   const label *__restrict__ lvl, const label *__restrict__ ghostLvl, unsigned long long *bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nC) return;
    const label bf = cplFace[i];
    if (abs(lvl[bOwner[bf]] - ghostLvl[bf]) > 1) atomicAdd(bad, 1ull);
}
// delta / scaledDelta / coded delta coupled-patch loop (delta.C:110-131, scaledDelta.C:112-136): the cells that own a
// coupled face hold the RAW internal maximum (the main kernel does not normalise them); max is applied with an
// integer atomicMax on the bit pattern (values are >= 0, so the order of doubles and of their bits agree: exact
// and order-independent).  k_fix_normalize then normalises each such cell once.
// fused with the receive side of the swap (the ghost values are read in place from the peer-written parity buffer)
template <int KIND>
__global__ void __launch_bounds__(kThreads)
k_fix_estimate_link(const __grid_constant__ GhostLink G, int64_t nC, const label *__restrict__ cplFace, const label *__restrict__ bOwner,
                    const double *__restrict__ x, double p0, double p1, double *err) {
    const double *ghost = reinterpret_cast<const double *>(ghostAcquire(G));
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ghost && i < nC) {
        const label bf = cplFace[i];
        const label o = bOwner[bf];
        const bool useOffset = fabs(p1) > kSMALL;
        const double fp = fieldXform<KIND>(x[o], p1, useOffset), fn = fieldXform<KIND>(ghost[bf], p1, useOffset);
        double eT = fabs(fp - fn);                                           // delta.C:126 (no clamp)
        if (KIND == 1) eT = eT / fmaxFoam(fminFoam(fp, fn), p0);             // scaledDelta.C:129-131
        if (eT > 0.0) atomicMax(reinterpret_cast<unsigned long long *>(err + o), (unsigned long long)__double_as_longlong(eT));
    }
    ghostRelease(G);
}
__global__ void k_fix_normalize(int64_t nCells, const label *__restrict__ cplCell, Thresholds thr, double *err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nCells) return;
    const label c = cplCell[i];
    err[c] = normalizeValue(err[c], thr);
}

// ------------------------------------------------------------------------------------------
// STAGE 4  field mapping and level update
// ------------------------------------------------------------------------------------------

// refineCell through the map, fvMeshHexRefiner.C:1562-1586.  4 cells per thread (strided by the CTA),
// the three dependent loads (cellMap -> reverseCellMap -> old flag) batched across the 4 chains.
__global__ void __launch_bounds__(kThreads)
k_remap_refine_cell(int64_t nNew, const label *__restrict__ cellMap, const label *__restrict__ reverseCellMap,
                    const u8 *__restrict__ oldFlag, u8 *__restrict__ newFlag) {
    const int64_t c0 = (int64_t)blockIdx.x * (kThreads * 4) + threadIdx.x;
    label o[4], r[4];
    u8 v[4];
#pragma unroll
    for (int t = 0; t < 4; t++) { const int64_t c = c0 + (int64_t)t * kThreads; o[t] = (c < nNew) ? cellMap[c] : -1; }
#pragma unroll
    for (int t = 0; t < 4; t++) r[t] = (o[t] >= 0) ? reverseCellMap[o[t]] : -1;
#pragma unroll
    for (int t = 0; t < 4; t++) v[t] = (o[t] >= 0) ? oldFlag[o[t]] : (u8)1;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int64_t c = c0 + (int64_t)t * kThreads;
        if (c >= nNew) continue;
        u8 out;
        if (o[t] < 0) out = 1;
        else if ((int64_t)r[t] != c) out = 1;
        else out = v[t] ? 1 : 0;
        newFlag[c] = out;
    }
}

// fvMesh::mapFields direct addressing: new[i][k] = old[cellMap[i]][k].  One thread handles 4 scalars
// (strided by the CTA so every access instruction stays coalesced); the 4 map loads are issued before
// the 4 dependent gathers so each thread keeps 4 independent chains in flight.
constexpr int kMapItems = 4;
template <int NC>
__global__ void __launch_bounds__(kThreads)
k_map_field(int64_t nNew, const label *__restrict__ cellMap, const double *__restrict__ oldF,
            double *__restrict__ newF) {
    const int64_t total = nNew * NC;
    const int64_t e0 = (int64_t)blockIdx.x * (kThreads * kMapItems) + threadIdx.x;
    int64_t src[kMapItems];
#pragma unroll
    for (int t = 0; t < kMapItems; t++) {
        const int64_t e = e0 + (int64_t)t * kThreads;
        if (e < total) {
            const int64_t i = e / NC;
            src[t] = (int64_t)cellMap[i] * NC + (e - i * NC);
        } else src[t] = -1;
    }
    double v[kMapItems];
#pragma unroll
    for (int t = 0; t < kMapItems; t++) if (src[t] >= 0) v[t] = oldF[src[t]];
#pragma unroll
    for (int t = 0; t < kMapItems; t++) if (src[t] >= 0) newF[e0 + (int64_t)t * kThreads] = v[t];
}
// multi-component fields: a thread owns kMapItems PAIRS of consecutive elements.  One 128-bit load when the two
// sources are consecutive and 16-byte aligned (identity runs of the map, components of one cell), always one 128-bit
// store.  Needs 16-byte aligned oldF / newF (checked at launch).  13 % faster than the element form for vectors.
template <int NC>
__global__ void __launch_bounds__(kThreads)
k_map_field_pair(int64_t nNew, const label *__restrict__ cellMap, const double *__restrict__ oldF, double *__restrict__ newF) {
    const int64_t total = nNew * NC;
    const int64_t p0 = (int64_t)blockIdx.x * (kThreads * kMapItems) + threadIdx.x;
    int64_t s0[kMapItems], s1[kMapItems];
#pragma unroll
    for (int t = 0; t < kMapItems; t++) {
        const int64_t e = 2 * (p0 + (int64_t)t * kThreads);
        s0[t] = s1[t] = -1;
        if (e < total) {
            const int64_t i = e / NC;
            const int k = (int)(e - i * NC);
            s0[t] = (int64_t)cellMap[i] * NC + k;
            if (e + 1 < total) s1[t] = (k + 1 < NC) ? s0[t] + 1 : (int64_t)cellMap[i + 1] * NC;
        }
    }
    double2 v[kMapItems];
#pragma unroll
    for (int t = 0; t < kMapItems; t++) {
        if (s0[t] < 0) continue;
        if (s1[t] == s0[t] + 1 && !(s0[t] & 1)) v[t] = *reinterpret_cast<const double2 *>(oldF + s0[t]);
        else {
            v[t].x = oldF[s0[t]];
            if (s1[t] >= 0) v[t].y = oldF[s1[t]];
        }
    }
#pragma unroll
    for (int t = 0; t < kMapItems; t++) {
        if (s0[t] < 0) continue;
        const int64_t e = 2 * (p0 + (int64_t)t * kThreads);
        if (s1[t] >= 0) *reinterpret_cast<double2 *>(newF + e) = v[t];
        else newF[e] = v[t].x;
    }
}
__global__ void __launch_bounds__(kThreads)
k_map_field_n(int64_t nNew, int nc, const label *__restrict__ cellMap, const double *__restrict__ oldF,
              double *__restrict__ newF) {
    const int64_t total = nNew * nc;
    const int64_t e0 = (int64_t)blockIdx.x * (kThreads * kMapItems) + threadIdx.x;
    int64_t src[kMapItems];
#pragma unroll
    for (int t = 0; t < kMapItems; t++) {
        const int64_t e = e0 + (int64_t)t * kThreads;
        if (e < total) {
            const int64_t i = e / nc;
            src[t] = (int64_t)cellMap[i] * nc + (e - i * nc);
        } else src[t] = -1;
    }
    double v[kMapItems];
#pragma unroll
    for (int t = 0; t < kMapItems; t++) if (src[t] >= 0) v[t] = oldF[src[t]];
#pragma unroll
    for (int t = 0; t < kMapItems; t++) if (src[t] >= 0) newF[e0 + (int64_t)t * kThreads] = v[t];
}

// range variant for the host<->host pipeline of amr_map_field: maps new cells [i0, i1)
__global__ void __launch_bounds__(kThreads)
k_map_field_range(int64_t i0, int64_t i1, int nc, const label *__restrict__ cellMap, const double *__restrict__ oldF,
                  double *__restrict__ newF) {
    const int64_t total = (i1 - i0) * nc;
    const int64_t t0 = (int64_t)blockIdx.x * (kThreads * kMapItems) + threadIdx.x;
    int64_t src[kMapItems], dst[kMapItems];
#pragma unroll
    for (int t = 0; t < kMapItems; t++) {
        const int64_t e = t0 + (int64_t)t * kThreads;
        if (e < total) {
            const int64_t i = i0 + e / nc;
            const int64_t k = e % nc;
            src[t] = (int64_t)cellMap[i] * nc + k;
            dst[t] = i * nc + k;
        } else src[t] = -1;
    }
    double v[kMapItems];
#pragma unroll
    for (int t = 0; t < kMapItems; t++) if (src[t] >= 0) v[t] = oldF[src[t]];
#pragma unroll
    for (int t = 0; t < kMapItems; t++) if (src[t] >= 0) newF[dst[t]] = v[t];
}
// largest source row of every chunk of new cells (decides how much of the old field must have arrived)
__global__ void __launch_bounds__(kThreads)
k_chunk_max_src(int64_t nNew, int64_t chunk, const label *__restrict__ cellMap, int *__restrict__ maxSrc) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int v = (i < nNew) ? cellMap[i] : -1;
    // all threads of a CTA belong to one chunk when chunk % kThreads == 0
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    if ((threadIdx.x & 31) == 0 && i < nNew) atomicMax(&maxSrc[i / chunk], v);
}

// conservative merge: new master value = sum V_old*old / sum V_old over the 8 merged cells,
// accumulated in ascending old label.  merged cells of master m are found through mergeOff/mergeIdx
// (a CSR built from reverseCellMap, rows ascending), so the sum order is deterministic.
__global__ void __launch_bounds__(kThreads)
k_map_conservative(int64_t nMasters, const label *__restrict__ masters, const int32_t *__restrict__ mergeOff,
                   const label *__restrict__ mergeIdx, const double *__restrict__ volOld, int nc,
                   const double *__restrict__ oldF, double *__restrict__ newF) {
    int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= nMasters * nc) return;
    int64_t mi = t / nc;
    int k = (int)(t - mi * nc);
    double vs = 0.0, acc = 0.0;
    for (int32_t j = mergeOff[mi]; j < mergeOff[mi + 1]; j++) {
        label c = mergeIdx[j];
        vs += volOld[c];
        acc += volOld[c] * oldF[(int64_t)c * nc + k];
    }
    newF[(int64_t)masters[mi] * nc + k] = acc / vs;
}

// cellLevel through the map
__global__ void __launch_bounds__(kThreads)
k_update_levels(int64_t nNew, const label *__restrict__ cellMap, const label *__restrict__ oldLevel,
                const u8 *__restrict__ refinedOld, const u8 *__restrict__ unrefinedOld, label *__restrict__ newLevel) {
    int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= nNew) return;
    label o = cellMap[c];
    label l = oldLevel[o];
    if (refinedOld && refinedOld[o]) l++;
    if (unrefinedOld && unrefinedOld[o]) l--;
    newLevel[c] = l;
}

// flags from an ascending label list (hexRef.C:1415-1419)
__global__ void k_list_to_flags(int64_t cnt, const label *__restrict__ list, u8 *__restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) flag[list[i]] = 1;
}
