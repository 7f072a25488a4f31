// lohner.cuh -- Lohner::update (Lohner.C:71-159) on the device.
//
// xf = cubic<scalar>(mesh).interpolate(x) is OpenFOAM core; the arithmetic below restates
// linear.H / cubic.H / gaussGrad.C of OpenFOAM-v2212 operation by operation (see the oracle's
// or_estimate_lohner for the formulas and the "parity unpinned" note).  The one order-sensitive
// piece is the Gauss-gradient SUM: the reference accumulates face contributions in ascending
// face index (internal faces, then boundary patches).  Each cell therefore walks its own faces in
// ascending face label (cell->face SELL, entry = 2*face + role, role 1 = cell is the face's
// neighbour) and adds/subtracts in exactly that order: deterministic, no atomics, bitwise equal to
// the host loop.  Two passes: gradient per cell, then the face indicator (xf needs both gradients).
#pragma once
#include "common.cuh"
#include "kernels.cuh"

struct Geom {
    const double *weights, *Sf, *magSf, *deltaCoeffs, *V;
};

__global__ void k_cf_deg(const label *__restrict__ owner, const label *__restrict__ neighbour, int64_t f, int64_t b,
                         int32_t *__restrict__ deg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f + b) return;
    atomicAdd(&deg[owner[i]], 1);
    if (i < f) atomicAdd(&deg[neighbour[i]], 1);
}
__global__ void k_cf_fill(const label *__restrict__ owner, const label *__restrict__ neighbour, int64_t f, int64_t b,
                          const int32_t *__restrict__ sliceOff, int32_t *__restrict__ cursor, int32_t *__restrict__ col) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f + b) return;
    label o = owner[i];
    int j = atomicAdd(&cursor[o], 1);
    col[((size_t)(sliceOff[o >> 5] + j) << 5) + (o & 31)] = (int32_t)(2 * i);
    if (i < f) {
        label q = neighbour[i];
        j = atomicAdd(&cursor[q], 1);
        col[((size_t)(sliceOff[q >> 5] + j) << 5) + (q & 31)] = (int32_t)(2 * i + 1);
    }
}

// fv::gaussGrad::gradf with linear interpolation (gaussGrad.C), per cell in ascending face order
__global__ void __launch_bounds__(kThreads)
k_gauss_grad(int64_t n, int64_t f, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
             const label *__restrict__ owner, const label *__restrict__ neighbour, Geom G,
             const double *__restrict__ x, const double *__restrict__ xbf, const double *__restrict__ xn,
             const u8 *__restrict__ bCoupled, double *__restrict__ grad) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const int32_t base = sliceOff[c >> 5];
    const int w = sliceOff[(c >> 5) + 1] - base;
    const int32_t *row = col + ((size_t)base << 5) + (c & 31);
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    for (int j = 0; j < w; j++) {
        const int32_t e = row[(size_t)j << 5];
        if (e < 0) break;
        const int64_t face = e >> 1;
        const double s0 = G.Sf[3 * face], s1 = G.Sf[3 * face + 1], s2 = G.Sf[3 * face + 2];
        if (face < f) {
            const label P = owner[face], N = neighbour[face];
            const double ssf = G.weights[face] * (x[P] - x[N]) + x[N];
            if (e & 1) { g0 -= s0 * ssf; g1 -= s1 * ssf; g2 -= s2 * ssf; }
            else { g0 += s0 * ssf; g1 += s1 * ssf; g2 += s2 * ssf; }
        } else {
            const int64_t i = face - f;
            double pssf;
            if (bCoupled[i]) pssf = G.weights[face] * x[c] + (1.0 - G.weights[face]) * xn[i];
            else pssf = xbf[i];
            g0 += s0 * pssf; g1 += s1 * pssf; g2 += s2 * pssf;
        }
    }
    const double V = G.V[c];
    grad[3 * c] = g0 / V; grad[3 * c + 1] = g1 / V; grad[3 * c + 2] = g2 / V;
}

// boundaryField of x: given values on physical patches (or zeroGradient), received neighbour values on
// processor patches
__global__ void k_boundary_field(int64_t b, const label *__restrict__ bOwner, const double *__restrict__ x,
                                 const double *__restrict__ xbIn, const double *__restrict__ xn,
                                 const u8 *__restrict__ bCoupled, double *__restrict__ xbf) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    if (bCoupled[i]) xbf[i] = xn[i];
    else xbf[i] = xbIn ? xbIn[i] : x[bOwner[i]];
}

template <bool SCALE>
__global__ void __launch_bounds__(kThreads)
k_lohner(int64_t n, int64_t f, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
         const label *__restrict__ owner, const label *__restrict__ neighbour, Geom G, const double *__restrict__ x,
         const double *__restrict__ grad, const double *__restrict__ xbf, const double *__restrict__ xn,
         const u8 *__restrict__ bCoupled, double epsilon, Thresholds thr, double *__restrict__ err) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const int32_t base = sliceOff[c >> 5];
    const int w = sliceOff[(c >> 5) + 1] - base;
    const int32_t *row = col + ((size_t)base << 5) + (c & 31);
    double e = 0.0;
    for (int j = 0; j < w; j++) {
        const int32_t en = row[(size_t)j << 5];
        if (en < 0) break;
        const int64_t face = en >> 1;
        double eT;
        if (face < f) {
            const label own = owner[face], nei = neighbour[face];
            const double l = G.weights[face];
            const double xo = x[own], xnb = x[nei];
            const double kSc = l * (1.0 - l * (3.0 - 2.0 * l));
            const double kVecP = ((1.0 - l) * (1.0 - l)) * l;
            const double kVecN = (l * l) * (l - 1.0);
            double corr = kSc * xo + (-kSc) * xnb;
            const double gi0 = kVecP * grad[3 * own] + kVecN * grad[3 * nei];
            const double gi1 = kVecP * grad[3 * own + 1] + kVecN * grad[3 * nei + 1];
            const double gi2 = kVecP * grad[3 * own + 2] + kVecN * grad[3 * nei + 2];
            const double dot = gi0 * G.Sf[3 * face] + gi1 * G.Sf[3 * face + 1] + gi2 * G.Sf[3 * face + 2];
            corr = corr + dot / G.magSf[face] / G.deltaCoeffs[face];
            const double xf = (l * (xo - xnb) + xnb) + corr;
            eT = sqrt(fabs(xnb - 2.0 * xf + xo)
                      / (fabs(xnb - xf) + fabs(xf - xo) + epsilon * (fabs(xnb) + 2.0 * fabs(xf) + fabs(xo))));
        } else {
            const int64_t i = face - f;
            if (!bCoupled[i]) continue;
            const double xp = x[c], xv = xn[i], xb = xbf[i];
            eT = sqrt(fabs(xv - 2.0 * xb + xp)
                      / (fabs(xv - xb) + fabs(xb - xp) + epsilon * (fabs(xv) + 2.0 * fabs(xb) + fabs(xp))));
        }
        e = fmaxFoam(e, eT);
    }
    err[c] = SCALE ? normalizeValue(e, thr) : e;
}

// ---- batched forms ---------------------------------------------------------------------------------------------------
// The walks above chase four dependent gathers per face (row entry -> owner/neighbour + geometry -> x / grad) one face
// at a time: latency-bound.  Here a cell takes its faces four at a time: all row entries, then every face-indexed
// value of the four faces, then the cell-indexed values, and only then the arithmetic -- in the same ascending face
// order, operation by operation, as the one-face walk (the sums and maxima are bitwise identical; tests compare both
// with the oracle).  Four faces keep the kernel under 128 registers.
constexpr int kFaceBatch = 4;

__global__ void __launch_bounds__(kThreads)
k_gauss_grad_batched(int64_t n, int64_t f, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                     const label *__restrict__ owner, const label *__restrict__ neighbour, Geom G,
                     const double *__restrict__ x, const double *__restrict__ xbf, const double *__restrict__ xn,
                     const u8 *__restrict__ bCoupled, double *__restrict__ grad) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const int32_t base = sliceOff[c >> 5];
    const int w = sliceOff[(c >> 5) + 1] - base;
    const int32_t *row = col + ((size_t)base << 5) + (c & 31);
    const double xc = x[c];
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    for (int j0 = 0; j0 < w; j0 += kFaceBatch) {
        int32_t e[kFaceBatch];
#pragma unroll
        for (int k = 0; k < kFaceBatch; k++) e[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : -1;
        label P[kFaceBatch], N[kFaceBatch];
        double wt[kFaceBatch], s0[kFaceBatch], s1[kFaceBatch], s2[kFaceBatch], bv[kFaceBatch];
        bool cpl[kFaceBatch];
#pragma unroll
        for (int k = 0; k < kFaceBatch; k++) {
            P[k] = N[k] = (label)c; wt[k] = 0.0; s0[k] = s1[k] = s2[k] = 0.0; bv[k] = 0.0; cpl[k] = false;
            if (e[k] < 0) continue;
            const int64_t face = e[k] >> 1;
            s0[k] = G.Sf[3 * face]; s1[k] = G.Sf[3 * face + 1]; s2[k] = G.Sf[3 * face + 2];
            if (face < f) { P[k] = owner[face]; N[k] = neighbour[face]; wt[k] = G.weights[face]; }
            else {
                const int64_t i = face - f;
                cpl[k] = bCoupled[i] != 0;
                if (cpl[k]) { wt[k] = G.weights[face]; bv[k] = xn[i]; } else bv[k] = xbf[i];
            }
        }
        double xP[kFaceBatch], xN[kFaceBatch];
#pragma unroll
        for (int k = 0; k < kFaceBatch; k++) { xP[k] = x[P[k]]; xN[k] = x[N[k]]; }
#pragma unroll
        for (int k = 0; k < kFaceBatch; k++) {
            if (e[k] < 0) continue;
            const int64_t face = e[k] >> 1;
            if (face < f) {
                const double ssf = wt[k] * (xP[k] - xN[k]) + xN[k];
                if (e[k] & 1) { g0 -= s0[k] * ssf; g1 -= s1[k] * ssf; g2 -= s2[k] * ssf; }
                else { g0 += s0[k] * ssf; g1 += s1[k] * ssf; g2 += s2[k] * ssf; }
            } else {
                const double pssf = cpl[k] ? wt[k] * xc + (1.0 - wt[k]) * bv[k] : bv[k];
                g0 += s0[k] * pssf; g1 += s1[k] * pssf; g2 += s2[k] * pssf;
            }
        }
    }
    const double V = G.V[c];
    grad[3 * c] = g0 / V; grad[3 * c + 1] = g1 / V; grad[3 * c + 2] = g2 / V;
}

template <bool SCALE>
__global__ void __launch_bounds__(kThreads)
k_lohner_batched(int64_t n, int64_t f, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                 const label *__restrict__ owner, const label *__restrict__ neighbour, Geom G, const double *__restrict__ x,
                 const double *__restrict__ grad, const double *__restrict__ xbf, const double *__restrict__ xn,
                 const u8 *__restrict__ bCoupled, double epsilon, Thresholds thr, double *__restrict__ err) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const int32_t base = sliceOff[c >> 5];
    const int w = sliceOff[(c >> 5) + 1] - base;
    const int32_t *row = col + ((size_t)base << 5) + (c & 31);
    const double xc = x[c];
    double e = 0.0;
    for (int j0 = 0; j0 < w; j0 += kFaceBatch) {
        int32_t en[kFaceBatch];
#pragma unroll
        for (int k = 0; k < kFaceBatch; k++) en[k] = (j0 + k < w) ? row[(size_t)(j0 + k) << 5] : -1;
        label P[kFaceBatch], N[kFaceBatch];
        double l[kFaceBatch], s0[kFaceBatch], s1[kFaceBatch], s2[kFaceBatch], ms[kFaceBatch], dc[kFaceBatch], bv[kFaceBatch], bn[kFaceBatch];
        int kind[kFaceBatch];                           // 0 skip, 1 internal, 2 coupled boundary
#pragma unroll
        for (int k = 0; k < kFaceBatch; k++) {
            P[k] = N[k] = (label)c; kind[k] = 0;
            l[k] = s0[k] = s1[k] = s2[k] = 0.0; ms[k] = dc[k] = 1.0; bv[k] = bn[k] = 0.0;
            if (en[k] < 0) continue;
            const int64_t face = en[k] >> 1;
            if (face < f) {
                kind[k] = 1;
                P[k] = owner[face]; N[k] = neighbour[face]; l[k] = G.weights[face];
                s0[k] = G.Sf[3 * face]; s1[k] = G.Sf[3 * face + 1]; s2[k] = G.Sf[3 * face + 2];
                ms[k] = G.magSf[face]; dc[k] = G.deltaCoeffs[face];
            } else {
                const int64_t i = face - f;
                if (bCoupled[i]) { kind[k] = 2; bn[k] = xn[i]; bv[k] = xbf[i]; }
            }
        }
        double xo[kFaceBatch], xnb[kFaceBatch], gP[kFaceBatch][3], gN[kFaceBatch][3];
#pragma unroll
        for (int k = 0; k < kFaceBatch; k++) {
            xo[k] = x[P[k]]; xnb[k] = x[N[k]];
#pragma unroll
            for (int d = 0; d < 3; d++) { gP[k][d] = grad[3 * (size_t)P[k] + d]; gN[k][d] = grad[3 * (size_t)N[k] + d]; }
        }
#pragma unroll
        for (int k = 0; k < kFaceBatch; k++) {
            if (kind[k] == 0) continue;
            double eT;
            if (kind[k] == 1) {
                const double lk = l[k];
                const double kSc = lk * (1.0 - lk * (3.0 - 2.0 * lk));
                const double kVecP = ((1.0 - lk) * (1.0 - lk)) * lk;
                const double kVecN = (lk * lk) * (lk - 1.0);
                double corr = kSc * xo[k] + (-kSc) * xnb[k];
                const double gi0 = kVecP * gP[k][0] + kVecN * gN[k][0];
                const double gi1 = kVecP * gP[k][1] + kVecN * gN[k][1];
                const double gi2 = kVecP * gP[k][2] + kVecN * gN[k][2];
                const double dot = gi0 * s0[k] + gi1 * s1[k] + gi2 * s2[k];
                corr = corr + dot / ms[k] / dc[k];
                const double xf = (lk * (xo[k] - xnb[k]) + xnb[k]) + corr;
                eT = sqrt(fabs(xnb[k] - 2.0 * xf + xo[k])
                          / (fabs(xnb[k] - xf) + fabs(xf - xo[k]) + epsilon * (fabs(xnb[k]) + 2.0 * fabs(xf) + fabs(xo[k]))));
            } else {
                const double xp = xc, xv = bn[k], xb = bv[k];
                eT = sqrt(fabs(xv - 2.0 * xb + xp)
                          / (fabs(xv - xb) + fabs(xb - xp) + epsilon * (fabs(xv) + 2.0 * fabs(xb) + fabs(xp))));
            }
            e = fmaxFoam(e, eT);
        }
    }
    err[c] = SCALE ? normalizeValue(e, thr) : e;
}
