// fields.cuh -- the rows either side of the decision path:
//   N1  label / protectedCell renumbering after a topology change   hexRef.C:2455-2560, fvMeshHexRefiner.C:229-240
//   N2  flux fix-up on new and split faces                          adaptiveFvMesh.C:122-291
//   N3  gradient selectors (Gauss linear, leastSquares), mesh size (meshSizeObject), cellSize,
//       multicomponent                                              gradientRange.C:87-110, meshSizeObject.C:66-160,
//                                                                   cellSize.C:133-210, multicomponent.C:95-157
// Every floating-point SUM walks the cell's own faces in the reference's order (cell->face SELL, rows
// ascending in face label; entry = 2*face + role, role 1 = the cell is the face's neighbour), so the
// results are bitwise those of the host loops: no atomics on doubles.
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "lohner.cuh"

// ---------------------------------------------------------------------------------------------- N1
// ListOps reorder(oldToNew, size, -1, list): hexRef::updateMesh, hexRef.C:2463 / :2529 (newV pre-filled with -1)
__global__ void __launch_bounds__(kThreads)
k_reorder_labels(int64_t nOld, const label *__restrict__ oldToNew, const label *__restrict__ oldV, label *__restrict__ newV) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= nOld) return;
    const label t = oldToNew[i];
    if (t >= 0) newV[t] = oldV[i];
}
// hexRef.C:2467-2484 / :2533-2556
__global__ void __launch_bounds__(kThreads)
k_gather_labels(int64_t nNew, const label *__restrict__ map, const label *__restrict__ oldV, label *__restrict__ newV) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= nNew) return;
    const label o = map[i];
    newV[i] = (o == -1) ? -1 : oldV[o];
}
// fvMeshHexRefiner.C:229-240, :283-297
__global__ void __launch_bounds__(kThreads)
k_gather_protected(int64_t nNew, const label *__restrict__ cellMap, const u8 *__restrict__ oldP, u8 *__restrict__ newP) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= nNew) return;
    const label o = cellMap[i];
    newP[i] = (o >= 0) ? oldP[o] : (u8)0;
}

// ---------------------------------------------------------------------------------------------- N2
// which faces get a recomputed flux (adaptiveFvMesh.C:131-150, :212-227): inflated faces, faces made from a
// master face, and the master faces themselves.  bad: a surviving face whose old face was removed (:139-145)
__global__ void __launch_bounds__(kThreads)
k_flux_mark(int64_t nf, const label *__restrict__ faceMap, const label *__restrict__ reverseFaceMap, u8 *__restrict__ fix,
            int *__restrict__ bad) {
    const int64_t facei = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (facei >= nf) return;
    const label oldFacei = faceMap[facei];
    if (oldFacei >= 0) {
        const label masterFacei = reverseFaceMap[oldFacei];
        if (masterFacei < 0) { *bad = 1; return; }
        if (masterFacei != facei) { fix[masterFacei] = 1; fix[facei] = 1; }
    } else if (oldFacei == -1) fix[facei] = 1;
}
// phi[face] = (linear interpolate(U) & Sf)[face] on the marked faces; empty patches carry no flux entries
__global__ void __launch_bounds__(kThreads)
k_flux_apply(int64_t nf, int64_t f, const label *__restrict__ owner, const label *__restrict__ neighbour,
             const double *__restrict__ weights, const double *__restrict__ Sf, const double *__restrict__ U,
             const double *__restrict__ Ub, const double *__restrict__ Un, const u8 *__restrict__ bCoupled,
             const u8 *__restrict__ bEmpty, const u8 *__restrict__ fix, double *__restrict__ phi, unsigned long long *count) {
    const int64_t facei = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (facei >= nf || !fix[facei]) return;
    double u0, u1, u2;
    const double w = weights[facei];
    if (facei < f) {
        const label P = owner[facei], N = neighbour[facei];
        u0 = w * (U[3 * P] - U[3 * N]) + U[3 * N];
        u1 = w * (U[3 * P + 1] - U[3 * N + 1]) + U[3 * N + 1];
        u2 = w * (U[3 * P + 2] - U[3 * N + 2]) + U[3 * N + 2];
    } else {
        const int64_t i = facei - f;
        if (bEmpty[i]) return;
        const label P = owner[facei];
        if (bCoupled[i]) {
            u0 = w * U[3 * P] + (1.0 - w) * Un[3 * i];
            u1 = w * U[3 * P + 1] + (1.0 - w) * Un[3 * i + 1];
            u2 = w * U[3 * P + 2] + (1.0 - w) * Un[3 * i + 2];
        } else if (Ub) {
            u0 = Ub[3 * i]; u1 = Ub[3 * i + 1]; u2 = Ub[3 * i + 2];
        } else {
            u0 = U[3 * P]; u1 = U[3 * P + 1]; u2 = U[3 * P + 2];
        }
    }
    phi[facei] = u0 * Sf[3 * facei] + u1 * Sf[3 * facei + 1] + u2 * Sf[3 * facei + 2];
    atomicAdd(count, 1ull);
}
// patchInternalField of a vector field over the boundary faces (NCCL transport of the swap)
__global__ void k_pack_vec3(int64_t b, const label *__restrict__ bOwner, const double *__restrict__ U, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    const label o = bOwner[i];
    out[3 * i] = U[3 * o]; out[3 * i + 1] = U[3 * o + 1]; out[3 * i + 2] = U[3 * o + 2];
}

// ---------------------------------------------------------------------------------------------- N3
// fv::leastSquaresGrad::calcGrad (leastSquaresGrad.C) with the geometric vectors of leastSquaresVectors
// supplied by the host: per cell in ascending face order, owner side += pVectors*(x_N - x_P),
// neighbour side -= nVectors*(x_N - x_P); boundary faces += pVectors*(x_b - x_P)
__global__ void __launch_bounds__(kThreads)
k_ls_grad(int64_t n, int64_t f, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
          const label *__restrict__ owner, const label *__restrict__ neighbour, const double *__restrict__ lsP,
          const double *__restrict__ lsN, const double *__restrict__ x, const double *__restrict__ xbf,
          double *__restrict__ grad) {
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
        if (face < f) {
            const double d = x[neighbour[face]] - x[owner[face]];
            if (e & 1) { g0 -= lsN[3 * face] * d; g1 -= lsN[3 * face + 1] * d; g2 -= lsN[3 * face + 2] * d; }
            else { g0 += lsP[3 * face] * d; g1 += lsP[3 * face + 1] * d; g2 += lsP[3 * face + 2] * d; }
        } else {
            const double d = xbf[face - f] - x[c];
            g0 += lsP[3 * face] * d; g1 += lsP[3 * face + 1] * d; g2 += lsP[3 * face + 2] * d;
        }
    }
    grad[3 * c] = g0; grad[3 * c + 1] = g1; grad[3 * c + 2] = g2;
}
// error = mag(grad) (gradientRange.C:100), optionally * V/tV (the coded flame selector)
__global__ void __launch_bounds__(kThreads)
k_mag_grad(int64_t n, const double *__restrict__ grad, const double *__restrict__ V, double tV, double *__restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const double g0 = grad[3 * c], g1 = grad[3 * c + 1], g2 = grad[3 * c + 2];
    double v = sqrt(g0 * g0 + g1 * g1 + g2 * g2);
    if (tV > 0) v *= V[c] / tV;
    out[c] = v;
}

// meshSizeObject::calcDx, non-3-D branch (meshSizeObject.C:80-126), incl. the reference's face count
// (an internal face counts twice on its owner and never on its neighbour)
__global__ void __launch_bounds__(kThreads)
k_mesh_dx_planar(int64_t n, int64_t f, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col,
                 const double *__restrict__ Sf, double v0, double v1, double v2, double *__restrict__ dx) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const int32_t base = sliceOff[c >> 5];
    const int w = sliceOff[(c >> 5) + 1] - base;
    const int32_t *row = col + ((size_t)base << 5) + (c & 31);
    double sum = 0.0;
    int nFaces = 0;
    for (int j = 0; j < w; j++) {
        const int32_t e = row[(size_t)j << 5];
        if (e < 0) break;
        const int64_t face = e >> 1;
        const double s0 = Sf[3 * face], s1 = Sf[3 * face + 1], s2 = Sf[3 * face + 2];
        const double magSf = sqrt(s0 * s0 + s1 * s1 + s2 * s2);
        const double dot = (s0 / magSf) * v0 + (s1 / magSf) * v1 + (s2 / magSf) * v2;
        if (!(fabs(dot) > 0.5)) continue;
        sum += magSf;
        if (face >= f) nFaces += 1;
        else if (!(e & 1)) nFaces += 2;
    }
    dx[c] = sqrt(sum / (double)nFaces);
}
// 3-D branch: cbrt(V) (meshSizeObject.C:128-131).  CUDA's cbrt and glibc's may differ in the last bit.
__global__ void __launch_bounds__(kThreads)
k_mesh_dx_cbrt(int64_t n, const double *__restrict__ V, double *__restrict__ dx) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c < n) dx[c] = cbrt(V[c]);
}
// meshSizeObject::calcDX (meshSizeObject.C:134-160): 2V / sum of cmptMag(Sf) over cells()[c]; primitiveMesh::calcCells
// lists the faces a cell owns first, then the faces it is neighbour of, each ascending
__global__ void __launch_bounds__(kThreads)
k_mesh_dX(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col, const double *__restrict__ Sf,
          const double *__restrict__ V, double *__restrict__ dX) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const int32_t base = sliceOff[c >> 5];
    const int w = sliceOff[(c >> 5) + 1] - base;
    const int32_t *row = col + ((size_t)base << 5) + (c & 31);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int role = 0; role < 2; role++)
        for (int j = 0; j < w; j++) {
            const int32_t e = row[(size_t)j << 5];
            if (e < 0) break;
            if ((e & 1) != role) continue;
            const int64_t face = e >> 1;
            s0 += fabs(Sf[3 * face]); s1 += fabs(Sf[3 * face + 1]); s2 += fabs(Sf[3 * face + 2]);
        }
    const double tv = 1.0 * (2.0 * V[c]);
    dX[3 * c] = tv / s0; dX[3 * c + 1] = tv / s1; dX[3 * c + 2] = tv / s2;
}

// errorEstimator::maxRefinement (errorEstimator.C:262-292), minDx branch
__global__ void __launch_bounds__(kThreads)
k_max_refinement(int64_t n, const label *__restrict__ lvl, double minDx, const double *__restrict__ dx,
                 const double *__restrict__ err, label *__restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    out[c] = lvl[c] + ((dx[c] > minDx && err[c] > 0) ? 1 : 0);
}
__global__ void __launch_bounds__(kThreads) k_fill_scalar(int64_t n, double v, double *__restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c < n) out[c] = v;
}
__global__ void __launch_bounds__(kThreads) k_fill_labels(int64_t n, label v, label *__restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c < n) out[c] = v;
}

// errorEstimators::cellSize::update (cellSize.C:133-210)
struct CellSizeArgs {
    int sizeType;            // 0 VOLUME, 1 CHARACTERISTIC, 2 MAG, 3 CMPT
    int geoValid[3];         // geometricD == 1
    double lowerUnrefine, lowerRefine;
    int nCmpts, cmpts[3];
    double minDX[3], maxDX[3];
};
__global__ void __launch_bounds__(kThreads)
k_cell_size(int64_t n, CellSizeArgs A, const double *__restrict__ V, const double *__restrict__ dx, const double *__restrict__ dX,
            double *__restrict__ err) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    double e;
    if (A.sizeType == 3) {
        e = err[c];
        for (int i = 0; i < A.nCmpts; i++) {
            const int k = A.cmpts[i];
            const double d = dX[3 * c + k];
            if (d > A.maxDX[k]) e = fmaxFoam(1.0, e);
            else if (d < A.minDX[k]) e = fmaxFoam(-1.0, e);
            else e = fmaxFoam(0.0, e);
        }
        err[c] = e;
        return;
    }
    if (A.sizeType == 0) e = V[c];
    else if (A.sizeType == 1) e = dx[c];
    else {
        e = 0.0;
        for (int k = 0; k < 3; k++) if (A.geoValid[k]) e += dX[3 * c + k] * dX[3 * c + k];
        e = sqrt(e);
    }
    err[c] = (e < A.lowerUnrefine) ? -1.0 : ((e > A.lowerRefine) ? 1.0 : 0.0);
}

// errorEstimators::multicomponent::update (multicomponent.C:113-117): error = max(error, errors_[i].error())
__global__ void __launch_bounds__(kThreads)
k_max_combine(int64_t n, const double *__restrict__ a, double *__restrict__ acc, int first) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    acc[c] = fmaxFoam(first ? -1.0 : acc[c], a[c]);
}

// errorEstimators::multicomponent::maxRefinement (multicomponent.C:122-157) is an in-place loop over ascending
// cells in which a cell that "fires" (error > 0.5 and maxRefinement_ < its level) also overwrites its
// cellCells.  The outcome is therefore defined by a recurrence over lower-numbered neighbours:
//   value seen by c at its turn = level of its highest-numbered fired neighbour a < c, else the value before the pass
//   fire(c) = error[c] > 0.5  &&  seen(c) < level[c]
//   after the pass: maxRefinement[j] = level[k*], k* = highest-numbered fired cell in {j} + cellCells(j)
// state: 0 not fired, 1 fired, 2 undetermined.  A wavefront resolves every cell whose highest-numbered
// undetermined lower neighbour lies below its highest-numbered fired lower neighbour; each launch settles at
// least the lowest undetermined cell, and decided states never change, so chaotic reads are safe.
__global__ void __launch_bounds__(kThreads)
k_mc_init(int64_t n, const double *__restrict__ err, u8 *__restrict__ state) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c < n) state[c] = (err[c] > 0.5) ? 2 : 0;
}
__global__ void __launch_bounds__(kThreads)
k_mc_wave(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col, const label *__restrict__ level,
          const label *__restrict__ before, volatile u8 *state, int *__restrict__ flags /* [0] progress, [1] remaining */) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n || state[c] != 2) return;
    const int32_t base = sliceOff[c >> 5];
    const int w = sliceOff[(c >> 5) + 1] - base;
    const int32_t *row = col + ((size_t)base << 5) + (c & 31);
    int32_t maxF = -1, maxU = -1;
    for (int j = 0; j < w; j++) {
        const int32_t a = row[(size_t)j << 5];
        if (a >= (int32_t)c) continue;            // padding (== c) and later cells
        const u8 s = state[a];
        if (s == 1) maxF = max(maxF, a);
        else if (s == 2) maxU = max(maxU, a);
    }
    if (maxU > maxF) { flags[1] = 1; return; }
    const label seen = (maxF >= 0) ? level[maxF] : before[c];
    state[c] = (seen < level[c]) ? 1 : 0;
    flags[0] = 1;
}
__global__ void __launch_bounds__(kThreads)
k_mc_apply(int64_t n, const int32_t *__restrict__ sliceOff, const int32_t *__restrict__ col, const label *__restrict__ level,
           const u8 *__restrict__ state, label *__restrict__ maxRefinement) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const int32_t base = sliceOff[c >> 5];
    const int w = sliceOff[(c >> 5) + 1] - base;
    const int32_t *row = col + ((size_t)base << 5) + (c & 31);
    int32_t last = state[c] == 1 ? (int32_t)c : -1;
    for (int j = 0; j < w; j++) {
        const int32_t a = row[(size_t)j << 5];
        if (state[a] == 1) last = max(last, a);
    }
    if (last >= 0) maxRefinement[c] = level[last];
}

// ---------------------------------------------------------------------------------------------- N4
// hexRefRefinementHistory::markCommonCells (hexRefRefinementHistory.C:398-444), order-free form: root[c] = top-most
// ancestor of the cell's history entry; a cluster is numbered by the rank of its LOWEST cell among the lowest cells of
// all clusters, which is the order of first appearance over ascending cells the reference produces.
__global__ void __launch_bounds__(kThreads)
k_history_root(int64_t n, const label *__restrict__ visible, const label *__restrict__ splitParent, label *__restrict__ root,
               int *__restrict__ firstCell) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    label index = visible[c];
    if (index >= 0) {
        label p;
        while ((p = splitParent[index]) != -1) index = p;
        atomicMin(&firstCell[index], (int)c);
    }
    root[c] = index >= 0 ? index : -1;
}
__global__ void __launch_bounds__(kThreads)
k_cluster_first_flag(int64_t n, const label *__restrict__ root, const int *__restrict__ firstCell, u8 *__restrict__ isFirst) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    isFirst[c] = (root[c] >= 0 && firstCell[root[c]] == (int)c) ? 1 : 0;
}
// firstList = ascending first cells (compaction of isFirst): cluster id = position in that list
__global__ void __launch_bounds__(kThreads)
k_cluster_of_root(const int32_t *__restrict__ total, const int32_t *__restrict__ firstList, const label *__restrict__ root,
                  int *__restrict__ clusterOfRoot) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= *total) return;
    clusterOfRoot[root[firstList[i]]] = (int)i;      // reuses the firstCell array: entries of non-roots are never read
}
__global__ void __launch_bounds__(kThreads)
k_cluster_cells(int64_t n, const label *__restrict__ root, const int *__restrict__ clusterOfRoot, label *__restrict__ cellToCluster) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    cellToCluster[c] = root[c] >= 0 ? clusterOfRoot[root[c]] : -1;
}
// hexRefRefinementHistory::add (:448-487) / refinement::add (refinement.C:1007-1046)
__global__ void __launch_bounds__(kThreads)
k_blocked_faces(int64_t nf, int64_t f, const label *__restrict__ owner, const label *__restrict__ neighbour,
                const label *__restrict__ cellToCluster, u8 *__restrict__ blocked, unsigned long long *nUnblocked) {
    const int64_t face = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (face >= nf) return;
    u8 bl = 1;
    if (face < f) {
        const label a = cellToCluster[owner[face]], b2 = cellToCluster[neighbour[face]];
        if (a != -1 && a == b2) { bl = 0; atomicAdd(nUnblocked, 1ull); }
    }
    blocked[face] = bl;
}
// hexRefRefinementHistory::apply (:490-545): the cluster's processor is the one given to the owner of its lowest
// in-cluster face; every cell on an in-cluster face follows
__global__ void __launch_bounds__(kThreads)
k_cluster_first_face(int64_t f, const label *__restrict__ owner, const label *__restrict__ neighbour,
                     const label *__restrict__ cellToCluster, int *__restrict__ firstFace) {
    const int64_t face = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (face >= f) return;
    const label a = cellToCluster[owner[face]];
    if (a != -1 && a == cellToCluster[neighbour[face]]) atomicMin(&firstFace[a], (int)face);
}
__global__ void __launch_bounds__(kThreads)
k_cluster_proc(int64_t nClusters, const int *__restrict__ firstFace, const label *__restrict__ owner,
               const label *__restrict__ decompIn, int *__restrict__ clusterToProc) {
    const int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (k >= nClusters) return;
    const int ff = firstFace[k];
    clusterToProc[k] = (ff == 0x7fffffff) ? -1 : decompIn[owner[ff]];
}
__global__ void __launch_bounds__(kThreads)
k_cluster_apply(int64_t f, const label *__restrict__ owner, const label *__restrict__ neighbour,
                const label *__restrict__ cellToCluster, const int *__restrict__ clusterToProc, const label *__restrict__ decompIn,
                label *__restrict__ decompOut, u8 *__restrict__ changed) {
    const int64_t face = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (face >= f) return;
    const label own = owner[face], nei = neighbour[face];
    const label a = cellToCluster[own];
    if (a == -1 || a != cellToCluster[nei]) return;
    const int proc = clusterToProc[a];
    if (decompIn[own] != proc) { decompOut[own] = proc; changed[own] = 1; }
    if (decompIn[nei] != proc) { decompOut[nei] = proc; changed[nei] = 1; }
}
// procLoadNew (fvMeshBalance.C:497-503): per-CTA shared histogram, then global atomics
__global__ void __launch_bounds__(kThreads)
k_proc_load(int64_t n, const label *__restrict__ distribution, int nProcs, unsigned long long *__restrict__ load, int *__restrict__ bad) {
    extern __shared__ unsigned int hist[];
    for (int q = threadIdx.x; q < nProcs; q += kThreads) hist[q] = 0;
    __syncthreads();
    for (int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x; c < n; c += (int64_t)gridDim.x * kThreads) {
        const label q = distribution[c];
        if (q < 0 || q >= nProcs) *bad = 1;
        else atomicAdd(&hist[q], 1u);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nProcs; q += kThreads) if (hist[q]) atomicAdd(&load[q], (unsigned long long)hist[q]);
}
__global__ void k_fill_int(int64_t n, int v, int *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

// ---- old-time volumes through a topology change (fvMesh::updateMesh, OpenFOAM core; SURVEY App. A "V0/V00 additionally
// receive merged-cell volume injection").  Direct map, then every merged-away old cell adds its volume to its master.
// The reference adds in ascending old label.  Here the merged cells of each master are collected into a small slot
// block (atomic cursor, any order), sorted ascending by the master's thread and added in that order: the sum order --
// and the result, bit for bit -- is the reference's.
__global__ void __launch_bounds__(kThreads)
k_map_old_volume(int64_t nNew, const label *__restrict__ cellMap, const double *__restrict__ oldV, double *__restrict__ newV) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= nNew) return;
    const label o = cellMap[i];
    newV[i] = o > -1 ? oldV[o] : 0.0;
}
// merged-away old cells (reverseCellMap < -1) flag their masters
__global__ void __launch_bounds__(kThreads)
k_merge_masters(int64_t nOld, const label *__restrict__ reverseCellMap, u8 *__restrict__ isMaster, unsigned long long *nMerged) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int hit = 0;
    if (c < nOld) {
        const label r = reverseCellMap[c];
        if (r < -1) { isMaster[-r - 2] = 1; hit = 1; }
    }
    const int cnt = __syncthreads_count(hit);
    if (threadIdx.x == 0 && cnt) atomicAdd(nMerged, (unsigned long long)cnt);
}
__global__ void __launch_bounds__(kThreads)
k_merge_index(const int32_t *__restrict__ nMasters, const label *__restrict__ masters, label *__restrict__ indexOf) {
    const int cnt = *nMasters;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < cnt; i += gridDim.x * kThreads) indexOf[masters[i]] = i;
}
// every merged cell takes a slot of its master's block; overflow (more than K merged cells for one master) is reported
__global__ void __launch_bounds__(kThreads)
k_merge_slots(int64_t nOld, const label *__restrict__ reverseCellMap, const label *__restrict__ indexOf, int K, int *__restrict__ cursor,
              label *__restrict__ slots, int *overflow) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= nOld) return;
    const label r = reverseCellMap[c];
    if (r >= -1) return;
    const label i = indexOf[-r - 2];
    const int pos = atomicAdd(&cursor[i], 1);
    if (pos < K) slots[(size_t)i * K + pos] = (label)c;
    else atomicMax(overflow, pos + 1);
}
__global__ void __launch_bounds__(kThreads)
k_merge_inject(const int32_t *__restrict__ nMasters, const label *__restrict__ masters, int K, const int *__restrict__ cursor,
               label *__restrict__ slots, const double *__restrict__ oldV, double *newV) {
    const int cnt = *nMasters;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < cnt; i += gridDim.x * kThreads) {
        label *row = slots + (size_t)i * K;
        const int d = min(cursor[i], K);
        for (int a = 1; a < d; a++) {                    // ascending old label
            const label v = row[a];
            int j = a - 1;
            while (j >= 0 && row[j] > v) { row[j + 1] = row[j]; j--; }
            row[j + 1] = v;
        }
        const label m = masters[i];
        double v = newV[m];
        for (int a = 0; a < d; a++) v += oldV[row[a]];
        newV[m] = v;
    }
}
