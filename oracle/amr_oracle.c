/*
 * amr_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, single thread per emulated MPI rank) of blastAMR's per-timestep
 * AMR decision path, function by function, for use as the parity oracle of the CUDA path and
 * as the reported CPU baseline ("port").  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the product path
 * (blastamr_b200/libamr_b200.so) never does.
 *
 * Parity status: the reference cannot be compiled in this environment (it needs ESI OpenFOAM
 * v2212 + wmake + MPI, none of which are installed) and ships no golden vectors or
 * known-answer tests (its tests assert cell-count inequalities only,
 * tests/adaptiveFvMeshTests/adaptiveFvMeshTest.C:205,229,347).  The restatement below follows
 * the reference sources line by line -- same loop order, same in-place (Gauss-Seidel) updates,
 * same compare operators -- and is pinned by (i) the reference's own mesh fixtures
 * (tests/testCases/hex3D, hex2D) and its test recipe (box field, fieldValue estimator,
 * thresholds 0.01), (ii) algorithm-independent invariants (2:1 after balance, closure
 * minimality/maximality, idempotence, ascending unique lists).  Numerical parity against a
 * real OpenFOAM build remains UNPINNED ("parity unpinned") and DESIGN.md says so.
 *
 * All file:line citations are relative to /root/reference.
 *
 * A "world" is a set of R emulated MPI ranks living in one process; syncTools calls become
 * direct copies between the ranks' arrays.  R = 1 is the serial reference run.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

typedef int32_t label;     /* WM_LABEL_SIZE=32 */
typedef double scalar;     /* WM_PRECISION_OPTION=DP */
typedef uint8_t boolean;   /* boolList element */

#define OR_API __attribute__((visibility("default")))

static const scalar SMALL = 1e-15;
static const scalar GREAT = 1e15;

/* Emulated ranks are independent between two syncTools calls, exactly like MPI ranks; the
   multi-core CPU baseline runs them on separate threads (one thread per rank, like one MPI
   process per core).  With or_set_threads(1) (default) everything is serial. */
#ifdef _OPENMP
#include <omp.h>
#define DO_PRAGMA(x) _Pragma(#x)
#define PAR_RANKS DO_PRAGMA(omp parallel for schedule(dynamic,1))
#define PAR_RANKS_SUM(v) DO_PRAGMA(omp parallel for schedule(dynamic,1) reduction(+:v))
#define PAR_RANKS_OR(v) DO_PRAGMA(omp parallel for schedule(dynamic,1) reduction(|:v))
OR_API void or_set_threads(int t) { omp_set_num_threads(t < 1 ? 1 : t); }
OR_API int or_max_threads(void) { return omp_get_max_threads(); }
#else
#define PAR_RANKS
#define PAR_RANKS_SUM(v)
#define PAR_RANKS_OR(v)
OR_API void or_set_threads(int t) { (void)t; }
OR_API int or_max_threads(void) { return 1; }
#endif

static inline scalar smax(scalar a, scalar b) { return (a > b) ? a : b; } /* Foam::max */
static inline scalar smin(scalar a, scalar b) { return (a < b) ? a : b; } /* Foam::min */
static inline scalar smag(scalar a) { return fabs(a); }                   /* Foam::mag */

typedef struct {
    int64_t n, f, b, p;
    const label *owner;      /* faceOwner()  [f+b] */
    const label *neighbour;  /* faceNeighbour() [f] */
    int npatch;
    const label *patchStart, *patchSize, *patchType, *patchNbr; /* type 2 = processor (coupled) */
    const label *pcOff, *pcCells;   /* pointCells() */
    const label *cpOff, *cpPts;     /* cellPoints() */
    const boolean *bndPoint;        /* point used by any boundary face (mesh_.faces()[bface]) */
    const label *cellLevel, *pointLevel;
    const label *visible;           /* history_.visibleCells() */
    const label *splitParent;       /* history_.splitCells()[i].parent_ */
    /* mesh.edges() / edgeCells() (hexRef2D: split elements are edges) */
    int64_t nedge;
    const label *edgePts, *ecOff, *ecCells;
    const boolean *bndEdge;       /* edge used by a boundary face (mesh_.faceEdges()[bface]) */
    label *ceOff, *ceEdges;       /* derived: cellEdges() restricted to the listed edges */
    /* mesh.faces() as CSR (polyhedralRefinement split-point rule) */
    const label *faceOff, *facePts;
    /* fvMesh geometry (Lohner only) */
    const scalar *weights, *Sf, *magSf, *deltaCoeffs, *V;
    /* derived: primitiveMesh::cells() */
    label *cfOff, *cfFaces;
} or_mesh;

typedef struct {
    int R;
    or_mesh *m;
} or_world;

OR_API or_world *or_world_create(int R) {
    or_world *w = (or_world *)calloc(1, sizeof(or_world));
    w->R = R;
    w->m = (or_mesh *)calloc(R, sizeof(or_mesh));
    return w;
}

OR_API void or_world_free(or_world *w) {
    if (!w) return;
    for (int r = 0; r < w->R; r++) { free(w->m[r].cfOff); free(w->m[r].cfFaces); free(w->m[r].ceOff); free(w->m[r].ceEdges); }
    free(w->m); free(w);
}

/* primitiveMesh::calcCells: first every face goes to its owner (ascending), then every internal face to its
   neighbour (ascending) -- the order matters only for floating-point sums over cells()[c] (meshSizeObject) */
static void build_cells(or_mesh *m) {
    free(m->cfOff); free(m->cfFaces);
    int64_t n = m->n, nf = m->f + m->b;
    m->cfOff = (label *)calloc(n + 1, sizeof(label));
    for (int64_t i = 0; i < nf; i++) m->cfOff[m->owner[i] + 1]++;
    for (int64_t i = 0; i < m->f; i++) m->cfOff[m->neighbour[i] + 1]++;
    for (int64_t c = 0; c < n; c++) m->cfOff[c + 1] += m->cfOff[c];
    m->cfFaces = (label *)malloc(sizeof(label) * (size_t)(m->cfOff[n] ? m->cfOff[n] : 1));
    label *fill = (label *)malloc(sizeof(label) * (size_t)(n + 1));
    memcpy(fill, m->cfOff, sizeof(label) * (size_t)(n + 1));
    for (int64_t i = 0; i < nf; i++) m->cfFaces[fill[m->owner[i]]++] = (label)i;
    for (int64_t i = 0; i < m->f; i++) m->cfFaces[fill[m->neighbour[i]]++] = (label)i;
    free(fill);
}

OR_API void or_world_set_mesh(or_world *w, int r, int64_t n, int64_t f, int64_t b, int64_t p,
                              const label *owner, const label *neighbour, int npatch,
                              const label *patchStart, const label *patchSize, const label *patchType,
                              const label *patchNbr) {
    or_mesh *m = &w->m[r];
    m->n = n; m->f = f; m->b = b; m->p = p;
    m->owner = owner; m->neighbour = neighbour; m->npatch = npatch;
    m->patchStart = patchStart; m->patchSize = patchSize; m->patchType = patchType; m->patchNbr = patchNbr;
    build_cells(m);
}

OR_API void or_world_set_points(or_world *w, int r, const label *pcOff, const label *pcCells,
                                const label *cpOff, const label *cpPts, const boolean *bndPoint) {
    or_mesh *m = &w->m[r];
    m->pcOff = pcOff; m->pcCells = pcCells; m->cpOff = cpOff; m->cpPts = cpPts; m->bndPoint = bndPoint;
}

OR_API void or_world_set_levels(or_world *w, int r, const label *cellLevel, const label *pointLevel,
                                const label *visible, const label *splitParent) {
    or_mesh *m = &w->m[r];
    m->cellLevel = cellLevel; m->pointLevel = pointLevel; m->visible = visible; m->splitParent = splitParent;
}

OR_API void or_world_set_edges(or_world *w, int r, int64_t nedge, const label *edgePts, const label *ecOff,
                               const label *ecCells, const boolean *bndEdge) {
    or_mesh *m = &w->m[r];
    m->nedge = nedge; m->edgePts = edgePts; m->ecOff = ecOff; m->ecCells = ecCells; m->bndEdge = bndEdge;
    free(m->ceOff); free(m->ceEdges);
    m->ceOff = (label *)calloc((size_t)(m->n + 1), sizeof(label));
    for (int64_t i = 0; i < ecOff[nedge]; i++) m->ceOff[ecCells[i] + 1]++;
    for (int64_t c = 0; c < m->n; c++) m->ceOff[c + 1] += m->ceOff[c];
    m->ceEdges = (label *)malloc(sizeof(label) * (size_t)(ecOff[nedge] + 1));
    label *fill = (label *)malloc(sizeof(label) * (size_t)(m->n + 1));
    memcpy(fill, m->ceOff, sizeof(label) * (size_t)(m->n + 1));
    for (int64_t e = 0; e < nedge; e++)
        for (label i = ecOff[e]; i < ecOff[e + 1]; i++) m->ceEdges[fill[ecCells[i]]++] = (label)e;
    free(fill);
}

OR_API void or_world_set_faces(or_world *w, int r, const label *faceOff, const label *facePts) {
    w->m[r].faceOff = faceOff; w->m[r].facePts = facePts;
}

OR_API void or_world_set_geometry(or_world *w, int r, const scalar *weights, const scalar *Sf, const scalar *magSf,
                                  const scalar *deltaCoeffs, const scalar *V) {
    or_mesh *m = &w->m[r];
    m->weights = weights; m->Sf = Sf; m->magSf = magSf; m->deltaCoeffs = deltaCoeffs; m->V = V;
}

/* --- syncTools restated for ranks living in one address space ----------------------------- */

/* patch on rank s that faces rank r (processor patch pairs list their faces in the same order) */
static int find_back_patch(const or_mesh *ms, int r) {
    for (int q = 0; q < ms->npatch; q++)
        if (ms->patchType[q] == 2 && ms->patchNbr[q] == r) return q;
    return -1;
}

/* syncTools::swapBoundaryFaceList on label lists of size b (one per rank):
   coupled faces receive the other side's value, all other boundary faces keep their own */
static void swap_boundary_label(const or_world *w, label **v) {
    int R = w->R;
    label **snd = (label **)malloc(sizeof(label *) * R);
    for (int r = 0; r < R; r++) {
        snd[r] = (label *)malloc(sizeof(label) * (size_t)(w->m[r].b ? w->m[r].b : 1));
        memcpy(snd[r], v[r], sizeof(label) * (size_t)w->m[r].b);
    }
    PAR_RANKS
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        for (int q = 0; q < m->npatch; q++) if (m->patchType[q] == 2) {
            int s = m->patchNbr[q];
            const or_mesh *ms = &w->m[s];
            int q2 = find_back_patch(ms, r);
            for (label i = 0; i < m->patchSize[q]; i++)
                v[r][m->patchStart[q] - m->f + i] = snd[s][ms->patchStart[q2] - ms->f + i];
        }
    }
    for (int r = 0; r < R; r++) free(snd[r]);
    free(snd);
}

static void swap_boundary_scalar(const or_world *w, scalar **v) {
    int R = w->R;
    scalar **snd = (scalar **)malloc(sizeof(scalar *) * R);
    for (int r = 0; r < R; r++) {
        snd[r] = (scalar *)malloc(sizeof(scalar) * (size_t)(w->m[r].b ? w->m[r].b : 1));
        memcpy(snd[r], v[r], sizeof(scalar) * (size_t)w->m[r].b);
    }
    PAR_RANKS
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        for (int q = 0; q < m->npatch; q++) if (m->patchType[q] == 2) {
            int s = m->patchNbr[q];
            const or_mesh *ms = &w->m[s];
            int q2 = find_back_patch(ms, r);
            for (label i = 0; i < m->patchSize[q]; i++)
                v[r][m->patchStart[q] - m->f + i] = snd[s][ms->patchStart[q2] - ms->f + i];
        }
    }
    for (int r = 0; r < R; r++) free(snd[r]);
    free(snd);
}

/* syncTools::syncFaceList(mesh, faceFlags, orEqOp<bool>()): flags are over all faces (f+b) */
static void sync_face_or(const or_world *w, boolean **face) {
    int R = w->R;
    boolean **snd = (boolean **)malloc(sizeof(boolean *) * R);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        snd[r] = (boolean *)malloc((size_t)(m->b ? m->b : 1));
        memcpy(snd[r], face[r] + m->f, (size_t)m->b);
    }
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        for (int q = 0; q < m->npatch; q++) if (m->patchType[q] == 2) {
            int s = m->patchNbr[q];
            const or_mesh *ms = &w->m[s];
            int q2 = find_back_patch(ms, r);
            for (label i = 0; i < m->patchSize[q]; i++)
                face[r][m->patchStart[q] + i] |= snd[s][ms->patchStart[q2] - ms->f + i];
        }
    }
    for (int r = 0; r < R; r++) free(snd[r]);
    free(snd);
}

/* ======================================================================================= */
/* STAGE 1: error estimators                                                               */
/* ======================================================================================= */

/* errorEstimator::normalize, errorEstimator.C:204-228 (probe part :230-259 = refineProbes,
   host-only, not restated: the configs run with refineProbes no / no probes registered) */
OR_API void or_normalize(int64_t n, scalar *error, scalar lowerRefine, scalar upperRefine,
                         scalar lowerUnrefine, scalar upperUnrefine) {
    for (int64_t celli = 0; celli < n; celli++) {
        if (error[celli] < lowerUnrefine || error[celli] > upperUnrefine) error[celli] = -1.0;
        else if (error[celli] > lowerRefine && error[celli] < upperRefine) error[celli] = 1.0;
        else error[celli] = 0.0;
    }
}

/* gradientEstimator::update threshold part, gradientRange.C:102-109: gMax/gMin over all ranks */
OR_API void or_range_thresholds(const or_world *w, const scalar *const *error, scalar refineThreshold,
                                scalar unrefineThreshold, scalar *thr /*[4]*/, scalar *minMax /*[2]*/) {
    int first = 1;
    scalar maxGrad = 0, minGrad = 0;
    for (int r = 0; r < w->R; r++)
        for (int64_t c = 0; c < w->m[r].n; c++) {
            scalar v = error[r][c];
            if (first) { maxGrad = minGrad = v; first = 0; }
            maxGrad = smax(maxGrad, v);
            minGrad = smin(minGrad, v);
        }
    thr[0] = minGrad + refineThreshold * (maxGrad - minGrad);   /* lowerRefine_ */
    thr[1] = GREAT;                                               /* upperRefine_ */
    thr[2] = minGrad + unrefineThreshold * (maxGrad - minGrad); /* lowerUnrefine_ */
    thr[3] = GREAT;                                               /* upperUnrefine_ */
    if (minMax) { minMax[0] = minGrad; minMax[1] = maxGrad; }
}

/* normalize on every rank of the world (each MPI rank normalises its own field) */
OR_API void or_normalize_world(const or_world *w, scalar **error, scalar lowerRefine, scalar upperRefine,
                               scalar lowerUnrefine, scalar upperUnrefine) {
    PAR_RANKS
    for (int r = 0; r < w->R; r++)
        or_normalize(w->m[r].n, error[r], lowerRefine, upperRefine, lowerUnrefine, upperUnrefine);
}

/*
 * kind 0: delta::update        delta.C:93-131        eT = min(|xo-xn|,0.5) internal, |xp-xn| coupled
 * kind 1: scaledDelta::update  scaledDelta.C:78-136  x = mag(field) - offset (if |offset|>SMALL);
 *                                                    eT = |xo-xn| / max(min(xo,xn), minVal)
 * kind 3: fieldValue::update   fieldValue.C:68-82    error = mag(field)
 * kind 4: coded delta of tutorials/damBreak2D/constant/dynamicMeshDict:24-65 (delta without clamp)
 * x[r] = cell values of rank r (already mag() for non-scalar fields: errorEstimatorTemplates.C:30-46).
 * Coupled patches: patchNeighbourField() == neighbour rank's patchInternalField().
 * params: p0 = minVal (scaledDelta), p1 = offset (scaledDelta)
 */
OR_API void or_estimate(const or_world *w, int kind, const scalar *const *xin, scalar **error,
                        scalar p0, scalar p1) {
    int R = w->R;
    scalar **x = (scalar **)malloc(sizeof(scalar *) * R);
    PAR_RANKS
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        x[r] = (scalar *)malloc(sizeof(scalar) * (size_t)(m->n ? m->n : 1));
        for (int64_t c = 0; c < m->n; c++) {
            scalar v = xin[r][c];
            if (kind == 1 || kind == 3) v = smag(v);           /* getFieldValue -> mag(field) */
            if (kind == 1 && smag(p1) > SMALL) v -= p1;        /* scaledDelta.C:89-93 */
            x[r][c] = v;
        }
    }
    if (kind == 3) {
        for (int r = 0; r < R; r++) { memcpy(error[r], x[r], sizeof(scalar) * (size_t)w->m[r].n); free(x[r]); }
        free(x);
        return;
    }
    /* patchNeighbourField: swap of patchInternalField */
    scalar **fn = (scalar **)malloc(sizeof(scalar *) * R);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        fn[r] = (scalar *)malloc(sizeof(scalar) * (size_t)(m->b ? m->b : 1));
        for (int64_t i = 0; i < m->b; i++) fn[r][i] = x[r][m->owner[m->f + i]];
    }
    swap_boundary_scalar(w, fn);
    PAR_RANKS
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        scalar *err = error[r];
        const scalar *xr = x[r];
        for (int64_t c = 0; c < m->n; c++) err[c] = 0.0;
        for (int64_t facei = 0; facei < m->f; facei++) {
            label own = m->owner[facei], nei = m->neighbour[facei];
            scalar eT;
            if (kind == 0) eT = smin(smag(xr[own] - xr[nei]), 0.5);                       /* delta.C:103 */
            else if (kind == 1) eT = smag(xr[own] - xr[nei]) / smax(smin(xr[own], xr[nei]), p0); /* scaledDelta.C:103-104 */
            else eT = smag(xr[own] - xr[nei]);
            err[own] = smax(err[own], eT);
            err[nei] = smax(err[nei], eT);
        }
        for (int q = 0; q < m->npatch; q++) if (m->patchType[q] == 2) {   /* .coupled() */
            for (label i = 0; i < m->patchSize[q]; i++) {
                int64_t bi = m->patchStart[q] - m->f + i;
                label cell = m->owner[m->patchStart[q] + i];              /* faceCells */
                scalar fp = xr[cell], fnv = fn[r][bi];
                scalar eT;
                if (kind == 1) eT = smag(fp - fnv) / smax(smin(fp, fnv), p0);   /* scaledDelta.C:129-131 */
                else eT = smag(fp - fnv);                                        /* delta.C:126 (no clamp) */
                err[cell] = smax(err[cell], eT);
            }
        }
    }
    for (int r = 0; r < R; r++) { free(x[r]); free(fn[r]); }
    free(x); free(fn);
}

/*
 * Lohner::update, Lohner.C:71-159.
 *
 * xf = cubic<scalar>(mesh_).interpolate(x) is OpenFOAM core (NOT in the reference tree).  Restated
 * from upstream OpenFOAM-v2212 -- PARITY UNPINNED, no reference test or fixture pins these values:
 *   surfaceInterpolationScheme::interpolate(vf, lambdas):   lambda*(vf[P] - vf[N]) + vf[N]
 *   ... (vf, lambdas, ys):                                  lambda*vf[P] + y*vf[N]
 *   coupled patch faces:                                    pLambda*patchInternalField + (1 - pLambda)*patchNeighbourField
 *   linear weights = mesh.weights(); corrected scheme: interpolate = linear + correction
 *   cubic::correction (cubic.H): kSc = l(1 - l(3 - 2l)), kVecP = sqr(1-l) l, kVecN = sqr(l)(l-1);
 *       corr = interpolate(vf, kSc, -kSc)
 *            + (interpolate(gaussGrad(vf), kVecP, kVecN) & Sf)/magSf/nonOrthDeltaCoeffs
 *       (boundary values of corr are zeroed on non-coupled patches; Lohner never reads xf's boundary)
 *   fv::gaussGrad::gradf: for internal faces in ascending order  g[own] += Sf*ssf; g[nei] -= Sf*ssf;
 *       then per patch  g[faceCells] += pSf*pssf;  then g /= V      (ssf = linear interpolate of vf)
 * xb[r]: boundaryField values of x over boundary faces (NULL = zeroGradient: owner cell value);
 * on processor patches the patch field holds the received neighbour values.
 */
OR_API void or_estimate_lohner(const or_world *w, const scalar *const *x, const scalar *const *xbIn, scalar epsilon,
                               scalar **error) {
    int R = w->R;
    scalar **xn = (scalar **)malloc(sizeof(scalar *) * R);   /* patchNeighbourField over boundary faces */
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        xn[r] = (scalar *)malloc(sizeof(scalar) * (size_t)(m->b + 1));
        for (int64_t i = 0; i < m->b; i++) xn[r][i] = x[r][m->owner[m->f + i]];
    }
    swap_boundary_scalar(w, xn);
    PAR_RANKS
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        const scalar *xr = x[r];
        int64_t n = m->n, f = m->f, b = m->b;
        unsigned char *coupled = (unsigned char *)calloc((size_t)(b + 1), 1);
        for (int q = 0; q < m->npatch; q++) if (m->patchType[q] == 2)
            for (label i = 0; i < m->patchSize[q]; i++) coupled[m->patchStart[q] - f + i] = 1;
        /* boundaryField of x */
        scalar *xbf = (scalar *)malloc(sizeof(scalar) * (size_t)(b + 1));
        for (int64_t i = 0; i < b; i++) {
            if (coupled[i]) xbf[i] = xn[r][i];
            else xbf[i] = (xbIn && xbIn[r]) ? xbIn[r][i] : xr[m->owner[f + i]];
        }
        /* gaussGrad(linear) */
        scalar *g = (scalar *)calloc((size_t)(3 * n + 3), sizeof(scalar));
        for (int64_t facei = 0; facei < f; facei++) {
            label P = m->owner[facei], N = m->neighbour[facei];
            scalar ssf = m->weights[facei] * (xr[P] - xr[N]) + xr[N];
            for (int k = 0; k < 3; k++) {
                scalar Sfssf = m->Sf[3 * facei + k] * ssf;
                g[3 * P + k] += Sfssf;
                g[3 * N + k] -= Sfssf;
            }
        }
        for (int64_t i = 0; i < b; i++) {
            int64_t facei = f + i;
            label P = m->owner[facei];
            scalar pssf;
            if (coupled[i]) pssf = m->weights[facei] * xr[P] + (1.0 - m->weights[facei]) * xn[r][i];
            else pssf = xbf[i];
            for (int k = 0; k < 3; k++) g[3 * P + k] += m->Sf[3 * facei + k] * pssf;
        }
        for (int64_t c = 0; c < n; c++) for (int k = 0; k < 3; k++) g[3 * c + k] /= m->V[c];
        /* Lohner.C:84-110 */
        scalar *err = error[r];
        for (int64_t c = 0; c < n; c++) err[c] = 0.0;
        for (int64_t facei = 0; facei < f; facei++) {
            label own = m->owner[facei], nei = m->neighbour[facei];
            scalar l = m->weights[facei];
            scalar kSc = l * (1.0 - l * (3.0 - 2.0 * l));
            scalar kVecP = ((1.0 - l) * (1.0 - l)) * l;
            scalar kVecN = (l * l) * (l - 1.0);
            scalar corr = kSc * xr[own] + (-kSc) * xr[nei];
            scalar gi[3];
            for (int k = 0; k < 3; k++) gi[k] = kVecP * g[3 * own + k] + kVecN * g[3 * nei + k];
            scalar dot = gi[0] * m->Sf[3 * facei] + gi[1] * m->Sf[3 * facei + 1] + gi[2] * m->Sf[3 * facei + 2];
            corr = corr + dot / m->magSf[facei] / m->deltaCoeffs[facei];
            scalar xf = (l * (xr[own] - xr[nei]) + xr[nei]) + corr;
            scalar eT = sqrt(smag(xr[nei] - 2.0 * xf + xr[own])
                             / (smag(xr[nei] - xf) + smag(xf - xr[own])
                                + epsilon * (smag(xr[nei]) + 2.0 * smag(xf) + smag(xr[own]))));
            err[own] = smax(err[own], eT);
            err[nei] = smax(err[nei], eT);
        }
        /* coupled patches, Lohner.C:113-154 */
        for (int64_t i = 0; i < b; i++) if (coupled[i]) {
            label cell = m->owner[f + i];
            scalar xp = xr[cell], xnv = xn[r][i], xb = xbf[i];
            scalar eT = sqrt(smag(xnv - 2.0 * xb + xp)
                             / (smag(xnv - xb) + smag(xb - xp) + epsilon * (smag(xnv) + 2.0 * smag(xb) + smag(xp))));
            err[cell] = smax(err[cell], eT);
        }
        free(coupled); free(xbf); free(g);
    }
    for (int r = 0; r < R; r++) free(xn[r]);
    free(xn);
}

/* ======================================================================================= */
/* STAGE 2: marking                                                                        */
/* ======================================================================================= */

/* fvMeshRefiner::extendMarkedCellsAcrossFaces, fvMeshRefiner.C:866-921
   (identical body: errorEstimator::extendMarkedCellsAcrossFaces, errorEstimator.C:367-424) */
static void extend_across_faces(const or_world *w, boolean **markedCells) {
    int R = w->R;
    boolean **markedFace = (boolean **)malloc(sizeof(boolean *) * R);
    PAR_RANKS
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        markedFace[r] = (boolean *)calloc((size_t)(m->f + m->b + 1), 1);
        for (int64_t cellI = 0; cellI < m->n; cellI++)
            if (markedCells[r][cellI])
                for (label i = m->cfOff[cellI]; i < m->cfOff[cellI + 1]; i++) markedFace[r][m->cfFaces[i]] = 1;
    }
    sync_face_or(w, markedFace);
    PAR_RANKS
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        for (int64_t faceI = 0; faceI < m->f; faceI++)
            if (markedFace[r][faceI]) { markedCells[r][m->owner[faceI]] = 1; markedCells[r][m->neighbour[faceI]] = 1; }
        for (int64_t faceI = m->f; faceI < m->f + m->b; faceI++)
            if (markedFace[r][faceI]) markedCells[r][m->owner[faceI]] = 1;
        free(markedFace[r]);
    }
    free(markedFace);
}

/* fvMeshRefiner::extendMarkedCells, fvMeshRefiner.C:794-863 */
static void extend_marked_cells(const or_world *w, boolean **markedCells, label *const *maxCellLevel,
                                int isTop, int force) {
    int R = w->R;
    boolean **markedFace = (boolean **)malloc(sizeof(boolean *) * R);
    PAR_RANKS
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        markedFace[r] = (boolean *)calloc((size_t)(m->f + m->b + 1), 1);
        for (int64_t celli = 0; celli < m->n; celli++) {
            int go;
            if (force) go = markedCells[r][celli] && (maxCellLevel[r][celli] > m->cellLevel[celli] || !isTop);
            else go = markedCells[r][celli];
            if (go)
                for (label i = m->cfOff[celli]; i < m->cfOff[celli + 1]; i++) markedFace[r][m->cfFaces[i]] = 1;
        }
    }
    sync_face_or(w, markedFace);
    PAR_RANKS
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        for (int64_t facei = 0; facei < m->f; facei++)
            if (markedFace[r][facei]) { markedCells[r][m->owner[facei]] = 1; markedCells[r][m->neighbour[facei]] = 1; }
        for (int64_t facei = m->f; facei < m->f + m->b; facei++)
            if (markedFace[r][facei]) markedCells[r][m->owner[facei]] = 1;
        free(markedFace[r]);
    }
    free(markedFace);
}

/* fvMeshRefiner::extendMarkedCellsAcrossPoints, fvMeshRefiner.C:924-977.
   syncPointList over processor-shared points is restated through global point ids when the
   world has more than one rank (pointGlobal[r] != NULL), else it is the identity. */
static void extend_across_points(const or_world *w, boolean **markedCells, const label *const *pointGlobal,
                                 int64_t nGlobalPoints) {
    int R = w->R;
    boolean **markedPoint = (boolean **)malloc(sizeof(boolean *) * R);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        markedPoint[r] = (boolean *)calloc((size_t)(m->p + 1), 1);
        for (int64_t cellI = 0; cellI < m->n; cellI++)
            if (markedCells[r][cellI])
                for (label i = m->cpOff[cellI]; i < m->cpOff[cellI + 1]; i++) markedPoint[r][m->cpPts[i]] = 1;
    }
    if (R > 1 && pointGlobal) {
        boolean *g = (boolean *)calloc((size_t)(nGlobalPoints + 1), 1);
        for (int r = 0; r < R; r++)
            for (int64_t q = 0; q < w->m[r].p; q++) if (markedPoint[r][q]) g[pointGlobal[r][q]] = 1;
        for (int r = 0; r < R; r++)
            for (int64_t q = 0; q < w->m[r].p; q++) if (g[pointGlobal[r][q]]) markedPoint[r][q] = 1;
        free(g);
    }
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        for (int64_t pointI = 0; pointI < m->p; pointI++)
            if (markedPoint[r][pointI])
                for (label i = m->pcOff[pointI]; i < m->pcOff[pointI + 1]; i++) markedCells[r][m->pcCells[i]] = 1;
        free(markedPoint[r]);
    }
    free(markedPoint);
}

/* errorEstimator::protectPatches, errorEstimator.C:329-365.  patchIds = indices of the patches
   whose names are listed in protectedPatches; nLayers = nPatchesBuffers_ +
   (maxRefinementLevel_-1)*nBufferCells_ (errorEstimator.C:347).  Returns nProtectedCells per rank
   in nProt (may be NULL). */
OR_API void or_protect_patches(const or_world *w, const label *patchIds, int nPatchIds, int nLayers,
                               scalar **error, int64_t *nProt) {
    int R = w->R;
    if (nPatchIds == 0) { if (nProt) for (int r = 0; r < R; r++) nProt[r] = 0; return; }
    boolean **prot = (boolean **)malloc(sizeof(boolean *) * R);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        prot[r] = (boolean *)calloc((size_t)(m->n + 1), 1);
        for (int k = 0; k < nPatchIds; k++) {
            int q = patchIds[k];
            for (label i = 0; i < m->patchSize[q]; i++) prot[r][m->owner[m->patchStart[q] + i]] = 1;
        }
    }
    for (int i = 0; i < nLayers; i++) extend_across_faces(w, prot);
    for (int r = 0; r < R; r++) {
        int64_t cnt = 0;
        for (int64_t c = 0; c < w->m[r].n; c++) if (prot[r][c]) { error[r][c] = 0.0; cnt++; }
        if (nProt) nProt[r] = cnt;
        free(prot[r]);
    }
    free(prot);
}

/* hexRef::faceConsistentRefinement, hexRef.C:700-783 (one in-place sweep per rank + swap).
   Returns the global nChanged (the reduce(sumOp) of hexRef.C:1424). */
static int64_t face_consistent_refinement(const or_world *w, int maxSet, boolean **refineCell) {
    int R = w->R;
    int64_t nChangedTot = 0;
    label **neiLevel = (label **)malloc(sizeof(label *) * R);
    PAR_RANKS_SUM(nChangedTot)
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        boolean *rc = refineCell[r];
        int64_t nChanged = 0;
        for (int64_t facei = 0; facei < m->f; facei++) {
            label own = m->owner[facei];
            label ownLevel = m->cellLevel[own] + rc[own];
            label nei = m->neighbour[facei];
            label neiL = m->cellLevel[nei] + rc[nei];
            if (ownLevel > (neiL + 1)) {
                if (maxSet) rc[nei] = 1; else rc[own] = 0;
                nChanged++;
            } else if (neiL > (ownLevel + 1)) {
                if (maxSet) rc[own] = 1; else rc[nei] = 0;
                nChanged++;
            }
        }
        neiLevel[r] = (label *)malloc(sizeof(label) * (size_t)(m->b ? m->b : 1));
        for (int64_t i = 0; i < m->b; i++) {
            label own = m->owner[i + m->f];
            neiLevel[r][i] = m->cellLevel[own] + rc[own];
        }
        nChangedTot += nChanged;
    }
    swap_boundary_label(w, neiLevel);
    PAR_RANKS_SUM(nChangedTot)
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        boolean *rc = refineCell[r];
        for (int64_t i = 0; i < m->b; i++) {
            label own = m->owner[i + m->f];
            label ownLevel = m->cellLevel[own] + rc[own];
            if (ownLevel > (neiLevel[r][i] + 1)) {
                if (!maxSet) { rc[own] = 0; nChangedTot++; }
            } else if (neiLevel[r][i] > (ownLevel + 1)) {
                if (maxSet) { rc[own] = 1; nChangedTot++; }
            }
        }
        free(neiLevel[r]);
    }
    free(neiLevel);
    return nChangedTot;
}

/* hexRef::consistentRefinement, hexRef.C:1402-1468 (flags in, flags out; the ascending list is
   produced by the caller).  Returns the number of sweeps the reference would have executed. */
OR_API int or_consistent_refinement(const or_world *w, int maxSet, boolean **refineCell) {
    int sweeps = 0;
    while (1) {
        int64_t nChanged = face_consistent_refinement(w, maxSet, refineCell);
        sweeps++;
        if (nChanged == 0) break;
    }
    return sweeps;
}

/* fvMeshHexRefiner::calculateProtectedCells, fvMeshHexRefiner.C:60-179.
   protectedCell[r] == NULL on all ranks -> unrefineable stays empty (returns 0). */
static int calculate_protected_cells(const or_world *w, const boolean *const *protectedCell,
                                     boolean **unrefineableCell) {
    int R = w->R;
    int64_t tot = 0;
    for (int r = 0; r < R; r++) if (protectedCell && protectedCell[r]) tot += w->m[r].n;
    if (!tot) return 0;
    label **neiLevel = (label **)malloc(sizeof(label *) * R);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        if (protectedCell[r]) memcpy(unrefineableCell[r], protectedCell[r], (size_t)m->n);
        else memset(unrefineableCell[r], 0, (size_t)m->n);
        neiLevel[r] = (label *)malloc(sizeof(label) * (size_t)(m->b ? m->b : 1));
        for (int64_t i = 0; i < m->b; i++) neiLevel[r][i] = m->cellLevel[m->owner[m->f + i]];
    }
    swap_boundary_label(w, neiLevel);
    boolean **seedFace = (boolean **)malloc(sizeof(boolean *) * R);
    while (1) {
        for (int r = 0; r < R; r++) {
            const or_mesh *m = &w->m[r];
            const label *cellLevel = m->cellLevel;
            boolean *u = unrefineableCell[r];
            seedFace[r] = (boolean *)calloc((size_t)(m->f + m->b + 1), 1);
            for (int64_t facei = 0; facei < m->f; facei++) {
                label own = m->owner[facei], nei = m->neighbour[facei];
                if (u[own] && (cellLevel[nei] > cellLevel[own])) seedFace[r][facei] = 1;
                else if (u[nei] && (cellLevel[own] > cellLevel[nei])) seedFace[r][facei] = 1;
            }
            for (int64_t facei = m->f; facei < m->f + m->b; facei++) {
                label own = m->owner[facei];
                if (u[own] && (neiLevel[r][facei - m->f] > cellLevel[own])) seedFace[r][facei] = 1;
            }
        }
        sync_face_or(w, seedFace);
        int hasExtended = 0;
        for (int r = 0; r < R; r++) {
            const or_mesh *m = &w->m[r];
            boolean *u = unrefineableCell[r];
            for (int64_t facei = 0; facei < m->f; facei++) if (seedFace[r][facei]) {
                label own = m->owner[facei];
                if (u[own] == 0) { u[own] = 1; hasExtended = 1; }
                label nei = m->neighbour[facei];
                if (u[nei] == 0) { u[nei] = 1; hasExtended = 1; }
            }
            for (int64_t facei = m->f; facei < m->f + m->b; facei++) if (seedFace[r][facei]) {
                label own = m->owner[facei];
                if (u[own] == 0) { u[own] = 1; hasExtended = 1; }
            }
            free(seedFace[r]);
        }
        if (!hasExtended) break;
    }
    for (int r = 0; r < R; r++) free(neiLevel[r]);
    free(neiLevel); free(seedFace);
    return 1;
}

/*
 * Refinement branch of fvMeshHexRefiner::refine, fvMeshHexRefiner.C:1497-1541:
 *   selectRefineCandidates (fvMeshRefiner.C:170-187)
 *   nRefinementBufferLayers x extendMarkedCells(refineCell, maxRefinement, i==0, true) (:1509-1512)
 *   clear cells on protectedPatches_ (:1514-1523) and protectedCell_ (:1524-1533)
 *   selectRefineCells (fvMeshHexRefiner.C:306-391) incl. calculateProtectedCells, the maxCells
 *   truncation and hexRef::consistentRefinement(candidates, true).
 * In:  error[r] (normalised), maxCellLevel[r] (per cell), protectedCell[r] (may be NULL),
 *      protectedPatchIds (refiner's protectedPatches_), maxCells.
 * Out: refineCell[r]  = the marking AFTER buffer layers/protections (what the reference keeps in
 *                       `refineCell` for the later unrefinement stage),
 *      refineSet[r]   = flags of the consistent set (cellsToRefine = ascending scan of it).
 * Returns the number of faceConsistentRefinement sweeps.
 */
OR_API int or_hex_select_refine(const or_world *w, const scalar *const *error, scalar lowerRefineLevel,
                                scalar upperRefineLevel, label *const *maxCellLevel,
                                int nRefinementBufferLayers, const boolean *const *protectedCell,
                                const label *protectedPatchIds, int nProtectedPatches, int64_t maxCells,
                                boolean **refineCell, boolean **refineSet) {
    int R = w->R;
    /* selectRefineCandidates, fvMeshRefiner.C:170-187 (inclusive compares) */
    PAR_RANKS
    for (int r = 0; r < R; r++)
        for (int64_t cellI = 0; cellI < w->m[r].n; cellI++)
            refineCell[r][cellI] = (error[r][cellI] >= lowerRefineLevel) && (error[r][cellI] <= upperRefineLevel);
    for (int i = 0; i < nRefinementBufferLayers; i++)
        extend_marked_cells(w, refineCell, maxCellLevel, i == 0, 1);
    int anyProt = 0;
    for (int r = 0; r < R; r++) if (protectedCell && protectedCell[r]) anyProt = 1;
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        for (int k = 0; k < nProtectedPatches; k++) {
            int q = protectedPatchIds[k];
            for (label i = 0; i < m->patchSize[q]; i++) refineCell[r][m->owner[m->patchStart[q] + i]] = 0;
        }
        if (anyProt && protectedCell[r])
            for (int64_t c = 0; c < m->n; c++) if (protectedCell[r][c]) refineCell[r][c] = 0;
    }
    /* selectRefineCells, fvMeshHexRefiner.C:306-391 */
    int64_t nTotalCells = 0;
    for (int r = 0; r < R; r++) nTotalCells += w->m[r].n;
    int64_t nTotToRefine = (maxCells - nTotalCells) / 7;
    boolean **unrefineable = (boolean **)malloc(sizeof(boolean *) * R);
    for (int r = 0; r < R; r++) unrefineable[r] = (boolean *)calloc((size_t)(w->m[r].n + 1), 1);
    int haveU = calculate_protected_cells(w, protectedCell, unrefineable);
    int64_t nCandidates = 0;
    PAR_RANKS_SUM(nCandidates)
    for (int r = 0; r < R; r++)
        for (int64_t c = 0; c < w->m[r].n; c++) nCandidates += refineCell[r][c] ? 1 : 0;
    for (int r = 0; r < R; r++) memset(refineSet[r], 0, (size_t)w->m[r].n);
    if (nCandidates < nTotToRefine) {
        PAR_RANKS
        for (int r = 0; r < R; r++) {
            const or_mesh *m = &w->m[r];
            for (int64_t celli = 0; celli < m->n; celli++)
                if (m->cellLevel[celli] < maxCellLevel[r][celli] && refineCell[r][celli]
                    && (!haveU || !unrefineable[r][celli]))
                    refineSet[r][celli] = 1;
        }
    } else {
        label maxLevel = 0; /* gMax(maxRefinement) */
        int first = 1;
        for (int r = 0; r < R; r++)
            for (int64_t c = 0; c < w->m[r].n; c++)
                if (first || maxCellLevel[r][c] > maxLevel) { maxLevel = maxCellLevel[r][c]; first = 0; }
        for (label level = 0; level < maxLevel; level++) {
            int64_t tot = 0;
            for (int r = 0; r < R; r++) {
                const or_mesh *m = &w->m[r];
                for (int64_t celli = 0; celli < m->n; celli++)
                    if (m->cellLevel[celli] == level && refineCell[r][celli] && (!haveU || !unrefineable[r][celli]))
                        refineSet[r][celli] = 1;
                for (int64_t celli = 0; celli < m->n; celli++) tot += refineSet[r][celli];
            }
            if (tot > nTotToRefine) break;
        }
    }
    for (int r = 0; r < R; r++) free(unrefineable[r]);
    free(unrefineable);
    return or_consistent_refinement(w, 1, refineSet);
}

/* Update of refineCell after the topology change, fvMeshHexRefiner.C:1562-1586 */
OR_API void or_remap_refine_cell(int64_t nNew, const label *cellMap, const label *reverseCellMap,
                                 const boolean *refineCell, boolean *newRefineCell) {
    for (int64_t celli = 0; celli < nNew; celli++) {
        label oldCelli = cellMap[celli];
        if (oldCelli < 0) newRefineCell[celli] = 1;
        else if (reverseCellMap[oldCelli] != celli) newRefineCell[celli] = 1;
        else newRefineCell[celli] = refineCell[oldCelli];
    }
}

/* ======================================================================================= */
/* STAGE 3b: unrefinement selection (3-D hex)                                              */
/* ======================================================================================= */

/* fvMeshRefiner::maxCellField, fvMeshRefiner.C:110-125 */
OR_API void or_max_cell_field(const or_world *w, int r, const scalar *vFld, scalar *pFld) {
    const or_mesh *m = &w->m[r];
    for (int64_t pointi = 0; pointi < m->p; pointi++) {
        pFld[pointi] = -GREAT;
        for (label i = m->pcOff[pointi]; i < m->pcOff[pointi + 1]; i++)
            pFld[pointi] = smax(pFld[pointi], vFld[m->pcCells[i]]);
    }
}

/* hexRef3D::getSplitElems, hexRef3D.C:2183-2321 -> flags over points (splitMaster >= 0) */
OR_API void or_get_split_points(const or_world *w, int r, boolean *isSplit) {
    const or_mesh *m = &w->m[r];
    label *splitMaster = (label *)malloc(sizeof(label) * (size_t)(m->p + 1));
    label *splitMasterLevel = (label *)calloc((size_t)(m->p + 1), sizeof(label));
    for (int64_t pointi = 0; pointi < m->p; pointi++) {
        splitMaster[pointi] = -1;
        if (m->pcOff[pointi + 1] - m->pcOff[pointi] != 8) splitMaster[pointi] = -2;
    }
    for (int64_t celli = 0; celli < m->n; celli++) {
        label vis = m->visible[celli];
        label parentIndex = (vis != -1) ? m->splitParent[vis] : -1;   /* history_.parentIndex(celli) */
        if (vis != -1 && parentIndex >= 0) {
            for (label i = m->cpOff[celli]; i < m->cpOff[celli + 1]; i++) {
                label pointi = m->cpPts[i];
                label masterCelli = splitMaster[pointi];
                if (masterCelli == -1) {
                    splitMaster[pointi] = parentIndex;
                    splitMasterLevel[pointi] = m->cellLevel[celli] - 1;
                } else if (masterCelli == -2) {
                } else if ((masterCelli != parentIndex) || (splitMasterLevel[pointi] != m->cellLevel[celli] - 1)) {
                    splitMaster[pointi] = -2;
                }
            }
        } else {
            for (label i = m->cpOff[celli]; i < m->cpOff[celli + 1]; i++) splitMaster[m->cpPts[i]] = -2;
        }
    }
    /* "Unmark boundary faces": every point of every boundary face (hexRef3D.C:2279-2294) */
    for (int64_t pointi = 0; pointi < m->p; pointi++) if (m->bndPoint[pointi]) splitMaster[pointi] = -2;
    for (int64_t pointi = 0; pointi < m->p; pointi++) isSplit[pointi] = splitMaster[pointi] >= 0;
    free(splitMaster); free(splitMasterLevel);
}

/* hexRef3D::consistentUnrefinement, hexRef3D.C:1719-1953 (maxSet = false).  Returns sweeps,
   or -1 on the reference's "problem" FatalError (hexRef3D.C:1811-1815). */
/* kind 0: elements are points (hexRef3D.C:1719-1953); kind 1: elements are edges (hexRef2D.C:1929-2164,
   same body with mesh_.edgeCells() in place of mesh_.pointCells()) */
static int consistent_unrefinement_elems(const or_world *w, int kind, boolean **unrefinePoint);
OR_API int or_consistent_unrefinement(const or_world *w, boolean **unrefinePoint) {
    return consistent_unrefinement_elems(w, 0, unrefinePoint);
}
OR_API int or_consistent_unrefinement_edges(const or_world *w, boolean **unrefineEdge) {
    return consistent_unrefinement_elems(w, 1, unrefineEdge);
}
#define EL_N(m) (kind ? (m)->nedge : (m)->p)
#define EL_OFF(m) (kind ? (m)->ecOff : (m)->pcOff)
#define EL_CELLS(m) (kind ? (m)->ecCells : (m)->pcCells)
static int consistent_unrefinement_elems(const or_world *w, int kind, boolean **unrefinePoint) {
    int R = w->R;
    int sweeps = 0;
    boolean **unrefineCell = (boolean **)malloc(sizeof(boolean *) * R);
    label **neiLevel = (label **)malloc(sizeof(label *) * R);
    for (int r = 0; r < R; r++) {
        unrefineCell[r] = (boolean *)malloc((size_t)(w->m[r].n + 1));
        neiLevel[r] = (label *)malloc(sizeof(label) * (size_t)(w->m[r].b + 1));
    }
    int fail = 0;
    while (!fail) {
        int64_t nChanged = 0;
        _Pragma("omp parallel for schedule(dynamic,1) reduction(+:nChanged) reduction(|:fail)")
        for (int r = 0; r < R; r++) {
            const or_mesh *m = &w->m[r];
            boolean *uc = unrefineCell[r];
            memset(uc, 0, (size_t)m->n);
            for (int64_t pointi = 0; pointi < EL_N(m); pointi++)
                if (unrefinePoint[r][pointi])
                    for (label j = EL_OFF(m)[pointi]; j < EL_OFF(m)[pointi + 1]; j++) uc[EL_CELLS(m)[j]] = 1;
            for (int64_t facei = 0; facei < m->f; facei++) {
                label own = m->owner[facei];
                label ownLevel = m->cellLevel[own] - uc[own];
                label nei = m->neighbour[facei];
                label neiL = m->cellLevel[nei] - uc[nei];
                if (ownLevel < (neiL - 1)) {
                    if (uc[own] == 0) { fail = 1; break; }
                    uc[own] = 0; nChanged++;
                } else if (neiL < (ownLevel - 1)) {
                    if (uc[nei] == 0) { fail = 1; break; }
                    uc[nei] = 0; nChanged++;
                }
            }
            for (int64_t i = 0; i < m->b; i++) {
                label own = m->owner[i + m->f];
                neiLevel[r][i] = m->cellLevel[own] - uc[own];
            }
        }
        if (fail) break;
        swap_boundary_label(w, neiLevel);
        for (int r = 0; r < R; r++) {
            const or_mesh *m = &w->m[r];
            boolean *uc = unrefineCell[r];
            for (int64_t i = 0; i < m->b; i++) {
                label own = m->owner[i + m->f];
                label ownLevel = m->cellLevel[own] - uc[own];
                if (ownLevel < (neiLevel[r][i] - 1)) {
                    if (uc[own] == 0) { fail = 1; break; }
                    uc[own] = 0; nChanged++;
                }
            }
        }
        if (fail) break;
        sweeps++;
        if (nChanged == 0) break;
        for (int r = 0; r < R; r++) {
            const or_mesh *m = &w->m[r];
            for (int64_t pointi = 0; pointi < EL_N(m); pointi++)
                if (unrefinePoint[r][pointi])
                    for (label j = EL_OFF(m)[pointi]; j < EL_OFF(m)[pointi + 1]; j++)
                        if (!unrefineCell[r][EL_CELLS(m)[j]]) { unrefinePoint[r][pointi] = 0; break; }
        }
    }
    for (int r = 0; r < R; r++) { free(unrefineCell[r]); free(neiLevel[r]); }
    free(unrefineCell); free(neiLevel);
    return fail ? -1 : sweeps;
}

/*
 * Unrefinement branch of fvMeshHexRefiner::refine, fvMeshHexRefiner.C:1591-1633 with
 * hexRef3D::selectUnrefineElems, hexRef3D.C:2324-2381:
 *   nUnrefinementBufferLayers x extendMarkedCellsAcrossFaces(refineCell)
 *   set protectedCell_ / cells on protectedPatches_ in refineCell
 *   pFld = maxCellField(error); splitPoints = getSplitElems();
 *   keep point if pFld < unrefineLevel and no pointCell is marked; consistentUnrefinement.
 * refineCell[r] is updated in place (as in the reference); unrefPoint[r] receives flags of the
 * final consistent split points; splitPoint[r] (may be NULL) the flags of all split points.
 * Returns sweeps or -1.
 */
OR_API int or_hex_select_unrefine(const or_world *w, const scalar *const *error, scalar unrefineLevel,
                                  int nUnrefinementBufferLayers, const boolean *const *protectedCell,
                                  const label *protectedPatchIds, int nProtectedPatches,
                                  boolean **refineCell, boolean **unrefPoint, boolean **splitPoint) {
    int R = w->R;
    for (int i = 0; i < nUnrefinementBufferLayers; i++) extend_across_faces(w, refineCell);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        if (protectedCell && protectedCell[r])
            for (int64_t c = 0; c < m->n; c++) if (protectedCell[r][c]) refineCell[r][c] = 1;
        for (int k = 0; k < nProtectedPatches; k++) {
            int q = protectedPatchIds[k];
            for (label i = 0; i < m->patchSize[q]; i++) refineCell[r][m->owner[m->patchStart[q] + i]] = 1;
        }
    }
    PAR_RANKS
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        scalar *pFld = (scalar *)malloc(sizeof(scalar) * (size_t)(m->p + 1));
        boolean *isSplit = (boolean *)malloc((size_t)(m->p + 1));
        or_max_cell_field(w, r, error[r], pFld);
        or_get_split_points(w, r, isSplit);
        for (int64_t pointi = 0; pointi < m->p; pointi++) {
            unrefPoint[r][pointi] = 0;
            if (!isSplit[pointi]) continue;
            if (pFld[pointi] < unrefineLevel) {
                int hasMarked = 0;
                for (label j = m->pcOff[pointi]; j < m->pcOff[pointi + 1]; j++)
                    if (refineCell[r][m->pcCells[j]]) { hasMarked = 1; break; }
                if (!hasMarked) unrefPoint[r][pointi] = 1;
            }
        }
        if (splitPoint && splitPoint[r]) memcpy(splitPoint[r], isSplit, (size_t)m->p);
        free(pFld); free(isSplit);
    }
    return or_consistent_unrefinement(w, unrefPoint);
}

/* hexRef2D::getSplitElems, hexRef2D.C:2202-2342 -> flags over edges */
OR_API void or_get_split_edges(const or_world *w, int r, boolean *isSplit) {
    const or_mesh *m = &w->m[r];
    label *splitMaster = (label *)malloc(sizeof(label) * (size_t)(m->nedge + 1));
    label *splitMasterLevel = (label *)calloc((size_t)(m->nedge + 1), sizeof(label));
    for (int64_t edgei = 0; edgei < m->nedge; edgei++) {
        splitMaster[edgei] = -1;
        if (m->ecOff[edgei + 1] - m->ecOff[edgei] != 4) splitMaster[edgei] = -2;
    }
    for (int64_t celli = 0; celli < m->n; celli++) {
        label vis = m->visible[celli];
        label parentIndex = (vis != -1) ? m->splitParent[vis] : -1;
        if (vis != -1 && parentIndex >= 0) {
            for (label i = m->ceOff[celli]; i < m->ceOff[celli + 1]; i++) {
                label edgei = m->ceEdges[i];
                label masterCelli = splitMaster[edgei];
                if (masterCelli == -1) {
                    splitMaster[edgei] = parentIndex;
                    splitMasterLevel[edgei] = m->cellLevel[celli] - 1;
                } else if (masterCelli == -2) {
                } else if ((masterCelli != parentIndex) || (splitMasterLevel[edgei] != m->cellLevel[celli] - 1)) {
                    splitMaster[edgei] = -2;
                }
            }
        } else {
            for (label i = m->ceOff[celli]; i < m->ceOff[celli + 1]; i++) splitMaster[m->ceEdges[i]] = -2;
        }
    }
    /* "Unmark boundary faces": every edge of every boundary face (hexRef2D.C:2296-2311) */
    for (int64_t edgei = 0; edgei < m->nedge; edgei++) if (m->bndEdge[edgei]) splitMaster[edgei] = -2;
    for (int64_t edgei = 0; edgei < m->nedge; edgei++) isSplit[edgei] = splitMaster[edgei] >= 0;
    free(splitMaster); free(splitMasterLevel);
}

/*
 * Unrefinement branch of fvMeshHexRefiner::refine (fvMeshHexRefiner.C:1591-1633) with the 2-D cutter:
 * hexRef2D::selectUnrefineElems, hexRef2D.C:1860-1927 -- a split edge is kept when EITHER of its end
 * points has pFld < unrefineLevel and no marked pointCell (the 3-D code tests the single split point).
 */
OR_API int or_hex2d_select_unrefine(const or_world *w, const scalar *const *error, scalar unrefineLevel,
                                    int nUnrefinementBufferLayers, const boolean *const *protectedCell,
                                    const label *protectedPatchIds, int nProtectedPatches,
                                    boolean **refineCell, boolean **unrefEdge, boolean **splitEdge) {
    int R = w->R;
    for (int i = 0; i < nUnrefinementBufferLayers; i++) extend_across_faces(w, refineCell);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        if (protectedCell && protectedCell[r])
            for (int64_t c = 0; c < m->n; c++) if (protectedCell[r][c]) refineCell[r][c] = 1;
        for (int k = 0; k < nProtectedPatches; k++) {
            int q = protectedPatchIds[k];
            for (label i = 0; i < m->patchSize[q]; i++) refineCell[r][m->owner[m->patchStart[q] + i]] = 1;
        }
    }
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        scalar *pFld = (scalar *)malloc(sizeof(scalar) * (size_t)(m->p + 1));
        boolean *isSplit = (boolean *)malloc((size_t)(m->nedge + 1));
        or_max_cell_field(w, r, error[r], pFld);
        or_get_split_edges(w, r, isSplit);
        for (int64_t edgej = 0; edgej < m->nedge; edgej++) {
            unrefEdge[r][edgej] = 0;
            if (!isSplit[edgej]) continue;
            for (int k = 0; k < 2; k++) {
                label pointi = m->edgePts[2 * edgej + k];
                int hasMarked = 1;
                if (pFld[pointi] < unrefineLevel) {
                    hasMarked = 0;
                    for (label j = m->pcOff[pointi]; j < m->pcOff[pointi + 1]; j++)
                        if (refineCell[r][m->pcCells[j]]) { hasMarked = 1; break; }
                }
                if (!hasMarked) { unrefEdge[r][edgej] = 1; break; }
            }
        }
        if (splitEdge && splitEdge[r]) memcpy(splitEdge[r], isSplit, (size_t)m->nedge);
        free(pFld); free(isSplit);
    }
    return or_consistent_unrefinement_edges(w, unrefEdge);
}

/* ======================================================================================= */
/* polyRefiner (fvMeshPolyRefiner + refinement + polyhedralRefinement), 3-D                */
/* ======================================================================================= */

/* refinement::faceConsistentRefinement, refinement.C:202-302 (counts real flips: boolList::set/unset
   return whether the value changed) */
static int64_t poly_face_consistent_refinement(const or_world *w, int maxSet, boolean **refineCell) {
    int R = w->R;
    int64_t nChangedTot = 0;
    label **neiLevel = (label **)malloc(sizeof(label *) * R);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        boolean *rc = refineCell[r];
        for (int64_t faceI = 0; faceI < m->f; faceI++) {
            label own = m->owner[faceI], nei = m->neighbour[faceI];
            label ownLevel = m->cellLevel[own] + rc[own];
            label neiL = m->cellLevel[nei] + rc[nei];
            if (ownLevel > (neiL + 1)) {
                if (maxSet) { nChangedTot += !rc[nei]; rc[nei] = 1; } else { nChangedTot += rc[own]; rc[own] = 0; }
            } else if (neiL > (ownLevel + 1)) {
                if (maxSet) { nChangedTot += !rc[own]; rc[own] = 1; } else { nChangedTot += rc[nei]; rc[nei] = 0; }
            }
        }
        neiLevel[r] = (label *)malloc(sizeof(label) * (size_t)(m->b + 1));
        for (int64_t i = 0; i < m->b; i++) { label own = m->owner[i + m->f]; neiLevel[r][i] = m->cellLevel[own] + rc[own]; }
    }
    swap_boundary_label(w, neiLevel);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        boolean *rc = refineCell[r];
        for (int64_t i = 0; i < m->b; i++) {
            label own = m->owner[i + m->f];
            label ownLevel = m->cellLevel[own] + rc[own];
            if (ownLevel > (neiLevel[r][i] + 1)) { if (!maxSet) { nChangedTot += rc[own]; rc[own] = 0; } }
            else if (neiLevel[r][i] > (ownLevel + 1)) { if (maxSet) { nChangedTot += !rc[own]; rc[own] = 1; } }
        }
        free(neiLevel[r]);
    }
    free(neiLevel);
    return nChangedTot;
}

/* refinement::edgeConsistentRefinement, refinement.C:305-387 (local: every unordered pair of edgeCells) */
static int64_t poly_edge_consistent_refinement(const or_world *w, int maxSet, boolean **refineCell) {
    int64_t nChanged = 0;
    for (int r = 0; r < w->R; r++) {
        const or_mesh *m = &w->m[r];
        boolean *rc = refineCell[r];
        for (int64_t edgeI = 0; edgeI < m->nedge; edgeI++)
            for (label i = m->ecOff[edgeI]; i < m->ecOff[edgeI + 1]; i++) {
                label cellI = m->ecCells[i];
                for (label j = i + 1; j < m->ecOff[edgeI + 1]; j++) {
                    label cellJ = m->ecCells[j];
                    label cellILevel = m->cellLevel[cellI] + rc[cellI];
                    label cellJLevel = m->cellLevel[cellJ] + rc[cellJ];
                    if (cellILevel > (cellJLevel + 1)) {
                        if (maxSet) { nChanged += !rc[cellJ]; rc[cellJ] = 1; } else { nChanged += rc[cellI]; rc[cellI] = 0; }
                    } else if (cellJLevel > (cellILevel + 1)) {
                        if (maxSet) { nChanged += !rc[cellI]; rc[cellI] = 1; } else { nChanged += rc[cellJ]; rc[cellJ] = 0; }
                    }
                }
            }
    }
    return nChanged;
}

/* refinement::consistentRefinement, refinement.C:744-810 */
OR_API int or_poly_consistent_refinement(const or_world *w, int maxSet, int edgeBasedConsistency, boolean **refineCell) {
    int sweeps = 0;
    while (1) {
        int64_t nChanged = 0;
        if (edgeBasedConsistency) nChanged += poly_edge_consistent_refinement(w, maxSet, refineCell);
        nChanged += poly_face_consistent_refinement(w, maxSet, refineCell);
        sweeps++;
        if (nChanged == 0) break;
    }
    return sweeps;
}

/* Refinement branch of fvMeshPolyRefiner::refine, fvMeshPolyRefiner.C:321-370: selectRefineCandidates,
   layers x extendMarkedCells(.., force_), clear protectedPatches, base selectRefineCells
   (fvMeshRefiner.C:190-219), refinement::consistentRefinement(.., true). */
OR_API int or_poly_select_refine(const or_world *w, const scalar *const *error, scalar lowerRefineLevel, scalar upperRefineLevel,
                                 label *const *maxCellLevel, int nRefinementBufferLayers, int force,
                                 const label *protectedPatchIds, int nProtectedPatches, int edgeBasedConsistency,
                                 boolean **refineCell, boolean **refineSet) {
    int R = w->R;
    for (int r = 0; r < R; r++)
        for (int64_t c = 0; c < w->m[r].n; c++)
            refineCell[r][c] = (error[r][c] >= lowerRefineLevel) && (error[r][c] <= upperRefineLevel);
    for (int i = 0; i < nRefinementBufferLayers; i++) extend_marked_cells(w, refineCell, maxCellLevel, i == 0, force);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        for (int k = 0; k < nProtectedPatches; k++) {
            int q = protectedPatchIds[k];
            for (label i = 0; i < m->patchSize[q]; i++) refineCell[r][m->owner[m->patchStart[q] + i]] = 0;
        }
        for (int64_t c = 0; c < m->n; c++) refineSet[r][c] = (m->cellLevel[c] < maxCellLevel[r][c]) && refineCell[r][c];
    }
    return or_poly_consistent_refinement(w, 1, edgeBasedConsistency, refineSet);
}

/* polyhedralRefinement::consistentUnrefinement PART 1 (polyhedralRefinement.C:2096-2184): split point <=>
   pointLevel >= 1, every point of every face using it has the same pointLevel, not on a boundary face */
OR_API void or_poly_split_points(const or_world *w, int r, boolean *isSplit) {
    const or_mesh *m = &w->m[r];
    boolean *bad = (boolean *)calloc((size_t)(m->p + 1), 1);
    int64_t nf = m->f + m->b;
    /* a point is rejected iff one of its faces carries a point of a different level <=> it lies on a face
       whose points do not all share one level (then every point of that face sees a different level) */
    for (int64_t faceI = 0; faceI < nf; faceI++) {
        label s = m->faceOff[faceI], e = m->faceOff[faceI + 1];
        int uniform = 1;
        for (label j = s + 1; j < e; j++) if (m->pointLevel[m->facePts[j]] != m->pointLevel[m->facePts[s]]) { uniform = 0; break; }
        if (!uniform) for (label j = s; j < e; j++) bad[m->facePts[j]] = 1;
        if (faceI >= m->f) for (label j = s; j < e; j++) bad[m->facePts[j]] = 1;   /* boundary faces, :2175-2184 */
    }
    for (int64_t pointI = 0; pointI < m->p; pointI++) isSplit[pointI] = (m->pointLevel[pointI] >= 1) && !bad[pointI];
    free(bad);
}

/* prismatic2DRefinement::consistentUnrefinement PART 1 (prismatic2DRefinement.C:2744-2849): split point <=>
   pointLevel > 0 and both end points of every edge using the point carry that same pointLevel; then the
   points of processor-patch faces are removed (no other boundary exclusion: in 2-D every point lies on
   the empty front/back patches) */
OR_API void or_prism_split_points(const or_world *w, int r, boolean *isSplit) {
    const or_mesh *m = &w->m[r];
    boolean *bad = (boolean *)calloc((size_t)(m->p + 1), 1);
    /* pointEdges walk restated edge-centric: a point fails iff one of its edges joins two different levels */
    for (int64_t edgeI = 0; edgeI < m->nedge; edgeI++) {
        label a = m->edgePts[2 * edgeI], b = m->edgePts[2 * edgeI + 1];
        if (m->pointLevel[a] != m->pointLevel[b]) { bad[a] = 1; bad[b] = 1; }
    }
    for (int64_t pointI = 0; pointI < m->p; pointI++) isSplit[pointI] = (m->pointLevel[pointI] > 0) && !bad[pointI];
    for (int q = 0; q < m->npatch; q++) {
        if (m->patchType[q] != 2) continue;   /* processor */
        for (label i = 0; i < m->patchSize[q]; i++) {
            label faceI = m->patchStart[q] + i;
            for (label j = m->faceOff[faceI]; j < m->faceOff[faceI + 1]; j++) isSplit[m->facePts[j]] = 0;
        }
    }
    free(bad);
}

/* refinement::faceConsistentUnrefinement (refinement.C:390-508) + edgeConsistentUnrefinement (:511-604),
   maxSet = false.  Returns nChanged, or -1 on the reference's "problem" FatalError. */
static int64_t poly_consistent_unrefinement_sweep(const or_world *w, int edgeBasedConsistency, boolean **unrefineCell) {
    int R = w->R;
    int64_t nChanged = 0;
    int fail = 0;
    label **neiLevel = (label **)malloc(sizeof(label *) * R);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        boolean *uc = unrefineCell[r];
        for (int64_t faceI = 0; faceI < m->f && !fail; faceI++) {
            label own = m->owner[faceI], nei = m->neighbour[faceI];
            label ownLevel = m->cellLevel[own] - uc[own], neiL = m->cellLevel[nei] - uc[nei];
            if (ownLevel < (neiL - 1)) { if (!uc[own]) fail = 1; else { uc[own] = 0; nChanged++; } }
            else if (neiL < (ownLevel - 1)) { if (!uc[nei]) fail = 1; else { uc[nei] = 0; nChanged++; } }
        }
        neiLevel[r] = (label *)malloc(sizeof(label) * (size_t)(m->b + 1));
        for (int64_t i = 0; i < m->b; i++) { label own = m->owner[i + m->f]; neiLevel[r][i] = m->cellLevel[own] - uc[own]; }
    }
    swap_boundary_label(w, neiLevel);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        boolean *uc = unrefineCell[r];
        for (int64_t i = 0; i < m->b && !fail; i++) {
            label own = m->owner[i + m->f];
            label ownLevel = m->cellLevel[own] - uc[own];
            if (ownLevel < (neiLevel[r][i] - 1)) { if (!uc[own]) fail = 1; else { uc[own] = 0; nChanged++; } }
        }
        free(neiLevel[r]);
        if (edgeBasedConsistency)
            for (int64_t edgeI = 0; edgeI < m->nedge && !fail; edgeI++)
                for (label i = m->ecOff[edgeI]; i < m->ecOff[edgeI + 1] && !fail; i++) {
                    label cellI = m->ecCells[i];
                    for (label j = i + 1; j < m->ecOff[edgeI + 1]; j++) {
                        label cellJ = m->ecCells[j];
                        label cellILevel = m->cellLevel[cellI] - uc[cellI], cellJLevel = m->cellLevel[cellJ] - uc[cellJ];
                        if (cellILevel < (cellJLevel - 1)) { if (!uc[cellI]) { fail = 1; break; } uc[cellI] = 0; nChanged++; }
                        else if (cellJLevel < (cellILevel - 1)) { if (!uc[cellJ]) { fail = 1; break; } uc[cellJ] = 0; nChanged++; }
                    }
                }
    }
    free(neiLevel);
    return fail ? -1 : nChanged;
}

/* Unrefinement branch of fvMeshPolyRefiner::refine (fvMeshPolyRefiner.C:419-472): layers x
   extendMarkedCellsAcrossPoints, set protectedPatches, selectUnrefinePoints (:48-96) ->
   polyhedralRefinement::consistentUnrefinement (polyhedralRefinement.C:2077-2362), or with prismatic2D != 0
   prismatic2DRefinement::consistentUnrefinement (prismatic2DRefinement.C:2729-2972; same loop, other
   split-point rule). */
OR_API int or_poly_select_unrefine(const or_world *w, const scalar *const *error, scalar unrefineLevel,
                                   int nUnrefinementBufferLayers, const label *protectedPatchIds, int nProtectedPatches,
                                   int edgeBasedConsistency, const label *const *pointGlobal, int64_t nGlobalPoints,
                                   boolean **refineCell, boolean **unrefPoint, boolean **splitPoint, int prismatic2D) {
    int R = w->R;
    for (int i = 0; i < nUnrefinementBufferLayers; i++) extend_across_points(w, refineCell, pointGlobal, nGlobalPoints);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        for (int k = 0; k < nProtectedPatches; k++) {
            int q = protectedPatchIds[k];
            for (label i = 0; i < m->patchSize[q]; i++) refineCell[r][m->owner[m->patchStart[q] + i]] = 1;
        }
    }
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        scalar *pFld = (scalar *)malloc(sizeof(scalar) * (size_t)(m->p + 1));
        boolean *isSplit = (boolean *)malloc((size_t)(m->p + 1));
        or_max_cell_field(w, r, error[r], pFld);
        if (prismatic2D) or_prism_split_points(w, r, isSplit);
        else or_poly_split_points(w, r, isSplit);
        for (int64_t pointI = 0; pointI < m->p; pointI++) {
            int cand = 0;
            if (pFld[pointI] < unrefineLevel) {                       /* selectUnrefinePoints */
                int hasMarked = 0;
                for (label j = m->pcOff[pointI]; j < m->pcOff[pointI + 1]; j++)
                    if (refineCell[r][m->pcCells[j]]) { hasMarked = 1; break; }
                cand = !hasMarked;
            }
            unrefPoint[r][pointI] = cand && isSplit[pointI];          /* PART 2 */
        }
        if (splitPoint && splitPoint[r]) memcpy(splitPoint[r], isSplit, (size_t)m->p);
        free(pFld); free(isSplit);
    }
    /* PART 4 */
    boolean **uc = (boolean **)malloc(sizeof(boolean *) * R);
    for (int r = 0; r < R; r++) uc[r] = (boolean *)malloc((size_t)(w->m[r].n + 1));
    int sweeps = 0, fail = 0;
    while (1) {
        for (int r = 0; r < R; r++) {
            const or_mesh *m = &w->m[r];
            memset(uc[r], 0, (size_t)m->n);
            for (int64_t pointI = 0; pointI < m->p; pointI++)
                if (unrefPoint[r][pointI])
                    for (label j = m->pcOff[pointI]; j < m->pcOff[pointI + 1]; j++) uc[r][m->pcCells[j]] = 1;
        }
        int64_t nChanged = poly_consistent_unrefinement_sweep(w, edgeBasedConsistency, uc);
        if (nChanged < 0) { fail = 1; break; }
        sweeps++;
        if (nChanged == 0) break;
        for (int r = 0; r < R; r++) {
            const or_mesh *m = &w->m[r];
            for (int64_t pointI = 0; pointI < m->p; pointI++)
                if (unrefPoint[r][pointI])
                    for (label j = m->pcOff[pointI]; j < m->pcOff[pointI + 1]; j++)
                        if (!uc[r][m->pcCells[j]]) { unrefPoint[r][pointI] = 0; break; }
        }
    }
    for (int r = 0; r < R; r++) free(uc[r]);
    free(uc);
    return fail ? -1 : sweeps;
}

/* hexRef::checkRefinementLevels (2:1 part), hexRef.C:2934-3025.  Returns #violating faces. */
OR_API int64_t or_check_levels(const or_world *w, const label *const *cellLevel) {
    int R = w->R;
    int64_t bad = 0;
    label **neiLevel = (label **)malloc(sizeof(label *) * R);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        for (int64_t facei = 0; facei < m->f; facei++)
            if (abs(cellLevel[r][m->owner[facei]] - cellLevel[r][m->neighbour[facei]]) > 1) bad++;
        neiLevel[r] = (label *)malloc(sizeof(label) * (size_t)(m->b + 1));
        for (int64_t i = 0; i < m->b; i++) neiLevel[r][i] = cellLevel[r][m->owner[i + m->f]];
    }
    swap_boundary_label(w, neiLevel);
    for (int r = 0; r < R; r++) {
        const or_mesh *m = &w->m[r];
        for (int64_t i = 0; i < m->b; i++)
            if (abs(cellLevel[r][m->owner[i + m->f]] - neiLevel[r][i]) > 1) bad++;
        free(neiLevel[r]);
    }
    free(neiLevel);
    return bad;
}

/* ======================================================================================= */
/* STAGE 4: field mapping and level update                                                 */
/* ======================================================================================= */

/* fvMesh::mapFields with direct addressing (reached from adaptiveFvMesh.C:293,346-348):
   new[i] = old[cellMap[i]] for each of nComp interleaved components */
OR_API void or_map_field(int64_t nNew, const label *cellMap, int nComp, const scalar *oldF, scalar *newF) {
    for (int64_t i = 0; i < nNew; i++)
        for (int k = 0; k < nComp; k++) newF[i * nComp + k] = oldF[(int64_t)cellMap[i] * nComp + k];
}

/* fvMesh::mapFields on every rank of the world */
OR_API void or_map_field_world(const or_world *w, const int64_t *nNew, const label *const *cellMap, int nComp,
                               const scalar *const *oldF, scalar **newF) {
    PAR_RANKS
    for (int r = 0; r < w->R; r++) or_map_field(nNew[r], cellMap[r], nComp, oldF[r], newF[r]);
}

/* Optional conservative (volume-weighted) merge, the north-star's "conservative averaging":
   new[m] = sum V_old[c] old[c] / sum V_old[c] over the master and every old cell merged into it
   (reverseCellMap[c] == -(m)-2); accumulation in ascending old-cell order.  No in-tree reference
   (parity unpinned); refinement children still take the parent value. */
/*
 * Old-time volumes V0 / V00 through a topology change: the part of fvMesh::updateMesh (OpenFOAM core, reached from
 * adaptiveFvMesh.C:293; restated from upstream fvMesh.C "Map the old volume. Just map to new cell labels." /
 * "Inject volume of merged cells" -- not in the reference tree, parity unpinned):
 *   V0new[i] = cellMap[i] > -1 ? V0old[cellMap[i]] : 0
 *   for oldCelli ascending: index = reverseCellMap[oldCelli]; if (index < -1) V0new[-index-2] += V0old[oldCelli]
 * so the master of a merge carries the volume of all the cells merged into it.
 */
OR_API int64_t or_map_old_volumes(int64_t nOld, int64_t nNew, const label *cellMap, const label *reverseCellMap,
                                  const scalar *V0old, scalar *V0new) {
    for (int64_t i = 0; i < nNew; i++) V0new[i] = cellMap[i] > -1 ? V0old[cellMap[i]] : 0.0;
    int64_t nMerged = 0;
    for (int64_t oldCelli = 0; oldCelli < nOld; oldCelli++) {
        label index = reverseCellMap[oldCelli];
        if (index < -1) { V0new[-index - 2] += V0old[oldCelli]; nMerged++; }
    }
    return nMerged;
}

OR_API void or_map_field_conservative(int64_t nOld, int64_t nNew, const label *cellMap,
                                      const label *reverseCellMap, const scalar *volOld, int nComp,
                                      const scalar *oldF, scalar *newF) {
    scalar *vs = (scalar *)calloc((size_t)(nNew + 1), sizeof(scalar));
    int hasMerge = 0;
    for (int64_t c = 0; c < nOld; c++) if (reverseCellMap[c] < -1) { hasMerge = 1; break; }
    or_map_field(nNew, cellMap, nComp, oldF, newF);
    if (!hasMerge) { free(vs); return; }
    boolean *merged = (boolean *)calloc((size_t)(nNew + 1), 1);
    for (int64_t c = 0; c < nOld; c++) if (reverseCellMap[c] < -1) merged[-reverseCellMap[c] - 2] = 1;
    for (int64_t i = 0; i < nNew; i++) if (merged[i]) for (int k = 0; k < nComp; k++) newF[i * nComp + k] = 0.0;
    for (int64_t c = 0; c < nOld; c++) {
        int64_t t = reverseCellMap[c] < -1 ? -reverseCellMap[c] - 2 : reverseCellMap[c];
        if (t < 0 || !merged[t]) continue;
        vs[t] += volOld[c];
        for (int k = 0; k < nComp; k++) newF[t * nComp + k] += volOld[c] * oldF[c * nComp + k];
    }
    for (int64_t i = 0; i < nNew; i++) if (merged[i]) for (int k = 0; k < nComp; k++) newF[i * nComp + k] /= vs[i];
    free(vs); free(merged);
}

/* Level update: refine: cell and its 7 added cells get old+1
   (hexRef3D.C:1238,1255) == level of every new cell whose cellMap target was refined;
   unrefine: cellLevel[pCells]-- (hexRef3D.C:2132-2135) then remap through cellMap. */
OR_API void or_update_levels(int64_t nNew, const label *cellMap, const label *oldLevel,
                             const boolean *refinedOld, const boolean *unrefinedOld, label *newLevel) {
    for (int64_t i = 0; i < nNew; i++) {
        label o = cellMap[i];
        newLevel[i] = oldLevel[o] + (refinedOld && refinedOld[o] ? 1 : 0) - (unrefinedOld && unrefinedOld[o] ? 1 : 0);
    }
}

/* ======================================================================================= */
/* NEXT ROWS: remap after a topology change (N1), flux fix-up (N2), gradient   */
/* estimators, mesh size, multicomponent (N3).  Same status as the rest: restatements,     */
/* parity unpinned (no reference test pins any of these).                                  */
/* ======================================================================================= */

/* ListOps reorder(oldToNew, size, -1, list) as called by hexRef::updateMesh for cellLevel_ / pointLevel_
   (hexRef.C:2463, :2529): entries with a negative target (merged away) are dropped, unfilled slots are -1 */
OR_API void or_reorder_labels(int64_t nOld, int64_t nNew, const label *oldToNew, const label *oldV, label *newV) {
    for (int64_t i = 0; i < nNew; i++) newV[i] = -1;
    for (int64_t i = 0; i < nOld; i++) { label t = oldToNew[i]; if (t >= 0) newV[t] = oldV[i]; }
}
/* the other branch of hexRef::updateMesh (hexRef.C:2467-2484, :2533-2556): new = old[map], -1 for inflated */
OR_API void or_map_labels(int64_t nNew, const label *map, const label *oldV, label *newV) {
    for (int64_t i = 0; i < nNew; i++) { label o = map[i]; newV[i] = (o == -1) ? -1 : oldV[o]; }
}
/* protectedCell_ renumbering, fvMeshHexRefiner.C:229-240 (refine) and :283-297 (unrefine: oldCelli >= 0 only) */
OR_API void or_map_protected(int64_t nNew, const label *cellMap, const boolean *oldP, boolean *newP) {
    for (int64_t i = 0; i < nNew; i++) { label o = cellMap[i]; newP[i] = (o >= 0) ? oldP[o] : 0; }
}

/* coupled flags over boundary faces */
static unsigned char *coupled_faces(const or_mesh *m) {
    unsigned char *coupled = (unsigned char *)calloc((size_t)(m->b + 1), 1);
    for (int q = 0; q < m->npatch; q++) if (m->patchType[q] == 2)
        for (label i = 0; i < m->patchSize[q]; i++) coupled[m->patchStart[q] - m->f + i] = 1;
    return coupled;
}

/* Flux fix-up of adaptiveFvMesh::updateMesh (adaptiveFvMesh.C:122-291) for one flux field of one rank, on the
   NEW mesh (geometry = new mesh).  phiU = fvc::interpolate(U) & Sf with linear interpolation;
   faces rewritten: inflated/appended (faceMap == -1), face-from-master (reverseFaceMap[old] != face), and
   the master faces themselves, except on empty patches (their patch fields have no entries).
   U [3n]; Ub [3b] boundary values (NULL: patch-internal value); Un [3b] patchNeighbourField on coupled
   faces.  Returns the number of rewritten faces. */
OR_API int64_t or_correct_fluxes(const or_world *w, int r, int64_t nOldFaces, const label *faceMap,
                                 const label *reverseFaceMap, const scalar *U, const scalar *Ub, const scalar *Un,
                                 scalar *phi) {
    const or_mesh *m = &w->m[r];
    int64_t f = m->f, nf = m->f + m->b, nDone = 0;
    (void)nOldFaces;
    unsigned char *coupled = coupled_faces(m);
    unsigned char *empty = (unsigned char *)calloc((size_t)(m->b + 1), 1);
    for (int q = 0; q < m->npatch; q++) if (m->patchType[q] == 1)
        for (label i = 0; i < m->patchSize[q]; i++) empty[m->patchStart[q] - f + i] = 1;
    unsigned char *fix = (unsigned char *)calloc((size_t)(nf + 1), 1);
    for (int64_t facei = 0; facei < nf; facei++) {
        label oldFacei = faceMap[facei];
        if (oldFacei >= 0) {
            label masterFacei = reverseFaceMap[oldFacei];
            if (masterFacei < 0) { free(coupled); free(empty); free(fix); return -1; }   /* :139-145 FatalError */
            if (masterFacei != facei) { fix[masterFacei] = 1; fix[facei] = 1; }          /* :146-149, :222-226 */
        } else if (oldFacei == -1) fix[facei] = 1;                                         /* :217-221 */
    }
    for (int64_t facei = 0; facei < nf; facei++) {
        if (!fix[facei]) continue;
        scalar Uf[3];
        if (facei < f) {
            label P = m->owner[facei], N = m->neighbour[facei];
            for (int k = 0; k < 3; k++) Uf[k] = m->weights[facei] * (U[3 * P + k] - U[3 * N + k]) + U[3 * N + k];
        } else {
            int64_t i = facei - f;
            if (empty[i]) continue;
            label P = m->owner[facei];
            for (int k = 0; k < 3; k++) {
                if (coupled[i]) Uf[k] = m->weights[facei] * U[3 * P + k] + (1.0 - m->weights[facei]) * Un[3 * i + k];
                else Uf[k] = Ub ? Ub[3 * i + k] : U[3 * P + k];
            }
        }
        phi[facei] = Uf[0] * m->Sf[3 * facei] + Uf[1] * m->Sf[3 * facei + 1] + Uf[2] * m->Sf[3 * facei + 2];
        nDone++;
    }
    free(coupled); free(empty); free(fix);
    return nDone;
}

/* Cell gradient of a scalar field for the gradient-based selectors (gradientRange.C:100, the coded
   selector of tutorials/poly_freelyPropFlame2D_H2/constant/dynamicMeshDict:35-40).
   scheme 0: Gauss linear  (gaussGrad.C gradf, as in or_estimate_lohner)
   scheme 1: leastSquares  (leastSquaresGrad.C calcGrad) with the geometric vectors of
             leastSquaresVectors::New(mesh) supplied by the caller: lsP [3(f+b)] = pVectors incl. boundary
             faces, lsN [3f] = nVectors.
   Xn = patchNeighbourField over boundary faces (caller swaps), xb = boundary values (NULL: zeroGradient).
   Output: magGrad[c] = mag(grad[c]) (* V[c]/tV when tV > 0), optionally grad [3n]. */
OR_API void or_gradient_mag(const or_world *w, int r, int scheme, const scalar *x, const scalar *xb, const scalar *xn,
                            const scalar *lsP, const scalar *lsN, scalar tV, scalar *grad, scalar *magGrad) {
    const or_mesh *m = &w->m[r];
    int64_t n = m->n, f = m->f, b = m->b;
    unsigned char *coupled = coupled_faces(m);
    scalar *g = (scalar *)calloc((size_t)(3 * n + 3), sizeof(scalar));
    if (scheme == 0) {
        for (int64_t facei = 0; facei < f; facei++) {
            label P = m->owner[facei], N = m->neighbour[facei];
            scalar ssf = m->weights[facei] * (x[P] - x[N]) + x[N];
            for (int k = 0; k < 3; k++) {
                scalar Sfssf = m->Sf[3 * facei + k] * ssf;
                g[3 * P + k] += Sfssf;
                g[3 * N + k] -= Sfssf;
            }
        }
        for (int64_t i = 0; i < b; i++) {
            int64_t facei = f + i;
            label P = m->owner[facei];
            scalar pssf;
            if (coupled[i]) pssf = m->weights[facei] * x[P] + (1.0 - m->weights[facei]) * xn[i];
            else pssf = xb ? xb[i] : x[P];
            for (int k = 0; k < 3; k++) g[3 * P + k] += m->Sf[3 * facei + k] * pssf;
        }
        for (int64_t c = 0; c < n; c++) for (int k = 0; k < 3; k++) g[3 * c + k] /= m->V[c];
    } else {
        for (int64_t facei = 0; facei < f; facei++) {
            label ownFacei = m->owner[facei], neiFacei = m->neighbour[facei];
            scalar deltaVsf = x[neiFacei] - x[ownFacei];
            for (int k = 0; k < 3; k++) {
                g[3 * ownFacei + k] += lsP[3 * facei + k] * deltaVsf;
                g[3 * neiFacei + k] -= lsN[3 * facei + k] * deltaVsf;
            }
        }
        for (int64_t i = 0; i < b; i++) {
            int64_t facei = f + i;
            label P = m->owner[facei];
            scalar other = coupled[i] ? xn[i] : (xb ? xb[i] : x[P]);
            scalar d = other - x[P];
            for (int k = 0; k < 3; k++) g[3 * P + k] += lsP[3 * facei + k] * d;
        }
    }
    for (int64_t c = 0; c < n; c++) {
        scalar v = sqrt(g[3 * c] * g[3 * c] + g[3 * c + 1] * g[3 * c + 1] + g[3 * c + 2] * g[3 * c + 2]);
        if (tV > 0) v *= m->V[c] / tV;
        magGrad[c] = v;
    }
    if (grad) memcpy(grad, g, sizeof(scalar) * (size_t)(3 * n));
    free(coupled); free(g);
}

/* meshSizeObject::calcDx (meshSizeObject.C:66-131).  geoD = mesh.geometricD() (+1 valid, -1 empty).
   The reference's bookkeeping is kept as written: an internal face adds its area to BOTH cells but
   counts twice on the owner and never on the neighbour (:97-104). */
OR_API void or_mesh_dx(const or_world *w, int r, const label *geoD, scalar *dx) {
    const or_mesh *m = &w->m[r];
    int64_t n = m->n, f = m->f, nf = m->f + m->b;
    int nGeo = 0;
    for (int k = 0; k < 3; k++) if (geoD[k] > 0) nGeo++;
    if (nGeo == 3) { for (int64_t c = 0; c < n; c++) dx[c] = cbrt(m->V[c]); return; }
    scalar validD[3];
    for (int k = 0; k < 3; k++) validD[k] = geoD[k] < 0 ? 1.0 : 0.0;
    label *nFaces = (label *)calloc((size_t)(n + 1), sizeof(label));
    for (int64_t c = 0; c < n; c++) dx[c] = 0.0;
    for (int64_t facei = 0; facei < nf; facei++) {
        const scalar *S = m->Sf + 3 * facei;
        scalar magSf = sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]);
        scalar dot = (S[0] / magSf) * validD[0] + (S[1] / magSf) * validD[1] + (S[2] / magSf) * validD[2];
        if (!(smag(dot) > 0.5)) continue;
        if (facei < f) {
            dx[m->owner[facei]] += magSf;
            dx[m->neighbour[facei]] += magSf;
            nFaces[m->owner[facei]]++;
            nFaces[m->owner[facei]]++;
        } else {
            dx[m->owner[facei]] += magSf;
            nFaces[m->owner[facei]]++;
        }
    }
    for (int64_t c = 0; c < n; c++) dx[c] = sqrt(dx[c] / (scalar)nFaces[c]);
    free(nFaces);
}

/* meshSizeObject::calcDX (meshSizeObject.C:134-160): dX = 2V / sum over cells()[c] of cmptMag(Sf), the sum
   in primitiveMesh::cells() order */
OR_API void or_mesh_dX(const or_world *w, int r, scalar *dX) {
    const or_mesh *m = &w->m[r];
    for (int64_t c = 0; c < m->n; c++) {
        scalar s[3] = {0, 0, 0};
        for (label j = m->cfOff[c]; j < m->cfOff[c + 1]; j++)
            for (int k = 0; k < 3; k++) s[k] += smag(m->Sf[3 * (int64_t)m->cfFaces[j] + k]);
        for (int k = 0; k < 3; k++) dX[3 * c + k] = (1.0 * (2.0 * m->V[c])) / s[k];
    }
}

/* errorEstimator::maxRefinement (errorEstimator.C:262-292) */
OR_API void or_max_refinement(const or_world *w, int r, label maxLevel, scalar minDx, const scalar *dx,
                              const scalar *error, label *out) {
    const or_mesh *m = &w->m[r];
    if (maxLevel >= 0 || !m->cellLevel) { for (int64_t c = 0; c < m->n; c++) out[c] = maxLevel; return; }
    for (int64_t c = 0; c < m->n; c++) {
        out[c] = m->cellLevel[c];
        if (dx[c] > minDx && error[c] > 0) out[c]++;
    }
}

/* errorEstimators::cellSize::update (cellSize.C:133-210).  sizeType 0 VOLUME, 1 CHARACTERISTIC, 2 MAG, 3 CMPT.
   CMPT accumulates into the error field it finds (the reference never resets it). */
OR_API void or_cell_size_error(const or_world *w, int r, int sizeType, const label *geoD, scalar lowerUnrefine,
                               scalar lowerRefine, const label *cmpts, int nCmpts, const scalar *minDX,
                               const scalar *maxDX, const scalar *dx, const scalar *dX, scalar *error) {
    const or_mesh *m = &w->m[r];
    int64_t n = m->n;
    if (sizeType == 0) for (int64_t c = 0; c < n; c++) error[c] = m->V[c];
    else if (sizeType == 1) for (int64_t c = 0; c < n; c++) error[c] = dx[c];
    else if (sizeType == 2) {
        for (int64_t c = 0; c < n; c++) error[c] = 0.0;
        for (int k = 0; k < 3; k++) if (geoD[k] == 1)
            for (int64_t c = 0; c < n; c++) error[c] += dX[3 * c + k] * dX[3 * c + k];
        for (int64_t c = 0; c < n; c++) error[c] = sqrt(error[c]);
    } else {
        for (int i = 0; i < nCmpts; i++) {
            label k = cmpts[i];
            for (int64_t c = 0; c < n; c++) {
                if (dX[3 * c + k] > maxDX[k]) error[c] = smax(1.0, error[c]);
                else if (dX[3 * c + k] < minDX[k]) error[c] = smax(-1.0, error[c]);
                else error[c] = smax(0, error[c]);
            }
        }
        return;
    }
    for (int64_t c = 0; c < n; c++) {
        if (error[c] < lowerUnrefine) error[c] = -1.0;
        else if (error[c] > lowerRefine) error[c] = 1.0;
        else error[c] = 0.0;
    }
}

/* errorEstimators::multicomponent::update (multicomponent.C:95-119): error = max over the sub-estimators, from -1 */
OR_API void or_multicomponent_error(int64_t n, int nEst, const scalar *const *errors, scalar *error) {
    for (int64_t c = 0; c < n; c++) error[c] = -1.0;
    for (int i = 0; i < nEst; i++)
        for (int64_t c = 0; c < n; c++) error[c] = smax(error[c], errors[i][c]);
}

/* errorEstimators::multicomponent::maxRefinement (multicomponent.C:122-157): in-place, ascending cells; a cell
   that fires also overwrites its cellCells -- order dependent, kept as written */
OR_API void or_multicomponent_max_refinement(const or_world *w, int r, int nEst, const scalar *const *errors,
                                             const label *const *maxCellLevels, label *maxRefinement) {
    const or_mesh *m = &w->m[r];
    int64_t n = m->n, f = m->f;
    for (int64_t c = 0; c < n; c++) maxRefinement[c] = 0;
    for (int i = 0; i < nEst; i++) {
        const scalar *errori = errors[i];
        const label *maxCellLevel = maxCellLevels[i];
        for (int64_t celli = 0; celli < n; celli++) {
            if (errori[celli] > 0.5 && maxRefinement[celli] < maxCellLevel[celli]) {
                maxRefinement[celli] = maxCellLevel[celli];
                for (label j = m->cfOff[celli]; j < m->cfOff[celli + 1]; j++) {       /* cellCells via cells() */
                    label facei = m->cfFaces[j];
                    if (facei >= f) continue;
                    label cellj = (m->owner[facei] == celli) ? m->neighbour[facei] : m->owner[facei];
                    maxRefinement[cellj] = maxCellLevel[celli];
                }
            }
        }
    }
}

/* ---- N4: inputs of the load-balance decision ------------------------------------------------------------ */

/* hexRefRefinementHistory::markCommonCells (hexRefRefinementHistory.C:398-444): every cell whose history entry
   exists gets the cluster of its top-most ancestor; clusters are numbered in order of first appearance over
   ascending cells.  (mark() paints the whole tree below the root; only the entries of visible cells are read back,
   and each of those reaches the root through parent_.)  Returns the number of clusters. */
OR_API label or_history_clusters(const or_world *w, int r, int64_t nSplit, label *cellToCluster) {
    const or_mesh *m = &w->m[r];
    label clusterI = 0;
    label *rootToCluster = (label *)malloc(sizeof(label) * (size_t)(nSplit + 1));
    for (int64_t i = 0; i < nSplit; i++) rootToCluster[i] = -1;
    for (int64_t cellI = 0; cellI < m->n; cellI++) {
        label index = m->visible[cellI];
        cellToCluster[cellI] = -1;
        if (index >= 0) {
            while (m->splitParent[index] != -1) index = m->splitParent[index];
            if (rootToCluster[index] == -1) rootToCluster[index] = clusterI++;
            cellToCluster[cellI] = rootToCluster[index];
        }
    }
    free(rootToCluster);
    return clusterI;
}

/* hexRefRefinementHistory::add (:448-487) / refinement::add (refinement.C:1007-1046): faces between two cells of one
   cluster are unblocked, everything else (boundary faces included) stays blocked; the syncFaceList(andEqOp) that
   follows cannot change a boundary face (both sides hold true). */
OR_API int64_t or_blocked_faces(const or_world *w, int r, const label *cellToCluster, boolean *blockedFace) {
    const or_mesh *m = &w->m[r];
    int64_t nUnblocked = 0;
    for (int64_t faceI = 0; faceI < m->f + m->b; faceI++) blockedFace[faceI] = 1;
    for (int64_t faceI = 0; faceI < m->f; faceI++) {
        label ownCluster = cellToCluster[m->owner[faceI]], neiCluster = cellToCluster[m->neighbour[faceI]];
        if (ownCluster != -1 && ownCluster == neiCluster) { blockedFace[faceI] = 0; nUnblocked++; }
    }
    return nUnblocked;
}

/* hexRefRefinementHistory::apply (:490-545): every cluster goes to the processor its first in-cluster face's owner
   was given.  Returns nChanged. */
OR_API int64_t or_constrain_decomposition(const or_world *w, int r, const label *cellToCluster, label nClusters,
                                          label *decomposition) {
    const or_mesh *m = &w->m[r];
    label *clusterToProc = (label *)malloc(sizeof(label) * (size_t)(nClusters + 1));
    for (label i = 0; i < nClusters; i++) clusterToProc[i] = -1;
    int64_t nChanged = 0;
    for (int64_t faceI = 0; faceI < m->f; faceI++) {
        label own = m->owner[faceI], nei = m->neighbour[faceI];
        label ownCluster = cellToCluster[own], neiCluster = cellToCluster[nei];
        if (ownCluster != -1 && ownCluster == neiCluster) {
            if (clusterToProc[ownCluster] == -1) clusterToProc[ownCluster] = decomposition[own];
            if (decomposition[own] != clusterToProc[ownCluster]) { decomposition[own] = clusterToProc[ownCluster]; nChanged++; }
            if (decomposition[nei] != clusterToProc[ownCluster]) { decomposition[nei] = clusterToProc[ownCluster]; nChanged++; }
        }
    }
    free(clusterToProc);
    return nChanged;
}

/* fvMeshBalance::canBalance (fvMeshBalance.C:462-468): maxImbalance / idealNCells over the world */
OR_API scalar or_imbalance(const or_world *w, int64_t *nGlobalCells) {
    int64_t tot = 0;
    for (int r = 0; r < w->R; r++) tot += w->m[r].n;
    scalar ideal = (scalar)tot / (scalar)w->R, mx = 0;
    for (int r = 0; r < w->R; r++) mx = smax(mx, smag((scalar)w->m[r].n - ideal));
    if (nGlobalCells) *nGlobalCells = tot;
    return mx / ideal;
}

/* procLoadNew (fvMeshBalance.C:497-503) of one rank; the caller sums over ranks */
OR_API void or_proc_load(int64_t n, const label *distribution, int nProcs, int64_t *procLoad) {
    for (int q = 0; q < nProcs; q++) procLoad[q] = 0;
    for (int64_t c = 0; c < n; c++) procLoad[distribution[c]]++;
}
