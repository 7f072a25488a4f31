/*
 * amr_b200.h -- C-ABI of the B200-native AMR decision path (libamr_b200.so).
 *
 * This is the drop-in boundary for blastAMR's per-timestep decision path
 * (estimate -> mark -> 2:1 balance -> map).  Every entry point replaces the body of one
 * reference method; the host (adaptiveFvMesh / errorEstimator / fvMeshRefiner / hexRef in the
 * reference, blastamr_b200/host in this repo) keeps its run-time-selectable names and
 * dynamicMeshDict keywords and calls in here.  Citations are relative to the reference tree.
 *
 * Conventions (OpenFOAM, WM_LABEL_SIZE=32, WM_PRECISION_OPTION=DP):
 *   label  = int32_t, scalar = double, boolList element = uint8_t (0/1).
 *   n cells, f internal faces, b boundary faces, p points.
 *   owner[f+b] = mesh.faceOwner(), neighbour[f] = mesh.faceNeighbour(); internal faces first,
 *   upper-triangular order; boundary faces grouped by patch (polyMesh/boundary startFace,nFaces).
 *   Arrays "over boundary faces" have length b and are indexed by (face - f), the layout of
 *   `labelList neiLevel(nFaces-nInternalFaces)` in hexRef.C:745.
 *
 * Memory: every data pointer may be a host pointer (the drop-in case; the library stages it
 * through the context's device buffers on the context's stream) or a device pointer on the
 * context's device (resident case; used as is, no copy).  The library detects which with
 * cudaPointerGetAttributes.  Scalar outputs (counts) are always host pointers and may be NULL.
 *
 * Errors: every call returns AMR_OK (0) or a negative code; amr_last_error() gives the text.
 * No C++ exception crosses the boundary.  The host wrapper turns a failure into
 * FatalErrorInFunction ... abort(FatalError), like the reference's own invariants
 * (hexRef3D.C:1811-1815).
 *
 * Threading: one context per MPI rank / GPU.  A call that was given any HOST data pointer is
 * synchronous (its results are complete on return, exactly like the reference method it replaces);
 * a call whose data pointers are all device pointers is stream-ordered on the context's stream
 * (amr_set_stream), the normal CUDA contract for device buffers.  Scalar outputs (counts) always
 * force completion.  With a
 * communicator attached, all ranks must enter the collective-bearing calls together (they
 * replace syncTools::* / reduce() calls that are already collective).
 *
 * There is no CPU fallback: without a CUDA device amr_ctx_create fails.
 *
 * Environment (read at amr_ctx_create / amr_comm_init; tuning and diagnosis only, results never change):
 *   AMR_B200_HALO=nccl            processor-patch swaps and reductions over NCCL instead of NVLink peer memory
 *   AMR_B200_PEER_TIMEOUT_S=s     bound of the device-side waits of the peer-memory transport (default 20 s); a
 *                                 timeout fails the call with AMR_ERR_STATE instead of hanging the GPU
 *   AMR_B200_NO_PIPELINE=1        direct kernels instead of the TMA-staged ones
 *   AMR_B200_PUSH_FRACTION=a/b    marked fraction below which a buffer layer / 2:1 sweep takes its sparse form
 *                                 (default 1/4; 0 = always dense)
 *   AMR_B200_PUSH_MIN_CELLS=n     mesh size below which the sparse forms are not tried (default 65536)
 *   AMR_B200_PULL_FRACTION=a/b    source fraction above which a buffer layer walks the rows of the UNMARKED cells only
 *                                 (default 1/2; 0 = never)
 *   AMR_B200_LINKED=0             closure loops across ranks as three launches per iteration (halo push, queue kernel,
 *                                 coupled-face fix-up) instead of one with push-on-change (same value on every rank)
 *   AMR_B200_POLY_QUEUE=0         polyRefiner: dense edge + face sweeps instead of the queue closure on the edge graph
 *   AMR_B200_PIPE_STAGES=2|3      ring depth of the TMA-staged kernels (default 3)
 *   AMR_B200_FACE_BATCH=1         Lohner / Gauss gradient: four faces of a cell at a time (bitwise identical, slower)
 *   AMR_B200_UPLOAD_LANES=n       host threads of the staged upload of pageable arrays
 *   AMR_B200_TRACE=prefix         CUDA event after every launch, time line written at amr_ctx_destroy
 *   AMR_B200_HOSTTIME=1           wall-clock marks of the mesh hand-over calls on stderr
 */
#ifndef AMR_B200_H
#define AMR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct amr_ctx amr_ctx;

enum {
    AMR_OK = 0,
    AMR_ERR_CUDA = -1,      /* CUDA runtime error (text in amr_last_error) */
    AMR_ERR_ARG = -2,       /* bad argument */
    AMR_ERR_STATE = -3,     /* call order: mesh / levels / points not set */
    AMR_ERR_NCCL = -4,      /* NCCL error */
    AMR_ERR_INVARIANT = -5  /* a reference FatalError condition ("problem", 2:1 violated ...) */
};

/* estimator kinds == run-time-selection names of the reference (errorEstimators/Make/files:6-14) */
enum {
    AMR_EST_DELTA = 0,        /* "delta"        delta.C:70-136 */
    AMR_EST_SCALED_DELTA = 1, /* "scaledDelta"  scaledDelta.C:71-141 */
    AMR_EST_LOHNER = 2,       /* "Lohner"       Lohner.C:71-159 (needs amr_mesh_set_geometry) */
    AMR_EST_FIELD_VALUE = 3,  /* "fieldValue"   fieldValue.C:68-82 */
    AMR_EST_CODED_DELTA = 4,  /* coded delta w/o clamp, tutorials/damBreak2D/constant/dynamicMeshDict:24-65 */
    AMR_EST_DENSITY_GRADIENT = 5 /* "densityGradient": named by the north star but ABSENT from the reference tree
                                    (0 grep hits; upstream blastFoam only).  Defined here as the scaled face
                                    delta of rho, |rho_o-rho_n| / max(min(rho_o,rho_n), minValue) -- i.e. the
                                    arithmetic of scaledDelta.C:78-136 applied to the density field.  No in-tree
                                    oracle: parity unpinned by construction. */
};

/* patch types for amr_mesh_set (only "coupled" matters on this path) */
enum { AMR_PATCH_PLAIN = 0, AMR_PATCH_EMPTY = 1, AMR_PATCH_PROCESSOR = 2 };

/* refiner kinds == `refiner` keyword (fvMeshRefinerNew.C:74-103) x hexRef::New dimension switch */
enum {
    AMR_REFINER_HEX3D = 0,   /* hexRefiner + hexRef3D */
    AMR_REFINER_HEX2D = 1,   /* hexRefiner + hexRef2D (split elements are edges; needs amr_mesh_set_edges) */
    AMR_REFINER_POLY3D = 2,  /* polyRefiner + polyhedralRefinement (needs amr_mesh_set_edges with ALL mesh edges and
                                amr_mesh_set_faces; across ranks also amr_mesh_set_patch_points for the point sync of the
                                unrefinement buffer layers, syncTools::syncPointList fvMeshRefiner.C:952-958) */
    AMR_REFINER_POLY2D = 3   /* polyRefiner + prismatic2DRefinement (fvMeshPolyRefiner.C:139-155): same marking and
                                consistency sweeps as POLY3D; split points from pointLevel over the point's edges and
                                only processor-patch points excluded (prismatic2DRefinement.C:2744-2849) */
};

/* field-map modes */
enum {
    AMR_MAP_DIRECT = 0,       /* fvMesh::mapFields direct addressing: new[i] = old[cellMap[i]] (parity mode) */
    AMR_MAP_CONSERVATIVE = 1  /* volume-weighted merge on unrefinement (extension; no in-tree reference) */
};

/* ---- context ---------------------------------------------------------------------------- */

/* Create a context on CUDA device `device` (one per MPI rank).  Fails (AMR_ERR_CUDA) when no
   CUDA device is usable: there is no CPU path. */
int amr_ctx_create(int device, amr_ctx **out);
int amr_ctx_destroy(amr_ctx *ctx);
const char *amr_last_error(const amr_ctx *ctx);

/* Use an existing cudaStream_t (e.g. torch's current stream) instead of the context's own. */
int amr_set_stream(amr_ctx *ctx, void *cuda_stream);

/* Multi-GPU: one rank per GPU, processor patches <-> NCCL send/recv, reduce() <-> allreduce.
   Replaces the Pstream layer under syncTools::swapBoundaryFaceList / syncFaceList / reduce
   (call sites: hexRef.C:756,1424; fvMeshRefiner.C:840,894; hexRef3D.C:1853,1891).
   id128 is an ncclUniqueId created on rank 0 by amr_comm_unique_id and broadcast by the host. */
int amr_comm_unique_id(void *id128);
int amr_comm_init(amr_ctx *ctx, int nranks, int rank, const void *id128);

/* ---- mesh / level state (re-sent after every topology change) --------------------------- */

/* primitiveMesh addressing.  patchType[i]==AMR_PATCH_PROCESSOR marks coupled patches;
   patchNbrRank[i] is the neighbour rank there.  The library derives its own cell->cell
   adjacency (the reference walks mesh.cells() + owner/neighbour: fvMeshRefiner.C:807-862). */
int amr_mesh_set(amr_ctx *ctx, int64_t n, int64_t f, int64_t b, int64_t p,
                 const int32_t *owner, const int32_t *neighbour,
                 int npatch, const int32_t *patchStart, const int32_t *patchSize,
                 const int32_t *patchType, const int32_t *patchNbrRank);

/* mesh.pointCells() as CSR (pcOff[p+1], pcCells[pcOff[p]]) and the flag "point is used by a
   boundary face" (what the loop hexRef3D.C:2279-2294 computes from mesh.faces()). */
int amr_mesh_set_points(amr_ctx *ctx, const int32_t *pcOff, const int32_t *pcCells,
                        const uint8_t *bndPoint);

/* N1, second half: incremental hand-over of the ADDRESSING after a topology change
   (hexRef::updateMesh hexRef.C:2422-2608, fvMeshHexRefiner::updateMesh fvMeshHexRefiner.C:229-240).  amr_mesh_set
   re-uploads the whole mesh (4.7 GB at 64 M cells); these two calls carry the adjacency the context already holds
   through the change.  The host sends the maps of mapPolyMesh and only the rows that changed:
     cellMap [nNew]          new cell -> the old cell it came from (mapPolyMesh::cellMap(); read for clean cells only)
     reverseCellMap [nOld]   old cell -> new label, < 0 = removed (mapPolyMesh::reverseCellMap()); NULL = labels kept,
                             cells only appended (hexRef3D::setRefinement)
     dirtyCells [nDirty]     ascending NEW labels of every cell whose set of face neighbours changed (cut or merged cells,
                             added cells, and the cells next to them), dirtyCellOff/dirtyCellNbrs = their rows
                             (internal-face neighbours, new labels; mesh.cellCells()), dirtyCellLevel = their
                             cellLevel, dirtyParentIndex = history.parentIndex() or -1 (NULL: no history)
     boundaryOwner [bNew]    faceOwner of the boundary faces (patch ranges as in amr_mesh_set)
   A cell that is NOT listed keeps the row of cellMap[cell] with every neighbour renumbered through reverseCellMap;
   a clean row that points at a removed cell, or a dirty list that is not ascending, fails with AMR_ERR_ARG.
   The owner/neighbour face lists are not rebuilt: entry points that need them (amr_mesh_set_geometry and what
   depends on it, the history-cluster queries) return AMR_ERR_STATE until the next amr_mesh_set.  The decision path
   (estimators over faces, marking, 2:1 sweeps, unrefinement, maps) runs on the carried adjacency.
   amr_mesh_update_points does the same for mesh.pointCells(), bndPoint and pointLevel_ (pointMap [pNew] = new point
   -> old point; dirty points = new points, the points of cut / merged cells); it must follow amr_mesh_update when
   the context held point addressing.  Collective like amr_mesh_set when a communicator is attached. */
int amr_mesh_update(amr_ctx *ctx, int64_t nNew, int64_t fNew, int64_t bNew, int64_t pNew, const int32_t *cellMap,
                    const int32_t *reverseCellMap, int64_t nDirtyCells, const int32_t *dirtyCells,
                    const int32_t *dirtyCellOff, const int32_t *dirtyCellNbrs, const int32_t *dirtyCellLevel,
                    const int32_t *dirtyParentIndex, const int32_t *boundaryOwner, int npatch,
                    const int32_t *patchStart, const int32_t *patchSize, const int32_t *patchType,
                    const int32_t *patchNbrRank);
int amr_mesh_update_points(amr_ctx *ctx, const int32_t *pointMap, int64_t nDirtyPoints, const int32_t *dirtyPoints,
                           const int32_t *dirtyPointOff, const int32_t *dirtyPointCells, const uint8_t *dirtyBndPoint,
                           const int32_t *dirtyPointLevel);

/* 2-D cutter (hexRef2D): split elements are mesh EDGES.  mesh.edges() end points [2*nEdges],
   mesh.edgeCells() as CSR, and the flag "edge belongs to a boundary face" (the loop over
   mesh_.faceEdges() at hexRef2D.C:2296-2311).  Only needed with AMR_REFINER_HEX2D. */
int amr_mesh_set_edges(amr_ctx *ctx, int64_t nEdges, const int32_t *edgePoints, const int32_t *ecOff,
                       const int32_t *ecCells, const uint8_t *bndEdge);

/* mesh.faces() as CSR over all faces (faceOff[f+b+1], facePts): polyhedralRefinement derives its split
   points from pointLevel and the faces (polyhedralRefinement.C:2110-2184).  Only with AMR_REFINER_POLY3D. */
int amr_mesh_set_faces(amr_ctx *ctx, const int32_t *faceOff, const int32_t *facePts);

/* options of the refiners that the reference reads from dynamicMeshDict */
enum {
    AMR_OPT_EDGE_BASED_CONSISTENCY = 0, /* `edgeBasedConsistency` (refinement.C:700-703), default 1 */
    AMR_OPT_POLY_FORCE = 1              /* fvMeshRefiner::force_ as passed by fvMeshPolyRefiner.C:338, default 0 */
};
/* Points of the processor patches, patch after patch (patchPointOff [npatch+1], zero-length ranges for other
   patches), each list in the order that pairs it with the neighbour's list: processorPolyPatch::meshPoints() against
   neighbPoints().  Needed only for point-based syncs across ranks (syncTools::syncPointList in
   fvMeshRefiner::extendMarkedCellsAcrossPoints, fvMeshRefiner.C:952-958: the polyRefiner's unrefinement buffer
   layers; the pointLevel check of hexRef.C:3027-3057).  With a communicator the call is COLLECTIVE (the ranks agree
   on the transport of the point swaps): every rank calls it, with empty lists when it has no processor patch. */
int amr_mesh_set_patch_points(amr_ctx *ctx, const int32_t *patchPointOff, const int32_t *patchPoints);

int amr_set_option(amr_ctx *ctx, int option, int64_t value);

/* hexRef::cellLevel_/pointLevel_ and the refinement history arrays visibleCells_ and
   splitCells_[i].parent_ (hexRefRefinementHistory.H:299-310).  pointLevel, visibleCells,
   splitParent may be NULL when no unrefinement is requested. */
int amr_levels_set(amr_ctx *ctx, const int32_t *cellLevel, const int32_t *pointLevel,
                   const int32_t *visibleCells, const int32_t *splitParent, int64_t nSplit);

/* fvMesh geometry needed by the Lohner estimator only (Lohner.C:78-82 -> cubic interpolation + Gauss
   gradient): surfaceInterpolation::weights() [f+b], Sf() [3(f+b)], magSf() [f+b],
   nonOrthDeltaCoeffs() [f+b], V() [n].  Boundary entries follow the boundary-face order. */
int amr_mesh_set_geometry(amr_ctx *ctx, const double *weights, const double *Sf, const double *magSf,
                          const double *nonOrthDeltaCoeffs, const double *V);

/* fvMeshHexRefiner::protectedCell_ (fvMeshHexRefiner.C:1014-1465 computes it on the host);
   NULL = empty list. */
int amr_protected_cells_set(amr_ctx *ctx, const uint8_t *protectedCell);

/* ---- stage 1: errorEstimator::update ---------------------------------------------------- */

/* errorEstimator::update(scale) for `kind`.
     x          cell values [n]; for non-scalar fields the host passes mag(field)
                (errorEstimatorTemplates.C:30-46)
     nbrValues  patchNeighbourField() values over boundary faces [b] (only coupled ranges are
                read) or NULL: then, with a communicator, the library swaps patchInternalField
                itself; without one there are no coupled faces.
     p0, p1     scaledDelta: minValue, offset; Lohner: epsilon, unused
     scale      0: raw error; 1: followed by normalize() with thr[4] =
                {lowerRefine, upperRefine, lowerUnrefine, upperUnrefine} (errorEstimator.C:204-228)
     errorOut   [n] or NULL to leave the result resident in the context (used by the later
                stages when they are passed error == NULL). */
int amr_estimate(amr_ctx *ctx, int kind, const double *x, const double *nbrValues,
                 double p0, double p1, int scale, const double *thr, double *errorOut);

/* Lohner::update (Lohner.C:71-159).  x [n]; xBoundary [b] = x.boundaryField() values on the physical
   patches (NULL = zeroGradient, the owner cell value); nbrValues [b] = patchNeighbourField() on coupled
   faces (NULL = swapped by the library / no coupled faces).  The OpenFOAM-core part (cubic, gaussGrad) is
   restated from upstream: parity unpinned (see oracle/amr_oracle.c:or_estimate_lohner). */
int amr_estimate_lohner(amr_ctx *ctx, const double *x, const double *xBoundary, const double *nbrValues,
                        double epsilon, int scale, const double *thr, double *errorOut);

/* errorEstimator::normalize on its own (errorEstimator.C:204-228).  errorInOut NULL = resident. */
int amr_normalize(amr_ctx *ctx, double lowerRefine, double upperRefine, double lowerUnrefine,
                  double upperUnrefine, double *errorInOut);

/* Threshold part of gradientEstimator::update (gradientRange.C:102-109) and of the coded selectors of
   the flame tutorials (tutorials/poly_freelyPropFlame2D_H2/constant/dynamicMeshDict:50-58):
     maxE = gMax(error), minE = gMin(error)           (all-reduced over ranks with a communicator)
     lowerRefine = minE + refineThreshold (maxE-minE), lowerUnrefine = minE + unrefineThreshold (maxE-minE),
     upperRefine = upperUnrefine = GREAT; then normalize() when scale != 0.
   errorInOut: raw error (e.g. mag(fvc::grad(T)) computed by the host) or NULL for the resident field.
   thrOut[4] (host, nullable) receives {lowerRefine, upperRefine, lowerUnrefine, upperUnrefine};
   minMaxOut[2] (host, nullable) receives {minE, maxE}. */
int amr_normalize_range(amr_ctx *ctx, double refineThreshold, double unrefineThreshold, int scale,
                        double *errorInOut, double *thrOut, double *minMaxOut);

/* errorEstimator::protectPatches (errorEstimator.C:329-365): seed = owners of the faces of the
   listed patches, grow nLayers face layers (:347), set error = 0 there.  nProtected (host,
   nullable) receives the local count (:355-361). */
int amr_protect_patches(amr_ctx *ctx, const int32_t *patchIds, int nPatchIds, int nLayers,
                        double *errorInOut, int64_t *nProtected);

/* ---- stages 2+3: refinement selection --------------------------------------------------- */

/* Refinement branch of fvMeshHexRefiner::refine (fvMeshHexRefiner.C:1497-1541):
   selectRefineCandidates(lo, hi) -> nBufferLayers x extendMarkedCells(force=true) -> clear
   protectedPatches/protectedCell -> selectRefineCells(maxCells) incl. calculateProtectedCells
   -> hexRef::consistentRefinement(maxSet=true).
     error            normalised error [n] or NULL (resident result of amr_estimate)
     maxCellLevel     per-cell maxRefinement [n] or NULL, then maxLevelUniform applies
     refineCellOut    [n] the marking after buffer layers/protections (the reference keeps it
                      for the unrefinement stage); NULL = keep resident only
     cellsOut, nOut   ascending labelList cellsToRefine (capacity n) and its size
     sweepsOut        number of 2:1 sweeps executed (diagnostic; differs from the reference's
                      Gauss-Seidel count by construction) */
int amr_select_refine(amr_ctx *ctx, const double *error, double lowerRefineLevel,
                      double upperRefineLevel, const int32_t *maxCellLevel, int32_t maxLevelUniform,
                      int nBufferLayers, const int32_t *protectedPatchIds, int nProtectedPatches,
                      int64_t maxCells, int refinerKind, uint8_t *refineCellOut,
                      int32_t *cellsOut, int64_t *nOut, int *sweepsOut);

/* hexRef::consistentRefinement(cellsToRefine, maxSet) on its own (hexRef.C:1402-1468). */
int amr_consistent_refinement(amr_ctx *ctx, const int32_t *cellsIn, int64_t nIn, int maxSet,
                              int32_t *cellsOut, int64_t *nOut, int *sweepsOut);

/* Update of refineCell through the topology change (fvMeshHexRefiner.C:1562-1586).
   cellMap [nNew], reverseCellMap [nOld], refineCell [nOld] or NULL = the resident marking of this
   context (then nOld must equal its size); the result (size nNew) replaces the resident marking
   (it survives the following amr_mesh_set of the new mesh) and is copied to newRefineCellOut
   when not NULL. */
int amr_remap_refine_cell(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap,
                          const int32_t *reverseCellMap, const uint8_t *refineCell,
                          uint8_t *newRefineCellOut);

/* Unrefinement branch (fvMeshHexRefiner.C:1591-1633 + hexRef3D::selectUnrefineElems
   hexRef3D.C:2324-2381): nBufferLayers x extendMarkedCellsAcrossFaces -> set
   protectedCell/protectedPatches -> maxCellField -> getSplitElems -> filter ->
   consistentUnrefinement.  refineCellInOut NULL = resident marking.  elemsOut (capacity p)
   receives the ascending split-point list. */
int amr_select_unrefine(amr_ctx *ctx, const double *error, double unrefineLevel,
                        uint8_t *refineCellInOut, int nBufferLayers,
                        const int32_t *protectedPatchIds, int nProtectedPatches, int refinerKind,
                        int32_t *elemsOut, int64_t *nOut, int *sweepsOut);

/* hexRef3D::getSplitElems (hexRef3D.C:2183-2321): ascending split points. */
int amr_get_split_elems(amr_ctx *ctx, int refinerKind, int32_t *elemsOut, int64_t *nOut);

/* fvMeshRefiner::maxCellField (fvMeshRefiner.C:110-125): pFld[p] = max over pointCells, init -GREAT */
int amr_max_cell_field(amr_ctx *ctx, const double *vFld, double *pFld);

/* hexRef::checkRefinementLevels 2:1 part (hexRef.C:2934-3025).  cellLevel NULL = resident.
   nBad (host) = number of violating faces; returns AMR_ERR_INVARIANT when > 0. */
int amr_check_levels(amr_ctx *ctx, const int32_t *cellLevel, int64_t *nBad);

/* "Check pointLevel is synchronized" of hexRef::checkRefinementLevels (hexRef.C:3027-3057): every rank that shares a
   processor-patch point must hold the same pointLevel.  pointLevel NULL = resident.  Needs amr_mesh_set_patch_points
   on decomposed meshes; collective.  nBad (host) = points of this rank that disagree with a neighbour; returns
   AMR_ERR_INVARIANT when any rank found one. */
int amr_check_point_levels(amr_ctx *ctx, const int32_t *pointLevel, int64_t *nBad);

/* ---- stage 4: field mapping ------------------------------------------------------------- */

/* fvMesh::mapFields, reached from adaptiveFvMesh.C:293,346-348: direct addressing through
   mapPolyMesh::cellMap().  nComp interleaved components per cell (scalar 1, vector 3,
   symmTensor 6, tensor 9).  mode AMR_MAP_CONSERVATIVE additionally needs reverseCellMap[nOld]
   and volOld[nOld]. */
int amr_map_field(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap, int nComp,
                  const double *oldF, double *newF, int mode, const int32_t *reverseCellMap,
                  const double *volOld);

/* All registered fields through one map (what fvMesh::mapFields does): nFields fields with nComp[i]
   interleaved components each, direct addressing.  Shares the cellMap upload; with host buffers field k+1
   is uploaded while field k is downloaded (both PCIe directions busy). */
int amr_map_fields(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap, int nFields,
                   const int *nComp, const double *const *oldF, double *const *newF);

/* fvMesh::mapFields (as amr_map_fields) AND the refiner's update of refineCell through the same map
   (fvMeshHexRefiner.C:1562-1586, as amr_remap_refine_cell) in ONE pass over cellMap: with device buffers the map is
   read from memory once for all fields and the marking.  reverseCellMap NULL: fields only.  refineCell [nOld] NULL =
   the resident marking of this context; the new marking (size nNew) becomes the resident one and is copied to
   newRefineCellOut when not NULL. */
int amr_map_fields_and_marking(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap, int nFields,
                               const int *nComp, const double *const *oldF, double *const *newF,
                               const int32_t *reverseCellMap, const uint8_t *refineCell, uint8_t *newRefineCellOut);

/* cellLevel through a refinement/unrefinement map (hexRef3D.C:1238,1255,2132-2135 +
   hexRef.C:2457-2485): new[i] = old[cellMap[i]] + refined[cellMap[i]] - unrefined[cellMap[i]].
   refinedOld / unrefinedOld: uint8 flags over old cells, nullable. */
int amr_update_levels(amr_ctx *ctx, int64_t nNew, const int32_t *cellMap, const int32_t *oldLevel,
                      const uint8_t *refinedOld, const uint8_t *unrefinedOld, int32_t *newLevel);

/* ---- diagnostics ------------------------------------------------------------------------ */

/* ---------------------------------------------------------------------------------------------------
   Rows either side of the decision path.  Same conventions as above.
   --------------------------------------------------------------------------------------------------- */

/* remap modes of amr_remap_labels */
enum {
    AMR_REMAP_REORDER = 0,   /* map = oldToNew[nOld] (reverseCellMap / reversePointMap): ListOps reorder(map, nNew, -1, list),
                                hexRef.C:2463, :2529 -- negative targets are dropped, unfilled slots are -1 */
    AMR_REMAP_GATHER = 1     /* map = newToOld[nNew] (cellMap / pointMap): new = old[map], -1 where map == -1,
                                hexRef.C:2467-2484, :2533-2556 */
};
/* hexRef::updateMesh renumbering of cellLevel_ / pointLevel_ after a topology change.  Independent of the mesh
   held by the context. */
int amr_remap_labels(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *map, int mode, const int32_t *oldValues,
                     int32_t *newValues);
/* protectedCell_ renumbering, fvMeshHexRefiner.C:229-240 (refine) and :283-297 (unrefine): new = old[cellMap] where
   cellMap >= 0, false elsewhere. */
int amr_remap_protected_cells(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap, const uint8_t *oldProtected,
                              uint8_t *newProtected);

/* Flux fix-up of adaptiveFvMesh::updateMesh (adaptiveFvMesh.C:122-291) for one flux field.  The context holds the NEW
   mesh and its geometry (amr_mesh_set + amr_mesh_set_geometry).  phi [f+b] (surface field flattened in face order;
   entries of empty patches are ignored) is rewritten with (linear interpolate(U) & Sf) on inflated faces
   (faceMap == -1), faces made from a master face (reverseFaceMap[faceMap[face]] != face) and the master faces
   themselves.  U [3n]; UBoundary [3b] boundary values of U (NULL: patch-internal value); nbrU [3b]
   patchNeighbourField on processor patches (NULL: the library swaps it).  AMR_ERR_INVARIANT = the reference's
   "should not have removed faces when refining".  nCorrected (nullable) = number of rewritten faces. */
int amr_correct_fluxes(amr_ctx *ctx, int64_t nOldFaces, const int32_t *faceMap, const int32_t *reverseFaceMap, const double *U,
                       const double *UBoundary, const double *nbrU, double *phi, int64_t *nCorrected);

/* Old-time volumes through a topology change: the V0 / V00 part of fvMesh::updateMesh (OpenFOAM core, reached from
   adaptiveFvMesh.C:293): newV[i] = oldV[cellMap[i]] (0 where cellMap == -1), then every merged-away old cell
   (reverseCellMap < -1) adds its old volume to its master -(reverseCellMap)-2, in ascending old label like the
   reference's loop (bit-identical sums).  nVolumes fields (V0, V00) share the map.  Restated from upstream
   fvMesh.C: parity unpinned.  nMerged (nullable) = number of merged-away cells. */
int amr_map_old_volumes(amr_ctx *ctx, int64_t nOld, int64_t nNew, const int32_t *cellMap, const int32_t *reverseCellMap,
                        int nVolumes, const double *const *oldV, double *const *newV, int64_t *nMerged);

/* gradient schemes of amr_estimate_gradient */
enum { AMR_GRAD_GAUSS_LINEAR = 0, AMR_GRAD_LEAST_SQUARES = 1 };
/* leastSquaresVectors::New(mesh).pVectors() [3(f+b)] (internal + boundary, face order) and nVectors() [3f]: purely
   geometric, computed by the host once per topology. */
int amr_mesh_set_ls_vectors(amr_ctx *ctx, const double *pVectors, const double *nVectors);
/* gradientEstimator::update (gradientRange.C:87-110) and the coded gradient selectors of the flame tutorials
   (tutorials/poly_freelyPropFlame2D_H2/constant/dynamicMeshDict:35-60) in one call:
   error = mag(fvc::grad(x)) [* V/totalVolume when totalVolume > 0], gMin/gMax over all ranks, thresholds
   min + threshold*(max - min) (upper ones GREAT), normalize when scale != 0.  Gauss linear restates
   gaussGrad.C (as Lohner does); leastSquares restates leastSquaresGrad.C on the host's vectors.
   xBoundary / nbrValues as in amr_estimate_lohner.  thrOut[4], gradOut [3n] nullable. */
int amr_estimate_gradient(amr_ctx *ctx, int scheme, const double *x, const double *xBoundary, const double *nbrValues,
                          double totalVolume, double refineThreshold, double unrefineThreshold, int scale, double *thrOut,
                          double *gradOut, double *errorOut);

/* meshSizeObject (meshSizeObject.C:66-160): dx [n] (cbrt(V) in 3-D -- CUDA's cbrt, may differ from glibc's in the
   last bit; the face-area form with the reference's face count otherwise) and dX [3n] = 2V / sum cmptMag(Sf).
   geometricD[3] = mesh.geometricD() (+1 / -1).  Either output nullable. */
int amr_mesh_size(amr_ctx *ctx, const int32_t *geometricD, double *dx, double *dX);
/* errorEstimator::maxRefinement (errorEstimator.C:262-292): maxLevel >= 0 -> uniform list, else
   cellLevel + (dx > minDx && error > 0).  error NULL = resident error. */
int amr_max_refinement(amr_ctx *ctx, int32_t maxLevel, double minDx, const double *dx, const double *error,
                       int32_t *maxRefinementOut);
/* cellSize estimator (cellSize.C:133-210) */
enum { AMR_SIZE_VOLUME = 0, AMR_SIZE_CHARACTERISTIC = 1, AMR_SIZE_MAG = 2, AMR_SIZE_CMPT = 3 };
int amr_estimate_cell_size(amr_ctx *ctx, int sizeType, const int32_t *geometricD, double lowerUnrefine, double lowerRefine,
                           const int32_t *cmpts, int nCmpts, const double *minDX, const double *maxDX, const double *dx,
                           const double *dX, double *errorInOut);
/* multicomponent estimator: update() = max over the sub-estimators' error fields from -1 (multicomponent.C:95-119) and
   maxRefinement() (multicomponent.C:122-157), whose in-place ascending-cell loop is reproduced exactly by a
   wavefront over lower-numbered neighbours.  errorOut NULL = resident error; maxRefinementOut nullable (then
   maxCellLevels is not read); wavesOut nullable. */
int amr_multicomponent(amr_ctx *ctx, int nEstimators, const double *const *errors, const int32_t *const *maxCellLevels,
                       double *errorOut, int32_t *maxRefinementOut, int *wavesOut);

/* ---- inputs of the load-balance decision; the decomposer and the redistribution stay host ---- */

/* fvMeshBalance::canBalance imbalance metric (fvMeshBalance.C:462-468): max over ranks of |nCells - ideal| / ideal with
   ideal = nGlobalCells / nProcs.  Collective when a communicator is attached. */
int amr_imbalance(amr_ctx *ctx, double *maxImbalanceRatio, int64_t *nGlobalCells);
/* hexRefRefinementHistory::markCommonCells (hexRefRefinementHistory.C:398-444): cellToCluster [n] (cluster of the cell's
   top-most ancestor, numbered in order of first appearance over ascending cells, -1 without history) and
   ::add (:448-487): blockedFace [f+b] (1 = blocked; only internal faces between two cells of one cluster are
   unblocked).  cellToClusterIn != NULL: use the host's clusters instead (polyRefiner, refinement::getCellClusters
   refinement.C:607-647, whose numbering is a hash-table order; refinement::add :1007-1046 needs only equality).
   All outputs nullable; nClusters = -1 when the clusters were given. */
int amr_history_clusters(amr_ctx *ctx, const int32_t *cellToClusterIn, int32_t *cellToClusterOut, uint8_t *blockedFace,
                         int64_t *nClusters, int64_t *nUnblocked);
/* hexRefRefinementHistory::apply (:490-545): decomposition [n] in/out -- every cluster goes to the processor given to
   the owner of its lowest in-cluster face.  nChanged (nullable) as counted by the reference. */
int amr_history_constrain_decomposition(amr_ctx *ctx, int32_t *decomposition, int64_t *nChanged);
/* procLoadNew of fvMeshBalance::canBalance (fvMeshBalance.C:497-503): cells per target processor, summed over ranks */
int amr_proc_load(amr_ctx *ctx, const int32_t *distribution, int nProcs, int64_t *procLoad);

/* Wait for everything enqueued on the context's stream and report what calls that were only stream-ordered (all data
   pointers on the device) could not: a peer-memory halo swap or reduction that timed out because a neighbour rank
   did not enter the same collective call (AMR_ERR_STATE), or a CUDA error. */
int amr_synchronize(amr_ctx *ctx);

/* number of kernels launched by this context since creation */
int64_t amr_launch_count(const amr_ctx *ctx);
/* bytes currently held in device memory by this context */
int64_t amr_device_bytes(const amr_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* AMR_B200_H */
