/*
 * amr_host.h -- C entry points of libamrhost.so, the C++ host layer above libamr_b200.so
 * (blastamr_b200/host: adaptiveFvMesh, errorEstimator + delta/scaledDelta/Lohner/fieldValue/
 * gradientRange/densityGradient, fvMeshRefiner + hexRefiner, dictionary).  These exist so that a
 * non-C++ embedder (the Python tests) can drive the same classes a solver would use through
 * `mesh.update()`; an OpenFOAM build would link the classes directly.
 * All functions return 0 or a negative code; amrhost_last_error() gives the FatalError text.
 */
#ifndef AMR_HOST_H
#define AMR_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct amrhost amrhost;

/* topology change callbacks (polyTopoChange is host bookkeeping).  The callback must install the new
   mesh with amrhost_set_mesh/... and the map with amrhost_set_map before returning 0. */
typedef int (*amrhost_topo_cb)(void *user, const int32_t *elems, int64_t n);

int amrhost_create(const char *dynamicMeshDictText, int device, amrhost **out);
void amrhost_destroy(amrhost *h);
const char *amrhost_last_error(const amrhost *h);          /* h may be NULL: error of the last failed create */
int amrhost_set_dict(amrhost *h, const char *dynamicMeshDictText);
int amrhost_set_topo_callbacks(amrhost *h, amrhost_topo_cb refine, amrhost_topo_cb unrefine, void *user);

int amrhost_set_mesh(amrhost *h, int64_t n, int64_t f, int64_t b, int64_t p, const int32_t *owner, const int32_t *neighbour,
                     int npatch, const char *const *patchNames, const int32_t *patchStart, const int32_t *patchSize,
                     const int32_t *patchType, const int32_t *patchNbr, int nSolutionD);
int amrhost_set_points(amrhost *h, const int32_t *pcOff, const int32_t *pcCells, const uint8_t *bndPoint);
int amrhost_set_edges(amrhost *h, int64_t nEdges, const int32_t *edgePts, const int32_t *ecOff, const int32_t *ecCells,
                      const uint8_t *bndEdge);
int amrhost_set_faces(amrhost *h, const int32_t *faceOff, const int32_t *facePts);
int amrhost_set_geometry(amrhost *h, const double *weights, const double *Sf, const double *magSf, const double *deltaCoeffs,
                         const double *V);
int amrhost_set_levels(amrhost *h, const int32_t *cellLevel, const int32_t *pointLevel, const int32_t *visibleCells,
                       const int32_t *splitParent, int64_t nSplit);
int amrhost_set_map(amrhost *h, int64_t nOld, int64_t nNew, const int32_t *cellMap, const int32_t *reverseCellMap);
int amrhost_mesh_changed(amrhost *h);                       /* upload addressing after the initial amrhost_set_* calls */

int amrhost_set_field(amrhost *h, const char *name, int nComp, const double *values, int64_t nCells);
int amrhost_get_field(amrhost *h, const char *name, double *values, int64_t capacity, int64_t *nValues);
int amrhost_set_time(amrhost *h, int32_t timeIndex, double timeValue);

int amrhost_update(amrhost *h, int *changed);               /* adaptiveFvMesh::update() */
int64_t amrhost_n_cells(const amrhost *h);
int amrhost_last_lists(amrhost *h, int32_t *cellsToRefine, int64_t *nRefine, int32_t *elemsToUnrefine, int64_t *nUnrefine);
int amrhost_get_error(amrhost *h, double *error, int64_t capacity);
/* keys of the parsed dictionary, space separated (dictionary unit tests, no GPU needed) */
int amrhost_parse_dict(const char *text, char *out, int64_t capacity);
/* names in the run-time selection tables, space separated: which=0 errorEstimator, 1 fvMeshRefiner */
int amrhost_selection_table(int which, char *out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif
