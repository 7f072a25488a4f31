/*
 * meshgen.c -- synthetic polyMesh generator (host bookkeeping, plain C).
 *
 * Stands in for the parts of the workflow the AMR decision path does NOT own:
 *   - blockMesh (single hex block, OpenFOAM cell/face/point numbering),
 *   - the hexRef3D/polyTopoChange topology change (cells cut 1->8, merged 8->1),
 *   - decomposePar ("simple"-style split into processor patches).
 * It only produces the arrays the reference reads on its hot path
 * (owner/neighbour/boundary, pointCells/cellPoints, cellLevel/pointLevel,
 * refinement-history visibleCells/parent) so that the CUDA path and the oracle
 * can be fed identical inputs.  Numbering conventions follow OpenFOAM:
 *   - cells of a block: c = i + nx*(j + ny*k); points likewise on (nx+1)(ny+1)(nz+1);
 *   - internal faces upper-triangular (by owner, then neighbour), owner<neighbour;
 *   - refining cell c keeps label c for child 0 and appends 7 new labels
 *     (reference: hexRef3D.C:1238-1255 "polyAddCell ... master cell celli");
 *   - merging keeps the lowest-numbered child as master and compacts labels
 *     (reference: hexRef3D.C:2078,2130);
 *   - history: hexRefRefinementHistory.C:1673-1732 (storeSplit / combineCells).
 * Nothing here is timed or graded as a kernel; it is the "host topology engine".
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

typedef int32_t i32;
typedef int64_t i64;
typedef uint8_t u8;

#define MG_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ mesh */

typedef struct mg_mesh {
    i64 n, f, b, p;
    i32 *owner;      /* f+b */
    i32 *neighbour;  /* f   */
    int npatch;
    i32 *patchStart; /* absolute face index */
    i32 *patchSize;
    i32 *patchType;  /* 0 = wall/patch, 1 = empty, 2 = processor */
    i32 *patchNbr;   /* neighbour rank for processor patches, else -1 */
    double *cc;      /* 3n cell centres */
    double *vol;     /* n  cell volumes */
    i32 *pcOff, *pcCells; /* pointCells CSR  (p+1, nnz) */
    i32 *cpOff, *cpPts;   /* cellPoints CSR  (n+1, nnz) */
    u8  *bndPoint;        /* p: point lies on a boundary face */
    i32 *cellLevel;       /* n */
    i32 *pointLevel;      /* p */
    i32 *visible;         /* n : history visibleCells_ */
    i32 *splitParent;     /* nsplit : history splitCells_[i].parent_ */
    i64 nsplit;
    i32 *cellGlobal;      /* n : global cell id (decomposed meshes), else NULL */
    i32 *pointGlobal;     /* p : global point id (decomposed meshes), else NULL */
    i32 *ppOff, *ppPts;   /* npatch+1 / ... : points of each processor patch, same order on both sides (decomposed meshes) */
    i32 *faceGlobal;      /* f+b: global face id (decomposed meshes), else NULL */
    i32 *pxyz;            /* 3p integer point coordinates in fine units (octree meshes) */
    /* fvMesh geometry (octree meshes): surfaceInterpolation::weights(), Sf(), magSf(), nonOrthDeltaCoeffs() */
    double *Sf;           /* 3(f+b) face area vectors, owner -> neighbour / outward */
    double *magSf;        /* f+b */
    double *weights;      /* f+b owner-side linear weights (1 on boundary faces) */
    double *deltaCoeffs;  /* f+b nonOrthDeltaCoeffs */
    /* 2-D (one cell thick) meshes: the edges normal to the plane, mesh.edgeCells() restricted to them */
    i64 nedge;
    i32 *edgePts;         /* 2*nedge: mesh.edges()[e] */
    i32 *ecOff, *ecCells; /* edgeCells CSR */
    u8 *bndEdge;          /* edge belongs to a boundary face of a non-empty patch */
    i32 *faceOff, *facePts; /* mesh.faces() as CSR (f+b+1, nnz): points around each face, hanging nodes included */
} mg_mesh;

static void *xmalloc(size_t s) {
    void *q = malloc(s ? s : 1);
    if (!q) { fprintf(stderr, "meshgen: out of memory (%zu bytes)\n", s); abort(); }
    return q;
}
static void *xcalloc(size_t n, size_t s) {
    void *q = calloc(n ? n : 1, s ? s : 1);
    if (!q) { fprintf(stderr, "meshgen: out of memory (%zu x %zu)\n", n, s); abort(); }
    return q;
}
static void *xrealloc(void *p, size_t s) {
    void *q = realloc(p, s ? s : 1);
    if (!q) { fprintf(stderr, "meshgen: out of memory (%zu bytes)\n", s); abort(); }
    return q;
}

MG_API void mg_mesh_free(mg_mesh *m) {
    if (!m) return;
    free(m->owner); free(m->neighbour); free(m->patchStart); free(m->patchSize);
    free(m->patchType); free(m->patchNbr); free(m->cc); free(m->vol);
    free(m->pcOff); free(m->pcCells); free(m->cpOff); free(m->cpPts);
    free(m->bndPoint); free(m->cellLevel); free(m->pointLevel); free(m->visible);
    free(m->splitParent); free(m->cellGlobal); free(m->pointGlobal); free(m->faceGlobal); free(m->ppOff); free(m->ppPts);
    free(m->pxyz); free(m->Sf); free(m->magSf); free(m->weights); free(m->deltaCoeffs);
    free(m->edgePts); free(m->ecOff); free(m->ecCells); free(m->bndEdge); free(m->faceOff); free(m->facePts);
    free(m);
}

/* sizes: n f b p npatch nsplit nnz(pointCells) nnz(cellPoints) */
MG_API void mg_mesh_sizes(const mg_mesh *m, i64 *out) {
    out[0] = m->n; out[1] = m->f; out[2] = m->b; out[3] = m->p;
    out[4] = m->npatch; out[5] = m->nsplit;
    out[6] = m->pcOff ? m->pcOff[m->p] : 0;
    out[7] = m->cpOff ? m->cpOff[m->n] : 0;
    out[8] = m->nedge;
    out[9] = m->ecOff ? m->ecOff[m->nedge] : 0;
    out[10] = m->faceOff ? m->faceOff[m->f + m->b] : 0;
    out[11] = m->ppOff ? m->ppOff[m->npatch] : 0;
}

/* array access by id (Python wraps the pointers without copying) */
MG_API void *mg_mesh_array(const mg_mesh *m, int id) {
    switch (id) {
    case 0: return m->owner;     case 1: return m->neighbour;
    case 2: return m->patchStart; case 3: return m->patchSize;
    case 4: return m->patchType; case 5: return m->patchNbr;
    case 6: return m->cc;        case 7: return m->vol;
    case 8: return m->pcOff;     case 9: return m->pcCells;
    case 10: return m->cpOff;    case 11: return m->cpPts;
    case 12: return m->bndPoint; case 13: return m->cellLevel;
    case 14: return m->pointLevel; case 15: return m->visible;
    case 16: return m->splitParent; case 17: return m->cellGlobal;
    case 18: return m->pointGlobal; case 19: return m->faceGlobal;
    case 20: return m->pxyz;
    case 21: return m->Sf; case 22: return m->magSf; case 23: return m->weights; case 24: return m->deltaCoeffs;
    case 25: return m->edgePts; case 26: return m->ecOff; case 27: return m->ecCells; case 28: return m->bndEdge;
    case 29: return m->faceOff; case 30: return m->facePts;
    case 31: return m->ppOff; case 32: return m->ppPts;
    }
    return NULL;
}

/* ---------------------------------------------------------------- octree */

typedef struct mg_oct {
    int nx, ny, nz, lcap;
    int dim2;            /* 1: quadtree in (x,y), cells keep their z extent (hexRef2D-style 1->4 cut) */
    double L[3];
    /* nodes */
    i64 nnode, capnode;
    i32 *nodeChild, *nodeCell, *nodeParent;
    i32 *freeBlocks; i64 nfree, capfree;      /* recycled 8-child blocks */
    /* cells (leaves) */
    i64 n, capn;
    i32 *cellNode;
    u8  *cellLvl;
    i32 *cx, *cy, *cz;                        /* coordinates at own level */
    /* history (hexRefRefinementHistory restated: visibleCells_, splitCells_.parent_) */
    i32 *visible;
    i32 *splitParent; i64 nsplit, capsplit;
    i32 *freeSplit; i64 nfreeSplit, capfreeSplit;
    /* boundary: ordered list of (side, patch) ; side 0..5 = xmin xmax ymin ymax zmin zmax */
    int nside; int sideId[6], sidePatch[6];
    int npatch; int patchType[8]; int patchNbr[8];
} mg_oct;

MG_API mg_oct *mg_oct_create(int nx, int ny, int nz, double lx, double ly, double lz,
                             int lcap, int nside, const int *sideId, const int *sidePatch,
                             int npatch, const int *patchType, const int *patchNbr) {
    mg_oct *o = (mg_oct *)xcalloc(1, sizeof(mg_oct));
    o->nx = nx; o->ny = ny; o->nz = nz; o->lcap = lcap;
    o->L[0] = lx; o->L[1] = ly; o->L[2] = lz;
    i64 n = (i64)nx * ny * nz;
    o->nnode = o->capnode = n;
    o->nodeChild = (i32 *)xmalloc(n * 4);
    o->nodeCell = (i32 *)xmalloc(n * 4);
    o->nodeParent = (i32 *)xmalloc(n * 4);
    o->n = o->capn = n;
    o->cellNode = (i32 *)xmalloc(n * 4);
    o->cellLvl = (u8 *)xcalloc(n, 1);
    o->cx = (i32 *)xmalloc(n * 4); o->cy = (i32 *)xmalloc(n * 4); o->cz = (i32 *)xmalloc(n * 4);
    o->visible = (i32 *)xmalloc(n * 4);
    #pragma omp parallel for schedule(static)
    for (i64 c = 0; c < n; c++) {
        o->nodeChild[c] = -1; o->nodeCell[c] = (i32)c; o->nodeParent[c] = -1;
        o->cellNode[c] = (i32)c;
        o->cx[c] = (i32)(c % nx); o->cy[c] = (i32)((c / nx) % ny); o->cz[c] = (i32)(c / ((i64)nx * ny));
        o->visible[c] = -1;
    }
    o->nside = nside; o->npatch = npatch;
    for (int s = 0; s < nside; s++) { o->sideId[s] = sideId[s]; o->sidePatch[s] = sidePatch[s]; }
    for (int q = 0; q < npatch; q++) { o->patchType[q] = patchType[q]; o->patchNbr[q] = patchNbr ? patchNbr[q] : -1; }
    return o;
}

MG_API void mg_oct_free(mg_oct *o) {
    if (!o) return;
    free(o->nodeChild); free(o->nodeCell); free(o->nodeParent); free(o->freeBlocks);
    free(o->cellNode); free(o->cellLvl); free(o->cx); free(o->cy); free(o->cz);
    free(o->visible); free(o->splitParent); free(o->freeSplit);
    free(o);
}

MG_API i64 mg_oct_ncells(const mg_oct *o) { return o->n; }
MG_API void mg_oct_set_2d(mg_oct *o, int on) { o->dim2 = on; }
/* z extent of a cell in its own level's units */
static inline i32 zext(const mg_oct *o, i64 c) { return o->dim2 ? (1 << o->cellLvl[c]) : 1; }

/* descend towards the cell (l; ix,iy,iz); returns node, *depth = level reached */
static inline i32 oct_locate(const mg_oct *o, int l, i32 ix, i32 iy, i32 iz, int *depth) {
    i32 rx = ix >> l, ry = iy >> l, rz = iz >> l;
    i32 node = (i32)(rx + (i64)o->nx * (ry + (i64)o->ny * rz));
    int d = 0;
    while (d < l && o->nodeChild[node] >= 0) {
        int sh = l - d - 1;
        int oc = ((ix >> sh) & 1) | (((iy >> sh) & 1) << 1) | (((iz >> sh) & 1) << 2);
        node = o->nodeChild[node] + oc;
        d++;
    }
    *depth = d;
    return node;
}

/* face neighbour of cell c in direction dir (0..5 = -x +x -y +y -z +z)
 * returns leaf cell id (same level or coarser), -1 = domain boundary, -2 = finer cells */
static inline i32 oct_face_nbr(const mg_oct *o, i64 c, int dir) {
    int l = o->cellLvl[c];
    i32 ix = o->cx[c], iy = o->cy[c], iz = o->cz[c];
    switch (dir) {
    case 0: ix--; break; case 1: ix++; break;
    case 2: iy--; break; case 3: iy++; break;
    case 4: iz--; break; default: iz += (o->dim2 ? (1 << l) : 1); break;
    }
    if (ix < 0 || iy < 0 || iz < 0 || ix >= (o->nx << l) || iy >= (o->ny << l) || iz >= (o->nz << l))
        return -1;
    int d;
    i32 node = oct_locate(o, l, ix, iy, iz, &d);
    if (o->nodeChild[node] >= 0) return -2;
    return o->nodeCell[node];
}

static i32 alloc_split(mg_oct *o, i32 parent) {
    i32 idx;
    if (o->nfreeSplit > 0) idx = o->freeSplit[--o->nfreeSplit];
    else {
        if (o->nsplit == o->capsplit) {
            o->capsplit = o->capsplit ? o->capsplit * 2 : 1024;
            o->splitParent = (i32 *)xrealloc(o->splitParent, o->capsplit * 4);
        }
        idx = (i32)o->nsplit++;
    }
    o->splitParent[idx] = parent;
    return idx;
}
static void free_split(mg_oct *o, i32 idx) {
    if (o->nfreeSplit == o->capfreeSplit) {
        o->capfreeSplit = o->capfreeSplit ? o->capfreeSplit * 2 : 1024;
        o->freeSplit = (i32 *)xrealloc(o->freeSplit, o->capfreeSplit * 4);
    }
    o->splitParent[idx] = -2; /* freed marker (reference: splitCell8 parent_ = -2 for free) */
    o->freeSplit[o->nfreeSplit++] = idx;
}

static void grow_cells(mg_oct *o, i64 need) {
    if (need <= o->capn) return;
    i64 cap = o->capn * 2; if (cap < need) cap = need;
    o->cellNode = (i32 *)xrealloc(o->cellNode, cap * 4);
    o->cellLvl = (u8 *)xrealloc(o->cellLvl, cap);
    o->cx = (i32 *)xrealloc(o->cx, cap * 4); o->cy = (i32 *)xrealloc(o->cy, cap * 4);
    o->cz = (i32 *)xrealloc(o->cz, cap * 4);
    o->visible = (i32 *)xrealloc(o->visible, cap * 4);
    o->capn = cap;
}
static i32 alloc_block(mg_oct *o) {
    if (o->nfree > 0) return o->freeBlocks[--o->nfree];
    if (o->nnode + 8 > o->capnode) {
        i64 cap = o->capnode * 2; if (cap < o->nnode + 8) cap = o->nnode + 8;
        o->nodeChild = (i32 *)xrealloc(o->nodeChild, cap * 4);
        o->nodeCell = (i32 *)xrealloc(o->nodeCell, cap * 4);
        o->nodeParent = (i32 *)xrealloc(o->nodeParent, cap * 4);
        o->capnode = cap;
    }
    i32 base = (i32)o->nnode; o->nnode += 8;
    return base;
}

/*
 * Refine the listed cells (ascending, unique, already 2:1 consistent: the caller
 * is the decision path).  cellMap (size n_new, may be NULL) receives new->old labels
 * exactly like mapPolyMesh::cellMap() for a hexRef3D cut: child 0 keeps the label,
 * the 7 added cells follow at the end in list order.  Returns the new cell count,
 * or -1 if a cell would exceed the level cap.
 */
MG_API i64 mg_oct_refine(mg_oct *o, const i32 *cells, i64 ncells, i32 *cellMap) {
    for (i64 k = 0; k < ncells; k++)
        if (o->cellLvl[cells[k]] >= o->lcap) return -1;
    const int NCH = o->dim2 ? 4 : 8;
    i64 nold = o->n, nnew = nold + (NCH - 1) * ncells;
    grow_cells(o, nnew);
    if (cellMap) for (i64 c = 0; c < nold; c++) cellMap[c] = (i32)c;
    for (i64 k = 0; k < ncells; k++) {
        i32 c = cells[k];
        i32 node = o->cellNode[c];
        int l = o->cellLvl[c];
        i32 px = o->cx[c], py = o->cy[c], pz = o->cz[c];
        i32 base = alloc_block(o);
        o->nodeChild[node] = base; o->nodeCell[node] = -1;
        for (int j = 0; j < 8; j++) { o->nodeChild[base + j] = -1; o->nodeCell[base + j] = -1; o->nodeParent[base + j] = node; }
        /* history: storeSplit (hexRefRefinementHistory.C:1673-1708) */
        i32 parentIndex;
        if (o->visible[c] != -1) { parentIndex = o->visible[c]; o->visible[c] = -1; }
        else parentIndex = alloc_split(o, -1);
        for (int j = 0; j < NCH; j++) {
            i32 id = (j == 0) ? c : (i32)(nold + (NCH - 1) * k + (j - 1));
            i32 nd = base + j;
            o->nodeChild[nd] = -1; o->nodeCell[nd] = id; o->nodeParent[nd] = node;
            o->cellNode[id] = nd; o->cellLvl[id] = (u8)(l + 1);
            o->cx[id] = 2 * px + (j & 1); o->cy[id] = 2 * py + ((j >> 1) & 1); o->cz[id] = 2 * pz + ((j >> 2) & 1);
            o->visible[id] = alloc_split(o, parentIndex);
            if (cellMap) cellMap[id] = c;
        }
    }
    o->n = nnew;
    return nnew;
}

/*
 * Merge the 8 siblings around each listed split point.  The point is given by the
 * 8 cell labels around it (pointCells row of the current mesh); cells8 has 8*nsp
 * entries.  cellMap (size n_new) = new->old; reverseMap (size n_old) = old->new for
 * surviving cells, or -(newMaster)-2 for merged-away cells (mapPolyMesh convention).
 * Returns new cell count or -1 on inconsistent input.
 */
MG_API i64 mg_oct_unrefine(mg_oct *o, const i32 *cells8, i64 nsp, i32 *cellMap, i32 *reverseMap) {
    i64 nold = o->n;
    const int NCH = o->dim2 ? 4 : 8;
    i32 *master = (i32 *)xmalloc(nold * 4); /* old label -> old master label, or self */
    u8 *removed = (u8 *)xcalloc(nold, 1);
    for (i64 c = 0; c < nold; c++) master[c] = (i32)c;
    for (i64 s = 0; s < nsp; s++) {
        const i32 *cs = cells8 + NCH * s;
        i32 pnode = o->nodeParent[o->cellNode[cs[0]]];
        if (pnode < 0) { free(master); free(removed); return -1; }
        i32 m = cs[0];
        for (int j = 0; j < NCH; j++) {
            if (o->nodeParent[o->cellNode[cs[j]]] != pnode) { free(master); free(removed); return -1; }
            if (cs[j] < m) m = cs[j];
        }
        /* history: combineCells (hexRefRefinementHistory.C:1711-1732) */
        i32 parentIndex = o->splitParent[o->visible[m]];
        for (int j = 0; j < NCH; j++) { free_split(o, o->visible[cs[j]]); o->visible[cs[j]] = -1; }
        o->visible[m] = parentIndex;
        int l = o->cellLvl[m];
        i32 base = o->nodeChild[pnode];
        /* parent becomes a leaf again carrying the master label */
        o->nodeChild[pnode] = -1; o->nodeCell[pnode] = m;
        i32 nx_ = o->cx[cs[0]] >> 1, ny_ = o->cy[cs[0]] >> 1, nz_ = o->cz[cs[0]] >> 1;
        o->cellNode[m] = pnode; o->cellLvl[m] = (u8)(l - 1);
        o->cx[m] = nx_; o->cy[m] = ny_; o->cz[m] = nz_;
        for (int j = 0; j < NCH; j++) if (cs[j] != m) { removed[cs[j]] = 1; master[cs[j]] = m; }
        if (o->nfree == o->capfree) {
            o->capfree = o->capfree ? o->capfree * 2 : 1024;
            o->freeBlocks = (i32 *)xrealloc(o->freeBlocks, o->capfree * 4);
        }
        o->freeBlocks[o->nfree++] = base;
    }
    /* compact labels, preserving order */
    i32 *newId = (i32 *)xmalloc(nold * 4);
    i64 nn = 0;
    for (i64 c = 0; c < nold; c++) newId[c] = removed[c] ? -1 : (i32)nn++;
    for (i64 c = 0; c < nold; c++) {
        if (removed[c]) { if (reverseMap) reverseMap[c] = -newId[master[c]] - 2; continue; }
        i32 id = newId[c];
        if (reverseMap) reverseMap[c] = id;
        if (cellMap) cellMap[id] = (i32)c;
        if (id != c) {
            o->cellNode[id] = o->cellNode[c]; o->cellLvl[id] = o->cellLvl[c];
            o->cx[id] = o->cx[c]; o->cy[id] = o->cy[c]; o->cz[id] = o->cz[c];
            o->visible[id] = o->visible[c];
        }
        o->nodeCell[o->cellNode[id]] = id;
    }
    o->n = nn;
    free(master); free(removed); free(newId);
    return nn;
}

/* ------------------------------------------------------- point hash (non-lattice points) */

typedef struct { uint64_t *key; i32 *val; i64 cap, cnt; } phash;
static inline uint64_t hmix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static void ph_init(phash *h, i64 cap) {
    i64 c = 1024; while (c < cap) c <<= 1;
    h->cap = c; h->cnt = 0;
    h->key = (uint64_t *)xmalloc(c * 8); h->val = (i32 *)xmalloc(c * 4);
    memset(h->key, 0xff, c * 8);
}
static void ph_free(phash *h) { free(h->key); free(h->val); }
static i32 ph_get(const phash *h, uint64_t k) {
    i64 i = hmix(k) & (h->cap - 1);
    while (h->key[i] != UINT64_MAX) { if (h->key[i] == k) return h->val[i]; i = (i + 1) & (h->cap - 1); }
    return -1;
}
static void ph_put_raw(phash *h, uint64_t k, i32 v) {
    i64 i = hmix(k) & (h->cap - 1);
    while (h->key[i] != UINT64_MAX) i = (i + 1) & (h->cap - 1);
    h->key[i] = k; h->val[i] = v; h->cnt++;
}
static void ph_put(phash *h, uint64_t k, i32 v) {
    if ((h->cnt + 1) * 2 > h->cap) {
        phash g; ph_init(&g, h->cap * 2);
        for (i64 i = 0; i < h->cap; i++) if (h->key[i] != UINT64_MAX) ph_put_raw(&g, h->key[i], h->val[i]);
        ph_free(h); *h = g;
    }
    ph_put_raw(h, k, v);
}
static inline uint64_t pkey(i32 X, i32 Y, i32 Z) {
    return ((uint64_t)(uint32_t)X) | ((uint64_t)(uint32_t)Y << 21) | ((uint64_t)(uint32_t)Z << 42);
}

/* boundary leaves on one side in blockMesh order (refined: depth-first below each root face) */
static void side_collect_rec(const mg_oct *o, i32 node, int side, i32 **out, i64 *cnt, i64 *cap) {
    if (o->nodeChild[node] < 0) {
        if (*cnt == *cap) { *cap = *cap ? *cap * 2 : 4096; *out = (i32 *)xrealloc(*out, *cap * 4); }
        (*out)[(*cnt)++] = o->nodeCell[node];
        return;
    }
    int axis = side >> 1, hi = side & 1;
    for (int j = 0; j < (o->dim2 ? 4 : 8); j++) {
        int bit = (j >> axis) & 1;
        if (!(o->dim2 && axis == 2) && bit != hi) continue;
        side_collect_rec(o, o->nodeChild[node] + j, side, out, cnt, cap);
    }
}
static void side_collect(const mg_oct *o, int side, i32 **out, i64 *cnt, i64 *cap) {
    int nx = o->nx, ny = o->ny, nz = o->nz;
    /* blockMesh patch-face order (checked against tests/testCases/hex3D of the reference):
       x sides: k outer, j inner; y sides: i outer, k inner; z sides: i outer, j inner */
    if (side < 2) {
        int i = side == 0 ? 0 : nx - 1;
        for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++)
            side_collect_rec(o, (i32)(i + (i64)nx * (j + (i64)ny * k)), side, out, cnt, cap);
    } else if (side < 4) {
        int j = side == 2 ? 0 : ny - 1;
        for (int i = 0; i < nx; i++) for (int k = 0; k < nz; k++)
            side_collect_rec(o, (i32)(i + (i64)nx * (j + (i64)ny * k)), side, out, cnt, cap);
    } else {
        int k = side == 4 ? 0 : nz - 1;
        for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++)
            side_collect_rec(o, (i32)(i + (i64)nx * (j + (i64)ny * k)), side, out, cnt, cap);
    }
}

static int cmp_i32(const void *a, const void *b) {
    i32 x = *(const i32 *)a, y = *(const i32 *)b; return (x > y) - (x < y);
}

/* ---- faces() and edges() of a 3-D octree mesh (polyRefiner needs them) ---------------------------------- */
typedef struct {
    const mg_oct *o; const phash *H; i64 npx, npy; int FS;
    i32 *buf; i64 cnt, cap;
} facewalk;
static inline i32 point_id(const facewalk *W, i32 X, i32 Y, i32 Z) {
    if (((X | Y | Z) & (W->FS - 1)) == 0) return (i32)((X / W->FS) + W->npx * ((Y / W->FS) + W->npy * (i64)(Z / W->FS)));
    return ph_get(W->H, pkey(X, Y, Z));
}
static void walk_push(facewalk *W, i32 id) {
    if (W->cnt == W->cap) { W->cap = W->cap ? W->cap * 2 : 1 << 16; W->buf = (i32 *)xrealloc(W->buf, W->cap * 4); }
    W->buf[W->cnt++] = id;
}
/* points on the open segment A -> B (hanging nodes), by bisection: a finer point implies the midpoint */
static void walk_side(facewalk *W, const i32 *A, const i32 *B) {
    i32 M[3] = {(A[0] + B[0]) / 2, (A[1] + B[1]) / 2, (A[2] + B[2]) / 2};
    i32 len = abs(B[0] - A[0]) + abs(B[1] - A[1]) + abs(B[2] - A[2]);
    if (len < 2 || (len & 1)) return;
    i32 id = point_id(W, M[0], M[1], M[2]);
    if (id < 0) return;
    walk_side(W, A, M);
    walk_push(W, id);
    walk_side(W, M, B);
}
/* integer box of a cell in fine units */
static inline void cell_box(const mg_oct *o, i64 c, i32 *lo, i32 *hi) {
    int sh = o->lcap - o->cellLvl[c];
    lo[0] = o->cx[c] << sh; lo[1] = o->cy[c] << sh; lo[2] = o->cz[c] << sh;
    hi[0] = (o->cx[c] + 1) << sh; hi[1] = (o->cy[c] + 1) << sh; hi[2] = (o->cz[c] + zext(o, c)) << sh;
}

/* mesh.edges() / edgeCells() / "edge of a boundary face" from faces() and pointCells(): consecutive perimeter points,
   unique, numbered in order of first appearance over ascending faces; cells around an edge = pointCells(a) ^ pointCells(b) */
static void edges_from_faces(mg_mesh *m) {
    const i64 f = m->f, nfaces = m->f + m->b;
    const i32 *pcnt = m->pcOff;
    phash E; ph_init(&E, m->faceOff[nfaces] + 16);
    i64 ne = 0, ecap = 0;
    i32 *ep = NULL; u8 *be = NULL;
    for (i64 i = 0; i < nfaces; i++) {
        i32 s0 = m->faceOff[i], s1 = m->faceOff[i + 1];
        for (i32 k = s0; k < s1; k++) {
            i32 a = m->facePts[k], c2 = m->facePts[k + 1 < s1 ? k + 1 : s0];
            i32 lo_ = a < c2 ? a : c2, hi_ = a < c2 ? c2 : a;
            uint64_t key = ((uint64_t)(uint32_t)lo_ << 32) | (uint32_t)hi_;
            i32 id = ph_get(&E, key);
            if (id < 0) {
                if (ne == ecap) { ecap = ecap ? ecap * 2 : 1 << 16; ep = (i32 *)xrealloc(ep, ecap * 8); be = (u8 *)xrealloc(be, ecap); }
                id = (i32)ne++;
                ph_put(&E, key, id);
                ep[2 * id] = lo_; ep[2 * id + 1] = hi_; be[id] = 0;
            }
            if (i >= f) be[id] = 1;
        }
    }
    ph_free(&E);
    m->nedge = ne; m->edgePts = ep; m->bndEdge = be;
    m->ecOff = (i32 *)xcalloc(ne + 1, 4);
    for (int pass = 0; pass < 2; pass++) {
        for (i64 e = 0; e < ne; e++) {
            i32 q = ep[2 * e], partner = ep[2 * e + 1];
            i32 a = pcnt[q], ae = pcnt[q + 1], c2 = pcnt[partner], ce = pcnt[partner + 1];
            i32 k = pass ? m->ecOff[e] : 0;
            while (a < ae && c2 < ce) {
                i32 pa = m->pcCells[a], pc = m->pcCells[c2];
                if (pa == pc) { if (pass) m->ecCells[k] = pa; k++; a++; c2++; } else if (pa < pc) a++; else c2++;
            }
            if (!pass) m->ecOff[e + 1] = k;
        }
        if (!pass) {
            for (i64 e = 0; e < ne; e++) m->ecOff[e + 1] += m->ecOff[e];
            m->ecCells = (i32 *)xmalloc((m->ecOff[ne] ? m->ecOff[ne] : 1) * 4);
        }
    }
}

/*
 * Build the polyMesh-style arrays for the current octree.
 * flags: bit0 = build point addressing (pointCells, cellPoints, bndPoint, pointLevel).
 */
MG_API mg_mesh *mg_oct_build(const mg_oct *o, int flags) {
    mg_mesh *m = (mg_mesh *)xcalloc(1, sizeof(mg_mesh));
    i64 n = o->n;
    m->n = n;
    const int FS = 1 << o->lcap;

    /* --- neighbour table */
    i32 *nb = (i32 *)xmalloc(n * 6 * 4);
    #pragma omp parallel for schedule(static)
    for (i64 c = 0; c < n; c++)
        for (int d = 0; d < 6; d++) nb[6 * c + d] = oct_face_nbr(o, c, d);

    /* --- internal faces: emitted from the finer side, or from the lower label when equal level */
    i32 *cnt = (i32 *)xcalloc(n + 1, 4);
    for (i64 c = 0; c < n; c++) {
        int l = o->cellLvl[c];
        for (int d = 0; d < 6; d++) {
            i32 q = nb[6 * c + d];
            if (q < 0) continue;
            int lq = o->cellLvl[q];
            if (lq == l) { if (c < q) cnt[c + 1]++; }
            else cnt[(c < q ? c : q) + 1]++;
        }
    }
    for (i64 c = 0; c < n; c++) cnt[c + 1] += cnt[c];
    i64 f = cnt[n];
    i32 *nei = (i32 *)xmalloc(f * 4);
    i32 *own = NULL;
    {
        i32 *fill = (i32 *)xmalloc((n + 1) * 4);
        memcpy(fill, cnt, (n + 1) * 4);
        for (i64 c = 0; c < n; c++) {
            int l = o->cellLvl[c];
            for (int d = 0; d < 6; d++) {
                i32 q = nb[6 * c + d];
                if (q < 0) continue;
                int lq = o->cellLvl[q];
                if (lq == l) { if (c < q) nei[fill[c]++] = q; }
                else { i32 a = c < q ? (i32)c : q, bq = c < q ? q : (i32)c; nei[fill[a]++] = bq; }
            }
        }
        free(fill);
        #pragma omp parallel for schedule(dynamic, 4096)
        for (i64 c = 0; c < n; c++) {
            i32 s = cnt[c], e = cnt[c + 1];
            /* insertion sort of the (short) row */
            for (i32 i = s + 1; i < e; i++) {
                i32 v = nei[i], j = i - 1;
                while (j >= s && nei[j] > v) { nei[j + 1] = nei[j]; j--; }
                nei[j + 1] = v;
            }
        }
    }
    /* --- boundary faces in patch order */
    i32 *bcell = NULL; i64 bcnt = 0, bcap = 0;
    m->npatch = o->npatch;
    m->patchStart = (i32 *)xcalloc(o->npatch, 4); m->patchSize = (i32 *)xcalloc(o->npatch, 4);
    m->patchType = (i32 *)xcalloc(o->npatch, 4); m->patchNbr = (i32 *)xcalloc(o->npatch, 4);
    u8 *bside = NULL;
    {
        int curPatch = -1;
        for (int s = 0; s < o->nside; s++) {
            int q = o->sidePatch[s];
            i64 before = bcnt;
            side_collect(o, o->sideId[s], &bcell, &bcnt, &bcap);
            if (q != curPatch) { m->patchStart[q] = (i32)(f + before); curPatch = q; }
            m->patchSize[q] += (i32)(bcnt - before);
            bside = (u8 *)xrealloc(bside, bcnt ? bcnt : 1);
            for (i64 i = before; i < bcnt; i++) bside[i] = (u8)o->sideId[s];
        }
        for (int q = 0; q < o->npatch; q++) { m->patchType[q] = o->patchType[q]; m->patchNbr[q] = o->patchNbr[q]; }
    }
    i64 b = bcnt;
    m->f = f; m->b = b;
    own = (i32 *)xmalloc((f + b) * 4);
    #pragma omp parallel for schedule(static)
    for (i64 c = 0; c < n; c++) for (i32 i = cnt[c]; i < cnt[c + 1]; i++) own[i] = (i32)c;
    for (i64 i = 0; i < b; i++) own[f + i] = bcell[i];
    m->owner = own; m->neighbour = nei;
    free(cnt); free(bcell);

    /* --- geometry, levels, history */
    m->cc = (double *)xmalloc(n * 3 * 8); m->vol = (double *)xmalloc(n * 8);
    m->cellLevel = (i32 *)xmalloc(n * 4); m->visible = (i32 *)xmalloc(n * 4);
    double h0x = o->L[0] / o->nx, h0y = o->L[1] / o->ny, h0z = o->L[2] / o->nz;
    #pragma omp parallel for schedule(static)
    for (i64 c = 0; c < n; c++) {
        int l = o->cellLvl[c];
        double s = 1.0 / (double)(1 << l);
        m->cc[3 * c + 0] = (o->cx[c] + 0.5) * s * h0x;
        m->cc[3 * c + 1] = (o->cy[c] + 0.5) * s * h0y;
        m->cc[3 * c + 2] = (o->cz[c] + 0.5 * zext(o, c)) * s * h0z;
        m->vol[c] = (h0x * s) * (h0y * s) * (h0z * s * zext(o, c));
        m->cellLevel[c] = l; m->visible[c] = o->visible[c];
    }
    m->nsplit = o->nsplit;
    m->splitParent = (i32 *)xmalloc((o->nsplit ? o->nsplit : 1) * 4);
    memcpy(m->splitParent, o->splitParent, o->nsplit * 4);

    if (flags & 2) {
        /* face geometry from the cells' boxes: the face is the contact rectangle of two cubes */
        i64 nf = f + b;
        m->Sf = (double *)xcalloc(nf * 3, 8); m->magSf = (double *)xmalloc(nf * 8);
        m->weights = (double *)xmalloc(nf * 8); m->deltaCoeffs = (double *)xmalloc(nf * 8);
        double h0[3] = {h0x, h0y, h0z};
        #pragma omp parallel for schedule(static)
        for (i64 i = 0; i < nf; i++) {
            i64 c = own[i];
            int lc = o->cellLvl[c];
            double sc = 1.0 / (double)(1 << lc);
            double lo_c[3] = {o->cx[c] * sc * h0x, o->cy[c] * sc * h0y, o->cz[c] * sc * h0z};
            double hi_c[3] = {(o->cx[c] + 1) * sc * h0x, (o->cy[c] + 1) * sc * h0y, (o->cz[c] + zext(o, c)) * sc * h0z};
            double Cc[3] = {m->cc[3 * c], m->cc[3 * c + 1], m->cc[3 * c + 2]};
            double Cf[3], S[3] = {0, 0, 0}, Cn[3];
            if (i < f) {
                i64 q = nei[i];
                int lq = o->cellLvl[q];
                double sq = 1.0 / (double)(1 << lq);
                i32 cq[3] = {o->cx[q], o->cy[q], o->cz[q]};
                i32 ccv[3] = {o->cx[c], o->cy[c], o->cz[c]};
                int axis = -1, sign = 0;
                double area = 1.0;
                for (int a = 0; a < 3; a++) {
                    /* integer comparison in fine units decides the contact axis */
                    i64 loC = (i64)ccv[a] << (o->lcap - lc), hiC = (i64)(ccv[a] + (a == 2 ? zext(o, c) : 1)) << (o->lcap - lc);
                    i64 loQ = (i64)cq[a] << (o->lcap - lq), hiQ = (i64)(cq[a] + (a == 2 ? zext(o, q) : 1)) << (o->lcap - lq);
                    if (hiC == loQ) { axis = a; sign = 1; }
                    else if (hiQ == loC) { axis = a; sign = -1; }
                }
                for (int a = 0; a < 3; a++) {
                    double loQ = cq[a] * sq * h0[a], hiQ = (cq[a] + (a == 2 ? zext(o, q) : 1)) * sq * h0[a];
                    if (a == axis) { Cf[a] = sign > 0 ? hi_c[a] : lo_c[a]; continue; }
                    double lo = lo_c[a] > loQ ? lo_c[a] : loQ, hi = hi_c[a] < hiQ ? hi_c[a] : hiQ;
                    area *= (hi - lo); Cf[a] = 0.5 * (lo + hi);
                }
                S[axis] = sign * area;
                Cn[0] = m->cc[3 * q]; Cn[1] = m->cc[3 * q + 1]; Cn[2] = m->cc[3 * q + 2];
                double magS = fabs(area);
                double SfdOwn = fabs(S[0] * (Cf[0] - Cc[0]) + S[1] * (Cf[1] - Cc[1]) + S[2] * (Cf[2] - Cc[2]));
                double SfdNei = fabs(S[0] * (Cn[0] - Cf[0]) + S[1] * (Cn[1] - Cf[1]) + S[2] * (Cn[2] - Cf[2]));
                m->weights[i] = SfdNei / (SfdOwn + SfdNei);          /* surfaceInterpolation::makeWeights */
                double d[3] = {Cn[0] - Cc[0], Cn[1] - Cc[1], Cn[2] - Cc[2]};
                double nd = (S[0] * d[0] + S[1] * d[1] + S[2] * d[2]) / magS;
                double md = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                m->deltaCoeffs[i] = 1.0 / (nd > 0.05 * md ? nd : 0.05 * md);   /* makeNonOrthDeltaCoeffs */
                m->magSf[i] = magS;
            } else {
                int side = bside[i - f];
                int axis = side >> 1, hi = side & 1;
                double area = 1.0;
                for (int a = 0; a < 3; a++) if (a != axis) area *= (hi_c[a] - lo_c[a]);
                S[axis] = hi ? area : -area;
                m->magSf[i] = area;
                m->weights[i] = 1.0;
                double dn = 0.5 * (hi_c[axis] - lo_c[axis]);
                m->deltaCoeffs[i] = 1.0 / dn;
            }
            m->Sf[3 * i] = S[0]; m->Sf[3 * i + 1] = S[1]; m->Sf[3 * i + 2] = S[2];
        }
    }
    u8 *bsideKeep = bside;

    if (!(flags & 1)) { free(nb); free(bsideKeep); return m; }

    /* --- points: lattice points first (blockMesh order), then hanging/new points in order of
           first appearance over ascending cells */
    i64 npx = o->nx + 1, npy = o->ny + 1, npz = o->nz + 1;
    i64 plat = npx * npy * npz;
    phash H; i64 nfine = 0;
    for (i64 c = 0; c < n; c++) if (o->cellLvl[c] > 0) nfine++;
    ph_init(&H, nfine * 4 + 16);
    i64 p = plat;
    i32 *extra = NULL; i64 excap = 0; /* coords of non-lattice points */
    for (i64 c = 0; c < n; c++) {
        int l = o->cellLvl[c];
        if (l == 0) continue;
        int s = FS >> l;
        for (int j = 0; j < 8; j++) {
            i32 X = (o->cx[c] + (j & 1)) * s, Y = (o->cy[c] + ((j >> 1) & 1)) * s, Z = (o->cz[c] + ((j >> 2) & 1) * zext(o, c)) * s;
            if (((X | Y | Z) & (FS - 1)) == 0) continue;
            uint64_t k = pkey(X, Y, Z);
            if (ph_get(&H, k) < 0) {
                ph_put(&H, k, (i32)p);
                i64 e = p - plat;
                if (e == excap) { excap = excap ? excap * 2 : 4096; extra = (i32 *)xrealloc(extra, excap * 3 * 4); }
                extra[3 * e] = X; extra[3 * e + 1] = Y; extra[3 * e + 2] = Z;
                p++;
            }
        }
    }
    m->p = p;
    m->pxyz = (i32 *)xmalloc(p * 3 * 4);
    #pragma omp parallel for schedule(static)
    for (i64 q = 0; q < plat; q++) {
        m->pxyz[3 * q] = (i32)(q % npx) * FS; m->pxyz[3 * q + 1] = (i32)((q / npx) % npy) * FS;
        m->pxyz[3 * q + 2] = (i32)(q / (npx * npy)) * FS;
    }
    if (p > plat) memcpy(m->pxyz + 3 * plat, extra, (p - plat) * 3 * 4);
    free(extra);

    /* --- pointCells: every leaf whose closed cube contains the point (8 octants) */
    i32 *tmp = (i32 *)xmalloc(p * 8 * 4);
    i32 *pcnt = (i32 *)xcalloc(p + 1, 4);
    m->bndPoint = (u8 *)xmalloc(p); m->pointLevel = (i32 *)xmalloc(p * 4);
    i32 MX = o->nx * FS, MY = o->ny * FS, MZ = o->nz * FS;
    int lcap = o->lcap;
    #pragma omp parallel for schedule(static)
    for (i64 q = 0; q < p; q++) {
        i32 X = m->pxyz[3 * q], Y = m->pxyz[3 * q + 1], Z = m->pxyz[3 * q + 2];
        m->bndPoint[q] = (X == 0 || Y == 0 || Z == 0 || X == MX || Y == MY || Z == MZ);
        int pl = 0; { i32 t = X | Y | Z; int tz = 0; while (tz < lcap && !((t >> tz) & 1)) tz++; pl = lcap - tz; }
        m->pointLevel[q] = pl;
        i32 *row = tmp + 8 * q; int k = 0;
        for (int j = 0; j < 8; j++) {
            i32 fx = X - ((j & 1) ? 0 : 1), fy = Y - (((j >> 1) & 1) ? 0 : 1), fz = Z - (((j >> 2) & 1) ? 0 : 1);
            if (fx < 0 || fy < 0 || fz < 0 || fx >= MX || fy >= MY || fz >= MZ) continue;
            int d; i32 node = oct_locate(o, lcap, fx, fy, o->dim2 ? (fz & ~(FS - 1)) : fz, &d);
            i32 c = o->nodeCell[node];
            int dup = 0; for (int t = 0; t < k; t++) if (row[t] == c) { dup = 1; break; }
            if (!dup) row[k++] = c;
        }
        qsort(row, k, 4, cmp_i32);
        pcnt[q + 1] = k;
    }
    for (i64 q = 0; q < p; q++) pcnt[q + 1] += pcnt[q];
    i64 nnz = pcnt[p];
    m->pcOff = pcnt; m->pcCells = (i32 *)xmalloc(nnz * 4);
    #pragma omp parallel for schedule(static)
    for (i64 q = 0; q < p; q++) memcpy(m->pcCells + pcnt[q], tmp + 8 * q, (size_t)(pcnt[q + 1] - pcnt[q]) * 4);
    free(tmp);
    /* --- cellPoints = transpose (rows ascending in point label) */
    i32 *cpo = (i32 *)xcalloc(n + 1, 4);
    for (i64 i = 0; i < nnz; i++) cpo[m->pcCells[i] + 1]++;
    for (i64 c = 0; c < n; c++) cpo[c + 1] += cpo[c];
    i32 *cpp = (i32 *)xmalloc(nnz * 4);
    {
        i32 *fill = (i32 *)xmalloc((n + 1) * 4); memcpy(fill, cpo, (n + 1) * 4);
        for (i64 q = 0; q < p; q++) for (i32 i = pcnt[q]; i < pcnt[q + 1]; i++) cpp[fill[m->pcCells[i]]++] = (i32)q;
        free(fill);
    }
    m->cpOff = cpo; m->cpPts = cpp;
    /* --- faces() and edges() (flag 4): perimeter walk of every face with hanging nodes.  Also valid for the
           2-D quadtree: a through-thickness side has no midpoint, in-plane sides bisect as in 3-D */
    if (flags & 4) {
        facewalk W = {o, &H, npx, npy, FS, NULL, 0, 0};
        i64 nfaces = f + b;
        m->faceOff = (i32 *)xcalloc(nfaces + 1, 4);
        for (i64 i = 0; i < nfaces; i++) {
            i32 lo[3], hi[3];
            cell_box(o, own[i], lo, hi);
            int axis = -1;
            if (i < f) {
                i32 lq[3], hq[3];
                cell_box(o, nei[i], lq, hq);
                for (int a = 0; a < 3; a++) {
                    if (hi[a] == lq[a]) { axis = a; lo[a] = hi[a]; }
                    else if (hq[a] == lo[a]) { axis = a; hi[a] = lo[a]; }
                    else { if (lq[a] > lo[a]) lo[a] = lq[a]; if (hq[a] < hi[a]) hi[a] = hq[a]; }
                }
            }
            if (i >= f) {
                int side = bsideKeep[i - f];
                axis = side >> 1;
                if (side & 1) lo[axis] = hi[axis]; else hi[axis] = lo[axis];
            }
            int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
            i32 C[4][3];
            for (int k = 0; k < 4; k++) {
                C[k][axis] = lo[axis];
                C[k][a1] = (k == 0 || k == 3) ? lo[a1] : hi[a1];
                C[k][a2] = (k < 2) ? lo[a2] : hi[a2];
            }
            for (int k = 0; k < 4; k++) {
                walk_push(&W, point_id(&W, C[k][0], C[k][1], C[k][2]));
                walk_side(&W, C[k], C[(k + 1) & 3]);
            }
            m->faceOff[i + 1] = (i32)W.cnt;
        }
        m->facePts = W.buf;
        edges_from_faces(m);
    }
    /* --- 2-D without flag 4: only the edges normal to the plane (one per point column and layer; what the
           2-D cutter selects from), edgeCells = cells around it */
    if (o->dim2 && !(flags & 4)) {
        i64 ne = 0;
        for (i64 q = 0; q < p; q++) if (m->pxyz[3 * q + 2] < MZ) ne++;
        m->nedge = ne;
        m->edgePts = (i32 *)xmalloc((ne ? ne : 1) * 2 * 4);
        m->bndEdge = (u8 *)xmalloc(ne ? ne : 1);
        m->ecOff = (i32 *)xcalloc(ne + 1, 4);
        i64 e = 0;
        for (i64 q = 0; q < p; q++) {
            i32 X = m->pxyz[3 * q], Y = m->pxyz[3 * q + 1], Z = m->pxyz[3 * q + 2];
            if (Z >= MZ) continue;
            i32 Z2 = Z + FS, partner;
            if (((X | Y) & (FS - 1)) == 0) partner = (i32)((X / FS) + npx * ((Y / FS) + npy * (i64)(Z2 / FS)));
            else partner = ph_get(&H, pkey(X, Y, Z2));
            m->edgePts[2 * e] = (i32)q; m->edgePts[2 * e + 1] = partner;
            m->bndEdge[e] = (X == 0 || Y == 0 || X == MX || Y == MY);
            /* cells around the edge = pointCells(q) ^ pointCells(partner) (sorted rows) */
            i32 a = pcnt[q], ae = pcnt[q + 1], c2 = pcnt[partner], ce = pcnt[partner + 1], k = 0;
            while (a < ae && c2 < ce) {
                i32 pa = m->pcCells[a], pc = m->pcCells[c2];
                if (pa == pc) { k++; a++; c2++; } else if (pa < pc) a++; else c2++;
            }
            m->ecOff[e + 1] = k;
            e++;
        }
        for (i64 i = 0; i < ne; i++) m->ecOff[i + 1] += m->ecOff[i];
        m->ecCells = (i32 *)xmalloc((m->ecOff[ne] ? m->ecOff[ne] : 1) * 4);
        for (i64 i = 0; i < ne; i++) {
            i32 q = m->edgePts[2 * i], partner = m->edgePts[2 * i + 1];
            i32 a = pcnt[q], ae = pcnt[q + 1], c2 = pcnt[partner], ce = pcnt[partner + 1], k = m->ecOff[i];
            while (a < ae && c2 < ce) {
                i32 pa = m->pcCells[a], pc = m->pcCells[c2];
                if (pa == pc) { m->ecCells[k++] = pa; a++; c2++; } else if (pa < pc) a++; else c2++;
            }
        }
    }
    ph_free(&H);
    free(nb); free(bsideKeep);
    return m;
}

/* ------------------------------------------------------------ decomposition */

/*
 * Split a global mesh by a per-cell rank array (decomposePar-like):
 *  - local cells/points keep ascending global order,
 *  - local internal faces keep global order (stays upper-triangular),
 *  - global patches first (same ids), then one processor patch per neighbour rank
 *    (ascending rank), faces in ascending global face index so that both sides of a
 *    processor patch list their faces in the same order.
 * out[r] receives the local mesh of rank r.
 */
static int decompose_impl(const mg_mesh *g, const i32 *cellRank, int nranks, mg_mesh **out, int only) {
    i64 n = g->n, f = g->f, b = g->b, p = g->p;
    i32 *loc = (i32 *)xmalloc(n * 4);
    i64 *nloc = (i64 *)xcalloc(nranks, 8);
    for (i64 c = 0; c < n; c++) loc[c] = (i32)nloc[cellRank[c]]++;
    for (int r = 0; r < nranks; r++) {
        if (only >= 0 && r != only) continue;
        mg_mesh *m = (mg_mesh *)xcalloc(1, sizeof(mg_mesh));
        out[only >= 0 ? 0 : r] = m;
        i64 nl = nloc[r];
        m->n = nl;
        m->cellGlobal = (i32 *)xmalloc(nl * 4);
        /* counts */
        i64 fl = 0;
        i64 *pcount = (i64 *)xcalloc(nranks, 8); /* cut faces per neighbour rank */
        for (i64 i = 0; i < f; i++) {
            int ro = cellRank[g->owner[i]], rn = cellRank[g->neighbour[i]];
            if (ro == r && rn == r) fl++;
            else if (ro == r) pcount[rn]++;
            else if (rn == r) pcount[ro]++;
        }
        int nproc = 0; for (int q = 0; q < nranks; q++) if (pcount[q]) nproc++;
        int np = g->npatch + nproc;
        m->npatch = np;
        m->patchStart = (i32 *)xcalloc(np, 4); m->patchSize = (i32 *)xcalloc(np, 4);
        m->patchType = (i32 *)xcalloc(np, 4); m->patchNbr = (i32 *)xcalloc(np, 4);
        i64 bl = 0;
        for (int q = 0; q < g->npatch; q++) {
            i64 cntq = 0;
            for (i64 i = 0; i < g->patchSize[q]; i++)
                if (cellRank[g->owner[g->patchStart[q] + i]] == r) cntq++;
            m->patchStart[q] = (i32)(fl + bl); m->patchSize[q] = (i32)cntq;
            m->patchType[q] = g->patchType[q]; m->patchNbr[q] = g->patchNbr ? g->patchNbr[q] : -1;
            bl += cntq;
        }
        i64 *pstart = (i64 *)xcalloc(nranks, 8);
        { int q2 = g->npatch;
          for (int q = 0; q < nranks; q++) if (pcount[q]) {
              m->patchStart[q2] = (i32)(fl + bl); m->patchSize[q2] = (i32)pcount[q];
              m->patchType[q2] = 2; m->patchNbr[q2] = q; pstart[q] = fl + bl; bl += pcount[q]; q2++;
          } }
        m->f = fl; m->b = bl;
        m->owner = (i32 *)xmalloc((fl + bl) * 4); m->neighbour = (i32 *)xmalloc((fl ? fl : 1) * 4);
        m->faceGlobal = (i32 *)xmalloc((fl + bl) * 4);
        u8 *flip = (u8 *)xcalloc(fl + bl + 1, 1);   /* local owner is the global neighbour */
        i64 fi = 0;
        i64 *pfill = (i64 *)xmalloc(nranks * 8); memcpy(pfill, pstart, nranks * 8);
        for (i64 i = 0; i < f; i++) {
            i32 o_ = g->owner[i], n_ = g->neighbour[i];
            int ro = cellRank[o_], rn = cellRank[n_];
            if (ro == r && rn == r) { m->owner[fi] = loc[o_]; m->neighbour[fi] = loc[n_]; m->faceGlobal[fi] = (i32)i; fi++; }
            else if (ro == r) { i64 k = pfill[rn]++; m->owner[k] = loc[o_]; m->faceGlobal[k] = (i32)i; }
            else if (rn == r) { i64 k = pfill[ro]++; m->owner[k] = loc[n_]; m->faceGlobal[k] = (i32)i; flip[k] = 1; }
        }
        for (int q = 0; q < g->npatch; q++) {
            i64 k = m->patchStart[q];
            for (i64 i = 0; i < g->patchSize[q]; i++) {
                i64 gf = g->patchStart[q] + i;
                if (cellRank[g->owner[gf]] == r) { m->owner[k] = loc[g->owner[gf]]; m->faceGlobal[k] = (i32)gf; k++; }
            }
        }
        /* face geometry: cut faces seen from the global neighbour get Sf reversed and the
           complementary weight (coupledFvPatch weights are computed from the local side) */
        if (g->Sf) {
            i64 nfl = fl + bl;
            m->Sf = (double *)xmalloc(nfl * 3 * 8); m->magSf = (double *)xmalloc(nfl * 8);
            m->weights = (double *)xmalloc(nfl * 8); m->deltaCoeffs = (double *)xmalloc(nfl * 8);
            for (i64 k = 0; k < nfl; k++) {
                i64 gf = m->faceGlobal[k];
                double sg = flip[k] ? -1.0 : 1.0;
                m->Sf[3 * k] = sg * g->Sf[3 * gf]; m->Sf[3 * k + 1] = sg * g->Sf[3 * gf + 1]; m->Sf[3 * k + 2] = sg * g->Sf[3 * gf + 2];
                m->magSf[k] = g->magSf[gf]; m->deltaCoeffs[k] = g->deltaCoeffs[gf];
                m->weights[k] = flip[k] ? 1.0 - g->weights[gf] : g->weights[gf];
            }
        }
        free(flip);
        /* cells */
        m->cc = (double *)xmalloc(nl * 3 * 8); m->vol = (double *)xmalloc(nl * 8);
        m->cellLevel = (i32 *)xmalloc(nl * 4); m->visible = (i32 *)xmalloc(nl * 4);
        for (i64 c = 0; c < n; c++) if (cellRank[c] == r) {
            i32 l = loc[c];
            m->cellGlobal[l] = (i32)c;
            m->cc[3 * l] = g->cc[3 * c]; m->cc[3 * l + 1] = g->cc[3 * c + 1]; m->cc[3 * l + 2] = g->cc[3 * c + 2];
            m->vol[l] = g->vol[c]; m->cellLevel[l] = g->cellLevel[c]; m->visible[l] = g->visible[c];
        }
        m->nsplit = g->nsplit;
        m->splitParent = (i32 *)xmalloc((g->nsplit ? g->nsplit : 1) * 4);
        memcpy(m->splitParent, g->splitParent, g->nsplit * 4);
        /* points */
        if (g->pcOff) {
            i32 *ploc = (i32 *)xmalloc(p * 4);
            i64 pl = 0;
            for (i64 q = 0; q < p; q++) {
                int used = 0;
                for (i32 i = g->pcOff[q]; i < g->pcOff[q + 1]; i++) if (cellRank[g->pcCells[i]] == r) { used = 1; break; }
                ploc[q] = used ? (i32)pl++ : -1;
            }
            m->p = pl;
            m->pointGlobal = (i32 *)xmalloc(pl * 4);
            m->pcOff = (i32 *)xcalloc(pl + 1, 4);
            m->bndPoint = (u8 *)xcalloc(pl, 1); m->pointLevel = (i32 *)xmalloc(pl * 4);
            if (g->pxyz) m->pxyz = (i32 *)xmalloc(pl * 3 * 4);
            i64 nnz = 0;
            for (i64 q = 0; q < p; q++) if (ploc[q] >= 0) {
                i32 l = ploc[q];
                m->pointGlobal[l] = (i32)q; m->bndPoint[l] = g->bndPoint[q]; m->pointLevel[l] = g->pointLevel[q];
                if (g->pxyz) { m->pxyz[3 * l] = g->pxyz[3 * q]; m->pxyz[3 * l + 1] = g->pxyz[3 * q + 1]; m->pxyz[3 * l + 2] = g->pxyz[3 * q + 2]; }
                for (i32 i = g->pcOff[q]; i < g->pcOff[q + 1]; i++) if (cellRank[g->pcCells[i]] == r) nnz++;
                m->pcOff[l + 1] = (i32)nnz;
            }
            m->pcCells = (i32 *)xmalloc((nnz ? nnz : 1) * 4);
            i64 k = 0;
            for (i64 q = 0; q < p; q++) if (ploc[q] >= 0)
                for (i32 i = g->pcOff[q]; i < g->pcOff[q + 1]; i++)
                    if (cellRank[g->pcCells[i]] == r) m->pcCells[k++] = loc[g->pcCells[i]];
            /* cellPoints transpose */
            m->cpOff = (i32 *)xcalloc(nl + 1, 4);
            for (i64 i = 0; i < nnz; i++) m->cpOff[m->pcCells[i] + 1]++;
            for (i64 c = 0; c < nl; c++) m->cpOff[c + 1] += m->cpOff[c];
            m->cpPts = (i32 *)xmalloc((nnz ? nnz : 1) * 4);
            { i32 *fill = (i32 *)xmalloc((nl + 1) * 4); memcpy(fill, m->cpOff, (nl + 1) * 4);
              for (i64 q = 0; q < pl; q++) for (i32 i = m->pcOff[q]; i < m->pcOff[q + 1]; i++) m->cpPts[fill[m->pcCells[i]]++] = (i32)q;
              free(fill); }
            /* points of cut faces are boundary points of the sub-domain:
               face points = cellPoints(owner) ^ cellPoints(neighbour) (rows are sorted) */
            for (i64 i = 0; i < f; i++) {
                i32 o_ = g->owner[i], n_ = g->neighbour[i];
                int ro = cellRank[o_], rn = cellRank[n_];
                if ((ro == r) == (rn == r)) continue;
                i32 a = g->cpOff[o_], ae = g->cpOff[o_ + 1], c2 = g->cpOff[n_], ce = g->cpOff[n_ + 1];
                while (a < ae && c2 < ce) {
                    i32 pa = g->cpPts[a], pc = g->cpPts[c2];
                    if (pa == pc) { if (ploc[pa] >= 0) m->bndPoint[ploc[pa]] = 1; a++; c2++; }
                    else if (pa < pc) a++; else c2++;
                }
            }
            /* faces(), edges() and the point lists of the processor patches (polyRefiner across ranks) */
            if (g->faceOff) {
                i64 nfl = fl + bl, tot = 0;
                m->faceOff = (i32 *)xcalloc(nfl + 1, 4);
                for (i64 k2 = 0; k2 < nfl; k2++) { i64 gf = m->faceGlobal[k2]; tot += g->faceOff[gf + 1] - g->faceOff[gf]; m->faceOff[k2 + 1] = (i32)tot; }
                m->facePts = (i32 *)xmalloc((tot ? tot : 1) * 4);
                for (i64 k2 = 0; k2 < nfl; k2++) {
                    i64 gf = m->faceGlobal[k2];
                    i32 *dst = m->facePts + m->faceOff[k2];
                    for (i32 j = g->faceOff[gf]; j < g->faceOff[gf + 1]; j++) *dst++ = ploc[g->facePts[j]];
                }
                edges_from_faces(m);
                /* syncTools::syncPointList pairing: the points of a processor patch in ascending GLOBAL label, the same
                   order on both sides */
                m->ppOff = (i32 *)xcalloc(np + 1, 4);
                i32 *tmpp = (i32 *)xmalloc((tot ? tot : 1) * 4);
                i64 ntot = 0;
                for (int q = 0; q < np; q++) {
                    i64 cnt = 0;
                    if (m->patchType[q] == 2) {
                        for (i64 k2 = m->patchStart[q]; k2 < m->patchStart[q] + m->patchSize[q]; k2++)
                            for (i32 j = m->faceOff[k2]; j < m->faceOff[k2 + 1]; j++) tmpp[ntot + cnt++] = m->pointGlobal[m->facePts[j]];
                        qsort(tmpp + ntot, (size_t)cnt, 4, cmp_i32);
                        i64 u = 0;
                        for (i64 i = 0; i < cnt; i++) if (i == 0 || tmpp[ntot + i] != tmpp[ntot + i - 1]) tmpp[ntot + u++] = tmpp[ntot + i];
                        cnt = u;
                    }
                    ntot += cnt;
                    m->ppOff[q + 1] = (i32)ntot;
                }
                m->ppPts = (i32 *)xmalloc((ntot ? ntot : 1) * 4);
                for (i64 i = 0; i < ntot; i++) m->ppPts[i] = ploc[tmpp[i]];
                free(tmpp);
            }
            free(ploc);
        }
        free(pcount); free(pstart); free(pfill);
    }
    free(loc); free(nloc);
    return 0;
}

MG_API int mg_decompose(const mg_mesh *g, const i32 *cellRank, int nranks, mg_mesh **out) {
    return decompose_impl(g, cellRank, nranks, out, -1);
}
/* the local mesh of ONE rank (what a rank of a decomposed run holds); same numbering as mg_decompose */
MG_API mg_mesh *mg_decompose_one(const mg_mesh *g, const i32 *cellRank, int nranks, int rank) {
    mg_mesh *out = NULL;
    if (rank < 0 || rank >= nranks) return NULL;
    decompose_impl(g, cellRank, nranks, &out, rank);
    return out;
}

/* "simple" decomposition of an octree mesh by cell-centre position: (px py pz) slabs */
MG_API void mg_simple_ranks(const mg_mesh *g, double lx, double ly, double lz, int px, int py, int pz, i32 *cellRank) {
    #pragma omp parallel for schedule(static)
    for (i64 c = 0; c < g->n; c++) {
        int i = (int)(g->cc[3 * c] / lx * px), j = (int)(g->cc[3 * c + 1] / ly * py), k = (int)(g->cc[3 * c + 2] / lz * pz);
        if (i >= px) i = px - 1; if (j >= py) j = py - 1; if (k >= pz) k = pz - 1;
        cellRank[c] = i + px * (j + py * k);
    }
}
