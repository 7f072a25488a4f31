// capi_host.cpp -- extern "C" face of the host layer (include/amr_host.h)
#include <cstring>
#include <string>

#include "amrHost.H"
#include "amr_host.h"

using namespace amr;

struct amrhost : topoChanger {
    std::unique_ptr<adaptiveFvMesh> mesh;
    std::string err;
    amrhost_topo_cb refineCb = nullptr, unrefineCb = nullptr;
    void *user = nullptr;
    mapPolyMesh pendingMap;
    bool haveMap = false;

    mapPolyMesh run(amrhost_topo_cb cb, const labelList &elems, const char *what) {
        if (!cb) throw FatalError(std::string("no ") + what + " callback attached");
        haveMap = false;
        int rc = cb(user, elems.data(), (int64_t)elems.size());
        if (rc != 0) throw FatalError(std::string(what) + " callback failed");
        if (!haveMap) throw FatalError(std::string(what) + " callback did not install a map (amrhost_set_map)");
        return pendingMap;
    }
    mapPolyMesh refine(fvMeshLite &, const labelList &cells) override { return run(refineCb, cells, "refine"); }
    mapPolyMesh unrefine(fvMeshLite &, const labelList &elems) override { return run(unrefineCb, elems, "unrefine"); }
};

static std::string g_createError;

#define GUARD(body)                                              \
    try { body; return 0; }                                      \
    catch (const std::exception &e) { h->err = e.what(); return -1; }

extern "C" {
#define API __attribute__((visibility("default")))

API int amrhost_create(const char *text, int device, amrhost **out) {
    if (!out || !text) return -2;
    *out = nullptr;
    try {
        amrhost *h = new amrhost();
        h->mesh.reset(new adaptiveFvMesh(text, device));
        h->mesh->setTopoChanger(h);
        *out = h;
        return 0;
    } catch (const std::exception &e) { g_createError = e.what(); return -1; }
}
API void amrhost_destroy(amrhost *h) { delete h; }
API const char *amrhost_last_error(const amrhost *h) { return h ? h->err.c_str() : g_createError.c_str(); }
API int amrhost_set_dict(amrhost *h, const char *text) { GUARD(h->mesh->setDict(text)) }
API int amrhost_set_topo_callbacks(amrhost *h, amrhost_topo_cb r, amrhost_topo_cb u, void *user) {
    h->refineCb = r; h->unrefineCb = u; h->user = user; return 0;
}
API int amrhost_set_mesh(amrhost *h, int64_t n, int64_t f, int64_t b, int64_t p, const int32_t *owner, const int32_t *neighbour, int npatch,
                         const char *const *names, const int32_t *st, const int32_t *sz, const int32_t *ty, const int32_t *nb, int nSolutionD) {
    GUARD({
        fvMeshLite &m = h->mesh->mesh();
        m.nCells = n; m.nInternalFaces = f; m.nBoundaryFaces = b; m.nPoints = p; m.nSolutionD = nSolutionD;
        m.owner.assign(owner, owner + f + b); m.neighbour.assign(neighbour, neighbour + f);
        m.patches.clear();
        for (int i = 0; i < npatch; i++) m.patches.push_back({names[i], ty[i], st[i], sz[i], nb ? nb[i] : -1});
        m.pcOff.clear(); m.pcCells.clear(); m.bndPoint.clear();
        m.nEdges = 0; m.edgePts.clear(); m.ecOff.clear(); m.ecCells.clear(); m.bndEdge.clear();
        m.Sf.clear(); m.weights.clear(); m.magSf.clear(); m.deltaCoeffs.clear(); m.V.clear();
        m.faceOff.clear(); m.facePts.clear();
    })
}
API int amrhost_set_points(amrhost *h, const int32_t *off, const int32_t *cells, const uint8_t *bnd) {
    GUARD({
        fvMeshLite &m = h->mesh->mesh();
        m.pcOff.assign(off, off + m.nPoints + 1); m.pcCells.assign(cells, cells + off[m.nPoints]); m.bndPoint.assign(bnd, bnd + m.nPoints);
    })
}
API int amrhost_set_edges(amrhost *h, int64_t ne, const int32_t *ep, const int32_t *off, const int32_t *cells, const uint8_t *bnd) {
    GUARD({
        fvMeshLite &m = h->mesh->mesh();
        m.nEdges = ne; m.edgePts.assign(ep, ep + 2 * ne); m.ecOff.assign(off, off + ne + 1); m.ecCells.assign(cells, cells + off[ne]);
        m.bndEdge.assign(bnd, bnd + ne);
    })
}
API int amrhost_set_faces(amrhost *h, const int32_t *off, const int32_t *pts) {
    GUARD({
        fvMeshLite &m = h->mesh->mesh();
        const int64_t nf = m.nInternalFaces + m.nBoundaryFaces;
        m.faceOff.assign(off, off + nf + 1); m.facePts.assign(pts, pts + off[nf]);
    })
}
API int amrhost_set_geometry(amrhost *h, const double *w, const double *Sf, const double *magSf, const double *dc, const double *V) {
    GUARD({
        fvMeshLite &m = h->mesh->mesh();
        const int64_t nf = m.nInternalFaces + m.nBoundaryFaces;
        m.weights.assign(w, w + nf); m.Sf.assign(Sf, Sf + 3 * nf); m.magSf.assign(magSf, magSf + nf); m.deltaCoeffs.assign(dc, dc + nf);
        m.V.assign(V, V + m.nCells);
    })
}
API int amrhost_set_levels(amrhost *h, const int32_t *cl, const int32_t *pl, const int32_t *vis, const int32_t *sp, int64_t nSplit) {
    GUARD({
        fvMeshLite &m = h->mesh->mesh();
        m.cellLevel.assign(cl, cl + m.nCells);
        if (pl) m.pointLevel.assign(pl, pl + m.nPoints); else m.pointLevel.clear();
        if (vis) m.visibleCells.assign(vis, vis + m.nCells); else m.visibleCells.clear();
        if (sp && nSplit > 0) m.splitParent.assign(sp, sp + nSplit); else m.splitParent.clear();
    })
}
API int amrhost_set_map(amrhost *h, int64_t nOld, int64_t nNew, const int32_t *cellMap, const int32_t *rev) {
    GUARD({
        h->pendingMap.cellMap.assign(cellMap, cellMap + nNew);
        h->pendingMap.reverseCellMap.assign(rev, rev + nOld);
        h->haveMap = true;
    })
}
API int amrhost_mesh_changed(amrhost *h) { GUARD(h->mesh->meshChanged()) }
API int amrhost_set_field(amrhost *h, const char *name, int nComp, const double *v, int64_t nCells) {
    GUARD({
        volField &f = h->mesh->mesh().fields[name];
        f.nComp = nComp; f.v.assign(v, v + nCells * nComp);
    })
}
API int amrhost_get_field(amrhost *h, const char *name, double *v, int64_t cap, int64_t *nv) {
    GUARD({
        volField &f = h->mesh->mesh().lookupField(name);
        if (nv) *nv = (int64_t)f.v.size();
        if (v) { if ((int64_t)f.v.size() > cap) throw FatalError("amrhost_get_field: buffer too small"); std::memcpy(v, f.v.data(), f.v.size() * 8); }
    })
}
API int amrhost_set_time(amrhost *h, int32_t ti, double tv) { h->mesh->mesh().timeIndex = ti; h->mesh->mesh().timeValue = tv; return 0; }
API int amrhost_update(amrhost *h, int *changed) { GUARD({ bool c = h->mesh->update(); if (changed) *changed = c ? 1 : 0; }) }
API int64_t amrhost_n_cells(const amrhost *h) { return h->mesh->mesh().nCells; }
API int amrhost_last_lists(amrhost *h, int32_t *cr, int64_t *nr, int32_t *eu, int64_t *nu) {
    GUARD({
        fvMeshHexRefiner *r = dynamic_cast<fvMeshHexRefiner *>(&h->mesh->refiner());
        if (!r) throw FatalError("refiner is not a hexRefiner");
        if (nr) *nr = (int64_t)r->lastCellsToRefine.size();
        if (nu) *nu = (int64_t)r->lastElemsToUnrefine.size();
        if (cr) std::memcpy(cr, r->lastCellsToRefine.data(), r->lastCellsToRefine.size() * 4);
        if (eu) std::memcpy(eu, r->lastElemsToUnrefine.data(), r->lastElemsToUnrefine.size() * 4);
    })
}
API int amrhost_get_error(amrhost *h, double *e, int64_t cap) {
    GUARD({
        scalarField &er = h->mesh->error().error();
        if ((int64_t)er.size() > cap) throw FatalError("amrhost_get_error: buffer too small");
        std::memcpy(e, er.data(), er.size() * 8);
    })
}
static int joinTo(const std::vector<std::string> &v, char *out, int64_t cap) {
    std::string s;
    for (auto &w : v) { if (!s.empty()) s += " "; s += w; }
    if ((int64_t)s.size() + 1 > cap) return -2;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return 0;
}
API int amrhost_parse_dict(const char *text, char *out, int64_t cap) {
    try { return joinTo(dictionary::parse(text).optionalSubDict("adaptiveFvMeshCoeffs").toc(), out, cap); }
    catch (const std::exception &e) { g_createError = e.what(); return -1; }
}
API int amrhost_selection_table(int which, char *out, int64_t cap) {
    std::vector<std::string> v;
    if (which == 0) for (auto &e : errorEstimator::dictionaryConstructorTable()) v.push_back(e.first);
    else for (auto &e : fvMeshRefiner::dictionaryConstructorTable()) v.push_back(e.first);
    return joinTo(v, out, cap);
}
}
