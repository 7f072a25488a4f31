"""Shared scenario builders for the parity tests (synthetic meshes + analytic fields)."""
from __future__ import annotations

import numpy as np

from blastamr_b200 import mesh as M


def front_field(cc, centre=(0.5, 0.5, 0.5), r0=0.3, w=0.02):
    """rho(x) = 1 + 0.5 (1 + tanh((|x-c| - r0)/w))  -- SURVEY.md 8(d) config C5"""
    d = np.sqrt(((cc - np.asarray(centre)) ** 2).sum(axis=1))
    return 1.0 + 0.5 * (1.0 + np.tanh((d - r0) / w))


def box_field(cc, lo=(0.02, 0.025, -1.0), hi=(0.04, 0.035, 1.0)):
    """the reference test recipe: 1 inside a box, 0 outside (adaptiveFvMeshTest.C:156-170)"""
    inside = np.all((cc > np.asarray(lo)) & (cc < np.asarray(hi)), axis=1)
    return inside.astype(np.float64)


def graded_octree(n0=8, max_level=2, centre=(0.5, 0.5, 0.5), r0=0.3, band=0.08, sides=None, patch_types=None,
                  lengths=(1.0, 1.0, 1.0)):
    """Octree refined towards a spherical shell, level by level, with a crude 2:1 closure done by
    repeated flag growth on the generator side (NOT the code under test)."""
    oc = M.Octree(n0, n0, n0, lengths, max_level, sides, patch_types)
    for lev in range(max_level):
        m = oc.build(points=False)
        d = np.sqrt(((m.cc - np.asarray(centre)) ** 2).sum(axis=1))
        want = (np.abs(d - r0) < band / (lev + 1)) & (m.cell_level == lev)
        flags = want.copy()
        # closure: a face neighbour more than one level coarser than (level+1) must refine too
        own, nei = m.owner[:m.f], m.neighbour
        while True:
            lv = m.cell_level + flags
            a = lv[own] > lv[nei] + 1
            b = lv[nei] > lv[own] + 1
            if not a.any() and not b.any():
                break
            flags[nei[a]] = True
            flags[own[b]] = True
        oc.refine(np.flatnonzero(flags).astype(np.int32))
    return oc


def poly_balanced_octree(O, n0=10, max_level=2):
    """an octree that is 2:1 balanced over faces AND edges (what polyRefiner produces and requires): grown from a
    uniform box by the oracle's own polyRefiner selection around a spherical front"""
    oc = M.Octree(n0, n0, n0, (1.0, 1.0, 1.0), max_level)
    for _ in range(max_level):
        m = oc.build(faces=True)
        w = O.World([m])
        e = w.estimate("delta", [front_field(m.cc, centre=(0.5, 0.5, 0.5), r0=0.3, w=0.03)])
        w.normalize(e, 0.01, 1e15, 0.01, 1e15)
        _, rs, _ = w.poly_select_refine(e, max_level, 2)
        oc.refine(np.flatnonzero(rs[0]).astype(np.int32))
    return oc


class OracleBackend:
    """Same call surface as blastamr_b200.capi.Context, answered by the oracle (R = 1).
    Used on CPU only to exercise the test drivers themselves; never by the product."""

    def __init__(self):
        from oracle import oracle as O
        self.O = O
        self.w = None
        self.prot = None

    def set_mesh(self, m):
        self.m = m
        self.w = self.O.World([m])
        self.n, self.p, self.f, self.b = m.n, m.p, m.f, m.b

    def set_protected_cells(self, prot):
        self.prot = prot

    def estimate(self, kind, x, p0=1e-15, p1=0.0, thresholds=None, **_):
        e = self.w.estimate(kind, [x], min_val=p0, offset=p1)
        if thresholds is not None:
            self.w.normalize(e, *thresholds)
        return e[0]

    def normalize(self, err, *thr):
        self.w.normalize([err], *thr)
        return err

    def protect_patches(self, ids, n_layers, err):
        return int(self.w.protect_patches([err], ids, n_layers)[0])

    def select_refine(self, err, max_level, n_layers, protected_patch_ids=(), max_cells=2147483647, refine_cell_out=None, kind=0, **_):
        if err is None:
            err = self.err
        if not np.isscalar(max_level):
            max_level = [max_level]
        if kind in (2, 3):
            rc, rs, sw = self.w.poly_select_refine([err], max_level, n_layers, protected_patch_ids=protected_patch_ids)
            if refine_cell_out is not None:
                refine_cell_out[:] = rc[0]
            return np.flatnonzero(rs[0]).astype(np.int32), sw
        rc, rs, sw = self.w.hex_select_refine([err], max_level, n_layers, protected_cell=None if self.prot is None else [self.prot],
                                              protected_patch_ids=protected_patch_ids, max_cells=max_cells)
        if refine_cell_out is not None:
            refine_cell_out[:] = rc[0]
        return np.flatnonzero(rs[0]).astype(np.int32), sw

    def consistent_refinement(self, cells, max_set=True):
        f = np.zeros(self.n, np.uint8)
        f[cells] = 1
        sw = self.w.consistent_refinement([f], max_set)
        return np.flatnonzero(f).astype(np.int32), sw

    def remap_refine_cell(self, cell_map, rev, refine_cell, out):
        out[:] = self.O.remap_refine_cell(cell_map, rev, refine_cell)
        return out

    def map_field(self, cell_map, old, ncomp=1, mode=0, reverse_map=None, vol_old=None, **_):
        if mode == 1:
            return self.O.map_field_conservative(cell_map, reverse_map, vol_old, old, ncomp)
        return self.O.map_field(cell_map, old, ncomp)

    def update_levels(self, cell_map, old_level, refined_old=None, unrefined_old=None):
        return self.O.update_levels(cell_map, old_level, refined_old, unrefined_old)

    def check_levels(self, lv=None):
        return self.w.check_levels([self.m.cell_level if lv is None else lv])

    def split_elems(self, kind=0):
        if kind == 1:
            return np.flatnonzero(self.w.split_edges(0)).astype(np.int32)
        if kind == 2:
            return np.flatnonzero(self.w.poly_split_points(0)).astype(np.int32)
        if kind == 3:
            return np.flatnonzero(self.w.prism_split_points(0)).astype(np.int32)
        return np.flatnonzero(self.w.split_points(0)).astype(np.int32)

    def select_unrefine(self, err, n_layers, refine_cell=None, protected_patch_ids=(), kind=0, **_):
        if kind in (2, 3):
            up, sp, sw = self.w.poly_select_unrefine([err], [refine_cell], n_layers, protected_patch_ids=protected_patch_ids,
                                                     prismatic=(kind == 3))
            return np.flatnonzero(up[0]).astype(np.int32), sw
        fn = self.w.hex2d_select_unrefine if kind == 1 else self.w.hex_select_unrefine
        up, sp, sw = fn([err], [refine_cell], n_layers, protected_cell=None if self.prot is None else [self.prot],
                                                protected_patch_ids=protected_patch_ids)
        return np.flatnonzero(up[0]).astype(np.int32), sw

    def max_cell_field(self, v):
        return self.w.max_cell_field(0, v)

    # rows either side of the path
    def reorder_labels(self, old_to_new, n_new, old_values):
        return self.O.reorder_labels(old_to_new, n_new, old_values)

    def gather_labels(self, new_to_old, old_values):
        return self.O.gather_labels(new_to_old, old_values)

    def remap_protected_cells(self, cell_map, old_protected):
        return self.O.remap_protected_cells(cell_map, old_protected)

    def correct_fluxes(self, face_map, reverse_face_map, U, phi, U_boundary=None, nbr_U=None):
        k = self.w.correct_fluxes(0, face_map, reverse_face_map, U, phi, U_boundary=U_boundary, nbr_U=nbr_U)
        if k < 0:
            from blastamr_b200.capi import AmrError
            raise AmrError(-5, "Problem: should not have removed faces when refining.")
        return k

    def set_ls_vectors(self, p_vectors, n_vectors):
        self.ls = (p_vectors, n_vectors)

    def estimate_gradient(self, x, scheme=0, x_boundary=None, nbr_values=None, total_volume=0.0, refine_threshold=0.5,
                          unrefine_threshold=0.4, scale=True, want_grad=False):
        kw = dict(ls_p=self.ls[0], ls_n=self.ls[1]) if scheme == 1 else {}
        mag, g = self.w.gradient_mag(0, x, scheme=scheme, xb=x_boundary, xn=nbr_values, total_volume=total_volume, **kw)
        thr, _ = self.w.range_thresholds([mag], refine_threshold, unrefine_threshold)
        if scale:
            self.w.normalize([mag], *thr)
        self.err = mag
        return mag, np.asarray(thr), (g if want_grad else None)

    def mesh_size(self, geometric_d, want_dx=True, want_dX=True):
        return self.w.mesh_dx(0, geometric_d), self.w.mesh_dX(0)

    def max_refinement(self, max_level, min_dx=0.0, dx=None, error=None):
        return self.w.max_refinement(0, max_level, min_dx, dx, error)

    def estimate_cell_size(self, size_type, geometric_d, **kw):
        return self.w.cell_size_error(0, size_type, geometric_d, **kw)

    def imbalance(self):
        return self.w.imbalance()

    def history_clusters(self, clusters_in=None, want_blocked=True):
        if clusters_in is None:
            c2c, ncl = self.w.history_clusters(0)
        else:
            c2c, ncl = clusters_in, -1
        bl, nu = self.w.blocked_faces(0, c2c)
        return c2c, bl, ncl, nu

    def constrain_decomposition(self, decomposition):
        c2c, ncl = self.w.history_clusters(0)
        return self.w.constrain_decomposition(0, c2c, ncl, decomposition)

    def proc_load(self, distribution, n_procs):
        if len(distribution) and (distribution.min() < 0 or distribution.max() >= n_procs):
            from blastamr_b200.capi import AmrError
            raise AmrError(-2, "amr_proc_load: distribution holds a processor outside [0, nProcs)")
        return self.O.proc_load(distribution, n_procs)

    def multicomponent(self, errors, max_cell_levels=None):
        e, mr = self.w.multicomponent(0, errors, max_cell_levels)
        self.err = e
        return e, mr, 3


# ---- helpers for the rows either side of the path (flux fix-up, gradients) ---------------------------------

def _face_dir(Sf):
    """0..5 = -x,+x,-y,+y,-z,+z for axis-aligned faces"""
    ax = np.argmax(np.abs(Sf), axis=1)
    return 2 * ax + (Sf[np.arange(len(Sf)), ax] > 0)


def face_maps(m_old, m_new, cell_map):
    """mapPolyMesh faceMap / reverseFaceMap of a REFINEMENT step, reconstructed from the two meshes (the
    topology engine rebuilds faces instead of tracking them): a new face descends from the old face between
    the old cells of its two sides (or on the same patch side of its owner's old cell); faces inside a split
    cell are inflated (-1).  The master of an old face is its lowest-numbered descendant."""
    own_o, nei_o = m_old.owner, m_old.neighbour
    key_int = {(int(own_o[i]), int(nei_o[i])): i for i in range(m_old.f)}
    dir_o = _face_dir(m_old.Sf.reshape(-1, 3)[m_old.f:])
    key_bnd = {(int(own_o[m_old.f + i]), int(dir_o[i])): m_old.f + i for i in range(m_old.b)}
    nf_new = m_new.f + m_new.b
    face_map = np.full(nf_new, -1, np.int32)
    dir_n = _face_dir(m_new.Sf.reshape(-1, 3)[m_new.f:])
    for i in range(m_new.f):
        a, b = int(cell_map[m_new.owner[i]]), int(cell_map[m_new.neighbour[i]])
        if a != b:
            face_map[i] = key_int[(min(a, b), max(a, b))]
    for i in range(m_new.b):
        face_map[m_new.f + i] = key_bnd[(int(cell_map[m_new.owner[m_new.f + i]]), int(dir_n[i]))]
    rev = np.full(m_old.f + m_old.b, -1, np.int32)
    for i in range(nf_new - 1, -1, -1):
        if face_map[i] >= 0:
            rev[face_map[i]] = i
    return face_map, rev


def ls_vectors(m, cc_nbr=None):
    """leastSquaresVectors (OpenFOAM-v2212 leastSquaresVectors.C, restated in numpy as TEST INPUT: the library
    takes these vectors from the host, so parity only needs both sides to see the same numbers)."""
    n, f, b = m.n, m.f, m.b
    C, w = m.cc, m.weights
    magSf = m.magSf
    Sf = m.Sf.reshape(-1, 3)
    own, nei = m.owner, m.neighbour
    dd = np.zeros((n, 3, 3))
    d = C[nei] - C[own[:f]]
    wdd = (magSf[:f] / (d * d).sum(1))[:, None, None] * d[:, :, None] * d[:, None, :]
    np.add.at(dd, own[:f], (1 - w[:f])[:, None, None] * wdd)
    np.add.at(dd, nei, w[:f][:, None, None] * wdd)
    # boundary d-vectors: face centre - cell centre for plain patches (n * 1/deltaCoeffs on these box meshes),
    # neighbour centre - cell centre on processor patches
    nhat = Sf[f:] / magSf[f:, None]
    pd = nhat / m.delta_coeffs[f:, None]
    st, sz, ty, nb = m.patch_arrays()
    coupled = np.zeros(b, bool)
    for q in range(len(st)):
        if ty[q] == 2:
            coupled[st[q] - f:st[q] - f + sz[q]] = True
    if cc_nbr is not None:
        pd[coupled] = (cc_nbr - C[own[f:]])[coupled]
    fac = np.where(coupled, 1 - w[f:], 1.0) * magSf[f:] / (pd * pd).sum(1)
    np.add.at(dd, own[f:], fac[:, None, None] * pd[:, :, None] * pd[:, None, :])
    # 2-D meshes: the empty direction carries no information; regularise like inv(symmTensorField) does
    for k in range(3):
        if np.all(np.abs(dd[:, k, k]) < 1e-300):
            dd[:, k, k] = 1.0
    inv = np.linalg.inv(dd)
    s = magSf[:f] / (d * d).sum(1)
    ls_p = np.zeros((f + b, 3))
    ls_p[:f] = ((1 - w[:f]) * s)[:, None] * np.einsum("fij,fj->fi", inv[own[:f]], d)
    ls_n = (-w[:f] * s)[:, None] * np.einsum("fij,fj->fi", inv[nei], d)
    ls_p[f:] = fac[:, None] * np.einsum("fij,fj->fi", inv[own[f:]], pd)
    return np.ascontiguousarray(ls_p), np.ascontiguousarray(ls_n)


def unrefinement_maps(n0=8, maxlev=2):
    """a real merge of the host topology engine: (mesh before, cellMap, reverseCellMap) with several merged groups"""
    from oracle import oracle as O
    oc = graded_octree(n0, maxlev)
    m = oc.build()
    w = O.World([m])
    up, _, _ = w.hex_select_unrefine([np.full(m.n, -1.0)], [np.zeros(m.n, np.uint8)], 0)
    pts = np.flatnonzero(up[0]).astype(np.int32)
    assert len(pts) > 3
    cm, rev = oc.unrefine(m, pts)
    return m, cm, rev


def synthetic_merge_maps(n_old=5000, seed=3, max_group=12):
    """general merges (groups of 2..max_group old cells with arbitrary labels into their lowest member)"""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n_old)
    master_of = np.arange(n_old)
    i = 0
    while i < n_old // 2:
        g = int(rng.integers(2, max_group + 1))
        grp = perm[i:i + g]
        master_of[grp] = grp.min()
        i += g
    keep = master_of == np.arange(n_old)
    cm = np.flatnonzero(keep).astype(np.int32)
    new_id = np.cumsum(keep) - 1
    rev = np.where(keep, new_id, -new_id[master_of] - 2).astype(np.int32)
    return cm, rev
