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
        self.n, self.p = m.n, m.p

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
