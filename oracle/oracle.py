"""ctypes binding of the CPU oracle (oracle/amr_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
_lib = None

SMALL = 1e-15
GREAT = 1e15
SQRT_SMALL = float(np.sqrt(1e-15))

KIND = {"delta": 0, "scaledDelta": 1, "Lohner": 2, "fieldValue": 3, "codedDelta": 4, "densityGradient": 1}


def build(force: bool = False) -> Path:
    src = sorted(HERE.glob("*.c"))
    if not force and LIB.exists() and all(s.stat().st_mtime <= LIB.stat().st_mtime for s in src):
        return LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-std=gnu11", "-pthread", "-o", str(LIB)] + \
          [str(s) for s in src] + ["-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("oracle build failed")
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        L.or_world_create.restype = C.c_void_p
        L.or_world_create.argtypes = [C.c_int]
        L.or_world_free.argtypes = [C.c_void_p]
        L.or_world_set_mesh.argtypes = [C.c_void_p, C.c_int] + [C.c_int64] * 4 + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 4
        L.or_world_set_points.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.or_world_set_levels.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.or_normalize.argtypes = [C.c_int64, C.c_void_p] + [C.c_double] * 4
        L.or_world_set_edges.argtypes = [C.c_void_p, C.c_int, C.c_int64] + [C.c_void_p] * 4
        L.or_get_split_edges.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.or_hex2d_select_unrefine.restype = C.c_int
        L.or_hex2d_select_unrefine.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_world_set_faces.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.or_poly_consistent_refinement.restype = C.c_int
        L.or_poly_consistent_refinement.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.or_poly_select_refine.restype = C.c_int
        L.or_poly_select_refine.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                            C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.or_poly_split_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.or_poly_select_unrefine.restype = C.c_int
        L.or_poly_select_unrefine.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                              C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.or_prism_split_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.or_world_set_geometry.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.or_estimate_lohner.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.or_range_thresholds.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.or_normalize_world.argtypes = [C.c_void_p, C.c_void_p] + [C.c_double] * 4
        L.or_map_field_world.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.or_estimate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.or_protect_patches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.or_consistent_refinement.restype = C.c_int
        L.or_consistent_refinement.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.or_hex_select_refine.restype = C.c_int
        L.or_hex_select_refine.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.or_remap_refine_cell.argtypes = [C.c_int64] + [C.c_void_p] * 4
        L.or_max_cell_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.or_get_split_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.or_consistent_unrefinement.restype = C.c_int
        L.or_consistent_unrefinement.argtypes = [C.c_void_p, C.c_void_p]
        L.or_hex_select_unrefine.restype = C.c_int
        L.or_hex_select_unrefine.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_check_levels.restype = C.c_int64
        L.or_check_levels.argtypes = [C.c_void_p, C.c_void_p]
        L.or_map_field.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.or_map_field_conservative.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                C.c_void_p, C.c_void_p]
        L.or_map_old_volumes.restype = C.c_int64
        L.or_map_old_volumes.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 4
        L.or_update_levels.argtypes = [C.c_int64] + [C.c_void_p] * 5
        L.or_reorder_labels.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_map_labels.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_map_protected.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_correct_fluxes.restype = C.c_int64
        L.or_correct_fluxes.argtypes = [C.c_void_p, C.c_int, C.c_int64] + [C.c_void_p] * 6
        L.or_gradient_mag.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_double, C.c_void_p, C.c_void_p]
        L.or_mesh_dx.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.or_mesh_dX.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.or_max_refinement.argtypes = [C.c_void_p, C.c_int, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_cell_size_error.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_multicomponent_error.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        L.or_multicomponent_max_refinement.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_history_clusters.restype = C.c_int32
        L.or_history_clusters.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p]
        L.or_blocked_faces.restype = C.c_int64
        L.or_blocked_faces.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.or_constrain_decomposition.restype = C.c_int64
        L.or_constrain_decomposition.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int32, C.c_void_p]
        L.or_imbalance.restype = C.c_double
        L.or_imbalance.argtypes = [C.c_void_p, C.c_void_p]
        L.or_proc_load.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
        L.or_set_threads.argtypes = [C.c_int]
        L.or_max_threads.restype = C.c_int
        L.or_set_threads(1)
        _lib = L
    return _lib


def _pp(arrs: Sequence[Optional[np.ndarray]]):
    """array of pointers (one per rank); None -> NULL"""
    return (C.c_void_p * len(arrs))(*[None if a is None else a.ctypes.data for a in arrs])


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class World:
    """R emulated MPI ranks (R=1: the serial reference run) over PolyMesh objects."""

    def __init__(self, meshes):
        L = lib()
        self.meshes = list(meshes)
        self.R = len(self.meshes)
        self._h = L.or_world_create(self.R)
        self._keep = []
        for r, m in enumerate(self.meshes):
            st, sz, ty, nb = m.patch_arrays()
            own, nei = _i32(m.owner), _i32(m.neighbour)
            self._keep += [st, sz, ty, nb, own, nei]
            L.or_world_set_mesh(self._h, r, m.n, m.f, m.b, m.p, own.ctypes.data, nei.ctypes.data, len(st),
                                st.ctypes.data, sz.ctypes.data, ty.ctypes.data, nb.ctypes.data)
            if m.pc_off is not None:
                a = [_i32(m.pc_off), _i32(m.pc_cells), _i32(m.cp_off), _i32(m.cp_pts),
                     np.ascontiguousarray(m.bnd_point, dtype=np.uint8)]
                self._keep += a
                L.or_world_set_points(self._h, r, *[x.ctypes.data for x in a])
            if getattr(m, "n_edges", 0):
                a = [_i32(m.edge_pts), _i32(m.ec_off), _i32(m.ec_cells), np.ascontiguousarray(m.bnd_edge, dtype=np.uint8)]
                self._keep += a
                L.or_world_set_edges(self._h, r, m.n_edges, *[v.ctypes.data for v in a])
            if getattr(m, "face_off", None) is not None:
                a = [_i32(m.face_off), _i32(m.face_pts)]
                self._keep += a
                L.or_world_set_faces(self._h, r, a[0].ctypes.data, a[1].ctypes.data)
            if getattr(m, "Sf", None) is not None and m.vol is not None:
                a = [np.ascontiguousarray(v, dtype=np.float64) for v in (m.weights, m.Sf, m.magSf, m.delta_coeffs, m.vol)]
                self._keep += a
                L.or_world_set_geometry(self._h, r, *[v.ctypes.data for v in a])
            if m.cell_level is not None:
                sp = _i32(m.split_parent) if m.split_parent is not None and len(m.split_parent) else np.full(1, -1, np.int32)
                a = [_i32(m.cell_level), _i32(m.point_level) if m.point_level is not None else np.zeros(max(m.p, 1), np.int32),
                     _i32(m.visible) if m.visible is not None else np.full(max(m.n, 1), -1, np.int32), sp]
                self._keep += a
                L.or_world_set_levels(self._h, r, *[x.ctypes.data for x in a])

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.or_world_free(h)

    # ---- stage 1
    def estimate(self, kind: str, x: List[np.ndarray], min_val: float = SMALL, offset: float = 0.0):
        x = [np.ascontiguousarray(a, dtype=np.float64) for a in x]
        err = [np.empty(m.n, dtype=np.float64) for m in self.meshes]
        lib().or_estimate(self._h, KIND[kind], _pp(x), _pp(err), min_val, offset)
        return err

    def estimate_lohner(self, x: List[np.ndarray], epsilon: float, xb=None):
        """Lohner.C:71-159 (needs meshes built with geometry=True); xb = boundaryField values or None (zeroGradient)"""
        x = [np.ascontiguousarray(a, dtype=np.float64) for a in x]
        err = [np.empty(m.n, dtype=np.float64) for m in self.meshes]
        lib().or_estimate_lohner(self._h, _pp(x), None if xb is None else _pp(xb), epsilon, _pp(err))
        return err

    @staticmethod
    def normalize(err: List[np.ndarray], lower_refine, upper_refine=GREAT, lower_unrefine=None, upper_unrefine=GREAT):
        for e in err:
            lib().or_normalize(len(e), e.ctypes.data, lower_refine, upper_refine, lower_unrefine, upper_unrefine)
        return err

    def range_thresholds(self, err, refine_threshold=0.5, unrefine_threshold=0.4):
        thr = np.zeros(4)
        mm = np.zeros(2)
        lib().or_range_thresholds(self._h, _pp(err), refine_threshold, unrefine_threshold, thr.ctypes.data, mm.ctypes.data)
        return thr, mm

    def normalize_world(self, err, lower_refine, upper_refine, lower_unrefine, upper_unrefine):
        """same, all ranks at once (rank-parallel under set_threads)"""
        lib().or_normalize_world(self._h, _pp(err), lower_refine, upper_refine, lower_unrefine, upper_unrefine)
        return err

    def map_field_world(self, cell_maps, olds, news, ncomp):
        nn = np.array([len(c) for c in cell_maps], dtype=np.int64)
        lib().or_map_field_world(self._h, nn.ctypes.data, _pp(cell_maps), ncomp, _pp(olds), _pp(news))
        return news

    def protect_patches(self, err: List[np.ndarray], patch_ids, n_layers: int):
        ids = _i32(patch_ids)
        cnt = np.zeros(self.R, dtype=np.int64)
        lib().or_protect_patches(self._h, ids.ctypes.data, len(ids), n_layers, _pp(err), cnt.ctypes.data)
        return cnt

    # ---- stage 2/3
    def consistent_refinement(self, flags: List[np.ndarray], max_set: bool = True) -> int:
        return lib().or_consistent_refinement(self._h, 1 if max_set else 0, _pp(flags))

    def hex_select_refine(self, err, max_cell_level, n_layers, lower=SQRT_SMALL, upper=GREAT, protected_cell=None,
                          protected_patch_ids=(), max_cells=2147483647):
        mcl = [np.full(m.n, max_cell_level, dtype=np.int32) if np.isscalar(max_cell_level) else _i32(max_cell_level[r])
               for r, m in enumerate(self.meshes)]
        rc = [np.zeros(m.n, dtype=np.uint8) for m in self.meshes]
        rs = [np.zeros(m.n, dtype=np.uint8) for m in self.meshes]
        pc = protected_cell if protected_cell is not None else [None] * self.R
        ids = _i32(protected_patch_ids)
        sweeps = lib().or_hex_select_refine(self._h, _pp(err), lower, upper, _pp(mcl), n_layers, _pp(pc), ids.ctypes.data,
                                            len(ids), max_cells, _pp(rc), _pp(rs))
        return rc, rs, sweeps

    def hex_select_unrefine(self, err, refine_cell, n_layers, unrefine_level=-SQRT_SMALL, protected_cell=None,
                            protected_patch_ids=()):
        up = [np.zeros(m.p, dtype=np.uint8) for m in self.meshes]
        sp = [np.zeros(m.p, dtype=np.uint8) for m in self.meshes]
        pc = protected_cell if protected_cell is not None else [None] * self.R
        ids = _i32(protected_patch_ids)
        sweeps = lib().or_hex_select_unrefine(self._h, _pp(err), unrefine_level, n_layers, _pp(pc), ids.ctypes.data, len(ids),
                                              _pp(refine_cell), _pp(up), _pp(sp))
        if sweeps < 0:
            raise RuntimeError("reference FatalError 'problem' (hexRef3D.C:1811)")
        return up, sp, sweeps

    # ---- polyRefiner (3-D)
    def poly_consistent_refinement(self, flags, max_set=True, edge_based=True) -> int:
        return lib().or_poly_consistent_refinement(self._h, 1 if max_set else 0, 1 if edge_based else 0, _pp(flags))

    def poly_select_refine(self, err, max_cell_level, n_layers, lower=SQRT_SMALL, upper=GREAT, force=False,
                           protected_patch_ids=(), edge_based=True):
        mcl = [np.full(m.n, max_cell_level, dtype=np.int32) if np.isscalar(max_cell_level) else _i32(max_cell_level[r])
               for r, m in enumerate(self.meshes)]
        rc = [np.zeros(m.n, dtype=np.uint8) for m in self.meshes]
        rs = [np.zeros(m.n, dtype=np.uint8) for m in self.meshes]
        ids = _i32(protected_patch_ids)
        sweeps = lib().or_poly_select_refine(self._h, _pp(err), lower, upper, _pp(mcl), n_layers, 1 if force else 0, ids.ctypes.data,
                                             len(ids), 1 if edge_based else 0, _pp(rc), _pp(rs))
        return rc, rs, sweeps

    def poly_split_points(self, r):
        out = np.zeros(self.meshes[r].p, dtype=np.uint8)
        lib().or_poly_split_points(self._h, r, out.ctypes.data)
        return out

    def prism_split_points(self, r):
        """prismatic2DRefinement split points (pointEdges rule)"""
        out = np.zeros(self.meshes[r].p, dtype=np.uint8)
        lib().or_prism_split_points(self._h, r, out.ctypes.data)
        return out

    def poly_select_unrefine(self, err, refine_cell, n_layers, unrefine_level=-SQRT_SMALL, protected_patch_ids=(), edge_based=True,
                             prismatic=False):
        up = [np.zeros(m.p, dtype=np.uint8) for m in self.meshes]
        sp = [np.zeros(m.p, dtype=np.uint8) for m in self.meshes]
        ids = _i32(protected_patch_ids)
        pg, n_glob = None, 0
        if self.R > 1:
            # syncTools::syncPointList over ranks: shared points are identified by their global label
            gl = [_i32(m.point_global) for m in self.meshes]
            self._keep += gl
            pg, n_glob = _pp(gl), int(max(int(g.max()) for g in gl if len(g)) + 1)
        sweeps = lib().or_poly_select_unrefine(self._h, _pp(err), unrefine_level, n_layers, ids.ctypes.data, len(ids),
                                               1 if edge_based else 0, pg, n_glob, _pp(refine_cell), _pp(up), _pp(sp),
                                               1 if prismatic else 0)
        if sweeps < 0:
            raise RuntimeError("reference FatalError 'problem' (refinement.C:417)")
        return up, sp, sweeps

    def hex2d_select_unrefine(self, err, refine_cell, n_layers, unrefine_level=-SQRT_SMALL, protected_cell=None,
                              protected_patch_ids=()):
        """2-D cutter (hexRef2D): split elements are edges"""
        ue = [np.zeros(m.n_edges, dtype=np.uint8) for m in self.meshes]
        se = [np.zeros(m.n_edges, dtype=np.uint8) for m in self.meshes]
        pc = protected_cell if protected_cell is not None else [None] * self.R
        ids = _i32(protected_patch_ids)
        sweeps = lib().or_hex2d_select_unrefine(self._h, _pp(err), unrefine_level, n_layers, _pp(pc), ids.ctypes.data, len(ids),
                                                _pp(refine_cell), _pp(ue), _pp(se))
        if sweeps < 0:
            raise RuntimeError("reference FatalError 'problem' (hexRef2D.C)")
        return ue, se, sweeps

    def split_edges(self, r):
        out = np.zeros(self.meshes[r].n_edges, dtype=np.uint8)
        lib().or_get_split_edges(self._h, r, out.ctypes.data)
        return out

    def max_cell_field(self, r, v):
        out = np.empty(self.meshes[r].p, dtype=np.float64)
        lib().or_max_cell_field(self._h, r, v.ctypes.data, out.ctypes.data)
        return out

    def split_points(self, r):
        out = np.zeros(self.meshes[r].p, dtype=np.uint8)
        lib().or_get_split_points(self._h, r, out.ctypes.data)
        return out

    def check_levels(self, levels) -> int:
        lv = [_i32(a) for a in levels]
        return int(lib().or_check_levels(self._h, _pp(lv)))

    # -- rows either side of the path
    def swap_owner_values(self, vals):
        """patchNeighbourField of a cell field over the boundary faces of every rank (python restatement of the swap:
        processor patch pairs list their faces in the same order)"""
        out = []
        for r, m in enumerate(self.meshes):
            v = np.ascontiguousarray(vals[r])
            out.append(v[m.owner[m.f:]].copy())
        st = [m.patch_arrays() for m in self.meshes]
        for r, m in enumerate(self.meshes):
            s0, sz, ty, nb = st[r]
            for q in range(len(s0)):
                if ty[q] != 2:
                    continue
                s = int(nb[q])
                ms = self.meshes[s]
                t0, tz, tty, tnb = st[s]
                back = [k for k in range(len(t0)) if tty[k] == 2 and tnb[k] == r][0]
                src = np.ascontiguousarray(vals[s])[ms.owner[t0[back]:t0[back] + tz[back]]]
                out[r][s0[q] - m.f:s0[q] - m.f + sz[q]] = src
        return out

    def correct_fluxes(self, r, face_map, reverse_face_map, U, phi, U_boundary=None, nbr_U=None):
        face_map, reverse_face_map = _i32(face_map), _i32(reverse_face_map)
        U = np.ascontiguousarray(U, dtype=np.float64)
        return int(lib().or_correct_fluxes(self._h, r, len(reverse_face_map), face_map.ctypes.data, reverse_face_map.ctypes.data,
                                           U.ctypes.data, None if U_boundary is None else U_boundary.ctypes.data,
                                           None if nbr_U is None else nbr_U.ctypes.data, phi.ctypes.data))

    def gradient_mag(self, r, x, scheme=0, xb=None, xn=None, ls_p=None, ls_n=None, total_volume=0.0):
        m = self.meshes[r]
        grad = np.empty((m.n, 3), dtype=np.float64)
        mag = np.empty(m.n, dtype=np.float64)
        if xn is None:
            xn = np.zeros(max(m.b, 1))
        lib().or_gradient_mag(self._h, r, scheme, x.ctypes.data, None if xb is None else xb.ctypes.data, xn.ctypes.data,
                              None if ls_p is None else ls_p.ctypes.data, None if ls_n is None else ls_n.ctypes.data,
                              total_volume, grad.ctypes.data, mag.ctypes.data)
        return mag, grad

    def mesh_dx(self, r, geometric_d):
        gd = _i32(geometric_d)
        out = np.empty(self.meshes[r].n, dtype=np.float64)
        lib().or_mesh_dx(self._h, r, gd.ctypes.data, out.ctypes.data)
        return out

    def mesh_dX(self, r):
        out = np.empty((self.meshes[r].n, 3), dtype=np.float64)
        lib().or_mesh_dX(self._h, r, out.ctypes.data)
        return out

    def max_refinement(self, r, max_level, min_dx=0.0, dx=None, error=None):
        out = np.empty(self.meshes[r].n, dtype=np.int32)
        lib().or_max_refinement(self._h, r, max_level, min_dx, None if dx is None else dx.ctypes.data,
                                None if error is None else error.ctypes.data, out.ctypes.data)
        return out

    def cell_size_error(self, r, size_type, geometric_d, lower_unrefine=0.0, lower_refine=0.0, cmpts=(), min_dX=None, max_dX=None,
                        dx=None, dX=None, error=None):
        gd, cm = _i32(geometric_d), _i32(cmpts)
        lo = np.zeros(3) if min_dX is None else np.asarray(min_dX, dtype=np.float64)
        hi = np.zeros(3) if max_dX is None else np.asarray(max_dX, dtype=np.float64)
        if error is None:
            error = np.zeros(self.meshes[r].n, dtype=np.float64)
        lib().or_cell_size_error(self._h, r, size_type, gd.ctypes.data, lower_unrefine, lower_refine, cm.ctypes.data, len(cm),
                                 lo.ctypes.data, hi.ctypes.data, None if dx is None else dx.ctypes.data,
                                 None if dX is None else dX.ctypes.data, error.ctypes.data)
        return error

    def history_clusters(self, r):
        m = self.meshes[r]
        c2c = np.empty(m.n, dtype=np.int32)
        nsp = 0 if m.split_parent is None else len(m.split_parent)
        ncl = lib().or_history_clusters(self._h, r, nsp, c2c.ctypes.data)
        return c2c, int(ncl)

    def blocked_faces(self, r, cell_to_cluster):
        m = self.meshes[r]
        c2c = _i32(cell_to_cluster)
        bl = np.empty(m.f + m.b, dtype=np.uint8)
        nu = lib().or_blocked_faces(self._h, r, c2c.ctypes.data, bl.ctypes.data)
        return bl, int(nu)

    def constrain_decomposition(self, r, cell_to_cluster, n_clusters, decomposition):
        c2c = _i32(cell_to_cluster)
        return int(lib().or_constrain_decomposition(self._h, r, c2c.ctypes.data, n_clusters, decomposition.ctypes.data))

    def imbalance(self):
        g = C.c_int64(0)
        return float(lib().or_imbalance(self._h, C.byref(g))), g.value

    def multicomponent(self, r, errors, max_cell_levels=None):
        n = self.meshes[r].n
        err = np.empty(n, dtype=np.float64)
        lib().or_multicomponent_error(n, len(errors), _pp(errors), err.ctypes.data)
        mr = None
        if max_cell_levels is not None:
            lv = [_i32(a) for a in max_cell_levels]
            mr = np.empty(n, dtype=np.int32)
            lib().or_multicomponent_max_refinement(self._h, r, len(errors), _pp(errors), _pp(lv), mr.ctypes.data)
        return err, mr


def remap_refine_cell(cell_map, reverse_map, refine_cell):
    cell_map, reverse_map = _i32(cell_map), _i32(reverse_map)
    out = np.zeros(len(cell_map), dtype=np.uint8)
    lib().or_remap_refine_cell(len(cell_map), cell_map.ctypes.data, reverse_map.ctypes.data, refine_cell.ctypes.data,
                               out.ctypes.data)
    return out


def map_field(cell_map, old, ncomp=1):
    cell_map = _i32(cell_map)
    old = np.ascontiguousarray(old, dtype=np.float64)
    new = np.empty((len(cell_map),) + old.shape[1:], dtype=np.float64)
    lib().or_map_field(len(cell_map), cell_map.ctypes.data, ncomp, old.ctypes.data, new.ctypes.data)
    return new


def map_field_conservative(cell_map, reverse_map, vol_old, old, ncomp=1):
    cell_map, reverse_map = _i32(cell_map), _i32(reverse_map)
    old = np.ascontiguousarray(old, dtype=np.float64)
    vol_old = np.ascontiguousarray(vol_old, dtype=np.float64)
    new = np.empty((len(cell_map),) + old.shape[1:], dtype=np.float64)
    lib().or_map_field_conservative(len(reverse_map), len(cell_map), cell_map.ctypes.data, reverse_map.ctypes.data,
                                    vol_old.ctypes.data, ncomp, old.ctypes.data, new.ctypes.data)
    return new


def map_old_volumes(cell_map, reverse_map, v0_old):
    """V0 / V00 through a topology change (fvMesh::updateMesh): direct map + volume of the merged cells injected"""
    cell_map, reverse_map = _i32(cell_map), _i32(reverse_map)
    v0_old = np.ascontiguousarray(v0_old, dtype=np.float64)
    new = np.empty(len(cell_map), dtype=np.float64)
    n = lib().or_map_old_volumes(len(reverse_map), len(cell_map), cell_map.ctypes.data, reverse_map.ctypes.data, v0_old.ctypes.data,
                                 new.ctypes.data)
    return new, int(n)


def update_levels(cell_map, old_level, refined_old=None, unrefined_old=None):
    cell_map, old_level = _i32(cell_map), _i32(old_level)
    out = np.empty(len(cell_map), dtype=np.int32)
    lib().or_update_levels(len(cell_map), cell_map.ctypes.data, old_level.ctypes.data,
                           None if refined_old is None else refined_old.ctypes.data,
                           None if unrefined_old is None else unrefined_old.ctypes.data, out.ctypes.data)
    return out


def reorder_labels(old_to_new, n_new, old_values):
    old_to_new, old_values = _i32(old_to_new), _i32(old_values)
    out = np.empty(n_new, dtype=np.int32)
    lib().or_reorder_labels(len(old_values), n_new, old_to_new.ctypes.data, old_values.ctypes.data, out.ctypes.data)
    return out


def gather_labels(new_to_old, old_values):
    new_to_old, old_values = _i32(new_to_old), _i32(old_values)
    out = np.empty(len(new_to_old), dtype=np.int32)
    lib().or_map_labels(len(new_to_old), new_to_old.ctypes.data, old_values.ctypes.data, out.ctypes.data)
    return out


def remap_protected_cells(cell_map, old_protected):
    cell_map = _i32(cell_map)
    old_protected = np.ascontiguousarray(old_protected, dtype=np.uint8)
    out = np.empty(len(cell_map), dtype=np.uint8)
    lib().or_map_protected(len(cell_map), cell_map.ctypes.data, old_protected.ctypes.data, out.ctypes.data)
    return out


def proc_load(distribution, n_procs):
    d = _i32(distribution)
    out = np.zeros(n_procs, dtype=np.int64)
    lib().or_proc_load(len(d), d.ctypes.data, n_procs, out.ctypes.data)
    return out


def set_threads(t: int):
    lib().or_set_threads(int(t))


def max_threads() -> int:
    return int(lib().or_max_threads())
