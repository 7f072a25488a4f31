"""ctypes binding of the C-ABI in include/amr_b200.h (libamr_b200.so).

Arguments may be numpy arrays (host pointers: the drop-in case), torch CUDA tensors or
raw integer device addresses (resident case).  The library detects the memory space itself.
There is no fallback: if the CUDA library is missing or no GPU is present, this raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _build

AMR_OK = 0
EST = {"delta": 0, "scaledDelta": 1, "Lohner": 2, "fieldValue": 3, "codedDelta": 4, "densityGradient": 5}
HEX3D, HEX2D, POLY3D, POLY2D = 0, 1, 2, 3
MAP_DIRECT, MAP_CONSERVATIVE = 0, 1

SMALL = 1e-15
GREAT = 1e15
SQRT_SMALL = float(np.sqrt(1e-15))
LABEL_MAX = 2147483647

_lib = None

_SIGS = {
    "amr_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "amr_ctx_destroy": (C.c_int, [C.c_void_p]),
    "amr_last_error": (C.c_char_p, [C.c_void_p]),
    "amr_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "amr_comm_unique_id": (C.c_int, [C.c_void_p]),
    "amr_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "amr_mesh_set": (C.c_int, [C.c_void_p] + [C.c_int64] * 4 + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 4),
    "amr_mesh_set_points": (C.c_int, [C.c_void_p] * 4),
    "amr_mesh_update": (C.c_int, [C.c_void_p] + [C.c_int64] * 4 + [C.c_void_p] * 2 + [C.c_int64] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 4),
    "amr_mesh_update_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 5),
    "amr_mesh_set_edges": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 4),
    "amr_mesh_set_faces": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_mesh_set_patch_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "amr_mesh_set_geometry": (C.c_int, [C.c_void_p] * 6),
    "amr_estimate_lohner": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]),
    "amr_levels_set": (C.c_int, [C.c_void_p] * 5 + [C.c_int64]),
    "amr_protected_cells_set": (C.c_int, [C.c_void_p, C.c_void_p]),
    "amr_estimate": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                               C.c_void_p, C.c_void_p]),
    "amr_normalize": (C.c_int, [C.c_void_p] + [C.c_double] * 4 + [C.c_void_p]),
    "amr_normalize_range": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_protect_patches": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "amr_select_refine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_int,
                                    C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "amr_consistent_refinement": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "amr_remap_refine_cell": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_select_unrefine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_get_split_elems": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "amr_max_cell_field": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_check_levels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_check_point_levels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_map_field": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p]),
    "amr_map_fields": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_map_fields_and_marking": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int] + [C.c_void_p] * 6),
    "amr_update_levels": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 5),
    "amr_remap_labels": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "amr_remap_protected_cells": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_map_old_volumes": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_correct_fluxes": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 7),
    "amr_mesh_set_ls_vectors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_estimate_gradient": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                        C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_mesh_size": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_max_refinement": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_estimate_cell_size": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_multicomponent": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 5),
    "amr_imbalance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_history_clusters": (C.c_int, [C.c_void_p] * 6),
    "amr_history_constrain_decomposition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "amr_proc_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "amr_synchronize": (C.c_int, [C.c_void_p]),
    "amr_launch_count": (C.c_int64, [C.c_void_p]),
    "amr_device_bytes": (C.c_int64, [C.c_void_p]),
}

EXPORTS = tuple(_SIGS)


def load(build: bool = True):
    """dlopen libamr_b200.so (building it first if needed) and declare every entry point."""
    global _lib
    if _lib is None:
        path = _build.build_cuda() if build else _build.LIB_CUDA
        L = C.CDLL(str(path))
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)     # AttributeError = a symbol of include/amr_b200.h is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class AmrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"amr_b200 error {code}: {msg}")
        self.code = code


def _ptr(a):
    """host numpy array / torch tensor / int device address / None -> void*"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


def _i32(a):
    return None if a is None else (a if not isinstance(a, (list, tuple)) else np.asarray(a, dtype=np.int32))


class Context:
    """One amr_ctx (one MPI rank / one GPU)."""

    def __init__(self, device: int = 0):
        L = load()
        h = C.c_void_p()
        rc = L.amr_ctx_create(device, C.byref(h))
        if rc != AMR_OK:
            raise AmrError(rc, "amr_ctx_create failed (no usable CUDA device; this library has no CPU path)")
        self._h = h
        self.L = L
        self.n = self.f = self.b = self.p = 0

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.L.amr_ctx_destroy(h)

    __del__ = close

    def _ck(self, rc):
        if rc != AMR_OK:
            raise AmrError(rc, self.L.amr_last_error(self._h).decode())

    # -- plumbing
    def set_stream(self, stream_ptr: int):
        self._ck(self.L.amr_set_stream(self._h, stream_ptr))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = load().amr_comm_unique_id(buf)
        if rc != AMR_OK:
            raise AmrError(rc, "ncclGetUniqueId failed")
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        self._ck(self.L.amr_comm_init(self._h, nranks, rank, C.create_string_buffer(uid, 128)))

    def prepare(self, name, *args):
        """A call of C entry point `name` with its arguments converted once (arrays / tensors -> pointers): what a host
        code does when it calls the C-ABI in a time loop.  Returns a function that makes the call and checks the
        return code; the converted arguments (and the buffers behind them) stay referenced by it."""
        fn = getattr(self.L, name)
        conv = tuple(_ptr(a) if (a is None or isinstance(a, np.ndarray) or hasattr(a, "data_ptr")) else a for a in args)
        keep = args
        h, ck = self._h, self._ck

        def call():
            rc = fn(h, *conv)
            if rc != AMR_OK:
                ck(rc)
        call.keep = keep
        return call

    def synchronize(self):
        """wait for the context's stream; raises what stream-ordered calls could not report (peer timeout)"""
        self._ck(self.L.amr_synchronize(self._h))

    @property
    def launches(self) -> int:
        return int(self.L.amr_launch_count(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self.L.amr_device_bytes(self._h))

    # -- mesh state
    def set_mesh(self, mesh):
        st, sz, ty, nb = mesh.patch_arrays()
        self._ck(self.L.amr_mesh_set(self._h, mesh.n, mesh.f, mesh.b, mesh.p, _ptr(mesh.owner), _ptr(mesh.neighbour),
                                     len(st), _ptr(st), _ptr(sz), _ptr(ty), _ptr(nb)))
        self.n, self.f, self.b, self.p = mesh.n, mesh.f, mesh.b, mesh.p
        if mesh.pc_off is not None:
            self._ck(self.L.amr_mesh_set_points(self._h, _ptr(mesh.pc_off), _ptr(mesh.pc_cells), _ptr(mesh.bnd_point)))
        self.n_edges = int(getattr(mesh, "n_edges", 0) or 0)
        if self.n_edges:
            self._ck(self.L.amr_mesh_set_edges(self._h, mesh.n_edges, _ptr(mesh.edge_pts), _ptr(mesh.ec_off), _ptr(mesh.ec_cells),
                                               _ptr(mesh.bnd_edge)))
        if getattr(mesh, "face_off", None) is not None:
            self._ck(self.L.amr_mesh_set_faces(self._h, _ptr(mesh.face_off), _ptr(mesh.face_pts)))
        if getattr(mesh, "patch_pt_off", None) is not None:
            self._ck(self.L.amr_mesh_set_patch_points(self._h, _ptr(mesh.patch_pt_off), _ptr(mesh.patch_pts)))
        if getattr(mesh, "Sf", None) is not None and mesh.vol is not None:
            self._ck(self.L.amr_mesh_set_geometry(self._h, _ptr(mesh.weights), _ptr(mesh.Sf), _ptr(mesh.magSf),
                                                  _ptr(mesh.delta_coeffs), _ptr(mesh.vol)))
        if mesh.cell_level is not None:
            sp = mesh.split_parent
            nsp = 0 if sp is None else len(sp)
            self._ck(self.L.amr_levels_set(self._h, _ptr(mesh.cell_level), _ptr(mesh.point_level), _ptr(mesh.visible),
                                           _ptr(sp) if nsp else None, nsp))

    def update_mesh(self, mesh, delta):
        """Incremental hand-over after a topology change (amr_mesh_update + amr_mesh_update_points): `mesh` is the NEW mesh
        (sizes, boundary owners, patches), `delta` a MeshDelta (blastamr_b200/mesh.py: mesh_delta)."""
        st, sz, ty, nb = mesh.patch_arrays()
        d = delta
        bown = np.ascontiguousarray(mesh.owner[mesh.f:], dtype=np.int32)
        self._ck(self.L.amr_mesh_update(self._h, mesh.n, mesh.f, mesh.b, mesh.p, _ptr(d.cell_map), _ptr(d.reverse_cell_map), len(d.dirty_cells),
                                        _ptr(d.dirty_cells), _ptr(d.dirty_cell_off), _ptr(d.dirty_cell_nbrs), _ptr(d.dirty_cell_level),
                                        _ptr(d.dirty_parent_index), _ptr(bown), len(st), _ptr(st), _ptr(sz), _ptr(ty), _ptr(nb)))
        self.n, self.f, self.b, self.p = mesh.n, mesh.f, mesh.b, mesh.p
        if d.point_map is not None:
            self._ck(self.L.amr_mesh_update_points(self._h, _ptr(d.point_map), len(d.dirty_points), _ptr(d.dirty_points), _ptr(d.dirty_point_off),
                                                   _ptr(d.dirty_point_cells), _ptr(d.dirty_bnd_point), _ptr(d.dirty_point_level)))

    def set_mesh_raw(self, n, f, b, p, owner, neighbour, patch_start, patch_size, patch_type, patch_nbr):
        self._ck(self.L.amr_mesh_set(self._h, n, f, b, p, _ptr(owner), _ptr(neighbour), len(patch_start),
                                     _ptr(patch_start), _ptr(patch_size), _ptr(patch_type), _ptr(patch_nbr)))
        self.n, self.f, self.b, self.p = n, f, b, p

    def set_points(self, pc_off, pc_cells, bnd_point):
        self._ck(self.L.amr_mesh_set_points(self._h, _ptr(pc_off), _ptr(pc_cells), _ptr(bnd_point)))

    def set_levels(self, cell_level, point_level=None, visible=None, split_parent=None):
        nsp = 0 if split_parent is None else len(split_parent)
        self._ck(self.L.amr_levels_set(self._h, _ptr(cell_level), _ptr(point_level), _ptr(visible),
                                       _ptr(split_parent) if nsp else None, nsp))

    def set_option(self, option: int, value: int):
        """0 = edgeBasedConsistency, 1 = polyRefiner force_"""
        self._ck(self.L.amr_set_option(self._h, option, value))

    def set_protected_cells(self, prot):
        self._ck(self.L.amr_protected_cells_set(self._h, _ptr(prot)))

    # -- stage 1
    def estimate(self, kind, x, nbr_values=None, p0=SMALL, p1=0.0, thresholds=None, out=None, keep_resident=False):
        """thresholds = (lowerRefine, upperRefine, lowerUnrefine, upperUnrefine) -> normalised; None -> raw."""
        k = EST[kind] if isinstance(kind, str) else kind
        thr = None if thresholds is None else np.asarray(thresholds, dtype=np.float64)
        if out is None and not keep_resident:
            out = np.empty(self.n, dtype=np.float64)
        self._ck(self.L.amr_estimate(self._h, k, _ptr(x), _ptr(nbr_values), p0, p1, 0 if thr is None else 1, _ptr(thr),
                                     _ptr(out)))
        return out

    def estimate_lohner(self, x, epsilon, x_boundary=None, nbr_values=None, thresholds=None, out=None, keep_resident=False):
        thr = None if thresholds is None else np.asarray(thresholds, dtype=np.float64)
        if out is None and not keep_resident:
            out = np.empty(self.n, dtype=np.float64)
        self._ck(self.L.amr_estimate_lohner(self._h, _ptr(x), _ptr(x_boundary), _ptr(nbr_values), epsilon,
                                            0 if thr is None else 1, _ptr(thr), _ptr(out)))
        return out

    def normalize(self, err, lower_refine, upper_refine, lower_unrefine, upper_unrefine):
        self._ck(self.L.amr_normalize(self._h, lower_refine, upper_refine, lower_unrefine, upper_unrefine, _ptr(err)))
        return err

    def normalize_range(self, err, refine_threshold=0.5, unrefine_threshold=0.4, scale=True):
        """gradientRange.C:102-109: thresholds from gMin/gMax of the raw error, then normalize."""
        thr = np.zeros(4)
        mm = np.zeros(2)
        self._ck(self.L.amr_normalize_range(self._h, refine_threshold, unrefine_threshold, 1 if scale else 0, _ptr(err),
                                            _ptr(thr), _ptr(mm)))
        return err, thr, mm

    def protect_patches(self, patch_ids, n_layers, err=None) -> int:
        ids = np.asarray(patch_ids, dtype=np.int32)
        cnt = C.c_int64(0)
        self._ck(self.L.amr_protect_patches(self._h, _ptr(ids), len(ids), n_layers, _ptr(err), C.byref(cnt)))
        return cnt.value

    # -- stages 2+3
    def select_refine(self, err, max_level, n_layers, lower=SQRT_SMALL, upper=GREAT, protected_patch_ids=(),
                      max_cells=LABEL_MAX, kind=HEX3D, refine_cell_out=None, cells_out=None, want_count=True):
        mcl = None if np.isscalar(max_level) else max_level
        uni = int(max_level) if np.isscalar(max_level) else 0
        ids = np.asarray(protected_patch_ids, dtype=np.int32)
        if cells_out is None:
            cells_out = np.empty(max(self.n, 1), dtype=np.int32)
        cnt = C.c_int64(0)
        sweeps = C.c_int(0)
        self._ck(self.L.amr_select_refine(self._h, _ptr(err), lower, upper, _ptr(mcl), uni, n_layers, _ptr(ids), len(ids),
                                          max_cells, kind, _ptr(refine_cell_out), _ptr(cells_out),
                                          C.byref(cnt) if want_count else None, C.byref(sweeps)))
        if isinstance(cells_out, np.ndarray):
            return cells_out[:cnt.value], sweeps.value
        return cnt.value, sweeps.value

    def consistent_refinement(self, cells, max_set=True):
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        out = np.empty(max(self.n, 1), dtype=np.int32)
        cnt = C.c_int64(0)
        sweeps = C.c_int(0)
        self._ck(self.L.amr_consistent_refinement(self._h, _ptr(cells), len(cells), 1 if max_set else 0, _ptr(out),
                                                  C.byref(cnt), C.byref(sweeps)))
        return out[:cnt.value], sweeps.value

    def remap_refine_cell(self, cell_map, reverse_map, refine_cell=None, out=None):
        n_new = len(cell_map) if isinstance(cell_map, np.ndarray) else cell_map.numel()
        n_old = len(reverse_map) if isinstance(reverse_map, np.ndarray) else reverse_map.numel()
        self._ck(self.L.amr_remap_refine_cell(self._h, n_old, n_new, _ptr(cell_map), _ptr(reverse_map), _ptr(refine_cell),
                                              _ptr(out)))
        return out

    def select_unrefine(self, err, n_layers, unrefine_level=-SQRT_SMALL, refine_cell=None, protected_patch_ids=(),
                        kind=HEX3D, elems_out=None, want_count=True):
        ids = np.asarray(protected_patch_ids, dtype=np.int32)
        if elems_out is None:
            elems_out = np.empty(max(self.p, getattr(self, "n_edges", 0), 1), dtype=np.int32)
        cnt = C.c_int64(0)
        sweeps = C.c_int(0)
        self._ck(self.L.amr_select_unrefine(self._h, _ptr(err), unrefine_level, _ptr(refine_cell), n_layers, _ptr(ids),
                                            len(ids), kind, _ptr(elems_out), C.byref(cnt) if want_count else None,
                                            C.byref(sweeps)))
        if isinstance(elems_out, np.ndarray):
            return elems_out[:cnt.value], sweeps.value
        return cnt.value, sweeps.value

    def split_elems(self, kind=HEX3D):
        out = np.empty(max(self.p, getattr(self, "n_edges", 0), 1), dtype=np.int32)
        cnt = C.c_int64(0)
        self._ck(self.L.amr_get_split_elems(self._h, kind, _ptr(out), C.byref(cnt)))
        return out[:cnt.value]

    def max_cell_field(self, v):
        out = np.empty(self.p, dtype=np.float64)
        self._ck(self.L.amr_max_cell_field(self._h, _ptr(v), _ptr(out)))
        return out

    def check_levels(self, cell_level=None) -> int:
        cnt = C.c_int64(0)
        rc = self.L.amr_check_levels(self._h, _ptr(cell_level), C.byref(cnt))
        if rc not in (AMR_OK, -5):
            self._ck(rc)
        return cnt.value

    def check_point_levels(self, point_level=None) -> int:
        """hexRef.C:3027-3057; returns this rank's number of inconsistent points (the call itself reports -5 when any rank has one)"""
        cnt = C.c_int64(0)
        rc = self.L.amr_check_point_levels(self._h, _ptr(point_level), C.byref(cnt))
        if rc not in (AMR_OK, -5):
            self._ck(rc)
        return cnt.value if rc == AMR_OK else max(cnt.value, -1) if cnt.value else -1

    # -- stage 4
    def map_field(self, cell_map, old, n_old=None, n_new=None, ncomp=1, out=None, mode=MAP_DIRECT, reverse_map=None,
                  vol_old=None):
        if n_new is None:
            n_new = len(cell_map)
        if n_old is None:
            n_old = old.shape[0]
        if out is None:
            out = np.empty((n_new,) + tuple(old.shape[1:]), dtype=np.float64)
        self._ck(self.L.amr_map_field(self._h, n_old, n_new, _ptr(cell_map), ncomp, _ptr(old), _ptr(out), mode,
                                      _ptr(reverse_map), _ptr(vol_old)))
        return out

    def map_fields(self, cell_map, olds, n_old, n_new, ncomps, outs):
        """fvMesh::mapFields over a list of registered fields (one shared cellMap)"""
        k = len(olds)
        nc = (C.c_int * k)(*[int(c) for c in ncomps])
        po = (C.c_void_p * k)(*[_ptr(a) for a in olds])
        pn = (C.c_void_p * k)(*[_ptr(a) for a in outs])
        self._ck(self.L.amr_map_fields(self._h, n_old, n_new, _ptr(cell_map), k, nc, po, pn))
        return outs

    def map_fields_and_marking(self, cell_map, olds, n_old, n_new, ncomps, outs, reverse_map=None, refine_cell=None, out_marking=None):
        """fvMesh::mapFields + the refiner's refineCell update in one pass over cellMap (reverse_map None: fields only)"""
        k = len(olds)
        nc = (C.c_int * max(k, 1))(*[int(c) for c in ncomps])
        po = (C.c_void_p * max(k, 1))(*[_ptr(a) for a in olds])
        pn = (C.c_void_p * max(k, 1))(*[_ptr(a) for a in outs])
        self._ck(self.L.amr_map_fields_and_marking(self._h, n_old, n_new, _ptr(cell_map), k, nc, po, pn, _ptr(reverse_map),
                                                   _ptr(refine_cell), _ptr(out_marking)))
        return outs

    def update_levels(self, cell_map, old_level=None, refined_old=None, unrefined_old=None):
        out = np.empty(len(cell_map), dtype=np.int32)
        self._ck(self.L.amr_update_levels(self._h, len(cell_map), _ptr(cell_map), _ptr(old_level), _ptr(refined_old),
                                          _ptr(unrefined_old), _ptr(out)))
        return out

    # -- rows either side of the path
    def reorder_labels(self, old_to_new, n_new, old_values):
        """hexRef::updateMesh, reorder(reverseMap, nNew, -1, list)"""
        out = np.empty(n_new, dtype=np.int32)
        self._ck(self.L.amr_remap_labels(self._h, len(old_values), n_new, _ptr(old_to_new), 0, _ptr(old_values), _ptr(out)))
        return out

    def gather_labels(self, new_to_old, old_values):
        out = np.empty(len(new_to_old), dtype=np.int32)
        self._ck(self.L.amr_remap_labels(self._h, len(old_values), len(new_to_old), _ptr(new_to_old), 1, _ptr(old_values), _ptr(out)))
        return out

    def remap_protected_cells(self, cell_map, old_protected):
        out = np.empty(len(cell_map), dtype=np.uint8)
        self._ck(self.L.amr_remap_protected_cells(self._h, len(old_protected), len(cell_map), _ptr(cell_map), _ptr(old_protected),
                                                  _ptr(out)))
        return out

    def map_old_volumes(self, cell_map, reverse_map, olds):
        """V0 / V00 through a topology change (fvMesh::updateMesh): returns (new volumes, number of merged-away cells)"""
        k = len(olds)
        n_new = len(cell_map)
        outs = [np.empty(n_new, dtype=np.float64) for _ in olds]
        po = (C.c_void_p * max(k, 1))(*[_ptr(a) for a in olds])
        pn = (C.c_void_p * max(k, 1))(*[_ptr(a) for a in outs])
        nm = C.c_int64(0)
        self._ck(self.L.amr_map_old_volumes(self._h, len(reverse_map), n_new, _ptr(cell_map), _ptr(reverse_map), k, po, pn, C.byref(nm)))
        return outs, nm.value

    def correct_fluxes(self, face_map, reverse_face_map, U, phi, U_boundary=None, nbr_U=None):
        """adaptiveFvMesh::updateMesh flux fix-up on the mesh held by the context; phi is rewritten in place"""
        cnt = C.c_int64(0)
        self._ck(self.L.amr_correct_fluxes(self._h, len(reverse_face_map), _ptr(face_map), _ptr(reverse_face_map), _ptr(U),
                                           _ptr(U_boundary), _ptr(nbr_U), _ptr(phi), C.byref(cnt)))
        return cnt.value

    def set_ls_vectors(self, p_vectors, n_vectors):
        self._ck(self.L.amr_mesh_set_ls_vectors(self._h, _ptr(p_vectors), _ptr(n_vectors)))

    def estimate_gradient(self, x, scheme=0, x_boundary=None, nbr_values=None, total_volume=0.0, refine_threshold=0.5,
                          unrefine_threshold=0.4, scale=True, want_grad=False):
        """gradientRange::update: returns (error, thresholds[4], grad or None)"""
        out = np.empty(self.n, dtype=np.float64)
        thr = np.zeros(4)
        grad = np.empty((self.n, 3), dtype=np.float64) if want_grad else None
        self._ck(self.L.amr_estimate_gradient(self._h, scheme, _ptr(x), _ptr(x_boundary), _ptr(nbr_values), total_volume,
                                              refine_threshold, unrefine_threshold, 1 if scale else 0, _ptr(thr), _ptr(grad), _ptr(out)))
        return out, thr, grad

    def mesh_size(self, geometric_d, want_dx=True, want_dX=True):
        gd = np.asarray(geometric_d, dtype=np.int32)
        dx = np.empty(self.n, dtype=np.float64) if want_dx else None
        dX = np.empty((self.n, 3), dtype=np.float64) if want_dX else None
        self._ck(self.L.amr_mesh_size(self._h, _ptr(gd), _ptr(dx), _ptr(dX)))
        return dx, dX

    def max_refinement(self, max_level, min_dx=0.0, dx=None, error=None):
        out = np.empty(self.n, dtype=np.int32)
        self._ck(self.L.amr_max_refinement(self._h, max_level, min_dx, _ptr(dx), _ptr(error), _ptr(out)))
        return out

    def estimate_cell_size(self, size_type, geometric_d, lower_unrefine=0.0, lower_refine=0.0, cmpts=(), min_dX=None, max_dX=None,
                           dx=None, dX=None, error=None):
        gd = np.asarray(geometric_d, dtype=np.int32)
        cm = np.asarray(cmpts, dtype=np.int32)
        lo = None if min_dX is None else np.asarray(min_dX, dtype=np.float64)
        hi = None if max_dX is None else np.asarray(max_dX, dtype=np.float64)
        if error is None:
            error = np.zeros(self.n, dtype=np.float64)
        self._ck(self.L.amr_estimate_cell_size(self._h, size_type, _ptr(gd), lower_unrefine, lower_refine, _ptr(cm), len(cm),
                                               _ptr(lo), _ptr(hi), _ptr(dx), _ptr(dX), _ptr(error)))
        return error

    def imbalance(self):
        r, g = C.c_double(0), C.c_int64(0)
        self._ck(self.L.amr_imbalance(self._h, C.byref(r), C.byref(g)))
        return r.value, g.value

    def history_clusters(self, clusters_in=None, want_blocked=True):
        """markCommonCells + add: returns (cellToCluster, blockedFace, nClusters, nUnblocked)"""
        c2c = np.empty(self.n, dtype=np.int32)
        bl = np.empty(self.f + self.b, dtype=np.uint8) if want_blocked else None
        nc, nu = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.amr_history_clusters(self._h, _ptr(clusters_in), _ptr(c2c), _ptr(bl), C.byref(nc), C.byref(nu)))
        return c2c, bl, nc.value, nu.value

    def constrain_decomposition(self, decomposition):
        nch = C.c_int64(0)
        self._ck(self.L.amr_history_constrain_decomposition(self._h, _ptr(decomposition), C.byref(nch)))
        return nch.value

    def proc_load(self, distribution, n_procs):
        out = np.zeros(n_procs, dtype=np.int64)
        self._ck(self.L.amr_proc_load(self._h, _ptr(distribution), n_procs, _ptr(out)))
        return out

    def multicomponent(self, errors, max_cell_levels=None):
        """multicomponent::update + maxRefinement: returns (error, maxRefinement or None, waves)"""
        k = len(errors)
        pe = (C.c_void_p * max(k, 1))(*[_ptr(e) for e in errors])
        out = np.empty(self.n, dtype=np.float64)
        mr, pl = None, None
        if max_cell_levels is not None:
            pl = (C.c_void_p * max(k, 1))(*[_ptr(l) for l in max_cell_levels])
            mr = np.empty(self.n, dtype=np.int32)
        waves = C.c_int(0)
        self._ck(self.L.amr_multicomponent(self._h, k, pe, pl, _ptr(out), _ptr(mr), C.byref(waves)))
        return out, mr, waves.value
