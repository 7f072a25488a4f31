"""ctypes binding of libamrhost.so (include/amr_host.h): the C++ host layer that mirrors the
reference's adaptiveFvMesh / errorEstimator / fvMeshRefiner classes above the CUDA C-ABI.
`AdaptiveFvMesh` below plays the solver: it owns the (synthetic) topology engine, registers fields and
calls `update()` once per time step -- the call a reference user makes (`mesh.update()`)."""
from __future__ import annotations

import ctypes as C
import re
from typing import Dict, Optional

import numpy as np

from . import _build
from .mesh import Octree, PolyMesh

_lib = None
TOPO_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.c_int64)


def lib():
    global _lib
    if _lib is None:
        _build.build_cuda()
        L = C.CDLL(str(_build.build_host()))
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        L.amrhost_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
        L.amrhost_destroy.argtypes = [vp]
        L.amrhost_last_error.restype = C.c_char_p
        L.amrhost_last_error.argtypes = [vp]
        L.amrhost_set_dict.argtypes = [vp, C.c_char_p]
        L.amrhost_set_topo_callbacks.argtypes = [vp, TOPO_CB, TOPO_CB, vp]
        L.amrhost_set_mesh.argtypes = [vp, i64, i64, i64, i64, vp, vp, C.c_int, C.POINTER(C.c_char_p), vp, vp, vp, vp, C.c_int]
        L.amrhost_set_points.argtypes = [vp, vp, vp, vp]
        L.amrhost_set_edges.argtypes = [vp, i64, vp, vp, vp, vp]
        L.amrhost_set_faces.argtypes = [vp, vp, vp]
        L.amrhost_set_geometry.argtypes = [vp] * 6
        L.amrhost_set_levels.argtypes = [vp, vp, vp, vp, vp, i64]
        L.amrhost_set_map.argtypes = [vp, i64, i64, vp, vp]
        L.amrhost_mesh_changed.argtypes = [vp]
        L.amrhost_set_field.argtypes = [vp, C.c_char_p, C.c_int, vp, i64]
        L.amrhost_get_field.argtypes = [vp, C.c_char_p, vp, i64, C.POINTER(i64)]
        L.amrhost_set_time.argtypes = [vp, i32, dbl]
        L.amrhost_update.argtypes = [vp, C.POINTER(C.c_int)]
        L.amrhost_n_cells.restype = i64
        L.amrhost_n_cells.argtypes = [vp]
        L.amrhost_last_lists.argtypes = [vp, vp, C.POINTER(i64), vp, C.POINTER(i64)]
        L.amrhost_get_error.argtypes = [vp, vp, i64]
        L.amrhost_parse_dict.argtypes = [C.c_char_p, C.c_char_p, i64]
        L.amrhost_selection_table.argtypes = [C.c_int, C.c_char_p, i64]
        _lib = L
    return _lib


class FatalError(RuntimeError):
    pass


def parse_dict_keys(text: str):
    buf = C.create_string_buffer(1 << 16)
    if lib().amrhost_parse_dict(text.encode(), buf, len(buf)) != 0:
        raise FatalError(lib().amrhost_last_error(None).decode())
    return buf.value.decode().split()


def selection_table(which: str):
    buf = C.create_string_buffer(1 << 12)
    lib().amrhost_selection_table(0 if which == "errorEstimator" else 1, buf, len(buf))
    return buf.value.decode().split()


class AdaptiveFvMesh:
    """adaptiveFvMesh + a solver's time loop, with the octree engine as polyTopoChange."""

    def __init__(self, dict_text: str, octree: Octree, device: int = 0, geometry: bool = False, faces: bool = False):
        self.L = lib()
        h = C.c_void_p()
        if self.L.amrhost_create(dict_text.encode(), device, C.byref(h)) != 0:
            raise FatalError(self.L.amrhost_last_error(None).decode())
        self._h = h
        self.oc = octree
        self.geometry = geometry
        self.faces = faces
        # polyRefiner on a 2-D mesh selects split POINTS (prismatic2DRefinement); the 2-D topology engine merges around
        # the through-thickness edge joining the two split points
        self.prismatic = bool(getattr(octree, "two_d", False)) and re.search(r"refiner\s+polyRefiner", dict_text) is not None
        self.mesh: Optional[PolyMesh] = None
        self.time_index = 0
        self._cb_r = TOPO_CB(self._on_refine)
        self._cb_u = TOPO_CB(self._on_unrefine)
        self.L.amrhost_set_topo_callbacks(h, self._cb_r, self._cb_u, None)
        self._install(octree.build(geometry=geometry, faces=faces))
        self._ck(self.L.amrhost_mesh_changed(h))
        self.history = []

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.L.amrhost_destroy(h)

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise FatalError(self.L.amrhost_last_error(self._h).decode())

    def _install(self, m: PolyMesh):
        self.mesh = m
        st, sz, ty, nb = m.patch_arrays()
        names = (C.c_char_p * len(m.patches))(*[q.name.encode() for q in m.patches])
        nsd = 2 if getattr(self.oc, "two_d", False) else 3
        self._ck(self.L.amrhost_set_mesh(self._h, m.n, m.f, m.b, m.p, m.owner.ctypes.data, m.neighbour.ctypes.data, len(st), names,
                                         st.ctypes.data, sz.ctypes.data, ty.ctypes.data, nb.ctypes.data, nsd))
        self._ck(self.L.amrhost_set_points(self._h, m.pc_off.ctypes.data, m.pc_cells.ctypes.data, m.bnd_point.ctypes.data))
        if m.n_edges:
            self._ck(self.L.amrhost_set_edges(self._h, m.n_edges, m.edge_pts.ctypes.data, m.ec_off.ctypes.data, m.ec_cells.ctypes.data,
                                              m.bnd_edge.ctypes.data))
        if m.face_off is not None:
            self._ck(self.L.amrhost_set_faces(self._h, m.face_off.ctypes.data, m.face_pts.ctypes.data))
        if self.geometry:
            self._ck(self.L.amrhost_set_geometry(self._h, m.weights.ctypes.data, m.Sf.ctypes.data, m.magSf.ctypes.data,
                                                 m.delta_coeffs.ctypes.data, m.vol.ctypes.data))
        sp = m.split_parent
        self._ck(self.L.amrhost_set_levels(self._h, m.cell_level.ctypes.data, m.point_level.ctypes.data, m.visible.ctypes.data,
                                           sp.ctypes.data if len(sp) else None, len(sp)))

    # ---- polyTopoChange stand-ins (called back from fvMeshHexRefiner::refine)
    def _on_refine(self, user, elems, n):
        try:
            cells = np.ctypeslib.as_array(elems, shape=(n,)).copy()
            n_old = self.mesh.n
            cm = self.oc.refine(cells)
            rev = np.arange(n_old, dtype=np.int32)
            self._install(self.oc.build(geometry=self.geometry, faces=self.faces))
            self._ck(self.L.amrhost_set_map(self._h, n_old, len(cm), cm.ctypes.data, rev.ctypes.data))
            self.history.append(("refine", int(n)))
            return 0
        except Exception as e:      # noqa
            self._cb_error = e
            return 1

    def _on_unrefine(self, user, elems, n):
        try:
            el = np.ctypeslib.as_array(elems, shape=(n,)).copy()
            n_old = self.mesh.n
            if self.prismatic:
                inset = np.zeros(self.mesh.p, bool)
                inset[el] = True
                ep = self.mesh.edge_pts.reshape(-1, 2)
                el = np.flatnonzero(inset[ep[:, 0]] & inset[ep[:, 1]]).astype(np.int32)
            cm, rev = self.oc.unrefine(self.mesh, el)
            self._install(self.oc.build(geometry=self.geometry, faces=self.faces))
            self._ck(self.L.amrhost_set_map(self._h, n_old, len(cm), cm.ctypes.data, rev.ctypes.data))
            self.history.append(("unrefine", int(n)))
            return 0
        except Exception as e:      # noqa
            self._cb_error = e
            return 1

    # ---- solver-side API
    def set_dict(self, text: str):
        self._ck(self.L.amrhost_set_dict(self._h, text.encode()))

    def set_field(self, name: str, values: np.ndarray):
        v = np.ascontiguousarray(values, dtype=np.float64)
        ncomp = 1 if v.ndim == 1 else v.shape[1]
        self._ck(self.L.amrhost_set_field(self._h, name.encode(), ncomp, v.ctypes.data, v.shape[0]))

    def get_field(self, name: str, ncomp: int = 1) -> np.ndarray:
        nv = C.c_int64(0)
        self._ck(self.L.amrhost_get_field(self._h, name.encode(), None, 0, C.byref(nv)))
        out = np.empty(nv.value, dtype=np.float64)
        self._ck(self.L.amrhost_get_field(self._h, name.encode(), out.ctypes.data, nv.value, C.byref(nv)))
        return out if ncomp == 1 else out.reshape(-1, ncomp)

    @property
    def n_cells(self) -> int:
        return int(self.L.amrhost_n_cells(self._h))

    def update(self) -> bool:
        """runTime++; mesh.update()"""
        self.time_index += 1
        self.L.amrhost_set_time(self._h, self.time_index, float(self.time_index))
        ch = C.c_int(0)
        self._ck(self.L.amrhost_update(self._h, C.byref(ch)))
        return bool(ch.value)

    def last_lists(self):
        nr, nu = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.amrhost_last_lists(self._h, None, C.byref(nr), None, C.byref(nu)))
        cr = np.empty(max(nr.value, 1), np.int32)
        eu = np.empty(max(nu.value, 1), np.int32)
        self._ck(self.L.amrhost_last_lists(self._h, cr.ctypes.data, C.byref(nr), eu.ctypes.data, C.byref(nu)))
        return cr[:nr.value], eu[:nu.value]
