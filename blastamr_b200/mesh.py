"""Host-side mesh container and the synthetic topology engine (ctypes over libamrmesh.so).

`PolyMesh` carries exactly the OpenFOAM arrays the AMR decision path reads
(owner/neighbour/boundary, pointCells, cellLevel/pointLevel,
refinement-history visibleCells/parent).  `Octree` plays the role of
blockMesh + hexRef3D::setRefinement/setUnrefinement + polyTopoChange: it is host
bookkeeping that turns the *decisions* of the GPU path into the next mesh and the
`cellMap` used by the field-mapping kernel.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _build

_lib = None

SIDES = {"xmin": 0, "xmax": 1, "ymin": 2, "ymax": 3, "zmin": 4, "zmax": 5}
PATCH_WALL, PATCH_EMPTY, PATCH_PROCESSOR = 0, 1, 2


def lib():
    global _lib
    if _lib is None:
        path = _build.build_mesh()
        L = C.CDLL(str(path))
        L.mg_oct_create.restype = C.c_void_p
        L.mg_oct_create.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                                     C.c_int, C.c_void_p, C.c_void_p]
        L.mg_oct_free.argtypes = [C.c_void_p]
        L.mg_oct_set_2d.argtypes = [C.c_void_p, C.c_int]
        L.mg_oct_ncells.restype = C.c_int64
        L.mg_oct_ncells.argtypes = [C.c_void_p]
        L.mg_oct_refine.restype = C.c_int64
        L.mg_oct_refine.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.mg_oct_unrefine.restype = C.c_int64
        L.mg_oct_unrefine.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.mg_oct_build.restype = C.c_void_p
        L.mg_oct_build.argtypes = [C.c_void_p, C.c_int]
        L.mg_mesh_free.argtypes = [C.c_void_p]
        L.mg_mesh_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.mg_mesh_array.restype = C.c_void_p
        L.mg_mesh_array.argtypes = [C.c_void_p, C.c_int]
        L.mg_decompose.restype = C.c_int
        L.mg_decompose.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.mg_decompose_one.restype = C.c_void_p
        L.mg_decompose_one.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.mg_simple_ranks.argtypes = [C.c_void_p] + [C.c_double] * 3 + [C.c_int] * 3 + [C.c_void_p]
        _lib = L
    return _lib


@dataclass
class Patch:
    name: str
    type: int          # PATCH_WALL / PATCH_EMPTY / PATCH_PROCESSOR
    start: int         # absolute face index
    size: int
    nbr_rank: int = -1

    @property
    def coupled(self) -> bool:
        return self.type == PATCH_PROCESSOR


@dataclass
class PolyMesh:
    n: int
    f: int
    b: int
    p: int
    owner: np.ndarray                 # int32 [f+b]
    neighbour: np.ndarray             # int32 [f]
    patches: List[Patch]
    cc: Optional[np.ndarray] = None   # float64 [n,3]
    vol: Optional[np.ndarray] = None  # float64 [n]
    pc_off: Optional[np.ndarray] = None   # pointCells CSR
    pc_cells: Optional[np.ndarray] = None
    cp_off: Optional[np.ndarray] = None   # cellPoints CSR
    cp_pts: Optional[np.ndarray] = None
    bnd_point: Optional[np.ndarray] = None  # uint8 [p]
    cell_level: Optional[np.ndarray] = None
    point_level: Optional[np.ndarray] = None
    visible: Optional[np.ndarray] = None
    split_parent: Optional[np.ndarray] = None
    cell_global: Optional[np.ndarray] = None
    point_global: Optional[np.ndarray] = None
    patch_pt_off: Optional[np.ndarray] = None   # points of each processor patch (same order on both sides), CSR over patches
    patch_pts: Optional[np.ndarray] = None
    face_global: Optional[np.ndarray] = None
    pxyz: Optional[np.ndarray] = None
    Sf: Optional[np.ndarray] = None            # float64 [f+b,3]
    magSf: Optional[np.ndarray] = None
    weights: Optional[np.ndarray] = None       # owner-side linear weights [f+b]
    delta_coeffs: Optional[np.ndarray] = None  # nonOrthDeltaCoeffs [f+b]
    n_edges: int = 0                           # 2-D meshes: edges normal to the plane
    edge_pts: Optional[np.ndarray] = None      # int32 [n_edges,2]
    ec_off: Optional[np.ndarray] = None        # edgeCells CSR
    ec_cells: Optional[np.ndarray] = None
    bnd_edge: Optional[np.ndarray] = None      # uint8 [n_edges]
    face_off: Optional[np.ndarray] = None      # faces() CSR [f+b+1]
    face_pts: Optional[np.ndarray] = None
    _handle: Optional[int] = field(default=None, repr=False)

    # patch arrays in the layout the C-ABI takes
    def patch_arrays(self):
        st = np.array([q.start for q in self.patches], dtype=np.int32)
        sz = np.array([q.size for q in self.patches], dtype=np.int32)
        ty = np.array([q.type for q in self.patches], dtype=np.int32)
        nb = np.array([q.nbr_rank for q in self.patches], dtype=np.int32)
        return st, sz, ty, nb

    def patch_id(self, name: str) -> int:
        for i, q in enumerate(self.patches):
            if q.name == name:
                return i
        raise KeyError(name)

    def __del__(self):
        h, self._handle = self._handle, None
        if h is not None and _lib is not None:
            _lib.mg_mesh_free(h)


def _wrap(handle: int, names: Sequence[str], copy: bool = False) -> PolyMesh:
    L = lib()
    sz = np.zeros(12, dtype=np.int64)
    L.mg_mesh_sizes(handle, sz.ctypes.data)
    n, f, b, p, npatch, nsplit, nnz_pc, nnz_cp, nedge, nnz_ec, nnz_fp, nnz_pp = (int(x) for x in sz)

    def arr(idx, dtype, count, shape=None):
        ptr = L.mg_mesh_array(handle, idx)
        if not ptr or count == 0:
            return None if not ptr else np.zeros(shape or (0,), dtype=dtype)
        ct = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dtype).itemsize,))
        a = ct.view(dtype)
        if shape:
            a = a.reshape(shape)
        return a.copy() if copy else a

    pst = arr(2, np.int32, npatch); psz = arr(3, np.int32, npatch)
    pty = arr(4, np.int32, npatch); pnb = arr(5, np.int32, npatch)
    patches = []
    for i in range(npatch):
        nm = names[i] if i < len(names) else f"procBoundaryTo{int(pnb[i])}"
        patches.append(Patch(nm, int(pty[i]), int(pst[i]), int(psz[i]), int(pnb[i])))
    m = PolyMesh(
        n=n, f=f, b=b, p=p,
        owner=arr(0, np.int32, f + b), neighbour=arr(1, np.int32, f) if f else np.zeros(0, np.int32),
        patches=patches,
        cc=arr(6, np.float64, 3 * n, (n, 3)), vol=arr(7, np.float64, n),
        pc_off=arr(8, np.int32, p + 1), pc_cells=arr(9, np.int32, nnz_pc),
        cp_off=arr(10, np.int32, n + 1), cp_pts=arr(11, np.int32, nnz_cp),
        bnd_point=arr(12, np.uint8, p), cell_level=arr(13, np.int32, n), point_level=arr(14, np.int32, p),
        visible=arr(15, np.int32, n), split_parent=arr(16, np.int32, max(nsplit, 0)),
        cell_global=arr(17, np.int32, n), point_global=arr(18, np.int32, p), face_global=arr(19, np.int32, f + b),
        pxyz=arr(20, np.int32, 3 * p, (p, 3)),
        Sf=arr(21, np.float64, 3 * (f + b), (f + b, 3)), magSf=arr(22, np.float64, f + b),
        weights=arr(23, np.float64, f + b), delta_coeffs=arr(24, np.float64, f + b),
        n_edges=nedge, edge_pts=arr(25, np.int32, 2 * nedge, (nedge, 2)) if nedge else None,
        ec_off=arr(26, np.int32, nedge + 1) if nedge else None, ec_cells=arr(27, np.int32, nnz_ec) if nedge else None,
        bnd_edge=arr(28, np.uint8, nedge) if nedge else None,
        face_off=arr(29, np.int32, f + b + 1) if nnz_fp else None, face_pts=arr(30, np.int32, nnz_fp) if nnz_fp else None,
        patch_pt_off=arr(31, np.int32, npatch + 1) if nnz_pp else None, patch_pts=arr(32, np.int32, nnz_pp) if nnz_pp else None,
    )
    if m.split_parent is None:
        m.split_parent = np.zeros(0, np.int32)
    if copy:
        L.mg_mesh_free(handle)
    else:
        m._handle = handle
    return m


# default boundary layout = the reference's own fixture tests/testCases/hex3D (blockMeshDict):
#   movingWall (ymax), fixedWalls (xmin xmax ymin), frontAndBack (zmin zmax)
CAVITY_SIDES = [("ymax", "movingWall"), ("xmin", "fixedWalls"), ("xmax", "fixedWalls"), ("ymin", "fixedWalls"),
                ("zmin", "frontAndBack"), ("zmax", "frontAndBack")]


class Octree:
    """A single hex block that can be cut 1->8 and merged 8->1 (the host 'topology engine')."""

    def __init__(self, nx, ny, nz, lengths=(1.0, 1.0, 1.0), max_level=3, sides=None, patch_types=None,
                 patch_nbr=None, two_d=False):
        sides = sides or CAVITY_SIDES
        names: List[str] = []
        for _, nm in sides:
            if nm not in names:
                names.append(nm)
        patch_types = patch_types or {}
        sid = np.array([SIDES[s] for s, _ in sides], dtype=np.int32)
        spa = np.array([names.index(nm) for _, nm in sides], dtype=np.int32)
        pty = np.array([patch_types.get(nm, PATCH_WALL) for nm in names], dtype=np.int32)
        pnb = np.array([(patch_nbr or {}).get(nm, -1) for nm in names], dtype=np.int32)
        self.names = names
        self.dims = (nx, ny, nz)
        self.lengths = tuple(float(x) for x in lengths)
        self.max_level = max_level
        self._h = lib().mg_oct_create(nx, ny, nz, *self.lengths, max_level, len(sid), sid.ctypes.data,
                                      spa.ctypes.data, len(names), pty.ctypes.data, pnb.ctypes.data)
        self.two_d = bool(two_d)
        if two_d:
            lib().mg_oct_set_2d(self._h, 1)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.mg_oct_free(h)

    @property
    def ncells(self) -> int:
        return int(lib().mg_oct_ncells(self._h))

    def build(self, points: bool = True, copy: bool = False, geometry: bool = False, faces: bool = False) -> PolyMesh:
        """faces=True additionally builds mesh.faces() and mesh.edges()/edgeCells() (3-D; polyRefiner inputs)"""
        h = lib().mg_oct_build(self._h, (1 if points or faces else 0) | (2 if geometry else 0) | (4 if faces else 0))
        return _wrap(h, self.names, copy=copy)

    def refine(self, cells: np.ndarray) -> np.ndarray:
        """Cut the listed cells (ascending); returns cellMap (new -> old label)."""
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        n_new = self.ncells + (3 if self.two_d else 7) * len(cells)
        cell_map = np.empty(n_new, dtype=np.int32)
        r = lib().mg_oct_refine(self._h, cells.ctypes.data, len(cells), cell_map.ctypes.data)
        if r < 0:
            raise ValueError("refinement beyond the level cap")
        return cell_map

    def unrefine(self, mesh: PolyMesh, split_elems: np.ndarray):
        """Merge the cells around each split element of `mesh` (the mesh last built): split points
        (3-D, 8 cells) or split edges (2-D, 4 cells).
        Returns (cellMap new->old, reverseCellMap old->new | -(master)-2)."""
        split_elems = np.ascontiguousarray(split_elems, dtype=np.int32)
        n_old = self.ncells
        k = 4 if self.two_d else 8
        off, idx = (mesh.ec_off, mesh.ec_cells) if self.two_d else (mesh.pc_off, mesh.pc_cells)
        rows = np.empty((len(split_elems), k), dtype=np.int32)
        for i, el in enumerate(split_elems):
            s, e = off[el], off[el + 1]
            if e - s != k:
                raise ValueError(f"element {el} is not a split element (|cells|={e - s})")
            rows[i] = idx[s:e]
        cell_map = np.empty(n_old - (k - 1) * len(split_elems), dtype=np.int32)
        rev = np.empty(n_old, dtype=np.int32)
        r = lib().mg_oct_unrefine(self._h, rows.ctypes.data, len(split_elems), cell_map.ctypes.data, rev.ctypes.data)
        if r < 0:
            raise ValueError("inconsistent split elements")
        return cell_map, rev


def box_mesh(nx, ny, nz, lengths=(1.0, 1.0, 1.0), points=True, sides=None, patch_types=None, patch_nbr=None) -> PolyMesh:
    """blockMesh-equivalent single hex block (level 0)."""
    return Octree(nx, ny, nz, lengths, 0, sides, patch_types, patch_nbr).build(points=points)


def subdomain_sides(rank: int, procs):
    """Boundary layout of rank `rank` of a `simple (px py pz)` split of a box: outer sides are walls,
    inner sides are processor patches towards the neighbouring rank (one patch per neighbour,
    faces in the same order on both sides).  Returns (sides, patch_types, patch_nbr, (ix,iy,iz))."""
    px, py, pz = procs
    ix, iy, iz = rank % px, (rank // px) % py, rank // (px * py)
    sides, types, nbr = [], {}, {}
    walls, procsides = [], []
    for name, axis, hi, idx, cnt in (("xmin", 0, 0, ix, px), ("xmax", 0, 1, ix, px), ("ymin", 1, 0, iy, py),
                                     ("ymax", 1, 1, iy, py), ("zmin", 2, 0, iz, pz), ("zmax", 2, 1, iz, pz)):
        outer = (idx == 0) if not hi else (idx == cnt - 1)
        if outer:
            walls.append((name, "walls"))
        else:
            d = [0, 0, 0]
            d[axis] = 1 if hi else -1
            r2 = (ix + d[0]) + px * ((iy + d[1]) + py * (iz + d[2]))
            procsides.append((r2, name))
    sides = walls[:]
    types["walls"] = PATCH_WALL
    for r2, name in sorted(procsides):
        pn = f"procBoundary{rank}to{r2}"
        sides.append((name, pn))
        types[pn] = PATCH_PROCESSOR
        nbr[pn] = r2
    return sides, types, nbr, (ix, iy, iz)


def simple_ranks(mesh: PolyMesh, lengths, procs) -> np.ndarray:
    out = np.empty(mesh.n, dtype=np.int32)
    lib().mg_simple_ranks(mesh._handle, *[float(x) for x in lengths], *[int(x) for x in procs], out.ctypes.data)
    return out


def decompose(mesh: PolyMesh, cell_rank: np.ndarray, nranks: int) -> List[PolyMesh]:
    """decomposePar stand-in: one local mesh (with processor patches) per rank."""
    cell_rank = np.ascontiguousarray(cell_rank, dtype=np.int32)
    outs = (C.c_void_p * nranks)()
    if mesh._handle is None:
        raise ValueError("decompose needs a generator-owned mesh")
    lib().mg_decompose(mesh._handle, cell_rank.ctypes.data, nranks, outs)
    names = [q.name for q in mesh.patches]
    return [_wrap(outs[r], names) for r in range(nranks)]


def decompose_one(mesh: PolyMesh, cell_rank: np.ndarray, nranks: int, rank: int) -> PolyMesh:
    """the local mesh of one rank of `decompose` (a rank of a decomposed run never holds the others)"""
    cell_rank = np.ascontiguousarray(cell_rank, dtype=np.int32)
    if mesh._handle is None:
        raise ValueError("decompose needs a generator-owned mesh")
    h = lib().mg_decompose_one(mesh._handle, cell_rank.ctypes.data, nranks, rank)
    return _wrap(h, [q.name for q in mesh.patches])


# ------------------------------------------------------------------------------------------------------------------
# What a topology engine hands to amr_mesh_update / amr_mesh_update_points (include/amr_b200.h, row N1): the maps of the
# change and the rows of the addressing that changed.  polyTopoChange knows the cells it touched; this stand-in derives
# the same lists from the old and the new mesh (host bookkeeping, numpy).

@dataclass
class MeshDelta:
    cell_map: np.ndarray                    # [nNew] new -> old
    reverse_cell_map: Optional[np.ndarray]  # [nOld] old -> new (< 0 removed) or None (labels kept, cells appended)
    dirty_cells: np.ndarray
    dirty_cell_off: np.ndarray
    dirty_cell_nbrs: np.ndarray
    dirty_cell_level: np.ndarray
    dirty_parent_index: Optional[np.ndarray]
    point_map: Optional[np.ndarray] = None  # [pNew] new -> old or -1
    dirty_points: Optional[np.ndarray] = None
    dirty_point_off: Optional[np.ndarray] = None
    dirty_point_cells: Optional[np.ndarray] = None
    dirty_bnd_point: Optional[np.ndarray] = None
    dirty_point_level: Optional[np.ndarray] = None

    def nbytes(self) -> int:
        return int(sum(a.nbytes for a in self.__dict__.values() if isinstance(a, np.ndarray)))


def parent_index(mesh: PolyMesh) -> np.ndarray:
    """history.parentIndex(c): splitCells_[visibleCells_[c]].parent_ or -1 (hexRefRefinementHistory.H:299-310)"""
    vis = mesh.visible
    out = np.full(mesh.n, -1, dtype=np.int32)
    if vis is None or mesh.split_parent is None or len(mesh.split_parent) == 0:
        return out
    ok = vis >= 0
    out[ok] = mesh.split_parent[vis[ok]]
    return out


def _rows_of(cells_sorted: np.ndarray, n: int, owner_int: np.ndarray, neighbour: np.ndarray):
    """CSR of the internal-face neighbours of the listed cells (ascending), from the face lists"""
    flag = np.zeros(n, dtype=bool)
    flag[cells_sorted] = True
    mo, mn = flag[owner_int], flag[neighbour]
    a = np.concatenate([owner_int[mo], neighbour[mn]])
    b = np.concatenate([neighbour[mo], owner_int[mn]])
    order = np.lexsort((b, a))
    a, b = a[order], b[order]
    slot = np.full(n, -1, dtype=np.int64)
    slot[cells_sorted] = np.arange(len(cells_sorted))
    cnt = np.bincount(slot[a], minlength=len(cells_sorted))
    off = np.zeros(len(cells_sorted) + 1, dtype=np.int32)
    off[1:] = np.cumsum(cnt)
    return off, np.ascontiguousarray(b, dtype=np.int32)


def mesh_delta(old: PolyMesh, new: PolyMesh, cell_map: np.ndarray, reverse_cell_map: Optional[np.ndarray] = None,
               points: bool = True) -> MeshDelta:
    """old -> new through a cut (reverse_cell_map None: cells appended, labels kept) or a merge (reverse_cell_map given:
    < 0 = removed, -(master)-2).  Dirty cells: every new cell whose set of face neighbours differs from its old row."""
    n0, n1 = old.n, new.n
    cell_map = np.ascontiguousarray(cell_map, dtype=np.int32)
    own0, nei0 = old.owner[:old.f], old.neighbour
    changed_old = np.zeros(n0, dtype=bool)                    # old cells that were cut / merged away / merged into
    touched_new = np.zeros(n1, dtype=bool)
    if reverse_cell_map is None:
        added = np.arange(n0, n1)
        cut = np.unique(cell_map[added]) if len(added) else np.zeros(0, np.int64)
        changed_old[cut] = True
        touched_new[added] = True
        touched_new[cut] = True                               # the cut cell keeps its label (child 0)
        rev = None
    else:
        rev = np.ascontiguousarray(reverse_cell_map, dtype=np.int32)
        removed = rev < 0
        changed_old[removed] = True
        masters_new = (-rev[removed] - 2)
        touched_new[masters_new] = True
        changed_old[cell_map[masters_new]] = True
    # old neighbours of changed cells: their rows lose / gain entries
    m = changed_old[own0] | changed_old[nei0]
    nb_old = np.unique(np.concatenate([own0[m], nei0[m]]))
    nb_old = nb_old[~changed_old[nb_old]]
    touched_new[nb_old if rev is None else rev[nb_old]] = True
    dirty = np.flatnonzero(touched_new).astype(np.int32)
    off, nbrs = _rows_of(dirty, n1, new.owner[:new.f], new.neighbour)
    d = MeshDelta(cell_map=cell_map, reverse_cell_map=rev, dirty_cells=dirty, dirty_cell_off=off, dirty_cell_nbrs=nbrs,
                  dirty_cell_level=np.ascontiguousarray(new.cell_level[dirty], dtype=np.int32),
                  dirty_parent_index=np.ascontiguousarray(parent_index(new)[dirty], dtype=np.int32))
    if not points or new.pc_off is None or old.pc_off is None:
        return d
    # points: matched by their integer coordinates (the engine renumbers the non-lattice points on every build)
    def key(m_):
        q = m_.pxyz.astype(np.int64)
        return (q[:, 0] << 42) | (q[:, 1] << 21) | q[:, 2]
    k0, k1 = key(old), key(new)
    o0 = np.argsort(k0, kind="stable")
    k0s = k0[o0]
    o1 = np.argsort(k1, kind="stable")                        # sorted queries: the search walks k0s once (cache friendly)
    pos = np.empty(len(k1), dtype=np.int64)
    pos[o1] = np.searchsorted(k0s, k1[o1])
    pos[pos >= len(k0)] = 0
    hit = k0s[pos] == k1 if len(k0) else np.zeros(len(k1), bool)
    pmap = np.where(hit, o0[pos], -1).astype(np.int32)
    # a kept point is clean when every cell around it survived unchanged (same count, none touched)
    cnt1 = np.diff(new.pc_off)
    bad = np.zeros(new.p, dtype=bool)
    bad[pmap < 0] = True
    cells_touched = np.zeros(n1, dtype=bool)
    if rev is None:
        cells_touched[np.arange(n0, n1)] = True
        cells_touched[np.flatnonzero(changed_old)] = True
    else:
        cells_touched[masters_new] = True
    pt_of_entry = np.repeat(np.arange(new.p), cnt1)
    bad[pt_of_entry[cells_touched[new.pc_cells]]] = True
    kept = np.flatnonzero(~bad)
    bad[kept[np.diff(old.pc_off)[pmap[kept]] != cnt1[kept]]] = True
    dp = np.flatnonzero(bad).astype(np.int32)
    starts = new.pc_off[dp].astype(np.int64)
    lens = cnt1[dp].astype(np.int64)
    poff = np.zeros(len(dp) + 1, dtype=np.int32)
    poff[1:] = np.cumsum(lens)
    idx = np.repeat(starts - poff[:-1], lens) + np.arange(int(poff[-1]))
    d.point_map = pmap
    d.dirty_points = dp
    d.dirty_point_off = poff
    d.dirty_point_cells = np.ascontiguousarray(new.pc_cells[idx], dtype=np.int32)
    d.dirty_bnd_point = np.ascontiguousarray(new.bnd_point[dp], dtype=np.uint8)
    d.dirty_point_level = np.ascontiguousarray(new.point_level[dp], dtype=np.int32)
    return d
