"""Loaders for the committed golden fixtures (tests/golden/*.npz)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from blastamr_b200.mesh import Patch, PolyMesh, PATCH_EMPTY, PATCH_WALL

GOLDEN = Path(__file__).resolve().parent / "golden"


def add_edges(m: PolyMesh, foff, fpts) -> PolyMesh:
    """mesh.faces(), mesh.edges(), mesh.edgeCells() the way primitiveMesh derives them: an edge joins two consecutive
    points of a face; edges are numbered in order of first appearance over ascending faces; the cells of an edge are
    the cells of the faces that use it."""
    nf, f = len(foff) - 1, m.f
    face_of = np.repeat(np.arange(nf), np.diff(foff))
    nxt = np.arange(len(fpts)) + 1
    last = foff[1:] - 1
    nxt[last] = foff[:-1]
    a, b = fpts, fpts[nxt]
    lo, hi = np.minimum(a, b).astype(np.int64), np.maximum(a, b).astype(np.int64)
    key = lo * (m.p + 1) + hi
    uk, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # first appearance over ascending faces
    rank = np.empty(len(uk), np.int64)
    rank[order] = np.arange(len(uk))
    eid = rank[inv]
    ne = len(uk)
    edge_pts = np.stack([lo[first[order]], hi[first[order]]], 1).astype(np.int32)
    pairs = [np.stack([eid, m.owner[face_of]], 1)]
    internal = face_of < f
    pairs.append(np.stack([eid[internal], m.neighbour[face_of[internal]]], 1))
    ec = np.unique(np.concatenate(pairs), axis=0)
    ec_off = np.zeros(ne + 1, np.int32)
    np.add.at(ec_off, ec[:, 0] + 1, 1)
    bnd_edge = np.zeros(ne, np.uint8)
    bnd_edge[eid[~internal]] = 1
    m.n_edges, m.edge_pts, m.ec_off, m.ec_cells, m.bnd_edge = ne, edge_pts, np.cumsum(ec_off).astype(np.int32), ec[:, 1].astype(np.int32), bnd_edge
    m.face_off, m.face_pts = foff.astype(np.int32), fpts.astype(np.int32)
    return m


def load_reference_mesh(case: str, edges: bool = False) -> PolyMesh:
    """The reference's own test mesh (tests/testCases/<case>/constant/polyMesh) with the point
    addressing derived the way primitiveMesh does (pointCells / cellPoints from faces); edges=True adds
    faces() / edges() / edgeCells() (what the polyRefiner reads)."""
    z = np.load(GOLDEN / f"{case}_mesh.npz")
    owner, nei = z["owner"], z["neighbour"]
    foff, fpts, pts = z["face_off"], z["face_pts"], z["points"]
    f, nf = len(nei), len(owner)
    n, p = int(owner.max()) + 1, len(pts)
    patches = []
    for nm, ty, st, sz in zip(z["patch_names"], z["patch_types"], z["patch_start"], z["patch_size"]):
        patches.append(Patch(str(nm), PATCH_EMPTY if str(ty) == "empty" else PATCH_WALL, int(st), int(sz)))
    # point -> cells through faces
    face_of = np.repeat(np.arange(nf), np.diff(foff))
    pairs = [np.stack([fpts, owner[face_of]], 1)]
    internal = face_of < f
    pairs.append(np.stack([fpts[internal], nei[face_of[internal]]], 1))
    pc = np.unique(np.concatenate(pairs), axis=0)
    pc_off = np.zeros(p + 1, np.int32)
    np.add.at(pc_off, pc[:, 0] + 1, 1)
    pc_off = np.cumsum(pc_off).astype(np.int32)
    pc_cells = pc[:, 1].astype(np.int32)
    cp = pc[np.lexsort((pc[:, 0], pc[:, 1]))]
    cp_off = np.zeros(n + 1, np.int32)
    np.add.at(cp_off, cp[:, 1] + 1, 1)
    cp_off = np.cumsum(cp_off).astype(np.int32)
    cp_pts = cp[:, 0].astype(np.int32)
    bnd = np.zeros(p, np.uint8)
    bnd[fpts[foff[f]:]] = 1
    # cell centres = mean of the cell's points (exact for these hex blocks), volumes not needed
    cc = np.zeros((n, 3))
    np.add.at(cc, cp[:, 1], pts[cp[:, 0]])
    cc /= np.diff(cp_off)[:, None]
    m = PolyMesh(n=n, f=f, b=nf - f, p=p, owner=owner.astype(np.int32), neighbour=nei.astype(np.int32), patches=patches,
                 cc=cc, vol=None, pc_off=pc_off, pc_cells=pc_cells, cp_off=cp_off, cp_pts=cp_pts, bnd_point=bnd,
                 cell_level=np.zeros(n, np.int32), point_level=np.zeros(p, np.int32),
                 visible=np.full(n, -1, np.int32), split_parent=np.zeros(0, np.int32))
    return add_edges(m, foff, fpts) if edges else m


def load_reference_faces(case: str) -> PolyMesh:
    """A reference fixture that ships owner / neighbour / boundary only (poly3D): a mesh for the face-list kernels
    (estimators, buffer layers, 2:1 closure on a level-0 mesh, patch protection); no point addressing, no geometry."""
    z = np.load(GOLDEN / f"{case}_faces.npz")
    owner, nei = z["owner"].astype(np.int32), z["neighbour"].astype(np.int32)
    f, nf = len(nei), len(owner)
    n = int(max(owner.max(), nei.max())) + 1            # the last cells own no face: every face of theirs has a lower owner
    patches = [Patch(str(nm), PATCH_EMPTY if str(ty) == "empty" else PATCH_WALL, int(st), int(sz))
               for nm, ty, st, sz in zip(z["patch_names"], z["patch_types"], z["patch_start"], z["patch_size"])]
    return PolyMesh(n=n, f=f, b=nf - f, p=0, owner=owner, neighbour=nei, patches=patches, cc=None, vol=None,
                    pc_off=None, pc_cells=None, cp_off=None, cp_pts=None, bnd_point=None,
                    cell_level=np.zeros(n, np.int32), point_level=None, visible=None, split_parent=np.zeros(0, np.int32))


def poly3d_field(n: int) -> np.ndarray:
    """deterministic field over cell labels (the fixture has no cell centres): smooth in the label with a front"""
    c = np.arange(n, dtype=np.float64)
    return 1.0 + 0.5 * (1.0 + np.tanh((c - 0.45 * n) / (0.02 * n))) + 0.05 * np.sin(c * 0.013)


def golden():
    return np.load(GOLDEN / "oracle_golden.npz")
