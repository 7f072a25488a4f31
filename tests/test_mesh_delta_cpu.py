"""N1 (second half): the delta that amr_mesh_update / amr_mesh_update_points take (blastamr_b200/mesh.py mesh_delta), checked on the
CPU by replaying the device's carry-over rule in numpy: a clean row is the old row of map[r] renumbered through
reverseCellMap, a dirty row comes from the host; the result must be the adjacency of the new mesh."""
import numpy as np

from blastamr_b200 import mesh as M
from oracle import oracle as O
from tests import amr_cases as cases


def cell_rows(m):
    rows = [[] for _ in range(m.n)]
    for o, q in zip(m.owner[:m.f].tolist(), m.neighbour.tolist()):
        rows[o].append(q)
        rows[q].append(o)
    return [sorted(r) for r in rows]


def point_rows(m):
    return [sorted(m.pc_cells[m.pc_off[i]:m.pc_off[i + 1]].tolist()) for i in range(m.p)]


def replay(old_rows, n_new, mp, rev, dirty, off, entries):
    slot = {int(c): i for i, c in enumerate(dirty.tolist())}
    out = []
    for r in range(n_new):
        if r in slot:
            i = slot[r]
            out.append(sorted(entries[off[i]:off[i + 1]].tolist()))
        else:
            src = old_rows[int(mp[r])]
            row = src if rev is None else [int(rev[e]) for e in src]
            assert all(e >= 0 for e in row), "a clean row points at a removed element"
            out.append(sorted(row))
    return out


def check(old, new, cm, rev):
    d = M.mesh_delta(old, new, cm, rev)
    assert np.all(np.diff(d.dirty_cells) > 0) and np.all(np.diff(d.dirty_points) > 0)
    assert replay(cell_rows(old), new.n, d.cell_map, d.reverse_cell_map, d.dirty_cells, d.dirty_cell_off, d.dirty_cell_nbrs) == cell_rows(new)
    assert replay(point_rows(old), new.p, d.point_map, d.reverse_cell_map, d.dirty_points, d.dirty_point_off, d.dirty_point_cells) == point_rows(new)
    # values carried with the rows
    clean_c = np.setdiff1d(np.arange(new.n), d.dirty_cells)
    assert np.array_equal(old.cell_level[d.cell_map[clean_c]], new.cell_level[clean_c])
    assert np.array_equal(M.parent_index(old)[d.cell_map[clean_c]] >= 0, M.parent_index(new)[clean_c] >= 0)
    clean_p = np.setdiff1d(np.arange(new.p), d.dirty_points)
    assert np.array_equal(old.point_level[d.point_map[clean_p]], new.point_level[clean_p])
    assert np.array_equal(old.bnd_point[d.point_map[clean_p]], new.bnd_point[clean_p])
    return d


def test_delta_of_a_cut_and_of_a_merge():
    oc = cases.graded_octree(8, 2)
    m0 = oc.build()
    cells = np.flatnonzero((m0.cell_level < 2) & (np.abs(m0.cc[:, 0] - 0.3) < 0.07) & (np.abs(m0.cc[:, 1] - 0.5) < 0.2)).astype(np.int32)
    w = O.World([m0])
    flags = np.zeros(m0.n, np.uint8)
    flags[cells] = 1
    w.consistent_refinement([flags])
    cells = np.flatnonzero(flags).astype(np.int32)
    assert len(cells)
    cm = oc.refine(cells)
    m1 = oc.build()
    d = check(m0, m1, cm, None)
    assert 0 < len(d.dirty_cells) < m1.n and d.nbytes() > 0
    # merge: a few valid split points of the refined mesh (2:1 closure through the oracle)
    w1 = O.World([m1])
    err = np.full(m1.n, -1.0)
    err[np.abs(m1.cc[:, 0] - 0.7) < 0.25] = 1.0
    up, _, _ = w1.hex_select_unrefine([err], [np.zeros(m1.n, np.uint8)], 0)
    pts = np.flatnonzero(up[0]).astype(np.int32)
    assert len(pts)
    cm2, rev2 = oc.unrefine(m1, pts)
    m2 = oc.build()
    d2 = check(m1, m2, cm2, rev2)
    assert 0 < len(d2.dirty_cells) < m2.n
