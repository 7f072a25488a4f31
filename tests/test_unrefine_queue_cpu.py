"""hexRef3D::consistentUnrefinement as queue iterations on a mesh-constant split-point table (kernels_sparse.cuh (3)), replayed
on the CPU with random work orders against the oracle's sweep loop (hexRef3D.C:1719-1953 via or_hex_select_unrefine):
  table   every split point i (getSplitElems) with its 8 cells; every cell belongs to at most one split point
  filter  a split point survives when none of its cells is marked for refinement and all have error < unrefineLevel
  sweep 0 every cell c of every surviving point: final level lvl[c]-1 against the final levels lvl[q]-uc[q] of its face
          neighbours; more than one below -> the point is knocked out at once (its 8 cells revert), goes to a queue
  iter j  the cells reverted by queue j against their still-marked neighbours
Flags only drop and every drop is forced, so any order reaches the reference's greatest fixed point."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import amr_cases as cases


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_split_point_queue_closure_equals_the_reference_sweeps(seed):
    rng = np.random.default_rng(seed)
    oc = cases.graded_octree(8, 2, band=0.12)
    m = oc.build()
    w = O.World([m])
    lvl = m.cell_level.astype(np.int64)
    # error: "unrefine" (-1) everywhere; refinement marks on a random part of the shell keep their cells, so that coarser
    # groups next to them cannot be merged and the closure has to knock them out
    err = np.full(m.n, -1.0)
    d = np.sqrt(((m.cc - np.array([0.5, 0.5, 0.5])) ** 2).sum(axis=1))
    rc = ((np.abs(d - 0.3) < 0.05) & (rng.random(m.n) < 0.3)).astype(np.uint8)
    up, _, sweeps = w.hex_select_unrefine([err], [rc.copy()], 0)
    ref = np.flatnonzero(up[0])
    # ---- the device algorithm
    sp = np.flatnonzero(w.split_points(0))
    cells_of = {int(p): m.pc_cells[m.pc_off[p]:m.pc_off[p + 1]].tolist() for p in sp.tolist()}
    assert all(len(v) == 8 for v in cells_of.values())
    cell_sp = {}
    for p, cs in cells_of.items():
        for c in cs:
            assert c not in cell_sp
            cell_sp[c] = p
    rows = [[] for _ in range(m.n)]
    for o, q in zip(m.owner[:m.f].tolist(), m.neighbour.tolist()):
        rows[o].append(q)
        rows[q].append(o)
    alive = {p for p, cs in cells_of.items() if all((not rc[c]) and err[c] < -1e-8 for c in cs)}
    uc = np.zeros(m.n, dtype=np.int64)
    for p in alive:
        uc[cells_of[p]] = 1
    n_filtered = len(alive)

    def knock(p, queue):
        if p in alive:
            alive.discard(p)
            uc[cells_of[p]] = 0
            queue.append(p)

    queue = []
    work = [(p, c) for p in list(alive) for c in cells_of[p]]
    rng.shuffle(work)
    for p, c in work:                                           # sweep 0
        if p in alive and lvl[c] - 1 < max(lvl[q] - uc[q] for q in rows[c]) - 1:
            knock(p, queue)
    iters = 1
    while queue:                                                # iteration j: cells reverted by queue j
        nxt = []
        work = [q for p in queue for q in cells_of[p]]
        rng.shuffle(work)
        for q in work:
            for c in rows[q]:
                if uc[c] and lvl[c] < lvl[q]:
                    knock(cell_sp[c], nxt)
        queue = nxt
        iters += 1
    assert np.array_equal(np.array(sorted(alive), dtype=np.int64), ref)
    assert 0 < len(ref) < n_filtered <= len(sp)             # the 2:1 closure itself knocked points out
    assert iters >= 2
