"""polyRefiner on the queue machinery: the claims behind it, checked on the CPU against the oracle.
  (1) two cells that share an internal face share an edge: the face pairs are contained in the edge pairs;
  (2) the closure of ONE rule ("a cell more than one level below a cell it shares an edge with is added") on the graph of cells
      that share an edge, run as queue iterations from the marked cells on an edge-balanced mesh, equals
      refinement::consistentRefinement = edge rule + face rule swept to convergence (refinement.C:202-387, 744-810);
  (3) the unrefinement sweep as one cell-centric rule on the same graph equals edgeConsistentUnrefinement +
      faceConsistentUnrefinement (refinement.C:390-604) swept to convergence."""
import numpy as np

from oracle import oracle as O
from tests import amr_cases as cases


def edge_graph(m):
    nb = [set() for _ in range(m.n)]
    for e in range(m.n_edges):
        cs = m.ec_cells[m.ec_off[e]:m.ec_off[e + 1]].tolist()
        for a in cs:
            for b in cs:
                if a != b:
                    nb[a].add(b)
    return nb


def test_face_pairs_are_edge_pairs_and_queue_closure_equals_the_sweeps():
    oc = cases.poly_balanced_octree(O, 8, 2)
    m = oc.build(faces=True)
    nb = edge_graph(m)
    for o, q in zip(m.owner[:m.f].tolist(), m.neighbour.tolist()):          # (1)
        assert q in nb[o] and o in nb[q]
    lvl = m.cell_level.astype(np.int64)
    for c in range(m.n):                                                    # the mesh is edge-balanced (what the queue form needs)
        assert all(abs(lvl[c] - lvl[q]) <= 1 for q in nb[c])
    # (2) candidates: a random fifth of the finest cells -- those at the rim of the refined band pull coarser cells in, which
    # pull in still coarser ones (several queue iterations)
    rng = np.random.default_rng(7)
    marks = ((lvl == lvl.max()) & (rng.random(m.n) < 0.2)).astype(np.uint8)
    ref = marks.copy()
    w = O.World([m])
    w.poly_consistent_refinement([ref], True, True)
    assert ref.sum() > marks.sum()
    st = marks.astype(bool).copy()
    queue = np.flatnonzero(st).tolist()
    iters = 0
    while queue:
        nxt = []
        for c in queue:
            for q in nb[c]:
                if lvl[c] + 1 > lvl[q] + st[q] + 1 and not st[q]:
                    st[q] = True
                    nxt.append(q)
        queue = nxt
        iters += 1
    assert np.array_equal(st.astype(np.uint8), ref) and iters >= 2
    # (3) unrefinement: cells marked for unrefinement (final level lvl - 1) against the final levels of their edge neighbours
    uc0 = ((lvl >= 1) & (rng.random(m.n) < 0.5)).astype(np.uint8)
    uc = uc0.astype(bool).copy()
    while True:
        changed = False
        for c in np.flatnonzero(uc).tolist():
            mx = max((lvl[q] - uc[q] for q in nb[c]), default=0)
            if lvl[c] - 1 < mx - 1:
                uc[c] = False
                changed = True
        if not changed:
            break
    # the reference rule, pair by pair over edges and faces, swept to convergence (refinement.C:390-604)
    ur = uc0.astype(bool).copy()
    own, nei = m.owner[:m.f].tolist(), m.neighbour.tolist()
    while True:
        changed = False
        for e in range(m.n_edges):
            cs = m.ec_cells[m.ec_off[e]:m.ec_off[e + 1]].tolist()
            for a in cs:
                for b in cs:
                    if a != b and ur[a] and (lvl[a] - 1) < (lvl[b] - ur[b]) - 1:
                        ur[a] = False
                        changed = True
        for a, b in list(zip(own, nei)) + list(zip(nei, own)):
            if ur[a] and (lvl[a] - 1) < (lvl[b] - ur[b]) - 1:
                ur[a] = False
                changed = True
        if not changed:
            break
    assert np.array_equal(uc, ur) and 0 < uc.sum() < uc0.sum()
