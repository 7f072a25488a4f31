"""CPU: the oracle's polyRefiner functions on the reference's polyhedral fixture tests/testCases/poly2D against numpy
restatements from the definitions (tests/np_reference.py)."""
import numpy as np

from oracle import oracle as O
from tests import np_reference as NP
from tests import poly2d_case as P

THR = (0.01, 1e15, 0.01, 1e15)


def test_fixture_is_a_closed_polyhedral_complex():
    m = P.load()
    assert m.p - m.n_edges + (m.f + m.b) - m.n == 1                       # Euler
    assert (m.owner[:m.f] < m.neighbour).all()
    assert np.diff(m.ec_off).min() >= 1 and np.diff(m.ec_off).max() == 3     # prisms: at most 3 cells around an edge
    assert sorted(set(np.diff(m.face_off))) == [4, 5, 6, 7, 8, 9]


def test_estimators_vs_numpy():
    m = P.load()
    w = O.World([m])
    x = P.front(m.cc, 0.03, 0.006)
    assert np.array_equal(w.estimate("delta", [x])[0], NP.delta(m, x))
    assert np.array_equal(w.estimate("codedDelta", [x])[0], NP.delta(m, x, clamp=None))
    assert np.array_equal(w.estimate("scaledDelta", [x], min_val=1e-3, offset=0.25)[0], NP.scaled_delta(m, x, 1e-3, 0.25))


def test_virtual_refinement_loop_vs_numpy():
    """marking (3 face layers, the 2-D minimum of fvMeshPolyRefiner.C:267-300) + face/edge closure, three rounds on the
    levels the previous round left; then the unrefinement branch (point layers, prismatic and polyhedral split-point
    rules, face/edge unrefinement closure)"""
    m = P.load()
    refined = 0
    for rnd, r0 in enumerate(P.BUILD_R0):
        w = O.World([m])
        x = P.front(m.cc, r0)
        e = w.estimate("codedDelta", [x])
        w.normalize(e, *THR)
        for edge_based in (True, False):
            rc, rs, _ = w.poly_select_refine(e, 3, 3, edge_based=edge_based)
            marked = e[0] == 1
            for _ in range(3):
                marked = NP.grow(m, marked)
            assert np.array_equal(rc[0].astype(bool), marked)
            cand = marked & (m.cell_level < 3)
            assert np.array_equal(rs[0].astype(bool), NP.closure_refine_poly(m, cand, edge_based))
        rc, rs, _ = w.poly_select_refine(e, 3, 3)
        m.cell_level = (m.cell_level + rs[0]).astype(np.int32)
        assert NP.balanced_poly(m, m.cell_level)
        refined += int(rs[0].sum())
    assert refined > 100 and m.cell_level.max() == 3
    m.point_level = P.point_levels(m)
    w = O.World([m])
    assert np.array_equal(w.prism_split_points(0).astype(bool), NP.prism_split_points(m))
    assert np.array_equal(w.poly_split_points(0).astype(bool), NP.poly_split_points(m))
    x = P.front(m.cc, P.STEP_R0)                   # the front moved on: cells behind it fall below the threshold
    e = w.estimate("codedDelta", [x])
    w.normalize(e, *THR)
    rc, _, _ = w.poly_select_refine(e, 3, 3)
    n_un = 0
    for prismatic, sp in ((True, NP.prism_split_points(m)), (False, NP.poly_split_points(m))):
        for edge_based in (True, False):
            rc2 = [rc[0].copy()]
            up, spo, _ = w.poly_select_unrefine(e, rc2, 2, edge_based=edge_based, prismatic=prismatic)
            marked = NP.grow_points(m, NP.grow_points(m, rc[0].astype(bool)))
            assert np.array_equal(rc2[0].astype(bool), marked)
            pf = NP.max_cell_field(m, e[0]) < -np.sqrt(1e-15)
            pts = np.repeat(np.arange(m.p), np.diff(m.pc_off))
            anym = np.zeros(m.p, bool)
            np.logical_or.at(anym, pts, marked[m.pc_cells])
            want = NP.closure_unrefine_poly(m, sp & pf & ~anym, edge_based)
            assert np.array_equal(up[0].astype(bool), want), (prismatic, edge_based)
            n_un += int(up[0].sum())
    assert n_un > 100
