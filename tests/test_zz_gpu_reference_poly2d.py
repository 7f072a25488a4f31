"""GPU parity on the reference's polyhedral 2-D fixture tests/testCases/poly2D (670 prismatic polyhedra with 4-9 sided
end faces; rows of 1-3 cells per edge, 2-3 cells per point): the polyRefiner path (AMR_REFINER_POLY2D = prismatic2DRefinement
and AMR_REFINER_POLY3D = polyhedralRefinement rules) against the oracle, bit for bit, on genuinely polyhedral point /
edge / face addressing.  tests/poly2d_case.py explains the virtual refinement that makes the level fields non-trivial."""
import numpy as np
import pytest

from tests import poly2d_case as P

pytestmark = pytest.mark.gpu
THR = (0.01, 1e15, 0.01, 1e15)
POLY3D, POLY2D = 2, 3


@pytest.fixture(scope="module")
def ctx():
    from blastamr_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def test_poly2D_estimators_and_patches(ctx, O):
    m = P.load()
    w = O.World([m])
    ctx.set_mesh(m)
    x = P.front(m.cc, 0.03, 0.006)
    for kind, kw_o, kw_g in (("delta", {}, {}), ("codedDelta", {}, {}), ("scaledDelta", dict(min_val=1e-3, offset=0.25), dict(p0=1e-3, p1=0.25))):
        assert np.array_equal(w.estimate(kind, [x], **kw_o)[0], ctx.estimate(kind, x, **kw_g)), kind
    e = w.estimate("codedDelta", [x])
    w.normalize(e, *THR)
    eg = ctx.estimate("codedDelta", x, thresholds=THR)
    assert np.array_equal(e[0], eg)
    pid = [m.patch_id("sides")]
    e2, e2g = e[0].copy(), eg.copy()
    assert w.protect_patches([e2], pid, 2)[0] == ctx.protect_patches(pid, 2, e2g) > 0
    assert np.array_equal(e2, e2g)
    assert np.array_equal(ctx.max_cell_field(x), w.max_cell_field(0, x))


@pytest.mark.parametrize("layers", [1, 3])
def test_poly2D_virtual_refinement_loop(ctx, O, layers):
    m = P.load()
    refined = 0
    for r0 in P.BUILD_R0:
        w = O.World([m])
        ctx.set_mesh(m)
        x = P.front(m.cc, r0)
        e = w.estimate("codedDelta", [x])
        w.normalize(e, *THR)
        eg = ctx.estimate("codedDelta", x, thresholds=THR)
        assert np.array_equal(e[0], eg)
        for kind in (POLY2D, POLY3D):
            for edge_based in (True, False):
                ctx.set_option(0, 1 if edge_based else 0)
                rc, rs, _ = w.poly_select_refine(e, 3, layers, edge_based=edge_based)
                rc_g = np.zeros(m.n, np.uint8)
                cells, _ = ctx.select_refine(eg, 3, layers, refine_cell_out=rc_g, kind=kind)
                assert np.array_equal(rc[0], rc_g), (r0, kind, edge_based)
                assert np.array_equal(np.flatnonzero(rs[0]), cells), (r0, kind, edge_based)
        ctx.set_option(0, 1)
        rc, rs, _ = w.poly_select_refine(e, 3, layers)
        m.cell_level = (m.cell_level + rs[0]).astype(np.int32)
        refined += int(rs[0].sum())
    assert refined > 100 and m.cell_level.max() == 3
    # unrefinement branch on the graded levels, front moved inwards
    m.point_level = P.point_levels(m)
    w = O.World([m])
    ctx.set_mesh(m)
    assert ctx.check_levels() == 0
    assert np.array_equal(ctx.split_elems(kind=POLY2D), np.flatnonzero(w.prism_split_points(0)))
    assert np.array_equal(ctx.split_elems(kind=POLY3D), np.flatnonzero(w.poly_split_points(0)))
    x = P.front(m.cc, P.STEP_R0)
    e = w.estimate("codedDelta", [x])
    w.normalize(e, *THR)
    eg = ctx.estimate("codedDelta", x, thresholds=THR)
    rc, _, _ = w.poly_select_refine(e, 3, layers)
    n_un = 0
    for kind, prismatic in ((POLY2D, True), (POLY3D, False)):
        for edge_based in (True, False):
            ctx.set_option(0, 1 if edge_based else 0)
            rc_ref = [rc[0].copy()]
            up, _, _ = w.poly_select_unrefine(e, rc_ref, 2, edge_based=edge_based, prismatic=prismatic)
            rc_gpu = rc[0].copy()
            pts, _ = ctx.select_unrefine(eg, 2, refine_cell=rc_gpu, kind=kind)
            assert np.array_equal(rc_ref[0], rc_gpu), (kind, edge_based)
            assert np.array_equal(np.flatnonzero(up[0]), pts), (kind, edge_based)
            n_un += len(pts)
    ctx.set_option(0, 1)
    assert n_un > 50
