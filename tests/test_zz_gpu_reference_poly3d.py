"""GPU parity on the reference's own polyhedral fixture tests/testCases/poly3D (69 407 cells, 4-25 face neighbours per
cell, renumbered so that neighbours are far apart in label space): every face-list kernel of the path -- estimators,
patch protection, buffer layers (dense and sparse forms), 2:1 closure with and without maxSet, level check, field map --
bit for bit against the oracle.  The fixture ships owner / neighbour / boundary only, so the point-based stages are
covered by the generated meshes instead."""
import numpy as np
import pytest

from tests import np_reference as NP
from tests.fixtures import load_reference_faces, poly3d_field

pytestmark = pytest.mark.gpu
GREAT = 1e15


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


def test_poly3D_fixture(ctx, O):
    m = load_reference_faces("poly3D")
    w = O.World([m])
    ctx.set_mesh(m)
    x = poly3d_field(m.n)
    for kind, kw_o, kw_g in (("delta", {}, {}), ("codedDelta", {}, {}), ("scaledDelta", dict(min_val=1e-3, offset=0.25), dict(p0=1e-3, p1=0.25))):
        assert np.array_equal(w.estimate(kind, [x], **kw_o)[0], ctx.estimate(kind, x, **kw_g)), kind
    thr = (0.45, GREAT, 0.05, GREAT)
    e = w.estimate("delta", [x])
    w.normalize(e, *thr)
    eg = ctx.estimate("delta", x, thresholds=thr)
    assert np.array_equal(e[0], eg)
    # patch protection: 2 face layers around the only patch
    pid = [m.patch_id("all")]
    e2, e2g = e[0].copy(), eg.copy()
    assert w.protect_patches([e2], pid, 2)[0] == ctx.protect_patches(pid, 2, e2g) > 0
    assert np.array_equal(e2, e2g)
    # marking with 1 and 3 buffer layers (27 % and more of the cells marked: both forms of the layer run)
    for layers in (1, 3):
        rc, rs, _ = w.hex_select_refine(e, 2, layers)
        rc_g = np.zeros(m.n, np.uint8)
        cells, _ = ctx.select_refine(eg, 2, layers, refine_cell_out=rc_g)
        assert np.array_equal(rc[0], rc_g), layers
        assert np.array_equal(np.flatnonzero(rs[0]), cells), layers
    # a sparse marking (0.5 % of the cells) takes the push form of the layer and of the sweep
    es = np.where(np.arange(m.n) % 200 == 0, 1.0, 0.0)
    rc, rs, _ = w.hex_select_refine([es], 2, 1)
    cells, _ = ctx.select_refine(es, 2, 1)
    assert np.array_equal(np.flatnonzero(rs[0]), cells) and 0 < len(cells) < m.n // 4
    assert ctx.check_levels() == 0
    # 2:1 closure on artificial, balanced levels 0/1/2
    seed = (np.arange(m.n) % 97 == 0)
    lv = NP.grow(m, seed).astype(np.int32) + NP.grow(m, NP.grow(m, seed)).astype(np.int32)
    m.cell_level = lv
    w = O.World([m])
    ctx.set_mesh(m)
    assert ctx.check_levels() == w.check_levels([lv]) == 0
    flags = ((np.arange(m.n) % 11 == 0) & (lv == 2)).astype(np.uint8)
    for max_set in (True, False):
        ref = flags.copy()
        w.consistent_refinement([ref], max_set)
        got, _ = ctx.consistent_refinement(np.flatnonzero(flags).astype(np.int32), max_set=max_set)
        assert np.array_equal(np.flatnonzero(ref), got), max_set
    bad = lv.copy()
    bad[::13] += 3
    assert ctx.check_levels(bad) == w.check_levels([bad]) > 0
    # field map through a scattered map
    rng = np.random.default_rng(0)
    cm = rng.integers(0, m.n, m.n + 5000).astype(np.int32)
    U = rng.random((m.n, 3))
    assert np.array_equal(ctx.map_field(cm, U, ncomp=3), U[cm])
    assert np.array_equal(ctx.map_field(cm, x, ncomp=1), x[cm])
