"""CPU tests: mesh generator vs the reference's fixtures, oracle vs golden vectors, oracle vs an
independent numpy restatement, and the order-independence invariants."""
from pathlib import Path

import numpy as np
import pytest

from blastamr_b200 import mesh as M
from oracle import oracle as O
from tests import amr_cases as cases
from tests import np_reference as NP
from tests.fixtures import GOLDEN, golden, load_reference_mesh

GREAT = 1e15


def fl(a):
    return np.flatnonzero(a).astype(np.int32)


# ---------------------------------------------------------------- generator vs reference fixture

def test_box_generator_reproduces_reference_hex3D():
    z = np.load(GOLDEN / "hex3D_mesh.npz")
    m = M.box_mesh(20, 20, 5, (0.1, 0.1, 0.01))
    assert np.array_equal(m.owner, z["owner"]) and np.array_equal(m.neighbour, z["neighbour"])
    assert [q.start for q in m.patches] == list(z["patch_start"]) and [q.size for q in m.patches] == list(z["patch_size"])
    assert [q.name for q in m.patches] == [str(s) for s in z["patch_names"]]
    # same point numbering: cell centres computed from the fixture's points agree
    ref = load_reference_mesh("hex3D")
    assert np.allclose(ref.cc, m.cc, atol=1e-12)
    assert np.array_equal(ref.pc_off, m.pc_off) and np.array_equal(ref.pc_cells, m.pc_cells)
    assert np.array_equal(ref.bnd_point, m.bnd_point)


@pytest.mark.skipif(not Path("/root/reference/tests/testCases/hex3D").exists(), reason="reference tree not mounted")
def test_golden_mesh_fixture_matches_reference_tree():
    import re
    t = Path("/root/reference/tests/testCases/hex3D/constant/polyMesh/owner").read_text()
    body = t[t.index("}") + 1:]
    m_ = re.search(r"(\d+)\s*\(", body)
    own = np.array(body[m_.end():body.rindex(")")].split(), dtype=np.int32)
    assert np.array_equal(own, np.load(GOLDEN / "hex3D_mesh.npz")["owner"])


# ---------------------------------------------------------------- oracle vs golden vectors

@pytest.mark.parametrize("case", ["hex3D", "hex2D"])
def test_oracle_matches_golden_estimators(case):
    g = golden()
    m = load_reference_mesh(case)
    w = O.World([m])
    x = cases.box_field(m.cc)
    e = w.estimate("fieldValue", [x])
    w.normalize(e, 0.01, GREAT, 0.01, GREAT)
    assert np.array_equal(e[0].astype(np.int8), g[f"{case}_fieldValue_norm"])
    y = cases.front_field(m.cc, centre=(0.05, 0.05, 0.005), r0=0.03, w=0.01)
    for kind, kw in (("delta", {}), ("scaledDelta", dict(min_val=1e-3, offset=0.25)), ("codedDelta", {})):
        got = w.estimate(kind, [y], **kw)[0]
        assert np.array_equal(got.view(np.int64), g[f"{case}_{kind}_raw"].view(np.int64))


@pytest.mark.parametrize("layers", [1, 2])
@pytest.mark.parametrize("maxlev", [1, 3])
def test_oracle_matches_golden_selection(layers, maxlev):
    """the reference's own test matrix (adaptiveFvMeshTest.C:77-80) on its hex3D mesh"""
    g = golden()
    m = load_reference_mesh("hex3D")
    w = O.World([m])
    e = w.estimate("fieldValue", [cases.box_field(m.cc)])
    w.normalize(e, 0.01, GREAT, 0.01, GREAT)
    rc, rs, _ = w.hex_select_refine(e, maxlev, layers)
    assert np.array_equal(fl(rc[0]), g[f"hex3D_refineCell_L{layers}_M{maxlev}"])
    assert np.array_equal(fl(rs[0]), g[f"hex3D_cellsToRefine_L{layers}_M{maxlev}"])
    # the reference asserts only that the box gets refined (cell count grows >= 4x inside the box)
    inside = cases.box_field(m.cc) > 0
    assert rs[0][inside].all() and rs[0].sum() > inside.sum()


# ---------------------------------------------------------------- oracle vs numpy second opinion

@pytest.fixture(scope="module")
def graded():
    oc = cases.graded_octree(8, 3)
    return oc, oc.build()


def test_estimators_vs_numpy(graded):
    _, m = graded
    w = O.World([m])
    x = cases.front_field(m.cc, w=0.05) - 1.2
    assert np.array_equal(w.estimate("delta", [x])[0], NP.delta(m, x))
    assert np.array_equal(w.estimate("codedDelta", [x])[0], NP.delta(m, x, clamp=None))
    assert np.array_equal(w.estimate("scaledDelta", [x], min_val=1e-3, offset=0.2)[0], NP.scaled_delta(m, x, 1e-3, 0.2))
    assert np.array_equal(w.estimate("fieldValue", [x])[0], np.abs(x))
    e = [NP.delta(m, x)]
    ref = NP.normalize(e[0], 0.02, 0.4, 0.01, GREAT)
    w.normalize(e, 0.02, 0.4, 0.01, GREAT)
    assert np.array_equal(e[0], ref)


def test_normalize_strict_compares():
    # unrefine test wins; all four compares are strict (errorEstimator.C:208-227)
    e = [np.array([0.01, 0.02, 0.5, 0.015, 0.005, 0.6])]
    O.World.normalize(e, 0.02, 0.5, 0.01, 0.55)
    assert list(e[0]) == [0.0, 0.0, 0.0, 0.0, -1.0, -1.0]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_refine_closure_order_independent(graded, seed):
    _, m = graded
    w = O.World([m])
    rng = np.random.default_rng(seed)
    f0 = (rng.random(m.n) < 0.03).astype(np.uint8)
    gs = f0.copy()
    w.consistent_refinement([gs], True)
    jac, _ = NP.closure_refine(m, f0)
    assert np.array_equal(gs.astype(bool), jac)
    # idempotent, 2:1 after, minimal (dropping any added cell breaks 2:1)
    again = gs.copy()
    assert w.consistent_refinement([again], True) == 1 and np.array_equal(again, gs)
    lv = m.cell_level + gs
    assert (np.abs(lv[m.owner[:m.f]] - lv[m.neighbour]) <= 1).all()
    added = fl(gs.astype(bool) & ~f0.astype(bool))
    for c in added[:25]:
        t = gs.copy()
        t[c] = 0
        lv = m.cell_level + t
        assert (np.abs(lv[m.owner[:m.f]] - lv[m.neighbour]) > 1).any()
    # maxSet = false
    gs2 = f0.copy()
    w.consistent_refinement([gs2], False)
    assert np.array_equal(gs2.astype(bool), NP.closure_unselect(m, f0))


def test_extend_layers_vs_numpy(graded):
    _, m = graded
    w = O.World([m])
    rng = np.random.default_rng(5)
    err = [np.where(rng.random(m.n) < 0.02, 1.0, 0.0)]
    for layers in (0, 1, 3):
        rc, _, _ = w.hex_select_refine(err, 3, layers)
        mark = err[0] > 0
        for i in range(layers):
            src = mark & (m.cell_level < 3) if i == 0 else mark      # force=true, isTop (fvMeshRefiner.C:809-813)
            mark = NP.grow(m, mark, src)
        assert np.array_equal(rc[0].astype(bool), mark)


def test_unrefine_stage_vs_numpy(graded):
    oc, m = graded
    w = O.World([m])
    assert np.array_equal(w.split_points(0).astype(bool), NP.split_points(m))
    assert w.split_points(0).sum() > 0
    rng = np.random.default_rng(9)
    err = rng.choice([-1.0, 0.0, 1.0], size=m.n, p=[0.8, 0.1, 0.1])
    assert np.array_equal(w.max_cell_field(0, err), NP.max_cell_field(m, err))
    marked = (rng.random(m.n) < 0.01).astype(np.uint8)
    rc = marked.copy()
    up, sp, _ = w.hex_select_unrefine([err], [rc], 1)
    m1 = NP.grow(m, marked.astype(bool))
    assert np.array_equal(rc.astype(bool), m1)
    pf = NP.max_cell_field(m, err)
    pts = np.repeat(np.arange(m.p), np.diff(m.pc_off))
    any_marked = np.zeros(m.p, bool)
    np.logical_or.at(any_marked, pts, m1[m.pc_cells])
    cand = NP.split_points(m) & (pf < -np.sqrt(1e-15)) & ~any_marked
    assert np.array_equal(up[0].astype(bool), NP.closure_unrefine(m, cand))
    assert up[0].sum() > 0
    # post-condition: merging the selected points keeps 2:1 and the lists are ascending/unique
    uc = np.zeros(m.n, bool)
    uc[m.pc_cells[up[0].astype(bool)[pts]]] = True
    lv = m.cell_level - uc
    assert (np.abs(lv[m.owner[:m.f]] - lv[m.neighbour]) <= 1).all()


def test_multirank_world_equals_serial():
    """decisions do not depend on the decomposition (processor-patch swaps restated)"""
    oc = cases.graded_octree(8, 2)
    m = oc.build()
    x = cases.front_field(m.cc, w=0.05)
    for procs, R in (((2, 1, 1), 2), ((2, 2, 1), 4), ((1, 3, 2), 6)):
        loc = M.decompose(m, M.simple_ranks(m, (1, 1, 1), procs), R)
        assert sum(l.n for l in loc) == m.n
        w1, wR = O.World([m]), O.World(loc)
        for kind in ("delta", "scaledDelta"):
            e1 = w1.estimate(kind, [x], min_val=1e-3)
            eR = wR.estimate(kind, [x[l.cell_global] for l in loc], min_val=1e-3)
            for l, e in zip(loc, eR):
                assert np.array_equal(e1[0][l.cell_global], e)
        w1.normalize(e1, 0.01, GREAT, 0.01, GREAT)
        wR.normalize(eR, 0.01, GREAT, 0.01, GREAT)
        rc1, rs1, _ = w1.hex_select_refine(e1, 3, 2)
        rcR, rsR, _ = wR.hex_select_refine(eR, 3, 2)
        for l, a, b in zip(loc, rcR, rsR):
            assert np.array_equal(rc1[0][l.cell_global], a) and np.array_equal(rs1[0][l.cell_global], b)
        assert w1.check_levels([m.cell_level]) == 0 and wR.check_levels([l.cell_level for l in loc]) == 0


def test_field_map_and_levels():
    oc = M.Octree(4, 4, 4, max_level=2)
    m = oc.build()
    cells = np.array([5, 21, 22], np.int32)
    cm = oc.refine(cells)
    x = np.arange(m.n, dtype=np.float64)
    assert np.array_equal(O.map_field(cm, x), x[cm])
    v = np.stack([x, -x, 2 * x], 1).copy()
    assert np.array_equal(O.map_field(cm, v, 3), v[cm])
    ref = np.zeros(m.n, np.uint8)
    ref[cells] = 1
    lv = O.update_levels(cm, m.cell_level, ref)
    m1 = oc.build()
    assert np.array_equal(lv, m1.cell_level)
    # unrefine one parent again: direct map keeps the master's own value; conservative averages
    w = O.World([m1])
    sp = fl(w.split_points(0))
    assert len(sp) == 3
    vol = m1.vol.copy()
    y = np.random.default_rng(0).random(m1.n)
    cm2, rev = oc.unrefine(m1, sp[:1])
    d = O.map_field(cm2, y)
    assert np.array_equal(d, y[cm2])
    c = O.map_field_conservative(cm2, rev, vol, y)
    master = -rev[rev < -1][0] - 2
    merged = np.flatnonzero((rev == -master - 2) | (rev == master))
    assert len(merged) == 8
    assert np.isclose(c[master], (vol[merged] * y[merged]).sum() / vol[merged].sum(), rtol=1e-15)
    keep = np.ones(len(cm2), bool)
    keep[master] = False
    assert np.array_equal(c[keep], d[keep])
    # conservation of sum(V phi)
    m2 = oc.build()
    assert np.isclose((m2.vol * c).sum(), (vol * y).sum(), rtol=1e-13)


def test_amr_loop_on_cpu_runs_both_branches():
    """the multi-step driver used by the GPU parity tests, answered by the oracle itself"""
    import tests.test_gpu_parity as T
    oc = M.Octree(8, 8, 8, (1.0, 1.0, 1.0), 2)
    st = T.amr_loop(cases.OracleBackend(), O, oc, T.moving_front(0.06), steps=5, maxlev=2, layers_ref=1, layers_unref=1)
    assert st["refined"] > 0 and st["unrefined"] > 0


def test_amr_loop_prismatic_on_cpu_runs_both_branches():
    """the 2-D polyRefiner loop of the GPU suite, answered by the oracle (checks the driver and the 2-D topology engine)"""
    import tests.test_gpu_parity as T
    oc = M.Octree(16, 16, 1, (1.0, 1.0, 0.05), 2, T.CAVITY2D, {"frontAndBack": M.PATCH_EMPTY}, two_d=True)
    st = T.amr_loop(cases.OracleBackend(), O, oc, T.moving_front_2d(0.06), steps=6, maxlev=2, layers_ref=3, layers_unref=2,
                    kind="codedDelta", refiner="poly")
    assert st["refined"] > 0 and st["unrefined"] > 0


def _graded_quadtree(n0=12, maxlev=3):
    sides = [("xmin", "movingWall"), ("xmax", "fixedWalls"), ("ymin", "fixedWalls"), ("ymax", "fixedWalls"),
             ("zmin", "frontAndBack"), ("zmax", "frontAndBack")]
    oc = M.Octree(n0, n0, 1, (1.0, 1.0, 0.05), maxlev, sides, {"frontAndBack": M.PATCH_EMPTY}, two_d=True)
    for it in range(maxlev):
        m = oc.build()
        d = np.sqrt((m.cc[:, 0] - 0.4) ** 2 + (m.cc[:, 1] - 0.5) ** 2)
        err = [np.where((np.abs(d - 0.25) < 0.08 / (it + 1)), 1.0, 0.0)]
        _, rs, _ = O.World([m]).hex_select_refine(err, maxlev, 1)
        oc.refine(np.flatnonzero(rs[0]).astype(np.int32))
    return oc


def test_prismatic_split_points_agree_with_history_rule():
    """prismatic2DRefinement finds split points from pointLevel over the point's edges
    (prismatic2DRefinement.C:2756-2817); hexRef2D finds split edges from the refinement history
    (hexRef2D.C:2202-2342).  On the same quadtree the split points must be the end points of the split
    edges -- two independent rules, one answer.  Also V - E + F - C = 1 for the 2-D faces()/edges()."""
    oc = _graded_quadtree()
    m_all = oc.build(faces=True)        # every mesh edge (polyRefiner)
    m_thr = oc.build()                  # through-thickness edges only (2-D cutter)
    assert m_all.p - m_all.n_edges + (m_all.f + m_all.b) - m_all.n == 1
    w_all, w_thr = O.World([m_all]), O.World([m_thr])
    sp = np.flatnonzero(w_all.prism_split_points(0))
    err = [np.full(m_all.n, -1.0)]
    _, se_thr, _ = w_thr.hex2d_select_unrefine(err, [np.zeros(m_thr.n, np.uint8)], 0)
    _, se_all, _ = w_all.hex2d_select_unrefine(err, [np.zeros(m_all.n, np.uint8)], 0)
    ends_thr = np.unique(m_thr.edge_pts.reshape(-1, 2)[np.flatnonzero(se_thr[0])].ravel())
    ends_all = np.unique(m_all.edge_pts.reshape(-1, 2)[np.flatnonzero(se_all[0])].ravel())
    assert len(sp) > 0
    assert np.array_equal(sp, ends_thr)
    assert np.array_equal(sp, ends_all)     # the 2-D cutter's rule does not care about the extra in-plane edges
    # unrefinement through the prismatic engine keeps 2:1 over faces and 4:1 over edges
    up, _, _ = w_all.poly_select_unrefine(err, [np.zeros(m_all.n, np.uint8)], 0, prismatic=True)
    uc = np.zeros(m_all.n, np.int32)
    for pt in np.flatnonzero(up[0]):
        uc[m_all.pc_cells[m_all.pc_off[pt]:m_all.pc_off[pt + 1]]] = 1
    lv = m_all.cell_level - uc
    assert up[0].sum() > 0 and up[0].sum() % 2 == 0
    assert np.abs(lv[m_all.owner[:m_all.f]] - lv[m_all.neighbour]).max() <= 1
    for e in range(m_all.n_edges):
        cs = m_all.ec_cells[m_all.ec_off[e]:m_all.ec_off[e + 1]]
        assert lv[cs].max() - lv[cs].min() <= 1


def test_poly_point_layers_do_not_depend_on_the_decomposition():
    """polyRefiner across ranks: the unrefinement buffer layers spread over POINTS and their marks cross processor
    patches through syncPointList (fvMeshRefiner.C:952-958) -- gathered back to global labels, the marking is the
    same for every decomposition (the unrefinement points themselves are not: points on processor patches can never
    be split points)"""
    oc = cases.poly_balanced_octree(O)
    g = oc.build(geometry=True, faces=True)
    x = cases.front_field(g.cc, centre=(0.62, 0.5, 0.5), r0=0.3, w=0.03)
    marks = []
    for world, procs in ((2, (2, 1, 1)), (4, (2, 2, 1)), (8, (2, 2, 2))):
        loc = M.decompose(g, M.simple_ranks(g, (1, 1, 1), procs), world)
        for r, m in enumerate(loc):                   # faces()/edges() of the sub-domains are closed polyhedral complexes
            assert m.p - m.n_edges + (m.f + m.b) - m.n == 1
        w = O.World(loc)
        e = w.estimate("delta", [x[l.cell_global].copy() for l in loc])
        w.normalize(e, 0.01, 1e15, 0.01, 1e15)
        rc, rs, _ = w.poly_select_refine(e, 2, 2)
        rc2 = [r.copy() for r in rc]
        up, _, _ = w.poly_select_unrefine(e, rc2, 2)
        assert sum(int(u.sum()) for u in up) > 0
        glob = np.zeros(g.n, np.uint8)
        for l, r in zip(loc, rc2):
            glob[l.cell_global] = r
        marks.append(glob)
    assert marks[0].sum() > 0 and np.array_equal(marks[0], marks[1]) and np.array_equal(marks[0], marks[2])


def test_reference_poly3D_fixture_face_kernels_vs_numpy():
    """the reference's polyhedral fixture tests/testCases/poly3D (69 407 cells, 4-25 face neighbours per cell; owner /
    neighbour / boundary only): estimators, buffer layers and the 2:1 closure of the oracle against the numpy
    restatements from the definitions"""
    from tests.fixtures import load_reference_faces, poly3d_field
    m = load_reference_faces("poly3D")
    assert (m.n, m.f, m.b) == (69407, 437334, 28778)
    assert (m.owner[:m.f] < m.neighbour).all()                      # upper-triangular order
    w = O.World([m])
    x = poly3d_field(m.n)
    assert np.array_equal(w.estimate("delta", [x])[0], NP.delta(m, x))
    assert np.array_equal(w.estimate("codedDelta", [x])[0], NP.delta(m, x, clamp=None))
    assert np.array_equal(w.estimate("scaledDelta", [x], min_val=1e-3, offset=0.25)[0], NP.scaled_delta(m, x, 1e-3, 0.25))
    e = w.estimate("delta", [x])
    w.normalize(e, 0.45, 1e15, 0.05, 1e15)
    assert 0 < (e[0] == 1).sum() < m.n // 2
    marked = (e[0] == 1)
    rc, rs, _ = w.hex_select_refine(e, 2, 2)
    assert np.array_equal(rc[0].astype(bool), NP.grow(m, NP.grow(m, marked)))
    assert np.array_equal(rs[0], rc[0])                              # level-0 mesh: nothing for the 2:1 closure to add
    # closure on artificial levels that are 2:1 balanced by construction (level = min over a 2-ring of a seed pattern)
    seed = (np.arange(m.n) % 97 == 0)
    lv = NP.grow(m, seed).astype(np.int32) + NP.grow(m, NP.grow(m, seed)).astype(np.int32)   # 0 / 1 / 2, steps of 1
    assert np.abs(lv[m.owner[:m.f]] - lv[m.neighbour]).max() <= 1
    m.cell_level = lv
    w = O.World([m])
    flags = ((np.arange(m.n) % 11 == 0) & (lv == 2)).astype(np.uint8)
    got = flags.copy()
    w.consistent_refinement([got], True)
    assert np.array_equal(got.astype(bool), NP.closure_refine(m, flags.astype(bool))[0])
    assert got.sum() > flags.sum()
