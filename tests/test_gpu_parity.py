"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle, bit for bit."""
import numpy as np
import pytest

from blastamr_b200 import mesh as M
from tests import amr_cases as cases

pytestmark = pytest.mark.gpu

SQ = float(np.sqrt(1e-15))
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


def flags_to_list(f):
    return np.flatnonzero(f).astype(np.int32)


@pytest.mark.parametrize("kind", ["delta", "scaledDelta", "fieldValue", "codedDelta"])
@pytest.mark.parametrize("meshname", ["hex3D", "box32", "graded"])
def test_estimators_bitwise(ctx, O, kind, meshname):
    if meshname == "hex3D":
        m = M.box_mesh(20, 20, 5, (0.1, 0.1, 0.01))
        x = cases.box_field(m.cc)
    elif meshname == "box32":
        m = M.box_mesh(32, 32, 32)
        x = cases.front_field(m.cc, w=2.0 / 32)
    else:
        m = cases.graded_octree(8, 2).build()
        x = cases.front_field(m.cc, w=0.05) - 1.2   # negative values exercise mag()
    w = O.World([m])
    ctx.set_mesh(m)
    kw = dict(p0=1e-3, p1=0.25) if kind == "scaledDelta" else {}
    ref = w.estimate(kind, [x], min_val=kw.get("p0", 1e-15), offset=kw.get("p1", 0.0))[0]
    got = ctx.estimate(kind, x, **kw)
    assert np.array_equal(ref.view(np.int64), got.view(np.int64))
    thr = (0.01, GREAT, 0.01, GREAT)
    refn = ref.copy()
    w.normalize([refn], thr[0], thr[1], thr[2], thr[3])
    gotn = ctx.estimate(kind, x, thresholds=thr, **kw)
    assert np.array_equal(refn, gotn)
    g2 = ctx.normalize(got.copy(), *thr)
    assert np.array_equal(refn, g2)


@pytest.mark.parametrize("layers", [0, 1, 2, 3])
@pytest.mark.parametrize("maxlev", [1, 3])
def test_select_refine_box(ctx, O, layers, maxlev):
    m = M.box_mesh(20, 20, 5, (0.1, 0.1, 0.01))
    x = cases.box_field(m.cc)
    w = O.World([m])
    ctx.set_mesh(m)
    err = w.estimate("fieldValue", [x])
    w.normalize(err, 0.01, GREAT, 0.01, GREAT)
    rc, rs, _ = w.hex_select_refine(err, maxlev, layers)
    rc_g = np.zeros(m.n, np.uint8)
    cells, sweeps = ctx.select_refine(err[0], maxlev, layers, refine_cell_out=rc_g)
    assert np.array_equal(rc[0], rc_g)
    assert np.array_equal(flags_to_list(rs[0]), cells)


def amr_loop(ctx, O, oc, field_fn, steps, maxlev, layers_ref, layers_unref, protected=(), thr=0.01, kind="delta",
             refiner="hex"):
    """A multi-step AMR run; every stage is executed by the oracle and by the GPU and compared.
    Works for the 3-D cutter (split points) and the 2-D cutter (split edges, oc.two_d)."""
    two_d = getattr(oc, "two_d", False)
    poly = refiner == "poly"
    prismatic = poly and two_d                    # polyRefiner picks prismatic2DRefinement on 2-D meshes
    rk = (3 if two_d else 2) if poly else (1 if two_d else 0)       # AMR_REFINER_POLY2D / POLY3D / HEX2D / HEX3D
    build = (lambda: oc.build(faces=True)) if poly else oc.build
    stats = dict(refined=0, unrefined=0, sweeps=[])
    for step in range(steps):
        m = build()
        w = O.World([m])
        ctx.set_mesh(m)
        x = field_fn(m.cc, step)
        e_ref = w.estimate(kind, [x])
        e_gpu = ctx.estimate(kind, x)
        assert np.array_equal(e_ref[0], e_gpu), f"step {step}: raw error differs"
        w.normalize(e_ref, thr, GREAT, thr, GREAT)
        e_gpu = ctx.estimate(kind, x, thresholds=(thr, GREAT, thr, GREAT))
        assert np.array_equal(e_ref[0], e_gpu)
        pid = [m.patch_id(nm) for nm in protected]
        if pid:
            c_ref = w.protect_patches(e_ref, pid, 1 + (maxlev - 1) * max(layers_ref, 1))
            c_gpu = ctx.protect_patches(pid, 1 + (maxlev - 1) * max(layers_ref, 1), e_gpu)
            assert c_ref[0] == c_gpu
            assert np.array_equal(e_ref[0], e_gpu)
        if poly:
            rc, rs, sw_ref = w.poly_select_refine(e_ref, maxlev, layers_ref, protected_patch_ids=pid)
        else:
            rc, rs, sw_ref = w.hex_select_refine(e_ref, maxlev, layers_ref, protected_patch_ids=pid)
        rc_g = np.zeros(m.n, np.uint8)
        cells, sw = ctx.select_refine(e_gpu, maxlev, layers_ref, protected_patch_ids=pid, refine_cell_out=rc_g, kind=rk)
        assert np.array_equal(rc[0], rc_g), f"step {step}: refineCell differs"
        ref_cells = flags_to_list(rs[0])
        assert np.array_equal(ref_cells, cells), f"step {step}: cellsToRefine differs"
        stats["sweeps"].append((sw_ref, sw))
        refine_cell = rc[0]
        err = e_ref[0]
        fld = np.stack([x, 2 * x, -x], axis=1).copy()
        if len(cells):
            cell_map = oc.refine(cells)
            assert len(cell_map) == m.n + (3 if two_d else 7) * len(cells)
            rev = np.arange(m.n, dtype=np.int32)   # every old cell survives under its own label
            new_rc = O.remap_refine_cell(cell_map, rev, refine_cell)
            g_rc = np.zeros(len(cell_map), np.uint8)
            ctx.remap_refine_cell(cell_map, rev, refine_cell, g_rc)
            assert np.array_equal(new_rc, g_rc)
            # field + error mapping, level update
            assert np.array_equal(O.map_field(cell_map, fld, 3), ctx.map_field(cell_map, fld, ncomp=3))
            err = O.map_field(cell_map, err, 1)
            assert np.array_equal(err, ctx.map_field(cell_map, e_ref[0], ncomp=1))
            lv_ref = O.update_levels(cell_map, m.cell_level, rs[0])
            lv_gpu = ctx.update_levels(cell_map, m.cell_level, rs[0])
            assert np.array_equal(lv_ref, lv_gpu)
            m = build()
            assert np.array_equal(lv_ref, m.cell_level)     # the topology engine agrees with L2
            w = O.World([m])
            ctx.set_mesh(m)
            assert ctx.check_levels() == 0 and w.check_levels([m.cell_level]) == 0
            refine_cell = new_rc
            stats["refined"] += len(cells)
        # unrefinement on the (possibly new) mesh
        rc_ref = refine_cell.copy()
        if poly:
            up, sp, _ = w.poly_select_unrefine([err], [rc_ref], layers_unref, protected_patch_ids=pid, prismatic=prismatic)
        elif two_d:
            up, sp, _ = w.hex2d_select_unrefine([err], [rc_ref], layers_unref, protected_patch_ids=pid)
        else:
            up, sp, _ = w.hex_select_unrefine([err], [rc_ref], layers_unref, protected_patch_ids=pid)
        assert np.array_equal(flags_to_list(sp[0]), ctx.split_elems(kind=rk))
        rc_gpu = refine_cell.copy()
        pts, _ = ctx.select_unrefine(err, layers_unref, refine_cell=rc_gpu, protected_patch_ids=pid, kind=rk)
        assert np.array_equal(rc_ref, rc_gpu), f"step {step}: refineCell after unrefine layers differs"
        ref_pts = flags_to_list(up[0])
        assert np.array_equal(ref_pts, pts), f"step {step}: split points differ"
        assert np.array_equal(w.max_cell_field(0, err), ctx.max_cell_field(err))
        if len(pts):
            vol = m.vol.copy()
            n_old = m.n
            elems = pts
            if prismatic:
                # the 2-D topology engine merges around the through-thickness edge that joins the two split points
                inset = np.zeros(m.p, bool)
                inset[pts] = True
                ep = m.edge_pts.reshape(-1, 2)
                elems = np.flatnonzero(inset[ep[:, 0]] & inset[ep[:, 1]]).astype(np.int32)
                assert 2 * len(elems) == len(pts)
            cell_map, rev = oc.unrefine(m, elems)
            unref_old = np.zeros(n_old, np.uint8)
            eoff, ecells = (m.ec_off, m.ec_cells) if (two_d and not poly) else (m.pc_off, m.pc_cells)
            for pt in pts:
                unref_old[ecells[eoff[pt]:eoff[pt + 1]]] = 1
            f_old = np.stack([x, 2 * x, -x], axis=1).copy() if n_old == len(x) else np.random.default_rng(1).random((n_old, 3))
            assert np.array_equal(O.map_field(cell_map, f_old, 3), ctx.map_field(cell_map, f_old, ncomp=3))
            c_ref = O.map_field_conservative(cell_map, rev, vol, f_old, 3)
            c_gpu = ctx.map_field(cell_map, f_old, ncomp=3, mode=1, reverse_map=rev, vol_old=vol)
            assert np.array_equal(c_ref, c_gpu)
            lv_ref = O.update_levels(cell_map, m.cell_level, None, unref_old)
            lv_gpu = ctx.update_levels(cell_map, m.cell_level, None, unref_old)
            assert np.array_equal(lv_ref, lv_gpu)
            m2 = oc.build()
            assert np.array_equal(lv_ref, m2.cell_level)
            stats["unrefined"] += len(pts)
    return stats


def moving_front(speed):
    def fn(cc, step):
        return cases.front_field(cc, centre=(0.35 + speed * step, 0.5, 0.5), r0=0.22, w=0.03)
    return fn


@pytest.mark.parametrize("n0,maxlev,layers", [(8, 2, (1, 1)), (10, 3, (2, 2)), (12, 2, (3, 6))])
def test_amr_loop_moving_front(ctx, O, n0, maxlev, layers):
    oc = M.Octree(n0, n0, n0, (1.0, 1.0, 1.0), maxlev)
    st = amr_loop(ctx, O, oc, moving_front(0.06), steps=6, maxlev=maxlev, layers_ref=layers[0], layers_unref=layers[1])
    assert st["refined"] > 0 and st["unrefined"] > 0


CAVITY2D = [("ymax", "movingWall"), ("xmin", "fixedWalls"), ("xmax", "fixedWalls"), ("ymin", "fixedWalls"),
            ("zmin", "frontAndBack"), ("zmax", "frontAndBack")]


def moving_front_2d(speed):
    def fn(cc, step):
        d = np.sqrt((cc[:, 0] - (0.35 + speed * step)) ** 2 + (cc[:, 1] - 0.5) ** 2)
        return 1.0 + 0.5 * (1.0 + np.tanh((d - 0.22) / 0.03))
    return fn


@pytest.mark.parametrize("n0,maxlev,layers", [(20, 2, (1, 1)), (24, 3, (3, 6))])
def test_amr_loop_2d_hexRef2D(ctx, O, n0, maxlev, layers):
    """2-D cutter: one cell thick, `empty` front/back (the layout of the reference's hex2D fixture and of
    tutorials/damBreak2D with `refiner hexRefiner`); split elements are edges"""
    oc = M.Octree(n0, n0, 1, (1.0, 1.0, 0.05), maxlev, CAVITY2D, {"frontAndBack": M.PATCH_EMPTY}, two_d=True)
    st = amr_loop(ctx, O, oc, moving_front_2d(0.06), steps=6, maxlev=maxlev, layers_ref=layers[0], layers_unref=layers[1],
                  kind="codedDelta")
    assert st["refined"] > 0 and st["unrefined"] > 0


@pytest.mark.parametrize("n0,maxlev,layers", [(8, 2, (1, 2)), (10, 3, (2, 2))])
def test_amr_loop_polyRefiner(ctx, O, n0, maxlev, layers):
    """polyRefiner + polyhedralRefinement decision path on 3-D meshes with hanging nodes: face AND edge
    consistency (refinement.C:202-387, 390-604), point buffer layers, split points from pointLevel"""
    oc = M.Octree(n0, n0, n0, (1.0, 1.0, 1.0), maxlev)
    st = amr_loop(ctx, O, oc, moving_front(0.06), steps=6, maxlev=maxlev, layers_ref=layers[0], layers_unref=layers[1],
                  refiner="poly")
    assert st["refined"] > 0 and st["unrefined"] > 0
    # the edge rule really is stronger than the face rule on these meshes
    m = oc.build(faces=True)
    lv = m.cell_level
    for e in range(0, m.n_edges, 7):
        cs = m.ec_cells[m.ec_off[e]:m.ec_off[e + 1]]
        assert lv[cs].max() - lv[cs].min() <= 1


def test_amr_loop_polyRefiner_dense_sweeps(ctx, O, monkeypatch):
    """the polyRefiner's refinement closure has two device forms: queue iterations on the graph of cells that share an edge
    (default on edge-balanced meshes) and dense edge + face sweeps; this forces the sweeps on the same loop"""
    monkeypatch.setenv("AMR_B200_POLY_QUEUE", "0")
    oc = M.Octree(8, 8, 8, (1.0, 1.0, 1.0), 2)
    st = amr_loop(ctx, O, oc, moving_front(0.06), steps=6, maxlev=2, layers_ref=1, layers_unref=2, refiner="poly")
    assert st["refined"] > 0 and st["unrefined"] > 0


@pytest.mark.parametrize("n0,maxlev,layers", [(16, 2, (3, 2)), (20, 3, (3, 2))])
def test_amr_loop_polyRefiner_2d_prismatic(ctx, O, n0, maxlev, layers):
    """polyRefiner on a 2-D mesh = prismatic2DRefinement (fvMeshPolyRefiner.C:139-155): same marking and
    face/edge consistency, split points from pointLevel over pointEdges, only processor-patch points
    excluded (prismatic2DRefinement.C:2729-2972).  Config: tutorials/damBreak2D as shipped (polyRefiner)."""
    oc = M.Octree(n0, n0, 1, (1.0, 1.0, 0.05), maxlev, CAVITY2D, {"frontAndBack": M.PATCH_EMPTY}, two_d=True)
    st = amr_loop(ctx, O, oc, moving_front_2d(0.06), steps=6, maxlev=maxlev, layers_ref=layers[0], layers_unref=layers[1],
                  kind="codedDelta", refiner="poly")
    assert st["refined"] > 0 and st["unrefined"] > 0


def test_amr_loop_protected_patches(ctx, O):
    oc = M.Octree(10, 10, 10, (1.0, 1.0, 1.0), 2)
    st = amr_loop(ctx, O, oc, moving_front(0.08), steps=5, maxlev=2, layers_ref=1, layers_unref=2,
                  protected=("movingWall",))
    assert st["refined"] > 0


def test_consistent_refinement_standalone(ctx, O):
    m = cases.graded_octree(8, 3).build()
    ctx.set_mesh(m)
    w = O.World([m])
    rng = np.random.default_rng(7)
    for max_set in (True, False):
        flags = (rng.random(m.n) < 0.05).astype(np.uint8)
        ref = flags.copy()
        w.consistent_refinement([ref], max_set)
        got, _ = ctx.consistent_refinement(flags_to_list(flags), max_set)
        assert np.array_equal(flags_to_list(ref), got)
        # post-condition of checkWantedRefinementLevels: 2:1 holds for level + flag
        lv = m.cell_level + ref
        assert (np.abs(lv[m.owner[:m.f]] - lv[m.neighbour]) <= 1).all()


def test_protected_cells_and_quota(ctx, O):
    m = cases.graded_octree(8, 2).build()
    ctx.set_mesh(m)
    w = O.World([m])
    rng = np.random.default_rng(3)
    prot = (rng.random(m.n) < 0.02).astype(np.uint8)
    ctx.set_protected_cells(prot)
    err = [np.where(rng.random(m.n) < 0.2, 1.0, -1.0)]
    for max_cells in (2147483647, m.n + 7 * 40):
        rc, rs, _ = w.hex_select_refine(err, 3, 1, protected_cell=[prot], max_cells=max_cells)
        rc_g = np.zeros(m.n, np.uint8)
        cells, _ = ctx.select_refine(err[0], 3, 1, max_cells=max_cells, refine_cell_out=rc_g)
        assert np.array_equal(rc[0], rc_g)
        assert np.array_equal(flags_to_list(rs[0]), cells)
    ctx.set_protected_cells(None)


@pytest.mark.parametrize("meshname", ["uniform", "graded"])
def test_lohner_bitwise(ctx, O, meshname):
    """Lohner.C:71-159 incl. the restated cubic/gaussGrad arithmetic: GPU == oracle bit for bit"""
    if meshname == "uniform":
        m = M.Octree(12, 10, 8, (1.0, 0.8, 0.7), 0).build(geometry=True)
    else:
        m = cases.graded_octree(8, 3).build(geometry=True)
    w = O.World([m])
    ctx.set_mesh(m)
    x = cases.front_field(m.cc, centre=(0.45, 0.4, 0.35), r0=0.25, w=0.06)
    rng = np.random.default_rng(2)
    xb = x[m.owner[m.f:]] + 0.01 * rng.standard_normal(m.b)      # fixedValue-like boundary field
    for eps, xbv in ((0.05, None), (0.2, xb)):
        ref = w.estimate_lohner([x], eps, None if xbv is None else [xbv])[0]
        got = ctx.estimate_lohner(x, eps, x_boundary=xbv)
        assert not np.isnan(ref).any()
        assert np.array_equal(ref.view(np.int64), got.view(np.int64))
        thr = (0.3, 1e15, 0.2, 1e15)
        refn = ref.copy()
        w.normalize([refn], *thr)
        assert np.array_equal(refn, ctx.estimate_lohner(x, eps, x_boundary=xbv, thresholds=thr))


def test_gradient_range_thresholds(ctx, O):
    """gradientRange.C:102-109 on a host-computed raw error (the gradient itself stays on the host)"""
    m = cases.graded_octree(8, 2).build()
    ctx.set_mesh(m)
    w = O.World([m])
    rng = np.random.default_rng(11)
    raw = rng.random(m.n) ** 3 * 7.0 + 0.25
    for rt, ut in ((0.5, 0.4), (0.3, 0.06)):
        thr, mm = w.range_thresholds([raw], rt, ut)
        ref = raw.copy()
        w.normalize([ref], *thr)
        got, gthr, gmm = ctx.normalize_range(raw.copy(), rt, ut)
        assert np.array_equal(thr, gthr) and np.array_equal(mm, gmm)
        assert np.array_equal(ref, got)
        assert (got == 1).sum() > 0 and (got == -1).sum() > 0


def test_density_gradient_alias(ctx, O):
    """densityGradient is absent from the reference; defined here as scaledDelta's arithmetic on rho"""
    m = cases.graded_octree(8, 2).build()
    ctx.set_mesh(m)
    w = O.World([m])
    rho = cases.front_field(m.cc, w=0.05)
    assert np.array_equal(w.estimate("densityGradient", [rho], min_val=1e-6)[0], ctx.estimate("densityGradient", rho, p0=1e-6))


def test_no_cpu_fallback_message():
    from blastamr_b200 import capi
    with pytest.raises(capi.AmrError):
        capi.Context(9999)


@pytest.mark.parametrize("fraction,pull", [("1/1", "0"), ("1/4", "1/2"), ("0", "0"), ("0", "1/1")])
def test_buffer_layers_sparse_push_form(O, monkeypatch, fraction, pull):
    """the buffer layers (fvMeshRefiner.C:794-921) have a dense (pull), a sparse push (rows of the sources) and a sparse pull
    (rows of the unmarked cells) device form chosen by the marked fraction; force each on the small test meshes and run the
    whole loop, protected patches included"""
    from blastamr_b200 import capi
    monkeypatch.setenv("AMR_B200_PUSH_FRACTION", fraction)
    monkeypatch.setenv("AMR_B200_PULL_FRACTION", pull)
    monkeypatch.setenv("AMR_B200_PUSH_MIN_CELLS", "0")
    c = capi.Context(0)
    try:
        oc = M.Octree(10, 10, 10, (1.0, 1.0, 1.0), 2)
        st = amr_loop(c, O, oc, moving_front(0.08), steps=5, maxlev=2, layers_ref=2, layers_unref=3, protected=("movingWall",))
        assert st["refined"] > 0
        oc = M.Octree(20, 20, 1, (1.0, 1.0, 0.05), 2, CAVITY2D, {"frontAndBack": M.PATCH_EMPTY}, two_d=True)
        st = amr_loop(c, O, oc, moving_front_2d(0.06), steps=5, maxlev=2, layers_ref=3, layers_unref=6, kind="codedDelta")
        assert st["refined"] > 0 and st["unrefined"] > 0
    finally:
        c.close()
