"""GPU parity for the rows either side of the path: label remaps after a topology change (N1),
flux fix-up (N2), gradient selectors / mesh size / cellSize / multicomponent (N3).  Through the C-ABI,
bit for bit against the oracle (tolerance stated where a libm function is involved)."""
import numpy as np
import pytest

from blastamr_b200 import mesh as M
from tests import amr_cases as cases

pytestmark = pytest.mark.gpu


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


def _graded(n0=10, maxlev=2):
    oc = cases.graded_octree(n0, maxlev)
    return oc, oc.build(geometry=True)


def test_label_remaps(ctx, O):
    rng = np.random.default_rng(5)
    n_old, n_new = 100_003, 140_017
    old = rng.integers(0, 6, n_old).astype(np.int32)
    rev = rng.permutation(n_new)[:n_old].astype(np.int32)
    rev[rng.random(n_old) < 0.1] = -7
    assert np.array_equal(O.reorder_labels(rev, n_new, old), ctx.reorder_labels(rev, n_new, old))
    cm = rng.integers(-1, n_old, n_new).astype(np.int32)
    assert np.array_equal(O.gather_labels(cm, old), ctx.gather_labels(cm, old))
    prot = (rng.random(n_old) < 0.2).astype(np.uint8)
    assert np.array_equal(O.remap_protected_cells(cm, prot), ctx.remap_protected_cells(cm, prot))
    # empty lists
    e = np.zeros(0, np.int32)
    assert len(ctx.gather_labels(e, old)) == 0


def test_levels_through_a_real_cut(ctx, O):
    """hexRef::updateMesh on a real refinement: cellLevel through cellMap; pointLevel reorder through a
    reversePointMap (old points keep their labels in the topology engine -> identity + new points = -1)"""
    oc, m = _graded(8, 2)
    cells = np.flatnonzero((m.cell_level == 0) & (m.cc[:, 0] < 0.3)).astype(np.int32)
    w = O.World([m])
    f = np.zeros(m.n, np.uint8); f[cells] = 1
    w.consistent_refinement([f], True)
    cells = np.flatnonzero(f).astype(np.int32)
    cell_map = oc.refine(cells)
    m2 = oc.build()
    assert np.array_equal(ctx.gather_labels(cell_map, m.cell_level), O.gather_labels(cell_map, m.cell_level))
    rev_pt = np.arange(m.p, dtype=np.int32)
    got = ctx.reorder_labels(rev_pt, m2.p, m.point_level)
    assert np.array_equal(got, O.reorder_labels(rev_pt, m2.p, m.point_level))
    assert np.array_equal(got[:m.p], m.point_level) and (got[m.p:] == -1).all()


@pytest.mark.parametrize("two_d", [False, True])
def test_gradient_range_estimator(ctx, O, two_d):
    if two_d:
        sides = [("xmin", "a"), ("xmax", "a"), ("ymin", "a"), ("ymax", "b"), ("zmin", "frontAndBack"), ("zmax", "frontAndBack")]
        oc = M.Octree(24, 24, 1, (1.0, 1.0, 0.05), 2, sides, {"frontAndBack": M.PATCH_EMPTY}, two_d=True)
        m0 = oc.build()
        oc.refine(np.flatnonzero(np.abs(m0.cc[:, 0] - 0.5) < 0.1).astype(np.int32))
        m = oc.build(geometry=True)
    else:
        oc, m = _graded(12, 2)
    w = O.World([m])
    ctx.set_mesh(m)
    x = cases.front_field(m.cc, r0=0.3, w=0.08)
    xb = x[m.owner[m.f:]] * 1.01
    for scheme in (0, 1):
        kw = {}
        if scheme == 1:
            ls_p, ls_n = cases.ls_vectors(m)
            ctx.set_ls_vectors(ls_p, ls_n)
            kw = dict(ls_p=ls_p, ls_n=ls_n)
        for tv in (0.0, float(m.vol.sum())):
            mag, g = w.gradient_mag(0, x, scheme=scheme, xb=xb, total_volume=tv, **kw)
            e_raw, thr_g, g_gpu = ctx.estimate_gradient(x, scheme=scheme, x_boundary=xb, total_volume=tv, scale=False, want_grad=True)
            assert np.array_equal(g, g_gpu), f"scheme {scheme}: gradient differs"
            assert np.array_equal(mag, e_raw)
            thr, _ = w.range_thresholds([mag], 0.3, 0.06)
            e_ref = [mag.copy()]
            w.normalize(e_ref, *thr)
            e_gpu, thr_g, _ = ctx.estimate_gradient(x, scheme=scheme, x_boundary=xb, total_volume=tv, refine_threshold=0.3,
                                                    unrefine_threshold=0.06)
            assert np.array_equal(np.asarray(thr), thr_g)
            assert np.array_equal(e_ref[0], e_gpu)
            assert (e_gpu == 1).any() and (e_gpu == -1).any()
    # the normalised field stays resident for the marking stage
    cells, _ = ctx.select_refine(None, 2, 1)
    rc, rs, _ = w.hex_select_refine(e_ref, 2, 1)
    assert np.array_equal(np.flatnonzero(rs[0]), cells)


def test_flux_fixup_after_refinement(ctx, O):
    oc, m = _graded(8, 2)
    w = O.World([m])
    U = np.stack([np.sin(3 * m.cc[:, 0]), np.cos(2 * m.cc[:, 1]), m.cc[:, 2] ** 2], axis=1).copy()
    sel = np.isin(np.arange(m.n), np.flatnonzero((m.cell_level < 2) & (np.abs(m.cc[:, 0] - 0.5) < 0.2)))
    _, rs, _ = w.hex_select_refine([sel.astype(np.float64)], 2, 0)
    cells = np.flatnonzero(rs[0]).astype(np.int32)
    cell_map = oc.refine(cells)
    m2 = oc.build(geometry=True)
    fm, rev = cases.face_maps(m, m2, cell_map)
    U2 = O.map_field(cell_map, U, 3)
    rng = np.random.default_rng(0)
    phi0 = rng.random(m2.f + m2.b)
    Ub = rng.random((m2.b, 3))
    w2 = O.World([m2])
    ctx.set_mesh(m2)
    for ub in (None, Ub):
        phi_ref, phi_gpu = phi0.copy(), phi0.copy()
        n_ref = w2.correct_fluxes(0, fm, rev, U2, phi_ref, U_boundary=ub)
        n_gpu = ctx.correct_fluxes(fm, rev, U2, phi_gpu, U_boundary=ub)
        assert n_ref == n_gpu > 0
        assert np.array_equal(phi_ref, phi_gpu)
        assert (phi_ref != phi0).sum() > 0 and (phi_ref == phi0).sum() > 0
    rev_bad = rev.copy()
    rev_bad[fm[np.flatnonzero(fm >= 0)[0]]] = -1
    from blastamr_b200.capi import AmrError
    with pytest.raises(AmrError, match="should not have removed faces"):
        ctx.correct_fluxes(fm, rev_bad, U2, phi0.copy())


def test_flux_fixup_skips_empty_patches(ctx, O):
    sides = [("xmin", "a"), ("xmax", "a"), ("ymin", "a"), ("ymax", "b"), ("zmin", "frontAndBack"), ("zmax", "frontAndBack")]
    oc = M.Octree(10, 10, 1, (1.0, 1.0, 0.05), 2, sides, {"frontAndBack": M.PATCH_EMPTY}, two_d=True)
    m = oc.build(geometry=True)
    cells = np.flatnonzero(np.abs(m.cc[:, 0] - 0.5) < 0.2).astype(np.int32)
    cell_map = oc.refine(cells)
    m2 = oc.build(geometry=True)
    fm, rev = cases.face_maps(m, m2, cell_map)
    U2 = np.stack([m2.cc[:, 1], -m2.cc[:, 0], 1.0 + m2.cc[:, 0]], axis=1).copy()
    phi0 = np.full(m2.f + m2.b, 3.0)
    phi_ref, phi_gpu = phi0.copy(), phi0.copy()
    w2 = O.World([m2])
    ctx.set_mesh(m2)
    assert w2.correct_fluxes(0, fm, rev, U2, phi_ref) == ctx.correct_fluxes(fm, rev, U2, phi_gpu)
    assert np.array_equal(phi_ref, phi_gpu)
    st, sz, ty, _ = m2.patch_arrays()
    q = [k for k in range(len(st)) if ty[k] == 1][0]
    assert (phi_gpu[st[q]:st[q] + sz[q]] == 3.0).all()        # split front/back faces are left alone
    assert (phi_gpu != 3.0).any()


def test_mesh_size_cell_size_and_max_refinement(ctx, O):
    oc, m = _graded(10, 2)
    w = O.World([m])
    ctx.set_mesh(m)
    dx_ref, dX_ref = w.mesh_dx(0, [1, 1, 1]), w.mesh_dX(0)
    dx, dX = ctx.mesh_size([1, 1, 1])
    assert np.array_equal(dX_ref, dX)
    # 3-D dx is cbrt(V): CUDA's cbrt vs glibc's -- tolerance 1 ulp (documented in include/amr_b200.h)
    assert np.all(np.abs(dx - dx_ref) <= np.spacing(dx_ref))
    err = np.where(m.cc[:, 0] > 0.5, 1.0, -1.0)
    assert np.array_equal(w.max_refinement(0, -1, 0.03, dx_ref, err), ctx.max_refinement(-1, 0.03, dx_ref, err))
    assert np.array_equal(ctx.max_refinement(3), np.full(m.n, 3, np.int32))
    for st, kw in ((0, dict(lower_unrefine=1e-4, lower_refine=5e-4)), (1, dict(lower_unrefine=0.03, lower_refine=0.06, dx=dx_ref)),
                   (2, dict(lower_unrefine=0.06, lower_refine=0.11, dX=dX_ref))):
        e_ref = w.cell_size_error(0, st, [1, 1, 1], **kw)
        e_gpu = ctx.estimate_cell_size(st, [1, 1, 1], **kw)
        assert np.array_equal(e_ref, e_gpu), f"cellSize type {st}"
        assert len(np.unique(e_gpu)) >= 2
    e0 = np.where(m.cc[:, 1] > 0.5, 1.0, -1.0)
    kw = dict(cmpts=[0, 2], min_dX=[0.03] * 3, max_dX=[0.06] * 3, dX=dX_ref)
    assert np.array_equal(w.cell_size_error(0, 3, [1, 1, 1], error=e0.copy(), **kw),
                          ctx.estimate_cell_size(3, [1, 1, 1], error=e0.copy(), **kw))
    # 2-D: face-area form, IEEE sqrt/div only -> bitwise
    sides = [("xmin", "a"), ("xmax", "a"), ("ymin", "a"), ("ymax", "a"), ("zmin", "frontAndBack"), ("zmax", "frontAndBack")]
    oc2 = M.Octree(16, 16, 1, (1.0, 1.0, 0.1), 2, sides, {"frontAndBack": M.PATCH_EMPTY}, two_d=True)
    oc2.refine(np.array([9, 10, 27, 100], np.int32))
    m2 = oc2.build(geometry=True)
    ctx.set_mesh(m2)
    w2 = O.World([m2])
    dx2, dX2 = ctx.mesh_size([1, 1, -1])
    assert np.array_equal(w2.mesh_dx(0, [1, 1, -1]), dx2)
    assert np.array_equal(w2.mesh_dX(0), dX2)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_multicomponent(ctx, O, seed):
    oc, m = _graded(10, 2)
    w = O.World([m])
    ctx.set_mesh(m)
    rng = np.random.default_rng(seed)
    blob = cases.front_field(m.cc, r0=0.3, w=0.15)
    errs = [np.where(blob + 0.3 * rng.random(m.n) > 1.6, 1.0, rng.choice([-1.0, 0.0], m.n)),
            np.where(rng.random(m.n) < 0.35, 1.0, 0.0),
            np.where(m.cc[:, 2] < 0.4, 1.0, -1.0)]
    lvls = [np.full(m.n, 2, np.int32), (m.cell_level + 1 + (rng.random(m.n) < 0.5)).astype(np.int32), np.full(m.n, 1, np.int32)]
    e_ref, mr_ref = w.multicomponent(0, errs, lvls)
    e_gpu, mr_gpu, waves = ctx.multicomponent(errs, lvls)
    assert np.array_equal(e_ref, e_gpu)
    assert np.array_equal(mr_ref, mr_gpu)
    assert waves >= 3
    # the combined error is resident for the marking stage
    cells, _ = ctx.select_refine(None, mr_gpu, 1)
    _, rs, _ = w.hex_select_refine([e_ref], [mr_ref], 1)
    assert np.array_equal(np.flatnonzero(rs[0]), cells)
    e_only, none, _ = ctx.multicomponent(errs[:1])
    assert none is None and np.array_equal(e_only, np.maximum(-1.0, errs[0]))


def test_load_balance_inputs(ctx, O):
    """N4: imbalance metric, refinement-history clusters + blockedFace, cluster-constrained decomposition, procLoadNew"""
    oc, m = _graded(10, 3)
    w = O.World([m])
    ctx.set_mesh(m)
    assert ctx.imbalance() == w.imbalance() == (0.0, m.n)
    c2c_ref, ncl_ref = w.history_clusters(0)
    bl_ref, nu_ref = w.blocked_faces(0, c2c_ref)
    c2c, bl, ncl, nu = ctx.history_clusters()
    assert ncl == ncl_ref > 0 and nu == nu_ref > 0
    assert np.array_equal(c2c_ref, c2c) and np.array_equal(bl_ref, bl)
    # a cluster = one level-0 cell's descendants: 8^k visible cells each (k = depth), unrefined cells have none
    assert (c2c[m.cell_level == 0] == -1).all() and (c2c[m.cell_level > 0] >= 0).all()
    assert (bl[m.f:] == 1).all()
    # clusters supplied by the host (polyRefiner): same blockedFace for any relabelling
    perm = np.random.default_rng(0).permutation(ncl).astype(np.int32)
    relabelled = np.where(c2c >= 0, perm[np.maximum(c2c, 0)], -1).astype(np.int32)
    _, bl2, ncl2, nu2 = ctx.history_clusters(clusters_in=relabelled)
    assert ncl2 == -1 and nu2 == nu and np.array_equal(bl, bl2)
    # decomposition constraint: a geometric split cuts through clusters; apply() repairs it
    for nprocs in (2, 5):
        dec0 = np.minimum((m.cc[:, 0] * nprocs).astype(np.int32), nprocs - 1)
        d_ref, d_gpu = dec0.copy(), dec0.copy()
        nch_ref = w.constrain_decomposition(0, c2c_ref, ncl_ref, d_ref)
        nch_gpu = ctx.constrain_decomposition(d_gpu)
        assert nch_ref == nch_gpu
        assert np.array_equal(d_ref, d_gpu)
        for k in np.unique(c2c[c2c >= 0])[::7]:
            assert len(np.unique(d_gpu[c2c == k])) == 1
        assert np.array_equal(O.proc_load(d_gpu, nprocs), ctx.proc_load(d_gpu, nprocs))
        assert ctx.proc_load(d_gpu, nprocs).sum() == m.n
    from blastamr_b200.capi import AmrError
    with pytest.raises(AmrError, match="outside"):
        ctx.proc_load(np.full(m.n, 9, np.int32), 4)


def test_old_volumes_through_unrefinement(ctx, O):
    """amr_map_old_volumes (V0 / V00 of fvMesh::updateMesh) bit for bit against the oracle: a real hexRef3D merge (7 cells
    into the 8th), general merges with groups of up to 12 cells (slot blocks retried with the real maximum), a pure
    refinement map (nothing merged), two volume fields in one call"""
    m, cm, rev = cases.unrefinement_maps()
    v0 = m.vol * (1.0 + 0.1 * np.sin(np.arange(m.n)))
    v00 = m.vol * (1.0 + 0.2 * np.cos(np.arange(m.n)))
    (g0, g00), nm = ctx.map_old_volumes(cm, rev, [v0, v00])
    r0, nr = O.map_old_volumes(cm, rev, v0)
    r00, _ = O.map_old_volumes(cm, rev, v00)
    assert nm == nr > 0 and np.array_equal(g0, r0) and np.array_equal(g00, r00)
    assert abs(g0.sum() - v0.sum()) < 1e-12 * v0.sum()
    cm2, rev2 = cases.synthetic_merge_maps()
    v = np.random.default_rng(0).random(len(rev2))
    (g,), nm2 = ctx.map_old_volumes(cm2, rev2, [v])
    r, nr2 = O.map_old_volumes(cm2, rev2, v)
    assert nm2 == nr2 and np.array_equal(g, r)
    cm3 = np.concatenate([np.arange(100), np.repeat(np.arange(10), 7)]).astype(np.int32)
    (g3,), nm3 = ctx.map_old_volumes(cm3, np.arange(100, dtype=np.int32), [v[:100].copy()])
    assert nm3 == 0 and np.array_equal(g3, v[:100][cm3])


def _decisions(c, O, m, x, thr=(0.01, 1e15, 0.01, 1e15)):
    """the decision path on whatever mesh the context holds (estimate, marking + 2:1 closure, unrefinement, level check)"""
    e = c.estimate("delta", x, thresholds=thr)
    e2 = c.estimate("scaledDelta", x, p0=1e-3)
    rc = np.zeros(m.n, np.uint8)
    cells, _ = c.select_refine(e, 2, 2, refine_cell_out=rc)
    pts, _ = c.select_unrefine(e, 1, refine_cell=rc.copy())
    return e, e2, rc, cells.copy(), pts.copy(), c.check_levels()


def test_incremental_mesh_update_equals_the_full_hand_over(O):
    """N1, second half (include/amr_b200.h amr_mesh_update / amr_mesh_update_points): a context that carries its adjacency, levels
    and point addressing through a cut and then through a merge must decide exactly like a context that received the whole
    new mesh -- and like the oracle."""
    from blastamr_b200 import capi
    oc = cases.graded_octree(10, 2)
    m0 = oc.build()
    a, b = capi.Context(0), capi.Context(0)
    try:
        a.set_mesh(m0)
        x0 = cases.front_field(m0.cc, centre=(0.55, 0.5, 0.5), r0=0.3, w=0.04)
        e0 = a.estimate("delta", x0, thresholds=(0.01, 1e15, 0.01, 1e15))
        cells, _ = a.select_refine(e0, 2, 1)
        assert len(cells)
        cm = oc.refine(cells.copy())
        m1 = oc.build()
        d1 = M.mesh_delta(m0, m1, cm)
        assert len(d1.dirty_cells) < m1.n and len(d1.dirty_points) < m1.p
        a.update_mesh(m1, d1)
        b.set_mesh(m1)
        x1 = cases.front_field(m1.cc, centre=(0.6, 0.5, 0.5), r0=0.3, w=0.04)
        ra, rb = _decisions(a, O, m1, x1), _decisions(b, O, m1, x1)
        for u, v in zip(ra, rb):
            assert np.array_equal(u, v)
        w = O.World([m1])
        e = w.estimate("delta", [x1]); w.normalize(e, 0.01, 1e15, 0.01, 1e15)
        rc, rs, _ = w.hex_select_refine(e, 2, 2)
        up, _, _ = w.hex_select_unrefine(e, [rc[0].copy()], 1)
        assert np.array_equal(e[0], ra[0]) and np.array_equal(np.flatnonzero(rs[0]), ra[3]) and np.array_equal(np.flatnonzero(up[0]), ra[4])
        assert len(ra[4]) > 0 and ra[5] == 0
        # the merge: labels are compacted (reverseCellMap), points renumbered
        cm2, rev2 = oc.unrefine(m1, ra[4])
        m2 = oc.build()
        d2 = M.mesh_delta(m1, m2, cm2, rev2)
        a.update_mesh(m2, d2)
        b.set_mesh(m2)
        x2 = cases.front_field(m2.cc, centre=(0.5, 0.55, 0.5), r0=0.3, w=0.04)
        ra, rb = _decisions(a, O, m2, x2), _decisions(b, O, m2, x2)
        for u, v in zip(ra, rb):
            assert np.array_equal(u, v)
        assert len(ra[3]) > 0
        # what the incremental hand-over does not carry says so
        with pytest.raises(capi.AmrError):
            a.history_clusters()
        # a delta that misses a changed cell is refused (a clean row would point at a removed cell), the context recovers
        # with a full hand-over
        cm3, rev3 = oc.unrefine(m2, ra[4][:1]) if len(ra[4]) else (None, None)
        if cm3 is not None:
            m3 = oc.build()
            d3 = M.mesh_delta(m2, m3, cm3, rev3)
            master = int(-rev3[rev3 < 0][0] - 2)                     # the merged cell: its old row holds the 7 removed siblings
            keep = d3.dirty_cells != master
            offs = d3.dirty_cell_off
            sel = np.repeat(keep, np.diff(offs))
            d3.dirty_cell_nbrs = np.ascontiguousarray(d3.dirty_cell_nbrs[sel])
            d3.dirty_cell_off = np.concatenate([[0], np.cumsum(np.diff(offs)[keep])]).astype(np.int32)
            d3.dirty_cells = np.ascontiguousarray(d3.dirty_cells[keep]); d3.dirty_cell_level = np.ascontiguousarray(d3.dirty_cell_level[keep])
            d3.dirty_parent_index = np.ascontiguousarray(d3.dirty_parent_index[keep])
            with pytest.raises(capi.AmrError):
                a.update_mesh(m3, d3)
            a.set_mesh(m3)
            x3 = cases.front_field(m3.cc, centre=(0.5, 0.5, 0.55), r0=0.3, w=0.04)
            b.set_mesh(m3)
            for u, v in zip(_decisions(a, O, m3, x3), _decisions(b, O, m3, x3)):
                assert np.array_equal(u, v)
    finally:
        a.close(); b.close()
