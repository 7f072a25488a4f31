"""Edge cases of the C-ABI on the GPU: degenerate meshes, empty selections, threshold equality, error paths."""
import numpy as np
import pytest

from blastamr_b200 import mesh as M
from tests import amr_cases as cases

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


def fl(a):
    return np.flatnonzero(a).astype(np.int32)


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (3, 3, 1), (33, 1, 1), (5, 7, 3)])
def test_tiny_and_ragged_meshes(ctx, O, dims):
    """single cell (no internal faces), rows shorter than a warp slice, sizes not multiples of 32/16/4"""
    oc = M.Octree(*dims, (1.0, 1.0, 1.0), 2)
    for rounds in range(2):
        m = oc.build(geometry=True)
        w = O.World([m])
        ctx.set_mesh(m)
        rng = np.random.default_rng(dims[0] * 7 + rounds)
        x = rng.random(m.n) * 2 - 0.5
        for kind in ("delta", "scaledDelta", "codedDelta", "fieldValue"):
            assert np.array_equal(w.estimate(kind, [x], min_val=1e-3)[0], ctx.estimate(kind, x, p0=1e-3))
        if m.f > 0:
            assert np.array_equal(w.estimate_lohner([x], 0.1)[0], ctx.estimate_lohner(x, 0.1))
        e = w.estimate("delta", [x])
        w.normalize(e, 0.3, GREAT, 0.1, GREAT)
        for layers in (0, 2):
            rc, rs, _ = w.hex_select_refine(e, 2, layers)
            rc_g = np.zeros(m.n, np.uint8)
            cells, _ = ctx.select_refine(e[0], 2, layers, refine_cell_out=rc_g)
            assert np.array_equal(rc[0], rc_g) and np.array_equal(fl(rs[0]), cells)
            up, sp, _ = w.hex_select_unrefine(e, [rc[0].copy()], layers)
            pts, _ = ctx.select_unrefine(e[0], layers, refine_cell=rc_g.copy())
            assert np.array_equal(fl(up[0]), pts)
        # refine everything once so that the second round has levels, history and hanging nodes
        oc.refine(np.arange(0, m.n, 2, dtype=np.int32)[:max(1, m.n // 2)] if rounds == 0 and m.n > 1 else np.zeros(0, np.int32))


def test_empty_selections_and_saturated_levels(ctx, O):
    m = cases.graded_octree(6, 2).build()
    w = O.World([m])
    ctx.set_mesh(m)
    # nothing marked
    err = np.zeros(m.n)
    cells, sweeps = ctx.select_refine(err, 2, 3)
    assert len(cells) == 0 and sweeps == 1
    # everything marked but every cell already at maxRefinement = its own level -> nothing selectable at max 0
    err = np.ones(m.n)
    cells, _ = ctx.select_refine(err, 0, 1)
    assert len(cells) == 0
    rc, rs, _ = w.hex_select_refine([err], 0, 1)
    assert rs[0].sum() == 0
    # everything marked, per-cell maxCellLevel array
    mcl = (m.cell_level + (np.arange(m.n) % 2)).astype(np.int32)
    rc, rs, _ = w.hex_select_refine([err], [mcl], 1)
    cells, _ = ctx.select_refine(err, mcl, 1)
    assert np.array_equal(fl(rs[0]), cells) and len(cells) > 0
    # unrefinement with everything marked: nothing; with nothing marked and error -1: all consistent split points
    pts, _ = ctx.select_unrefine(-np.ones(m.n), 0, refine_cell=np.ones(m.n, np.uint8))
    assert len(pts) == 0
    up, sp, _ = w.hex_select_unrefine([-np.ones(m.n)], [np.zeros(m.n, np.uint8)], 0)
    pts, _ = ctx.select_unrefine(-np.ones(m.n), 0, refine_cell=np.zeros(m.n, np.uint8))
    assert np.array_equal(fl(up[0]), pts) and len(pts) > 0


def test_threshold_equality_is_strict(ctx, O):
    """normalize compares strictly, selectRefineCandidates inclusively (errorEstimator.C:208-227, fvMeshRefiner.C:180)"""
    m = M.box_mesh(4, 4, 4)
    ctx.set_mesh(m)
    e = np.array([0.01, 0.02, 0.5, 0.015, 0.005, 0.6] + [0.0] * (m.n - 6))
    ref = [e.copy()]
    O.World.normalize(ref, 0.02, 0.5, 0.01, 0.55)
    got = ctx.normalize(e.copy(), 0.02, 0.5, 0.01, 0.55)
    assert np.array_equal(ref[0], got)
    assert list(got[:6]) == [0.0, 0.0, 0.0, 0.0, -1.0, -1.0]
    # inclusive candidate bounds: lo == e and e == hi are candidates
    err = np.zeros(m.n)
    err[3], err[7], err[9] = 0.25, 0.75, 0.7500001
    w = O.World([m])
    rc, rs, _ = w.hex_select_refine([err], 1, 0, lower=0.25, upper=0.75)
    cells, _ = ctx.select_refine(err, 1, 0, lower=0.25, upper=0.75)
    assert list(cells) == [3, 7] and np.array_equal(fl(rs[0]), cells)


def test_error_paths(ctx):
    from blastamr_b200 import capi
    c = capi.Context(0)
    with pytest.raises(capi.AmrError, match="mesh not set"):
        c.estimate("delta", np.zeros(4))
    m = M.box_mesh(3, 3, 3)
    c.set_mesh(m)
    with pytest.raises(capi.AmrError, match="not available"):
        c.estimate(2, np.zeros(m.n))                      # Lohner through amr_estimate: needs amr_estimate_lohner
    with pytest.raises(capi.AmrError, match="geometry not set"):
        c.estimate_lohner(np.zeros(m.n), 0.1)
    with pytest.raises(capi.AmrError, match="patch id out of range"):
        c.select_refine(np.zeros(m.n), 1, 1, protected_patch_ids=[99])
    with pytest.raises(capi.AmrError, match="needs amr_mesh_set_edges"):
        c.select_unrefine(np.zeros(m.n), 1, refine_cell=np.zeros(m.n, np.uint8), kind=capi.HEX2D)
    # a mesh that violates 2:1 is reported like hexRef::checkRefinementLevels does
    lv = np.zeros(m.n, np.int32)
    lv[13] = 2
    assert c.check_levels(lv) == 6
    c.close()


def test_device_pointer_arguments(ctx, O):
    """resident path: torch CUDA tensors in and out, same results as the host-pointer path"""
    import torch
    m = cases.graded_octree(8, 2).build()
    ctx.set_mesh(m)
    x = cases.front_field(m.cc, w=0.05)
    thr = (0.01, GREAT, 0.01, GREAT)
    e_host = ctx.estimate("delta", x, thresholds=thr)
    xd = torch.from_numpy(x).cuda()
    ed = torch.empty(m.n, dtype=torch.float64, device="cuda")
    ctx.estimate("delta", xd, thresholds=thr, out=ed)
    torch.cuda.synchronize()
    assert np.array_equal(e_host, ed.cpu().numpy())
    cells_h, _ = ctx.select_refine(e_host, 2, 1)
    cd = torch.empty(m.n, dtype=torch.int32, device="cuda")
    rcd = torch.empty(m.n, dtype=torch.uint8, device="cuda")
    cnt, _ = ctx.select_refine(ed, 2, 1, cells_out=cd, refine_cell_out=rcd)
    assert np.array_equal(cells_h, cd[:cnt].cpu().numpy())
    cm = torch.cat([torch.arange(m.n, dtype=torch.int32), torch.from_numpy(np.repeat(cells_h, 7))]).cuda()
    f3 = torch.stack([xd, -xd, 2 * xd], 1).contiguous()
    out = torch.empty((cm.numel(), 3), dtype=torch.float64, device="cuda")
    ctx.map_field(cm, f3, n_old=m.n, n_new=cm.numel(), ncomp=3, out=out)
    torch.cuda.synchronize()
    assert torch.equal(out, f3[cm.long()])


@pytest.mark.parametrize("kind", ["refinement", "random"])
def test_large_host_map_uses_both_pcie_directions(ctx, kind):
    """host -> host mapping above the pipelining threshold (chunked upload/download on two streams)"""
    rng = np.random.default_rng(5)
    n_old, ncomp = 3_000_000, 3
    if kind == "refinement":
        parents = np.sort(rng.choice(n_old, 40_000, replace=False)).astype(np.int32)
        cm = np.concatenate([np.arange(n_old, dtype=np.int32), np.repeat(parents, 7)])
    else:
        cm = rng.integers(0, n_old, size=n_old + 123_457, dtype=np.int32)
    old = rng.random((n_old, ncomp))
    got = ctx.map_field(cm, old, ncomp=ncomp)
    assert np.array_equal(got, old[cm])
    got1 = ctx.map_field(cm, old[:, 0].copy(), ncomp=1)
    assert np.array_equal(got1, old[cm, 0])
    # batched entry point (fvMesh::mapFields over all registered fields), host and device buffers
    a, b = old[:, 1].copy(), old.copy()
    oa, ob = np.empty(len(cm)), np.empty((len(cm), ncomp))
    ctx.map_fields(cm, [a, b], n_old, len(cm), [1, ncomp], [oa, ob])
    assert np.array_equal(oa, a[cm]) and np.array_equal(ob, b[cm])
    import torch
    cmd, ad, bd = torch.from_numpy(cm).cuda(), torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    oad, obd = torch.empty(len(cm), dtype=torch.float64, device="cuda"), torch.empty((len(cm), ncomp), dtype=torch.float64, device="cuda")
    ctx.map_fields(cmd, [ad, bd], n_old, len(cm), [1, ncomp], [oad, obd])
    torch.cuda.synchronize()
    assert np.array_equal(oad.cpu().numpy(), oa) and np.array_equal(obd.cpu().numpy(), ob)


@pytest.mark.parametrize("n_old,n_new", [(5000, 7777), (1024, 1024), (300, 3), (70000, 2 * 1024 + 1)])
def test_fused_map_of_all_fields_and_the_marking(ctx, O, n_old, n_new):
    """amr_map_fields_and_marking on device buffers (one pass over cellMap for every field and refineCell): fields of
    1 / 3 / 6 / 9 / 2 components, ten fields (more than one launch holds), ragged tile ends, against numpy and the
    oracle's refineCell rule (fvMeshHexRefiner.C:1562-1586)"""
    import torch
    rng = np.random.default_rng(n_old + n_new)
    cm = rng.integers(0, n_old, n_new).astype(np.int32)
    keep = min(n_old, n_new) // 2
    cm[:keep] = np.arange(keep)                       # an identity run, like a real refinement map
    rev = np.full(n_old, -1, np.int32)
    rev[cm[::-1]] = np.arange(n_new - 1, -1, -1)      # each old cell -> the LOWEST new cell made from it
    rc = (rng.random(n_old) < 0.3).astype(np.uint8)
    ncs = [1, 3, 6, 9, 2, 1, 3, 1, 1, 3]
    olds = [rng.random((n_old, nc)) if nc > 1 else rng.random(n_old) for nc in ncs]
    d_old = [torch.from_numpy(a).cuda() for a in olds]
    d_new = [torch.empty((n_new,) + a.shape[1:], dtype=torch.float64, device="cuda") for a in olds]
    d_cm, d_rev, d_rc = torch.from_numpy(cm).cuda(), torch.from_numpy(rev).cuda(), torch.from_numpy(rc).cuda()
    d_out = torch.empty(n_new, dtype=torch.uint8, device="cuda")
    ctx.set_mesh(cases.graded_octree(8, 1).build())   # any mesh: the call is independent of it (flag pool sized by it)
    ctx.map_fields_and_marking(d_cm, d_old, n_old, n_new, ncs, d_new, reverse_map=d_rev, refine_cell=d_rc, out_marking=d_out)
    torch.cuda.synchronize()
    for a, d in zip(olds, d_new):
        assert np.array_equal(d.cpu().numpy(), a[cm])
    assert np.array_equal(d_out.cpu().numpy(), O.remap_refine_cell(cm, rev, rc))
    # fields only (amr_map_fields), and the host-pointer route of the same entry point
    d_new2 = [torch.zeros_like(d) for d in d_new]
    ctx.map_fields(d_cm, d_old, n_old, n_new, ncs, d_new2)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(d_new, d_new2))
    h_new = [np.empty((n_new,) + a.shape[1:]) for a in olds]
    h_out = np.empty(n_new, np.uint8)
    ctx.map_fields_and_marking(cm, olds, n_old, n_new, ncs, h_new, reverse_map=rev, refine_cell=rc, out_marking=h_out)
    assert all(np.array_equal(h, a[cm]) for h, a in zip(h_new, olds))
    assert np.array_equal(h_out, O.remap_refine_cell(cm, rev, rc))
