"""Full-size checks (BASELINE.json's 64 M-cell configuration): the oracle is too slow here, so the
CUDA path is compared with what the DOMAIN gives for free on a uniform blockMesh box -- every face neighbour of
cell (i,j,k) is (i+-1,j,k), (i,j+-1,k), (i,j,k+-1), so the estimator is a stencil over a 3-D array and a buffer
layer is a 6-neighbourhood dilation, both written in numpy from the definitions (bit-exact: max/min/abs only) --
plus size-independent properties: ascending unique lists, idempotence of the 2:1 closure, map = gather, level
update.  `AMR_FULL_SIZE_N` shrinks the box (the CPU dry run of this file uses 24 with the oracle as backend)."""
import os

import numpy as np
import pytest

from blastamr_b200 import mesh as M

pytestmark = pytest.mark.gpu

N = int(os.environ.get("AMR_FULL_SIZE_N", "400"))
THR = 0.01
GREAT = 1e15


@pytest.fixture(scope="module")
def ctx():
    from blastamr_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def front(cc, n):
    d = np.sqrt(((cc - 0.5) ** 2).sum(axis=1))
    return 1.0 + 0.5 * (1.0 + np.tanh((d - 0.3) / (2.0 / n)))


def stencil_delta(x3):
    """delta::update on a uniform box (delta.C:93-111): max over the face neighbours of min(|x_o - x_n|, 0.5)"""
    e = np.zeros_like(x3)
    for ax in range(3):
        d = np.minimum(np.abs(np.diff(x3, axis=ax)), 0.5)
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[ax] = slice(0, -1)
        hi[ax] = slice(1, None)
        np.maximum(e[tuple(lo)], d, out=e[tuple(lo)])
        np.maximum(e[tuple(hi)], d, out=e[tuple(hi)])
    return e


def dilate6(f3):
    out = f3.copy()
    for ax in range(3):
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[ax] = slice(0, -1)
        hi[ax] = slice(1, None)
        out[tuple(lo)] |= f3[tuple(hi)]
        out[tuple(hi)] |= f3[tuple(lo)]
    return out


def run(ctx, n_side):
    m = M.box_mesh(n_side, n_side, n_side, (1.0, 1.0, 1.0), points=False)
    n = m.n
    assert n == n_side ** 3
    ctx.set_mesh(m)
    x = front(m.cc, n_side)
    x3 = x.reshape(n_side, n_side, n_side)          # cell c = i + nx (j + ny k): C order over (k, j, i)
    # E: raw estimator and the normalised one, bit for bit
    e_raw = ctx.estimate("delta", x)
    ref_raw = stencil_delta(x3).ravel()
    assert np.array_equal(ref_raw, e_raw)
    e = ctx.estimate("delta", x, thresholds=(THR, GREAT, THR, GREAT))
    ref = np.where((ref_raw < THR) | (ref_raw > GREAT), -1.0, np.where((ref_raw > THR) & (ref_raw < GREAT), 1.0, 0.0))
    assert np.array_equal(ref, e)
    # M1 + M2 (one buffer layer) + M4 + B1 on a level-0 mesh: candidates dilated once; no 2:1 additions possible
    rc = np.zeros(n, np.uint8)
    cells, sweeps = ctx.select_refine(e, 3, 1, refine_cell_out=rc)
    cand = (ref == 1.0).reshape(x3.shape)
    want = dilate6(cand).ravel()
    assert np.array_equal(rc.astype(bool), want)
    assert np.array_equal(np.flatnonzero(want), cells)
    assert cells.dtype == np.int32 and (np.diff(cells) > 0).all()          # ascending, unique
    # idempotence of the closure: the balanced list is its own closure, with and without maxSet
    again, _ = ctx.consistent_refinement(cells, max_set=True)
    assert np.array_equal(again, cells)
    again, _ = ctx.consistent_refinement(cells, max_set=False)
    assert np.array_equal(again, cells)
    # two layers = two dilations (the dense and the sparse form of the layer both get used at this size)
    cells2, _ = ctx.select_refine(e, 3, 2)
    assert np.array_equal(np.flatnonzero(dilate6(dilate6(cand)).ravel()), cells2)
    # F1 + L2 through the map of this cut (parents keep their label, 7 children appended per parent)
    cell_map = np.concatenate([np.arange(n, dtype=np.int32), np.repeat(cells, 7)])
    assert np.array_equal(ctx.map_field(cell_map, x, ncomp=1), x[cell_map])
    U = np.stack([x, 2 * x, -x], axis=1).copy()
    assert np.array_equal(ctx.map_field(cell_map, U, ncomp=3), U[cell_map])
    refined = np.zeros(n, np.uint8)
    refined[cells] = 1
    lv = ctx.update_levels(cell_map, m.cell_level, refined)
    assert lv.sum() == 8 * len(cells) and set(np.unique(lv)) <= {0, 1}
    rev = np.arange(n, dtype=np.int32)
    new_rc = np.zeros(len(cell_map), np.uint8)
    ctx.remap_refine_cell(cell_map, rev, rc, new_rc)
    assert np.array_equal(new_rc[:n], rc) and (new_rc[n:] == 1).all()
    return len(cells)


def test_full_size_box(ctx):
    assert run(ctx, N) > 0
