"""Oracle restatements of the rows either side of the path against independent numpy
restatements from the definitions -- no GPU."""
import numpy as np
import pytest

from blastamr_b200 import mesh as M
from oracle import oracle as O
from tests import amr_cases as cases
from tests import np_reference as NP


def _graded(n0=6, maxlev=2, geometry=True):
    oc = cases.graded_octree(n0, maxlev)
    return oc, oc.build(geometry=geometry)


def test_label_remaps():
    rng = np.random.default_rng(3)
    n_old, n_new = 1000, 1400
    old = rng.integers(0, 5, n_old).astype(np.int32)
    rev = rng.permutation(n_new)[:n_old].astype(np.int32)
    rev[rng.random(n_old) < 0.1] = -5                # merged away
    got = O.reorder_labels(rev, n_new, old)
    ref = np.full(n_new, -1, np.int32)
    ref[rev[rev >= 0]] = old[rev >= 0]
    assert np.array_equal(got, ref)
    cm = rng.integers(-1, n_old, n_new).astype(np.int32)
    assert np.array_equal(O.gather_labels(cm, old), np.where(cm == -1, -1, old[np.maximum(cm, 0)]))
    prot = (rng.random(n_old) < 0.2).astype(np.uint8)
    assert np.array_equal(O.remap_protected_cells(cm, prot), np.where(cm >= 0, prot[np.maximum(cm, 0)], 0))


def test_gauss_gradient_and_flux_vs_numpy():
    oc, m = _graded()
    w = O.World([m])
    x = cases.front_field(m.cc, r0=0.3, w=0.1)
    mag, g = w.gradient_mag(0, x)
    g_np = NP.gauss_grad(m, x)
    assert np.allclose(g, g_np, rtol=1e-11, atol=1e-11)
    assert np.array_equal(mag, np.sqrt(g[:, 0] * g[:, 0] + g[:, 1] * g[:, 1] + g[:, 2] * g[:, 2]))
    # a linear field has an exact Gauss gradient on interior cells of the uniform part; everywhere: bounded
    lin = m.cc @ np.array([1.0, -2.0, 0.5])
    _, gl = w.gradient_mag(0, lin, xb=(m.cc[m.owner[m.f:]] @ np.array([1.0, -2.0, 0.5])))
    assert np.isfinite(gl).all()
    # weighting by V / sum(V)
    tv = float(m.vol.sum())
    magw, _ = w.gradient_mag(0, x, total_volume=tv)
    assert np.array_equal(magw, mag * (m.vol / tv))
    # flux fix-up on a refinement step
    U = np.stack([np.sin(3 * m.cc[:, 0]), np.cos(2 * m.cc[:, 1]), m.cc[:, 2] ** 2], axis=1).copy()
    cells = np.flatnonzero((m.cell_level < 2) & (np.abs(m.cc[:, 0] - 0.5) < 0.2)).astype(np.int32)
    _, rs, _ = w.hex_select_refine([np.isin(np.arange(m.n), cells).astype(np.float64)], 2, 0)
    cells = np.flatnonzero(rs[0]).astype(np.int32)
    cell_map = oc.refine(cells)
    m2 = oc.build(geometry=True)
    fm, rev = cases.face_maps(m, m2, cell_map)
    assert (fm == -1).sum() == 12 * len(cells)            # 12 inflated faces inside every split hex
    U2 = O.map_field(cell_map, U, 3)
    phi = np.full(m2.f + m2.b, 7.0)
    w2 = O.World([m2])
    nfix = w2.correct_fluxes(0, fm, rev, U2, phi)
    ref = NP.flux(m2, U2)
    touched = phi != 7.0
    assert nfix == touched.sum() > 0
    assert np.array_equal(phi[touched], ref[touched])
    # exactly: inflated faces, faces of split old faces (all descendants incl. the master)
    cnt = np.bincount(fm[fm >= 0], minlength=m.f + m.b)
    expect = (fm == -1) | ((fm >= 0) & (cnt[np.maximum(fm, 0)] > 1))
    assert np.array_equal(touched, expect)
    # a surviving face whose old face claims to be removed is the reference's FatalError
    rev_bad = rev.copy()
    rev_bad[fm[np.flatnonzero(fm >= 0)[0]]] = -1
    assert w2.correct_fluxes(0, fm, rev_bad, U2, phi.copy()) == -1


def test_least_squares_gradient_is_exact_for_linear_fields():
    oc, m = _graded()
    w = O.World([m])
    ls_p, ls_n = cases.ls_vectors(m)
    coef = np.array([1.5, -0.75, 2.0])
    lin = m.cc @ coef
    # boundary values of the linear field at the boundary face centres
    nhat = m.Sf.reshape(-1, 3)[m.f:] / m.magSf[m.f:, None]
    xb = lin[m.owner[m.f:]] + (nhat / m.delta_coeffs[m.f:, None]) @ coef
    _, g = w.gradient_mag(0, lin, scheme=1, xb=xb, ls_p=ls_p, ls_n=ls_n)
    assert np.allclose(g, coef[None, :], rtol=1e-9, atol=1e-9)


def test_mesh_size_and_max_refinement():
    oc, m = _graded()
    w = O.World([m])
    dx = w.mesh_dx(0, [1, 1, 1])
    assert np.allclose(dx, np.cbrt(m.vol), rtol=1e-15)
    dX = w.mesh_dX(0)
    h = 1.0 / 6 / (1 << m.cell_level)
    assert np.allclose(dX, np.stack([h, h, h], axis=1), rtol=1e-12)
    err = np.where(m.cc[:, 0] > 0.5, 1.0, -1.0)
    mr = w.max_refinement(0, -1, min_dx=0.05, dx=dx, error=err)
    assert np.array_equal(mr, m.cell_level + ((dx > 0.05) & (err > 0)))
    assert np.array_equal(w.max_refinement(0, 3), np.full(m.n, 3, np.int32))
    # 2-D: the face-area form; every cell has its two empty faces on the boundary -> sqrt(area)
    sides = [("xmin", "a"), ("xmax", "a"), ("ymin", "a"), ("ymax", "a"), ("zmin", "frontAndBack"), ("zmax", "frontAndBack")]
    oc2 = M.Octree(8, 8, 1, (1.0, 1.0, 0.1), 2, sides, {"frontAndBack": M.PATCH_EMPTY}, two_d=True)
    oc2.refine(np.array([9, 10, 27], np.int32))
    m2 = oc2.build(geometry=True)
    dx2 = O.World([m2]).mesh_dx(0, [1, 1, -1])
    assert np.allclose(dx2, (1.0 / 8) / (1 << m2.cell_level), rtol=1e-12)
    # cellSize estimator
    e = w.cell_size_error(0, 1, [1, 1, 1], lower_unrefine=0.05, lower_refine=0.1, dx=dx)
    assert np.array_equal(e, np.where(dx < 0.05, -1.0, np.where(dx > 0.1, 1.0, 0.0)))
    e = w.cell_size_error(0, 0, [1, 1, 1], lower_unrefine=1e-4, lower_refine=1e-3)
    assert np.array_equal(e, np.where(m.vol < 1e-4, -1.0, np.where(m.vol > 1e-3, 1.0, 0.0)))
    e = w.cell_size_error(0, 2, [1, 1, 1], lower_unrefine=0.1, lower_refine=0.2, dX=dX)
    mag = np.sqrt(dX[:, 0] ** 2 + dX[:, 1] ** 2 + dX[:, 2] ** 2)
    assert np.array_equal(e, np.where(mag < 0.1, -1.0, np.where(mag > 0.2, 1.0, 0.0)))
    e0 = np.where(m.cc[:, 1] > 0.5, 1.0, -1.0)
    e = w.cell_size_error(0, 3, [1, 1, 1], cmpts=[0, 2], min_dX=[0.05] * 3, max_dX=[0.1] * 3, dX=dX, error=e0.copy())
    cls = np.where(dX[:, 0] > 0.1, 1.0, np.where(dX[:, 0] < 0.05, -1.0, 0.0))
    assert np.array_equal(e, np.maximum(e0, cls))


@pytest.mark.parametrize("seed", [0, 1])
def test_multicomponent_wavefront_equals_in_place_loop(seed):
    """the literal ascending in-place loop (oracle) against the lower-neighbour recurrence the GPU runs"""
    oc, m = _graded(n0=5, maxlev=1, geometry=False)
    w = O.World([m])
    rng = np.random.default_rng(seed)
    blob = cases.front_field(m.cc, r0=0.3, w=0.15)
    errs = [np.where(blob + 0.3 * rng.random(m.n) > 1.6, 1.0, rng.choice([-1.0, 0.0], m.n)),
            np.where(rng.random(m.n) < 0.35, 1.0, 0.0),
            np.where(m.cc[:, 2] < 0.4, 1.0, -1.0)]
    lvls = [np.full(m.n, 2, np.int32), (m.cell_level + 1 + (rng.random(m.n) < 0.5)).astype(np.int32), np.full(m.n, 1, np.int32)]
    err, mr = w.multicomponent(0, errs, lvls)
    assert np.array_equal(err, np.maximum(np.maximum(np.maximum(-1.0, errs[0]), errs[1]), errs[2]))
    mr_wave, waves = NP.multicomponent_wavefront(m, errs, lvls)
    assert np.array_equal(mr, mr_wave)
    assert waves > 3            # the recurrence really is sequential here


def test_gpu_test_drivers_run_on_the_oracle():
    """the GPU parity tests of tests/test_gpu_next_rows.py, answered by the oracle itself: checks the test code"""
    import tests.test_gpu_next_rows as T
    be = cases.OracleBackend()
    T.test_label_remaps(be, O)
    T.test_levels_through_a_real_cut(be, O)
    for two_d in (False, True):
        T.test_gradient_range_estimator(be, O, two_d)
    T.test_flux_fixup_after_refinement(be, O)
    T.test_flux_fixup_skips_empty_patches(be, O)
    T.test_mesh_size_cell_size_and_max_refinement(be, O)
    T.test_multicomponent(be, O, 0)
    T.test_load_balance_inputs(be, O)


def test_full_size_driver_on_the_oracle():
    """tests/test_gpu_full_size.py at 24^3 with the oracle as backend: the numpy stencil / dilation references of
    that file agree with the oracle, so at 400^3 they can stand in for it"""
    import tests.test_gpu_full_size as T
    assert T.run(cases.OracleBackend(), 24) > 0


def test_poly3D_gpu_driver_on_the_oracle():
    """tests/test_zz_gpu_reference_poly3d.py answered by the oracle: checks the test code itself"""
    import tests.test_zz_gpu_reference_poly3d as T
    T.test_poly3D_fixture(cases.OracleBackend(), O)


def test_old_volumes_carry_the_merged_cells():
    """V0 / V00 through an unrefinement (fvMesh::updateMesh: direct map + volume of the merged cells injected into the
    master): total old volume is conserved, survivors keep theirs, a master holds the volume of its whole group"""
    m, cm, rev = cases.unrefinement_maps()
    v0 = m.vol * (1.0 + 0.1 * np.sin(np.arange(m.n)))
    new, n_merged = O.map_old_volumes(cm, rev, v0)
    assert n_merged == (rev < -1).sum() > 0 and n_merged % 7 == 0
    assert abs(new.sum() - v0.sum()) < 1e-12 * v0.sum()
    ref = v0[cm].copy()
    np.add.at(ref, -rev[rev < -1] - 2, v0[rev < -1])
    assert np.allclose(new, ref, rtol=1e-15, atol=0)
    untouched = np.ones(len(cm), bool); untouched[-rev[rev < -1] - 2] = False
    assert np.array_equal(new[untouched], v0[cm][untouched])
    cm2, rev2 = cases.synthetic_merge_maps()
    v = np.random.default_rng(0).random(len(rev2))
    new2, _ = O.map_old_volumes(cm2, rev2, v)
    assert abs(new2.sum() - v.sum()) < 1e-11
