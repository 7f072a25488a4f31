"""The C++ host layer (blastamr_b200/host): dictionary + run-time selection on CPU; on the GPU the
reference's own test recipe (tests/adaptiveFvMeshTests/adaptiveFvMeshTest.C) driven through
`mesh.update()` with the dynamicMeshDict keywords unchanged."""
import numpy as np
import pytest

from blastamr_b200 import hostapi, mesh as M
from tests import amr_cases as cases

DICT = """
FoamFile { version 2.0; format ascii; class dictionary; object dynamicMeshDict; }
dynamicFvMesh   adaptiveFvMesh;
// comment
errorEstimator  fieldValue;
fieldName       test;
refiner         hexRefiner;
lowerRefineLevel 0.01;
unrefineLevel   0.01;
nBufferLayers   %(layers)d;
maxRefinement   %(maxlev)d;
refineProbes    no;
protectedPatches ( %(prot)s );
nPatchesBuffers 1;
/* block
   comment */
loadBalance { balance no; }
"""


def test_dictionary_keys_and_tables():
    keys = hostapi.parse_dict_keys(DICT % dict(layers=1, maxlev=3, prot=""))
    for k in ("dynamicFvMesh", "errorEstimator", "fieldName", "refiner", "lowerRefineLevel", "unrefineLevel", "nBufferLayers",
              "maxRefinement", "refineProbes", "protectedPatches", "nPatchesBuffers", "loadBalance"):
        assert k in keys
    assert "FoamFile" not in keys
    # optional wrapper dictionary (adaptiveFvMesh.C:95-98) and verbatim code blocks
    wrapped = "adaptiveFvMeshCoeffs { errorEstimator coded; name c; code #{ error_ = 0; { } #}; maxRefinement 2; }"
    assert set(hostapi.parse_dict_keys(wrapped)) == {"errorEstimator", "name", "code", "maxRefinement"}
    est = hostapi.selection_table("errorEstimator")
    for nm in ("delta", "scaledDelta", "Lohner", "fieldValue", "gradientRange", "coded", "densityGradient"):
        assert nm in est
    assert set(hostapi.selection_table("fvMeshRefiner")) >= {"hexRefiner", "polyRefiner"}
    with pytest.raises(hostapi.FatalError):
        hostapi.parse_dict_keys("a { b 1; ")
    assert {"cellSize", "multicomponent"} <= set(est)
    # PtrList<entry> values (multicomponent.C:52): the `;` inside the entries does not end the list
    multi = "errorEstimator multicomponent; errorEstimators ( a { errorEstimator delta; deltaField p; } b { errorEstimator cellSize; type volume; } ); maxRefinement 2;"
    assert set(hostapi.parse_dict_keys(multi)) == {"errorEstimator", "errorEstimators", "maxRefinement"}


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hostapi.FatalError, match="no usable CUDA device"):
        hostapi.AdaptiveFvMesh(DICT % dict(layers=1, maxlev=1, prot=""), M.Octree(4, 4, 4, max_level=1))


@pytest.mark.gpu
@pytest.mark.parametrize("layers", [1, 2])
@pytest.mark.parametrize("maxlev", [1, 3])
def test_reference_recipe_through_update(layers, maxlev):
    """adaptiveFvMeshTest.C:62-233 with hexRefiner on the 3-D fixture block: field = 1 inside the box,
    update() maxRefinement times -> the box is refined; zero the field -> cells are merged again."""
    oc = M.Octree(20, 20, 5, (0.1, 0.1, 0.01), maxlev)
    am = hostapi.AdaptiveFvMesh(DICT % dict(layers=layers, maxlev=maxlev, prot=""), oc)
    n0 = am.n_cells
    box0 = int(cases.box_field(am.mesh.cc).sum())
    for _ in range(maxlev):
        am.set_field("test", cases.box_field(am.mesh.cc))
        am.set_field("U", np.stack([am.mesh.cc[:, 0], am.mesh.cc[:, 1], am.mesh.cc[:, 2]], 1))
        changed = am.update()
        assert changed
        # registered fields are mapped by direct injection: children carry the parent value
        assert am.get_field("U", 3).shape == (am.n_cells, 3)
        assert am.get_field("test").shape == (am.n_cells,)
    n1 = am.n_cells
    box1 = int(cases.box_field(am.mesh.cc).sum())
    assert box1 >= 4 * box0 and n1 > n0                       # adaptiveFvMeshTest.C:205
    assert am.mesh.cell_level.max() == maxlev
    lv = am.mesh.cell_level
    assert (np.abs(lv[am.mesh.owner[:am.mesh.f]] - lv[am.mesh.neighbour]) <= 1).all()
    # unrefine: field = 0 everywhere
    for _ in range(maxlev + 2):
        am.set_field("test", np.zeros(am.n_cells))
        am.update()
    assert am.n_cells < n1                                    # adaptiveFvMeshTest.C:229
    assert any(k == "unrefine" for k, _ in am.history)
    am.close()


@pytest.mark.gpu
def test_update_matches_oracle_driver():
    """mesh.update() (C++ host classes -> C-ABI -> CUDA) against the oracle run step by step"""
    from oracle import oracle as O
    maxlev, layers = 2, 1
    text = DICT.replace("errorEstimator  fieldValue;", "errorEstimator  delta;").replace("fieldName       test;", "deltaField      rho;")
    text = text % dict(layers=layers, maxlev=maxlev, prot="movingWall")
    oc = M.Octree(10, 10, 10, (1.0, 1.0, 1.0), maxlev)
    oc_ref = M.Octree(10, 10, 10, (1.0, 1.0, 1.0), maxlev)
    am = hostapi.AdaptiveFvMesh(text, oc)
    thr = (0.01, 1e15, 0.01, 1e15)
    for step in range(5):
        m = oc_ref.build()
        assert m.n == am.n_cells and np.array_equal(m.owner, am.mesh.owner)
        rho = cases.front_field(m.cc, centre=(0.35 + 0.07 * step, 0.5, 0.5), r0=0.22, w=0.03)
        am.set_field("rho", rho)
        am.update()
        cells_g, pts_g = am.last_lists()
        # oracle, same control flow as adaptiveFvMesh::refine / fvMeshHexRefiner::refine
        w = O.World([m])
        e = w.estimate("delta", [rho]); w.normalize(e, *thr)
        pid = [m.patch_id("movingWall")]
        w.protect_patches(e, pid, 1 + (maxlev - 1) * max(layers, 1))
        rc, rs, _ = w.hex_select_refine(e, maxlev, layers, protected_patch_ids=pid)
        cells = np.flatnonzero(rs[0]).astype(np.int32)
        assert np.array_equal(cells, cells_g), f"step {step}"
        refine_cell, err = rc[0], e[0]
        if len(cells):
            cm = oc_ref.refine(cells)
            refine_cell = O.remap_refine_cell(cm, np.arange(m.n, dtype=np.int32), refine_cell)
            err = O.map_field(cm, err, 1)
            m = oc_ref.build()
            w = O.World([m])
        up, _, _ = w.hex_select_unrefine([err], [refine_cell], layers, protected_patch_ids=pid)
        pts = np.flatnonzero(up[0]).astype(np.int32)
        assert np.array_equal(pts, pts_g), f"step {step}"
        if len(pts):
            oc_ref.unrefine(m, pts)
    assert any(k == "unrefine" for k, _ in am.history) and any(k == "refine" for k, _ in am.history)
    am.close()


@pytest.mark.gpu
def test_polyRefiner_keyword_through_update():
    """`refiner polyRefiner;` (what the reference's tutorials and tests ship) on a 3-D mesh: update() refines the
    box, keeps face AND edge 2:1, and unrefines again; nUnrefinementBufferLayers is raised to 2 (fvMeshPolyRefiner.C:287-300)"""
    text = (DICT % dict(layers=1, maxlev=2, prot="")).replace("refiner         hexRefiner;", "refiner         polyRefiner;")
    oc = M.Octree(20, 20, 5, (0.1, 0.1, 0.01), 2)
    am = hostapi.AdaptiveFvMesh(text, oc, faces=True)
    n0 = am.n_cells
    for _ in range(2):
        am.set_field("test", cases.box_field(am.mesh.cc))
        assert am.update()
    n1 = am.n_cells
    assert n1 > n0 and am.mesh.cell_level.max() == 2
    m = am.mesh
    lv = m.cell_level
    for e in range(m.n_edges):
        cs = m.ec_cells[m.ec_off[e]:m.ec_off[e + 1]]
        assert lv[cs].max() - lv[cs].min() <= 1
    for _ in range(4):
        am.set_field("test", np.zeros(am.n_cells))
        am.update()
    assert am.n_cells < n1
    am.close()


@pytest.mark.gpu
def test_unknown_names_and_missing_keywords_raise():
    oc = M.Octree(4, 4, 4, max_level=1)
    am = hostapi.AdaptiveFvMesh("errorEstimator nonsense; lowerRefineLevel 1; unrefineLevel 1; maxRefinement 1;", oc)
    with pytest.raises(hostapi.FatalError, match="Unknown errorEstimator type nonsense"):
        am.update()
    am.close()
    am = hostapi.AdaptiveFvMesh("errorEstimator fieldValue; fieldName test; unrefineLevel 1; maxRefinement 1;", M.Octree(4, 4, 4, max_level=1))
    am.set_field("test", np.zeros(64))
    with pytest.raises(hostapi.FatalError, match="lowerRefineLevel"):
        am.update()
    am.close()
    am = hostapi.AdaptiveFvMesh("errorEstimator fieldValue; fieldName test; lowerRefineLevel 1; unrefineLevel 1;", M.Octree(4, 4, 4, max_level=1))
    am.set_field("test", np.zeros(64))
    with pytest.raises(hostapi.FatalError, match="Either maxRefinement or minDx"):
        am.update()
    am.close()


CAVITY2D = [("xmin", "movingWall"), ("xmax", "fixedWalls"), ("ymin", "fixedWalls"), ("ymax", "fixedWalls"),
            ("zmin", "frontAndBack"), ("zmax", "frontAndBack")]


@pytest.mark.gpu
def test_polyRefiner_2d_prismatic_through_update():
    """tutorials/damBreak2D as shipped: `refiner polyRefiner;` on a one-cell-thick mesh with empty front/back ->
    prismatic2DRefinement (fvMeshPolyRefiner.C:139-155); nRefinementBufferLayers raised to 3 (:267-283)"""
    text = (DICT % dict(layers=1, maxlev=2, prot="")).replace("refiner         hexRefiner;", "refiner         polyRefiner;")
    oc = M.Octree(24, 24, 1, (0.1, 0.1, 0.01), 2, CAVITY2D, {"frontAndBack": M.PATCH_EMPTY}, two_d=True)
    am = hostapi.AdaptiveFvMesh(text, oc, faces=True)
    n0 = am.n_cells
    for _ in range(2):
        am.set_field("test", cases.box_field(am.mesh.cc))
        assert am.update()
    n1 = am.n_cells
    assert n1 > n0 and am.mesh.cell_level.max() == 2
    m = am.mesh
    lv = m.cell_level
    assert np.abs(lv[m.owner[:m.f]] - lv[m.neighbour]).max() <= 1
    for e in range(m.n_edges):
        cs = m.ec_cells[m.ec_off[e]:m.ec_off[e + 1]]
        assert lv[cs].max() - lv[cs].min() <= 1
    for _ in range(4):
        am.set_field("test", np.zeros(am.n_cells))
        am.update()
    assert am.n_cells < n1
    assert any(k == "unrefine" for k, _ in am.history)
    am.close()


@pytest.mark.gpu
def test_gradientRange_on_device_through_update():
    """gradientRange with fvc::grad (Gauss linear) computed by the library: the steep part of a front is refined"""
    text = """errorEstimator gradientRange; gradField T; refineThreshold 0.3; unrefineThreshold 0.06;
              lowerRefineLevel 0; unrefineLevel 0; maxRefinement 2; nBufferLayers 1;"""
    oc = M.Octree(12, 12, 12, (1.0, 1.0, 1.0), 2)
    am = hostapi.AdaptiveFvMesh(text, oc, geometry=True)
    n0 = am.n_cells
    for _ in range(2):
        am.set_field("T", cases.front_field(am.mesh.cc, r0=0.3, w=0.05))
        assert am.update()
    m = am.mesh
    d = np.sqrt(((m.cc - 0.5) ** 2).sum(1))
    assert am.n_cells > n0
    assert (m.cell_level == 2).any()
    assert np.abs(d[m.cell_level == 2] - 0.3).max() < 0.25       # refinement hugs the front
    assert (m.cell_level[np.abs(d - 0.3) < 0.02] == 2).mean() > 0.9
    am.close()


@pytest.mark.gpu
def test_multicomponent_cellSize_and_minDx_through_update():
    """multicomponent of (fieldValue on a box with minDx instead of maxRefinement, cellSize characteristic):
    cells inside the box refine until they are narrower than minDx; cellSize keeps every other cell as it is"""
    text = """errorEstimator multicomponent;
              errorEstimators
              (
                  box  { errorEstimator fieldValue; fieldName test; lowerRefineLevel 0.01; unrefineLevel 0.01; minDx 0.03; }
                  size { errorEstimator cellSize; type characteristic; minDx 0.001; maxDx 10; }
              );
              lowerRefineLevel 0.01; unrefineLevel 0.01; maxRefinement 3; nBufferLayers 1;"""
    oc = M.Octree(10, 10, 10, (1.0, 1.0, 1.0), 3)
    am = hostapi.AdaptiveFvMesh(text, oc, geometry=True)
    n0 = am.n_cells
    changed = []
    for _ in range(4):
        inside = np.all((am.mesh.cc > 0.3) & (am.mesh.cc < 0.6), axis=1).astype(np.float64)
        am.set_field("test", inside)
        changed.append(am.update())
    m = am.mesh
    assert changed[0] and am.n_cells > n0
    inside = np.all((m.cc > 0.3) & (m.cc < 0.6), axis=1)
    # dx0 = 0.1 -> 0.05 -> 0.025 < minDx: exactly two levels, although maxRefinement of the outer dictionary is 3
    assert m.cell_level[inside].max() == 2 and (m.cell_level[inside] == 2).all()
    am.close()
