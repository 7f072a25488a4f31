"""Generates the committed golden fixtures.  Run in the build container only
(needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

1. reference mesh fixtures (data, not source): tests/testCases/{hex3D,hex2D}/constant/polyMesh of the
   reference, parsed from OpenFOAM ASCII into npz (owner, neighbour, faces CSR, points, patches);
2. oracle outputs on those meshes for the reference's own test recipe
   (tests/adaptiveFvMeshTests/adaptiveFvMeshTest.C:73-100,156-170: field = 1 inside the box
   (0.02 0.025 -1)(0.04 0.035 1), fieldValue estimator, lowerRefineLevel = unrefineLevel = 0.01,
   nBufferLayers in {1,2}, maxRefinement in {1,3}) and for delta / scaledDelta on an analytic field.
The reference has no golden vectors of its own; these pin the oracle against regressions and feed
the GPU parity tests with the reference's meshes.
"""
import re
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/tests/testCases")


def _body(path):
    t = Path(path).read_text()
    t = re.sub(r"/\*.*?\*/", "", t, flags=re.S)
    t = re.sub(r"//.*", "", t)
    t = t[t.index("}") + 1:]          # after FoamFile header
    m = re.search(r"(\d+)\s*\(", t)
    return int(m.group(1)), t[m.end():]


def read_labels(path):
    n, b = _body(path)
    a = np.array(b[:b.rindex(")")].split(), dtype=np.int32)
    assert len(a) == n
    return a


def read_faces(path):
    n, b = _body(path)
    off = [0]
    pts = []
    for m in re.finditer(r"(\d+)\(([^)]*)\)", b):
        k = int(m.group(1))
        v = [int(x) for x in m.group(2).split()]
        assert len(v) == k
        pts += v
        off.append(len(pts))
    assert len(off) == n + 1
    return np.array(off, np.int32), np.array(pts, np.int32)


def read_points(path):
    n, b = _body(path)
    v = re.findall(r"\(([^()]*)\)", b)
    a = np.array([[float(x) for x in s.split()] for s in v], dtype=np.float64)
    assert a.shape == (n, 3)
    return a


def read_boundary(path):
    t = Path(path).read_text()
    t = re.sub(r"/\*.*?\*/", "", t, flags=re.S)
    t = t[t.index("}") + 1:]
    out = []
    for m in re.finditer(r"(\w+)\s*\{([^}]*)\}", t):
        d = dict(re.findall(r"(\w+)\s+([^;]+);", m.group(2)))
        out.append((m.group(1), d["type"].strip(), int(d["startFace"]), int(d["nFaces"])))
    return out


def convert(case):
    pm = REF / case / "constant" / "polyMesh"
    owner = read_labels(pm / "owner")
    nei = read_labels(pm / "neighbour")
    foff, fpts = read_faces(pm / "faces")
    pts = read_points(pm / "points")
    bnd = read_boundary(pm / "boundary")
    np.savez_compressed(HERE / f"{case}_mesh.npz", owner=owner, neighbour=nei, face_off=foff, face_pts=fpts, points=pts,
                        patch_names=np.array([b[0] for b in bnd]), patch_types=np.array([b[1] for b in bnd]),
                        patch_start=np.array([b[2] for b in bnd], np.int32), patch_size=np.array([b[3] for b in bnd], np.int32))
    print(case, "cells", owner.max() + 1, "faces", len(owner), "internal", len(nei), "points", len(pts), bnd)


def convert_faces_only(case):
    """poly3D ships without faces/points in the reference checkout (.MISSING_LARGE_BLOBS): owner / neighbour /
    boundary are enough for every face-list kernel"""
    pm = REF / case / "constant" / "polyMesh"
    owner = read_labels(pm / "owner")
    nei = read_labels(pm / "neighbour")
    bnd = read_boundary(pm / "boundary")
    f = min(b[2] for b in bnd)                      # foam-extend writes neighbour with nFaces entries, -1 on boundary faces
    assert (nei[f:] == -1).all() and (nei[:f] >= 0).all() and (owner[:f] < nei[:f]).all()
    nei = nei[:f]
    np.savez_compressed(HERE / f"{case}_faces.npz", owner=owner, neighbour=nei,
                        patch_names=np.array([b[0] for b in bnd]), patch_types=np.array([b[1] for b in bnd]),
                        patch_start=np.array([b[2] for b in bnd], np.int32), patch_size=np.array([b[3] for b in bnd], np.int32))
    print(case, "cells", owner.max() + 1, "faces", len(owner), "internal", len(nei), bnd)


# ---- foam-extend binary polyMesh (tests/testCases/poly2D: *.gz, `format binary`) ------------------------------

def _bin_body(path):
    import gzip
    d = gzip.open(path).read()
    i = d.index(b"// * * *")
    return d[d.index(b"\n", i) + 1:]


def _read_count(b, pos):
    """skip white space, read an ASCII integer, skip to just after the following '('"""
    while b[pos:pos + 1].isspace():
        pos += 1
    j = pos
    while b[j:j + 1].isdigit():
        j += 1
    n = int(b[pos:j])
    k = b.index(b"(", j)
    assert b[j:k].strip() == b""
    return n, k + 1


def read_labels_bin(path):
    b = _bin_body(path)
    n, pos = _read_count(b, 0)
    a = np.frombuffer(b, dtype="<i4", count=n, offset=pos).copy()
    assert b[pos + 4 * n:pos + 4 * n + 1] == b")"
    return a


def read_faces_bin(path):
    """faceList written face by face: `k ( <k x int32> )`"""
    b = _bin_body(path)
    n, pos = _read_count(b, 0)
    off, pts = [0], []
    for _ in range(n):
        k, pos = _read_count(b, pos)
        pts.append(np.frombuffer(b, dtype="<i4", count=k, offset=pos))
        pos += 4 * k
        assert b[pos:pos + 1] == b")"
        pos += 1
        off.append(off[-1] + k)
    return np.array(off, np.int32), np.concatenate(pts).astype(np.int32)


def read_points_bin(path):
    b = _bin_body(path)
    n, pos = _read_count(b, 0)
    a = np.frombuffer(b, dtype="<f8", count=3 * n, offset=pos).reshape(n, 3).copy()
    assert b[pos + 24 * n:pos + 24 * n + 1] == b")"
    return a


def convert_poly2d():
    """the reference's polyhedral 2-D fixture (670 prismatic polyhedra, one cell thick): the mesh class of config C3
    (poly_freelyPropFlame2D_H2) and of the reference's own unit test (adaptiveFvMeshTest.C runs on every case)"""
    case = "poly2D"
    pm = REF / case / "constant" / "polyMesh"
    owner = read_labels_bin(pm / "owner.gz")
    nei = read_labels_bin(pm / "neighbour.gz")
    foff, fpts = read_faces_bin(pm / "faces.gz")
    pts = read_points_bin(pm / "points.gz")
    bnd = read_boundary(pm / "boundary")
    f = min(b[2] for b in bnd)
    assert len(owner) == len(foff) - 1 and len(nei) == f and (owner[:f] < nei).all() and fpts.max() == len(pts) - 1
    np.savez_compressed(HERE / f"{case}_mesh.npz", owner=owner, neighbour=nei, face_off=foff, face_pts=fpts, points=pts,
                        patch_names=np.array([b[0] for b in bnd]), patch_types=np.array([b[1] for b in bnd]),
                        patch_start=np.array([b[2] for b in bnd], np.int32), patch_size=np.array([b[3] for b in bnd], np.int32))
    print(case, "cells", owner.max() + 1, "faces", len(owner), "internal", len(nei), "points", len(pts), bnd)


def golden_outputs():
    from tests.fixtures import load_reference_mesh
    from tests import amr_cases as cases
    from oracle import oracle as O
    out = {}
    for case in ("hex3D", "hex2D"):
        m = load_reference_mesh(case)
        w = O.World([m])
        x = cases.box_field(m.cc)
        e = w.estimate("fieldValue", [x])
        w.normalize(e, 0.01, 1e15, 0.01, 1e15)
        out[f"{case}_fieldValue_norm"] = e[0].astype(np.int8)
        y = cases.front_field(m.cc, centre=(0.05, 0.05, 0.005), r0=0.03, w=0.01)
        for kind, kw in (("delta", {}), ("scaledDelta", dict(min_val=1e-3, offset=0.25)), ("codedDelta", {})):
            out[f"{case}_{kind}_raw"] = w.estimate(kind, [y], **kw)[0]
        if case == "hex3D":
            for layers in (1, 2):
                for maxlev in (1, 3):
                    rc, rs, _ = w.hex_select_refine(e, maxlev, layers)
                    out[f"{case}_refineCell_L{layers}_M{maxlev}"] = np.flatnonzero(rc[0]).astype(np.int32)
                    out[f"{case}_cellsToRefine_L{layers}_M{maxlev}"] = np.flatnonzero(rs[0]).astype(np.int32)
    np.savez_compressed(HERE / "oracle_golden.npz", **out)
    print("golden outputs:", sorted(out))


if __name__ == "__main__":
    if REF.exists():
        for c in ("hex3D", "hex2D"):
            convert(c)
        convert_faces_only("poly3D")
        convert_poly2d()
    golden_outputs()
