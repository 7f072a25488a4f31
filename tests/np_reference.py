"""Independent numpy restatement ("second opinion") of the order-free parts of the path.

The C oracle follows the reference's in-place face loops (Gauss-Seidel).  The functions here compute
the same quantities from their mathematical definition -- max over face neighbours, Jacobi fixed-point
iteration of the 2:1 rule, set-based split-point detection -- so agreement between the two is evidence
for the claim that the reference's results do not depend on sweep order.
"""
import numpy as np


def delta(mesh, x, clamp=0.5):
    own, nei = mesh.owner[:mesh.f], mesh.neighbour
    d = np.abs(x[own] - x[nei])
    if clamp is not None:
        d = np.minimum(d, clamp)
    e = np.zeros(mesh.n)
    np.maximum.at(e, own, d)
    np.maximum.at(e, nei, d)
    return e


def scaled_delta(mesh, x, min_val, offset):
    x = np.abs(x)
    if abs(offset) > 1e-15:
        x = x - offset
    own, nei = mesh.owner[:mesh.f], mesh.neighbour
    d = np.abs(x[own] - x[nei]) / np.maximum(np.minimum(x[own], x[nei]), min_val)
    e = np.zeros(mesh.n)
    np.maximum.at(e, own, d)
    np.maximum.at(e, nei, d)
    return e


def normalize(e, lr, ur, lu, uu):
    out = np.zeros_like(e)
    ref = (e > lr) & (e < ur)
    unr = (e < lu) | (e > uu)
    out[ref] = 1.0
    out[unr] = -1.0
    return out


def grow(mesh, marked, src=None):
    """one face layer: cell marked if it or a face neighbour (in src) is marked"""
    own, nei = mesh.owner[:mesh.f], mesh.neighbour
    src = marked if src is None else src
    out = marked.copy()
    np.logical_or.at(out, own, src[nei])
    np.logical_or.at(out, nei, src[own])
    return out


def closure_refine(mesh, flags):
    """least fixed point of the 2:1 rule (maxSet=true), Jacobi iteration"""
    own, nei = mesh.owner[:mesh.f], mesh.neighbour
    flags = flags.astype(bool).copy()
    it = 0
    while True:
        lv = mesh.cell_level + flags
        add_n = lv[own] > lv[nei] + 1
        add_o = lv[nei] > lv[own] + 1
        new = flags.copy()
        new[nei[add_n]] = True
        new[own[add_o]] = True
        it += 1
        if (new == flags).all():
            return flags, it
        flags = new


def closure_unselect(mesh, flags):
    """greatest fixed point (maxSet=false)"""
    own, nei = mesh.owner[:mesh.f], mesh.neighbour
    flags = flags.astype(bool).copy()
    while True:
        lv = mesh.cell_level + flags
        bad_o = lv[own] > lv[nei] + 1
        bad_n = lv[nei] > lv[own] + 1
        new = flags.copy()
        new[own[bad_o]] = False
        new[nei[bad_n]] = False
        if (new == flags).all():
            return flags
        flags = new


def split_points(mesh):
    """hexRef3D::getSplitElems from its definition"""
    p = mesh.p
    cnt = np.diff(mesh.pc_off)
    par = np.where(mesh.visible >= 0, mesh.split_parent[np.maximum(mesh.visible, 0)] if len(mesh.split_parent) else -1, -1)
    par = np.where(par < 0, -1, par)
    out = np.zeros(p, bool)
    for pt in np.flatnonzero((cnt == 8) & (mesh.bnd_point == 0)):
        cs = mesh.pc_cells[mesh.pc_off[pt]:mesh.pc_off[pt + 1]]
        if par[cs[0]] >= 0 and (par[cs] == par[cs[0]]).all() and (mesh.cell_level[cs] == mesh.cell_level[cs[0]]).all():
            out[pt] = True
    return out


def max_cell_field(mesh, v):
    out = np.full(mesh.p, -1e15)
    pts = np.repeat(np.arange(mesh.p), np.diff(mesh.pc_off))
    np.maximum.at(out, pts, v[mesh.pc_cells])
    return out


def closure_unrefine(mesh, pflags):
    """greatest set of points whose merged cells keep 2:1 (Jacobi)"""
    own, nei = mesh.owner[:mesh.f], mesh.neighbour
    pts = np.repeat(np.arange(mesh.p), np.diff(mesh.pc_off))
    pf = pflags.astype(bool).copy()
    while True:
        uc = np.zeros(mesh.n, bool)
        uc[mesh.pc_cells[pf[pts]]] = True
        lv = mesh.cell_level - uc
        bad = np.zeros(mesh.n, bool)
        bad[own[lv[own] < lv[nei] - 1]] = True
        bad[nei[lv[nei] < lv[own] - 1]] = True
        if not bad.any():
            return pf
        uc &= ~bad
        # knock out points with a cell that cannot be unrefined
        ok = np.ones(mesh.p, bool)
        np.logical_and.at(ok, pts, uc[mesh.pc_cells])
        pf &= ok


# ---- rows either side of the path: independent restatements from the definitions -------------------------

def coupled_mask(mesh):
    st, sz, ty, _ = mesh.patch_arrays()
    c = np.zeros(mesh.b, bool)
    e = np.zeros(mesh.b, bool)
    for q in range(len(st)):
        (c if ty[q] == 2 else e if ty[q] == 1 else np.zeros(mesh.b, bool))[st[q] - mesh.f:st[q] - mesh.f + sz[q]] = True
    return c, e


def gauss_grad(mesh, x, xb=None, xn=None):
    """Gauss linear gradient from its definition (sum of Sf*x_f over the cell's faces / V); the summation order
    differs from the reference's, so comparisons are to round-off, not bitwise"""
    f = mesh.f
    own, nei, w = mesh.owner, mesh.neighbour, mesh.weights
    Sf = mesh.Sf.reshape(-1, 3)
    g = np.zeros((mesh.n, 3))
    xf = w[:f] * x[own[:f]] + (1 - w[:f]) * x[nei]
    np.add.at(g, own[:f], Sf[:f] * xf[:, None])
    np.subtract.at(g, nei, Sf[:f] * xf[:, None])
    cpl, _ = coupled_mask(mesh)
    xbf = x[own[f:]].copy() if xb is None else xb.copy()
    if xn is not None:
        xbf[cpl] = (w[f:] * x[own[f:]] + (1 - w[f:]) * xn)[cpl]
    np.add.at(g, own[f:], Sf[f:] * xbf[:, None])
    return g / mesh.vol[:, None]


def flux(mesh, U, Ub=None, Un=None):
    """phiU = linear interpolate(U) & Sf on every face (empty-patch entries meaningless)"""
    f = mesh.f
    own, nei, w = mesh.owner, mesh.neighbour, mesh.weights
    Sf = mesh.Sf.reshape(-1, 3)
    Uf = np.empty((f + mesh.b, 3))
    Uf[:f] = w[:f, None] * (U[own[:f]] - U[nei]) + U[nei]
    Uf[f:] = U[own[f:]] if Ub is None else Ub
    cpl, _ = coupled_mask(mesh)
    if Un is not None and cpl.any():
        Uf[f:][cpl] = (w[f:, None] * U[own[f:]] + (1 - w[f:, None]) * Un)[cpl]
    return Uf[:, 0] * Sf[:, 0] + Uf[:, 1] * Sf[:, 1] + Uf[:, 2] * Sf[:, 2]


def multicomponent_wavefront(mesh, errors, levels):
    """the device algorithm for multicomponent::maxRefinement (fields.cuh k_mc_*), emulated with numpy: checks the
    recurrence against the oracle's literal in-place loop without a GPU"""
    n, f = mesh.n, mesh.f
    own, nei = mesh.owner[:f], mesh.neighbour
    nbr = [[] for _ in range(n)]
    for a, b in zip(own, nei):
        nbr[a].append(int(b))
        nbr[b].append(int(a))
    mr = np.zeros(n, np.int32)
    waves = 0
    for err, lv in zip(errors, levels):
        state = np.where(err > 0.5, 2, 0).astype(np.int8)
        while (state == 2).any():
            new = state.copy()                      # Jacobi sweep: the most conservative schedule
            for c in np.flatnonzero(state == 2):
                lower = [a for a in nbr[c] if a < c]
                max_f = max([a for a in lower if state[a] == 1], default=-1)
                max_u = max([a for a in lower if state[a] == 2], default=-1)
                if max_u > max_f:
                    continue
                seen = lv[max_f] if max_f >= 0 else mr[c]
                new[c] = 1 if seen < lv[c] else 0
            assert (new != state).any()
            state = new
            waves += 1
        out = mr.copy()
        for c in range(n):
            fired = [a for a in nbr[c] + [c] if state[a] == 1]
            if fired:
                out[c] = lv[max(fired)]
        mr = out
    return mr, waves


# ---- polyRefiner (face + edge rules) from the definitions -------------------------------------------------

def _edge_ids(mesh):
    return np.repeat(np.arange(mesh.n_edges), np.diff(mesh.ec_off))


def closure_refine_poly(mesh, flags, edge_based=True):
    """least fixed point of refinement::consistentRefinement (refinement.C:744-810): 2:1 over faces and, with
    edgeBasedConsistency, over every pair of cells sharing a mesh edge.  Jacobi iteration."""
    own, nei = mesh.owner[:mesh.f], mesh.neighbour
    eid = _edge_ids(mesh)
    flags = flags.astype(bool).copy()
    while True:
        lv = mesh.cell_level + flags
        new = flags.copy()
        new[nei[lv[own] > lv[nei] + 1]] = True
        new[own[lv[nei] > lv[own] + 1]] = True
        if edge_based:
            mx = np.zeros(mesh.n_edges, np.int64)
            np.maximum.at(mx, eid, lv[mesh.ec_cells])
            new[mesh.ec_cells[lv[mesh.ec_cells] + 1 < mx[eid]]] = True
        if (new == flags).all():
            return flags
        flags = new


def balanced_poly(mesh, lv):
    """2:1 over faces and over edge-cell pairs"""
    own, nei = mesh.owner[:mesh.f], mesh.neighbour
    if np.abs(lv[own] - lv[nei]).max() > 1:
        return False
    eid = _edge_ids(mesh)
    mx = np.zeros(mesh.n_edges, np.int64)
    mn = np.full(mesh.n_edges, 1 << 30, np.int64)
    np.maximum.at(mx, eid, lv[mesh.ec_cells])
    np.minimum.at(mn, eid, lv[mesh.ec_cells])
    return bool((mx - mn <= 1).all())


def grow_points(mesh, marked):
    """fvMeshRefiner::extendMarkedCellsAcrossPoints: cells -> points -> cells"""
    pts = np.repeat(np.arange(mesh.p), np.diff(mesh.pc_off))
    pm = np.zeros(mesh.p, bool)
    np.logical_or.at(pm, pts, marked[mesh.pc_cells])
    out = marked.copy()
    out[mesh.pc_cells[pm[pts]]] = True
    return out


def prism_split_points(mesh):
    """prismatic2DRefinement.C:2756-2849: pointLevel >= 1 and every edge at the point joins equal point levels"""
    bad = np.zeros(mesh.p, bool)
    e = mesh.edge_pts.reshape(-1, 2)
    d = mesh.point_level[e[:, 0]] != mesh.point_level[e[:, 1]]
    bad[e[d, 0]] = True
    bad[e[d, 1]] = True
    return (mesh.point_level >= 1) & ~bad


def poly_split_points(mesh):
    """polyhedralRefinement.C:2110-2184: pointLevel >= 1, every face at the point has one point level, no boundary face"""
    bad = np.zeros(mesh.p, bool)
    nf = mesh.f + mesh.b
    for i in range(nf):
        ps = mesh.face_pts[mesh.face_off[i]:mesh.face_off[i + 1]]
        if i >= mesh.f or len(set(mesh.point_level[ps].tolist())) > 1:
            bad[ps] = True
    return (mesh.point_level >= 1) & ~bad


def closure_unrefine_poly(mesh, pflags, edge_based=True):
    """greatest set of split points whose merged cells keep 2:1 over faces (and edges): Jacobi form of
    polyhedralRefinement.C:2257-2328"""
    own, nei = mesh.owner[:mesh.f], mesh.neighbour
    pts = np.repeat(np.arange(mesh.p), np.diff(mesh.pc_off))
    eid = _edge_ids(mesh)
    pf = pflags.astype(bool).copy()
    while True:
        uc = np.zeros(mesh.n, bool)
        uc[mesh.pc_cells[pf[pts]]] = True
        lv = mesh.cell_level - uc
        bad = np.zeros(mesh.n, bool)
        bad[own[lv[own] < lv[nei] - 1]] = True
        bad[nei[lv[nei] < lv[own] - 1]] = True
        if edge_based:
            mx = np.zeros(mesh.n_edges, np.int64)
            np.maximum.at(mx, eid, lv[mesh.ec_cells])
            bad[mesh.ec_cells[(lv[mesh.ec_cells] < mx[eid] - 1)]] = True
        bad &= uc
        if not bad.any():
            return pf
        uc &= ~bad
        ok = np.ones(mesh.p, bool)
        np.logical_and.at(ok, pts, uc[mesh.pc_cells])
        pf &= ok
