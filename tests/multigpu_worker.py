"""Launched by torchrun (one rank per GPU): the C-ABI path with NCCL processor-patch halos on a
decomposed graded mesh; every rank compares its slice with the serial oracle on the global mesh.
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
       --master-port P tests/multigpu_worker.py"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

PROCS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def main():
    import torch
    import torch.distributed as dist
    from blastamr_b200 import capi, mesh as M
    from oracle import oracle as O
    from tests import amr_cases as cases

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = capi.Context(lr)
    uid = [capi.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])

    oc = cases.graded_octree(12, 3, centre=(0.5, 0.5, 0.5), r0=0.3, band=0.1)
    g = oc.build(geometry=True)
    loc = M.decompose(g, M.simple_ranks(g, (1, 1, 1), PROCS[world]), world)
    me = loc[rank]
    cg = me.cell_global
    thr = (0.01, 1e15, 0.01, 1e15)
    w = O.World([g])
    wR = O.World(loc)
    ctx.set_mesh(me)
    fails = []
    for shift in (0.0, 0.04):
        x = cases.front_field(g.cc, centre=(0.5 + shift, 0.5, 0.5), r0=0.3, w=0.03)
        for kind in ("delta", "scaledDelta", "codedDelta"):
            ref = w.estimate(kind, [x], min_val=1e-3)[0]
            got = ctx.estimate(kind, x[cg].copy(), p0=1e-3)
            if not np.array_equal(ref[cg], got):
                fails.append(f"estimate {kind}")
        # Lohner against the multi-rank oracle world on the same decomposition
        refL = wR.estimate_lohner([x[l.cell_global].copy() for l in loc], 0.05)[rank]
        gotL = ctx.estimate_lohner(x[cg].copy(), 0.05)
        if not np.array_equal(refL, gotL):
            fails.append("Lohner")
        # gradient selector on the decomposed mesh (8-byte swap of x, gMin/gMax over ranks) and the flux fix-up with
        # every face rewritten (24-byte swap of U): both against the multi-rank oracle world
        xs = [x[l.cell_global].copy() for l in loc]
        xn = wR.swap_owner_values(xs)
        mags = [wR.gradient_mag(r, xs[r], xn=xn[r])[0] for r in range(world)]
        thr_w, _ = wR.range_thresholds(mags, 0.3, 0.06)
        refg = mags[rank].copy(); wR.normalize([refg], *thr_w)
        gotg, gthr2, _ = ctx.estimate_gradient(xs[rank], refine_threshold=0.3, unrefine_threshold=0.06)
        if not (np.array_equal(np.asarray(thr_w), gthr2) and np.array_equal(refg, gotg)):
            fails.append("estimate_gradient")
        Us = [np.stack([np.sin(3 * l.cc[:, 0]), np.cos(2 * l.cc[:, 1]), l.cc[:, 2] ** 2 + shift], axis=1).copy() for l in loc]
        Un = wR.swap_owner_values(Us)
        fmap = np.full(me.f + me.b, -1, np.int32)
        phi_ref = np.zeros(me.f + me.b); phi_gpu = np.zeros(me.f + me.b)
        n_ref = wR.correct_fluxes(rank, fmap, np.zeros(1, np.int32), Us[rank], phi_ref, nbr_U=np.ascontiguousarray(Un[rank]))
        n_gpu = ctx.correct_fluxes(fmap, np.zeros(1, np.int32), Us[rank], phi_gpu)
        if n_ref != n_gpu or not np.array_equal(phi_ref, phi_gpu):
            fails.append("correct_fluxes (vector halo)")
        # gradientRange thresholds (gMin/gMax over ranks)
        raw = w.estimate("delta", [x])[0]
        thr_g, mm_g = w.range_thresholds([raw], 0.5, 0.4)
        refn = raw.copy(); w.normalize([refn], *thr_g)
        gotn, gthr, gmm = ctx.normalize_range(raw[cg].copy(), 0.5, 0.4)
        if not (np.array_equal(thr_g, gthr) and np.array_equal(refn[cg], gotn)):
            fails.append("gradientRange thresholds")
        e = w.estimate("delta", [x]); w.normalize(e, *thr)
        eg = ctx.estimate("delta", x[cg].copy(), thresholds=thr)
        if not np.array_equal(e[0][cg], eg):
            fails.append("normalised error")
        for layers in (1, 3):
            rc, rs, _ = w.hex_select_refine(e, 3, layers)
            rc_g = np.zeros(me.n, np.uint8)
            cells, sweeps = ctx.select_refine(eg, 3, layers, refine_cell_out=rc_g)
            if not np.array_equal(rc[0][cg], rc_g):
                fails.append(f"refineCell layers={layers}")
            if not np.array_equal(np.flatnonzero(rs[0][cg]), cells):
                fails.append(f"cellsToRefine layers={layers}")
            # unrefinement: local split points exclude processor-boundary points -> compare with the
            # multi-rank oracle world (same decomposition), which restates exactly that
            eR = [e[0][l.cell_global].copy() for l in loc]
            rcR = [rc[0][l.cell_global].copy() for l in loc]
            upR, spR, _ = wR.hex_select_unrefine(eR, rcR, layers)
            rc_in = rc[0][cg].copy()
            pts, _ = ctx.select_unrefine(eg, layers, refine_cell=rc_in)
            if not np.array_equal(np.flatnonzero(upR[rank]), pts):
                fails.append(f"unrefine points layers={layers}")
            if not np.array_equal(rcR[rank], rc_in):
                fails.append(f"refineCell after unrefine layers layers={layers}")
        if ctx.check_levels() != 0:
            fails.append("check_levels")
        # protect patches across ranks
        pid = [me.patch_id("movingWall")]
        e2 = e[0].copy(); w.protect_patches([e2], pid, 3)
        e2g = eg.copy(); ctx.protect_patches(pid, 3, e2g)
        if not np.array_equal(e2[cg], e2g):
            fails.append("protect_patches")
    # ---- polyRefiner across ranks: face + (local) edge consistency with halos, and the unrefinement buffer layers over
    # POINTS, whose marks cross processor patches through syncTools::syncPointList (fvMeshRefiner.C:952-958)
    ocp = cases.poly_balanced_octree(O)
    gp = ocp.build(geometry=True, faces=True)
    locp = M.decompose(gp, M.simple_ranks(gp, (1, 1, 1), PROCS[world]), world)
    mep = locp[rank]
    wP = O.World(locp)
    ctx.set_mesh(mep)
    xp = cases.front_field(gp.cc, centre=(0.62, 0.5, 0.5), r0=0.3, w=0.03)
    eP = wP.estimate("delta", [xp[l.cell_global].copy() for l in locp])
    wP.normalize(eP, *thr)
    rcP, rsP, _ = wP.poly_select_refine(eP, 2, 2)
    rc_g = np.zeros(mep.n, np.uint8)
    cells, _ = ctx.select_refine(eP[rank], 2, 2, refine_cell_out=rc_g, kind=2)
    if not (np.array_equal(rcP[rank], rc_g) and np.array_equal(np.flatnonzero(rsP[rank]), cells)):
        fails.append("polyRefiner select_refine")
    rc_ref = [r.copy() for r in rcP]
    upP, _, _ = wP.poly_select_unrefine(eP, rc_ref, 2)
    rc_gpu = rcP[rank].copy()
    pts, _ = ctx.select_unrefine(eP[rank], 2, refine_cell=rc_gpu, kind=2)
    if not np.array_equal(rc_ref[rank], rc_gpu):
        fails.append("polyRefiner point buffer layers (syncPointList)")
    if not np.array_equal(np.flatnonzero(upP[rank]), pts):
        fails.append("polyRefiner unrefine points")
    if sum(int(u.sum()) for u in upP) == 0 or all(np.array_equal(a, b) for a, b in zip(rcP, rc_ref)):
        fails.append("polyRefiner case is trivial")
    # pointLevel must agree across processor patches (hexRef.C:3027-3057): consistent as built, caught when one rank differs
    if ctx.check_point_levels() != 0:
        fails.append("check_point_levels on a consistent mesh")
    pl = mep.point_level.copy()
    if rank == 0 and mep.patch_pts is not None and len(mep.patch_pts):
        pl[mep.patch_pts[0]] += 1
    if ctx.check_point_levels(pl) == 0:
        fails.append("check_point_levels missed an inconsistent point")
    t = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(t)
    if fails:
        print(f"rank {rank} FAILED: {fails}", flush=True)
    if rank == 0:
        print("MULTIGPU_PARITY", "OK" if int(t) == 0 else "FAIL", "world", world, flush=True)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(1 if int(t) else 0)


if __name__ == "__main__":
    main()
