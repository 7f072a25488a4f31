"""world_size-2 gloo test of the N>1 host logic: decomposition into processor patches, the halo
plan, and the swap protocol (pack at boundary-face owners -> per-patch send/recv -> ghost rule).
Each rank runs an order-free numpy version of the estimator, one buffer layer and the 2:1 closure
with real message passing; the result must equal the serial oracle on the undecomposed mesh."""
import os
import socket

import numpy as np
import pytest

from tests import amr_cases as cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from blastamr_b200 import mesh as M
    from tests.halo_host_mirror import halo_plan, swap_boundary
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oc = cases.graded_octree(8, 2)
        g = oc.build()
        loc = M.decompose(g, M.simple_ranks(g, (1, 1, 1), (world, 1, 1)), world)[rank]
        plan = halo_plan(loc)
        assert plan and all(n != rank for n, _, _ in plan)
        x = cases.front_field(g.cc, w=0.05)[loc.cell_global]
        own, nei = loc.owner[:loc.f], loc.neighbour
        bown = loc.owner[loc.f:]
        coupled = np.zeros(loc.b, bool)
        for _, off, cnt in plan:
            coupled[off:off + cnt] = True

        def swap(v):
            return swap_boundary(torch.from_numpy(np.ascontiguousarray(v)), plan).numpy()

        # delta estimator with patchNeighbourField from the swap (delta.C:93-131)
        e = np.zeros(loc.n)
        d = np.minimum(np.abs(x[own] - x[nei]), 0.5)
        np.maximum.at(e, own, d)
        np.maximum.at(e, nei, d)
        ghost = swap(x[bown])
        np.maximum.at(e, bown[coupled], np.abs(x[bown] - ghost)[coupled])
        mark = e > 0.01
        # one forced buffer layer (fvMeshRefiner.C:794-863), syncFaceList == OR with the ghost source
        src = mark & (loc.cell_level < 3)
        gsrc = swap(src[bown].astype(np.uint8)).astype(bool)
        grown = mark.copy()
        np.logical_or.at(grown, own, src[nei])
        np.logical_or.at(grown, nei, src[own])
        np.logical_or.at(grown, bown[coupled], gsrc[coupled])
        flags = grown & (loc.cell_level < 3)
        # 2:1 closure with swapBoundaryFaceList + reduce(nChanged) per sweep (hexRef.C:1421-1437)
        sweeps = 0
        while True:
            lv = (loc.cell_level + flags).astype(np.int32)
            glv = swap(lv[bown])
            new = flags.copy()
            new[nei[lv[own] > lv[nei] + 1]] = True
            new[own[lv[nei] > lv[own] + 1]] = True
            bb = coupled & (glv > lv[bown] + 1)
            new[bown[bb]] = True
            changed = torch.tensor([int((new != flags).sum())])
            dist.all_reduce(changed)
            flags = new
            sweeps += 1
            if int(changed) == 0:
                break
        q.put((rank, loc.cell_global.copy(), e, grown, flags, sweeps))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_serial_oracle():
    import torch.multiprocessing as mp
    from oracle import oracle as O
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = cases.graded_octree(8, 2).build()
    w = O.World([g])
    x = cases.front_field(g.cc, w=0.05)
    e = w.estimate("delta", [x])
    raw = e[0].copy()
    w.normalize(e, 0.01, 1e15, 0.01, 1e15)
    rc, rs, _ = w.hex_select_refine(e, 3, 1)
    seen = 0
    for rank, cg, e_r, grown_r, flags_r, sweeps in res:
        assert np.array_equal(raw[cg], e_r)
        assert np.array_equal(rc[0][cg].astype(bool), grown_r)
        assert np.array_equal(rs[0][cg].astype(bool), flags_r)
        seen += len(cg)
    assert seen == g.n


def _point_worker(rank, world, procs, port, q):
    import torch
    import torch.distributed as dist
    from blastamr_b200 import mesh as M
    from tests.halo_host_mirror import point_plan, sync_points_or
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oc = cases.poly_balanced_octree(O, n0=8, max_level=1)
        g = oc.build(geometry=True, faces=True)
        loc = M.decompose(g, M.simple_ranks(g, (1, 1, 1), procs), world)[rank]
        plan = point_plan(loc)
        marked = np.zeros(g.n, bool)
        marked[::37] = True                                  # the same global marking on every rank
        mc = marked[loc.cell_global]
        # extendMarkedCellsAcrossPoints (fvMeshRefiner.C:924-977): cells -> points, sync, points -> cells
        pts_of = np.repeat(np.arange(loc.p), np.diff(loc.pc_off))
        pflag = np.zeros(loc.p, np.uint8)
        np.logical_or.at(pflag, pts_of, mc[loc.pc_cells])
        t = torch.from_numpy(pflag)
        rounds = sync_points_or(t, plan)
        pflag = t.numpy().astype(bool)
        out = mc.copy()
        out[loc.pc_cells[pflag[pts_of]]] = True
        q.put((rank, loc.cell_global.copy(), out, rounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_four_rank_point_sync_reaches_edge_points():
    """the point sync protocol of syncPointsOr (pairwise OR over processor-patch point lists until stable) on a
    2 x 2 split, whose centre line is shared by FOUR ranks, two of which share no face: one point layer grown from a
    global marking must equal the serial growth"""
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from blastamr_b200 import mesh as M
    world, procs = 4, (2, 2, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_point_worker, args=(r, world, procs, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = cases.poly_balanced_octree(O, n0=8, max_level=1).build(faces=True)
    marked = np.zeros(g.n, bool)
    marked[::37] = True
    pts_of = np.repeat(np.arange(g.p), np.diff(g.pc_off))
    pf = np.zeros(g.p, bool)
    np.logical_or.at(pf, pts_of, marked[g.pc_cells])
    ref = marked.copy()
    ref[g.pc_cells[pf[pts_of]]] = True
    glob = np.zeros(g.n, bool)
    for rank, cg, out, rounds in res:
        glob[cg] = out
        assert rounds >= 2                      # at least one exchange plus the round that sees no change
    assert np.array_equal(glob, ref) and ref.sum() > marked.sum()
