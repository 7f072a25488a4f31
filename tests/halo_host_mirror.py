"""TEST SUPPORT (not part of the product): processor-patch halo plan on the host.  Mirrors what libamr_b200.so does over NCCL in
haloExchange(): for every processor patch send the boundary-face values of the patch's face range
and receive the neighbour's into the same range (syncTools::swapBoundaryFaceList semantics:
non-coupled boundary faces keep their own value).  The torch.distributed version below runs on any
backend (gloo on CPU for tests, nccl for device tensors)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from blastamr_b200.mesh import PATCH_PROCESSOR, PolyMesh


def halo_plan(mesh: PolyMesh) -> List[Tuple[int, int, int]]:
    """[(neighbour rank, offset into the boundary-face array, count)] in patch order"""
    return [(q.nbr_rank, q.start - mesh.f, q.size) for q in mesh.patches if q.type == PATCH_PROCESSOR and q.size > 0]


def swap_boundary(values, plan, group=None):
    """values: torch tensor over boundary faces [b, ...]; returns the swapped copy."""
    import torch
    import torch.distributed as dist
    out = values.clone()
    ops = []
    recv = []
    for nbr, off, cnt in plan:
        s = values[off:off + cnt].contiguous()
        r = torch.empty_like(s)
        recv.append((off, cnt, r))
        ops.append(dist.P2POp(dist.isend, s, nbr, group))
        ops.append(dist.P2POp(dist.irecv, r, nbr, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for off, cnt, r in recv:
        out[off:off + cnt] = r
    return out


def point_plan(mesh: PolyMesh) -> List[Tuple[int, np.ndarray]]:
    """[(neighbour rank, local point labels of the processor patch in the order shared with the neighbour)]"""
    if mesh.patch_pt_off is None:
        return []
    out = []
    for k, q in enumerate(mesh.patches):
        pts = mesh.patch_pts[mesh.patch_pt_off[k]:mesh.patch_pt_off[k + 1]]
        if q.type == PATCH_PROCESSOR and len(pts):
            out.append((q.nbr_rank, np.ascontiguousarray(pts)))
    return out


def sync_points_or(pflag, plan, group=None) -> int:
    """syncTools::syncPointList(pflag, orEqOp<bool>()) the way libamr_b200.so does it (syncPointsOr): pairwise OR over
    the processor patches' point lists, repeated until no rank changed, so that a point on an edge or corner of the
    decomposition reaches the ranks it shares no face with through a common neighbour.  pflag: uint8 torch tensor over
    the local points, updated in place.  Returns the number of rounds.  Collective."""
    import torch
    import torch.distributed as dist
    rounds = 0
    while True:
        ops, recv = [], []
        for nbr, pts in plan:
            idx = torch.from_numpy(pts.astype(np.int64))
            s = pflag[idx].contiguous()
            r = torch.empty_like(s)
            recv.append((idx, r))
            ops.append(dist.P2POp(dist.isend, s, nbr, group))
            ops.append(dist.P2POp(dist.irecv, r, nbr, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        changed = 0
        for idx, r in recv:
            new = pflag[idx] | r
            changed += int((new != pflag[idx]).sum())
            pflag[idx] = new
        t = torch.tensor([changed])
        dist.all_reduce(t, group=group)
        rounds += 1
        if int(t) == 0:
            return rounds
