"""Processor-patch halo plan (host side).  Mirrors what libamr_b200.so does over NCCL in
haloExchange(): for every processor patch send the boundary-face values of the patch's face range
and receive the neighbour's into the same range (syncTools::swapBoundaryFaceList semantics:
non-coupled boundary faces keep their own value).  The torch.distributed version below runs on any
backend (gloo on CPU for tests, nccl for device tensors)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .mesh import PATCH_PROCESSOR, PolyMesh


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
