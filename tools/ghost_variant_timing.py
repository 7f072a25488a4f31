"""Single-GPU timing of the GHOST=true kernel variants: a box whose xmax side is a processor patch, with the
neighbour values supplied by the host side (no communicator needed)."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from blastamr_b200 import capi, mesh as M

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
sides = [("xmin", "walls"), ("ymin", "walls"), ("ymax", "walls"), ("zmin", "walls"), ("zmax", "walls"), ("xmax", "proc")]
for coupled in (False, True):
    ty = {"proc": M.PATCH_PROCESSOR} if coupled else {}
    m = M.box_mesh(N, N, N, sides=sides, patch_types=ty, patch_nbr={"proc": 1} if coupled else None)
    ctx = capi.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_mesh(m)
    d = np.sqrt(((m.cc - 0.5) ** 2).sum(1))
    x = torch.from_numpy(1 + 0.5 * (1 + np.tanh((d - 0.3) / (2.0 / N)))).cuda()
    nbr = torch.ones(m.b, dtype=torch.float64, device="cuda") if coupled else None
    out = torch.empty(m.n, dtype=torch.float64, device="cuda")
    thr = (0.01, 1e15, 0.01, 1e15)
    for _ in range(3):
        ctx.estimate("delta", x, nbr_values=nbr, thresholds=thr, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ctx.estimate("delta", x, nbr_values=nbr, thresholds=thr, out=out)
    b.record(); torch.cuda.synchronize()
    print("coupled" if coupled else "plain  ", "estimate ms", a.elapsed_time(b) / 10)
    ctx.close()
