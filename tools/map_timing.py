"""Times the forms of the device-side field map (stage F of bench.py) on a refinement-like cellMap:
AMR_B200_MAP=fused (one launch, 56 registers) | fused40 (one launch, 40 registers) | separate (one launch per field).
Usage on the GPU box: python tools/map_timing.py [cells]"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def run(n):
    import numpy as np
    import torch
    from blastamr_b200 import capi
    from blastamr_b200 import mesh as M
    ctx = capi.Context(0)
    ctx.set_mesh(M.box_mesh(4, 4, 4))
    rng = np.random.default_rng(0)
    parents = np.sort(rng.choice(n, n // 45, replace=False)).astype(np.int32)
    cm = torch.from_numpy(np.concatenate([np.arange(n, dtype=np.int32), np.repeat(parents, 7)])).cuda()
    n1 = cm.numel()
    rev = torch.arange(n, dtype=torch.int32, device="cuda")
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    u = torch.rand((n, 3), dtype=torch.float64, device="cuda")
    e = torch.rand(n, dtype=torch.float64, device="cuda")
    rc = (torch.rand(n, device="cuda") < 0.6).to(torch.uint8)
    o = [torch.empty(n1, dtype=torch.float64, device="cuda"), torch.empty((n1, 3), dtype=torch.float64, device="cuda"),
         torch.empty(n1, dtype=torch.float64, device="cuda")]
    byts = 4 * n1 + 5 * (8 * n + 8 * n1) + 4 * n + n + n1
    for _ in range(3):
        ctx.map_fields_and_marking(cm, [x, u, e], n, n1, [1, 3, 1], o, reverse_map=rev, refine_cell=rc)
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ctx.map_fields_and_marking(cm, [x, u, e], n, n1, [1, 3, 1], o, reverse_map=rev, refine_cell=rc)
    z.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(z) / 10
    print(f"{os.environ.get('AMR_B200_MAP', 'fused'):9s} n={n} n'={n1}: {ms:.4f} ms, {byts / 1e9 / (ms / 1e3):.0f} GB/s of algorithmic bytes (cellMap once)")
    assert torch.equal(o[1], u[cm.long()])


if __name__ == "__main__":
    if os.environ.get("_MAP_TIMING_CHILD"):
        run(int(sys.argv[1]))
    else:
        n = sys.argv[1] if len(sys.argv) > 1 else "62584096"
        for form in ("fused", "fused40", "separate"):
            env = dict(os.environ, AMR_B200_MAP=form, _MAP_TIMING_CHILD="1")
            subprocess.run([sys.executable, __file__, n], env=env, check=False)
