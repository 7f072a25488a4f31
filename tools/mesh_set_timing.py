"""Cost of handing a new topology to the library (amr_mesh_set + amr_mesh_set_points + amr_levels_set): what a real
AMR loop pays once per topology change, outside the per-step decision path that bench.py times.
Usage: python tools/mesh_set_timing.py [cells_per_side]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from blastamr_b200 import capi, mesh as M  # noqa: E402


def main():
    n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    t0 = time.time()
    m = M.box_mesh(n_side, n_side, n_side, (1.0, 1.0, 1.0))
    print(f"host: blockMesh-equivalent build of {m.n} cells: {time.time() - t0:.2f} s")
    ctx = capi.Context(0)
    L, h, P = ctx.L, ctx._h, capi._ptr
    st, sz, ty, nb = m.patch_arrays()
    import torch
    for rep in range(3):
        torch.cuda.synchronize()
        t = [time.time()]
        ctx._ck(L.amr_mesh_set(h, m.n, m.f, m.b, m.p, P(m.owner), P(m.neighbour), len(st), P(st), P(sz), P(ty), P(nb)))
        torch.cuda.synchronize(); t.append(time.time())
        ctx._ck(L.amr_mesh_set_points(h, P(m.pc_off), P(m.pc_cells), P(m.bnd_point)))
        torch.cuda.synchronize(); t.append(time.time())
        sp = m.split_parent
        ctx._ck(L.amr_levels_set(h, P(m.cell_level), P(m.point_level), P(m.visible), P(sp) if len(sp) else None, len(sp)))
        torch.cuda.synchronize(); t.append(time.time())
        gb = (m.owner.nbytes + m.neighbour.nbytes + m.pc_off.nbytes + m.pc_cells.nbytes + m.bnd_point.nbytes + m.cell_level.nbytes
              + m.point_level.nbytes + m.visible.nbytes) / 1e9
        print(f"rep {rep}: amr_mesh_set {1e3 * (t[1] - t[0]):.1f} ms, amr_mesh_set_points {1e3 * (t[2] - t[1]):.1f} ms, "
              f"amr_levels_set {1e3 * (t[3] - t[2]):.1f} ms, total {1e3 * (t[3] - t[0]):.1f} ms for {gb:.2f} GB of host arrays (pageable)")
    ctx.n, ctx.f, ctx.b, ctx.p = m.n, m.f, m.b, m.p
    x = np.random.default_rng(0).random(m.n)
    ctx.estimate("delta", x)
    print("ok")
    ctx.close()


if __name__ == "__main__":
    main()
