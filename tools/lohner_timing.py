"""Lohner estimator (Lohner.C:71-159) and the Gauss-linear gradient selector (gradientRange.C:100) on a graded mesh: kernel time with
device-resident inputs against the algorithmic bytes of SURVEY 8(d) row E3
  8f (owner, neighbour) + 8f (weights) + 24(f+b) (Sf) + 8f (magSf) + 8f (deltaCoeffs) + 8n (V) + 8n (x) + 8n (err)   [+ 24n + 24n: grad written / read]
Run once per form:  AMR_B200_FACE_BATCH=0|1 python tools/lohner_timing.py [level-0 cells per side] [levels]
Prints one JSON line; the checksum must not depend on the form (bitwise-identical arithmetic)."""
import json
import os
import sys
import zlib
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from blastamr_b200 import capi  # noqa: E402
from tests import amr_cases as cases  # noqa: E402


def main():
    import torch
    n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 56
    lev = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    oc = cases.graded_octree(n0, lev, band=0.12)
    m = oc.build(points=False, geometry=True)
    ctx = capi.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_mesh(m)
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(cases.front_field(m.cc, w=0.05)).to(dev)
    xb = torch.from_numpy(np.ascontiguousarray(cases.front_field(m.cc, w=0.05)[m.owner[m.f:]])).to(dev)
    err = torch.empty(m.n, dtype=torch.float64, device=dev)
    thr = np.asarray([0.1, 1e15, 0.05, 1e15])
    thr4 = np.zeros(4)

    def lohner():
        ctx._ck(ctx.L.amr_estimate_lohner(ctx._h, x.data_ptr(), xb.data_ptr(), None, 0.05, 1, capi._ptr(thr), err.data_ptr()))

    def gauss():
        ctx._ck(ctx.L.amr_estimate_gradient(ctx._h, 0, x.data_ptr(), xb.data_ptr(), None, 0.0, 0.3, 0.06, 1, capi._ptr(thr4), None, err.data_ptr()))

    out = {"cells": int(m.n), "faces": int(m.f), "boundary_faces": int(m.b), "form": "batched" if os.environ.get("AMR_B200_FACE_BATCH", "1") != "0" else "one face at a time"}
    n, f, b = m.n, m.f, m.b
    byt = {"lohner": 8 * f + 8 * f + 24 * (f + b) + 8 * f + 8 * f + 8 * n + 8 * n + 8 * n + 48 * n + 4 * (2 * f + b),
           "gauss_gradient_selector": 8 * f + 8 * f + 24 * (f + b) + 8 * n + 8 * n + 8 * n + 48 * n + 4 * (2 * f + b)}
    for name, fn in (("lohner", lohner), ("gauss_gradient_selector", gauss)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        a.record()
        for _ in range(reps):
            fn()
        z.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(z) / reps
        out[name] = {"ms": ms, "algorithmic_bytes": int(byt[name]), "GBps": byt[name] / 1e9 / (ms / 1e3),
                     "checksum": int(zlib.crc32(err.cpu().numpy().tobytes()))}
    out["bytes_note"] = "SURVEY 8(d) E3 + the gradient written and read (48n) + the cell->face rows (4 B per entry)"
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
