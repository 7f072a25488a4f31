#!/usr/bin/env python
"""bench.py -- AMR decision step throughput (estimate + mark + 2:1 balance + map) on B200.

Workload (BASELINE.json configs[4], SURVEY.md 8(d) "C5/S1"): uniform level-0 blockMesh hex box,
N^3 cells (default N=400: 64 M cells, 191.5 M internal faces, 64.5 M points) per GPU, analytic
density front rho(x) = 1 + 0.5 (1 + tanh((|x-c|-r0)/w)), r0 = 0.3, w = 2/N; `delta` estimator on rho,
lowerRefineLevel = unrefineLevel = 0.01, maxRefinement 3, nBufferLayers 1 (refine) / 1 (unrefine),
hexRefiner.  One "step" = the per-timestep decision path of adaptiveFvMesh::refine:
  E   errorEstimator::update (delta + normalize)
  M   selectRefineCandidates + buffer layer + protections + selectRefineCells
  B   hexRef::consistentRefinement sweeps + ascending list
  F   fvMesh::mapFields through the real cellMap of the cut (hexRef3D numbering: child 0 keeps the label,
      7 children appended per refined parent): 1 scalar + 1 vector field, plus the error field and the
      refineCell marking (fvMeshHexRefiner.C:1562-1586), which the unrefinement branch needs on the new mesh
  U   unrefinement branch ON THE REFINED MESH M1 (what the reference does in the same update()):
      buffer layer, maxCellField, getSplitElems, filter, consistentUnrefinement
The topology change itself (M0 -> M1, done once by the host octree engine from the selected cells) is
host bookkeeping and outside the timed region (BASELINE.json north_star); both meshes stay resident.

`value`   = cells before the step (all ranks) / max-over-ranks CUDA-event time, inputs resident in HBM.
`e2e`     = same metric through the C-ABI with HOST (pinned) buffers: the field and the mapped fields
            cross PCIe inside the timed region, every step.
Multi-GPU = one rank per GPU; each rank owns an N^3 sub-box of a `simple (px py pz)` split with
            processor patches; halos of field values / marks / levels over NCCL ("weak" scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

THR = 0.01
GREAT = 1e15
MAX_LEVEL = 3
LAYERS_REF = 1
LAYERS_UNREF = 1
PROCS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def front_field(cc, n_side):
    d = np.sqrt(((cc - 0.5) ** 2).sum(axis=1))
    return 1.0 + 0.5 * (1.0 + np.tanh((d - 0.3) / (2.0 / n_side)))


def algorithmic_bytes(n, f, p, nnz_pc, n_new, sweeps_ref, sweeps_unref, m1=None):
    """each input array read once and each output written once per logical pass of
    the reference algorithm; label 4 B, scalar 8 B, flag 1 B."""
    E = 8 * f + 16 * n
    M1 = 9 * n
    M2 = LAYERS_REF * (8 * f + 2 * n + 8 * n)
    M4 = 10 * n
    B1 = sweeps_ref * (8 * f + 6 * n)
    F1 = 4 * (4 * n_new + 8 * n + 8 * n_new)          # 1 scalar + 1 vector field
    F1 += (4 * n_new + 8 * n + 8 * n_new)              # the error field (a registered volScalarField)
    F1 += (4 * n_new + 4 * n + n + n_new)              # refineCell through cellMap/reverseCellMap
    if m1 is not None:                                 # unrefinement branch runs on the refined mesh
        n, f, p, nnz_pc = m1
    csr = 4 * (p + 1) + 4 * nnz_pc
    U0 = LAYERS_UNREF * (8 * f + 2 * n)
    U1 = csr + 8 * n + 8 * p
    U2 = 8 * n + csr
    U3 = 9 * p + n
    B3 = sweeps_unref * ((8 * f + 6 * n) + 2 * csr + 2 * p)
    return {"E": E, "M": M1 + M2 + M4, "B": B1, "F": F1, "U": U0 + U1 + U2 + U3 + B3}


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason sampling during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def physical_device_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def measured_peak():
    pth = ROOT / "MEASURED_PEAKS.json"
    if pth.exists():
        try:
            return float(json.loads(pth.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------- CPU arms

def oracle_step(O, S):
    """The same step on the CPU oracle (all ranks of the world): E, M+B on M0; F through the cellMap;
    U on the refined mesh M1."""
    w, w1 = S["w"], S["w1"]
    e = w.estimate("delta", S["xs"])
    w.normalize_world(e, THR, GREAT, THR, GREAT)
    rc, rs, sw = w.hex_select_refine(e, MAX_LEVEL, LAYERS_REF)
    cms = S["cms"]
    w.map_field_world(cms, S["xs"], S["new1"], 1)
    w.map_field_world(cms, S["f3"], S["new3"], 3)
    w.map_field_world(cms, e, S["err1"], 1)
    rc1 = [O.remap_refine_cell(cm, rv, r) for cm, rv, r in zip(cms, S["revs"], rc)]
    up, sp, sw2 = w1.hex_select_unrefine(S["err1"], rc1, LAYERS_UNREF)
    return rs, up, sw, sw2


def cpu_sample(n_side, ranks, threads):
    """Oracle on a bounded sample of the workload: an n_side^3 box split in `ranks` slabs (and its
    refined successor M1 split the same way)."""
    from blastamr_b200 import mesh as M
    from oracle import oracle as O
    oc = M.Octree(n_side, n_side, n_side, (1.0, 1.0, 1.0), 1)
    g = oc.build()
    x = front_field(g.cc, n_side)

    def split(mesh):
        if ranks > 1:
            cr = M.simple_ranks(mesh, (1, 1, 1), (1, 1, ranks))
            return M.decompose(mesh, cr, ranks)
        return [mesh]

    loc = split(g)
    xs = [np.ascontiguousarray(x[l.cell_global]) if ranks > 1 else x for l in loc]
    w = O.World(loc)
    O.set_threads(threads)
    e = w.estimate("delta", xs)
    w.normalize_world(e, THR, GREAT, THR, GREAT)
    rc, rs, _ = w.hex_select_refine(e, MAX_LEVEL, LAYERS_REF)
    # global list of cells to refine -> the host topology engine -> M1 and the global cellMap
    if ranks > 1:
        cells = np.sort(np.concatenate([l.cell_global[np.flatnonzero(s_)] for l, s_ in zip(loc, rs)])).astype(np.int32)
    else:
        cells = np.flatnonzero(rs[0]).astype(np.int32)
    cm_g = oc.refine(cells)
    g1 = oc.build()
    loc1 = split(g1)
    w1 = O.World(loc1)
    cms, revs = [], []
    for l0, l1 in zip(loc, loc1):
        if ranks > 1:
            old_local = np.full(g.n, -1, np.int32)
            old_local[l0.cell_global] = np.arange(l0.n, dtype=np.int32)
            cm = old_local[cm_g[l1.cell_global]]
            assert (cm >= 0).all()
            new_local = np.full(g1.n, -1, np.int32)
            new_local[l1.cell_global] = np.arange(l1.n, dtype=np.int32)
            rv = new_local[l0.cell_global]          # old cell keeps its global label in M1
        else:
            cm, rv = cm_g, np.arange(g.n, dtype=np.int32)
        cms.append(np.ascontiguousarray(cm, dtype=np.int32))
        revs.append(np.ascontiguousarray(rv, dtype=np.int32))
    S = dict(w=w, w1=w1, xs=xs, cms=cms, revs=revs,
             f3=[np.stack([a, 2 * a, -a], axis=1).copy() for a in xs],
             new1=[np.empty(len(c)) for c in cms], new3=[np.empty((len(c), 3)) for c in cms],
             err1=[np.empty(len(c)) for c in cms], keep=(g, g1, loc, loc1, oc))
    return O, S, g.n


def run_reference(args):
    """--impl reference: the reference algorithm (restated: oracle/amr_oracle.c; the real reference needs
    OpenFOAM + MPI, absent here) on the host cores, one emulated MPI rank per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ranks = max(1, min(cores, 64))
    n_side = args.cpu_cells
    O, S, ncells = cpu_sample(n_side, ranks, ranks)
    for _ in range(args.warmup):
        oracle_step(O, S)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(O, S)
    dt = (time.perf_counter() - t0) / args.steps
    val = ncells / dt / 1e6
    out = {
        "impl": "reference", "metric": "AMR step Mcells/s (estimate+mark+2:1 balance+map)", "value": val,
        "unit": "Mcells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"synthetic blockMesh hex box, analytic density front, delta estimator, maxRefinement {MAX_LEVEL}; "
                               f"bounded CPU sample {n_side}^3 = {ncells} cells in {ranks} slabs"},
        "cpu_baseline": {"value": val, "unit": "Mcells/s", "cores": ranks, "kind": "port",
                         "sample": f"{n_side}^3-cell box, {ranks} emulated MPI ranks (one thread each), {args.steps} steps"},
        "e2e": {"value": val, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------- GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=400, help="cells per side of each rank's box (400 -> 64 M cells)")
    ap.add_argument("--cpu-cells", type=int, default=160, help="cells per side of the CPU baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from blastamr_b200 import capi, mesh as M

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = args.cells
    procs = PROCS.get(world, (world, 1, 1))

    # ---- mesh of this rank (host bookkeeping, untimed)
    t_setup = time.time()
    if world > 1:
        sides, ptypes, pnbr, _ = M.subdomain_sides(rank, procs)
        octree = M.Octree(N, N, N, (1.0, 1.0, 1.0), 1, sides, ptypes, pnbr)
    else:
        octree = M.Octree(N, N, N, (1.0, 1.0, 1.0), 1)
    mesh = octree.build()
    n, f, b, p = mesh.n, mesh.f, mesh.b, mesh.p
    nnz_pc = int(mesh.pc_off[p])
    x_host = front_field(mesh.cc, N)

    ctx = capi.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if world > 1:
        uid = [capi.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
    ctx.set_mesh(mesh)

    x = torch.from_numpy(x_host).to(dev)
    fld3 = torch.stack([x, 2 * x, -x], dim=1).contiguous()
    cells_dev = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    thr = (THR, GREAT, THR, GREAT)

    # first pass (untimed): select, then let the host topology engine cut M0 -> M1 and hand back the
    # real cellMap; M1 goes to a second context so that both meshes are resident
    ctx.estimate("delta", x, thresholds=thr, keep_resident=True)
    cells_h = np.empty(max(n, 1), dtype=np.int32)
    cells_first, _ = ctx.select_refine(None, MAX_LEVEL, LAYERS_REF, cells_out=cells_h)
    n_ref = len(cells_first)
    if world > 1:
        # each rank cuts its own sub-box: the processor faces of neighbouring ranks stay matched only if no
        # refined cell touches a processor patch (true for the per-rank front, which stays inside the box)
        touched = np.zeros(n, bool)
        for q in mesh.patches:
            if q.coupled:
                touched[mesh.owner[q.start:q.start + q.size]] = True
        if touched[cells_first].any():
            raise SystemExit("bench: refinement reaches a processor patch; the per-rank topology engine cannot keep patches matched")
    del mesh                         # M0's host arrays (~10 GB at 64 M cells) are resident on the device now; M1 is built next
    cell_map_h = octree.refine(cells_first.copy())
    mesh1 = octree.build()
    n1, f1, b1, p1 = mesh1.n, mesh1.f, mesh1.b, mesh1.p
    nnz_pc1 = int(mesh1.pc_off[p1])
    ctx1 = capi.Context(local_rank)
    ctx1.set_stream(torch.cuda.current_stream().cuda_stream)
    if world > 1:
        uid1 = [capi.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid1, src=0)
        ctx1.comm_init(world, rank, uid1[0])
    ctx1.set_mesh(mesh1)
    cell_map = torch.from_numpy(cell_map_h).to(dev)
    rev_map = torch.arange(n, dtype=torch.int32, device=dev)      # every old cell survives under its own label
    n_new = int(cell_map.numel())
    assert n_new == n1
    new1 = torch.empty(n_new, dtype=torch.float64, device=dev)
    new3 = torch.empty((n_new, 3), dtype=torch.float64, device=dev)
    err0 = torch.empty(n, dtype=torch.float64, device=dev)
    err1 = torch.empty(n_new, dtype=torch.float64, device=dev)
    rc0 = torch.empty(n, dtype=torch.uint8, device=dev)
    pts_dev = torch.empty(max(p1, 1), dtype=torch.int32, device=dev)
    del mesh1
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    # one event interval per C-ABI call; F is two calls (scalar field, vector field)
    call_names = ["E:estimate(delta+normalize)", "M+B:select_refine", "F:map_field(scalar)", "F:map_field(vector)",
                  "F:map_field(error)+remap_refine_cell", "U:select_unrefine(on M1)"]
    stage_of = [0, 1, 2, 2, 2, 3]
    stage_names = ["E", "M+B", "F", "U"]
    NEV = len(call_names) + 1
    info = {}

    def step(evs=None):
        def mark(i):
            if evs is not None:
                evs[i].record()
        mark(0)
        ctx.estimate("delta", x, thresholds=thr, out=err0)
        mark(1)
        nr, sw = ctx.select_refine(err0, MAX_LEVEL, LAYERS_REF, cells_out=cells_dev, refine_cell_out=rc0)
        mark(2)
        ctx.map_field(cell_map, x, n_old=n, n_new=n_new, ncomp=1, out=new1)
        mark(3)
        ctx.map_field(cell_map, fld3, n_old=n, n_new=n_new, ncomp=3, out=new3)
        mark(4)
        ctx.map_field(cell_map, err0, n_old=n, n_new=n_new, ncomp=1, out=err1)
        ctx1.remap_refine_cell(cell_map, rev_map, refine_cell=rc0)
        mark(5)
        npts, sw2 = ctx1.select_unrefine(err1, LAYERS_UNREF, elems_out=pts_dev)
        mark(6)
        info.update(n_refine=nr, sweeps_refine=sw, n_unrefine=npts, sweeps_unrefine=sw2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(physical_device_index(local_rank))
    sampler.start()
    l0 = ctx.launches + ctx1.launches
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(NEV)] for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for k in range(args.steps):
        step(evs[k])
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launches + ctx1.launches - l0
    total_ms = t_start.elapsed_time(t_end)
    call_ms = [float(np.mean([evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(args.steps)])) for i in range(NEV - 1)]
    ms_per_step = total_ms / args.steps

    # ---- e2e through the C-ABI with host (pinned) buffers
    e2e = None
    if not args.no_e2e:
        xh = torch.from_numpy(x_host).pin_memory()
        f3h = torch.stack([xh, 2 * xh, -xh], dim=1).contiguous().pin_memory()
        cmh = cell_map.cpu().pin_memory()
        n1h = torch.empty(n_new, dtype=torch.float64).pin_memory()
        n3h = torch.empty((n_new, 3), dtype=torch.float64).pin_memory()
        cells_h2 = np.empty(max(n, 1), dtype=np.int32)
        pts_h = np.empty(max(p1, 1), dtype=np.int32)

        def step_host():
            # Host buffers for everything the solver owns (the field in, the registered fields in and their
            # mapped versions out, the decision lists out); the path's own intermediates (error, refineCell)
            # stay on the device between the calls, as INTEGRATION.md prescribes for the drop-in.
            ctx.estimate("delta", xh.numpy(), thresholds=thr, out=err0)
            cl, _ = ctx.select_refine(err0, MAX_LEVEL, LAYERS_REF, cells_out=cells_h2, refine_cell_out=rc0)
            ctx.map_fields(cmh.numpy(), [xh.numpy(), f3h.numpy()], n, n_new, [1, 3], [n1h.numpy(), n3h.numpy()])   # fvMesh::mapFields
            ctx.map_field(cell_map, err0, n_old=n, n_new=n_new, ncomp=1, out=err1)
            ctx1.remap_refine_cell(cell_map, rev_map, refine_cell=rc0)
            pl, _ = ctx1.select_unrefine(err1, LAYERS_UNREF, elems_out=pts_h)
            return len(cl), len(pl)

        step_host()
        barrier()
        a = torch.cuda.Event(enable_timing=True)
        z = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.e2e_steps):
            ncl, npl = step_host()
        z.record()
        barrier()
        e2e_ms = a.elapsed_time(z) / args.e2e_steps
        h2d = (8 * n) + (4 * n_new) + (8 * n + 24 * n)          # x; cellMap; scalar + vector field
        d2h = (4 * ncl) + (8 * n_new + 24 * n_new) + (4 * npl)  # cellsToRefine; mapped fields; split points
        e2e = (e2e_ms, h2d, d2h)

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms_per_step] + call_ms + [e2e[0] if e2e else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t[0])
        call_ms = [float(v) for v in t[1:NEV]]
        if e2e:
            e2e = (float(t[NEV]), e2e[1], e2e[2])
        cnt = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt)
        n_total = int(cnt[0])
    else:
        n_total = n

    if rank == 0:
        peak, peak_src = measured_peak()
        ab = algorithmic_bytes(n, f, p, nnz_pc, n_new, info["sweeps_refine"], info["sweeps_unrefine"], (n1, f1, p1, nnz_pc1))
        stage_ms = [sum(ms for ms, st in zip(call_ms, stage_of) if st == i) for i in range(4)]
        stage_bytes = [ab["E"], ab["M"] + ab["B"], ab["F"], ab["U"]]
        stages = {nm: {"ms": ms, "GBps": (by / 1e9) / (ms / 1e3) if ms > 0 else None}
                  for nm, ms, by in zip(stage_names, stage_ms, stage_bytes)}
        calls = {nm: ms for nm, ms in zip(call_names, call_ms)}
        step_bytes = sum(stage_bytes)
        value = n_total / (ms_per_step / 1e3) / 1e6
        # Dominant kernel = the single-kernel call with the largest time.  estimate and map_field are
        # exactly one launch per call (resident path), so the CUDA-event interval is the launch time.
        single = {"E:estimate(delta+normalize)": ("kp_estimate<0, 1>", ab["E"], "8f + 16n  (delta + fused normalize, TMA-staged SELL)"),
                  "F:map_field(scalar)": ("k_map_field<1>", 4 * n_new + 8 * n + 8 * n_new, "4n' + 8n + 8n'"),
                  "F:map_field(vector)": ("k_map_field_pair<3>", 4 * n_new + 3 * (8 * n + 8 * n_new), "4n' + 3(8n + 8n')")}
        dom = max(single, key=lambda k: calls[k])
        kname, kbytes, kform = single[dom]
        k_ms = calls[dom]
        # DRAM traffic per launch of that kernel from the committed ncu capture of this same command line
        traffic = None
        tpath = ROOT / "profiles" / "r1_traffic.json"
        if tpath.exists() and N == 400 and world == 1:
            try:
                tj = json.loads(tpath.read_text())["traffic_bytes_per_launch"]
                traffic = tj.get(kname)
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": kname, "achieved": (kbytes / 1e9) / (k_ms / 1e3), "peak": peak, "unit": "GB/s",
                "frac": (kbytes / 1e9) / (k_ms / 1e3) / peak, "traffic": traffic, "algorithmic_bytes": kbytes,
                "formula": kform, "ms": k_ms, "peak_source": peak_src,
                "other_kernels": {single[k][0]: {"ms": calls[k], "GBps": (single[k][1] / 1e9) / (calls[k] / 1e3),
                                                 "frac": (single[k][1] / 1e9) / (calls[k] / 1e3) / peak} for k in single if k != dom},
                "step_GBps": (step_bytes / 1e9) / (ms_per_step / 1e3),
                "step_frac": (step_bytes / 1e9) / (ms_per_step / 1e3) / peak,
                "step_bytes_per_cell": step_bytes / n}
        out = {
            "metric": "AMR step Mcells/s (estimate+mark+2:1 balance+map)", "value": value, "unit": "Mcells/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic {N}^3-cell blockMesh hex box per GPU ({n} cells, {f} internal faces, {p} points), "
                                   f"analytic density front, delta estimator, thresholds {THR}, maxRefinement {MAX_LEVEL}, "
                                   f"buffer layers {LAYERS_REF}/{LAYERS_UNREF}, hexRefiner; 1 scalar + 1 vector field mapped",
                       "decomposition": f"simple {procs}", "l2_policy": "inputs larger than L2 (>= 0.5 GB per array)",
                       "cells_per_gpu": n, "n_refine": info["n_refine"], "n_new": n_new,
                       "refined_mesh": {"cells": n1, "internal_faces": f1, "points": p1},
                       "sweeps_refine": info["sweeps_refine"], "sweeps_unrefine": info["sweeps_unrefine"],
                       "n_unrefine": info["n_unrefine"], "setup_s": setup_s},
            "stages": stages, "calls_ms": calls, "roofline": roof, "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e:
            out["e2e"] = {"value": n_total / (e2e[0] / 1e3) / 1e6, "unit": "Mcells/s", "ms_per_step": e2e[0],
                          "h2d_bytes_per_step": int(e2e[1]), "d2h_bytes_per_step": int(e2e[2])}
        if not args.no_cpu and world == 1:
            ncpu = args.cpu_cells
            O, S, nc = cpu_sample(ncpu, 1, 1)
            oracle_step(O, S)
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                oracle_step(O, S)
            dt = (time.perf_counter() - t0) / reps
            out["cpu_baseline"] = {"value": nc / dt / 1e6, "unit": "Mcells/s", "cores": 1, "kind": "port",
                                   "sample": f"{ncpu}^3-cell box ({nc} cells), same field/thresholds, {reps} steps, 1 thread"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        ctx.close()
        ctx1.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
