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
  F   fvMesh::mapFields through cellMap (identity + 7 children per refined parent): 1 scalar + 1 vector
  U   unrefinement branch: buffer layer, maxCellField, getSplitElems, filter, consistentUnrefinement
Topology change itself is host bookkeeping and outside the timed region (BASELINE.json north_star).

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


def algorithmic_bytes(n, f, p, nnz_pc, n_new, sweeps_ref, sweeps_unref):
    """each input array read once and each output written once per logical pass of
    the reference algorithm; label 4 B, scalar 8 B, flag 1 B."""
    E = 8 * f + 16 * n
    M1 = 9 * n
    M2 = LAYERS_REF * (8 * f + 2 * n + 8 * n)
    M4 = 10 * n
    B1 = sweeps_ref * (8 * f + 6 * n)
    F1 = 4 * (4 * n_new + 8 * n + 8 * n_new)
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

def oracle_step(O, w, xs, n_new_maps):
    """The same step on the CPU oracle (all ranks of the world)."""
    e = w.estimate("delta", xs)
    w.normalize_world(e, THR, GREAT, THR, GREAT)
    rc, rs, sw = w.hex_select_refine(e, MAX_LEVEL, LAYERS_REF)
    cms = [m[0] for m in n_new_maps]
    w.map_field_world(cms, [m[1] for m in n_new_maps], [m[3] for m in n_new_maps], 1)
    w.map_field_world(cms, [m[2] for m in n_new_maps], [m[4] for m in n_new_maps], 3)
    up, sp, sw2 = w.hex_select_unrefine(e, rc, LAYERS_UNREF)
    return rs, up, sw, sw2


def cpu_sample(n_side, ranks, threads):
    """Oracle on a bounded sample of the workload: an n_side^3 box split in `ranks` slabs."""
    from blastamr_b200 import mesh as M
    from oracle import oracle as O
    g = M.box_mesh(n_side, n_side, n_side)
    x = front_field(g.cc, n_side)
    if ranks > 1:
        cell_rank = M.simple_ranks(g, (1, 1, 1), (1, 1, ranks))
        loc = M.decompose(g, cell_rank, ranks)
        xs = [np.ascontiguousarray(x[l.cell_global]) for l in loc]
    else:
        loc, xs = [g], [x]
    w = O.World(loc)
    O.set_threads(threads)
    # cellMap from a first (untimed) pass
    e = w.estimate("delta", xs)
    w.normalize(e, THR, GREAT, THR, GREAT)
    rc, rs, _ = w.hex_select_refine(e, MAX_LEVEL, LAYERS_REF)
    maps = []
    for l, s, xx in zip(loc, rs, xs):
        cells = np.flatnonzero(s).astype(np.int32)
        cm = np.concatenate([np.arange(l.n, dtype=np.int32), np.repeat(cells, 7)])
        maps.append((cm, xx, np.stack([xx, xx, xx], axis=1).copy(), np.empty(len(cm)), np.empty((len(cm), 3))))
    return O, w, xs, maps, g.n


def run_reference(args):
    """--impl reference: the reference algorithm (restated: oracle/amr_oracle.c; the real reference needs
    OpenFOAM + MPI, absent here) on the host cores, one emulated MPI rank per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ranks = max(1, min(cores, 64))
    n_side = args.cpu_cells
    O, w, xs, maps, ncells = cpu_sample(n_side, ranks, ranks)
    for _ in range(args.warmup):
        oracle_step(O, w, xs, maps)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(O, w, xs, maps)
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
        mesh = M.box_mesh(N, N, N, sides=sides, patch_types=ptypes, patch_nbr=pnbr)
    else:
        mesh = M.box_mesh(N, N, N)
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
    pts_dev = torch.empty(max(p, 1), dtype=torch.int32, device=dev)
    thr = (THR, GREAT, THR, GREAT)

    # first pass: build the synthetic cellMap (identity + 7 children per refined parent)
    ctx.estimate("delta", x, thresholds=thr, keep_resident=True)
    n_ref, _ = ctx.select_refine(None, MAX_LEVEL, LAYERS_REF, cells_out=cells_dev)
    cell_map = torch.cat([torch.arange(n, dtype=torch.int32, device=dev),
                          cells_dev[:n_ref].repeat_interleave(7)]).contiguous()
    n_new = int(cell_map.numel())
    new1 = torch.empty(n_new, dtype=torch.float64, device=dev)
    new3 = torch.empty((n_new, 3), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    # one event interval per C-ABI call; F is two calls (scalar field, vector field)
    call_names = ["E:estimate(delta+normalize)", "M+B:select_refine", "F:map_field(scalar)", "F:map_field(vector)",
                  "U:select_unrefine"]
    stage_of = [0, 1, 2, 2, 3]
    stage_names = ["E", "M+B", "F", "U"]
    NEV = len(call_names) + 1
    info = {}

    def step(evs=None):
        def mark(i):
            if evs is not None:
                evs[i].record()
        mark(0)
        ctx.estimate("delta", x, thresholds=thr, keep_resident=True)
        mark(1)
        nr, sw = ctx.select_refine(None, MAX_LEVEL, LAYERS_REF, cells_out=cells_dev)
        mark(2)
        ctx.map_field(cell_map, x, n_old=n, n_new=n_new, ncomp=1, out=new1)
        mark(3)
        ctx.map_field(cell_map, fld3, n_old=n, n_new=n_new, ncomp=3, out=new3)
        mark(4)
        npts, sw2 = ctx.select_unrefine(None, LAYERS_UNREF, elems_out=pts_dev)
        mark(5)
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
    l0 = ctx.launches
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
    launches = ctx.launches - l0
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
        cells_h = np.empty(max(n, 1), dtype=np.int32)
        pts_h = np.empty(max(p, 1), dtype=np.int32)

        def step_host():
            ctx.estimate("delta", xh.numpy(), thresholds=thr, keep_resident=True)
            cl, _ = ctx.select_refine(None, MAX_LEVEL, LAYERS_REF, cells_out=cells_h)
            ctx.map_field(cmh.numpy(), xh.numpy(), n_old=n, n_new=n_new, ncomp=1, out=n1h.numpy())
            ctx.map_field(cmh.numpy(), f3h.numpy(), n_old=n, n_new=n_new, ncomp=3, out=n3h.numpy())
            pl, _ = ctx.select_unrefine(None, LAYERS_UNREF, elems_out=pts_h)
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
        h2d = 8 * n + 2 * 4 * n_new + 8 * n + 24 * n     # x, cellMap (twice), scalar field, vector field
        d2h = 4 * ncl + 8 * n_new + 24 * n_new + 4 * npl
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
        ab = algorithmic_bytes(n, f, p, nnz_pc, n_new, info["sweeps_refine"], info["sweeps_unrefine"])
        stage_ms = [sum(ms for ms, st in zip(call_ms, stage_of) if st == i) for i in range(4)]
        stage_bytes = [ab["E"], ab["M"] + ab["B"], ab["F"], ab["U"]]
        stages = {nm: {"ms": ms, "GBps": (by / 1e9) / (ms / 1e3) if ms > 0 else None}
                  for nm, ms, by in zip(stage_names, stage_ms, stage_bytes)}
        calls = {nm: ms for nm, ms in zip(call_names, call_ms)}
        step_bytes = sum(stage_bytes)
        value = n_total / (ms_per_step / 1e3) / 1e6
        # Dominant kernel = the single-kernel call with the largest time.  estimate and map_field are
        # exactly one launch per call (resident path), so the CUDA-event interval is the launch time.
        single = {"E:estimate(delta+normalize)": ("kp_estimate<0,true,false> (delta + fused normalize, TMA-staged SELL)", ab["E"], "8f + 16n"),
                  "F:map_field(scalar)": ("k_map_field<1>", 4 * n_new + 8 * n + 8 * n_new, "4n' + 8n + 8n'"),
                  "F:map_field(vector)": ("k_map_field<3>", 4 * n_new + 3 * (8 * n + 8 * n_new), "4n' + 3(8n + 8n')")}
        dom = max(single, key=lambda k: calls[k])
        kname, kbytes, kform = single[dom]
        k_ms = calls[dom]
        roof = {"bound": "hbm", "kernel": kname, "achieved": (kbytes / 1e9) / (k_ms / 1e3), "peak": peak, "unit": "GB/s",
                "frac": (kbytes / 1e9) / (k_ms / 1e3) / peak, "traffic": None, "algorithmic_bytes": kbytes,
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
                       "sweeps_refine": info["sweeps_refine"], "sweeps_unrefine": info["sweeps_unrefine"],
                       "n_unrefine": info["n_unrefine"], "setup_s": setup_s},
            "stages": stages, "calls_ms": calls, "roofline": roof, "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e:
            out["e2e"] = {"value": n_total / (e2e[0] / 1e3) / 1e6, "unit": "Mcells/s", "ms_per_step": e2e[0],
                          "h2d_bytes_per_step": int(e2e[1]), "d2h_bytes_per_step": int(e2e[2])}
        if not args.no_cpu and world == 1:
            ncpu = args.cpu_cells
            O, w, xs, maps, nc = cpu_sample(ncpu, 1, 1)
            oracle_step(O, w, xs, maps)
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                oracle_step(O, w, xs, maps)
            dt = (time.perf_counter() - t0) / reps
            out["cpu_baseline"] = {"value": nc / dt / 1e6, "unit": "Mcells/s", "cores": 1, "kind": "port",
                                   "sample": f"{ncpu}^3-cell box ({nc} cells), same field/thresholds, {reps} steps, 1 thread"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
