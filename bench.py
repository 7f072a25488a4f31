#!/usr/bin/env python
"""bench.py -- AMR decision step throughput (estimate + mark + 2:1 balance + map) on B200.

Workload (BASELINE.json configs[4], SURVEY.md 8(d) "C5"):
  S2 (default)  the 64 M-cell hex box of C5 as the AMR loop itself leaves it: a blockMesh box of B^3 level-0 cells
                (default B = 176) refined to maxRefinement 3 around the analytic density front
                rho(x) = 1 + 0.5 (1 + tanh((|x-c|-r0)/w)), r0 = 0.3, w = 0.8/B (6.4 level-3 cells), by three
                refine-only update() cycles and one full (refine + unrefine) cycle of the path itself -- about 64 M
                leaves in four levels -- and then the front MOVED by `--shift-cells` (default 4) level-3 cells, so
                that in the timed step refinement, 2:1 propagation (several sweeps) and unrefinement all fire.
  S1 (--workload S1)  round 1's uniform level-0 box of N^3 cells (default N = 400), unshifted front.
  P2 (--workload P2)  the polyhedral engine on the S2 case: `refiner polyRefiner` (polyhedralRefinement: cells as general
                polyhedra over faces()/edges(), face AND edge 2:1 consistency, refinement.C:202-387; buffer layers across
                POINTS and pointLevel-based split points on unrefinement, fvMeshPolyRefiner.C:48-96,
                polyhedralRefinement.C:2110-2205; nUnrefinementBufferLayers >= 2, fvMeshPolyRefiner.C:287-300), default
                base 64 (~4 M leaves: the host topology engine has to build faces() and edges() as well).
`delta` estimator on rho, lowerRefineLevel = unrefineLevel = 0.01, maxRefinement 3, nBufferLayers 1/1, hexRefiner.
One "step" = what adaptiveFvMesh::update() asks of the decision path in one time step:
  E    errorEstimator::update (delta + normalize)                                            on mesh M0
  M+B  selectRefineCandidates + buffer layer + protections + selectRefineCells +
       hexRef::consistentRefinement sweeps + ascending cellsToRefine                          on M0
  F    fvMesh::mapFields through the real cellMap of the cut M0 -> M1 (hexRef3D numbering): 1 scalar + 1 vector
       field, the error field, and refineCell through cellMap/reverseCellMap (fvMeshHexRefiner.C:1562-1586)
  U    the unrefinement branch ON THE REFINED MESH M1: buffer layer, maxCellField, getSplitElems, filter,
       consistentUnrefinement sweeps, ascending split-point list
  F2   fvMesh::mapFields through the cellMap of the merge M1 -> M2 (the master cell keeps its value: parity mode)
The topology changes themselves (host octree engine, polyTopoChange in the reference) are host bookkeeping and
outside the timed region (BASELINE.json north_star); M0 and M1 stay resident in two contexts.

`value`    = cells of M0 (all ranks) / max-over-ranks CUDA-event time of the step, inputs resident in HBM.
`e2e`      = same metric through the C-ABI with HOST (pinned) buffers for everything the solver owns: the field and
             the registered fields cross PCIe inside the timed region, every step.
Multi-GPU  = STRONG scaling: the ONE box is split `simple` (1 1 1)/(2 1 1)/(2 2 1)/(2 2 2) into processor patches
             (rank 0 runs the host topology engine on the global mesh and hands every rank its sub-domain); the front
             is centred on the box, so it crosses every processor patch and every halo swap carries decisions.
After the timed loop the decisions are checked: the per-rank cellsToRefine must equal the serial run on the
undecomposed mesh (decomposition independence, full size) and a reduced-size instance of the same case is compared
with the CPU oracle, list by list (`parity`).
"""
from __future__ import annotations

import argparse
import datetime
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
THRS = (THR, GREAT, THR, GREAT)
METRIC = "AMR step Mcells/s (estimate+mark+2:1 balance+map)"


# ------------------------------------------------------------------------------- the case

class Case:
    """S2: graded box, moved front.  S1: uniform box, static front."""

    def __init__(self, workload, base, shift_cells):
        self.workload = workload
        self.base = int(base)
        self.poly = workload == "P2"
        self.graded = workload in ("S2", "P2")
        self.kind = 2 if self.poly else 0                       # AMR_REFINER_POLY3D / AMR_REFINER_HEX3D
        self.layers_ref = LAYERS_REF
        self.layers_unref = max(LAYERS_UNREF, 2) if self.poly else LAYERS_UNREF    # fvMeshPolyRefiner.C:287-300
        if self.graded:
            self.w = 0.8 / self.base
            self.h_fine = 1.0 / (self.base << MAX_LEVEL)
            self.shift = float(shift_cells) * self.h_fine
            self.shift_cells = float(shift_cells)
        else:
            self.w = 2.0 / self.base
            self.h_fine = 1.0 / self.base
            self.shift = 0.0
            self.shift_cells = 0.0
        self.c_build = (0.5, 0.5, 0.5)
        self.c_step = (0.5 + self.shift, 0.5, 0.5)

    def field(self, cc, centre=None):
        c = np.asarray(self.c_step if centre is None else centre)
        d = np.sqrt(((cc - c) ** 2).sum(axis=1))
        return 1.0 + 0.5 * (1.0 + np.tanh((d - 0.3) / self.w))

    def describe(self):
        if self.graded:
            engine = ("polyRefiner (polyhedralRefinement on faces()/edges(): face + edge consistency, point buffer layers, "
                      "pointLevel split points)") if self.poly else "hexRefiner"
            return (f"{self.workload}: synthetic blockMesh hex box of {self.base}^3 level-0 cells graded to maxRefinement {MAX_LEVEL} around the "
                    f"analytic density front by the AMR loop itself (3 refine cycles + 1 refine/unrefine cycle), front then moved "
                    f"by {self.shift_cells:g} level-{MAX_LEVEL} cells; delta estimator, thresholds {THR}, buffer layers "
                    f"{self.layers_ref}/{self.layers_unref}, {engine}; 1 scalar + 1 vector + the error field mapped on refinement and on "
                    f"unrefinement")
        return (f"S1: synthetic {self.base}^3-cell uniform blockMesh hex box, analytic density front, delta estimator, thresholds "
                f"{THR}, maxRefinement {MAX_LEVEL}, buffer layers {LAYERS_REF}/{LAYERS_UNREF}, hexRefiner; 1 scalar + 1 vector + the "
                f"error field mapped on refinement and on unrefinement")


def unrefine_maps(pc_off, pc_cells, pts, n1):
    """mapPolyMesh::cellMap / reverseCellMap of a hexRef3D merge (hexRef3D.C:2078,2130): the 8 cells around every split
    point collapse into the lowest-numbered one, labels are compacted preserving order.  Purely local: what the host
    topology engine returns (blastamr_b200/mesh.py Octree.unrefine), restated with numpy so that a rank of a decomposed
    run needs no global octree for it."""
    pts = np.asarray(pts, dtype=np.int64)
    removed = np.zeros(n1, dtype=bool)
    master_of = np.arange(n1, dtype=np.int64)
    if len(pts):
        rows = pc_cells[(pc_off[pts].astype(np.int64)[:, None] + np.arange(8)[None, :])]
        assert np.all(pc_off[pts + 1] - pc_off[pts] == 8)
        master = rows.min(axis=1)
        removed[rows.ravel()] = True
        removed[master] = False
        master_of[rows.ravel()] = np.repeat(master, 8)
    keep = ~removed
    cm2 = np.flatnonzero(keep).astype(np.int32)
    new_id = np.cumsum(keep) - 1
    rev2 = np.where(keep, new_id, -new_id[master_of] - 2).astype(np.int32)
    return cm2, rev2


def local_maps(n0_glob, n1_glob, cm_g, cg0, cg1):
    """cellMap / reverseCellMap of one rank from the global ones (decomposed run): local labels ascend with the global"""
    if cg0 is None:
        return np.ascontiguousarray(cm_g, dtype=np.int32), np.arange(n0_glob, dtype=np.int32)
    old_local = np.full(n0_glob, -1, np.int32)
    old_local[cg0] = np.arange(len(cg0), dtype=np.int32)
    cm = old_local[cm_g[cg1]]
    assert (cm >= 0).all()
    new_local = np.full(n1_glob, -1, np.int32)
    new_local[cg1] = np.arange(len(cg1), dtype=np.int32)
    rv = new_local[cg0]                       # an old cell keeps its global label in M1 (child 0)
    return np.ascontiguousarray(cm, dtype=np.int32), np.ascontiguousarray(rv, dtype=np.int32)


# ------------------------------------------------------------------------------- the path behind one interface
# (the build-up cycles run on the product for the GPU arm and on the oracle for the CPU arms)

class GpuPath:
    def __init__(self, device, case):
        from blastamr_b200 import capi
        self.capi = capi
        self.device = device
        self.case = case
        self.ctx = capi.Context(device)

    def refine_decision(self, mesh, x):
        ctx = self.ctx
        ctx.set_mesh(mesh)
        ctx.estimate("delta", x, thresholds=THRS, keep_resident=True)
        cells, _ = ctx.select_refine(None, MAX_LEVEL, self.case.layers_ref, kind=self.case.kind, cells_out=np.empty(max(mesh.n, 1), np.int32))
        return cells.copy()

    def unrefine_decision_same_mesh(self):
        pts, _ = self.ctx.select_unrefine(None, self.case.layers_unref, kind=self.case.kind)
        return pts.copy()

    def close(self):
        self.ctx.close()


def oracle_select_refine(w, case, e):
    if case.poly:
        return w.poly_select_refine(e, MAX_LEVEL, case.layers_ref)
    return w.hex_select_refine(e, MAX_LEVEL, case.layers_ref)


def oracle_select_unrefine(w, case, e, rc):
    if case.poly:
        return w.poly_select_unrefine(e, rc, case.layers_unref)
    return w.hex_select_unrefine(e, rc, case.layers_unref)


class OraclePath:
    def __init__(self, case):
        from oracle import oracle as O
        self.O = O
        self.case = case

    def refine_decision(self, mesh, x):
        w = self.O.World([mesh])
        e = w.estimate("delta", [x])
        w.normalize(e, *THRS)
        rc, rs, _ = oracle_select_refine(w, self.case, e)
        self._w, self._e, self._rc = w, e, rc
        return np.flatnonzero(rs[0]).astype(np.int32)

    def unrefine_decision_same_mesh(self):
        up, _, _ = oracle_select_unrefine(self._w, self.case, self._e, [self._rc[0].copy()])
        return np.flatnonzero(up[0]).astype(np.int32)

    def close(self):
        self._w = None


def build_initial(oc, path, case, log=None):
    """The mesh the timed step starts from.  S2: update() cycles of the path on the static front -- refine-only until
    maxRefinement is reached, then one full cycle whose unrefinement branch removes what the buffer layers over-refined."""
    if not case.graded:
        return {"cycles": 0}
    hist = []
    for _ in range(MAX_LEVEL):
        m = oc.build(points=case.poly, faces=case.poly)
        cells = path.refine_decision(m, case.field(m.cc, case.c_build))
        hist.append((int(m.n), int(len(cells)), 0))
        del m
        if len(cells):
            oc.refine(cells)
    m = oc.build(points=True, faces=case.poly)
    cells = path.refine_decision(m, case.field(m.cc, case.c_build))
    if len(cells):                  # rare on a static front: cut, then evaluate again on the cut mesh (build-up only; any
        oc.refine(cells)            # 2:1-balanced mesh is a valid starting point, a non-empty second list is left to the step)
        m = oc.build(points=True, faces=case.poly)
        path.refine_decision(m, case.field(m.cc, case.c_build))
    pts = path.unrefine_decision_same_mesh()
    hist.append((int(m.n), int(len(cells)), int(len(pts))))
    if len(pts):
        oc.unrefine(m, pts)
    del m
    if log:
        log(f"build-up cycles (cells, refined, split points merged): {hist} -> {oc.ncells} cells")
    return {"cycles": len(hist), "history": hist}


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
            time.sleep(0.004)               # the timed region is ~10 ms at 8 GPUs

    def reset(self):
        self.samples = []
        self.reasons = set()

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        if self.ok and not self.samples:           # a timed region shorter than one sampling period: the clock right at its end
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            except Exception:
                pass
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


def bind_host_threads(local_rank, world):
    """NUMA-aware placement of a rank's host threads (pinned staging, upload lanes, OpenMP of the host topology engine):
    the cores of the GPU's NUMA node, divided between the ranks that share them.  Best effort; returns a description."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        node_cpus = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(physical_device_index(local_rank))
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            node = int(Path(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node").read_text())
            if node >= 0:
                cpus = []
                for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus += list(range(int(a), int(b or a) + 1))
                node_cpus = [c for c in cpus if c in allowed]
        except Exception:
            node_cpus = None
        pool = node_cpus if node_cpus else allowed
        if world > 1 and len(pool) >= world:
            k = len(pool) // world
            mine = pool[local_rank * k:(local_rank + 1) * k]
        else:
            mine = pool
        if mine:
            os.sched_setaffinity(0, mine)
        return {"cores": len(mine), "numa_local": bool(node_cpus)}
    except Exception as e:      # noqa: BLE001
        return {"cores": None, "error": str(e)[:80]}


def measured_peak():
    pth = ROOT / "MEASURED_PEAKS.json"
    if pth.exists():
        try:
            return float(json.loads(pth.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def checksum(ids):
    """order-free 61-bit checksum of a set of (global) labels: equal sets <=> equal sums of a mixing hash"""
    a = np.asarray(ids, dtype=np.uint64)
    h = (a + np.uint64(0x9E3779B97F4A7C15)) * np.uint64(0xBF58476D1CE4E5B9)
    h ^= h >> np.uint64(31)
    h *= np.uint64(0x94D049BB133111EB)
    h ^= h >> np.uint64(29)
    return int(h.sum(dtype=np.uint64) & np.uint64((1 << 61) - 1))


def algorithmic_bytes(m0, m1, n2, sweeps_ref, sweeps_unref, case=None, edges=None):
    """each input array read once and each output written once per logical pass of the REFERENCE
    algorithm (dense face loops); label 4 B, scalar 8 B, flag 1 B.  m0 = (n, f), m1 = (n, f, p, nnz pointCells).
    This is what the reference's loops would move -- NOT what the kernels move (their sparse forms skip most of it);
    the measured DRAM bytes are reported separately."""
    n, f = m0
    n1, f1, p1, nnz = m1
    E = 8 * f + 16 * n
    M = 9 * n + LAYERS_REF * (8 * f + 2 * n + 8 * n) + 10 * n
    B = sweeps_ref * (8 * f + 6 * n)
    F = 5 * (4 * n1 + 8 * n + 8 * n1) + (4 * n1 + 4 * n + n + n1)
    csr = 4 * (p1 + 1) + 4 * nnz
    U = LAYERS_UNREF * (8 * f1 + 2 * n1) + (csr + 8 * n1 + 8 * p1) + (8 * n1 + csr) + (9 * p1 + n1) + \
        sweeps_unref * ((8 * f1 + 6 * n1) + 2 * csr + 2 * p1)
    F2 = 5 * (4 * n2 + 8 * n1 + 8 * n2)
    if case is not None and case.poly and edges is not None:
        # polyRefiner: B2 per sweep = B1 + edgeCells CSR; buffer layers across points = cells -> points -> cells over the
        # pointCells CSR; U3 split-point rule reads faces() + pointLevel; B4 per sweep = B3 + edgeCells CSR
        (e0, nnz_e0), (e1, nnz_e1), nnz_f1 = edges
        B = sweeps_ref * ((8 * f + 6 * n) + 4 * nnz_e0 + 4 * (e0 + 1))
        U = case.layers_unref * (2 * csr + 2 * n1 + 2 * p1) + (csr + 8 * n1 + 8 * p1) + (4 * nnz_f1 + 4 * (f1 + 1) + 4 * p1 + csr) + \
            (9 * p1 + n1) + sweeps_unref * ((8 * f1 + 6 * n1) + 2 * csr + 2 * p1 + 4 * nnz_e1 + 4 * (e1 + 1))
    return {"E": E, "M+B": M + B, "F": F, "U": U, "F2": F2}


# ------------------------------------------------------------------------------- CPU arms (the oracle)

def divisor_at_most(n, k):
    for r in range(min(n, k), 0, -1):
        if n % r == 0:
            return r
    return 1


def oracle_prepare(case, ranks, threads, log=None):
    """The same case on the CPU oracle: build-up cycles, slab decomposition into `ranks` emulated MPI ranks, first pass
    (host topology engine makes M1 and the merge map), everything a step needs."""
    from blastamr_b200 import mesh as M
    from oracle import oracle as O
    O.set_threads(threads)
    lcap = MAX_LEVEL if case.graded else 1
    oc = M.Octree(case.base, case.base, case.base, (1.0, 1.0, 1.0), lcap)
    path = OraclePath(case)
    build_initial(oc, path, case, log)
    path.close()
    g0 = oc.build(points=case.poly, faces=case.poly)

    def split(g):
        if ranks > 1:
            cr = M.simple_ranks(g, (1, 1, 1), (1, 1, ranks))
            return [M.decompose_one(g, cr, ranks, r) for r in range(ranks)]
        return [g]

    loc0 = split(g0)
    xs = [np.ascontiguousarray(case.field(l.cc)) for l in loc0]
    w = O.World(loc0)
    e = w.estimate("delta", xs)
    w.normalize_world(e, *THRS)
    rc, rs, sw = oracle_select_refine(w, case, e)
    if ranks > 1:
        cells = np.sort(np.concatenate([l.cell_global[np.flatnonzero(s_)] for l, s_ in zip(loc0, rs)])).astype(np.int32)
    else:
        cells = np.flatnonzero(rs[0]).astype(np.int32)
    cm_g = oc.refine(cells)
    g1 = oc.build(points=True, faces=case.poly)
    loc1 = split(g1)
    w1 = O.World(loc1)
    cms, revs = [], []
    for l0, l1 in zip(loc0, loc1):
        cm, rv = local_maps(g0.n, g1.n, cm_g, l0.cell_global if ranks > 1 else None, l1.cell_global if ranks > 1 else None)
        cms.append(cm); revs.append(rv)
    f3 = [np.stack([a, 2 * a, -a], axis=1).copy() for a in xs]
    S = dict(O=O, w=w, w1=w1, case=case, xs=xs, f3=f3, cms=cms, revs=revs, n0=int(g0.n),
             new1=[np.empty(len(c)) for c in cms], new3=[np.empty((len(c), 3)) for c in cms], err1=[np.empty(len(c)) for c in cms])
    # first pass of the unrefinement branch -> the merge maps
    w.map_field_world(cms, e, S["err1"], 1)
    rc1 = [O.remap_refine_cell(cm, rv, r) for cm, rv, r in zip(cms, revs, rc)]
    up, _, sw2 = oracle_select_unrefine(w1, case, S["err1"], rc1)
    cm2s = [unrefine_maps(l1.pc_off, l1.pc_cells, np.flatnonzero(u), l1.n)[0] for l1, u in zip(loc1, up)]
    S.update(cm2s=cm2s, x2=[np.empty(len(c)) for c in cm2s], u2=[np.empty((len(c), 3)) for c in cm2s], err2=[np.empty(len(c)) for c in cm2s],
             n_refine=int(len(cells)), n_unrefine=int(sum(int(u.sum()) for u in up)), sweeps_refine=int(sw), sweeps_unrefine=int(sw2),
             n1=int(g1.n), n2=int(sum(len(c) for c in cm2s)), keep=(g0, g1, loc0, loc1, oc))
    return S


def oracle_step(S):
    """one step on the CPU oracle (all emulated ranks): E, M+B on M0; F through the cellMap; U on M1; F2 through the merge map"""
    O, w, w1 = S["O"], S["w"], S["w1"]
    e = w.estimate("delta", S["xs"])
    w.normalize_world(e, *THRS)
    rc, rs, sw = oracle_select_refine(w, S["case"], e)
    cms = S["cms"]
    w.map_field_world(cms, S["xs"], S["new1"], 1)
    w.map_field_world(cms, S["f3"], S["new3"], 3)
    w.map_field_world(cms, e, S["err1"], 1)
    rc1 = [O.remap_refine_cell(cm, rv, r) for cm, rv, r in zip(cms, S["revs"], rc)]
    up, sp, sw2 = oracle_select_unrefine(w1, S["case"], S["err1"], rc1)
    cm2s = S["cm2s"]
    w1.map_field_world(cm2s, S["new1"], S["x2"], 1)
    w1.map_field_world(cm2s, S["new3"], S["u2"], 3)
    w1.map_field_world(cm2s, S["err1"], S["err2"], 1)
    return rs, up, sw, sw2


def run_reference(args):
    """--impl reference: the reference algorithm (restated: oracle/amr_oracle.c; the real reference needs OpenFOAM + MPI,
    absent here) on the host cores, one emulated MPI rank per core, on the SAME workload as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    base = args.ref_base or args.base
    case = Case(args.workload, base, args.shift_cells)
    ranks = divisor_at_most(base, max(1, min(cores, 64)))       # slabs of whole level-0 layers (>= 25 fine layers each)
    t0 = time.time()
    S = oracle_prepare(case, ranks, ranks, log=lambda s: sys.stderr.write(s + "\n"))
    setup_s = time.time() - t0
    for _ in range(args.warmup):
        oracle_step(S)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(S)
    dt = (time.perf_counter() - t0) / args.steps
    val = S["n0"] / dt / 1e6
    sample = "the full workload" if base == args.base else f"bounded sample: base {base}^3 instead of {args.base}^3"
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mcells/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": Case(args.workload, args.base, args.shift_cells).describe(), "cells": S["n0"], "n_refine": S["n_refine"],
                   "n_unrefine": S["n_unrefine"], "sweeps_refine_gauss_seidel": S["sweeps_refine"],
                   "sweeps_unrefine": S["sweeps_unrefine"], "refined_mesh_cells": S["n1"], "merged_mesh_cells": S["n2"],
                   "decomposition": f"simple (1 1 {ranks}) over the host cores", "setup_s": setup_s},
        "cpu_baseline": {"value": val, "unit": "Mcells/s", "cores": ranks, "kind": "port",
                         "sample": f"{sample}; {S['n0']} cells in {ranks} emulated MPI ranks (one thread each), {args.steps} steps"},
        "e2e": {"value": val, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------- multi-rank plumbing (host side)

MESH_KEYS = ("owner", "neighbour", "cc", "pc_off", "pc_cells", "bnd_point", "cell_level", "point_level", "visible", "split_parent",
             "cell_global", "point_global", "cp_off", "cp_pts", "edge_pts", "ec_off", "ec_cells", "bnd_edge", "face_off", "face_pts",
             "patch_pt_off", "patch_pts")


def mesh_pack(m):
    d = {k: getattr(m, k) for k in MESH_KEYS if getattr(m, k, None) is not None}
    d["_sizes"] = np.array([m.n, m.f, m.b, m.p, int(getattr(m, "n_edges", 0) or 0)], dtype=np.int64)
    st, sz, ty, nb = m.patch_arrays()
    d["_patches"] = np.stack([st, sz, ty, nb]).astype(np.int32)
    return d


def mesh_unpack(d):
    from blastamr_b200 import mesh as M
    n, f, b, p, ne = (int(v) for v in d["_sizes"])
    pa = d["_patches"]
    patches = [M.Patch(f"patch{i}" if int(pa[2, i]) != 2 else f"procBoundaryTo{int(pa[3, i])}", int(pa[2, i]), int(pa[0, i]),
                       int(pa[1, i]), int(pa[3, i])) for i in range(pa.shape[1])]
    m = M.PolyMesh(n=n, f=f, b=b, p=p, owner=d["owner"], neighbour=d["neighbour"], patches=patches)
    for k in MESH_KEYS[2:]:
        if k in d:
            setattr(m, k, d[k])
    if m.split_parent is None:
        m.split_parent = np.zeros(0, np.int32)
    m.n_edges = ne
    return m


def send_arrays(dist, torch, group, d, dst):
    meta = [(k, str(v.dtype), tuple(v.shape)) for k, v in d.items()]
    dist.send_object_list([meta], dst=dst, group=group)
    for k, v in d.items():
        if v.size:
            dist.send(torch.from_numpy(np.ascontiguousarray(v).reshape(-1).view(np.uint8)), dst=dst, group=group)


def recv_arrays(dist, torch, group, src):
    box = [None]
    dist.recv_object_list(box, src=src, group=group)
    out = {}
    for k, dt, shape in box[0]:
        a = np.empty(shape, dtype=np.dtype(dt))
        if a.size:
            dist.recv(torch.from_numpy(a.reshape(-1).view(np.uint8)), src=src, group=group)
        out[k] = a
    return out


# ------------------------------------------------------------------------------- GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="S2", choices=["S1", "S2", "P2"])
    ap.add_argument("--base", type=int, default=None, help="level-0 cells per side of the box (S2 default 176 -> ~64 M leaves; S1 default 400)")
    ap.add_argument("--cells", type=int, default=None, help="alias of --base (round-1 flag)")
    ap.add_argument("--shift-cells", type=float, default=4.0, help="S2: displacement of the front in level-3 cells")
    ap.add_argument("--ref-base", type=int, default=None, help="--impl reference: bounded sample (default: the full workload)")
    ap.add_argument("--cpu-base", type=int, default=None, help="base of the single-thread cpu_baseline sample (default base/4)")
    ap.add_argument("--parity-base", type=int, default=16, help="base of the reduced-size GPU-vs-oracle check")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-handover", action="store_true", help="skip the topology hand-over timing (amr_mesh_set vs amr_mesh_update)")
    args = ap.parse_args()
    if args.base is None:
        args.base = args.cells if args.cells else {"S2": 176, "P2": 64}.get(args.workload, 400)
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
    gl = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        gl = dist.new_group(backend="gloo", timeout=datetime.timedelta(hours=1))     # host-side hand-over of the sub-domains
    procs = PROCS.get(world, (world, 1, 1))
    case = Case(args.workload, args.base, args.shift_cells)
    lcap = MAX_LEVEL if case.graded else 1
    host_binding = bind_host_threads(local_rank, world) if (world > 1 and rank != 0) else {"cores": len(os.sched_getaffinity(0)), "note": "rank 0 keeps all cores for the host topology engine"}

    def log(s):
        if rank == 0:
            sys.stderr.write(f"[bench +{time.time() - t_setup:6.1f}s] {s}\n")
            sys.stderr.flush()

    def new_ctx():
        c = capi.Context(local_rank)
        c.set_stream(torch.cuda.current_stream().cuda_stream)
        if world > 1:
            uid = [capi.Context.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0, group=gl)
            c.comm_init(world, rank, uid[0])
        return c

    def hand_over(g, want):
        """rank 0: decompose the global mesh `g` and send every rank its sub-domain; returns this rank's mesh and (rank 0) the
        cell_global lists of all ranks"""
        if world == 1:
            return g, None
        if rank == 0:
            cr = M.simple_ranks(g, (1, 1, 1), procs)
            mine, cgs = None, []
            for r in range(world):
                l = M.decompose_one(g, cr, world, r)
                cgs.append(l.cell_global.copy())
                if r == 0:
                    mine = l
                else:
                    send_arrays(dist, torch, gl, mesh_pack(l), r)
                    del l
            return mine, cgs
        return mesh_unpack(recv_arrays(dist, torch, gl, 0)), None

    want_handover = (not args.no_handover) and world == 1 and not case.poly
    # ---- meshes (host bookkeeping, untimed): rank 0 owns the global octree
    t_setup = time.time()
    ctx, ctx1 = new_ctx(), new_ctx()
    oc = gser = None
    build_info = {}
    g0 = None
    if rank == 0:
        oc = M.Octree(case.base, case.base, case.base, (1.0, 1.0, 1.0), lcap)
        gser = GpuPath(local_rank, case)
        build_info = build_initial(oc, gser, case, log)
        g0 = oc.build(points=case.poly or want_handover, faces=case.poly)     # (the hand-over timing carries M0's point addressing)
        log(f"M0 built: {g0.n} cells")
    mesh0, cgs0 = hand_over(g0, "M0")
    n, f, b = mesh0.n, mesh0.f, mesh0.b
    edges0 = (int(getattr(mesh0, "n_edges", 0) or 0), int(mesh0.ec_off[-1]) if getattr(mesh0, "ec_off", None) is not None else 0)
    ctx.set_mesh(mesh0)
    x_host = np.ascontiguousarray(case.field(mesh0.cc))
    x = torch.from_numpy(x_host).to(dev)
    fld3 = torch.stack([x, 2 * x, -x], dim=1).contiguous()
    cells_dev = torch.empty(max(n, 1), dtype=torch.int32, device=dev)

    # first pass (untimed): the decision, then the host topology engine cuts M0 -> M1 and hands back the real cellMap
    ctx.estimate("delta", x, thresholds=THRS, keep_resident=True)
    cells_first, _ = ctx.select_refine(None, MAX_LEVEL, case.layers_ref, kind=case.kind, cells_out=np.empty(max(n, 1), dtype=np.int32))
    cells_first = cells_first.copy()
    gcells = mesh0.cell_global[cells_first] if world > 1 else cells_first
    decomposition_independent = None
    g1 = None
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(np.ascontiguousarray(gcells), gathered, dst=0, group=gl)
        if rank == 0:
            gcells_all = np.sort(np.concatenate(gathered)).astype(np.int32)
    else:
        gcells_all = cells_first
    n0_total = None
    if rank == 0:
        if world > 1:
            # the same decision by ONE context on the undecomposed mesh: the decomposed run must select exactly these cells
            cells_serial = gser.refine_decision(g0, case.field(g0.cc))
            decomposition_independent = bool(np.array_equal(cells_serial, gcells_all))
            log(f"decomposition independence of cellsToRefine at full size: {decomposition_independent}")
        n0_total = int(g0.n)
        n0_glob = int(g0.n)
        del g0
        gser.close()
        cm_g = oc.refine(gcells_all.copy())
        g1 = oc.build(points=True, faces=case.poly)
        n1_glob = int(g1.n)
        log(f"M1 built: {g1.n} cells")
    do_handover = want_handover
    if not do_handover:
        del mesh0
    mesh1, cgs1 = hand_over(g1, "M1")
    if world > 1:
        if rank == 0:
            mine = None
            for r in range(world):
                cm_r, rv_r = local_maps(n0_glob, n1_glob, cm_g, cgs0[r], cgs1[r])
                if r == 0:
                    mine = (cm_r, rv_r)
                else:
                    send_arrays(dist, torch, gl, {"cm": cm_r, "rv": rv_r}, r)
            cell_map_h, rev_h = mine
            del cm_g, cgs0, cgs1
        else:
            d = recv_arrays(dist, torch, gl, 0)
            cell_map_h, rev_h = d["cm"], d["rv"]
    else:
        cell_map_h, rev_h = local_maps(n0_glob, n1_glob, cm_g, None, None)
        del cm_g
    del g1
    n1, f1, b1, p1 = mesh1.n, mesh1.f, mesh1.b, mesh1.p
    nnz_pc1 = int(mesh1.pc_off[p1])
    edges1 = (int(getattr(mesh1, "n_edges", 0) or 0), int(mesh1.ec_off[-1]) if getattr(mesh1, "ec_off", None) is not None else 0)
    nnz_f1 = int(mesh1.face_off[-1]) if getattr(mesh1, "face_off", None) is not None else 0
    ctx1.set_mesh(mesh1)
    handover = None
    if do_handover:
        # N1: what the hand-over of the cut M0 -> M1 costs through the C-ABI -- the whole mesh (amr_mesh_set + points + levels)
        # against the incremental form (amr_mesh_update + amr_mesh_update_points: maps + dirty rows).  Host arrays are
        # pageable, as OpenFOAM's are.  The delta itself is what polyTopoChange knows; deriving it here is host bookkeeping.
        t0 = time.time()
        delta = M.mesh_delta(mesh0, mesh1, cell_map_h)
        t_delta = time.time() - t0
        ctxh = new_ctx()
        ctxh.set_mesh(mesh0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctxh.update_mesh(mesh1, delta)
        t_inc = time.perf_counter() - t0
        xe = np.ascontiguousarray(case.field(mesh1.cc))
        ea = ctx1.estimate("delta", xe, thresholds=THRS)
        eb = ctxh.estimate("delta", xe, thresholds=THRS)
        ca, _ = ctx1.select_refine(ea, MAX_LEVEL, case.layers_ref, kind=case.kind)
        cb, _ = ctxh.select_refine(eb, MAX_LEVEL, case.layers_ref, kind=case.kind)
        pa, _ = ctx1.select_unrefine(ea, case.layers_unref, kind=case.kind)
        pb, _ = ctxh.select_unrefine(eb, case.layers_unref, kind=case.kind)
        same = bool(np.array_equal(ea, eb) and np.array_equal(ca, cb) and np.array_equal(pa, pb))
        # the full hand-over, call by call, twice (the second pass runs with every staging buffer in place)
        st_, sz_, ty_, nb_ = mesh1.patch_arrays()
        full_parts = []
        for _rep in range(2):
            torch.cuda.synchronize()
            tt = [time.perf_counter()]
            ctxh.set_mesh_raw(mesh1.n, mesh1.f, mesh1.b, mesh1.p, mesh1.owner, mesh1.neighbour, st_, sz_, ty_, nb_)
            tt.append(time.perf_counter())
            ctxh.set_points(mesh1.pc_off, mesh1.pc_cells, mesh1.bnd_point)
            tt.append(time.perf_counter())
            ctxh.set_levels(mesh1.cell_level, mesh1.point_level, mesh1.visible, mesh1.split_parent)
            tt.append(time.perf_counter())
            full_parts.append([1e3 * (b_ - a_) for a_, b_ in zip(tt[:-1], tt[1:])])
        t_full = sum(full_parts[-1]) / 1e3
        # ... and the incremental one again on the warm context: back to M0, then the update
        ctxh.set_mesh(mesh0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctxh.update_mesh(mesh1, delta)
        t_inc = min(t_inc, time.perf_counter() - t0)
        ctxh.close()
        full_bytes = int(sum(getattr(mesh1, k).nbytes for k in ("owner", "neighbour", "pc_off", "pc_cells", "bnd_point", "cell_level",
                                                                "point_level", "visible", "split_parent") if getattr(mesh1, k, None) is not None))
        handover = {"cut": "M0 -> M1 of this step", "full_ms": t_full * 1e3, "incremental_ms": t_inc * 1e3, "full_bytes": full_bytes,
                    "incremental_bytes": int(delta.nbytes() + 4 * mesh1.b), "dirty_cells": int(len(delta.dirty_cells)),
                    "dirty_points": int(len(delta.dirty_points)), "cells_new": int(mesh1.n), "points_new": int(mesh1.p),
                    "decisions_equal_full_hand_over": same, "delta_host_s": t_delta,
                    "full_ms_by_call": {"passes": full_parts, "calls": ["amr_mesh_set", "amr_mesh_set_points", "amr_levels_set"]},
                    "note": "amr_mesh_set + amr_mesh_set_points + amr_levels_set of M1 against amr_mesh_update + amr_mesh_update_points from the "
                            "resident M0 (wall clock of the synchronous C-ABI calls, pageable host arrays); estimate / cellsToRefine / split "
                            "points on the carried adjacency compared with the fully handed-over context"}
        log(f"hand-over of the cut: full {t_full * 1e3:.1f} ms, incremental {t_inc * 1e3:.1f} ms ({handover['incremental_bytes'] / 1e6:.0f} MB instead of "
            f"{full_bytes / 1e6:.0f} MB), decisions equal: {same}; deriving the delta on the host took {t_delta:.1f} s")
        del delta, xe, ea, eb, mesh0
    cell_map = torch.from_numpy(cell_map_h).to(dev)
    rev_map = torch.from_numpy(rev_h).to(dev)
    assert int(cell_map.numel()) == n1
    new1 = torch.empty(n1, dtype=torch.float64, device=dev)
    new3 = torch.empty((n1, 3), dtype=torch.float64, device=dev)
    err0 = torch.empty(n, dtype=torch.float64, device=dev)
    err1 = torch.empty(n1, dtype=torch.float64, device=dev)
    rc0 = torch.empty(n, dtype=torch.uint8, device=dev)
    pts_dev = torch.empty(max(p1, 1), dtype=torch.int32, device=dev)
    # first pass of the unrefinement branch -> the merge map M1 -> M2 (local: the 8 cells around a split point collapse)
    ctx.estimate("delta", x, thresholds=THRS, out=err0)
    ctx.select_refine(err0, MAX_LEVEL, case.layers_ref, kind=case.kind, cells_out=cells_dev, refine_cell_out=rc0)
    ctx.map_field(cell_map, err0, n_old=n, n_new=n1, ncomp=1, out=err1)
    ctx1.remap_refine_cell(cell_map, rev_map, refine_cell=rc0)
    pts_first, _ = ctx1.select_unrefine(err1, case.layers_unref, kind=case.kind, elems_out=np.empty(max(p1, 1), dtype=np.int32))
    pts_first = pts_first.copy()
    cm2_h, _rev2 = unrefine_maps(mesh1.pc_off, mesh1.pc_cells, pts_first, n1)
    n2 = int(len(cm2_h))
    gpts = mesh1.point_global[pts_first] if world > 1 else pts_first
    del mesh1, _rev2
    cell_map2 = torch.from_numpy(cm2_h).to(dev)
    x2 = torch.empty(n2, dtype=torch.float64, device=dev)
    u2 = torch.empty((n2, 3), dtype=torch.float64, device=dev)
    err2 = torch.empty(n2, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup
    if world > 1 and rank == 0:
        host_binding = bind_host_threads(local_rank, world)      # the host topology engine is done: rank 0 takes its share too
    log(f"setup done: local cells {n} -> {n1} -> {n2}, refine {len(cells_first)}, split points {len(pts_first)}")

    # one event interval per stage
    stage_names = ["E", "M+B", "F", "U", "F2"]
    NEV = len(stage_names) + 1
    info = {}

    # the C-ABI calls of one step, arguments converted once (what a host code in a time loop does)
    import ctypes as C
    thr_arr = np.asarray(THRS, dtype=np.float64)
    no_ids = np.zeros(1, dtype=np.int32)
    cnt_r, sw_r, cnt_u, sw_u = C.c_int64(0), C.c_int(0), C.c_int64(0), C.c_int(0)

    def field_args(olds, news, ncs):
        k = len(olds)
        return ((C.c_int * k)(*ncs), (C.c_void_p * k)(*[a.data_ptr() for a in olds]), (C.c_void_p * k)(*[a.data_ptr() for a in news]))

    fa, fa2 = field_args([x, fld3, err0], [new1, new3, err1], [1, 3, 1]), field_args([new1, new3, err1], [x2, u2, err2], [1, 3, 1])
    call_E = ctx.prepare("amr_estimate", capi.EST["delta"], x, None, capi.SMALL, 0.0, 1, thr_arr, err0)
    call_R = ctx.prepare("amr_select_refine", err0, capi.SQRT_SMALL, capi.GREAT, None, MAX_LEVEL, case.layers_ref, no_ids, 0, capi.LABEL_MAX,
                         case.kind, rc0, cells_dev, C.byref(cnt_r), C.byref(sw_r))
    call_F = ctx1.prepare("amr_map_fields_and_marking", n, n1, cell_map, 3, fa[0], fa[1], fa[2], rev_map, rc0, None)
    call_U = ctx1.prepare("amr_select_unrefine", err1, -capi.SQRT_SMALL, None, case.layers_unref, no_ids, 0, case.kind, pts_dev,
                          C.byref(cnt_u), C.byref(sw_u))
    call_F2 = ctx1.prepare("amr_map_fields_and_marking", n1, n2, cell_map2, 3, fa2[0], fa2[1], fa2[2], None, None, None)

    def step(evs=None):
        if evs is None:
            call_E(); call_R(); call_F(); call_U(); call_F2()
        else:
            evs[0].record(); call_E()
            evs[1].record(); call_R()
            evs[2].record(); call_F()
            evs[3].record(); call_U()
            evs[4].record(); call_F2()
            evs[5].record()
        info.update(n_refine=cnt_r.value, sweeps_refine=sw_r.value, n_unrefine=cnt_u.value, sweeps_unrefine=sw_u.value)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Everything slow on the host (NVML start-up of the clock sampler: ~1 s, event creation) happens BEFORE the warm-up steps,
    # so that the timed steps follow the warm-up directly: an idle second between them lets the clocks drop and skews the
    # ranks by a millisecond, which the first timed step then pays (profiles/r2_n8_timeline.txt: 2.3 ms instead of 0.95 ms,
    # a tenth of the whole timed region at 8 GPUs).
    sampler = ClockSampler(physical_device_index(local_rank))
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(NEV)] for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    nvtx = os.environ.get("BENCH_NVTX") == "1"            # profiling aid: `ncu --nvtx --nvtx-include "amr_step/"` sees the timed steps only
    sampler.start()
    for _ in range(args.warmup):
        step()
    sampler.reset()                     # clocks / throttle reasons of the timed region (from its opening barrier on) only
    barrier()
    l0 = ctx.launches + ctx1.launches
    t_start.record()
    for k in range(args.steps):
        if nvtx:
            torch.cuda.nvtx.range_push("amr_step")
        step(evs[k])
        if nvtx:
            torch.cuda.nvtx.range_pop()
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launches + ctx1.launches - l0
    total_ms = t_start.elapsed_time(t_end)
    stage_ms = [float(np.mean([evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(args.steps)])) for i in range(NEV - 1)]
    ms_per_step = total_ms / args.steps

    # ---- the decisions of the timed loop: same lists as the first pass, checksums over GLOBAL labels
    ok_lists = (info["n_refine"] == len(cells_first) and info["n_unrefine"] == len(pts_first) and
                bool(torch.equal(cells_dev[:info["n_refine"]].cpu(), torch.from_numpy(cells_first))) and
                bool(torch.equal(pts_dev[:info["n_unrefine"]].cpu(), torch.from_numpy(pts_first))))
    sums = np.array([checksum(gcells), checksum(gpts), len(gcells), len(gpts), n, n1, n2, 0 if ok_lists else 1], dtype=np.int64)

    # ---- e2e through the C-ABI with host (pinned) buffers
    e2e = None
    if not args.no_e2e:
        xh = torch.from_numpy(x_host).pin_memory()
        f3h = torch.stack([xh, 2 * xh, -xh], dim=1).contiguous().pin_memory()
        cmh = cell_map.cpu().pin_memory()
        cm2h = cell_map2.cpu().pin_memory()
        n1h = torch.empty(n1, dtype=torch.float64).pin_memory()
        n3h = torch.empty((n1, 3), dtype=torch.float64).pin_memory()
        x2h = torch.empty(n2, dtype=torch.float64).pin_memory()
        u2h = torch.empty((n2, 3), dtype=torch.float64).pin_memory()
        cells_h2 = np.empty(max(n, 1), dtype=np.int32)
        pts_h = np.empty(max(p1, 1), dtype=np.int32)

        def step_host():
            # Host buffers for everything the solver owns (the field in, the registered fields in and their mapped versions
            # out -- after the refinement AND after the unrefinement, as fvMesh::mapFields delivers them -- and the decision
            # lists out); the path's own intermediates (error, refineCell) stay on the device between the calls, as
            # INTEGRATION.md prescribes for the drop-in.
            ctx.estimate("delta", xh.numpy(), thresholds=THRS, out=err0)
            cl, _ = ctx.select_refine(err0, MAX_LEVEL, case.layers_ref, kind=case.kind, cells_out=cells_h2, refine_cell_out=rc0)
            ctx.map_fields(cmh.numpy(), [xh.numpy(), f3h.numpy()], n, n1, [1, 3], [n1h.numpy(), n3h.numpy()])
            ctx.map_field(cell_map, err0, n_old=n, n_new=n1, ncomp=1, out=err1)
            ctx1.remap_refine_cell(cell_map, rev_map, refine_cell=rc0)
            pl, _ = ctx1.select_unrefine(err1, case.layers_unref, kind=case.kind, elems_out=pts_h)
            ctx1.map_fields(cm2h.numpy(), [n1h.numpy(), n3h.numpy()], n1, n2, [1, 3], [x2h.numpy(), u2h.numpy()])
            ctx1.map_field(cell_map2, err1, n_old=n1, n_new=n2, ncomp=1, out=err2)
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
        h2d = 8 * n + (4 * n1 + 32 * n) + (4 * n2 + 32 * n1)            # x; cellMap + fields; merge map + fields
        d2h = 4 * ncl + 32 * n1 + 4 * npl + 32 * n2                     # cellsToRefine; mapped fields; split points; merged fields
        e2e = (e2e_ms, h2d, d2h)
        del xh, f3h, cmh, cm2h, n1h, n3h, x2h, u2h

    # ---- reduced-size instance of the same case against the CPU oracle (every rank checks its own sub-domain)
    parity = None
    if not args.no_parity:
        parity = parity_check(args, case, world, rank, local_rank, procs, ctx, ctx1, capi, M)

    # ---- max / sums over ranks
    if world > 1:
        t = torch.tensor([ms_per_step] + stage_ms + [e2e[0] if e2e else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t[0])
        stage_ms = [float(v) for v in t[1:NEV]]
        if e2e:
            e2e = (float(t[NEV]), e2e[1], e2e[2])
        allsums = [None] * world if rank == 0 else None
        dist.gather_object((sums, launches, parity), allsums, dst=0, group=gl)
    else:
        allsums = [(sums, launches, parity)]

    if rank == 0:
        peak, peak_src = measured_peak()
        S = np.stack([a[0] for a in allsums])
        mask61 = (1 << 61) - 1
        n_total = int(S[:, 4].sum())
        n1_total, n2_total = int(S[:, 5].sum()), int(S[:, 6].sum())
        lists_ok = bool((S[:, 7] == 0).all())
        parities = [a[2] for a in allsums]
        value = n_total / (ms_per_step / 1e3) / 1e6
        ab = algorithmic_bytes((n, f), (n1, f1, p1, nnz_pc1), n2, info["sweeps_refine"], info["sweeps_unrefine"], case,
                               (edges0, edges1, nnz_f1))
        # measured DRAM bytes per stage of THIS command line (ncu --set full capture committed under profiles/), when the
        # capture matches the configuration; the roofline fractions of the stages and of the step come from these
        traffic = None
        tpath = ROOT / "profiles" / "r2_traffic.json"
        if tpath.exists() and world == 1:
            try:
                tj = json.loads(tpath.read_text())
                if tj.get("workload") == case.workload and int(tj.get("base", -1)) == case.base and \
                        float(tj.get("shift_cells", -1)) == case.shift_cells:
                    traffic = tj
            except Exception:
                traffic = None
        stages = {}
        for nm, ms in zip(stage_names, stage_ms):
            st = {"ms": ms, "reference_algorithm_bytes": int(ab[nm])}
            if traffic and nm in traffic.get("stage_dram_bytes", {}):
                by = float(traffic["stage_dram_bytes"][nm])
                st["dram_bytes"] = by
                st["GBps"] = (by / 1e9) / (ms / 1e3) if ms > 0 else None
                st["frac"] = st["GBps"] / peak if st["GBps"] else None
            stages[nm] = st
        # Dominant kernel: stage F is ONE launch (k_map_fields_fused: scalar + vector + error field + refineCell through one
        # read of cellMap), so its CUDA-event interval over the timed region is the launch duration.  Algorithmic bytes:
        # cellMap once, 5 components old + new, reverseCellMap + old and new flag.
        kname = "k_map_fields_fused"
        kbytes = 4 * n1 + 5 * (8 * n + 8 * n1) + (4 * n + n + n1)
        k_ms = stage_ms[2]
        roof = {"bound": "hbm", "kernel": kname, "achieved": (kbytes / 1e9) / (k_ms / 1e3), "peak": peak, "unit": "GB/s",
                "frac": (kbytes / 1e9) / (k_ms / 1e3) / peak,
                "traffic": next((v for k_, v in (traffic or {}).get("kernel_dram_bytes_per_launch", {}).items() if k_.startswith(kname)), None),
                "algorithmic_bytes": kbytes, "formula": "4n' + 5(8n + 8n') + (4n + n + n')", "ms": k_ms, "peak_source": peak_src,
                "algorithmic_step_bytes": int(sum(ab.values())),
                "algorithmic_step_note": "bytes of the REFERENCE's dense loops (SURVEY 8d), not what the kernels move"}
        if traffic and "step_dram_bytes" in traffic:
            sb = float(traffic["step_dram_bytes"])
            roof["step_dram_bytes"] = sb
            roof["step_GBps"] = (sb / 1e9) / (ms_per_step / 1e3)
            roof["step_frac"] = roof["step_GBps"] / peak
        out = {
            "metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": case.describe(), "cells": n_total, "decomposition": f"simple {procs} of the ONE box",
                       "l2_policy": "inputs larger than L2 (per-rank arrays >= 60 MB at 8 ranks; 0.5 GB at 1)",
                       "cells_per_gpu": [int(v) for v in S[:, 4]], "n_refine": int(S[:, 2].sum()), "n_unrefine": int(S[:, 3].sum()),
                       "refined_mesh_cells": n1_total, "merged_mesh_cells": n2_total,
                       "sweeps_refine": info["sweeps_refine"], "sweeps_unrefine": info["sweeps_unrefine"],
                       "build": build_info, "setup_s": setup_s, "host_threads": host_binding},
            "decisions": {"cells_to_refine_checksum": int(sum(int(v) for v in S[:, 0]) & mask61),
                          "split_points_checksum": int(sum(int(v) for v in S[:, 1]) & mask61),
                          "timed_lists_equal_first_pass": lists_ok,
                          "cells_to_refine_equal_serial_run": decomposition_independent,
                          "note": "checksums over GLOBAL labels; cellsToRefine is decomposition independent, the split points are "
                                  "not (points on processor patches cannot be unrefined, as in the reference)"},
            "parity": parities[0] if world == 1 else {"per_rank": parities, "ok": bool(all(p and p.get("ok") for p in parities))},
            "stages": stages, "roofline": roof, "gpu_launches": int(sum(a[1] for a in allsums)), "clocks": clocks,
        }
        if e2e:
            out["e2e"] = {"value": n_total / (e2e[0] / 1e3) / 1e6, "unit": "Mcells/s", "ms_per_step": e2e[0],
                          "h2d_bytes_per_step": int(e2e[1]), "d2h_bytes_per_step": int(e2e[2]),
                          "note": "per rank bytes of rank 0; host buffers for field, registered fields (both maps) and lists"}
            if handover:
                for k, nm in (("incremental_ms", "e2e_with_topology_handover"), ("full_ms", "e2e_with_full_mesh_upload")):
                    ms = e2e[0] + handover[k]
                    out[nm] = {"value": n_total / (ms / 1e3) / 1e6, "unit": "Mcells/s", "ms_per_step": ms,
                               "note": f"e2e step + the hand-over of the step's cut ({k}); the merge of the same step is not included"}
        if handover:
            out["handover"] = handover
        if not args.no_cpu and world == 1:
            cb = args.cpu_base or max(8, (case.base // 4) // 2 * 2)
            small = Case(case.workload, cb, case.shift_cells)
            Sc = oracle_prepare(small, 1, 1)
            oracle_step(Sc)
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                oracle_step(Sc)
            dt = (time.perf_counter() - t0) / reps
            out["cpu_baseline"] = {"value": Sc["n0"] / dt / 1e6, "unit": "Mcells/s", "cores": 1, "kind": "port",
                                   "sample": f"same case at base {cb}^3 ({Sc['n0']} cells), {reps} steps, 1 thread"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
    ctx.close()
    ctx1.close()
    if world > 1:
        dist.destroy_process_group()


def parity_check(args, case, world, rank, local_rank, procs, ctx, ctx1, capi, M):
    """A reduced-size instance of the same case (same build-up, same moved front, same decomposition), the CUDA path
    against the CPU oracle: error field, refineCell, cellsToRefine, mapped fields, remapped refineCell, split points,
    merged fields.  Every rank builds the small global mesh itself (deterministic), runs the oracle world of ALL ranks
    and compares its own sub-domain.  Collective."""
    from oracle import oracle as O
    workload = case.workload
    small = Case(workload, args.parity_base, args.shift_cells)
    lcap = MAX_LEVEL if small.graded else 1
    oc = M.Octree(small.base, small.base, small.base, (1.0, 1.0, 1.0), lcap)
    ser = GpuPath(local_rank, small)
    build_initial(oc, ser, small)
    ser.close()
    g0 = oc.build(points=small.poly, faces=small.poly)

    def split(g):
        if world > 1:
            return M.decompose(g, M.simple_ranks(g, (1, 1, 1), procs), world)
        return [g]

    loc0 = split(g0)
    w = O.World(loc0)
    xs = [np.ascontiguousarray(small.field(l.cc)) for l in loc0]
    e = w.estimate("delta", xs)
    w.normalize_world(e, *THRS)
    rc, rs, _ = oracle_select_refine(w, small, e)
    m0 = loc0[rank]
    ctx.set_mesh(m0)
    fails = []
    eg = ctx.estimate("delta", xs[rank], thresholds=THRS)
    if not np.array_equal(e[rank], eg):
        fails.append("error field")
    rc_g = np.zeros(m0.n, np.uint8)
    cells, _ = ctx.select_refine(eg, MAX_LEVEL, small.layers_ref, kind=small.kind, refine_cell_out=rc_g)
    if not np.array_equal(rc[rank], rc_g):
        fails.append("refineCell")
    if not np.array_equal(np.flatnonzero(rs[rank]), cells):
        fails.append("cellsToRefine")
    if world > 1:
        gc = np.sort(np.concatenate([l.cell_global[np.flatnonzero(s_)] for l, s_ in zip(loc0, rs)])).astype(np.int32)
    else:
        gc = np.flatnonzero(rs[0]).astype(np.int32)
    cm_g = oc.refine(gc)
    g1 = oc.build(points=True, faces=small.poly)
    loc1 = split(g1)
    w1 = O.World(loc1)
    cms, revs = [], []
    for l0, l1 in zip(loc0, loc1):
        cm, rv = local_maps(g0.n, g1.n, cm_g, l0.cell_global if world > 1 else None, l1.cell_global if world > 1 else None)
        cms.append(cm); revs.append(rv)
    m1 = loc1[rank]
    ctx1.set_mesh(m1)
    U = np.stack([xs[rank], 2 * xs[rank], -xs[rank]], axis=1).copy()
    new1 = ctx.map_field(cms[rank], xs[rank], ncomp=1)
    new3 = ctx.map_field(cms[rank], U, ncomp=3)
    err1 = ctx.map_field(cms[rank], eg, ncomp=1)
    if not (np.array_equal(new1, O.map_field(cms[rank], xs[rank], 1)) and np.array_equal(new3, O.map_field(cms[rank], U, 3))):
        fails.append("mapped fields (refinement)")
    err1s = [O.map_field(cm, ee, 1) for cm, ee in zip(cms, e)]
    rc1s = [O.remap_refine_cell(cm, rv, r) for cm, rv, r in zip(cms, revs, rc)]
    rc1_g = np.zeros(m1.n, np.uint8)
    ctx1.remap_refine_cell(cms[rank], revs[rank], refine_cell=rc_g, out=rc1_g)
    if not np.array_equal(rc1s[rank], rc1_g):
        fails.append("refineCell through the map")
    up, _, _ = oracle_select_unrefine(w1, small, err1s, [r.copy() for r in rc1s])
    pts, _ = ctx1.select_unrefine(err1, small.layers_unref, kind=small.kind, refine_cell=rc1_g.copy())
    if not np.array_equal(np.flatnonzero(up[rank]), pts):
        fails.append("split points")
    cm2, _ = unrefine_maps(m1.pc_off, m1.pc_cells, pts, m1.n)
    x2 = ctx1.map_field(cm2, new1, ncomp=1)
    if not np.array_equal(x2, O.map_field(cm2, new1, 1)):
        fails.append("mapped fields (unrefinement)")
    n_ref, n_unref = int(len(gc)), int(sum(int(u.sum()) for u in up))
    if n_ref == 0 or (small.graded and n_unref == 0):     # (S1 is a uniform level-0 box: nothing can be unrefined there)
        fails.append("trivial case")
    return {"ok": not fails, "failed": fails, "case": f"same case at base {small.base}^3: {g0.n} cells, {n_ref} cells refined, "
            f"{n_unref} split points merged, {world} rank(s)", "against": "oracle/amr_oracle.c (CPU restatement), bit for bit"}


if __name__ == "__main__":
    main()
