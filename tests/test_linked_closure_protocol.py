"""The cross-rank protocol of the linked closure loops (blastamr_b200/csrc/kernels_linked.cuh), replayed on the CPU with
adversarial timing: ONE halo swap per closure, afterwards a boundary cell stores its new value into the neighbours' ghost
arrays the moment it changes; a rank's coupled-face pass may read a ghost byte before or after the neighbour's store of the
same iteration (decided at random here); every iteration ends with an all-reduce of the queue sizes that also makes all
stores of the iteration visible.  Claim checked: whatever the interleaving, the loop stops with exactly the closure of the
serial algorithm on the undecomposed mesh (hexRef::consistentRefinement, maxSet = true; hexRef.C:700-783, 1402-1468), because
marks only grow, every flip is forced, and a flip that a neighbour missed sits in a queue, so another iteration follows."""
import numpy as np
import pytest

from blastamr_b200 import mesh as M
from oracle import oracle as O
from tests import amr_cases as cases
from tests.halo_host_mirror import halo_plan


class Rank:
    def __init__(self, mesh, marks):
        self.m = mesh
        n = mesh.n
        own, nei = mesh.owner[:mesh.f], mesh.neighbour
        rows = [[] for _ in range(n)]
        for o, q in zip(own.tolist(), nei.tolist()):
            rows[o].append(q)
            rows[q].append(o)
        self.rows = rows
        self.lvl = mesh.cell_level.astype(np.int64)
        self.set = marks.astype(bool).copy()
        self.lf = self.lvl + self.set                                   # level + flag, what the swap carries
        self.plan = halo_plan(mesh)                                     # [(nbr, offset into boundary faces, count)]
        self.bown = mesh.owner[mesh.f:]
        self.ghost = np.zeros(mesh.b, dtype=np.int64)                   # my ghost array (boundary-face layout), written by neighbours
        self.coupled = np.zeros(mesh.b, dtype=bool)
        for _, off, cnt in self.plan:
            self.coupled[off:off + cnt] = True
        # where a boundary cell's value goes: (neighbour rank, index in ITS ghost array) per coupled face it owns
        self.dst = {}
        self.queue = []


def link(ranks):
    """push-on-change addressing (linkPlan in amr_b200.cu): patch pairs list their faces in the same order on both sides"""
    for r, R in enumerate(ranks):
        for nbr, off, cnt in R.plan:
            their = [(o2, c2) for n2, o2, c2 in ranks[nbr].plan if n2 == r]
            assert len(their) == 1 and their[0][1] == cnt
            off2 = their[0][0]
            for k in range(cnt):
                R.dst.setdefault(int(R.bown[off + k]), []).append((nbr, off2 + k))


def run_linked(ranks, rng):
    link(ranks)
    for R in ranks:                                 # the ONE swap of the closure: state before sweep 0
        for c, targets in R.dst.items():
            for nbr, idx in targets:
                ranks[nbr].ghost[idx] = R.lf[c]
    pending = []                                    # remote stores of the running iteration not yet visible: (rank, idx, value)

    def flip(R, c):
        R.set[c] = True
        R.lf[c] = R.lvl[c] + 1
        R.queue_next.append(c)
        for nbr, idx in R.dst.get(c, []):
            if rng.random() < 0.5:
                ranks[nbr].ghost[idx] = R.lf[c]     # the neighbour may already see it in this iteration ...
            else:
                pending.append((nbr, idx, R.lf[c]))  # ... or only after the iteration's all-reduce

    iters = 0
    first = True
    while True:
        for R in ranks:
            R.queue_next = []
        work = []                                   # (rank, kind, item): queue rows and coupled faces, in ONE random order
        for r, R in enumerate(ranks):
            src = np.flatnonzero(R.set).tolist() if first else R.queue
            work += [(r, 0, c) for c in src]
            work += [(r, 1, int(i)) for i in np.flatnonzero(R.coupled)]
        rng.shuffle(work)
        for r, kind, item in work:
            R = ranks[r]
            if kind == 0:
                c = item
                if not R.set[c]:
                    continue
                for q in R.rows[c]:
                    if R.lf[c] > R.lvl[q] + 1 and not R.set[q]:
                        flip(R, q)
            else:
                o = int(R.bown[item])
                if R.ghost[item] > R.lvl[o] + 1 and not R.set[o]:
                    flip(R, o)
        for nbr, idx, v in pending:                 # the all-reduce: every store of the iteration is visible now
            ranks[nbr].ghost[idx] = max(ranks[nbr].ghost[idx], v)
        pending.clear()
        total = sum(len(R.queue_next) for R in ranks)
        for R in ranks:
            R.queue = R.queue_next
        first = False
        iters += 1
        if total == 0:
            return iters
        assert iters < 100


@pytest.mark.parametrize("nranks,procs,seed", [(2, (2, 1, 1), 1), (4, (2, 2, 1), 2), (8, (2, 2, 2), 3), (8, (2, 2, 2), 4)])
def test_linked_refinement_closure_equals_the_serial_closure(nranks, procs, seed):
    rng = np.random.default_rng(seed)
    oc = cases.graded_octree(8, 2, centre=(0.5, 0.5, 0.5), r0=0.3, band=0.1)
    g = oc.build()
    # candidates: the finest cells near a shifted shell (forces the closure to pull coarser neighbours in, across the patches)
    d = np.sqrt(((g.cc - np.array([0.56, 0.5, 0.5])) ** 2).sum(axis=1))
    marks = ((np.abs(d - 0.3) < 0.05) & (g.cell_level >= 1) & (g.cell_level < 2 + 1)).astype(np.uint8)
    ref = marks.copy()
    w = O.World([g])
    w.consistent_refinement([ref], True)
    assert ref.sum() > marks.sum()                  # the closure really adds cells
    loc = M.decompose(g, M.simple_ranks(g, (1, 1, 1), procs), nranks)
    ranks = [Rank(l, marks[l.cell_global]) for l in loc]
    iters = run_linked(ranks, rng)
    got = np.zeros(g.n, np.uint8)
    for l, R in zip(loc, ranks):
        got[l.cell_global] = R.set
    assert np.array_equal(got, ref)
    assert iters >= 2
