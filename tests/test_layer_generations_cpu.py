"""Buffer layers in place (kernels_sparse.cuh (1)): the marking is a byte per cell, layer k treats 0 < v <= gen as sources and
writes gen + 1, so ONE array is read and written by unordered threads.  Replayed here with random cell orders for the three
device forms -- push from the sources, pull into the unmarked cells, pull with the source mask frozen as bits before the walk
(k_pack_source_bits) -- against the reference's layer-by-layer growth (fvMeshRefiner::extendMarkedCellsAcrossFaces,
fvMeshRefiner.C:866-921: a marked cell marks all its face neighbours, from a copy of the marking)."""
import numpy as np

from tests import amr_cases as cases


def test_in_place_generation_layers_equal_the_copying_reference():
    m = cases.graded_octree(8, 2).build(points=False)
    own, nei = m.owner[:m.f], m.neighbour
    rows = [[] for _ in range(m.n)]
    for o, q in zip(own.tolist(), nei.tolist()):
        rows[o].append(q)
        rows[q].append(o)
    rng = np.random.default_rng(3)
    seed = rng.random(m.n) < 0.03
    layers = 3
    ref = seed.copy()
    for _ in range(layers):                                   # the reference: from a copy, layer by layer
        nxt = ref.copy()
        src = np.flatnonzero(ref)
        for c in src.tolist():
            for q in rows[c]:
                nxt[q] = True
        ref = nxt
    for form in ("push", "pull", "pull_bits"):
        cur = seed.astype(np.int64).copy()                    # candidates hold generation 1
        for gen in range(1, layers + 1):
            order = rng.permutation(m.n).tolist()
            frozen = (cur > 0) & (cur <= gen) if form == "pull_bits" else None
            for c in order:
                if form == "push":
                    if 0 < cur[c] <= gen:
                        for q in rows[c]:
                            if cur[q] == 0:
                                cur[q] = gen + 1
                else:
                    if cur[c] == 0:
                        hit = any(frozen[q] for q in rows[c]) if frozen is not None else any(0 < cur[q] <= gen for q in rows[c])
                        if hit:
                            cur[c] = gen + 1
        assert np.array_equal(cur > 0, ref), form
        assert cur.max() == layers + 1
