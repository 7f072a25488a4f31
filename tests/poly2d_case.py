"""The reference's polyhedral 2-D fixture (tests/testCases/poly2D: 670 prismatic polyhedra with 4-9 sided end faces,
2524 points, 5124 edges) as a test case for the polyRefiner kernels on genuinely polyhedral addressing.

The repo's host topology engine cuts octrees only, so the fixture cannot be refined for real.  The polyRefiner's
decision kernels, however, read nothing but (cellLevel, pointLevel, owner/neighbour, edgeCells, pointCells, faces):
a "virtual refinement" -- cellLevel += selected after every selection, the consistent refinement guaranteeing that
the new levels are 2:1 balanced over faces and edges again -- produces graded level fields on the unchanged
addressing, and pointLevel = max over pointCells(cellLevel) marks the interior points of equally refined regions as
split points.  That drives every polyRefiner stage (marking over faces and points, face + edge sweeps, split-point
rules, unrefinement sweeps) with non-trivial data."""
import numpy as np

from tests.fixtures import load_reference_mesh


W = 0.001
BUILD_R0 = (0.030, 0.031, 0.032)     # the front while the levels are built up
STEP_R0 = 0.018                        # the front of the step: moved inwards, the old band can unrefine


def front(cc, r0, w=W):
    d = np.sqrt((cc[:, 0] - 0.05) ** 2 + (cc[:, 1] - 0.05) ** 2)
    return 1.0 + 0.5 * (1.0 + np.tanh((d - r0) / w))


def load():
    m = load_reference_mesh("poly2D", edges=True)
    assert (m.n, m.f, m.b, m.p, m.n_edges) == (670, 1771, 1500, 2524, 5124)
    return m


def point_levels(m):
    pts = np.repeat(np.arange(m.p), np.diff(m.pc_off))
    pl = np.zeros(m.p, np.int32)
    np.maximum.at(pl, pts, m.cell_level[m.pc_cells])
    return pl
