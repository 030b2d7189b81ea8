"""TEST INFRASTRUCTURE -- recovers the 256-case triangle table the reference's extractor uses, by running the compiled
reference (oracle/_ref/libmc_ref.so) on one synthetic cell per case and reading back which cube edges each emitted triangle
touches.  A 3x3x3 volume has exactly one valid cell (centre voxel); its eight dual-grid corner values are the means of the
eight 2x2x2 voxel blocks, so any sign pattern is reachable by a minimum-norm solve.  Used to check csrc/mc_tables.inc
(the classic Lorensen-Cline / Bourke table) against the reference's behaviour:  python -m oracle.probe_mc_tables"""
import itertools

import numpy as np

from oracle import mc_ref

# cube-index bit -> corner offset (x, y, z) in {0,1}^3, as marching_cubes.cpp:192-199 tests them
BIT_CORNER = [(0, 1, 0), (1, 1, 0), (1, 0, 0), (0, 0, 0), (0, 1, 1), (1, 1, 1), (1, 0, 1), (0, 0, 1)]
# edge -> (corner bit a, corner bit b), as marching_cubes.cpp:234-245 interpolates them
EDGE_ENDS = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def corner_matrix():
    A = np.zeros((8, 27))
    for b, (cx, cy, cz) in enumerate(BIT_CORNER):
        for dx, dy, dz in itertools.product((0, 1), repeat=3):
            A[b, ((cx + dx) * 3 + (cy + dy)) * 3 + (cz + dz)] = 0.125
    return A


def probe():
    A = corner_matrix()
    pinv = np.linalg.pinv(A)
    mids = {}
    for e, (a, b) in enumerate(EDGE_ENDS):
        pa, pb = np.array(BIT_CORNER[a], float), np.array(BIT_CORNER[b], float)
        mids[tuple(np.round(0.5 + 0.5 * (pa + pb), 3))] = e          # cell spans [0.5, 1.5]^3
    table = []
    for case in range(256):
        target = np.array([-1.0 if (case >> b) & 1 else 1.0 for b in range(8)])
        vol = (pinv @ target).reshape(3, 3, 3)
        verts, faces = mc_ref.marching_cubes(vol, 0.0, 1e3)     # truncation out of the way; corner values are +-1
        row = []
        for f in faces:
            for vi in f:
                row.append(mids[tuple(np.round(verts[int(vi)], 3))])
        table.append(row)
    return table


if __name__ == '__main__':
    t = probe()
    for case, row in enumerate(t):
        print(case, row)
