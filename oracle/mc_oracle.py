"""TEST INFRASTRUCTURE -- numpy restatement of the reference's marching cubes
(third_parties/coslam/external/NumpyMarchingCubes/marching_cubes/src/marching_cubes.cpp).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline may import this.  Pinned to the reference's own extractor compiled where
it lies (oracle/_ref/libmc_ref.so, oracle/build_ref.py) and to tests/golden/mc_small.npz produced by it
(oracle/make_golden_mc.py).  The triangle table is read from naruto_b200/csrc/mc_tables.inc (data, see
oracle/make_mc_tables.py)."""
import os
import re

import numpy as np

f32 = np.float32
_INC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'naruto_b200', 'csrc', 'mc_tables.inc')
# cube-case bit -> corner offset (marching_cubes.cpp:192-199); edge -> its two corner bits (:234-245)
BIT_CORNER = [(0, 1, 0), (1, 1, 0), (1, 0, 0), (0, 0, 0), (0, 1, 1), (1, 1, 1), (1, 0, 1), (0, 0, 1)]
EDGE_ENDS = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def tri_table():
    words = [int(w, 16) for w in re.findall(r'0x([0-9A-Fa-f]{16})ull', open(_INC).read())]
    assert len(words) == 256
    rows = []
    for w in words:
        row = []
        for t in range(16):
            e = (w >> (4 * t)) & 0xF
            if e == 0xF:
                break
            row.append(e)
        rows.append(row)
    return rows


def corner_grid(vol, truncation):
    """trilerp (:96-118) at every dual-grid corner: mean of the 2x2x2 voxels around it, NaN where any of them is out of
    bounds, -inf or |d| >= truncation.  Shape (nx+1, ny+1, nz+1)."""
    nx, ny, nz = vol.shape
    c = np.full((nx + 1, ny + 1, nz + 1), np.nan, dtype=f32)
    if min(nx, ny, nz) < 2:
        return c
    dist = np.zeros((nx - 1, ny - 1, nz - 1), dtype=f32)
    ok = np.ones(dist.shape, dtype=bool)
    for ox, oy, oz in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)]:
        d = vol[ox:nx - 1 + ox, oy:ny - 1 + oy, oz:nz - 1 + oz]
        with np.errstate(invalid='ignore'):
            ok &= (d != -np.inf) & (np.abs(d) < f32(truncation))
        dist = (dist + f32(0.125) * d).astype(f32)
    c[1:nx, 1:ny, 1:nz] = np.where(ok, dist, f32(np.nan))
    return c


def triangle_soup(volume, isovalue, truncation, thresh=10.0):
    """run_marching_cubes_internal (:403-416): [T,3,3] fp32 vertices in scan order."""
    vol = np.asarray(volume, dtype=np.float64).astype(f32)             # `float d = accessor(...)` narrows the double
    iso, th = f32(isovalue), f32(thresh)
    nx, ny, nz = vol.shape
    cg = corner_grid(vol, truncation)
    table = tri_table()
    d = np.stack([cg[cx:cx + nx, cy:cy + ny, cz:cz + nz] for cx, cy, cz in BIT_CORNER])          # [8, nx, ny, nz]
    valid = ~np.isnan(d).any(0)
    with np.errstate(invalid='ignore'):
        cube = sum(((d[b] < iso).astype(np.int32) << b) for b in range(8))
        emit = valid.copy()
        for a in range(8):
            for b in range(8):
                neg = (d[a] * d[b]).astype(f32) < 0
                big = np.where(neg, (np.abs(d[a]) + np.abs(d[b])).astype(f32) > th, np.abs((d[a] - d[b]).astype(f32)) > th)
                emit &= ~big
            emit &= ~(np.abs(d[a]) > th)
    mask = np.zeros(cube.shape, dtype=np.int32)
    for e, (a, b) in enumerate(EDGE_ENDS):
        mask |= ((((cube >> a) ^ (cube >> b)) & 1) << e)
    emit &= (mask != 0) & (mask != 255)
    tris = []
    for i, j, k in np.argwhere(emit):                                  # argwhere is in scan order (i, j, k)
        dd = d[:, i, j, k]
        pos = np.array([i, j, k], dtype=f32)
        verts = {}
        for e in set(table[cube[i, j, k]]):
            a, b = EDGE_ENDS[e]
            p1 = pos + np.array([0.5 if o else -0.5 for o in BIT_CORNER[a]], dtype=f32)
            p2 = pos + np.array([0.5 if o else -0.5 for o in BIT_CORNER[b]], dtype=f32)
            d1, d2 = dd[a], dd[b]
            if abs(f32(iso - d1)) < f32(0.00001):                      # vertexInterp (:122-141)
                v = p1
            elif abs(f32(iso - d2)) < f32(0.00001):
                v = p2
            elif abs(f32(d1 - d2)) < f32(0.00001):
                v = p1
            else:
                mu = f32(f32(iso - d1) / f32(d2 - d1))
                v = (p1 + (mu * (p2 - p1).astype(f32)).astype(f32)).astype(f32)
            verts[e] = v
        row = table[cube[i, j, k]]
        for t in range(0, len(row), 3):
            tris.append([verts[row[t]], verts[row[t + 1]], verts[row[t + 2]]])
    return np.array(tris, dtype=f32).reshape(-1, 3, 3)


def _sgn(v):
    return int(0.0 < v) - int(v < 0.0)


def merge(soup):
    """merge_close_vertices(approx=True) + remove_degenerate_faces + remove_duplicate_faces (:253-416)."""
    cell = f32(0.00001)
    flat = soup.reshape(-1, 3)
    seen, lookup, new_verts = {}, [], []
    for v in flat:
        c = tuple(int(f32(f32(x / cell) + f32(0.5) * f32(_sgn(x)))) for x in v)
        found = None
        for i in (-1, 0, 1):
            for j in (-1, 0, 1):
                for k in (-1, 0, 1):
                    if found is None:
                        found = seen.get((c[0] + i, c[1] + j, c[2] + k))
        if found is None:
            seen[c] = len(new_verts)
            lookup.append(len(new_verts))
            new_verts.append(v)
        else:
            lookup.append(found)
    faces, dup = [], set()
    for f in range(0, len(lookup), 3):
        a, b, c = lookup[f:f + 3]
        if a == b or a == c or b == c:
            continue
        key = tuple(sorted((a, b, c)))
        if key in dup:
            continue
        dup.add(key)
        faces.append((a, b, c))
    return (np.array(new_verts, dtype=np.float64).reshape(-1, 3), np.array(faces, dtype=np.uint64).reshape(-1, 3))


def marching_cubes(volume, isovalue, truncation):
    """mcubes.marching_cubes (:418-462 + _mcubes.pyx:20-25): vertices float64 [V,3], triangles uint64 [F,3]."""
    return merge(triangle_soup(volume, isovalue, truncation))
