"""Table-driven marching cubes on the host -- TEST INFRASTRUCTURE (never imported by the product).

The reference extracts its meshes with `skimage.measure.marching_cubes(volume, level=0.0, spacing=...)`
(wild_completion/utils.py:576-578; scikit-image is third-party, unpinned in README.md:47, and not installed in this image, so
mesh-level parity with skimage itself is unpinned -- SURVEY.md 8c).  This module restates the published algorithm
(Lorensen & Cline 1987: one sign bit per cube corner -> 256 cases, vertices by linear interpolation on the cube edges) so that
the device extractor (csrc/mesher.cu, marching tetrahedra) can be held against a marching-CUBES surface of the same grid at
Chamfer level, and so that the open3d/skimage stand-ins of the host-script tests have a `measure.marching_cubes`.

The 256-entry case table is GENERATED, not typed in: on every cube face the crossed edges are joined by segments (a face with
four crossings is resolved by cutting off its two inside corners -- a rule that depends on the face's signs only, so the two
cubes sharing a face agree and the surface is watertight, which the classic 15-case table does not guarantee); the segments
close into loops, each loop is fan-triangulated and oriented towards increasing values.  skimage's default (Lewiner et al.
2003) differs from it only in how ambiguous faces / interiors are resolved, i.e. on isolated cells.
"""
from __future__ import annotations

import numpy as np

# corner c of the unit cube sits at (c & 1, (c >> 1) & 1, (c >> 2) & 1) in (i, j, k) index order
CORNERS = np.array([[(c >> 0) & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)], np.int64)
# 12 edges: (corner a, corner b) with b = a + one axis bit, grouped by axis
EDGES = [(a, a | (1 << ax)) for ax in range(3) for a in range(8) if not a & (1 << ax)]
EDGE_AXIS = [ax for ax in range(3) for a in range(8) if not a & (1 << ax)]
EDGE_ID = {e: i for i, e in enumerate(EDGES)}
# 6 faces as cyclic corner quadruples
FACES = []
for ax in range(3):
    u, v = [(1 << a) for a in range(3) if a != ax]
    for side in (0, 1 << ax):
        FACES.append([side, side | u, side | u | v, side | v])


def _edge(a, b):
    return EDGE_ID[(min(a, b), max(a, b))]


def _build_case(config: int):
    inside = [(config >> c) & 1 for c in range(8)]
    links = {}
    for f in FACES:
        cr = [_edge(f[i], f[(i + 1) % 4]) for i in range(4) if inside[f[i]] != inside[f[(i + 1) % 4]]]
        if len(cr) == 2:
            segs = [(cr[0], cr[1])]
        elif len(cr) == 4:           # alternating signs: cut off each inside corner with its two adjacent edges
            segs = []
            for i in range(4):
                if inside[f[i]]:
                    segs.append((_edge(f[i - 1], f[i]), _edge(f[i], f[(i + 1) % 4])))
        else:
            segs = []
        for a, b in segs:
            links.setdefault(a, []).append(b)
            links.setdefault(b, []).append(a)
    mid = {e: (CORNERS[EDGES[e][0]] + CORNERS[EDGES[e][1]]) / 2.0 for e in links}
    tris, seen = [], set()
    for start in sorted(links):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start          # every crossed edge lies on two faces -> exactly two links -> closed loops
        seen.add(start)
        while True:
            a, b = links[cur]
            nxt = a if a != prev else b
            if nxt == start:
                break
            loop.append(nxt)
            seen.add(nxt)
            prev, cur = cur, nxt
        # orientation: normal away from the inside corners (towards increasing values)
        pts = np.array([mid[e] for e in loop])
        n = np.zeros(3)
        for i in range(len(loop)):
            n += np.cross(pts[i], pts[(i + 1) % len(loop)])
        s = 0.0
        for e, p in zip(loop, pts):
            a, b = EDGES[e]
            s += float(n @ (p - CORNERS[a if inside[a] else b]))
        if s < 0:
            loop = loop[::-1]
        tris += [(loop[0], loop[i], loop[i + 1]) for i in range(1, len(loop) - 1)]
    return tris


CASES = [_build_case(c) for c in range(256)]
MAX_TRIS = max(len(t) for t in CASES)


def marching_cubes(volume: np.ndarray, level: float = 0.0, spacing=(1.0, 1.0, 1.0)):
    """-> (verts (V,3) float64 in index space times `spacing`, faces (F,3) int64).  Vertices are welded (one per crossed grid
    edge).  A corner counts as inside when its value is < level."""
    vol = np.asarray(volume, np.float64)
    nx, ny, nz = vol.shape
    inside = vol < level
    cfg = np.zeros((nx - 1, ny - 1, nz - 1), np.int64)
    for c in range(8):
        i, j, k = CORNERS[c]
        cfg |= inside[i:nx - 1 + i, j:ny - 1 + j, k:nz - 1 + k].astype(np.int64) << c
    lin = np.arange(nx * ny * nz, dtype=np.int64).reshape(nx, ny, nz)
    base = lin[:-1, :-1, :-1]
    strides = np.array([ny * nz, nz, 1], np.int64)
    corner_off = CORNERS @ strides
    keys = []
    for case in np.unique(cfg):
        tris = CASES[int(case)]
        if not tris:
            continue
        cells = base[cfg == case]
        t = np.asarray(tris, np.int64)                                  # (T,3) cube-edge ids
        lower = np.array([corner_off[EDGES[e][0]] for e in range(12)], np.int64)
        axis = np.array(EDGE_AXIS, np.int64)
        k = (cells[:, None, None] + lower[t][None]) * 3 + axis[t][None]  # global edge key = lower grid point * 3 + axis
        keys.append(k.reshape(-1, 3))
    if not keys:
        return np.zeros((0, 3)), np.zeros((0, 3), np.int64)
    keys = np.concatenate(keys, 0)
    uniq, inv = np.unique(keys.reshape(-1), return_inverse=True)
    faces = inv.reshape(-1, 3)
    p0, ax = uniq // 3, uniq % 3
    p1 = p0 + strides[ax]
    v0, v1 = vol.reshape(-1)[p0], vol.reshape(-1)[p1]
    t = (level - v0) / (v1 - v0)
    ijk = np.stack(np.unravel_index(p0, vol.shape), 1).astype(np.float64)
    ijk[np.arange(len(ax)), ax] += t
    return ijk * np.asarray(spacing, np.float64), faces
