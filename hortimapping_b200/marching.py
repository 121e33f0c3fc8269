"""Iso-surface extraction for the mesher (the step AFTER the hot path).

The reference calls skimage.measure.marching_cubes (Lewiner) on the host
(wild_completion/utils.py:565-588).  When scikit-image is importable it is used the same way, so the
mesh is exactly what the reference would produce from the same SDF grid; otherwise a vectorised
marching-tetrahedra extractor (numpy) is used.  Mesh-level parity with skimage is not pinned (third-party,
unpinned version); parity is pinned at the SDF-grid level and checked at Chamfer level.
"""
from __future__ import annotations

import numpy as np

# the 6 tetrahedra of a cube around its main diagonal (corner 0 -> corner 6); corner numbering:
# c = (dx, dy, dz): 0=(0,0,0) 1=(1,0,0) 2=(1,1,0) 3=(0,1,0) 4=(0,0,1) 5=(1,0,1) 6=(1,1,1) 7=(0,1,1)
_CORNERS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
_TETS = np.array([[0, 5, 1, 6], [0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6]])
# tetra edges (pairs of local vertices) and, per inside-mask (bit i = vertex i inside), the crossed edges of
# up to 2 triangles
_EDGES = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]])
_E = {(0, 1): 0, (0, 2): 1, (0, 3): 2, (1, 2): 3, (1, 3): 4, (2, 3): 5}
_TRI = {
    0b0001: [(0, 1, 2)], 0b0010: [(0, 3, 4)], 0b0100: [(1, 3, 5)], 0b1000: [(2, 4, 5)],
    0b0011: [(1, 2, 3), (2, 3, 4)], 0b0101: [(0, 2, 3), (2, 3, 5)], 0b1001: [(0, 1, 4), (1, 4, 5)],
}
for _m in list(_TRI):
    _TRI[0b1111 ^ _m] = _TRI[_m]
_TRI_TABLE = np.full((16, 2, 3), -1, np.int64)
for _m, _ts in _TRI.items():
    for _i, _t in enumerate(_ts):
        _TRI_TABLE[_m, _i] = _t


def marching_tetrahedra(vol: np.ndarray, level: float = 0.0, spacing=(1.0, 1.0, 1.0)):
    """Vertices (V,3) float32 in index*spacing coordinates and faces (F,3) int32, wound so that the normal
    points towards increasing field values (outside of an SDF)."""
    vol = np.asarray(vol, np.float64) - level
    nx, ny, nz = vol.shape
    ii, jj, kk = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), np.arange(nz - 1), indexing="ij")
    base = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], 1)                        # (C,3)
    cidx = base[:, None, :] + _CORNERS[None, :, :]                                  # (C,8,3)
    cval = vol[cidx[..., 0], cidx[..., 1], cidx[..., 2]]                            # (C,8)
    mixed = (cval.min(1) < 0) & (cval.max(1) >= 0)
    cidx, cval = cidx[mixed], cval[mixed]
    if cidx.shape[0] == 0:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)
    tv = cval[:, _TETS]                                                             # (C,6,4)
    tp = cidx[:, _TETS, :]                                                          # (C,6,4,3)
    tv, tp = tv.reshape(-1, 4), tp.reshape(-1, 4, 3)
    inside = tv < 0
    mask = (inside * np.array([1, 2, 4, 8])).sum(1)
    keep = (mask != 0) & (mask != 15)
    tv, tp, mask, inside = tv[keep], tp[keep], mask[keep], inside[keep]
    tris = _TRI_TABLE[mask]                                                         # (T,2,3) edge ids
    valid = tris[..., 0] >= 0                                                       # (T,2)
    tsel, which = np.nonzero(valid)
    e = tris[tsel, which]                                                           # (F,3) edge ids
    a_loc, b_loc = _EDGES[e, 0], _EDGES[e, 1]                                       # (F,3)
    rows = tsel[:, None]
    pa, pb = tp[rows, a_loc], tp[rows, b_loc]                                       # (F,3,3) grid indices
    va, vb = tv[rows, a_loc], tv[rows, b_loc]
    t = va / (va - vb)
    pos = pa + (pb - pa) * t[..., None]                                             # (F,3,3)
    # orient: normal towards the positive (outside) vertices of the tetrahedron
    w_out = (~inside[tsel]).astype(np.float64)
    w_in = inside[tsel].astype(np.float64)
    c_out = (tp[tsel] * w_out[..., None]).sum(1) / w_out.sum(1, keepdims=True)
    c_in = (tp[tsel] * w_in[..., None]).sum(1) / w_in.sum(1, keepdims=True)
    n = np.cross(pos[:, 1] - pos[:, 0], pos[:, 2] - pos[:, 0])
    flip = (n * (c_out - c_in)).sum(1) < 0
    pos[flip] = pos[flip][:, [0, 2, 1]]
    pa_f, pb_f = pa.copy(), pb.copy()
    pa_f[flip], pb_f[flip] = pa[flip][:, [0, 2, 1]], pb[flip][:, [0, 2, 1]]
    # weld vertices by their (unordered) grid edge
    lin = lambda p: (p[..., 0] * ny + p[..., 1]) * nz + p[..., 2]
    ka, kb = lin(pa_f), lin(pb_f)
    key = np.minimum(ka, kb) * (nx * ny * nz) + np.maximum(ka, kb)
    uniq, first, inv = np.unique(key.ravel(), return_index=True, return_inverse=True)
    verts = pos.reshape(-1, 3)[first] * np.asarray(spacing, np.float64)
    faces = inv.reshape(-1, 3)
    good = (faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 0] != faces[:, 2])
    return verts.astype(np.float32), faces[good].astype(np.int32)


def extract_isosurface(vol: np.ndarray, level: float, spacing):
    """skimage's marching_cubes when available (exactly the reference's call, utils.py:576-578)."""
    try:
        from skimage import measure  # type: ignore
        verts, faces, _, _ = measure.marching_cubes(vol, level=level, spacing=list(spacing))
        return verts, faces
    except ImportError:
        return marching_tetrahedra(vol, level, spacing)
