"""N2 pinned at surface level: the reference meshes with skimage's marching cubes (third-party, absent here);
oracle/marching_cubes.py restates marching cubes with a generated case table.  CPU: the table itself (watertight, oriented,
analytic sphere) and the product's marching-tetrahedra extractor (hortimapping_b200/marching.py, the host twin of csrc/mesher.cu)
against it on the same grids: the two surfaces must coincide to a small fraction of a voxel."""
import numpy as np
import pytest

from oracle import marching_cubes as MC
from tests.helpers import point_to_mesh_distance


def sphere_grid(n, r=0.7):
    g = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    return np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - r


def test_generated_case_table():
    assert len(MC.CASES) == 256 and MC.CASES[0] == [] and MC.CASES[255] == []
    assert MC.MAX_TRIS == 5                                        # as in the classic table
    for c in range(256):                                           # (ambiguous faces cut off the INSIDE corners: c and 255 - c may differ in topology)
        used = {e for t in MC.CASES[c] for e in t}
        crossed = {i for i, (a, b) in enumerate(MC.EDGES) if ((c >> a) & 1) != ((c >> b) & 1)}
        assert used == crossed                                     # every crossed cube edge carries a vertex, no other does


def test_sphere_is_watertight_oriented_and_accurate():
    n = 48
    v, f = MC.marching_cubes(sphere_grid(n), 0.0, (2 / (n - 1),) * 3)
    v = v - 1
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    key = e[:, 0] * len(v) + e[:, 1]
    rev = e[:, 1] * len(v) + e[:, 0]
    assert len(np.unique(key)) == len(key) and np.array_equal(np.sort(key), np.sort(rev))      # every edge once in each direction
    vol = np.einsum("ij,ij->i", v[f[:, 0]], np.cross(v[f[:, 1]], v[f[:, 2]])).sum() / 6
    assert abs(vol / (4 / 3 * np.pi * 0.7 ** 3) - 1) < 5e-3 and vol > 0
    assert np.abs(np.linalg.norm(v, axis=1) - 0.7).max() < 1e-3


def test_random_volume_has_no_interior_holes():
    rng = np.random.default_rng(0)
    n = 14
    v, f = MC.marching_cubes(rng.standard_normal((n, n, n)), 0.0)
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), 1)
    uniq, cnt = np.unique(e, axis=0, return_counts=True)
    assert cnt.max() == 2
    on_border = lambda p: np.any((p < 1e-9) | (p > n - 1 - 1e-9), axis=1)
    open_edges = uniq[cnt == 1]
    assert np.all(on_border(v[open_edges[:, 0]]) & on_border(v[open_edges[:, 1]]))          # ambiguous faces are resolved consistently


@pytest.mark.parametrize("n", [24, 40])
def test_marching_tetrahedra_surface_coincides_with_marching_cubes(n):
    from hortimapping_b200.marching import marching_tetrahedra
    g = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    vol = np.sqrt((X / 0.8) ** 2 + (Y / 0.6) ** 2 + (Z / 0.7) ** 2) - 1 + 0.08 * np.sin(5 * X) * np.cos(4 * Y)      # a bumpy ellipsoid
    h = 2 / (n - 1)
    vc, fc = MC.marching_cubes(vol, 0.0, (h,) * 3)
    vt, ft = marching_tetrahedra(vol, 0.0, (h,) * 3)
    d_tc = point_to_mesh_distance(vt, vc, fc) / h
    d_ct = point_to_mesh_distance(vc, vt, ft) / h
    chamfer = 0.5 * (d_tc.mean() + d_ct.mean())
    assert chamfer < 0.05 and max(d_tc.max(), d_ct.max()) < 0.5, (chamfer, d_tc.max(), d_ct.max())
