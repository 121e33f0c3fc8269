"""GPU parity of the device iso-surface extractor (hm_isosurface) against the host extractor
hortimapping_b200/marching.py, which it mirrors index by index: faces are compared exactly (integer work), vertices to
fp32 rounding (both interpolate in fp64; the host keeps the first of several equivalent evaluations of a welded vertex)."""
import numpy as np
import pytest
import torch

from hortimapping_b200.marching import marching_tetrahedra
from tests.helpers import load_npz

pytestmark = pytest.mark.gpu


def _sphere(n, r=0.6, c=(0.05, -0.1, 0.02)):
    g = np.linspace(-1, 1, n, dtype=np.float32)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    return (np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - r).astype(np.float32)


def _compare(dec, vol, level=0.0, spacing=1.0):
    v_ref, f_ref = marching_tetrahedra(vol, level, (spacing,) * 3)
    v, f = dec.isosurface(torch.from_numpy(vol).cuda(), level, spacing)
    v, f = v.cpu().numpy(), f.cpu().numpy()
    assert v.shape == v_ref.shape and f.shape == f_ref.shape, (v.shape, v_ref.shape, f.shape, f_ref.shape)
    np.testing.assert_array_equal(f, f_ref)
    if v.size:
        scale = max(1.0, float(np.abs(v_ref).max()))
        assert np.abs(v - v_ref).max() <= 2 * np.finfo(np.float32).eps * scale
        assert (v != v_ref).any(axis=1).mean() <= 0.01          # all but a handful are bit-identical
    return v, f


def test_isosurface_matches_host_extractor_on_analytic_fields():
    from tests.gpu_helpers import pepper_decoder
    dec = pepper_decoder()
    for n in (2, 3, 9, 20, 33):
        _compare(dec, _sphere(n), 0.0, 2.0 / (n - 1))
    g = np.random.default_rng(3)
    noise = g.standard_normal((17, 17, 17)).astype(np.float32)      # every tetrahedron case, many surface sheets
    _compare(dec, noise, 0.0, 1.0)
    _compare(dec, noise, 0.25, 0.5)
    flat = _sphere(12)
    flat[flat > 0.3] = 0.0                                           # exact zeros on the outside (t = 0 / 1 interpolation)
    _compare(dec, flat, 0.0, 1.0)


def test_isosurface_empty_and_orientation():
    from tests.gpu_helpers import pepper_decoder
    dec = pepper_decoder()
    v, f = dec.isosurface(torch.ones(8, 8, 8).cuda(), 0.0, 1.0)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    v, f = dec.isosurface(-torch.ones(8, 8, 8).cuda(), 0.0, 1.0)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    # closed, outward-oriented surface of a sphere: signed volume = +4/3 pi r^3 within the discretisation error
    n, r = 48, 0.6
    v, f = dec.isosurface(torch.from_numpy(_sphere(n, r, (0, 0, 0))).cuda(), 0.0, 2.0 / (n - 1), affine_radius=1.0)
    v, f = v.cpu().numpy().astype(np.float64), f.cpu().numpy()
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    vol = (a * np.cross(b, c)).sum() / 6.0
    assert abs(vol - 4 / 3 * np.pi * r ** 3) / (4 / 3 * np.pi * r ** 3) < 0.02
    edges = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    _, counts = np.unique(edges, axis=0, return_counts=True)
    assert (counts == 2).all(), "the mesh must be watertight (every edge shared by exactly two faces)"
    assert np.abs(np.linalg.norm(v, axis=1) - r).max() < 2.0 / (n - 1)


def test_mesh_extractor_device_path_matches_host_path_on_a_decoder_grid():
    """MeshExtractor.extract_mesh_from_code (wild_completion/mesher.py:14-24) with the device extractor vs the host
    extractor on the same decoder SDF grid: same faces, same vertices."""
    from hortimapping_b200.mesher import MeshExtractor, convert_sdf_voxels_to_mesh
    from hortimapping_b200 import marching
    from tests.gpu_helpers import pepper_decoder
    dec = pepper_decoder()
    m = load_npz("misc")
    lat = torch.from_numpy(m["grid_lat"]).cuda()
    mx = MeshExtractor(dec, code_len=32, voxels_dim=40, cube_radius=0.08, iso="device")
    mesh = mx.extract_mesh_from_code(lat)
    sdf = mx.sdf_grid(lat)
    v_ref, f_ref = marching.marching_tetrahedra(sdf.cpu().numpy(), 0.0, (2.0 / 39,) * 3)
    v_ref = ((v_ref.astype(np.float64) + -1.0) * 0.08).astype(np.float32)
    assert mesh.vertices.dtype == np.float32 and mesh.faces.dtype == np.int32
    np.testing.assert_array_equal(mesh.faces, f_ref)
    np.testing.assert_allclose(mesh.vertices, v_ref, rtol=0, atol=1e-8)
    assert mesh.vertices.shape[0] > 1000 and np.abs(mesh.vertices).max() < 0.08


def test_fused_grid_equals_explicit_points_and_full_size_grid():
    """hm_sdf_grid generates the voxel-grid points inside the decoder kernel (fused grid sample + decode): the result must be
    bit-identical to decoding the explicit points of hm_voxel_grid, and the BASELINE config-1 grid (128^3) must run through
    grid + iso-surface (size-independent properties: finite SDF, vertices inside the cube, every face index valid)."""
    from tests.gpu_helpers import pepper_decoder
    dec = pepper_decoder()
    m = load_npz("misc")
    lat = torch.from_numpy(m["grid_lat"]).cuda()
    for n in (20, 40, 41):
        fused = dec.sdf_grid(lat, n, 0.08).reshape(-1)
        explicit = dec.sdf(lat, dec.voxel_grid(n, 0.08))
        assert torch.equal(fused, explicit), n
    n = 128
    sdf = dec.sdf_grid(lat, n, 0.08)
    assert bool(torch.isfinite(sdf).all()) and float(sdf.min()) < 0 < float(sdf.max())
    v, f = dec.isosurface(sdf, 0.0, 2.0 / (n - 1), affine_radius=0.08)
    assert v.shape[0] > 10000 and f.shape[0] > 20000
    assert float(v.abs().max()) <= 0.08 and int(f.min()) == 0 and int(f.max()) == v.shape[0] - 1
    # the vertices sit on the level set up to the reference's grid shear (create_voxel_grid quirk, SURVEY.md 7.5: the samples
    # are taken at sheared positions but meshed as a regular grid -- a few millimetres at this resolution)
    s = dec.sdf(lat, v).abs().max()
    assert float(s) < 0.005


@pytest.mark.parametrize("n", [40, 128])
def test_device_isosurface_coincides_with_marching_cubes(n):
    """N2 at surface level (wild_completion/utils.py:565-588): the reference runs skimage's marching cubes on the N^3 grid; the
    device extractor (marching tetrahedra) must give the same surface.  Both are extracted from the SAME device SDF grid of the
    shipped model (40^3 = the shipped configs' grid, 128^3 = BASELINE.json configs[0]) and compared by exact point-to-triangle
    distances in voxel units: symmetric mean (Chamfer) < 0.1 voxel, as VERDICT r01 item 8 asks."""
    import torch
    from oracle import marching_cubes as MC
    from tests.gpu_helpers import pepper_decoder
    from tests.helpers import pepper_weights, point_to_mesh_distance
    dec = pepper_decoder()
    _, _, codes = pepper_weights()
    lat = torch.from_numpy(codes[5]).cuda()
    sdf = dec.sdf_grid(lat, n, 0.08)
    h = 2.0 / (n - 1)
    vd, fd = dec.isosurface(sdf, 0.0, h)
    vd, fd = vd.cpu().numpy().astype(np.float64), fd.cpu().numpy()
    vc, fc = MC.marching_cubes(sdf.cpu().numpy(), 0.0, (h,) * 3)
    assert len(fd) > 1000 and len(fc) > 1000
    d1 = point_to_mesh_distance(vd, vc, fc) / h
    d2 = point_to_mesh_distance(vc, vd, fd) / h
    chamfer = 0.5 * (d1.mean() + d2.mean())
    assert chamfer < 0.1 and max(d1.max(), d2.max()) < 1.0, (chamfer, d1.max(), d2.max())
    print(f"{n}^3: {len(fd)} device faces vs {len(fc)} marching-cubes faces, Chamfer {chamfer:.4f} voxel, max {max(d1.max(), d2.max()):.3f} voxel")
