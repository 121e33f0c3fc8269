"""Host-side mirror of wild_completion/mesher.py `MeshExtractor` on top of the C ABI.

The N^3 SDF grid (create_voxel_grid(N) * cube_radius, the reference's sheared grid, utils.py:542-562) is
evaluated on the GPU by hm_sdf_grid; iso-surface extraction follows utils.py:565-588.
"""
from __future__ import annotations

import numpy as np
import torch

from .decoder import Decoder
from .marching import extract_isosurface


class ForceKeyErrorDict(dict):
    """wild_completion/utils.py:524-526: attribute access that raises KeyError for missing keys."""

    def __getattr__(self, name):
        return self[name]

    __setattr__ = dict.__setitem__

    def __missing__(self, name):
        raise KeyError(name)


def convert_sdf_voxels_to_mesh(sdf_3d: torch.Tensor, cube_radius: float):
    """wild_completion/utils.py:565-588."""
    vol = sdf_3d.detach().cpu().numpy()
    n = vol.shape[0]
    voxel_size = 2.0 / (n - 1)
    verts, faces = extract_isosurface(vol, 0.0, [voxel_size] * 3)
    verts = np.array(verts, dtype=np.float64)
    verts += np.array([-1.0, -1.0, -1.0])
    verts *= cube_radius
    return verts, faces


def _skimage_available() -> bool:
    try:
        from skimage import measure  # type: ignore  # noqa: F401
        return True
    except ImportError:
        return False


class MeshExtractor(object):
    """`iso` selects the iso-surface extractor: "device" = hm_isosurface (marching tetrahedra on the GPU, only vertices and
    faces leave the device), "skimage" = the reference's own host call, "auto" (default) = skimage when it is importable
    (the mesh is then exactly what the reference builds from the same grid), else the device extractor."""

    def __init__(self, decoder: Decoder, code_len=64, voxels_dim=64, cube_radius=1.0, iso: str = None):
        import os
        iso = iso or os.environ.get("HM_MESHER_ISO", "auto")      # the host scripts construct it without `iso`: the environment decides
        self.decoder = decoder
        self.code_len = code_len
        self.voxels_dim = voxels_dim
        self.cube_radius = cube_radius
        assert iso in ("auto", "device", "skimage")
        self.iso = ("skimage" if _skimage_available() else "device") if iso == "auto" else iso
        with torch.no_grad():
            self.voxel_points = decoder.voxel_grid(self.voxels_dim, self.cube_radius)   # mesher.py:12

    def sdf_grid(self, code: torch.Tensor) -> torch.Tensor:
        return self.decoder.sdf_grid(code, self.voxels_dim, self.cube_radius)

    def extract_mesh_from_code(self, code):
        """mesher.py:14-24."""
        sdf = self.sdf_grid(code)
        if self.iso == "device":
            v, f = self.decoder.isosurface(sdf, 0.0, 2.0 / (self.voxels_dim - 1), affine_radius=self.cube_radius)
            return ForceKeyErrorDict(vertices=v.cpu().numpy(), faces=f.cpu().numpy())
        vertices, faces = convert_sdf_voxels_to_mesh(sdf.view(self.voxels_dim, self.voxels_dim, self.voxels_dim), self.cube_radius)
        return ForceKeyErrorDict(vertices=vertices.astype("float32"), faces=np.asarray(faces).astype("int32"))

    def complete_mesh_arrays(self, latent, transform):
        """Vertices transformed by the 4x4 `transform` and faces, without open3d."""
        m = self.extract_mesh_from_code(latent)
        T = np.asarray(transform, np.float64)
        v = m.vertices.astype(np.float64) @ T[:3, :3].T + T[:3, 3]
        return v.astype(np.float32), m.faces

    def complete_mesh(self, latent, transform, color):
        """mesher.py:26-33 (needs open3d, like the reference)."""
        import open3d as o3d
        cur_mesh = self.extract_mesh_from_code(latent)
        mesh_o3d = o3d.geometry.TriangleMesh(o3d.utility.Vector3dVector(cur_mesh.vertices), o3d.utility.Vector3iVector(cur_mesh.faces))
        mesh_o3d.compute_vertex_normals()
        mesh_o3d.paint_uniform_color(color)
        mesh_o3d = mesh_o3d.transform(transform)
        return mesh_o3d
