"""numpy stand-in for the few open3d calls of the reference's host scripts (see ../README.md).  Not open3d."""
from __future__ import annotations

import copy as _copy
import types

import numpy as np

__version__ = "0.0-standin"
_rng = np.random.default_rng(0)


class _Vec(np.ndarray):
    pass


def _vec(a, dtype, cols=3):
    arr = np.array(a, dtype=dtype).reshape(-1, cols) if len(a) else np.zeros((0, cols), dtype)
    return arr


class AxisAlignedBoundingBox:
    def __init__(self, min_bound=(0, 0, 0), max_bound=(0, 0, 0)):
        self.min_bound, self.max_bound = np.asarray(min_bound, np.float64), np.asarray(max_bound, np.float64)

    def get_center(self):
        return (self.min_bound + self.max_bound) * 0.5

    def get_extent(self):
        return self.max_bound - self.min_bound


class PointCloud:
    def __init__(self, points=None):
        self.points = _vec(points if points is not None else [], np.float64)
        self.colors = np.zeros((0, 3))
        self.normals = np.zeros((0, 3))

    def _take(self, idx):
        out = PointCloud(self.points[idx])
        if len(self.colors) == len(self.points):
            out.colors = self.colors[idx]
        if len(self.normals) == len(self.points):
            out.normals = self.normals[idx]
        return out

    def select_by_index(self, indices, invert=False):
        idx = np.asarray(indices, np.int64)
        if invert:
            m = np.ones(len(self.points), bool)
            m[idx] = False
            idx = np.nonzero(m)[0]
        return self._take(idx)

    def get_axis_aligned_bounding_box(self):
        return AxisAlignedBoundingBox(self.points.min(0), self.points.max(0))

    def crop(self, box):
        m = np.all((self.points >= box.min_bound) & (self.points <= box.max_bound), axis=1)
        return self._take(np.nonzero(m)[0])

    def voxel_down_sample(self, voxel_size):
        if not len(self.points):
            return PointCloud()
        origin = self.points.min(0) - voxel_size * 0.5
        key = np.floor((self.points - origin) / voxel_size).astype(np.int64)
        _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
        inv = inv.reshape(-1)
        out = PointCloud(np.stack([np.bincount(inv, self.points[:, c]) / cnt for c in range(3)], 1))
        if len(self.colors) == len(self.points):
            out.colors = np.stack([np.bincount(inv, self.colors[:, c]) / cnt for c in range(3)], 1)
        return out

    def cluster_dbscan(self, eps, min_points, print_progress=False):
        from oracle.preprocess_oracle import cluster_dbscan
        return cluster_dbscan(self.points, eps, min_points).tolist()

    def paint_uniform_color(self, color):
        self.colors = np.tile(np.asarray(color, np.float64), (len(self.points), 1))
        return self

    def transform(self, T):
        T = np.asarray(T, np.float64)
        self.points = self.points @ T[:3, :3].T + T[:3, 3]
        return self


class TriangleMesh:
    def __init__(self, vertices=None, triangles=None):
        self.vertices = _vec(vertices if vertices is not None else [], np.float64)
        self.triangles = _vec(triangles if triangles is not None else [], np.int32)
        self.vertex_colors = np.zeros((0, 3))
        self.vertex_normals = np.zeros((0, 3))

    def compute_vertex_normals(self):
        v, f = self.vertices, self.triangles
        n = np.zeros_like(v)
        if len(f):
            fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
            for c in range(3):
                np.add.at(n, f[:, c], fn)
            n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
        self.vertex_normals = n
        return self

    def paint_uniform_color(self, color):
        self.vertex_colors = np.tile(np.asarray(color, np.float64), (len(self.vertices), 1))
        return self

    def transform(self, T):
        T = np.asarray(T, np.float64)
        self.vertices = self.vertices @ T[:3, :3].T + T[:3, 3]
        return self

    def sample_points_uniformly(self, number_of_points=100, use_triangle_normal=False):
        v, f = self.vertices, self.triangles
        a = 0.5 * np.linalg.norm(np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]), axis=1)
        t = _rng.choice(len(f), size=number_of_points, p=a / a.sum())
        r1, r2 = np.sqrt(_rng.random(number_of_points)), _rng.random(number_of_points)
        w = np.stack([1 - r1, r1 * (1 - r2), r1 * r2], 1)
        out = PointCloud((v[f[t]] * w[:, :, None]).sum(1))
        if len(self.vertex_colors) == len(v):
            out.colors = (self.vertex_colors[f[t]] * w[:, :, None]).sum(1)
        return out

    def cluster_connected_triangles(self):
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        f = self.triangles
        nv = len(self.vertices)
        rows = np.concatenate([f[:, 0], f[:, 1]])
        cols = np.concatenate([f[:, 1], f[:, 2]])
        _, lab = connected_components(coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(nv, nv)), directed=False)
        tl = lab[f[:, 0]]
        uniq, tl = np.unique(tl, return_inverse=True)
        a = 0.5 * np.linalg.norm(np.cross(self.vertices[f[:, 1]] - self.vertices[f[:, 0]], self.vertices[f[:, 2]] - self.vertices[f[:, 0]]), axis=1)
        return tl.tolist(), np.bincount(tl).tolist(), np.bincount(tl, a).tolist()

    def remove_triangles_by_mask(self, mask):
        self.triangles = self.triangles[~np.asarray(mask, bool)]

    @staticmethod
    def create_coordinate_frame(size=1.0, origin=(0, 0, 0)):
        return TriangleMesh()


# ---- PLY (ascii and binary_little_endian; vertex x y z [nx ny nz] [red green blue], face vertex_indices)
_PLY_T = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1", "char": "i1", "int": "<i4", "int32": "<i4",
          "uint": "<u4", "uint32": "<u4", "short": "<i2", "ushort": "<u2"}


def _read_ply(path):
    with open(path, "rb") as fh:
        assert fh.readline().strip() == b"ply"
        fmt, elems = None, []
        while True:
            line = fh.readline().decode().strip()
            if line == "end_header":
                break
            tok = line.split()
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elems.append([tok[1], int(tok[2]), []])
            elif tok[0] == "property":
                elems[-1][2].append(tok[1:])
        data = {}
        for name, count, props in elems:
            if any(p[0] == "list" for p in props):
                rows = []
                if fmt == "ascii":
                    for _ in range(count):
                        t = fh.readline().split()
                        rows.append([int(x) for x in t[1:1 + int(t[0])]])
                else:
                    lp = [p for p in props if p[0] == "list"][0]
                    ct, it = np.dtype(_PLY_T[lp[1]]), np.dtype(_PLY_T[lp[2]])
                    for _ in range(count):
                        k = int(np.frombuffer(fh.read(ct.itemsize), ct)[0])
                        rows.append(np.frombuffer(fh.read(it.itemsize * k), it).tolist())
                data[name] = rows
            else:
                dt = np.dtype([(p[1], _PLY_T[p[0]]) for p in props])
                if fmt == "ascii":
                    arr = np.loadtxt([fh.readline().decode() for _ in range(count)], ndmin=2) if count else np.zeros((0, len(props)))
                    data[name] = {p[1]: arr[:, i] for i, p in enumerate(props)}
                else:
                    arr = np.frombuffer(fh.read(dt.itemsize * count), dt)
                    data[name] = {n: arr[n] for n in dt.names}
    return data


def _write_ply(path, verts, colors=None, faces=None):
    verts = np.asarray(verts, np.float64)
    with open(path, "wb") as fh:
        hdr = ["ply", "format binary_little_endian 1.0", "comment hortimapping_b200 open3d stand-in", f"element vertex {len(verts)}",
               "property double x", "property double y", "property double z"]
        has_c = colors is not None and len(colors) == len(verts)
        if has_c:
            hdr += ["property uchar red", "property uchar green", "property uchar blue"]
        if faces is not None:
            hdr += [f"element face {len(faces)}", "property list uchar int vertex_indices"]
        fh.write(("\n".join(hdr) + "\nend_header\n").encode())
        dt = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")] + ([("r", "u1"), ("g", "u1"), ("b", "u1")] if has_c else [])
        arr = np.zeros(len(verts), dt)
        arr["x"], arr["y"], arr["z"] = verts[:, 0], verts[:, 1], verts[:, 2]
        if has_c:
            c = np.clip(np.rint(np.asarray(colors) * 255), 0, 255).astype(np.uint8)
            arr["r"], arr["g"], arr["b"] = c[:, 0], c[:, 1], c[:, 2]
        fh.write(arr.tobytes())
        if faces is not None and len(faces):
            fa = np.zeros(len(faces), [("n", "u1"), ("i", "<i4", (3,))])
            fa["n"], fa["i"] = 3, np.asarray(faces, np.int32)
            fh.write(fa.tobytes())


def _xyz_colors(v):
    pts = np.stack([v["x"], v["y"], v["z"]], 1).astype(np.float64) if len(v.get("x", [])) else np.zeros((0, 3))
    col = np.stack([v["red"], v["green"], v["blue"]], 1).astype(np.float64) / 255.0 if "red" in v else np.zeros((0, 3))
    return pts, col


def _read_triangle_mesh(path):
    d = _read_ply(path)
    pts, col = _xyz_colors(d.get("vertex", {}))
    m = TriangleMesh(pts, [f[:3] for f in d.get("face", [])])
    m.vertex_colors = col
    return m


def _read_point_cloud(path):
    d = _read_ply(path)
    pts, col = _xyz_colors(d.get("vertex", {}))
    p = PointCloud(pts)
    p.colors = col
    return p


def _write_triangle_mesh(path, mesh, **kw):
    _write_ply(path, mesh.vertices, mesh.vertex_colors, mesh.triangles)
    return True


def _write_point_cloud(path, pcd, **kw):
    _write_ply(path, pcd.points, pcd.colors, None)
    return True


def _seed(s):
    global _rng
    _rng = np.random.default_rng(int(s))


geometry = types.SimpleNamespace(PointCloud=PointCloud, TriangleMesh=TriangleMesh, AxisAlignedBoundingBox=AxisAlignedBoundingBox)
utility = types.SimpleNamespace(Vector3dVector=lambda a=(): _vec(a, np.float64), Vector3iVector=lambda a=(): _vec(a, np.int32),
                                random=types.SimpleNamespace(seed=_seed))
io = types.SimpleNamespace(read_triangle_mesh=_read_triangle_mesh, write_triangle_mesh=_write_triangle_mesh, read_point_cloud=_read_point_cloud,
                           write_point_cloud=_write_point_cloud)
visualization = types.SimpleNamespace()
