"""Host-side mirror of the point-cloud preparation of wild_completion/utils.py:389-459 on top of the C ABI:
`clean_mesh`, `clean_pcd`, `get_pose_init` with the reference's signatures, return values and print-outs.

The geometry containers stay whatever the host script uses (open3d's PointCloud / TriangleMesh, or any object with the same
few members: `.points`, `.select_by_index`, `.sample_points_uniformly`, mesh cluster helpers); the arithmetic -- DBSCAN
labelling, bounding boxes, the crop of the background cloud and its mean offset -- runs in libhortimapping_b200.so.
`dropin.install` rebinds these names on the reference's own `wild_completion.utils` module.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import Counter

import numpy as np
import torch

from . import _lib
from ._lib import check


def _dev_points(points) -> torch.Tensor:
    if not torch.cuda.is_available():
        raise RuntimeError("hortimapping_b200.preprocess needs a CUDA device (no CPU fallback)")
    a = points if isinstance(points, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(points, np.float64).reshape(-1, 3)))
    return a.to(device="cuda", dtype=torch.float64).contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def dbscan_labels(points, eps: float, min_points: int) -> np.ndarray:
    """open3d `PointCloud.cluster_dbscan(eps, min_points)` labels (int32, -1 = noise) computed by hm_dbscan."""
    p = _dev_points(points)
    n = int(p.shape[0])
    labels = torch.empty(n, dtype=torch.int32, device=p.device)
    nc = C.c_int32(0)
    check(_lib.lib().hm_dbscan(None, p.data_ptr(), n, float(eps), int(min_points), labels.data_ptr(), C.byref(nc), _stream()), "hm_dbscan")
    return labels.cpu().numpy()


def cloud_bounds(points):
    p = _dev_points(points)
    mn, mx = (C.c_double * 3)(), (C.c_double * 3)()
    check(_lib.lib().hm_cloud_bounds(None, p.data_ptr(), int(p.shape[0]), mn, mx, _stream()), "hm_cloud_bounds")
    return np.array(mn[:]), np.array(mx[:])


def crop_mean_offset(points, box_min, box_max, center):
    p = _dev_points(points)
    arr = lambda v: (C.c_double * 3)(*[float(x) for x in v])
    cnt, mean = C.c_int64(0), (C.c_double * 3)()
    check(_lib.lib().hm_crop_mean_offset(None, p.data_ptr(), int(p.shape[0]), arr(box_min), arr(box_max), arr(center), C.byref(cnt), mean, _stream()),
          "hm_crop_mean_offset")
    return int(cnt.value), np.array(mean[:])


def clean_pcd(cur_pcd, cluster_dist_thre=0.01, outlier_point_ratio=0.02):
    """utils.py:408-419: keep the most frequent DBSCAN label (noise included in the vote, like the reference's Counter)."""
    cur_point_count = len(cur_pcd.points)
    min_instance_pts = int(cur_point_count * outlier_point_ratio)
    cur_cluster_labels = dbscan_labels(np.asarray(cur_pcd.points), cluster_dist_thre, min_instance_pts).astype(int)
    cluster_counter = Counter(cur_cluster_labels.tolist())
    mode_label = cluster_counter.most_common(1)[0][0]
    main_cluster_indices = np.where(cur_cluster_labels == mode_label)[0].tolist()
    return cur_pcd.select_by_index(main_cluster_indices)


def clean_mesh(cur_mesh, sample_point_count=5000, cluster_dist_thre=0.01, outlier_point_ratio=0.02,
               filter_isolated_mesh=False, filter_cluster_min_tri=20):
    """utils.py:389-406.  Surface sampling stays the mesh object's own `sample_points_uniformly` (open3d's RNG)."""
    if filter_isolated_mesh:
        triangle_clusters, cluster_n_triangles, cluster_area = cur_mesh.cluster_connected_triangles()
        triangle_clusters = np.asarray(triangle_clusters)
        cluster_n_triangles = np.asarray(cluster_n_triangles)
        triangles_to_remove = cluster_n_triangles[triangle_clusters] < filter_cluster_min_tri
        cur_mesh.remove_triangles_by_mask(triangles_to_remove)
    cur_pcd = cur_mesh.sample_points_uniformly(number_of_points=sample_point_count)
    return clean_pcd(cur_pcd, cluster_dist_thre, outlier_point_ratio)


def get_pose_init(cur_pcd, bg_pcd, bbx_pad=0.01, min_bbx_size=0.03, max_bbx_size=0.16, min_nearby_bg_pts=10, max_init_rot_deg=45):
    """utils.py:422-459 -> (cur_center, init_rot_y_rad, bbx_size, valid_flag)."""
    valid_flag = True
    mn, mx = cloud_bounds(np.asarray(cur_pcd.points))
    cur_center, cur_extent = (mn + mx) * 0.5, mx - mn          # AxisAlignedBoundingBox.get_center / get_extent
    bbx_size = max(cur_extent) + bbx_pad
    print("Init bbx size (m):", bbx_size)
    if bbx_size > max_bbx_size:
        print("Too large bbx, could not be a valid object, skip")
        valid_flag = False
    if bbx_size < min_bbx_size:
        print("Too small bbx, could not be a valid object, skip")
        valid_flag = False
    init_rot_y_rad = 0.
    max_init_rot = max_init_rot_deg / 180. * math.pi
    if valid_flag:
        cur_center[1] += ((bbx_size - cur_extent[1]) * 0.5)
        if cur_extent[1] == max(cur_extent):
            cur_center[1] += 0.01
        box_bg_min = [cur_center[0] - 0.6 * bbx_size, cur_center[1] - 0.8 * bbx_size, cur_center[2] + 0.2 * bbx_size]
        box_bg_max = [cur_center[0] + 0.6 * bbx_size, cur_center[1] + 1.0 * bbx_size, cur_center[2] + 1.2 * bbx_size]
        n_bg, rot_vec = crop_mean_offset(np.asarray(bg_pcd.points), box_bg_min, box_bg_max, cur_center)
        if n_bg > min_nearby_bg_pts:
            init_rot_y_rad = 0.5 * math.pi - np.arctan2(rot_vec[2], rot_vec[0])
            init_rot_y_rad = max(min(init_rot_y_rad, max_init_rot), -max_init_rot)
        print("Init rot around y axis (deg):", init_rot_y_rad * 180. / math.pi)
    return cur_center, init_rot_y_rad, bbx_size, valid_flag
