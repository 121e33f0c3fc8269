"""CPU oracle of the point-cloud preparation (wild_completion/utils.py:389-459) -- TEST INFRASTRUCTURE, never imported by the
product.

`cluster_dbscan` restates open3d's PointCloud::ClusterDBSCAN (open3d==0.17, README.md:47; third-party and NOT installed here,
so parity with open3d itself is unpinned -- the restatement follows the published algorithm line by line: radius search with
squared distances `< eps^2` that includes the query point, sequential scan over the point indices, breadth-first expansion
through core points, noise points absorbed as border points by the first cluster that reaches them).  Because it is written
as the SEQUENTIAL algorithm it is an independent check of the order-independent parallel formulation in csrc/preprocess.cu.
`clean_pcd` / `get_pose_init` restate utils.py:408-459 on plain numpy arrays.
"""
from __future__ import annotations

import math
from collections import Counter

import numpy as np
from scipy.spatial import cKDTree


def cluster_dbscan(points: np.ndarray, eps: float, min_points: int) -> np.ndarray:
    pts = np.asarray(points, np.float64).reshape(-1, 3)
    n = pts.shape[0]
    tree = cKDTree(pts)
    cand = tree.query_ball_point(pts, eps * (1 + 1e-9) + 1e-300)          # superset; the exact test below decides
    eps2 = eps * eps
    nbs = []
    for i in range(n):
        c = np.asarray(cand[i], np.int64)
        d = pts[c] - pts[i]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        nbs.append(c[d2 < eps2])
    labels = np.full(n, -2, np.int64)
    cluster = 0
    for idx in range(n):
        if labels[idx] != -2:
            continue
        if len(nbs[idx]) < min_points:
            labels[idx] = -1
            continue
        nxt = list(nbs[idx])
        seen = {idx}
        labels[idx] = cluster
        while nxt:
            nb = int(nxt.pop())
            seen.add(nb)
            if labels[nb] == -1:
                labels[nb] = cluster
            if labels[nb] != -2:
                continue
            labels[nb] = cluster
            if len(nbs[nb]) >= min_points:
                nxt.extend(int(q) for q in nbs[nb] if q not in seen)
        cluster += 1
    return labels.astype(np.int32)


def clean_pcd(points: np.ndarray, cluster_dist_thre=0.01, outlier_point_ratio=0.02) -> np.ndarray:
    """utils.py:408-419 -> indices of the kept points."""
    n = len(points)
    labels = cluster_dbscan(points, cluster_dist_thre, int(n * outlier_point_ratio)).astype(int)
    mode = Counter(labels.tolist()).most_common(1)[0][0]
    return np.where(labels == mode)[0]


def get_pose_init(points: np.ndarray, bg_points: np.ndarray, bbx_pad=0.01, min_bbx_size=0.03, max_bbx_size=0.16,
                  min_nearby_bg_pts=10, max_init_rot_deg=45):
    """utils.py:422-459 on arrays -> (center, init_rot_y_rad, bbx_size, valid)."""
    pts = np.asarray(points, np.float64)
    mn, mx = pts.min(0), pts.max(0)
    center, extent = (mn + mx) * 0.5, mx - mn
    bbx_size = max(extent) + bbx_pad
    valid = not (bbx_size > max_bbx_size or bbx_size < min_bbx_size)
    rot = 0.0
    max_rot = max_init_rot_deg / 180. * math.pi
    if valid:
        center[1] += (bbx_size - extent[1]) * 0.5
        if extent[1] == max(extent):
            center[1] += 0.01
        lo = np.array([center[0] - 0.6 * bbx_size, center[1] - 0.8 * bbx_size, center[2] + 0.2 * bbx_size])
        hi = np.array([center[0] + 0.6 * bbx_size, center[1] + 1.0 * bbx_size, center[2] + 1.2 * bbx_size])
        bg = np.asarray(bg_points, np.float64)
        inside = np.all((bg >= lo) & (bg <= hi), axis=1)
        if inside.sum() > min_nearby_bg_pts:
            v = np.mean(bg[inside] - center, 0)
            rot = 0.5 * math.pi - np.arctan2(v[2], v[0])
            rot = max(min(rot, max_rot), -max_rot)
    return center, rot, bbx_size, valid
