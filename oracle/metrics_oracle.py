"""TEST INFRASTRUCTURE ONLY (see oracle/hm_oracle.py): CPU restatement of the reference's evaluation metrics.

metrics_3d/chamfer_distance.py:16-26 and metrics_3d/precision_recall.py:19-50 call open3d's
`PointCloud.compute_point_cloud_distance` (third-party, open3d==0.17, README.md:47; absent from this image): for every
point the Euclidean distance to its exact nearest neighbour in the other cloud, in double precision.  An exact
nearest-neighbour search has one answer, so it is restated with scipy.spatial.cKDTree (SURVEY.md 8c iii) and, for the
self-check in tests/test_oracle_golden.py, by brute force.  Parity with open3d itself is unpinned (it cannot run here).
"""
import numpy as np
from scipy.spatial import cKDTree


def point_cloud_distance(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """open3d PointCloud.compute_point_cloud_distance(src -> dst)."""
    return cKDTree(np.asarray(dst, np.float64)).query(np.asarray(src, np.float64), k=1)[0]


def point_cloud_distance_bruteforce(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    d = np.asarray(src, np.float64)[:, None, :] - np.asarray(dst, np.float64)[None, :, :]
    return np.sqrt((d * d).sum(-1).min(1))


def chamfer(gt: np.ndarray, pt: np.ndarray) -> float:
    """chamfer_distance.py:16-26 for one (gt, prediction) pair."""
    if len(pt) == 0:
        return 0
    return (np.mean(point_cloud_distance(gt, pt)) + np.mean(point_cloud_distance(pt, gt))) / 2


def precision_recall(gt: np.ndarray, pt: np.ndarray, thresholds: np.ndarray):
    """precision_recall.py:19-50 for one pair -> (precision[], recall[], fscore[]) in percent."""
    if len(pt) == 0:
        z = [0] * len(thresholds)
        return z, z, z
    d_p = point_cloud_distance(pt, gt)
    d_r = point_cloud_distance(gt, pt)
    pr, re, f1 = [], [], []
    for t in thresholds:
        p = 100 / len(d_p) * len(np.where(d_p < t)[0])
        r = 100 / len(d_r) * len(np.where(d_r < t)[0])
        pr.append(p)
        re.append(r)
        f1.append(0 if (p == 0 or r == 0) else 2 * p * r / (p + r))
    return pr, re, f1
