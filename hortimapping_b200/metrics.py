"""Host-side mirror of the reference's evaluation metrics (metrics_3d/chamfer_distance.py, metrics_3d/precision_recall.py,
metrics_3d/metric.py) on top of the C ABI: the nearest-neighbour distances that open3d's
`compute_point_cloud_distance` produces on the host are computed exactly, in fp64, by `hm_nn_distance` on the GPU; only
means and threshold counts come back.  Same class names, methods and bookkeeping as the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check


def _points(geom) -> np.ndarray:
    """metrics_3d/metric.py:36-58 convert_to_pcd, without requiring open3d: numpy / torch arrays use their first three
    columns; open3d point clouds their `.points`; open3d meshes are sampled with open3d's own sample_points_uniformly
    (1 000 000 points, exactly the reference's call)."""
    if isinstance(geom, torch.Tensor):
        return geom.detach().cpu().numpy()[:, :3].astype(np.float64)
    if isinstance(geom, np.ndarray):
        return np.asarray(geom[:, :3], np.float64)
    if hasattr(geom, "points"):
        return np.asarray(geom.points, np.float64).reshape(-1, 3)
    if hasattr(geom, "vertices") and hasattr(geom, "sample_points_uniformly"):
        return np.asarray(geom.sample_points_uniformly(1000000).points, np.float64).reshape(-1, 3)
    raise AssertionError("{} type not supported".format(type(geom)))


def nn_distance(query, target, device=None) -> torch.Tensor:
    """Distance of every query point to its nearest target point (fp64 CUDA tensor)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    q = torch.as_tensor(np.ascontiguousarray(query, np.float64) if not isinstance(query, torch.Tensor) else query).to(dev, torch.float64).contiguous()
    t = torch.as_tensor(np.ascontiguousarray(target, np.float64) if not isinstance(target, torch.Tensor) else target).to(dev, torch.float64).contiguous()
    out = torch.empty(q.shape[0], device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        check(_lib.lib().hm_nn_distance(None, q.data_ptr(), q.shape[0], t.data_ptr(), t.shape[0], out.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream), "hm_nn_distance")
    return out


class Metrics3D:
    def prediction_is_empty(self, geom):                       # metric.py:15-33
        if isinstance(geom, (np.ndarray, torch.Tensor)):
            return self.is_empty(len(geom[:, :3]))
        if hasattr(geom, "points"):
            return self.is_empty(len(geom.points))
        if hasattr(geom, "vertices"):
            return self.is_empty(len(geom.vertices))
        assert False, "{} type not supported".format(type(geom))

    convert_to_pcd = staticmethod(_points)

    @staticmethod
    def is_empty(length):
        return not bool(length)


class ChamferDistance(Metrics3D):
    """metrics_3d/chamfer_distance.py:11-36."""

    def __init__(self):
        self.cd_array = []

    def update(self, gt, pt):
        if self.prediction_is_empty(pt):
            self.cd_array.append(0)
            return
        g, p = _points(gt), _points(pt)
        dist_pt_2_gt = nn_distance(p, g)
        dist_gt_2_pt = nn_distance(g, p)
        d = (float(dist_gt_2_pt.mean().item()) + float(dist_pt_2_gt.mean().item())) / 2
        self.cd_array.append(d)

    def reset(self):
        self.cd_array = []

    def compute(self):
        return sum(self.cd_array) / len(self.cd_array)


class PrecisionRecall(Metrics3D):
    """metrics_3d/precision_recall.py:11-107."""

    def __init__(self, min_t, max_t, num):
        self.thresholds = np.linspace(min_t, max_t, num)
        self.reset()

    def update(self, gt, pt):
        if self.prediction_is_empty(pt):
            for t in self.thresholds:
                self.pr_dict[t].append(0)
                self.re_dict[t].append(0)
                self.f1_dict[t].append(0)
            return
        g, p = _points(gt), _points(pt)
        dist_pt_2_gt = nn_distance(p, g)                          # precision: predicted --> ground truth
        dist_gt_2_pt = nn_distance(g, p)                          # recall: ground truth --> predicted
        thr = torch.from_numpy(self.thresholds).to(dist_pt_2_gt.device)
        n_p = (dist_pt_2_gt[:, None] < thr[None, :]).sum(0).cpu().numpy()
        n_r = (dist_gt_2_pt[:, None] < thr[None, :]).sum(0).cpu().numpy()
        for k, t in enumerate(self.thresholds):
            pr = 100 / len(dist_pt_2_gt) * int(n_p[k])
            re = 100 / len(dist_gt_2_pt) * int(n_r[k])
            self.pr_dict[t].append(pr)
            self.re_dict[t].append(re)
            self.f1_dict[t].append(0 if (pr == 0 or re == 0) else 2 * pr * re / (pr + re))

    def reset(self):
        self.pr_dict = {t: [] for t in self.thresholds}
        self.re_dict = {t: [] for t in self.thresholds}
        self.f1_dict = {t: [] for t in self.thresholds}

    def compute_at_threshold(self, threshold):
        t = self.find_nearest_threshold(threshold)
        pr = sum(self.pr_dict[t]) / len(self.pr_dict[t])
        re = sum(self.re_dict[t]) / len(self.re_dict[t])
        f1 = sum(self.f1_dict[t]) / len(self.f1_dict[t])
        return pr, re, f1, t

    def compute_auc(self):
        import scipy.integrate
        dx = self.thresholds[1] - self.thresholds[0]
        perfect_predictor = scipy.integrate.simpson(np.ones_like(self.thresholds), dx=dx)
        pr, re, f1 = self.compute_at_all_thresholds()
        return (scipy.integrate.simpson(pr, dx=dx) / perfect_predictor, scipy.integrate.simpson(re, dx=dx) / perfect_predictor,
                scipy.integrate.simpson(f1, dx=dx) / perfect_predictor)

    def compute_at_all_thresholds(self):
        pr = [sum(self.pr_dict[t]) / len(self.pr_dict[t]) for t in self.thresholds]
        re = [sum(self.re_dict[t]) / len(self.re_dict[t]) for t in self.thresholds]
        f1 = [sum(self.f1_dict[t]) / len(self.f1_dict[t]) for t in self.thresholds]
        return pr, re, f1

    def find_nearest_threshold(self, value):
        idx = (np.abs(self.thresholds - value)).argmin()
        return self.thresholds[idx]
