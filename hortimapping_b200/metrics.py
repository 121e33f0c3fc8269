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
    """metrics_3d/chamfer_distance.py:11-36: per sample the mean of the two directed mean nearest-neighbour distances
    (an empty prediction scores 0, :17-19); `compute()` averages over the samples seen since the last `reset()`."""

    def __init__(self):
        self.cd_array = []

    def update(self, gt, pt):
        if self.prediction_is_empty(pt):
            self.cd_array.append(0)
            return
        g, p = _points(gt), _points(pt)
        both = torch.stack([nn_distance(g, p).mean(), nn_distance(p, g).mean()])      # gt -> prediction, prediction -> gt
        self.cd_array.append(float(both.sum().item()) / 2)

    def reset(self):
        self.cd_array = []

    def compute(self):
        return float(np.sum(self.cd_array)) / len(self.cd_array)


class PrecisionRecall(Metrics3D):
    """metrics_3d/precision_recall.py:11-107: precision (prediction -> ground truth), recall (ground truth -> prediction) and
    F-score in percent at `num` distance thresholds.  One row of shape (3, num) is kept per sample; the threshold counts are
    taken on the device, only 2 x num integers come back per sample."""

    def __init__(self, min_t, max_t, num):
        self.thresholds = np.linspace(min_t, max_t, num)
        self._rows = []

    def update(self, gt, pt):
        if self.prediction_is_empty(pt):                           # :20-25: an empty prediction scores 0 everywhere
            self._rows.append(np.zeros((3, len(self.thresholds))))
            return
        g, p = _points(gt), _points(pt)
        d_p, d_g = nn_distance(p, g), nn_distance(g, p)
        thr = torch.from_numpy(self.thresholds).to(d_p.device)
        frac = [100 / len(d) * (d[:, None] < thr[None, :]).sum(0).cpu().numpy().astype(np.float64) for d in (d_p, d_g)]
        pr, re = frac
        with np.errstate(divide="ignore", invalid="ignore"):
            f1 = np.where((pr == 0) | (re == 0), 0.0, 2 * pr * re / (pr + re))      # :44-48
        self._rows.append(np.stack([pr, re, f1]))

    def reset(self):
        self._rows = []

    def _mean(self):
        return np.mean(np.stack(self._rows), axis=0)               # (3, num): mean over samples

    def _as_dict(self, which):
        rows = np.stack(self._rows) if self._rows else np.zeros((0, 3, len(self.thresholds)))
        return {t: rows[:, which, k].tolist() for k, t in enumerate(self.thresholds)}

    # the reference keeps three {threshold: [per-sample values]} dicts; offered read-only for host code that looks at them
    pr_dict = property(lambda self: self._as_dict(0))
    re_dict = property(lambda self: self._as_dict(1))
    f1_dict = property(lambda self: self._as_dict(2))

    def find_nearest_threshold(self, value):
        return self.thresholds[int(np.argmin(np.abs(self.thresholds - value)))]

    def compute_at_threshold(self, threshold):
        t = self.find_nearest_threshold(threshold)
        k = int(np.argmin(np.abs(self.thresholds - threshold)))
        pr, re, f1 = self._mean()[:, k]
        return float(pr), float(re), float(f1), t

    def compute_at_all_thresholds(self):
        m = self._mean()
        return m[0].tolist(), m[1].tolist(), m[2].tolist()

    def compute_auc(self):
        """Area under the three curves over the threshold range, normalised by the area of a perfect predictor (:69-90)."""
        from scipy.integrate import simpson
        dx = self.thresholds[1] - self.thresholds[0]
        unit = simpson(np.ones_like(self.thresholds), dx=dx)
        return tuple(float(simpson(c, dx=dx) / unit) for c in self._mean())
