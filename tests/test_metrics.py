"""Evaluation metrics: the oracle restatement is self-checked on CPU; the device nearest-neighbour
kernel and the ChamferDistance / PrecisionRecall mirrors are checked against it on the GPU."""
import numpy as np
import pytest

from oracle import metrics_oracle as MO


def _clouds(seed, n_gt, n_pt, noise=0.002):
    g = np.random.default_rng(seed)
    d = g.standard_normal((n_gt, 3))
    gt = 0.04 * d / np.linalg.norm(d, axis=1, keepdims=True)
    d = g.standard_normal((n_pt, 3))
    pt = 0.04 * d / np.linalg.norm(d, axis=1, keepdims=True) + noise * g.standard_normal((n_pt, 3)) + np.array([0.001, 0, 0])
    return gt, pt


def test_oracle_kdtree_equals_bruteforce():
    for seed, (a, b) in enumerate([(1, 1), (7, 3), (500, 800), (1200, 900)]):
        gt, pt = _clouds(seed, a, b)
        np.testing.assert_allclose(MO.point_cloud_distance(pt, gt), MO.point_cloud_distance_bruteforce(pt, gt), rtol=1e-13, atol=1e-18)
    gt, pt = _clouds(9, 300, 200)
    assert MO.chamfer(gt, pt[:0]) == 0                                    # chamfer_distance.py:17-19
    assert MO.precision_recall(gt, pt[:0], np.linspace(1e-3, 1e-2, 5))[2] == [0] * 5
    assert MO.chamfer(gt, gt) == 0.0
    pr, re, f1 = MO.precision_recall(gt, gt, np.linspace(1e-3, 1e-2, 5))
    assert pr == [100.0] * 5 and re == [100.0] * 5 and f1 == [100.0] * 5


@pytest.mark.gpu
def test_device_nn_distance_matches_oracle():
    import torch
    from hortimapping_b200.metrics import nn_distance
    for seed, (a, b) in enumerate([(1, 1), (3, 1), (1, 5), (255, 1025), (257, 1023), (5000, 20000), (20000, 3000)]):
        gt, pt = _clouds(seed, a, b)
        d = nn_distance(pt, gt).cpu().numpy()
        np.testing.assert_allclose(d, MO.point_cloud_distance(pt, gt), rtol=1e-12, atol=1e-17)
    gt, pt = _clouds(3, 1000, 1000)
    pt[::7] = gt[::7]                                                      # exact duplicates -> distance 0
    d = nn_distance(pt, gt).cpu().numpy()
    assert (d[::7] == 0).all()
    assert nn_distance(np.zeros((0, 3)), gt).shape == (0,)
    f32 = torch.from_numpy(pt.astype(np.float32)).cuda()                   # float32 CUDA tensors are widened exactly
    np.testing.assert_allclose(nn_distance(f32, gt).cpu().numpy(), MO.point_cloud_distance(pt.astype(np.float32), gt), rtol=1e-12, atol=1e-17)


@pytest.mark.gpu
def test_metric_classes_match_oracle_bookkeeping():
    from hortimapping_b200.metrics import ChamferDistance, PrecisionRecall
    cd, pr = ChamferDistance(), PrecisionRecall(min_t=0.001, max_t=0.01, num=100)      # run_shape_completion_challenge.py:82-83
    cds, prs = [], []
    for seed in range(3):
        gt, pt = _clouds(seed, 4000 + 100 * seed, 3000, noise=0.001 * (seed + 1))
        cd.update(gt, pt)
        pr.update(gt, pt)
        cds.append(MO.chamfer(gt, pt))
        prs.append(MO.precision_recall(gt, pt, pr.thresholds))
    gt, _ = _clouds(5, 100, 10)
    cd.update(gt, np.zeros((0, 3)))                                        # empty prediction (chamfer_distance.py:17-19)
    pr.update(gt, np.zeros((0, 3)))
    cds.append(0)
    prs.append(([0] * 100,) * 3)
    assert abs(cd.compute() - sum(cds) / 4) <= 1e-15
    p_all, r_all, f_all = pr.compute_at_all_thresholds()
    np.testing.assert_allclose(p_all, np.mean([x[0] for x in prs], 0), rtol=1e-12)
    np.testing.assert_allclose(r_all, np.mean([x[1] for x in prs], 0), rtol=1e-12)
    np.testing.assert_allclose(f_all, np.mean([x[2] for x in prs], 0), rtol=1e-12)
    p5, r5, f5, t5 = pr.compute_at_threshold(0.005)
    k = int(np.abs(pr.thresholds - 0.005).argmin())
    assert t5 == pr.thresholds[k] and p5 == p_all[k] and r5 == r_all[k] and f5 == f_all[k]
    auc = pr.compute_auc()
    assert all(0 <= a <= 100 for a in auc)
    cd.reset()
    pr.reset()
    assert cd.cd_array == [] and all(v == [] for v in pr.pr_dict.values())
