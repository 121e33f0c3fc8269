"""N4: clean_pcd / get_pose_init (wild_completion/utils.py:389-459).

CPU: the sequential DBSCAN restatement (oracle/preprocess_oracle.py, open3d's algorithm) against scikit-learn's independent
implementation.  GPU: hm_dbscan / hm_cloud_bounds / hm_crop_mean_offset and the host mirrors against the oracle -- labels are
integer work and must be bit-identical, incl. noise, border points between two clusters, ties and degenerate sizes."""
import numpy as np
import pytest

from oracle import preprocess_oracle as PO


def clouds():
    g = np.random.default_rng(3)
    main = g.normal(0, 0.012, (1500, 3)) + [0.3, 0.1, 0.5]
    blob = g.normal(0, 0.004, (260, 3)) + [0.36, 0.1, 0.5]
    bridge = np.linspace([0.318, 0.1, 0.5], [0.35, 0.1, 0.5], 9) + g.normal(0, 0.0005, (9, 3))
    noise = g.uniform(-0.1, 0.1, (240, 3)) + [0.3, 0.1, 0.5]
    out = {"fruit_like": (np.concatenate([main, blob, bridge, noise]), 0.01, 40),
           "two_touching": (np.concatenate([g.normal(0, 0.003, (300, 3)), g.normal(0, 0.003, (300, 3)) + [0.013, 0, 0]]), 0.004, 12),
           "all_noise": (g.uniform(0, 1, (200, 3)), 0.01, 5),
           "min_points_0": (g.uniform(0, 0.05, (100, 3)), 0.01, 0),
           "single": (np.zeros((1, 3)), 0.01, 1),
           "duplicates": (np.repeat(g.uniform(0, 0.02, (40, 3)), 5, 0), 0.003, 6)}
    return out


@pytest.mark.parametrize("name", list(clouds()))
def test_oracle_dbscan_matches_sklearn(name):
    from sklearn.cluster import DBSCAN
    pts, eps, mp = clouds()[name]
    ours = PO.cluster_dbscan(pts, eps, mp)
    ref = DBSCAN(eps=np.nextafter(eps, 0), min_samples=max(mp, 1), algorithm="brute").fit(pts).labels_    # sklearn: <= eps, open3d: < eps
    np.testing.assert_array_equal(ours, ref)


def test_oracle_clean_pcd_and_pose_init():
    pts, eps, mp = clouds()["fruit_like"]
    keep = PO.clean_pcd(pts, 0.01, 0.02)
    assert 1300 < len(keep) < 1600             # the main blob: the small blob, the thin bridge and the scattered points are dropped
    assert (keep < 1500).mean() > 0.99
    g = np.random.default_rng(0)
    fruit = g.normal(0, 0.015, (2000, 3)) + [0.3, 0.1, 0.5]
    bg = np.concatenate([g.uniform(-0.5, 0.5, (5000, 3)) + [0.3, 0.1, 0.5], g.normal(0, 0.004, (400, 3)) + [0.33, 0.16, 0.56]])
    c, rot, size, valid = PO.get_pose_init(fruit, bg)
    assert valid and 0.03 < size < 0.16 and abs(rot) <= np.pi / 4 + 1e-12
    _, _, _, v2 = PO.get_pose_init(fruit * 10, bg)
    assert not v2


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(clouds()))
def test_device_dbscan_is_bit_identical_to_the_sequential_algorithm(name):
    from hortimapping_b200 import preprocess as PP
    pts, eps, mp = clouds()[name]
    np.testing.assert_array_equal(PP.dbscan_labels(pts, eps, mp), PO.cluster_dbscan(pts, eps, mp))


@pytest.mark.gpu
def test_device_dbscan_large_and_empty():
    from hortimapping_b200 import preprocess as PP
    g = np.random.default_rng(9)
    pts = np.concatenate([g.normal(0, 0.02, (6000, 3)), g.normal(0, 0.01, (3000, 3)) + [0.12, 0, 0], g.uniform(-0.3, 0.3, (1000, 3))])
    np.testing.assert_array_equal(PP.dbscan_labels(pts, 0.008, 30), PO.cluster_dbscan(pts, 0.008, 30))
    assert PP.dbscan_labels(np.zeros((0, 3)), 0.01, 3).shape == (0,)


class _Cloud:
    def __init__(self, p):
        self.points = np.asarray(p, np.float64)

    def select_by_index(self, idx):
        return _Cloud(self.points[np.asarray(idx, np.int64)])


@pytest.mark.gpu
def test_clean_pcd_and_get_pose_init_match_the_oracle():
    from hortimapping_b200 import preprocess as PP
    pts, _, _ = clouds()["fruit_like"]
    kept = PP.clean_pcd(_Cloud(pts), 0.01, 0.02)
    np.testing.assert_array_equal(kept.points, pts[PO.clean_pcd(pts, 0.01, 0.02)])
    g = np.random.default_rng(0)
    fruit = g.normal(0, 0.015, (2000, 3)) + [0.3, 0.1, 0.5]
    bg = np.concatenate([g.uniform(-0.5, 0.5, (50000, 3)) + [0.3, 0.1, 0.5], g.normal(0, 0.004, (400, 3)) + [0.33, 0.16, 0.56]])
    for f, b in ((fruit, bg), (fruit, bg[:20]), (fruit * 10, bg), (fruit * 0.1 + 0.3, bg)):
        c, rot, size, valid = PP.get_pose_init(_Cloud(f), _Cloud(b))
        c0, rot0, size0, valid0 = PO.get_pose_init(f, b)
        assert valid == valid0 and size == size0
        np.testing.assert_array_equal(c, c0)                       # min / max / midpoint: exact
        assert abs(rot - rot0) < 1e-12                             # mean of offsets: summation order differs
