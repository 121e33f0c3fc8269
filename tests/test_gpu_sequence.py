"""BASELINE.json configs[4] plumbing, end to end on the GPU box: the reference's UNCHANGED host script test_wild_completion.py
(staged under baseline/_ref by scripts/vendor_reference.py) run through `python -m hortimapping_b200.dropin` on a synthetic
sequence in the BUP20 on-disk layout (tests/synth_sequence.py), with the numpy stand-ins of tests/standins/ for open3d /
scikit-image / addict / plyfile / pyquaternion (none of them is installed here; SURVEY.md 8c last row).  Checks what the script
writes: one completed mesh, cleaned cloud and pose per fruit, poses and completed surfaces close to the ground truth."""
import os
import subprocess
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def test_unchanged_host_script_runs_on_a_synthetic_bup20_sequence(tmp_path):
    script = os.path.join(REF, "test_wild_completion.py")
    model_dir = os.path.join(REF, "deepsdf", "models", "sweetpepper_32")
    if not os.path.isfile(script) or not os.path.isdir(model_dir):
        pytest.skip("the unmodified reference is not staged under baseline/_ref (scripts/vendor_reference.py)")
    import torch
    from scipy.spatial import cKDTree
    from tests.gpu_helpers import pepper_decoder
    from tests.helpers import pepper_weights
    from tests.synth_sequence import make_sequence
    _, _, codes = pepper_weights()
    root = str(tmp_path / "seq")
    cfg_path, fruits = make_sequence(root, pepper_decoder(), codes, model_dir)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "standins"), ROOT]), HM_MESHER_ISO="device")
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-m", "hortimapping_b200.dropin", script, "-c", cfg_path], cwd=REF, env=env, capture_output=True,
                       text=True, timeout=900)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    sys.path.insert(0, os.path.join(ROOT, "tests", "standins"))
    import open3d as o3d
    gt = np.load(os.path.join(root, "gt.npz"))
    report = []
    for fid, T_gt in zip(gt["ids"], gt["T_wo"]):
        name = f"{int(fid):05d}_SweetPepper"
        pose_file = os.path.join(root, "submaps_pose", name + ".npy")
        assert os.path.isfile(pose_file), (name, r.stdout[-2000:])
        T = np.load(pose_file)
        mesh = o3d.io.read_triangle_mesh(os.path.join(root, "submaps_complete", name + ".ply"))
        clean = o3d.io.read_point_cloud(os.path.join(root, "submaps_clean", name + ".ply"))
        assert len(mesh.triangles) > 500 and 1000 < len(clean.points) <= 2000
        s, s_gt = np.linalg.det(T[:3, :3]) ** (1 / 3), np.linalg.det(T_gt[:3, :3]) ** (1 / 3)
        e_t = np.linalg.norm(T[:3, 3] - T_gt[:3, 3])
        surf = gt[f"surface_{int(fid)}"]
        d1 = cKDTree(surf).query(mesh.vertices)[0]
        d2 = cKDTree(mesh.vertices).query(surf)[0]
        chamfer = 0.5 * (d1.mean() + d2.mean())
        report.append((int(fid), float(s / s_gt), float(e_t), float(chamfer)))
    print(f"sequence of {len(report)} fruits completed in {wall:.1f} s wall (process start, model load and image IO included): "
          + "; ".join(f"id {i}: scale ratio {a:.3f}, |dt| {b * 1e3:.1f} mm, chamfer {c * 1e3:.2f} mm" for i, a, b, c in report))
    # Partial view (front half only), 30 LM iterations, 4 mm mesher voxels: the completed fruit must sit where the hidden truth is.
    # The Sim(3) scale itself is not identifiable (the latent space also encodes size: a larger shape at a smaller scale is the
    # same surface), so it is reported and only loosely bounded; the surface distance and the position carry the check.
    for fid, ratio, e_t, chamfer in report:
        assert 0.5 < ratio < 1.5 and e_t < 0.015 and chamfer < 0.008, report
