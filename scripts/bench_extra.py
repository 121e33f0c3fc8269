#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): the other rows of SURVEY.md section 8 on one B200, each next to
the CPU implementation of the same step.  Prints one JSON object per measurement; `scripts/gpu_round.sh` stores them under
gpurun_out/extra.log -> profiles/.

  grid      hm_sdf_grid 128^3 (BASELINE config 1 grid; fused voxel-grid + decoder forward)      tensor roofline
  iso       hm_isosurface on that grid (device marching tetrahedra) vs the numpy host extractor  HBM-bound integer work
  nn        hm_nn_distance 100k x 100k (Chamfer / precision-recall arithmetic) vs scipy cKDTree
  joint     shape_pose_joint_opt, wild_pepper.yaml sizes (10 frames x 400 rays x 30 samples + 2048 points), 32 fruits
"""
import copy
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B                                            # noqa: E402
from hortimapping_b200.decoder import Decoder                # noqa: E402
from hortimapping_b200.optimizer import Optimizer            # noqa: E402
from hortimapping_b200 import marching, metrics, synth       # noqa: E402

FLOP_FWD = 3_671_040


def cuda_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    which = set(sys.argv[1:]) or {"grid", "iso", "nn", "joint"}
    W, b, codes = B.load_weights()
    dec = Decoder(W, b, device=0)
    g = np.random.default_rng(0)
    cal = np.concatenate([codes[g.integers(0, codes.shape[0], 8192)], ((g.random((8192, 3)) * 2 - 1) * 0.15).astype(np.float32)], 1)
    dec.calibrate(torch.from_numpy(cal))
    lat = torch.from_numpy(codes.mean(0).astype(np.float32)).cuda()
    n = 128
    if "grid" in which or "iso" in which:
        ms = cuda_ms(lambda: dec.sdf_grid(lat, n, 0.08))
        sdf = dec.sdf_grid(lat, n, 0.08)
        print(json.dumps({"what": "hm_sdf_grid 128^3 (fused voxel grid + decoder forward)", "ms": ms, "rows": n ** 3,
                          "algorithmic_tflops": n ** 3 * FLOP_FWD / ms / 1e9, "bound": "tensor"}))
    if "iso" in which:
        ms = cuda_ms(lambda: dec.isosurface(sdf, 0.0, 2.0 / (n - 1), affine_radius=0.08))
        v, f = dec.isosurface(sdf, 0.0, 2.0 / (n - 1), affine_radius=0.08)
        vol = sdf.cpu().numpy()
        t0 = time.perf_counter()
        vr, fr = marching.marching_tetrahedra(vol, 0.0, (2.0 / (n - 1),) * 3)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        alg = n ** 3 * 4 + v.numel() * 4 + f.numel() * 4
        print(json.dumps({"what": "hm_isosurface 128^3 (count + 2 scans + emit + host sync)", "ms": ms, "verts": int(v.shape[0]), "faces": int(f.shape[0]),
                          "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6, "bound": "hbm (multi-pass + sync latency at this size)",
                          "cpu_numpy_ms": cpu_ms, "faces_equal_cpu": bool(np.array_equal(f.cpu().numpy(), fr))}))
    if "nn" in which:
        a = g.standard_normal((100000, 3)) * 0.04
        c = g.standard_normal((100000, 3)) * 0.04
        da, dc = torch.from_numpy(a).cuda(), torch.from_numpy(c).cuda()
        ms = cuda_ms(lambda: metrics.nn_distance(da, dc), reps=3, warm=1)
        from scipy.spatial import cKDTree
        t0 = time.perf_counter()
        ref = cKDTree(c).query(a, k=1)[0]
        cpu_ms = (time.perf_counter() - t0) * 1e3
        err = float(np.abs(metrics.nn_distance(da, dc).cpu().numpy() - ref).max())
        print(json.dumps({"what": "hm_nn_distance 100k x 100k (fp64 brute force)", "ms": ms, "pair_evals_per_s": 1e10 / ms * 1e3,
                          "fp64_gflops": 1e10 * 8 / ms / 1e6, "cpu_ckdtree_ms": cpu_ms, "max_abs_diff_vs_ckdtree": err}))
    if "joint" in which:
        def sdf_jac(latent, pts):
            y, gr = dec.sdf_jacobian(torch.from_numpy(np.asarray(latent, np.float32)), torch.from_numpy(np.asarray(pts, np.float32)))
            return y.reshape(-1).cpu().numpy(), gr.reshape(-1, 35)[:, 32:].cpu().numpy()
        fruits = [synth.make_fruit(sdf_jac, codes, 7, i, n_pts=2048, with_rays=True, leaf_fraction=0.2) for i in range(4)]
        n_f, iters = 32, 20
        cfg = copy.deepcopy(B.WILD_CFG)
        cfg["opt"]["converge"]["max_iter"] = iters
        opt = Optimizer(cfg, dec, None, None)
        rds = [fruits[i % 4].render_data for i in range(n_f)]
        pts = [fruits[i % 4].points_w for i in range(n_f)]
        lat0 = torch.from_numpy(np.tile(codes.mean(0).astype(np.float32), (n_f, 1))).cuda()
        T0 = torch.eye(4).repeat(n_f, 1, 1).cuda()

        def run():
            return opt.shape_pose_joint_opt_batch(lat0.clone(), T0.clone(), rds, pts, 0.08, pose_known=False)
        run()
        torch.cuda.synchronize()
        c0 = dec.counters()
        t0 = time.perf_counter()
        _, _, it, st = run()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        c1 = dec.counters()
        rf, rj = c1["rows_forward"] - c0["rows_forward"], c1["rows_jacobian"] - c0["rows_jacobian"]
        print(json.dumps({"what": f"shape_pose_joint_opt batch: {n_f} fruits x 10 frames x 400 rays x 30 samples + 2048 pts, {iters} LM iterations (host packing included)",
                          "s": dt, "fruits_per_s_at_200_iters": n_f / (dt * 200 / iters), "iterations_run": int(it.min().item()),
                          "iter_counts": it.cpu().tolist()[:8], "status": [hex(x) for x in st.cpu().tolist()[:8]],
                          "rows_forward": rf, "rows_jacobian_upper_bound": rj,
                          "launches": c1["kernel_launches"] - c0["kernel_launches"]}))


if __name__ == "__main__":
    main()
