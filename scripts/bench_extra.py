#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): the other rows of SURVEY.md section 8 on one B200, each next to
the CPU implementation of the same step.  Prints one JSON object per measurement; `scripts/gpu_round.sh` stores them under
gpurun_out/extra.log -> profiles/.

  grid      hm_sdf_grid 128^3 (BASELINE config 1 grid; fused voxel-grid + decoder forward)      tensor roofline
  iso       hm_isosurface on that grid (device marching tetrahedra) vs the numpy host extractor  HBM-bound integer work
  nn        hm_nn_distance 100k x 100k (Chamfer / precision-recall arithmetic) vs scipy cKDTree
  render_data  get_render_data for 20 fruits x 10 frames of 720 x 1280 images vs the numpy restatement of the reference
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
    which = set(sys.argv[1:]) or {"grid", "iso", "nn", "joint", "render_data"}
    W, b, codes = B.load_weights()
    dec = Decoder(W, b, device=0)
    g = np.random.default_rng(0)
    from hortimapping_b200.decoder import calibration_rows
    dec.calibrate(calibration_rows(codes, 0.15))
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
    if "render_data" in which:
        from hortimapping_b200 import render_data as RD
        from oracle import render_data_oracle as RO           # CPU baseline leg only (never on the product path)
        Hh, Ww, n_fr, n_id = 720, 1280, 10, 20
        vv, uu = np.mgrid[0:Hh, 0:Ww]
        id_imgs, depth_imgs, poses = {}, {}, {}
        for k in range(n_fr):
            img = np.zeros((Hh, Ww), np.int32)
            depth = (0.5 + 0.0001 * ((vv * 3 + uu * 5 + k) % 64)).astype(np.float32)
            for i in range(1, n_id + 1):
                cv, cu = 90 + 130 * ((i - 1) // 5) + 3 * k, 130 + 250 * ((i - 1) % 5) - 2 * k
                m = ((vv - cv) / 45.0) ** 2 + ((uu - cu) / 38.0) ** 2 <= 1.0
                img[m] = i
                depth[m] = np.float32(0.3 + 0.005 * i)
            depth[(vv * 7 + uu * 13 + k * 5) % 17 == 0] = 0.0
            id_imgs[k], depth_imgs[k], poses[k] = img, depth, np.eye(4)
        invK = np.linalg.inv(np.array([[900.0, 0, 640], [0, 900.0, 360], [0, 0, 1]]))
        cfg = {"device": "cuda", "opt": {"render": B.WILD_CFG["opt"]["render"]}}

        def run_all(fn):
            np.random.seed(0)
            return [fn(i, id_imgs, depth_imgs, poses, (Hh, Ww), invK, cfg) for i in range(1, n_id + 1)]
        run_all(RD.get_render_data)                             # uploads + per-frame id tables (cached per image object)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev_out = run_all(RD.get_render_data)
        torch.cuda.synchronize()
        t_dev = time.perf_counter() - t0
        RD._frames.clear()
        t0 = time.perf_counter()
        dev_cold = run_all(RD.get_render_data)
        torch.cuda.synchronize()
        t_cold = time.perf_counter() - t0
        t0 = time.perf_counter()
        cpu_out = run_all(RO.get_render_data)
        t_cpu = time.perf_counter() - t0
        same = all(np.array_equal(a["rays_fg"][j].cpu().numpy(), b2["rays_fg"][j]) and np.array_equal(a["depth_bg"][j].cpu().numpy(), b2["depth_bg"][j])
                   for a, b2 in zip(dev_out, cpu_out) for j in range(a["count"]))
        print(json.dumps({"what": f"get_render_data: {n_id} fruits x {n_fr} frames of {Hh}x{Ww} (ids + depth)", "device_s_images_resident": t_dev,
                          "device_s_including_upload_and_id_tables": t_cold, "cpu_numpy_s": t_cpu, "frames_found": sum(a["count"] for a in dev_out),
                          "bit_identical_to_cpu": bool(same)}))
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
