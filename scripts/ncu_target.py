#!/usr/bin/env python
"""Profiling target: every kernel of the library once, inside a cudaProfilerStart/Stop window (ncu --profile-from-start off).

  latent-only loop   64 fruits x 2048 points, 2 LM iterations   (tc_decoder_kernel<true>, normal_eq, solve, transform_points)
  joint loop         8 fruits x (10 frames x 400 rays x 30 samples + 2048 points), 2 iterations
                     (frame_setup, sample, tc_decoder_kernel<false>, composite, scan_blocks, scatter, ray_jacobian, point_jacobian, ...)
  mesher             hm_sdf_grid 128^3 (fused grid + forward decoder), hm_isosurface
  metrics / N1       hm_nn_distance 20k x 20k, get_render_data kernels on one synthetic frame
Used by scripts/gpu_round.sh; the per-kernel summary goes to profiles/ (scripts/summarise_ncu_kernels.py).
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B                                            # noqa: E402
from hortimapping_b200 import metrics, synth                 # noqa: E402
from hortimapping_b200.decoder import Decoder                # noqa: E402
from hortimapping_b200.optimizer import Optimizer            # noqa: E402


def main():
    W, b, codes = B.load_weights()
    dec = Decoder(W, b, device=0)
    g = np.random.default_rng(0)
    from hortimapping_b200.decoder import calibration_rows
    dec.calibrate(calibration_rows(codes, 0.15))
    cfg = copy.deepcopy(B.WILD_CFG)
    cfg["opt"]["converge"]["max_iter"] = 2
    opt = Optimizer(cfg, dec, None, None)
    n_f = 64
    pts = [((g.random((B.N_PTS, 3)) * 2 - 1) * 0.045).astype(np.float32) for _ in range(n_f)]
    lat0 = torch.from_numpy(np.tile(codes.mean(0).astype(np.float32), (n_f, 1))).cuda()
    T0 = torch.eye(4).repeat(n_f, 1, 1).cuda()
    fruits = [synth.make_fruit(B.product_sdf_jac(dec), codes, 7, i, n_pts=B.N_PTS, with_rays=True, leaf_fraction=0.2) for i in range(2)]
    rds, jpts = [fruits[i % 2].render_data for i in range(8)], [fruits[i % 2].points_w for i in range(8)]
    lat = torch.from_numpy(codes.mean(0).astype(np.float32)).cuda()
    qa, qb = torch.from_numpy(g.standard_normal((20000, 3)) * 0.04).cuda(), torch.from_numpy(g.standard_normal((20000, 3)) * 0.04).cuda()

    def region():
        opt.shape_opt_deepsdf_batch(lat0.clone(), T0.clone(), pts)
        opt.shape_pose_joint_opt_batch(lat0[:8].clone(), T0[:8].clone(), rds, jpts, 0.08, False)
        sdf = dec.sdf_grid(lat, 128, 0.08)
        dec.isosurface(sdf, 0.0, 2.0 / 127, affine_radius=0.08)
        metrics.nn_distance(qa, qb)

    region()                                   # warm-up (workspace allocation, lazy module load)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    region()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled region done")


if __name__ == "__main__":
    main()
