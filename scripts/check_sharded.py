#!/usr/bin/env python
"""L9 of the test ladder on real GPUs: the sharded sequence driver over N ranks (NCCL) returns, on every
rank, exactly what one rank computes for the whole list.  Run under torchrun:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sharded.py"""
import copy
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import cfg_of, load_npz, pepper_weights, render_data_of       # noqa: E402
from hortimapping_b200.decoder import Decoder                                     # noqa: E402
from hortimapping_b200.optimizer import Optimizer                                 # noqa: E402
from hortimapping_b200.shard import optimize_sharded                              # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    W, b, codes = pepper_weights()
    dec = Decoder(W, b, device=local)
    c = load_npz("fruit_wild")
    cfg = copy.deepcopy(cfg_of(c))
    cfg["device"] = "cuda"
    cfg["opt"]["converge"]["max_iter"] = 6
    for k in ("epsilon_g", "epsilon_c", "epsilon_t", "epsilon_r", "epsilon_s"):
        cfg["opt"]["converge"][k] = 0
    opt = Optimizer(cfg, dec, None, None)
    n = 7                                                            # ragged over 2 ranks
    g = np.random.default_rng(3)
    pts = [(c["points_w"] + g.normal(0, 2e-4, c["points_w"].shape)).astype(np.float32) for _ in range(n)]
    rds = [render_data_of(c)] * n
    lat0 = torch.from_numpy(np.tile(c["init_latent"], (n, 1)).astype(np.float32)).cuda()
    lat0 += 0.01 * torch.arange(n, device="cuda").reshape(-1, 1)
    T0 = torch.from_numpy(np.tile(c["init_T_ow"], (n, 1, 1)).astype(np.float32)).cuda()
    ok = True
    for rd in (None, rds):
        lat, T, it, st = optimize_sharded(opt, lat0, T0, pts, rd, cube_radius=float(c["cube_radius"]) if "cube_radius" in c else 0.08, pose_known=False)
        if rd is None:
            lr, Tr, ir, sr = opt.shape_opt_deepsdf_batch(lat0.clone(), T0.clone(), pts)
        else:
            lr, Tr, ir, sr = opt.shape_pose_joint_opt_batch(lat0.clone(), T0.clone(), rds, pts, float(c["cube_radius"]) if "cube_radius" in c else 0.08, False)
        same = torch.equal(lat, lr) and torch.equal(T, Tr.reshape(-1, 4, 4)) and torch.equal(it, ir.to(torch.int32)) and torch.equal(st, sr.to(torch.int32))
        print(f"rank {dist.get_rank()}: {'joint' if rd is not None else 'shape'} sharded == single-rank: {same}  iters {it.tolist()}", flush=True)
        ok &= bool(same)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
