"""Diagnostic (GPU box): where does the latent-only loop leave the fp64 oracle -- a wrong step, or amplification of rounding?

Fruit 0 of the bench workload, shape_opt_deepsdf (optimizer.py:306-429).  (1) Trajectories: 40 iterations with the tensor-core
engine, the fp32 CUDA-core engine, the numpy oracle in fp32 and in fp64: distance to the fp64 trajectory per iteration.
(2) Step replay: every iteration i is run ONCE on the device from the fp64 oracle's own state before it (iter_offset = i) and its
H, b, dx and new state are compared with the oracle's.  oracle/hm_oracle.py is the checker only.
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from hortimapping_b200 import synth  # noqa: E402
from hortimapping_b200.decoder import Decoder  # noqa: E402
from hortimapping_b200.optimizer import Optimizer  # noqa: E402

K = int(os.environ.get("DIAG_ITERS", "40"))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    from oracle import hm_oracle as O
    W, b, codes = B.load_weights()
    init = codes.mean(0).astype(np.float32)
    dec = Decoder(W, b, device=0)
    g = np.random.default_rng(0)
    from hortimapping_b200.decoder import calibration_rows
    dec.calibrate(calibration_rows(codes, 0.15))
    fr = synth.make_fruit(B.product_sdf_jac(dec), codes, int(os.environ.get("DIAG_SEED", "7")), int(os.environ.get("DIAG_FRUIT", "0")), n_pts=B.N_PTS, with_rays=False)
    pts = fr.points_w
    T_ow = np.linalg.inv(fr.T_wo_gt.astype(np.float64)).astype(np.float32)
    mm, _ = B._torch_mm()
    O.set_matmul(mm)
    cfg = copy.deepcopy(B.WILD_CFG)
    cfg["opt"]["converge"]["max_iter"] = K
    tr = {}
    for name, dt in (("o64", np.float64), ("o32", np.float32)):
        od = O.DecoderOracle(W, b, (4,), dt)
        t = O.OptTrace()
        O.shape_opt_deepsdf(od, cfg, init.astype(dt).copy(), T_ow.astype(dt), pts, trace=t)
        tr[name] = t
    opt = Optimizer(cfg, dec, None, None)
    traj = {}
    for name in ("tc", "simt"):
        dec.set_engine(name)
        out = []
        l = torch.from_numpy(init.copy()).cuda().reshape(1, 32)
        T = torch.from_numpy(T_ow.copy()).cuda().reshape(1, 4, 4)
        for i in range(K):                       # one iteration per call, continuing from the device's own state
            opt.shape_opt_deepsdf_batch(l, T, [pts], iter_offset=i, max_iter=1)
            out.append(l[0].cpu().numpy().copy())
        traj[name] = out
        # the same as ONE call of K iterations: must give the same bits
        l2 = torch.from_numpy(init.copy()).cuda().reshape(1, 32)
        opt.shape_opt_deepsdf_batch(l2, T, [pts], max_iter=K)
        print(f"{name}: K single-iteration calls == one K-iteration call: {np.array_equal(l2[0].cpu().numpy(), out[-1])}")
    print("trajectory: distance to the fp64 oracle's state after iteration i, and the size of the fp64 step")
    for i in range(K):
        t64 = tr["o64"].latent[i]
        print(f"  i={i:2d} |dx64| {np.abs(tr['o64'].dx[i]).max():.2e}  o32 {rel(tr['o32'].latent[i], t64):.2e}  tc {rel(traj['tc'][i], t64):.2e}  "
              f"simt {rel(traj['simt'][i], t64):.2e}")
    print("step replay from the fp64 oracle's state before iteration i: H, b, dx, state after")
    for name in ("tc", "simt"):
        dec.set_engine(name)
        worst = [0.0] * 4
        for i in range(K):
            s = (init if i == 0 else tr["o64"].latent[i - 1]).astype(np.float32)
            l = torch.from_numpy(s.copy()).cuda().reshape(1, 32)
            T = torch.from_numpy(T_ow.copy()).cuda().reshape(1, 4, 4)
            opt.shape_opt_deepsdf_batch(l, T, [pts], iter_offset=i, max_iter=1)
            H, bb, dx = (t.cpu().numpy()[0] for t in opt.last_system(1, joint=False))
            # the oracle's step from the SAME fp32-rounded state, in fp64
            od = O.DecoderOracle(W, b, (4,), np.float64)
            c1 = copy.deepcopy(cfg)
            c1["opt"]["converge"]["max_iter"] = 1
            t1 = O.OptTrace()
            l64 = s.astype(np.float64).copy()
            O.shape_opt_deepsdf(od, c1, l64, T_ow.astype(np.float64), pts, trace=t1, iter_offset=i)
            e = [rel(H, t1.H[0]), rel(bb, t1.b[0]), rel(dx, t1.dx[0]), rel(l[0].cpu().numpy(), l64)]
            worst = [max(a, c) for a, c in zip(worst, e)]
            if i < 3 or i % 5 == 0 or max(e[:2]) > 1e-4:
                print(f"  {name} i={i:2d}: H {e[0]:.2e} b {e[1]:.2e} dx {e[2]:.2e} state {e[3]:.2e}  |b| {np.abs(t1.b[0]).max():.2e}")
        print(f"  {name} worst over {K} iterations: H {worst[0]:.2e} b {worst[1]:.2e} dx {worst[2]:.2e} state {worst[3]:.2e}")
    dec.set_engine("tc")


if __name__ == "__main__":
    main()
