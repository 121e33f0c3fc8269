"""Diagnostic (GPU box): how many tiles leave the sparse plan during the bench's latent-only loop, per calibration set, and what it costs."""
import copy, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
from hortimapping_b200.decoder import Decoder
from hortimapping_b200.optimizer import Optimizer, PackedBatch, opt_params_from_cfg


def cal_rows(codes, kind, n, g):
    z = codes[g.integers(0, codes.shape[0], n)].astype(np.float32)
    mean = codes.mean(0).astype(np.float32)
    if "interp" in kind:
        t = g.random((n, 1)).astype(np.float32)
        z = mean + t * (z - mean)
    if "jitter" in kind:
        z = z + (0.03 * g.standard_normal(z.shape)).astype(np.float32)
    x = ((g.random((n, 3)) * 2 - 1) * 0.15).astype(np.float32)
    return np.concatenate([z, x], 1)


def main():
    W, b, codes = B.load_weights()
    dec = Decoder(W, b, device=0)
    g = np.random.default_rng(0)
    dec.calibrate(torch.from_numpy(cal_rows(codes, "codes", 65536, g)))
    pts, T_ow, init_lat = B.make_inputs(dec, codes, 0, 0)
    cfg = copy.deepcopy(B.WILD_CFG)
    opt = Optimizer(cfg, dec, None, None)
    pk = PackedBatch([p for p in pts], None, 0, np.zeros(B.N_FRUITS, np.float32), np.zeros(B.N_FRUITS, bool))
    params = opt_params_from_cfg(cfg["opt"])
    for kind, n in (("codes", 65536), ("codes+jitter", 65536), ("codes+interp", 65536), ("codes+interp+jitter", 262144), ("off", 0)):
        g = np.random.default_rng(0)
        if kind == "off":
            dec.set_sparse_plan(False)
        else:
            dec.calibrate(torch.from_numpy(cal_rows(codes, kind, n, g)))
        res = []
        for rep in range(2):
            lat, T = torch.from_numpy(init_lat).cuda(), torch.from_numpy(T_ow).cuda()
            c0 = dec.counters()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            opt._run(pk, lat, T, params)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            c1 = dec.counters()
            res.append((dt, c1["tiles_redone_jacobian"] - c0["tiles_redone_jacobian"], c1["tiles_jacobian"] - c0["tiles_jacobian"]))
        pi = dec.plan_info()
        print(f"{kind:22s} n={n:6d}: step {res[-1][0] * 1e3:7.1f} ms  redone {res[-1][1]}/{res[-1][2]} tiles  alive chunks {pi['alive_chunks_per_layer']}  "
              f"issued MFLOP/row {pi['issued_flop_per_row']['sparse_jacobian'] / 1e6:.2f}", flush=True)


if __name__ == "__main__":
    main()
