"""Diagnostic (GPU box): per-tile time and fixed launch cost of the decoder kernel: t(k tiles per CTA) = a + b k."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import pepper_weights
from hortimapping_b200.decoder import Decoder, calibration_rows

W, b, codes = pepper_weights()
dec = Decoder(W, b)
dec.calibrate(calibration_rows(codes, 0.15))
g = np.random.default_rng(0)
lat = torch.from_numpy(codes.mean(0).astype(np.float32)).cuda()


def ms_of(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for sparse in (True, False):
    dec.set_sparse_plan(sparse)
    for jac in (False, True):
        ks, ts = [1, 2, 4, 8, 16, 32], []
        for k in ks:
            n = 148 * 64 * k
            x = torch.from_numpy(((g.random((n, 3)) * 2 - 1) * 0.05).astype(np.float32)).cuda()
            ts.append(ms_of((lambda: dec.sdf_jacobian(lat, x)) if jac else (lambda: dec.sdf(lat, x))))
        A = np.stack([np.ones(len(ks)), np.array(ks, float)], 1)
        (a, bb), *_ = np.linalg.lstsq(A, np.array(ts), rcond=None)
        print(f"sparse {sparse} jac {jac}: " + " ".join(f"k={k}:{t * 1e3:.0f}us" for k, t in zip(ks, ts)) + f"  -> fixed {a * 1e3:.1f} us, per tile {bb * 1e3:.2f} us = {bb * 1e-3 * 1.965e9 / 1e3:.1f} k cycles @1.965 GHz")
