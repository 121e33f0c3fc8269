#!/usr/bin/env python
"""Design aids for the tensor-core decoder kernel, run against the TEST-ONLY library (libhortimapping_b200_testing.so: the product
sources compiled with -DHM_TESTING; hortimapping_b200/_testing.py).  Not part of the product, not used by bench.py.

    python scripts/probe_decoder.py wait      per-role wait-cycle breakdown (producer / MMA issuer / epilogue) + ms per launch
    python scripts/probe_decoder.py trace     clock64 timeline of the first CTA pair -> gpurun_out/trace.npy (scripts/trace_view.py)
    python scripts/probe_decoder.py pair      cta_group::2 probe: where an M = 128 pair MMA puts its accumulator, and its cycle count
    python scripts/probe_decoder.py speed     ms per launch of the PRODUCT library, shortcut on / off, forward and forward+gradient
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hortimapping_b200 import _testing                      # noqa: E402
from tests.helpers import pepper_weights                    # noqa: E402

N = 131072


def make_rows(codes, n=N, seed=0):
    g = np.random.default_rng(seed)
    return np.concatenate([codes[g.integers(0, codes.shape[0], n)], ((g.random((n, 3)) * 2 - 1) * 0.05).astype(np.float32)], 1)


def calibrated(dec, codes):
    from hortimapping_b200.decoder import calibration_rows
    dec.calibrate(calibration_rows(codes, 0.15))
    return dec


def ms_of(fn, reps=5):
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
    what = sys.argv[1] if len(sys.argv) > 1 else "speed"
    W, b, codes = pepper_weights()
    t = torch.from_numpy(make_rows(codes)).cuda()
    if what == "speed":
        from hortimapping_b200.decoder import Decoder
        dec = calibrated(Decoder(W, b), codes)
        for on in (True, False):
            dec.set_sparse_plan(on)
            c0 = dec.counters()
            f, j = ms_of(lambda: dec._eval_rows(t, with_jac=False)), ms_of(lambda: dec._eval_rows(t, with_jac=True))
            c1 = dec.counters()
            print(f"sparse plan {'on ' if on else 'off'}: forward {f:.3f} ms ({N * 3.67104e6 / f / 1e9:.0f} TFLOP/s algorithmic)  forward+gradient {j:.3f} ms "
                  f"({N * 7.34208e6 / j / 1e9:.0f} TFLOP/s)  re-evaluated tiles {c1['tiles_redone_forward'] - c0['tiles_redone_forward']}/{c1['tiles_forward'] - c0['tiles_forward']} fwd, "
                  f"{c1['tiles_redone_jacobian'] - c0['tiles_redone_jacobian']}/{c1['tiles_jacobian'] - c0['tiles_jacobian']} jac")
        return
    if what == "sweep":                        # ms per decode call against the number of rows: fixed cost of a call vs cost per tile
        from hortimapping_b200.decoder import Decoder
        dec = calibrated(Decoder(W, b), codes)
        for n in (9472, 18944, 66304, 132608, 265216, 1060864):          # multiples of 148 x 64 rows: whole rounds of tiles
            tt = torch.from_numpy(make_rows(codes, n)).cuda()
            f, j = ms_of(lambda: dec._eval_rows(tt, with_jac=False)), ms_of(lambda: dec._eval_rows(tt, with_jac=True))
            print(f"rows {n:8d} ({n // 9472:3d} tiles per CTA): forward {f:.4f} ms  forward+gradient {j:.4f} ms")
        return
    dec = calibrated(_testing.testing_decoder(W, b), codes)
    L = _testing.lib()
    if what == "wait":
        out = (C.c_ulonglong * 13)()
        for jac in (True, False):
            dec._eval_rows(t, with_jac=jac)
            torch.cuda.synchronize()
            L.hm_debug_tc_wait_cycles(dec.handle, out)
            ms = ms_of(lambda: dec._eval_rows(t, with_jac=jac), reps=1)
            L.hm_debug_tc_wait_cycles(dec.handle, out)
            v = [int(x) for x in out]
            tot, nlead = max(v[4], 1), 74 * 2          # (two launches since the reset: warm-up + timed)
            for nm, o in (("leader", 5), ("peer", 9)):
                if v[o + 3]:
                    print("   epilogue[%s]: total %.0f  wait_full %.1f%%  promote %.1f%%  finalize %.1f%%" %
                          (nm, v[o + 3] / nlead, 100 * v[o] / v[o + 3], 100 * v[o + 1] / v[o + 3], 100 * v[o + 2] / v[o + 3]))
            print("jac", jac, "ms", round(ms, 3), "(instrumented build) per-CTA avg cycles: total %.0f  wait_A %.1f%%  wait_part %.1f%%  wait_W %.1f%%  "
                  "(producer wait_empty %.1f%%)" % (tot / nlead, 100 * v[1] / tot, 100 * v[2] / tot, 100 * v[3] / tot, 100 * v[0] / tot))
    elif what == "trace":
        # PROBE_LATENT=1: one latent for all rows + an xyz array (the optimisers' input mode) instead of full [n][35] rows
        if os.environ.get("PROBE_LATENT"):
            lat1, xyz1 = t[0, :32].contiguous(), t[:, 32:].contiguous()
            run = lambda: dec.sdf_jacobian(lat1, xyz1)
        else:
            run = lambda: dec._eval_rows(t, with_jac=not os.environ.get("PROBE_FWD"))     # PROBE_FWD=1: forward-only launch
        run()
        torch.cuda.synchronize()
        L.hm_debug_tc_trace(dec.handle, 1, None)
        run()
        torch.cuda.synchronize()
        out = np.zeros(3 * 8192 * 2, np.uint32)
        L.hm_debug_tc_trace(dec.handle, 0, out.ctypes.data)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.save(os.path.join(ROOT, "gpurun_out", os.environ.get("PROBE_OUT", "trace.npy")), out.reshape(3, 8192, 2))
        print("trace saved", [int((out.reshape(3, 8192, 2)[r, :, 0] != 0).sum()) for r in range(3)])
        busy = out.reshape(3, 8192, 2)[2, 4096:4096 + 148, 1].astype(np.int64)          # per-CTA busy cycles of the traced launch
        if busy.any():
            print("per-CTA busy cycles: min %d  mean %.0f  max %d   (leaders %s ...)" % (busy.min(), busy.mean(), busy.max(), busy[0:16:2].tolist()))
            np.save(os.path.join(ROOT, "gpurun_out", "cta_busy_" + os.environ.get("PROBE_OUT", "trace.npy")), busy)
    elif what == "pair":
        for m_rows, n_cols in ((64, 256), (128, 256), (64, 64)):
            A = np.zeros((2, m_rows, 64), np.float16)
            for c in range(2):
                for r in range(m_rows):
                    A[c, r, r % 64] = 1
            B = np.tile(np.arange(1, 65, dtype=np.float16)[None, :], (n_cols, 1))
            out = np.zeros((2, 128, 256), np.float32)
            cyc = (C.c_longlong * 2)()
            for reps in (1, 64):
                rc = L.hm_debug_tc_pair_probe(dec.handle, A.view(np.uint16).ctypes.data, B.view(np.uint16).ctypes.data, out.ctypes.data, cyc, m_rows, n_cols, reps)
                assert rc == 0, L.hm_last_error()
                print(f"M = {2 * m_rows} ({m_rows} rows per CTA) x N = {n_cols}: {reps} x 4 MMAs: issue {cyc[0]} cycles, complete {cyc[1]} cycles"
                      + (f" -> {(cyc[1]) / (4 * reps):.0f} cycles per MMA" if reps > 1 else ""))
            lanes = [int(np.isfinite(out[c]).any(1).sum()) for c in range(2)]
            cols = [int(np.isfinite(out[c]).any(0).sum()) for c in range(2)]
            print(f"   accumulator footprint per CTA: {lanes} TMEM lanes x {cols} columns")


if __name__ == "__main__":
    main()
