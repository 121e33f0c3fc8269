"""GPU parity tests of the decoder engines (call through the C ABI).  Run with -m gpu on the B200 box.

Tolerance: BASELINE.json north_star asks for 1e-4 fp32 relative; SDF values cross zero (the surface),
so the tests use |err| <= 1e-4 * |ref| + 1e-6 (1 micrometre absolute) on SDF values and the same rtol
with a 2e-6 floor on Jacobian entries (|entries| <= ~1).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import hm_oracle as O
from tests.helpers import cfg_of, load_npz, oracle_decoder

pytestmark = pytest.mark.gpu

SDF_TOL = dict(rtol=1e-4, atol=1e-6)
JAC_TOL = dict(rtol=1e-4, atol=2e-6)


def assert_jac_close(actual, ref, rtol=1e-4, atol=2e-6, max_bad_rows=3e-3, what=""):
    """Row-wise Jacobian comparison.  d sdf / d input of a ReLU network is piecewise constant in the
    activation pattern: a hidden unit whose pre-activation is within fp32 rounding of zero (measured: the
    closest of 4096 units x 1000 rows sits at 8e-8 of its operand scale) switches sides between two
    correct fp32 evaluations and changes that row's whole gradient by O(10 %).  The reference has the same
    property between its CPU and CUDA runs.  So: all rows but a vanishing fraction must agree to the stated
    tolerance; the rest are reported."""
    actual = np.asarray(actual).reshape(-1, 35)
    ref = np.asarray(ref).reshape(-1, 35)
    bad = (np.abs(actual - ref) > rtol * np.abs(ref) + atol).any(1)
    allowed = max(1, int(np.ceil(max_bad_rows * len(bad))))
    assert bad.sum() <= allowed, f"{what}: {bad.sum()} of {len(bad)} rows off (allowed {allowed}); max abs err " \
                                 f"{np.abs(actual - ref).max():.3e}"
    good = ~bad
    if good.any():
        np.testing.assert_allclose(actual[good], ref[good], rtol=rtol, atol=atol)


def test_tcgen05_selftest_layouts():
    """One 64x128x64 fp16 GEMM through the kernel's descriptor / TMEM-layout building blocks."""
    from hortimapping_b200 import _lib, _testing
    from tests.helpers import pepper_weights
    W, b, _ = pepper_weights()
    dec = _testing.testing_decoder(W, b)          # the probes live in the test-only library, not in the product
    g = np.random.default_rng(3)
    A = g.integers(-4, 5, (64, 64)).astype(np.float16)
    B = g.integers(-4, 5, (128, 64)).astype(np.float16)
    ref = A.astype(np.float32) @ B.astype(np.float32).T          # exact in fp32
    L = _testing.lib()
    for lane_off, col_off in ((0, 0), (16, 0), (0, 128), (16, 128)):
        out = np.zeros((128, 256), np.float32)
        _lib.check(L.hm_debug_tc_selftest(dec.handle, A.view(np.uint16).ctypes.data, B.view(np.uint16).ctypes.data,
                                          out.ctypes.data, lane_off, col_off, 1), "selftest")
        # M = 64 accumulator layout: row r lives in TMEM lane 32*(r/16) + r%16 (+ lane_off)
        lanes = np.array([32 * (r // 16) + r % 16 + lane_off for r in range(64)])
        got = out[lanes, col_off:col_off + 128]
        np.testing.assert_array_equal(got, ref, err_msg=f"lane_off={lane_off} col_off={col_off}")
        mask = np.ones((128, 256), bool)
        mask[np.ix_(lanes, np.arange(col_off, col_off + 128))] = False
        assert np.all(out[mask] == 0), "MMA wrote outside its accumulator window"


@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_decoder_rows_vs_reference_golden(engine):
    from tests.gpu_helpers import pepper_decoder
    dec = pepper_decoder()
    dec.set_engine(engine)
    try:
        g = load_npz("decoder_rows")
        rows = torch.from_numpy(g["rows"]).cuda()
        y2 = dec(rows)
        assert tuple(y2.shape) == g["sdf2d"].shape
        np.testing.assert_allclose(y2.cpu().numpy(), g["sdf2d"], **SDF_TOL)
        y3 = dec(rows.unsqueeze(1))
        assert tuple(y3.shape) == g["sdf3d"].shape
        np.testing.assert_allclose(y3.cpu().numpy(), g["sdf3d"], **SDF_TOL)
        # autograd path used by the reference's get_gradient (utils.py:112-122)
        inp = rows.unsqueeze(1).clone().requires_grad_(True)
        y = dec(inp)
        (gr,) = torch.autograd.grad(y, inp, torch.ones_like(y))
        assert_jac_close(gr.cpu().numpy(), g["jac"], what="autograd")
        # fp64 reference: the device result must be as close to the truth as the fp32 reference is (x4 slack)
        e_dev = np.abs(y.detach().cpu().numpy().reshape(-1) - g["sdf64"].reshape(-1)).max()
        e_ref = np.abs(g["sdf3d"].reshape(-1) - g["sdf64"].reshape(-1)).max()
        assert e_dev < 4 * e_ref + 5e-8, (e_dev, e_ref)
        lat, xyz = torch.from_numpy(g["lat"]).cuda(), torch.from_numpy(g["xyz"]).cuda()
        from hortimapping_b200.decoder import decode_sdf, get_batch_sdf_jacobian
        np.testing.assert_allclose(decode_sdf(dec, lat, xyz).cpu().numpy(), g["decode_sdf"], **SDF_TOL)
        yb, gb = get_batch_sdf_jacobian(dec, lat, xyz)
        assert tuple(yb.shape) == g["batch_y"].shape and tuple(gb.shape) == g["batch_g"].shape
        np.testing.assert_allclose(yb.cpu().numpy(), g["batch_y"], **SDF_TOL)
        assert_jac_close(gb.cpu().numpy(), g["batch_g"], what="batch")
    finally:
        dec.set_engine("tc")


@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_alive_random_decoder_vs_oracle(engine):
    """The shipped checkpoints have a dead lin3 (SURVEY probe), so layers 0-3 are pinned with a random
    decoder whose layers are all alive, against the fp64 oracle."""
    from tests.gpu_helpers import random_decoder, random_rows
    dec, W, b = random_decoder(1)
    rows = random_rows(4096, seed=5)
    dec.calibrate(torch.from_numpy(random_rows(4096, seed=6)))
    dec.set_engine(engine)
    try:
        orc = O.DecoderOracle(W, b, (4,), np.float64)
        y_ref, j_ref = orc.forward_jac(rows.astype(np.float64))
        masks_alive = [(np.maximum(rows.astype(np.float64) @ W[0].T.astype(np.float64) + b[0], 0) > 0).mean()]
        assert masks_alive[0] > 0.2
        t = torch.from_numpy(rows).cuda().requires_grad_(True)
        y = dec(t)
        (gr,) = torch.autograd.grad(y, t, torch.ones_like(y))
        scale = np.abs(j_ref).max()
        np.testing.assert_allclose(y.detach().cpu().numpy(), y_ref, rtol=1e-4, atol=2e-6)
        # numpy fp32 vs fp64 of the SAME oracle already flips 3 of these 4096 rows (all 4096 hidden units alive,
        # larger weights than the shipped model); the engines' operand rounding (2^-22 for the fp16 split)
        # is ~4x fp32's 2^-24, hence ~4x the flips.
        assert_jac_close(gr.cpu().numpy(), j_ref, rtol=1e-4, atol=2e-6 * max(scale, 1.0), max_bad_rows=1e-2, what="alive")
    finally:
        dec.set_engine("tc")


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 127, 129, 1000])
def test_ragged_row_counts(n):
    from tests.gpu_helpers import pepper_decoder, random_rows
    dec = pepper_decoder()
    rows = random_rows(max(n, 1), seed=n)[:n]
    t = torch.from_numpy(rows).cuda().reshape(n, 35)
    y = dec(t)
    assert tuple(y.shape) == (n, 1)
    if n:
        ref = oracle_decoder().forward(rows)
        np.testing.assert_allclose(y.cpu().numpy(), ref, **SDF_TOL)
        _, gr = dec._eval_rows(t, with_jac=True)
        _, jref = oracle_decoder().forward_jac(rows)
        assert_jac_close(gr.cpu().numpy(), jref, what=f"n={n}")


def test_voxel_grid_bit_exact_and_grid_sdf():
    from tests.gpu_helpers import pepper_decoder
    dec = pepper_decoder()
    m = load_npz("misc")
    for n in (8, 20):
        np.testing.assert_array_equal(dec.voxel_grid(n, 1.0).cpu().numpy(), m[f"grid_{n}"])
    g40 = dec.voxel_grid(40, 1.0).cpu().numpy()
    np.testing.assert_array_equal(g40[m["grid_40_sample_idx"]], m["grid_40_sample"])
    np.testing.assert_array_equal(dec.voxel_grid(20, 0.08).cpu().numpy(), m["grid_20"] * np.float32(0.08))
    sdf = dec.sdf_grid(torch.from_numpy(m["grid_lat"]).cuda(), 20, 0.08)
    assert tuple(sdf.shape) == (20, 20, 20)
    np.testing.assert_allclose(sdf.cpu().numpy(), m["grid_sdf_20"], **SDF_TOL)


def test_tc_and_simt_engines_agree_on_a_large_batch():
    """Size-independent property at full tile counts (several waves of the persistent grid): both device
    engines agree, and the result does not depend on how rows are grouped into tiles."""
    from tests.gpu_helpers import pepper_decoder, random_rows
    dec = pepper_decoder()
    _, _, codes = __import__("tests.helpers", fromlist=["pepper_weights"]).pepper_weights()
    n = 148 * 64 * 3 + 17
    g = np.random.default_rng(11)
    rows = np.concatenate([codes[g.integers(0, codes.shape[0], n)], ((g.random((n, 3)) * 2 - 1) * 0.08).astype(np.float32)], 1)
    t = torch.from_numpy(rows).cuda()
    y_tc, j_tc = dec._eval_rows(t, with_jac=True)
    dec.set_engine("simt")
    try:
        y_s, j_s = dec._eval_rows(t, with_jac=True)
    finally:
        dec.set_engine("tc")
    np.testing.assert_allclose(y_tc.cpu().numpy(), y_s.cpu().numpy(), **SDF_TOL)
    assert_jac_close(j_tc.cpu().numpy(), j_s.cpu().numpy(), what="tc vs simt")
    perm = torch.randperm(n, device="cuda")
    y_p, _ = dec._eval_rows(t[perm], with_jac=False)
    assert torch.equal(y_p, y_tc[perm]), "a row's SDF must not depend on its tile neighbours"


def test_fp16_saturation_is_reported_not_silent():
    """The tensor-core engine scales operands into the fp16 range with calibrated powers of two (64x headroom).  Rows far
    outside the calibration set saturate the conversion: the library must say so (hm_saturation_count, HM_STATUS_F16_SATURATED)
    instead of returning silently degraded values, and a calibration on representative rows must clear the condition."""
    import warnings
    from hortimapping_b200 import _lib
    from hortimapping_b200.decoder import Decoder
    from hortimapping_b200.optimizer import Optimizer
    from tests.helpers import random_decoder_weights
    W, b = random_decoder_weights(5)
    dec = Decoder(W, b)
    g = np.random.default_rng(2)
    small = np.concatenate([g.normal(0, 0.01, (4096, 32)), (g.random((4096, 3)) * 2 - 1) * 1e-3], 1).astype(np.float32)
    big = np.concatenate([g.normal(0, 0.1, (4096, 32)), (g.random((4096, 3)) * 2 - 1) * 5.0], 1).astype(np.float32)
    dec.calibrate(torch.from_numpy(small))
    dec.saturation_count()
    dec._eval_rows(torch.from_numpy(small).cuda(), with_jac=True)
    assert dec.saturation_count() == 0
    dec._eval_rows(torch.from_numpy(big).cuda(), with_jac=True)
    assert dec.saturation_count() > 0
    assert dec.saturation_count() == 0                                   # reading resets
    cfg = cfg_of(load_npz("fruit_wild"))
    cfg["device"] = "cuda"
    cfg["opt"]["converge"]["max_iter"] = 2
    opt = Optimizer(cfg, dec, None, None)
    lat = torch.zeros(2, 32, device="cuda")
    T = torch.eye(4, device="cuda").repeat(2, 1, 1)
    pts = [big[:300, 32:], big[300:700, 32:]]
    _, _, _, st = opt.shape_opt_deepsdf_batch(lat, T, pts)
    assert all(int(s) & _lib.STATUS["F16_SATURATED"] for s in st.cpu().tolist())
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        opt._report(st.cpu().numpy())
    assert any("F16_SATURATED" in str(w.message) for w in wlist)
    assert dec.saturation_count() > 0                                    # the optimiser's launches count too; reading resets
    # calibrated on rows that represent what is evaluated: clean again, and fp32-grade
    from tests.gpu_helpers import random_rows
    mid, mid_cal = random_rows(4096, seed=5), random_rows(4096, seed=6)
    dec.calibrate(torch.from_numpy(mid_cal))
    y, _ = dec._eval_rows(torch.from_numpy(mid).cuda(), with_jac=False)
    assert dec.saturation_count() == 0
    y_ref = O.DecoderOracle(W, b, (4,), np.float64).forward(mid.astype(np.float64))
    np.testing.assert_allclose(y.cpu().numpy().reshape(-1), y_ref.reshape(-1), rtol=1e-4, atol=2e-6)


def _counters_delta(dec, fn):
    c0 = dec.counters()
    out = fn()
    c1 = dec.counters()
    return out, {k: c1[k] - c0[k] for k in c1}


@pytest.mark.parametrize("n", [64, 129, 9000, 148 * 64 + 33])
def test_sparse_plan_is_taken_and_bit_identical(n):
    """Both shipped models are very sparse after their ReLUs (lin3 dead outright, 12 .. 320 of 512 units ever alive elsewhere).
    The tensor-core engine orders the hidden units alive-first (from the calibration rows) and drops the MMAs on all-zero
    64-wide activation chunks; every tile's ReLU bits are checked against that assumption.  The dropped products are exact
    zeros, so the results must be BIT-identical with the plan switched off; on rows like the calibration rows no tile may need
    the full plan; the row / tile counters are exact."""
    from tests.gpu_helpers import pepper_decoder
    dec = pepper_decoder()
    _, _, codes = __import__("tests.helpers", fromlist=["pepper_weights"]).pepper_weights()
    g = np.random.default_rng(n)
    rows = np.concatenate([codes[g.integers(0, codes.shape[0], n)], ((g.random((n, 3)) * 2 - 1) * 0.08).astype(np.float32)], 1)
    t = torch.from_numpy(rows).cuda()
    tiles = (n + 63) // 64
    try:
        dec.set_sparse_plan(True)
        (y1, j1), d1 = _counters_delta(dec, lambda: dec._eval_rows(t, with_jac=True))
        (f1, _), e1 = _counters_delta(dec, lambda: dec._eval_rows(t, with_jac=False))
        dec.set_sparse_plan(False)
        (y0, j0), d0 = _counters_delta(dec, lambda: dec._eval_rows(t, with_jac=True))
        (f0, _), e0 = _counters_delta(dec, lambda: dec._eval_rows(t, with_jac=False))
    finally:
        dec.set_sparse_plan(True)
    assert torch.equal(y1, y0) and torch.equal(j1, j0) and torch.equal(f1, f0) and torch.equal(f1, y1)
    assert d1["rows_jacobian"] == n and d1["rows_forward"] == 0 and e1["rows_forward"] == n and e1["rows_jacobian"] == 0
    assert d1["tiles_jacobian"] == tiles and e1["tiles_forward"] == tiles
    # rows drawn like the calibration rows: (almost) every tile stays on the sparse plan -- a unit that calibration never saw alive
    # may still fire on a new row, that tile then takes the full plan
    assert d1["tiles_redone_jacobian"] <= max(1, tiles // 10) and e1["tiles_redone_forward"] == d1["tiles_redone_jacobian"], (d1, e1)
    assert d0["tiles_redone_jacobian"] == 0 and e0["tiles_redone_forward"] == 0
    assert d1["kernel_launches"] == 2 and d0["kernel_launches"] == 1              # sparse pass + (empty) redo pass vs one full pass
    y_ref, j_ref = oracle_decoder().forward_jac(rows)
    np.testing.assert_allclose(y1.cpu().numpy(), y_ref, **SDF_TOL)
    assert_jac_close(j1.cpu().numpy(), j_ref, what="sparse plan")


def test_tiles_that_contradict_the_sparse_plan_are_re_evaluated():
    """The plan's assumptions are checked, not trusted.  (a) A decoder whose lin3 is alive for SOME rows only, calibrated on rows
    where it is dead: rows are ordered by the pre-activation of the one revived unit, so the first 64-row tiles pass the checks,
    the last ones fail them and one tile straddles the boundary; exactly the tiles that contain a live row must go through the
    full plan, and the outputs must equal the plan-off run bit for bit and match the oracle.  (b) The shipped decoder calibrated
    on a tiny neighbourhood and then evaluated far outside it (units alive that calibration never saw)."""
    from hortimapping_b200.decoder import Decoder
    from tests.helpers import pepper_weights
    W, b, codes = pepper_weights()
    W, b = [w.copy() for w in W], [x.copy() for x in b]
    n = 64 * 40
    g = np.random.default_rng(77)
    rows = np.concatenate([codes[g.integers(0, codes.shape[0], n)], ((g.random((n, 3)) * 2 - 1) * 0.06).astype(np.float32)], 1).astype(np.float32)
    orc = O.DecoderOracle(W, b, (4,), np.float64)
    h = rows.astype(np.float64)
    for l in range(3):
        h = np.maximum(h @ orc.weights[l].T + orc.biases[l], 0)
    pre = h @ orc.weights[3].T + orc.biases[3]
    assert (pre > 0).sum() == 0                                  # the shipped lin3 is dead on these rows
    j = int(np.argmax(pre.max(0)))                               # revive the unit closest to zero for the upper ~40 % of the rows
    shift = np.float32(-np.quantile(pre[:, j], 0.6))
    b[3][j] += shift
    order = np.argsort(pre[:, j], kind="stable")
    rows = rows[order]
    alive_rows = (pre[order, j] + np.float64(shift)) > 0
    assert 0.3 < alive_rows.mean() < 0.5
    dec = Decoder(W, b)
    dec.calibrate(torch.from_numpy(rows[:64 * 16]))              # calibration sees dead rows only -> the plan assumes a dead lin3
    t = torch.from_numpy(rows).cuda()
    (y1, j1), d1 = _counters_delta(dec, lambda: dec._eval_rows(t, with_jac=True))
    (f1, _), e1 = _counters_delta(dec, lambda: dec._eval_rows(t, with_jac=False))
    dec.set_sparse_plan(False)
    y0, j0 = dec._eval_rows(t, with_jac=True)
    assert torch.equal(y1, y0) and torch.equal(j1, j0) and torch.equal(f1, y1)
    tiles_alive = int(alive_rows.reshape(-1, 64).any(1).sum())
    assert d1["tiles_redone_jacobian"] >= tiles_alive and e1["tiles_redone_forward"] >= tiles_alive, (d1, e1, tiles_alive)
    assert d1["tiles_redone_jacobian"] < d1["tiles_jacobian"]     # ... and the dead tiles stay on the sparse plan
    orc2 = O.DecoderOracle(W, b, (4,), np.float64)
    y_ref, j_ref = orc2.forward_jac(rows.astype(np.float64))
    np.testing.assert_allclose(y1.cpu().numpy(), y_ref, **SDF_TOL)
    assert_jac_close(j1.cpu().numpy(), j_ref, max_bad_rows=5e-3, what="mixed")
    # (b) shipped weights, calibration far too narrow
    W, b, codes = pepper_weights()
    dec = Decoder(W, b)
    near = np.concatenate([np.tile(codes[3], (2048, 1)), ((g.random((2048, 3)) * 2 - 1) * 0.002).astype(np.float32)], 1).astype(np.float32)
    dec.calibrate(torch.from_numpy(near))
    far = np.concatenate([codes[g.integers(0, codes.shape[0], 4096)] * 2.0, ((g.random((4096, 3)) * 2 - 1) * 0.2).astype(np.float32)], 1).astype(np.float32)
    # keep the fp16 operand scales of a sane calibration out of the picture: only the alive-unit assumption is under test
    (y1, j1), d1 = _counters_delta(dec, lambda: dec._eval_rows(torch.from_numpy(far).cuda(), with_jac=True))
    sat = dec.saturation_count()
    dec.set_sparse_plan(False)
    y0, j0 = dec._eval_rows(torch.from_numpy(far).cuda(), with_jac=True)
    assert torch.equal(y1, y0) and torch.equal(j1, j0)
    assert d1["tiles_redone_jacobian"] > 0
    if sat == 0:
        y_ref, j_ref = oracle_decoder(np.float64).forward_jac(far.astype(np.float64))
        np.testing.assert_allclose(y1.cpu().numpy(), y_ref, **SDF_TOL)


def test_strawberry_model_rows_and_grid():
    """The second shipped model (deepsdf/models/strawberry_32, configs/lab_berry.yaml: cube radius 0.04, 80^3 grid): decoder rows,
    Jacobians and mesher-grid SDF against golden vectors of the unmodified reference (oracle/gen_golden_full.py).  The fp16
    operand scales are model dependent: the decoder is calibrated on this model's own codes."""
    from hortimapping_b200.decoder import Decoder
    z = load_npz("strawberry_32")
    W, b = [z[f"W{l}"] for l in range(9)], [z[f"b{l}"] for l in range(9)]
    codes = z["latent_codes"]
    dec = Decoder(W, b)
    g = np.random.default_rng(0)
    cal = np.concatenate([codes[g.integers(0, codes.shape[0], 65536)] + 0.02 * g.standard_normal((65536, 32)),
                          (g.random((65536, 3)) * 2 - 1) * 0.075], 1).astype(np.float32)      # like the golden rows: codes + N(0, 0.02)
    dec.calibrate(torch.from_numpy(cal))
    rows = torch.from_numpy(z["rows_rows"]).cuda()
    (y, jac), d = _counters_delta(dec, lambda: dec._eval_rows(rows, with_jac=True))
    assert dec.saturation_count() == 0
    np.testing.assert_allclose(y.cpu().numpy(), z["rows_sdf"], **SDF_TOL)
    assert_jac_close(jac.cpu().numpy(), z["rows_jac"], what="strawberry")
    e_dev = np.abs(y.cpu().numpy().reshape(-1) - z["rows_sdf64"].reshape(-1)).max()
    e_ref = np.abs(z["rows_sdf"].reshape(-1) - z["rows_sdf64"].reshape(-1)).max()
    assert e_dev < 4 * e_ref + 5e-8, (e_dev, e_ref)
    assert d["tiles_redone_jacobian"] <= d["tiles_jacobian"] // 4 and d["kernel_launches"] == 2      # the sparse plan holds for this model too
    sdf = dec.sdf_grid(torch.from_numpy(z["rows_grid_lat"]).cuda(), 80, 0.04).reshape(-1)
    np.testing.assert_allclose(sdf.cpu().numpy()[z["rows_grid_idx"]], z["rows_grid_sdf"], **SDF_TOL)


def test_fp16_saturation_is_attributed_per_fruit_and_underflow_is_bounded():
    """(1) HM_STATUS_F16_SATURATED lands on the fruit whose rows left the calibrated range, not on its batch neighbours.
    (2) The other end of the fp16 range: rows whose operands sit far BELOW the calibrated maximum lose lo bits to fp16 underflow
    (hi keeps 11 bits).  With the 64x headroom of the calibration, inputs 2^-14 of the calibrated scale still give fp32-grade
    SDF values (the first layer's bias dominates such rows); the test pins that bound."""
    from hortimapping_b200 import _lib
    from hortimapping_b200.decoder import Decoder
    from hortimapping_b200.optimizer import Optimizer
    from tests.helpers import random_decoder_weights
    W, b = random_decoder_weights(5)
    dec = Decoder(W, b)
    g = np.random.default_rng(4)
    small = np.concatenate([g.normal(0, 0.01, (4096, 32)), (g.random((4096, 3)) * 2 - 1) * 1e-3], 1).astype(np.float32)
    dec.calibrate(torch.from_numpy(small))
    cfg = cfg_of(load_npz("fruit_wild"))
    cfg["device"] = "cuda"
    cfg["opt"]["converge"]["max_iter"] = 2
    opt = Optimizer(cfg, dec, None, None)
    lat = torch.zeros(3, 32, device="cuda")
    T = torch.eye(4, device="cuda").repeat(3, 1, 1)
    ok_pts = small[:500, 32:]
    bad_pts = ((g.random((300, 3)) * 2 - 1) * 5.0).astype(np.float32)
    _, _, _, st = opt.shape_opt_deepsdf_batch(lat, T, [ok_pts, bad_pts, ok_pts])
    sat = [bool(int(s) & _lib.STATUS["F16_SATURATED"]) for s in st.cpu().tolist()]
    assert sat == [False, True, False], sat
    assert dec.saturation_count() > 0                                    # (reading resets the event counter)
    # underflow side
    from tests.gpu_helpers import random_rows
    cal = random_rows(4096, seed=6)
    dec.calibrate(torch.from_numpy(cal))
    tiny = random_rows(2048, seed=9) * np.float32(2.0 ** -14)
    y, _ = dec._eval_rows(torch.from_numpy(tiny).cuda(), with_jac=False)
    y_ref = O.DecoderOracle(W, b, (4,), np.float64).forward(tiny.astype(np.float64))
    np.testing.assert_allclose(y.cpu().numpy().reshape(-1), y_ref.reshape(-1), rtol=1e-4, atol=2e-6)
    assert dec.saturation_count() == 0


def test_plan_variants_are_bit_identical_over_many_tiles_per_cta():
    """The plan decides where two things live: F0's operand (written for the NEXT tile during the tile's last op when a chunk is
    free, HM_TC_NO_X0_EARLY switches that off) and the ReLU bits of the forward + gradient pass (the plan's free A chunk, or the
    global scratch with HM_TC_NO_SMEM_MASKS).  Neither may change a bit of the output.  Several tiles per CTA and a ragged tail, so
    that the tile hand-over and the chunk reuse across tiles are exercised; forward-only, forward + gradient, and the
    per-fruit latent-table input mode the optimisers use."""
    import os
    from hortimapping_b200.decoder import Decoder, calibration_rows
    from tests.helpers import pepper_weights
    W, b, codes = pepper_weights()
    n = 148 * 64 * 5 + 77
    g = np.random.default_rng(5)
    rows = np.concatenate([codes[g.integers(0, codes.shape[0], n)], ((g.random((n, 3)) * 2 - 1) * 0.06).astype(np.float32)], 1)
    t = torch.from_numpy(rows).cuda()
    cal = calibration_rows(codes, 0.15, n=int(os.environ.get("HM_TEST_CAL_ROWS", "65536")))

    def run(env):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update({k: "1" for k in env})
        try:
            dec = Decoder(W, b)            # the plan is built (and the switches are read) when the engine is calibrated
            dec.calibrate(cal)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        f, _ = dec._eval_rows(t, with_jac=False)
        y, j = dec._eval_rows(t, with_jac=True)
        y1, j1 = dec.sdf_jacobian(t[0, :32].contiguous(), t[:, 32:].contiguous())
        return f, y, j, y1, j1

    base = run(())
    assert torch.equal(base[0], base[1])
    for env in (("HM_TC_NO_X0_EARLY",), ("HM_TC_NO_SMEM_MASKS",), ("HM_TC_NO_X0_EARLY", "HM_TC_NO_SMEM_MASKS")):
        other = run(env)
        for a, o in zip(base, other):
            assert torch.equal(a, o), env


def test_testing_library_decoder_is_the_product_decoder():
    """The TEST-ONLY library compiles the same kernels with timeline / wait-cycle instrumentation.  Different register allocation and
    timing, same protocol: its forward-only and forward + gradient launches over many tiles per CTA must finish and give the
    product's bits (a protocol that only works for one instruction schedule shows up here, scripts/check_testing_decoder.py)."""
    import os
    from hortimapping_b200 import _testing
    from hortimapping_b200.decoder import calibration_rows
    from tests.gpu_helpers import pepper_decoder
    from tests.helpers import pepper_weights
    W, b, codes = pepper_weights()
    ref = pepper_decoder()
    dec = _testing.testing_decoder(W, b)
    dec.calibrate(calibration_rows(codes, 0.15, n=int(os.environ.get("HM_TEST_CAL_ROWS", "262144"))))
    n = 148 * 64 * 6 + 5
    g = np.random.default_rng(11)
    rows = np.concatenate([codes[g.integers(0, codes.shape[0], n)], ((g.random((n, 3)) * 2 - 1) * 0.05).astype(np.float32)], 1)
    t = torch.from_numpy(rows).cuda()
    for jac in (False, True):
        a, r = dec._eval_rows(t, with_jac=jac), ref._eval_rows(t, with_jac=jac)
        torch.cuda.synchronize()
        assert torch.equal(a[0], r[0])
        if jac:
            assert torch.equal(a[1], r[1])
