"""GPU parity tests of the device-side LM loop (wild_completion/optimizer.py) through the C ABI."""
import copy

import numpy as np
import pytest
import torch

from oracle import hm_oracle as O
from tests.helpers import cfg_of, load_npz, oracle_decoder, render_data_of
from tests.test_oracle_golden import joint_cost

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def make_opt(cfg, engine="tc"):
    from hortimapping_b200.optimizer import Optimizer
    from tests.gpu_helpers import pepper_decoder
    dec = pepper_decoder()
    dec.set_engine(engine)
    cfg = copy.deepcopy(cfg)
    cfg["device"] = "cuda"
    return Optimizer(cfg, dec, None, None), dec


def zero_eps(cfg, max_iter):
    cfg = copy.deepcopy(cfg)
    cfg["opt"]["converge"]["max_iter"] = max_iter
    for k in ("epsilon_g", "epsilon_c", "epsilon_t", "epsilon_r", "epsilon_s"):
        cfg["opt"]["converge"][k] = 0
    return cfg


def last_system(dec, n_fruits, est):
    from hortimapping_b200 import _lib
    H = torch.empty(n_fruits, est, est, device="cuda")
    b = torch.empty(n_fruits, est, device="cuda")
    dx = torch.empty(n_fruits, est, device="cuda")
    _lib.check(dec._L.hm_get_last_system(dec.handle, n_fruits, H.data_ptr(), b.data_ptr(), dx.data_ptr(), torch.cuda.current_stream().cuda_stream), "last")
    return H.cpu().numpy(), b.cpu().numpy(), dx.cpu().numpy()


@pytest.mark.parametrize("case_name", ["fruit_wild", "fruit_challenge"])
def test_lm_every_iteration_replayed_from_reference_state(case_name):
    """Step-level parity from IDENTICAL state: iteration i of the reference's run is replayed as ONE device
    iteration from the reference's own state (iter_offset = i).  Same flip-tolerant contract as the oracle
    test (tests/test_oracle_golden.py): typical iteration within `tol`, none beyond a single-sample flip."""
    c = load_npz(case_name)
    cfg = zero_eps(cfg_of(c), 1)
    opt, dec = make_opt(cfg)
    pk = bool(c["pose_known"])
    est = (7 if cfg["opt"]["scale_on"] else 6) + 32
    n = c["trace_H"].shape[0]
    tol = 2e-4 if case_name == "fruit_wild" else 5e-4                 # (measured: fruit_challenge H 5e-6 .. 1.6e-4, b 2e-6 .. 2.6e-4)
    clean_tol = 1e-4                                                  # typical iteration without a membership flip
    flip_tol = 5e-3
    eH, eb, edx, elat, eT, flips = [], [], [], [], [], []
    rd = render_data_of(c)
    orc = oracle_decoder(np.float32)
    for i in range(n):
        lat0 = c["init_latent"] if i == 0 else c[f"after{i}_latent"]
        T0 = c["init_T_ow"] if i == 0 else c[f"after{i}_T_ow"]
        lat = torch.from_numpy(lat0.copy()).cuda().reshape(1, 32)
        T = torch.from_numpy(T0.copy()).cuda().reshape(1, 4, 4)
        c0 = dec.counters()
        _, _, iters, status = opt.shape_pose_joint_opt_batch(lat, T, [rd], [c["points_w"]], float(c["cube_radius"]), pk,
                                                             iter_offset=i, max_iter=1)
        c1 = dec.counters()
        assert int(iters.item()) == 1
        H, b, dx = last_system(dec, 1, est)
        eH.append(rel(H[0], c["trace_H"][i]))
        eb.append(rel(b[0], c["trace_b"][i]))
        edx.append(rel(dx[0], c["trace_dx"][i]))
        elat.append(rel(lat.cpu().numpy()[0], c[f"after{i + 1}_latent"]))
        eT.append(rel(T.cpu().numpy()[0], c[f"after{i + 1}_T_ow"]))
        # sample membership (in-sphere, in-band, occlusion: hard thresholds): rows the oracle selects from the same state vs the
        # rows the device evaluated (exact device-side counters)
        tr = O.OptTrace()
        O.shape_pose_joint_opt(orc, cfg, lat0.copy(), T0.copy(), rd, c["points_w"], float(c["cube_radius"]), pk, trace=tr, iter_offset=i)
        d_rows = (c1["rows_forward"] - c0["rows_forward"] - tr.rows_fwd,
                  (c1["rows_jacobian"] - c0["rows_jacobian"]) + (c1["rows_backward"] - c0["rows_backward"]) - tr.rows_grad)
        flips.append(abs(d_rows[0]) + abs(d_rows[1]))
    print(f"{case_name}: per-iteration membership flips {flips}; H errors {['%.1e' % e for e in eH]}; b errors {['%.1e' % e for e in eb]}")
    clean = [i for i in range(n) if flips[i] == 0]
    assert len(clean) >= n // 2, flips
    # With the same samples selected, the typical iteration is held to north_star's 1e-4 (measured: H 4e-7 .. 3e-6, b 1e-6 .. 8e-6).
    # What remains are ReLU-kink events: a hidden unit within rounding of zero gives one row another (piecewise constant) gradient;
    # when that row is a high-weight in-band sample an entry of H moves by up to ~1e-3 (fruit_wild iteration 6: 8.5e-4) although
    # no sample changed membership.  They are counted, and bounded by the single-sample tolerance.
    eHc, ebc = [eH[i] for i in clean], [eb[i] for i in clean]
    assert np.median(eHc) < clean_tol and np.median(ebc) < clean_tol, (clean, eH, eb)
    assert sum(e > clean_tol for e in eHc) <= max(1, len(clean) // 3), (clean, eH)
    assert eH[0] < tol and eb[0] < tol, (eH, eb)
    assert np.median(eH) < tol and max(eH) < flip_tol, eH
    assert np.median(eb) < tol and max(eb) < 4 * flip_tol, eb
    assert np.median(edx) < 20 * tol, edx
    assert np.median(elat) < 10 * tol and max(elat) < 10 * flip_tol, elat
    assert np.median(eT) < 10 * tol and max(eT) < flip_tol, eT


def test_solve_matches_fp64_oracle_from_identical_state():
    """dx of the device solve (fp64 elimination) against the fp64 oracle from the same state: the reference's
    own fp32 `torch.inverse` is the less accurate of the two (cond(H) ~ 1e5)."""
    c = load_npz("fruit_wild")
    cfg = zero_eps(cfg_of(c), 1)
    opt, dec = make_opt(cfg)
    rd = render_data_of(c)
    lat = torch.from_numpy(c["init_latent"].copy()).cuda().reshape(1, 32)
    T = torch.from_numpy(c["init_T_ow"].copy()).cuda().reshape(1, 4, 4)
    opt.shape_pose_joint_opt_batch(lat, T, [rd], [c["points_w"]], float(c["cube_radius"]), False, max_iter=1)
    H, b, dx = last_system(dec, 1, 39)
    tr = O.OptTrace()
    O.shape_pose_joint_opt(oracle_decoder(np.float64), cfg, c["init_latent"].astype(np.float64), c["init_T_ow"].astype(np.float64),
                           rd, c["points_w"], float(c["cube_radius"]), False, trace=tr)
    assert rel(H[0], tr.H[0]) < 1e-4
    assert rel(b[0], tr.b[0]) < 1e-4
    assert rel(dx[0], tr.dx[0]) < 1e-3
    assert rel(lat.cpu().numpy()[0], tr.latent[0]) < 1e-4


@pytest.mark.parametrize("var", ["se3", "lmeye", "linocc", "noocc", "gn"])
def test_config_variants_first_step(var):
    c = load_npz("fruit_wild")
    cfg = zero_eps(cfg_of(c, f"var_{var}_cfg_json"), 1)
    opt, dec = make_opt(cfg)
    est = (7 if cfg["opt"]["scale_on"] else 6) + 32
    lat = torch.from_numpy(c["init_latent"].copy()).cuda().reshape(1, 32)
    T = torch.from_numpy(c["init_T_ow"].copy()).cuda().reshape(1, 4, 4)
    opt.shape_pose_joint_opt_batch(lat, T, [render_data_of(c)], [c["points_w"]], float(c["cube_radius"]), False, max_iter=1)
    H, b, dx = last_system(dec, 1, est)
    tol = 1e-3 if var == "linocc" else 2e-4
    assert H[0].shape == c[f"var_{var}_H"][0].shape
    assert rel(H[0], c[f"var_{var}_H"][0]) < tol
    assert rel(b[0], c[f"var_{var}_b"][0]) < tol


@pytest.mark.parametrize("case_name", ["fruit_wild", "fruit_challenge"])
def test_reference_signature_and_trajectory_equivalence(case_name):
    """shape_pose_joint_opt with the reference's call signature (test_wild_completion.py:226): latent updated
    IN PLACE and returned, T_ow returned as a new tensor, iter_count an int.  Trajectories are chaotic
    (SURVEY.md 7.4), so the 12-iteration result is compared through the objective it reaches."""
    c = load_npz(case_name)
    cfg = zero_eps(cfg_of(c), 12)
    opt, dec = make_opt(cfg)
    latent = torch.from_numpy(c["init_latent"].copy()).cuda()
    T_in = torch.from_numpy(c["init_T_ow"].copy()).cuda()
    rd = {k: [torch.from_numpy(a).cuda() for a in v] for k, v in render_data_of(c).items()}
    out, T, it = opt.shape_pose_joint_opt(latent, T_in, rd, torch.from_numpy(c["points_w"]).cuda(), float(c["cube_radius"]),
                                          [0.5, 0.5, 0.5], bool(c["pose_known"]))
    assert out is latent and isinstance(it, int) and it == 12
    assert T is not T_in and torch.equal(T_in.cpu(), torch.from_numpy(c["init_T_ow"]))
    assert tuple(T.shape) == (4, 4) and T.dtype == torch.float32 and T.is_cuda
    ref_cfg = cfg_of(c)
    c_ref = joint_cost(c, ref_cfg, c["final_latent"], c["final_T_ow"])
    c_ref64 = joint_cost(c, ref_cfg, c["final64_latent"], c["final64_T_ow"])
    c_ours = joint_cost(c, ref_cfg, latent.cpu().numpy(), T.cpu().numpy())
    assert c_ours < 1.25 * max(c_ref, c_ref64), (c_ours, c_ref, c_ref64)


@pytest.mark.parametrize("case_name", ["fruit_wild", "fruit_challenge"])
def test_stop_rules(case_name):
    c = load_npz(case_name)
    base = cfg_of(c)
    for ename in ("epsilon_g", "epsilon_c", "epsilon_s"):
        cfg = zero_eps(base, 10)
        cfg["opt"]["converge"][ename] = 1e9
        if ename == "epsilon_s":
            cfg["opt"]["converge"]["epsilon_t"] = 1e9
            cfg["opt"]["converge"]["epsilon_r"] = 1e9
        opt, dec = make_opt(cfg)
        latent = torch.from_numpy(c["init_latent"].copy()).cuda()
        _, _, it = opt.shape_pose_joint_opt(latent, torch.from_numpy(c["init_T_ow"]).cuda(), render_data_of(c),
                                            torch.from_numpy(c["points_w"]).cuda(), float(c["cube_radius"]), [0.5] * 3, bool(c["pose_known"]))
        assert it == int(c[f"stop_{ename}_iters"]), ename
        bit = {"epsilon_g": 0x01, "epsilon_c": 0x02, "epsilon_s": 0x04}[ename]
        if it < 10:
            assert opt.last_status[0] & bit


def test_shape_opt_deepsdf_vs_reference():
    c = load_npz("fruit_wild")
    cfg = zero_eps(cfg_of(c), 1)
    opt, dec = make_opt(cfg)
    # iteration-0 system against the reference's capture
    lat = torch.from_numpy(c["init_latent"].copy()).cuda().reshape(1, 32)
    T = torch.from_numpy(c["init_T_ow"].copy()).cuda().reshape(1, 4, 4)
    opt.shape_opt_deepsdf_batch(lat, T, [c["points_w"]], max_iter=1)
    H, b, dx = last_system(dec, 1, 32)
    assert rel(H[0], c["shape_H"][0]) < 1e-4
    assert rel(b[0], c["shape_b"][0]) < 1e-4
    assert rel(dx[0], c["shape_dx"][0]) < 1e-3
    # the convergent phase (|dx| falls from 8e-2 to 1e-4 in ~6 iterations): each of the first 5 LM steps, replayed on the device
    # from the fp64 oracle's own state, lands within north_star's 1e-4 of the fp64 step -- both engines.  (Typical 1.4e-6; a row
    # that crosses a ReLU kink between two correct evaluations moves the step by ~2.5e-5.  Whole trajectories cannot be held to
    # 1e-4: the loop amplifies such a crossing, scripts/diag_parity.py.)
    cfg5 = zero_eps(cfg_of(c), 5)
    tr = O.OptTrace()
    O.shape_opt_deepsdf(oracle_decoder(np.float64), cfg5, c["init_latent"].astype(np.float64).copy(), c["init_T_ow"].astype(np.float64), c["points_w"],
                        trace=tr)
    for engine in ("tc", "simt"):
        opt5, dec5 = make_opt(cfg5, engine)
        try:
            for i in range(5):
                s_i = (c["init_latent"] if i == 0 else tr.latent[i - 1]).astype(np.float32)
                l5 = torch.from_numpy(s_i.copy()).cuda().reshape(1, 32)
                opt5.shape_opt_deepsdf_batch(l5, T.clone(), [c["points_w"]], iter_offset=i, max_iter=1)
                assert rel(l5[0].cpu().numpy(), tr.latent[i]) < 1e-4, (engine, i)
        finally:
            dec5.set_engine("tc")
    # 30 iterations: past convergence the loop keeps moving by ~1e-4 per iteration along flat directions (|b| at its fp32 cancellation
    # floor, single rows crossing ReLU kinks) and any two correct runs drift ~1e-3 apart (scripts/diag_parity.py: the reference's
    # fp32 vs fp64 runs, numpy fp32, both device engines) -> held to the reference's own fp32-vs-fp64 distance
    cfg30 = zero_eps(cfg_of(c), 30)
    opt30, _ = make_opt(cfg30)
    latent = torch.from_numpy(c["init_latent"].copy()).cuda()
    out, T_out, it = opt30.shape_opt_deepsdf(latent, torch.from_numpy(c["init_T_ow"]).cuda(), torch.from_numpy(c["points_w"]).cuda(), [0.5] * 3)
    assert out is latent and it == 30
    e_ours = rel(latent.cpu().numpy(), c["shape64_latent30"])
    e_ref = rel(c["shape_latent30"], c["shape64_latent30"])
    assert e_ours < max(20 * e_ref, 2e-3), (e_ours, e_ref)


def test_batch_equals_single_and_is_deterministic():
    """Fruits are independent: a batch of 3 gives bit-identical results to 3 single calls, twice in a row."""
    c = load_npz("fruit_wild")
    c2 = load_npz("fruit_challenge")
    cfg = zero_eps(cfg_of(c), 4)
    opt, dec = make_opt(cfg)
    fruits = [(c, False), (c2, False), (c, True)]
    lat_s, T_s = [], []
    for cc, pk in fruits:
        lat = torch.from_numpy(cc["init_latent"].copy()).cuda().reshape(1, 32)
        T = torch.from_numpy(cc["init_T_ow"].copy()).cuda().reshape(1, 4, 4)
        opt.shape_pose_joint_opt_batch(lat, T, [render_data_of(cc)], [cc["points_w"]], float(cc["cube_radius"]), pk)
        lat_s.append(lat.clone()); T_s.append(T.clone())
    for _ in range(2):
        lat = torch.stack([torch.from_numpy(cc["init_latent"].copy()) for cc, _ in fruits]).cuda()
        T = torch.stack([torch.from_numpy(cc["init_T_ow"].copy()) for cc, _ in fruits]).cuda()
        _, _, iters, status = opt.shape_pose_joint_opt_batch(lat, T, [render_data_of(cc) for cc, _ in fruits],
                                                             [cc["points_w"] for cc, _ in fruits],
                                                             [float(cc["cube_radius"]) for cc, _ in fruits], [pk for _, pk in fruits])
        assert iters.cpu().tolist() == [4, 4, 4]
        for f in range(3):
            assert torch.equal(lat[f], lat_s[f][0]) and torch.equal(T[f], T_s[f][0]), f


def test_mask_reuse_is_bit_identical_and_saves_the_second_forward():
    """Joint loop: the in-band ray samples are evaluated forward among all in-sphere samples and then differentiated
    (loss.py:47-49, :185-215).  With hm_set_mask_reuse on (the default) the forward launch keeps every row's ReLU bits and a
    gradient-only launch starts from them; with it off the in-band rows go through forward + gradient again, as in the reference.
    Same operands, same products: states, iteration counts and the last LM system must be BIT-identical, and the row counters
    must show where the rows went (observed points -> forward + gradient, in-band samples -> gradient only)."""
    c = load_npz("fruit_wild")
    c2 = load_npz("fruit_challenge")
    cfg = zero_eps(cfg_of(c), 5)
    opt, dec = make_opt(cfg)
    fruits = [(c, False), (c2, False), (c, True)]
    n_pts = sum(cc["points_w"].shape[0] for cc, _ in fruits)
    res = {}
    try:
        for on in (True, False):
            dec.set_mask_reuse(on)
            lat = torch.stack([torch.from_numpy(cc["init_latent"].copy()) for cc, _ in fruits]).cuda()
            T = torch.stack([torch.from_numpy(cc["init_T_ow"].copy()) for cc, _ in fruits]).cuda()
            c0 = dec.counters()
            _, _, iters, status = opt.shape_pose_joint_opt_batch(lat, T, [render_data_of(cc) for cc, _ in fruits], [cc["points_w"] for cc, _ in fruits],
                                                                 [float(cc["cube_radius"]) for cc, _ in fruits], [pk for _, pk in fruits])
            c1 = dec.counters()
            H, b, dx = last_system(dec, 3, 39)
            res[on] = (lat.clone(), T.clone(), iters.clone(), status.clone(), H, b, dx, {k: c1[k] - c0[k] for k in c1})
    finally:
        dec.set_mask_reuse(True)
    a, z = res[True], res[False]
    assert a[2].cpu().tolist() == [5, 5, 5]
    for i in range(4):
        assert torch.equal(a[i], z[i]), i
    for i in (4, 5, 6):
        np.testing.assert_array_equal(a[i], z[i])
    da, dz = a[7], z[7]
    assert dz["rows_backward"] == 0 and da["rows_backward"] > 0
    assert da["rows_jacobian"] == 5 * n_pts                                    # only the observed points are evaluated forward + gradient
    assert da["rows_jacobian"] + da["rows_backward"] == dz["rows_jacobian"]     # the same gradient rows either way
    assert da["rows_forward"] == dz["rows_forward"]


def test_invalid_submap_keeps_state():
    """No frame with >= 100 in-sphere samples -> "This submap is not valid": state untouched, iter_count 0
    (optimizer.py:139-141); other fruits of the batch are unaffected."""
    c = load_npz("fruit_wild")
    cfg = zero_eps(cfg_of(c), 3)
    opt, dec = make_opt(cfg)
    lat = torch.stack([torch.from_numpy(c["init_latent"].copy())] * 2).cuda()
    T = torch.stack([torch.from_numpy(c["init_T_ow"].copy())] * 2).cuda()
    _, _, iters, status = opt.shape_pose_joint_opt_batch(lat, T, [render_data_of(c)] * 2, [c["points_w"]] * 2, [1e-4, float(c["cube_radius"])], False)
    assert iters.cpu().tolist() == [0, 3]
    assert status.cpu().numpy()[0] & 0x20 and status.cpu().numpy()[0] & 0x10
    np.testing.assert_array_equal(lat[0].cpu().numpy(), c["init_latent"])
    np.testing.assert_array_equal(T[0].cpu().numpy(), c["init_T_ow"])
    assert not np.array_equal(lat[1].cpu().numpy(), c["init_latent"])


def test_full_size_batch_properties():
    """BASELINE.json configs[1] at full size (64 fruits x 2048 points), a few iterations, through size-independent properties:
    every fruit of the batch gets bit-identically what a single-fruit call gives it (fruits are independent; several waves
    of the persistent decoder grid, ragged tile assignment), permuting the fruits permutes the result, the weighted SDF
    objective the loop minimises does not increase, and the device loop agrees with the fp64 oracle for a sampled fruit."""
    from hortimapping_b200 import synth
    c = load_npz("fruit_wild")
    cfg = zero_eps(cfg_of(c), 4)
    opt, dec = make_opt(cfg)
    _, _, codes = __import__("tests.helpers", fromlist=["pepper_weights"]).pepper_weights()
    n_f, n_p = 64, 2048
    g = np.random.default_rng(21)
    # surface-like points: mean-shape level set points jittered, one cloud per fruit
    base = ((g.random((n_f, n_p, 3)) * 2 - 1) * 0.045).astype(np.float32)
    lat0 = (codes.mean(0)[None, :] + 0.02 * g.standard_normal((n_f, 32))).astype(np.float32)
    pts = [base[i] for i in range(n_f)]
    T0 = torch.eye(4).repeat(n_f, 1, 1).cuda()
    lat = torch.from_numpy(lat0).cuda()
    lat_b, _, iters, status = opt.shape_opt_deepsdf_batch(lat.clone(), T0.clone(), pts)
    assert iters.cpu().tolist() == [4] * n_f and all(int(s) == 0x8 for s in status.cpu().tolist())
    for f in (0, 17, 63):                                             # batch == single, bit for bit
        l1, _, _, _ = opt.shape_opt_deepsdf_batch(lat[f:f + 1].clone(), T0[f:f + 1].clone(), [pts[f]])
        assert torch.equal(l1[0], lat_b[f]), f
    perm = g.permutation(n_f)
    lat_p, _, _, _ = opt.shape_opt_deepsdf_batch(lat[perm].clone(), T0.clone(), [pts[i] for i in perm])
    assert torch.equal(lat_p, lat_b[perm])
    # objective of the latent-only loop (optimizer.py:306-429): mean squared SDF + code regulariser, before vs after
    def objective(lv, f):
        s = dec.sdf(lv, torch.from_numpy(pts[f]).cuda()).double()
        return float((s * s).mean()) * float(cfg["opt"]["weight"]["w_recon"]) + float(cfg["opt"]["weight"]["w_codereg"]) * float((lv.double() ** 2).sum())
    worse = sum(objective(lat_b[f], f) > objective(lat[f], f) * (1 + 1e-6) for f in range(0, n_f, 7))
    assert worse == 0
    f = 5
    ref = lat0[f].astype(np.float64).copy()
    O.shape_opt_deepsdf(oracle_decoder(np.float64), cfg, ref, np.eye(4), pts[f])
    np.testing.assert_allclose(lat_b[f].cpu().numpy(), ref, rtol=2e-3, atol=2e-5)


@pytest.mark.parametrize("case_name,model", [("fruit_full", "sweetpepper_32"), ("fruit_berry", "strawberry_32")])
def test_full_size_joint_step_replay_counts_membership_flips(case_name, model):
    """Step replay at BASELINE.json's FULL joint sizes (fruit_full: 10 frames x 400 rays x 30 samples + 2048 points, wild_pepper.yaml)
    and on the second shipped model (fruit_berry: strawberry_32, lab_berry.yaml).  Each iteration of the unmodified reference's run
    is replayed as ONE device iteration from the reference's own state.  Sample membership is decided by hard thresholds
    (in-sphere, |sdf| < th, de_do > 1e-6, occlusion): the goldens record how many decoder rows the reference selected per iteration
    (forward hook), the device counts its rows exactly -- when the two agree no sample flipped and H, b must match to 1e-4
    (north_star's tolerance); a flipped sample changes H, b by up to its own weight and is then counted, not hidden."""
    from hortimapping_b200.decoder import Decoder
    from hortimapping_b200.optimizer import Optimizer
    c = load_npz(case_name)
    cfg = zero_eps(cfg_of(c), 1)
    cfg["device"] = "cuda"
    if model == "sweetpepper_32":
        opt, dec = make_opt(cfg)
    else:
        z = load_npz(model)
        dec = Decoder([z[f"W{l}"] for l in range(9)], [z[f"b{l}"] for l in range(9)])
        g = np.random.default_rng(0)
        codes = z["latent_codes"]
        cal = np.concatenate([codes[g.integers(0, codes.shape[0], 8192)], ((g.random((8192, 3)) * 2 - 1) * 0.075).astype(np.float32)], 1)
        dec.calibrate(torch.from_numpy(cal))
        opt = Optimizer(cfg, dec, None, None)
    rd = render_data_of(c)
    n = c["trace_H"].shape[0]
    flips, worst_clean = 0, 0.0
    for i in range(n):
        lat0 = c["init_latent"] if i == 0 else c[f"after{i}_latent"]
        T0 = c["init_T_ow"] if i == 0 else c[f"after{i}_T_ow"]
        lat = torch.from_numpy(lat0.copy()).cuda().reshape(1, 32)
        T = torch.from_numpy(T0.copy()).cuda().reshape(1, 4, 4)
        c0 = dec.counters()
        _, _, iters, status = opt.shape_pose_joint_opt_batch(lat, T, [rd], [c["points_w"]], float(c["cube_radius"]), bool(c["pose_known"]),
                                                             iter_offset=i, max_iter=1)
        c1 = dec.counters()
        assert int(iters.item()) == 1 and not (int(status.item()) & 0x40)
        H, b, dx = (t.cpu().numpy()[0] for t in opt.last_system(1))
        d_fwd = c1["rows_forward"] - c0["rows_forward"] - int(c["trace_rows_fwd"][i])
        # gradient rows = observed points (forward + gradient) + in-band samples (gradient only, from the forward pass's ReLU bits)
        d_jac = (c1["rows_jacobian"] - c0["rows_jacobian"]) + (c1["rows_backward"] - c0["rows_backward"]) - int(c["trace_rows_jac"][i])
        eH, eb = rel(H, c["trace_H"][i]), rel(b, c["trace_b"][i])
        if d_fwd == 0 and d_jac == 0:
            worst_clean = max(worst_clean, eH, eb)
            assert eH < 1e-4 and eb < 1e-4, (case_name, i, eH, eb)
        else:
            flips += abs(d_fwd) + abs(d_jac)
            assert eH < 5e-3 and eb < 2e-2, (case_name, i, d_fwd, d_jac, eH, eb)
        e_lat = rel(lat.cpu().numpy()[0], c[f"after{i + 1}_latent"])
        assert e_lat < 5e-2, (case_name, i, e_lat, d_fwd, d_jac, eH, eb, rel(dx, c["trace_dx"][i]), hex(int(status.item())))
    assert flips <= 2 * n, (case_name, flips)
    print(f"{case_name}: {n} iterations replayed, {flips} membership flips, worst H/b error without flips {worst_clean:.2e}")


def test_lm_device_functions_vs_reference_vectors():
    """exp_sim3 / exp_se3 (utils.py:220-324 incl. the theta <= eps, s == 0 and `c = 0 for s <= 0` branches) and the squared Huber
    weights (utils.py:327-358 incl. the exact-zero residual) evaluated by the device functions of the LM step themselves, through
    the test-only library's hooks, against vectors written by the unmodified reference (tests/golden/misc.npz)."""
    import ctypes as C
    from hortimapping_b200 import _lib, _testing
    from tests.helpers import pepper_weights
    W, b, _ = pepper_weights()
    dec = _testing.testing_decoder(W, b)
    L = _testing.lib()
    m = load_npz("misc")
    x = np.ascontiguousarray(m["exp_x"], np.float32)
    n = x.shape[0]
    for pd, key in ((7, "exp_sim3"), (6, "exp_se3")):
        T = np.zeros((n, 16), np.float32)
        _lib.check(L.hm_debug_exp_pose(dec.handle, x.ctypes.data_as(_lib.c_float_p), n, pd, T.ctypes.data_as(_lib.c_float_p)), "exp")
        np.testing.assert_allclose(T.reshape(n, 4, 4), m[key], rtol=2e-6, atol=2e-7)
    r = np.ascontiguousarray(m["huber_res"], np.float32)
    w2 = np.zeros_like(r)
    _lib.check(L.hm_debug_huber_w2(dec.handle, r.ctypes.data_as(_lib.c_float_p), r.size, 0.02, w2.ctypes.data_as(_lib.c_float_p)), "huber")
    np.testing.assert_allclose(w2, m["huber_w2"].reshape(-1), rtol=2e-6, atol=0)
    assert w2[5] == 0.0                                             # the exact-zero residual gets weight 0 (utils.py:337-338)


def test_batched_api_rejects_bad_state_tensors():
    """The C ABI takes raw device pointers: the batched API must refuse anything but contiguous float32 CUDA tensors of the
    right shape on the decoder's GPU instead of corrupting memory."""
    c = load_npz("fruit_wild")
    opt, dec = make_opt(zero_eps(cfg_of(c), 1))
    lat = torch.from_numpy(c["init_latent"].copy()).reshape(1, 32)
    T = torch.from_numpy(c["init_T_ow"].copy()).reshape(1, 4, 4)
    good = (lat.cuda(), T.cuda())
    for bad_lat, bad_T in ((lat, T.cuda()), (lat.cuda().double(), T.cuda()), (lat.cuda().repeat(1, 2)[:, ::2], T.cuda()),
                           (lat.cuda().reshape(32), T.cuda()), (lat.cuda(), T.cuda().reshape(16))):
        with pytest.raises(ValueError):
            opt.shape_opt_deepsdf_batch(bad_lat, bad_T, [c["points_w"]])
    with pytest.raises(ValueError):
        opt.shape_opt_deepsdf_batch(good[0], good[1], [c["points_w"], c["points_w"]])
    opt.shape_opt_deepsdf_batch(good[0], good[1], [c["points_w"]])


def test_degenerate_batches():
    """A fruit without surface points stops with SUBMAP_INVALID and untouched state; its batch neighbours are unaffected."""
    c = load_npz("fruit_wild")
    opt, dec = make_opt(zero_eps(cfg_of(c), 2))
    lat = torch.stack([torch.from_numpy(c["init_latent"].copy())] * 2).cuda()
    T = torch.stack([torch.from_numpy(c["init_T_ow"].copy())] * 2).cuda()
    _, _, iters, status = opt.shape_opt_deepsdf_batch(lat, T, [np.zeros((0, 3), np.float32), c["points_w"]])
    assert iters.cpu().tolist() == [0, 2] and int(status[0].item()) & 0x20
    np.testing.assert_array_equal(lat[0].cpu().numpy(), c["init_latent"])
    assert not np.array_equal(lat[1].cpu().numpy(), c["init_latent"])
