"""CPU tests of the host-side logic of the product package (no GPU, no compute calls into the library)."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

from tests.helpers import cfg_of, load_npz, pepper_weights, render_data_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_opt_params_mirror_reference_config_reads():
    from hortimapping_b200.optimizer import opt_params_from_cfg
    cfg = cfg_of(load_npz("fruit_wild"))
    # YAML parses `5e-2` as str in the reference's configs; float() casts must accept both
    cfg["opt"]["weight"]["w_depth"] = "5e-2"
    p = opt_params_from_cfg(cfg["opt"], iter_offset=3, max_iter=7)
    assert p.max_iter == 7 and p.iter_offset == 3
    assert p.w_depth == 0.05 and p.w_mask == 5e-4 and p.w_recon == 1.0 and p.w_codereg == 5e-4
    assert p.n_depth_samples == 30 and p.log_sdf_occ == 1 and p.occlusion_on == 1 and p.scale_on == 1
    assert p.lm_on == 1 and p.lm_eye == 0 and p.lm_lambda_0 == 0.1 and p.s_damp == 1e-3
    assert p.robust_iter == 5 and p.robust_th_recon == 0.01 and p.robust_th_depth == 0.05
    assert (p.occlusion_th, p.min_valid_sample, p.min_grad_thre) == (0.03, 100, 1e-6)     # loss.py:11


def test_frame_selection_matches_reference_linspace():
    from hortimapping_b200.optimizer import select_frames
    for n, mx in ((4, 10), (10, 10), (23, 10), (7, 5), (1, 10)):
        np.testing.assert_array_equal(select_frames(n, mx), np.linspace(0, n - 1, min(mx, n)).astype(np.int32))


def test_packed_batch_layout():
    from hortimapping_b200.optimizer import PackedBatch
    c = load_npz("fruit_wild")
    rd = render_data_of(c)
    rd_t = {k: [torch.from_numpy(a) for a in v] for k, v in rd.items()}          # torch inputs like the reference host
    pk = PackedBatch([c["points_w"], c["points_w"][:100]], [rd, rd_t], 3, [0.08, 0.07], [False, True])
    assert pk.n_fruits == 2 and pk.point_offsets.tolist() == [0, 512, 612]
    assert pk.frame_offsets.tolist() == [0, 3, 6]                                 # 4 frames sub-sampled to 3
    sel = np.linspace(0, 3, 3).astype(np.int32)
    for j, idx in enumerate(sel):
        lo, hi = pk.ray_offsets[j], pk.ray_offsets[j + 1]
        nfg = rd["rays_fg"][idx].shape[0]
        assert pk.n_fg[j] == nfg and hi - lo == nfg + rd["rays_bg"][idx].shape[0]
        np.testing.assert_array_equal(pk.rays[lo:lo + nfg], rd["rays_fg"][idx])   # fg first, then bg (optimizer.py:113)
        np.testing.assert_array_equal(pk.rays[lo + nfg:hi], rd["rays_bg"][idx])
        np.testing.assert_array_equal(pk.depth_obs[lo:lo + nfg], rd["depth_fg"][idx])
        np.testing.assert_array_equal(pk.T_wc[j].reshape(4, 4), rd["T_wc"][idx])
    np.testing.assert_array_equal(pk.rays[pk.ray_offsets[3]:], pk.rays[:pk.ray_offsets[3]])
    assert pk.pose_known.tolist() == [0, 1] and pk.cube_radius.dtype == np.float32
    # empty render data (no matched frame) is representable: the device loop then reports "submap not valid"
    # scalars / stride-0 broadcasts must become real per-fruit arrays (the C side indexes them by fruit)
    pk3 = PackedBatch([c["points_w"]] * 3, [rd] * 3, 3, np.broadcast_to(np.float32(0.08), (3,)), np.broadcast_to(False, (3,)))
    assert pk3.cube_radius.flags["C_CONTIGUOUS"] and pk3.cube_radius.strides == (4,) and pk3.cube_radius.tolist() == [np.float32(0.08)] * 3
    assert pk3.pose_known.strides == (1,) and pk3.pose_known.tolist() == [0, 0, 0]
    pk2 = PackedBatch([c["points_w"]], [{k: [] for k in rd}], 10, [0.08], [False])
    assert pk2.frame_offsets.tolist() == [0, 0] and pk2.rays.shape == (0, 3)


def test_marching_tetrahedra_sphere_is_watertight_and_outward():
    from hortimapping_b200.marching import marching_tetrahedra
    n = 24
    g = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    v, f = marching_tetrahedra(np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.6, 0.0, [2 / (n - 1)] * 3)
    v = v - 1
    assert np.abs(np.linalg.norm(v, axis=1) - 0.6).max() < 5e-3
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), 1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert np.all(cnt == 2)
    nrm = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    assert np.all((nrm * v[f].mean(1)).sum(1) > 0)
    assert abs(0.5 * np.linalg.norm(nrm, axis=1).sum() - 4 * np.pi * 0.36) < 0.05
    ve, fe = marching_tetrahedra(np.ones((8, 8, 8)), 0.0)
    assert ve.shape == (0, 3) and fe.shape == (0, 3)


def test_force_key_error_dict_and_dropin_module_paths():
    from hortimapping_b200.mesher import ForceKeyErrorDict
    d = ForceKeyErrorDict(vertices=1, faces=2)
    assert d.vertices == 1 and d["faces"] == 2
    with pytest.raises(KeyError):
        d.missing
    from hortimapping_b200 import dropin
    saved = {k: sys.modules.get(k) for k in ("wild_completion", "wild_completion.optimizer", "wild_completion.mesher",
                                            "wild_completion.loss", "deepsdf", "deepsdf.deep_sdf", "deepsdf.deep_sdf.workspace",
                                            "metrics_3d", "metrics_3d.chamfer_distance", "metrics_3d.precision_recall")}
    try:
        dropin.install(None)
        from wild_completion.optimizer import Optimizer
        from wild_completion.mesher import MeshExtractor
        from deepsdf.deep_sdf.workspace import config_decoder, load_latent_vectors
        import hortimapping_b200.optimizer as ho
        assert Optimizer is ho.Optimizer and callable(config_decoder) and callable(load_latent_vectors) and MeshExtractor
        from metrics_3d.chamfer_distance import ChamferDistance            # run_shape_completion_challenge.py:15-16
        from metrics_3d.precision_recall import PrecisionRecall
        import hortimapping_b200.metrics as hmx
        assert ChamferDistance is hmx.ChamferDistance and PrecisionRecall is hmx.PrecisionRecall
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_decoder_refuses_to_run_without_gpu():
    """The product path fails loudly instead of falling back to a CPU implementation."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hortimapping_b200.decoder import Decoder
    from tests.helpers import pepper_weights
    W, b, _ = pepper_weights()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Decoder(W, b)


def test_shard_ranges_cover_everything():
    from hortimapping_b200.shard import shard_range
    for n, w in ((512, 8), (64, 1), (10, 4), (3, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_metric_input_conversion_and_bookkeeping_without_gpu():
    """metrics_3d/metric.py:15-58 input handling and the reference's bookkeeping for empty predictions (no device call)."""
    from hortimapping_b200.metrics import ChamferDistance, Metrics3D, PrecisionRecall, _points

    class Pcd:                      # duck-typed open3d.geometry.PointCloud
        def __init__(self, p):
            self.points = p

    class Mesh:                     # duck-typed open3d.geometry.TriangleMesh: sampled with ITS OWN sample_points_uniformly
        vertices = [0, 1, 2]

        def sample_points_uniformly(self, n):
            assert n == 1000000     # metric.py:43
            return Pcd(np.ones((5, 3)))

    a = np.arange(12, dtype=np.float32).reshape(4, 3)
    assert _points(np.concatenate([a, a], 1)).shape == (4, 3) and _points(a).dtype == np.float64
    np.testing.assert_array_equal(_points(torch.from_numpy(a)), a.astype(np.float64))
    np.testing.assert_array_equal(_points(Pcd(a.tolist())), a.astype(np.float64))
    assert _points(Mesh()).shape == (5, 3)
    m = Metrics3D()
    assert m.prediction_is_empty(np.zeros((0, 3))) and not m.prediction_is_empty(a)
    assert m.prediction_is_empty(Pcd([])) and not m.prediction_is_empty(Mesh())
    cd, pr = ChamferDistance(), PrecisionRecall(0.001, 0.01, 10)
    cd.update(a, np.zeros((0, 3)))                                   # chamfer_distance.py:17-19
    pr.update(a, torch.zeros(0, 3))                                  # precision_recall.py:20-25
    assert cd.compute() == 0 and pr.compute_at_threshold(0.005)[:3] == (0, 0, 0)
    assert pr.find_nearest_threshold(0.0049) == pr.thresholds[4]


def _model_dirs():
    out = []
    for base in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        for model in ("sweetpepper_32", "strawberry_32"):
            d = os.path.join(base, "deepsdf", "models", model)
            if os.path.isfile(os.path.join(d, "ModelParameters", "latest.pth")):
                out.append((model, d))
        if out:
            break
    return out


@pytest.mark.parametrize("model,model_dir", _model_dirs() or [pytest.param(None, None, marks=pytest.mark.skip(reason="no shipped checkpoint reachable"))])
def test_product_checkpoint_loader_matches_the_reference_module(model, model_dir):
    """a2 (workspace.py:82-114,203-225): the product's own loader -- specs check, weight-norm folding, latent codes -- run on the
    SHIPPED .pth files must reproduce the weights the unmodified reference module materialises (tests/golden/<model>.npz was
    exported from `lin.weight` of the reference's Decoder after its weight_norm pre-forward hook, oracle/gen_golden.py)."""
    from hortimapping_b200.decoder import load_decoder_weights, load_latent_vectors
    from tests.helpers import load_npz
    W, b, specs = load_decoder_weights(model_dir, "latest")
    z = load_npz(model)
    assert specs["CodeLength"] == 32
    for l in range(9):
        assert W[l].dtype == np.float32 and W[l].shape == z[f"W{l}"].shape
        np.testing.assert_array_equal(W[l], z[f"W{l}"])                      # same op as torch's weight_norm hook: bit-identical
        np.testing.assert_array_equal(b[l], z[f"b{l}"])
    codes = load_latent_vectors(model_dir, "latest")
    np.testing.assert_array_equal(codes.numpy(), z["latent_codes"])


def test_checkpoint_loader_errors(tmp_path):
    from hortimapping_b200.decoder import load_decoder_weights, load_latent_vectors
    with pytest.raises(Exception, match="specs.json"):
        load_decoder_weights(str(tmp_path))
    import json
    json.dump({"CodeLength": 64, "NetworkSpecs": {"dims": [512] * 8, "latent_in": [4], "weight_norm": True}}, open(tmp_path / "specs.json", "w"))
    with pytest.raises(ValueError, match="shipped DeepSDF architecture"):
        load_decoder_weights(str(tmp_path))
    with pytest.raises(Exception, match="latent code file"):
        load_latent_vectors(str(tmp_path))


def test_calibration_rows_cover_the_optimisers_working_region():
    """decoder.calibration_rows: deterministic, [n][35] float32, latents on the segments between the mean training code (the initial
    latent of every fruit, run_shape_completion_challenge.py:51-52) and the training codes plus a little jitter, a quarter of them the
    training codes themselves; query points inside the requested cube."""
    from hortimapping_b200.decoder import calibration_rows
    codes = pepper_weights()[2]
    a = calibration_rows(codes, 0.15, n=4096)
    b = calibration_rows(torch.from_numpy(codes), 0.15, n=4096)
    assert a.dtype == torch.float32 and tuple(a.shape) == (4096, 35) and torch.equal(a, b)
    assert float(a[:, 32:].abs().max()) <= 0.15
    z, mean = a[:, :32].numpy(), codes.mean(0)
    # rows of the first quarter: a training code + N(0, 0.03) jitter
    d = np.abs(z[:1024, None, :] - codes[None, :, :]).max(-1).min(-1)
    assert d.max() < 0.2 and np.median(d) < 0.12
    # the rest: within the jitter of the segment [mean, code] -> never farther from the mean than the farthest code (+ jitter)
    r_codes = np.linalg.norm(codes - mean, axis=1).max()
    assert np.linalg.norm(z - mean, axis=1).max() < r_codes + 0.03 * 6 * np.sqrt(32)
    assert np.linalg.norm(z[1024:] - mean, axis=1).mean() < np.linalg.norm(z[:1024] - mean, axis=1).mean()


# ---------------------------------------------------------------- tensor-core engine: the plan and its stage program (host logic)
def _tc_plan(amask):
    import ctypes as C
    from hortimapping_b200 import _testing
    L = _testing.lib()
    info, rec = np.zeros(9, np.int32), np.zeros(480, np.uint32)
    ops, off = np.zeros((16, 9), np.uint8), np.zeros(16, np.int64)
    am = np.asarray(amask, np.uint8)
    assert L.hm_debug_tc_plan(am.ctypes.data, info.ctypes.data, rec.ctypes.data, ops.ctypes.data, off.ctypes.data) == 0
    keys = ("n_rec_fwd", "n_rec_all", "last_op_fwd", "last_op_jac", "sparse", "x0_chunk", "x0_early", "mask_layers", "mask_chunk")
    return dict(zip(keys, (int(v) for v in info))), rec, ops, off


def _groups(n_kchunks, n_nblocks):
    """(step, half) of the accumulation groups of an op in issue order."""
    if n_kchunks == 1:
        return [(0, g) for g in range(n_nblocks)]
    if n_nblocks == 1:
        return [(g, 0) for g in range(4)]
    return [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (3, 0), (2, 1), (3, 1)]


def _chunk_of(step, which):
    return (step & 1) + 4 * (step >> 1) + 2 * which


@pytest.mark.parametrize("amask", [
    [0xFF] * 8,                                                   # a model without dead units: the full plan
    [0xFF, 0x05, 0x01, 0x00, 0x1F, 0x1F, 0x07, 0x05],             # sweetpepper_32: 8 / 2 / 1 / 0 / 5 / 5 / 3 / 2 alive chunks, lin3 dead
    [0x7F, 0x0F, 0x05, 0x00, 0x3F, 0xFF, 0x0F, 0x01],             # lin3 dead, a layer that needs all eight chunks after it
    [0x1F, 0x07, 0x05, 0xFF, 0x1F, 0x05, 0x07, 0x05],             # sparse, lin3 alive: the gradient pass runs down to lin0
])
def test_tc_stage_program_is_the_plan_flattened(amask):
    """hm_tc_plan::rec (what the MMA issuer, the weight producer and the peer's forwarder walk) re-derived from the plan's masks:
    every issued stage once, in blob order, with the right A chunk, flags, A_READY requirement and blob offset; and the plan's
    chunk assignments (F0 operand, ReLU-bit home) never collide with a chunk some op of the window touches."""
    info, rec, ops, off = _tc_plan(amask)
    cut = amask[3] == 0
    assert info["last_op_fwd"] == 7 and info["last_op_jac"] == (11 if cut else 15)
    assert info["sparse"] == int(any(m != 0xFF for m in amask))
    want, n_fwd = [], None
    for op in range(16):
        if op == 8:
            n_fwd = len(want)
        nk, nb, rows64, cmask, hmask, gmask, need_out, valive, is_last = (int(v) for v in ops[op])
        assert (nk, nb, rows64) == ((1 if op == 0 else 8), (1 if op == 15 else 2), (1 if op == 15 else 4))
        groups = _groups(nk, nb)
        nwhich = 1 if nk == 1 else 2
        # the group mask is what the chunk / half masks imply
        for g, (step, nh) in enumerate(groups):
            chunks = [0] if nk == 1 else [_chunk_of(step, 0), _chunk_of(step, 1)]
            assert bool((gmask >> g) & 1) == bool(((hmask >> nh) & 1) and any((cmask >> c) & 1 for c in chunks))
        first_of_op = True
        stages_of_op = []
        for g, (step, nh) in enumerate(groups):
            if not (gmask >> g) & 1:
                continue
            st = []
            for part in range(2):
                for which in range(nwhich):
                    chunk = 0 if nk == 1 else _chunk_of(step, which)
                    if not (cmask >> chunk) & 1:
                        continue
                    sidx = (g * 2 + part) * nwhich + which
                    st.append(dict(op=op, g=g, chunk=(info["x0_chunk"] if op == 0 else chunk), part=part, narrow=rows64 == 1,
                                   src=(int(off[op]) + sidx * rows64 * 64 * 128) // 8192, first=False, last=False, need=0))
            st[0]["first"], st[-1]["last"] = True, True
            st[0]["need"] = (7 if g == 0 else 0) if op == 0 else step + 1
            assert st[0]["part"] == 0                              # the group's first MMA (which overwrites the buffer) is a lo-tile MMA
            stages_of_op += st
        if stages_of_op:
            stages_of_op[0]["op_first"], stages_of_op[-1]["op_last"] = True, True
        want += stages_of_op
    assert info["n_rec_fwd"] == n_fwd and info["n_rec_all"] == len(want) <= 480
    for r, w in zip(rec[:len(want)], want):
        r = int(r)
        got = dict(chunk=r & 7, part=(r >> 3) & 1, first=bool(r & 0x10), last=bool(r & 0x20), need=(r >> 6) & 7, op_first=bool(r & 0x200),
                   op_last=bool(r & 0x400), narrow=bool(r & 0x800), op=(r >> 12) & 15, src=(r >> 16) & 0xFFF, g=r >> 28)
        for k, v in got.items():
            assert v == w.get(k, False), (k, got, w)
    assert not rec[len(want):].any()
    # blob offsets: every stage of every group, lo and hi tiles, issued or not
    assert int(off[0]) == 0 and all(int(off[op + 1]) - int(off[op]) == (4 if op == 0 else 32) * 256 * 128 for op in range(15))
    # F0's operand chunk: free during the last op of a forward-only pass and of a forward + gradient pass, or chunk 0 / written at the tile start
    busy = int(ops[7][3]) | int(ops[info["last_op_jac"]][3])
    if info["x0_early"]:
        assert not (busy >> info["x0_chunk"]) & 1
    else:
        assert busy == 0xFF and info["x0_chunk"] == 0
    # ReLU bits the gradient pass reads, and their shared-memory home
    layers = [l for l in range(7) if int(ops[14 - l][5]) != 0 and not int(ops[14 - l][8])]
    assert info["mask_layers"] == sum(1 << l for l in layers)
    if info["mask_chunk"] >= 0:
        assert 0 < len(layers) <= 4 and info["mask_chunk"] != info["x0_chunk"]
        for op in range(min(layers), 14 - min(layers) + 1):
            touched = int(ops[op][3]) | (int(ops[op][6]) if int(ops[op][5]) else 0)
            assert not (touched >> info["mask_chunk"]) & 1, op
    if amask == [0xFF] * 8:
        assert info["mask_chunk"] == -1 and info["mask_layers"] == 0x7F and not info["x0_early"]
    if cut:
        assert layers == [4, 5, 6]


@pytest.mark.parametrize("model", ["sweetpepper_32", "strawberry_32"])
def test_shipped_models_get_the_fast_plan(model):
    """Both shipped DeepSDF models, calibrated on the product's own calibration rows (alive sets measured here with the fp32 oracle,
    on the device with the CUDA-core engine): lin3 is dead, so the gradient pass ends at lin4 and reads the ReLU bits of lin4..6
    only; the plan has a free chunk for the next tile's F0 operand; the issued tensor work is well below the dense count."""
    from hortimapping_b200.decoder import calibration_rows
    from oracle import hm_oracle as O
    z = load_npz(model)
    W, b = [z[f"W{l}"] for l in range(9)], [z[f"b{l}"] for l in range(9)]
    rows = calibration_rows(z["latent_codes"], 0.15, n=16384).numpy()
    _, (masks, _) = O.DecoderOracle(W, b, (4,), np.float32).forward(rows, return_cache=True)
    alive = [int(np.asarray(m).any(0).sum()) for m in masks[:8]]
    assert alive[3] == 0 and all(a > 0 for i, a in enumerate(alive) if i != 3), alive
    fill = [0, 2, 1, 3, 4, 6, 5, 7]                    # chunk fill order of the alive-first unit permutation (decoder_tc.cu kChunkFill)
    amask = []
    for l, a in enumerate(alive):
        c = 0 if l == 3 else max(1, -(-a // 64))
        amask.append(sum(1 << fill[i] for i in range(c)))
    info, rec, ops, off = _tc_plan(amask)
    assert info["sparse"] == 1 and info["last_op_jac"] == 11
    assert info["x0_early"] == 1 and info["mask_layers"] == 0b1110000
    # sweetpepper_32 leaves a chunk free for the ReLU bits of lin4..6; strawberry_32's lin4 keeps six chunks alive, its bits go to the
    # global scratch (same time)
    assert (info["mask_chunk"] >= 0) == (model == "sweetpepper_32"), (info, amask)
    # issued fp16 tensor work per row: every record is a (lo | hi) stage of 4 | 8 MMAs of 128 rows x stage_rows x 16
    def flop(lo, hi):
        return sum((8 if (int(r) >> 3) & 1 else 4) * 2.0 * (64 if int(r) & 0x800 else 256) * 16 for r in rec[lo:hi])
    fwd, jac = flop(0, info["n_rec_fwd"]), flop(0, info["n_rec_all"])
    dense_fwd = 3.0 * 2 * (64 * 512 + 7 * 512 * 512)          # three fp16 products per fp32 product, K of lin0 padded to 64
    assert fwd < 0.6 * dense_fwd and fwd < jac < 2.2 * fwd, (fwd, jac, dense_fwd)
    if model == "sweetpepper_32":
        assert [bin(m).count("1") for m in amask] == [8, 2, 1, 0, 5, 5, 3, 2] or sum(bin(m).count("1") for m in amask) <= 26, amask
