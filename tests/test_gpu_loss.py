"""GPU parity tests of the two loss terms (wild_completion/loss.py) through the C ABI, fed the SAME
T_oc / depth samples the reference used (stored in the golden files), so that no threshold can flip."""
import numpy as np
import pytest
import torch

from tests.helpers import cfg_of, load_npz, render_data_of

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("engine", ["tc", "simt"])
@pytest.mark.parametrize("case_name", ["fruit_wild", "fruit_challenge"])
@pytest.mark.parametrize("state", ["it0", "it3"])
def test_render_loss_per_frame_vs_reference(case_name, state, engine):
    from hortimapping_b200.optimizer import compute_render_loss
    from tests.gpu_helpers import pepper_decoder
    dec = pepper_decoder()
    dec.set_engine(engine)
    try:
        c = load_npz(case_name)
        o = cfg_of(c)["opt"]
        rd = render_data_of(c)
        latent = c["init_latent"] if state == "it0" else c["after3_latent"]
        T_ow = c["init_T_ow"] if state == "it0" else c["after3_T_ow"]
        cur_scale = np.float32(np.linalg.det(T_ow[:3, :3])) ** np.float32(-1 / 3)
        n_checked = 0
        for j, idx in enumerate(c[f"{state}_frame_ind"]):
            rays = np.concatenate([rd["rays_fg"][idx], rd["rays_bg"][idx]], 0)
            r = compute_render_loss(dec, torch.from_numpy(latent).cuda(), torch.from_numpy(rays).cuda(),
                                    torch.from_numpy(rd["depth_fg"][idx]).cuda(), torch.from_numpy(rd["depth_bg"][idx]).cuda(),
                                    torch.from_numpy(c[f"{state}_f{j}_T_oc"]), torch.from_numpy(c[f"{state}_f{j}_depths"]),
                                    o["scale_on"], o["render"]["log_sdf_occ"], float(o["render"]["occ_cutoff_m"]),
                                    float(np.float32(c["cube_radius"]) * cur_scale), o["render"]["occlusion_on"])
            assert (r is None) == bool(c[f"{state}_f{j}_none"])
            if r is None:
                continue
            for name, t in zip(("res_d", "J_d_pose", "J_d_code", "res_m", "J_m_pose", "J_m_code"), r):
                ref = c[f"{state}_f{j}_{name}"]
                assert tuple(t.shape) == ref.shape, (name, tuple(t.shape), ref.shape)
                # the linear-occupancy config divides by (1 - o_k) -> 0 next to the band edge (loss.py:102,107)
                tol = 1e-4 if o["render"]["log_sdf_occ"] else 1e-3
                assert rel(t.cpu().numpy(), ref) < tol, (name, j, rel(t.cpu().numpy(), ref))
            n_checked += 1
        assert n_checked >= 3
    finally:
        dec.set_engine("tc")


def test_render_loss_returns_none_below_min_valid_samples():
    from hortimapping_b200.optimizer import compute_render_loss
    from tests.gpu_helpers import pepper_decoder
    dec = pepper_decoder()
    c = load_npz("fruit_wild")
    rd = render_data_of(c)
    rays = np.concatenate([rd["rays_fg"][0], rd["rays_bg"][0]], 0)
    r = compute_render_loss(dec, torch.from_numpy(c["init_latent"]).cuda(), torch.from_numpy(rays).cuda(),
                            torch.from_numpy(rd["depth_fg"][0]).cuda(), torch.from_numpy(rd["depth_bg"][0]).cuda(),
                            torch.from_numpy(c["it0_f0_T_oc"]), torch.from_numpy(c["it0_f0_depths"]), True, True, 0.01, 1e-4, True)
    assert r is None          # loss.py:43-45


@pytest.mark.parametrize("case_name", ["fruit_wild", "fruit_challenge"])
def test_sdf_loss_vs_reference(case_name):
    from hortimapping_b200.optimizer import compute_sdf_loss
    from tests.gpu_helpers import pepper_decoder
    dec = pepper_decoder()
    c = load_npz(case_name)
    cfg = cfg_of(c)
    T = c["init_T_ow"]
    pts_o = ((c["points_w"][..., None, :] * T[:3, :3]).sum(-1) + T[:3, 3]).astype(np.float32)
    res, jp, jc = compute_sdf_loss(dec, torch.from_numpy(c["init_latent"]).cuda(), torch.from_numpy(pts_o).cuda(), cfg["opt"]["scale_on"])
    for t, ref in ((res, c["it0_recon_res"]), (jp, c["it0_recon_J_pose"]), (jc, c["it0_recon_J_code"])):
        assert tuple(t.shape) == ref.shape
    np.testing.assert_allclose(res.cpu().numpy(), c["it0_recon_res"], rtol=1e-4, atol=1e-6)
    bad = (np.abs(jc.cpu().numpy() - c["it0_recon_J_code"]) > 1e-4 * np.abs(c["it0_recon_J_code"]) + 2e-6).reshape(len(pts_o), -1).any(1)
    assert bad.sum() <= 2, bad.sum()          # ReLU-kink rows, see test_gpu_decoder.assert_jac_close
    assert rel(jp.cpu().numpy()[~bad], c["it0_recon_J_pose"][~bad]) < 1e-4
