"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference ships no tests or fixtures for this path, so these files are
the anchor that pins both the numpy oracle (oracle/hm_oracle.py) and the CUDA path.  Everything
here calls the reference's own functions through oracle/ref_shim.py; H, b and dx of the LM step
(wild_completion/optimizer.py:210-234) are observed by wrapping torch.inverse / torch.mv while the
reference runs -- the reference code itself is not modified.
"""
from __future__ import annotations

import copy
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
import torch  # noqa: E402
import yaml  # noqa: E402
from deepsdf.deep_sdf.workspace import config_decoder, load_latent_vectors  # noqa: E402
from wild_completion import loss as ref_loss  # noqa: E402
from wild_completion import utils as ref_utils  # noqa: E402
from wild_completion.optimizer import Optimizer  # noqa: E402

from hortimapping_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = ref_shim.REFERENCE_ROOT
torch.set_num_threads(8)


def load_model(name):
    d = os.path.join(REF, "deepsdf", "models", name)
    dec = config_decoder(d, "latest")
    codes = load_latent_vectors(d, "latest")
    return dec, codes


def export_weights(dec, codes, name):
    """Folded weights exactly as the reference module uses them: the weight_norm pre-forward hook
    materialises `lin.weight = g * v / ||v||` (deep_sdf_decoder.py:49-54)."""
    with torch.no_grad():
        dec(torch.zeros(2, 35))
    out = {"latent_codes": codes.numpy().astype(np.float32)}
    for l in range(9):
        lin = getattr(dec, f"lin{l}")
        out[f"W{l}"] = lin.weight.detach().numpy().astype(np.float32)
        out[f"b{l}"] = lin.bias.detach().numpy().astype(np.float32)
    specs = json.load(open(os.path.join(REF, "deepsdf", "models", name, "specs.json")))
    out["specs_json"] = np.frombuffer(json.dumps(specs).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), **out)


def make_sdf_jac(dec):
    def sdf_jac(latent, pts):
        y, g = ref_utils.get_batch_sdf_jacobian(dec, torch.from_numpy(np.asarray(latent, np.float32)),
                                                torch.from_numpy(np.asarray(pts, np.float32)))
        return y.reshape(-1).numpy(), g.reshape(-1, 35)[:, 32:].numpy()
    return sdf_jac


def to_torch_rd(rd, dtype=torch.float32):
    out = dict(rd)
    for k in ("T_wc", "rays_fg", "rays_bg", "depth_fg", "depth_bg"):
        out[k] = [torch.from_numpy(np.asarray(a)).to(dtype) for a in rd[k]]
    return out


class Capture:
    """Observe H / b / dx of optimizer.py:234 (`torch.mv(torch.inverse(H), b)`) from outside."""

    def __init__(self, est):
        self.est, self.H, self.b, self.dx = est, [], [], []

    def __enter__(self):
        self._inv, self._mv = torch.inverse, torch.mv
        cap = self

        def inv(x):
            if x.shape[0] == cap.est:
                cap.H.append(x.detach().clone().numpy())
            return cap._inv(x)

        def mv(a, b):
            r = cap._mv(a, b)
            if a.shape[0] == cap.est:
                cap.b.append(b.detach().clone().numpy())
                cap.dx.append(r.detach().clone().numpy())
            return r

        torch.inverse, torch.mv = inv, mv
        return self

    def __exit__(self, *a):
        torch.inverse, torch.mv = self._inv, self._mv


def run_joint(dec, cfg, fruit, cube_radius, pose_known, max_iter, dtype=torch.float32, zero_eps=True):
    cfg = copy.deepcopy(cfg)
    cfg["opt"]["converge"]["max_iter"] = max_iter
    if zero_eps:
        for k in ("epsilon_g", "epsilon_c", "epsilon_t", "epsilon_r", "epsilon_s"):
            cfg["opt"]["converge"][k] = 0
    opt = Optimizer(cfg, dec, None, None)
    opt.dtype = dtype
    latent = torch.from_numpy(fruit.init_latent.copy()).to(dtype)
    T_ow = torch.from_numpy(fruit.init_T_ow.copy()).to(dtype)
    rd = to_torch_rd(fruit.render_data, dtype)
    pts = torch.from_numpy(fruit.points_w).to(dtype)
    est = (7 if cfg["opt"]["scale_on"] else 6) + 32
    torch.set_default_dtype(dtype)          # the reference creates tensors with the default dtype
    try:
        with Capture(est) as cap:
            lat, T, it = opt.shape_pose_joint_opt(latent, T_ow, rd, pts, cube_radius, [0.5, 0.5, 0.5], pose_known)
    finally:
        torch.set_default_dtype(torch.float32)
    return lat.numpy().copy(), T.numpy().copy(), it, cap


def run_shape(dec, cfg, fruit, max_iter, dtype=torch.float32, zero_eps=True):
    cfg = copy.deepcopy(cfg)
    cfg["opt"]["converge"]["max_iter"] = max_iter
    if zero_eps:
        for k in ("epsilon_g", "epsilon_c"):
            cfg["opt"]["converge"][k] = 0
    opt = Optimizer(cfg, dec, None, None)
    opt.dtype = dtype
    latent = torch.from_numpy(fruit.init_latent.copy()).to(dtype)
    T_ow = torch.from_numpy(fruit.init_T_ow.copy()).to(dtype)
    pts = torch.from_numpy(fruit.points_w).to(dtype)
    torch.set_default_dtype(dtype)
    try:
        with Capture(32) as cap:
            lat, T, it = opt.shape_opt_deepsdf(latent, T_ow, pts, [0.5, 0.5, 0.5])
    finally:
        torch.set_default_dtype(torch.float32)
    return lat.numpy().copy(), it, cap


def pack_fruit(fruit):
    out = {"gt_latent": fruit.gt_latent, "T_wo_gt": fruit.T_wo_gt, "points_w": fruit.points_w,
           "init_latent": fruit.init_latent, "init_T_ow": fruit.init_T_ow,
           "n_frames": np.int32(len(fruit.render_data["T_wc"]))}
    for k in ("T_wc", "rays_fg", "rays_bg", "depth_fg", "depth_bg"):
        for i, a in enumerate(fruit.render_data[k]):
            out[f"rd_{k}_{i}"] = np.asarray(a, np.float32)
    return out


def frame_outputs(dec, cfg, fruit, cube_radius, latent=None, T_ow=None):
    """compute_render_loss per frame at the given state, called exactly as optimizer.py:104-118 does."""
    o = cfg["opt"]
    latent = torch.from_numpy(fruit.init_latent.copy() if latent is None else latent)
    T_ow = torch.from_numpy(fruit.init_T_ow.copy() if T_ow is None else T_ow)
    rd = to_torch_rd(fruit.render_data)
    out = {}
    cur_scale = torch.det(T_ow[:3, :3]) ** (-1 / 3)
    n = len(rd["T_wc"])
    ind = np.linspace(0, n - 1, min(o["render"]["n_frame"], n)).astype(np.int32)
    out["frame_ind"] = ind
    for j, idx in enumerate(ind):
        T_oc = T_ow @ rd["T_wc"][idx]
        T_co = torch.inverse(T_oc)
        depth_range = cube_radius * cur_scale
        dmin, dmax = T_co[2, 3] - 1.0 * depth_range, T_co[2, 3] + 0.8 * depth_range
        depths = torch.linspace(dmin, dmax, o["render"]["n_sample_on_ray"], dtype=torch.float32)
        rays = torch.cat((rd["rays_fg"][idx], rd["rays_bg"][idx]), 0)
        r = ref_loss.compute_render_loss(dec, latent, rays, rd["depth_fg"][idx].clone(), rd["depth_bg"][idx].clone(),
                                         T_oc, depths, o["scale_on"], o["render"]["log_sdf_occ"],
                                         float(o["render"]["occ_cutoff_m"]), depth_range, o["render"]["occlusion_on"])
        out[f"f{j}_depths"] = depths.numpy()
        out[f"f{j}_T_oc"] = T_oc.numpy()
        out[f"f{j}_none"] = np.bool_(r is None)
        if r is not None:
            for name, t in zip(("res_d", "J_d_pose", "J_d_code", "res_m", "J_m_pose", "J_m_code"), r):
                out[f"f{j}_{name}"] = t.detach().numpy()
    return out


def gen_fruit_case(dec, dec64, codes, name, cfg, seed, index, *, n_pts, n_frames, n_fg, n_bg, cube_radius,
                   pose_known, leaf, n_trace, n_final, with_shape_opt=False, variants=()):
    fruit = synth.make_fruit(make_sdf_jac(dec), codes.numpy(), seed, index, n_pts=n_pts, n_frames=n_frames,
                             n_fg=n_fg, n_bg=n_bg, leaf_fraction=leaf, noise_m=0.0005)
    out = pack_fruit(fruit)
    out["cfg_json"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    out["cube_radius"] = np.float32(cube_radius)
    out["pose_known"] = np.bool_(pose_known)
    # per-frame render outputs + recon outputs at the initial state
    out.update({f"it0_{k}": v for k, v in frame_outputs(dec, cfg, fruit, cube_radius).items()})
    T0 = torch.from_numpy(fruit.init_T_ow)
    pw = torch.from_numpy(fruit.points_w)
    pts_o = (pw[..., None, :] * T0[:3, :3]).sum(-1) + T0[:3, 3]
    res, jp, jc = ref_loss.compute_sdf_loss(dec, torch.from_numpy(fruit.init_latent), pts_o, cfg["opt"]["scale_on"])
    out["it0_recon_res"], out["it0_recon_J_pose"], out["it0_recon_J_code"] = res.numpy(), jp.numpy(), jc.numpy()
    # H, b, dx trace
    lat, T, it, cap = run_joint(dec, cfg, fruit, cube_radius, pose_known, n_trace)
    out["trace_H"], out["trace_b"], out["trace_dx"] = np.stack(cap.H), np.stack(cap.b), np.stack(cap.dx)
    out["trace_final_latent"], out["trace_final_T_ow"], out["trace_iters"] = lat, T, np.int32(it)
    # state after k = 1..n_trace iterations (identical start; the reference is deterministic on CPU), so
    # that every iteration of the trace can be replayed as ONE step from the reference's own state
    for k in range(1, n_trace + 1):
        lat, T, it, _ = run_joint(dec, cfg, fruit, cube_radius, pose_known, k)
        out[f"after{k}_latent"], out[f"after{k}_T_ow"] = lat, T
    # stop rules (optimizer.py:276-291): a huge epsilon must stop the loop at i == 2 (the `i > 1` guard)
    for ename in ("epsilon_g", "epsilon_c", "epsilon_s"):
        c2 = copy.deepcopy(cfg)
        for k in ("epsilon_g", "epsilon_c", "epsilon_t", "epsilon_r", "epsilon_s"):
            c2["opt"]["converge"][k] = 0
        c2["opt"]["converge"][ename] = 1e9
        if ename == "epsilon_s":
            c2["opt"]["converge"]["epsilon_t"] = 1e9
            c2["opt"]["converge"]["epsilon_r"] = 1e9
        lat, T, it, _ = run_joint(dec, c2, fruit, cube_radius, pose_known, 10, zero_eps=False)
        out[f"stop_{ename}_iters"] = np.int32(it)
    # render outputs at the state after 3 iterations (pose no longer identity)
    out.update({f"it3_{k}": v for k, v in frame_outputs(dec, cfg, fruit, cube_radius, out["after3_latent"], out["after3_T_ow"]).items()})
    # longer run, fp32 and fp64 reference
    lat, T, it, _ = run_joint(dec, cfg, fruit, cube_radius, pose_known, n_final)
    out["final_latent"], out["final_T_ow"], out["final_iters"] = lat, T, np.int32(it)
    lat, T, it, _ = run_joint(dec64, cfg, fruit, cube_radius, pose_known, n_final, dtype=torch.float64)
    out["final64_latent"], out["final64_T_ow"] = lat, T
    # natural convergence with the shipped epsilons
    lat, T, it, _ = run_joint(dec, cfg, fruit, cube_radius, pose_known, cfg["opt"]["converge"]["max_iter"], zero_eps=False)
    out["natural_latent"], out["natural_T_ow"], out["natural_iters"] = lat, T, np.int32(it)
    for vname, edit in variants:
        c2 = copy.deepcopy(cfg)
        edit(c2)
        lat, T, it, cap = run_joint(dec, c2, fruit, cube_radius, pose_known, 3)
        out[f"var_{vname}_cfg_json"] = np.frombuffer(json.dumps(c2).encode(), dtype=np.uint8)
        out[f"var_{vname}_H"], out[f"var_{vname}_b"], out[f"var_{vname}_dx"] = np.stack(cap.H), np.stack(cap.b), np.stack(cap.dx)
        out[f"var_{vname}_latent"], out[f"var_{vname}_T_ow"], out[f"var_{vname}_iters"] = lat, T, np.int32(it)
    if with_shape_opt:
        lat, it, cap = run_shape(dec, cfg, fruit, 8)
        out["shape_H"], out["shape_b"], out["shape_dx"] = np.stack(cap.H), np.stack(cap.b), np.stack(cap.dx)
        out["shape_latent8"] = lat
        lat, it, _ = run_shape(dec, cfg, fruit, 30)
        out["shape_latent30"] = lat
        lat, it, _ = run_shape(dec64, cfg, fruit, 30, dtype=torch.float64)
        out["shape64_latent30"] = lat
        lat, it, _ = run_shape(dec, cfg, fruit, cfg["opt"]["converge"]["max_iter"], zero_eps=False)
        out["shape_natural_latent"], out["shape_natural_iters"] = lat, np.int32(it)
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), **out)
    print(name, "done: trace iters", out["trace_iters"], "natural iters", out["natural_iters"])
    return fruit


def gen_decoder_rows(dec, dec64, codes):
    g = torch.Generator().manual_seed(1234)
    n = 3000
    z = codes[torch.randint(0, codes.shape[0], (n,), generator=g)] + 0.02 * torch.randn(n, 32, generator=g)
    x = (torch.rand(n, 3, generator=g) - 0.5) * 0.16
    rows = torch.cat([z, x], 1)
    with torch.no_grad():
        sdf2d = dec(rows)                       # (n,1): Decoder.forward with 2-D input (utils.py:165-166)
        sdf3d = dec(rows.unsqueeze(1))          # (n,1,1): 3-D input (utils.py:187-189)
    inp = rows.unsqueeze(1).clone().requires_grad_(True)
    y = dec(inp)
    jac = ref_utils.get_gradient(inp, y).detach()
    inp64 = rows.double().unsqueeze(1).clone().requires_grad_(True)
    y64 = dec64(inp64)
    jac64 = ref_utils.get_gradient(inp64, y64).detach()
    lat = codes.mean(0)
    sdf_b = ref_utils.decode_sdf(dec, lat, x)
    yb, gb = ref_utils.get_batch_sdf_jacobian(dec, lat, x)
    np.savez_compressed(os.path.join(GOLD, "decoder_rows.npz"), rows=rows.numpy(), sdf2d=sdf2d.numpy(),
                        sdf3d=sdf3d.numpy(), jac=jac.numpy(), sdf64=y64.detach().numpy(), jac64=jac64.numpy(),
                        lat=lat.numpy(), xyz=x.numpy(), decode_sdf=sdf_b.numpy(), batch_y=yb.numpy(), batch_g=gb.numpy())


def gen_misc(dec, codes):
    out = {}
    g = torch.Generator().manual_seed(7)
    # exp maps incl. the quirk branches (utils.py:220-324)
    xs = [torch.randn(7, generator=g) * 0.1 for _ in range(6)]
    xs.append(torch.tensor([0.01, -0.02, 0.03, 0., 0., 0., 0.05]))      # theta <= eps, s != 0
    xs.append(torch.tensor([0.01, -0.02, 0.03, 0., 0., 0., 0.]))        # theta <= eps, s == 0
    xs.append(torch.tensor([0.01, -0.02, 0.03, 0.1, -0.2, 0.05, -0.04]))  # negative s -> c = 0
    xs.append(torch.tensor([0.01, -0.02, 0.03, 0.1, -0.2, 0.05, 0.04]))
    out["exp_x"] = torch.stack(xs).numpy()
    out["exp_sim3"] = torch.stack([ref_utils.exp_sim3(x) for x in xs]).numpy()
    out["exp_se3"] = torch.stack([ref_utils.exp_se3(x[:6]) for x in xs]).numpy()
    # huber (utils.py:327-358) incl. an exact zero
    r = torch.randn(64, generator=g) * 0.03
    r[5] = 0.
    rr, w2 = ref_utils.get_robust_res(r.clone(), 0.02)
    out["huber_res"], out["huber_out"], out["huber_w2"] = r.numpy(), rr.numpy(), w2.numpy()
    # pose Jacobians (utils.py:197-276)
    p = torch.randn(16, 3, generator=g) * 0.05
    out["pj_pts"] = p.numpy()
    out["pj_sim3"] = ref_utils.get_points_to_pose_jacobian_sim3(p).numpy()
    out["pj_se3"] = ref_utils.get_points_to_pose_jacobian_se3(p).numpy()
    # linspace (optimizer.py:111)
    a, b = torch.tensor(0.3217), torch.tensor(0.4711)
    for m in (15, 20, 30):
        out[f"linspace_{m}"] = torch.linspace(a - 0.07, b + 0.051, m, dtype=torch.float32).numpy()
    out["linspace_ab"] = np.array([(a - 0.07).item(), (b + 0.051).item()], np.float32)
    # voxel grids (utils.py:542-562) and grid SDF (mesher.py:12-18)
    for n in (8, 20):
        out[f"grid_{n}"] = ref_utils.create_voxel_grid(n).numpy()
    g40 = ref_utils.create_voxel_grid(40).numpy()
    out["grid_40_sample_idx"] = np.arange(0, 64000, 97)
    out["grid_40_sample"] = g40[::97]
    lat = codes[17]
    pts = ref_utils.create_voxel_grid(20) * 0.08
    out["grid_lat"] = lat.numpy()
    out["grid_sdf_20"] = ref_utils.decode_sdf(dec, lat, pts).view(20, 20, 20).numpy()
    np.savez_compressed(os.path.join(GOLD, "misc.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    dec, codes = load_model("sweetpepper_32")
    dec64, _ = load_model("sweetpepper_32")
    dec64 = dec64.double()
    export_weights(dec, codes, "sweetpepper_32")
    gen_decoder_rows(dec, dec64, codes)
    gen_misc(dec, codes)

    wild = yaml.safe_load(open(os.path.join(REF, "configs", "wild_pepper.yaml")))
    chal = yaml.safe_load(open(os.path.join(REF, "configs", "shape_completion_challenge_pepper.yaml")))
    for c in (wild, chal):
        c["device"] = "cpu"
        c["vis"]["vis_on"] = False
        c["vis"]["log_on"] = False

    def v_se3(c): c["opt"]["scale_on"] = False
    def v_lmeye(c): c["opt"]["lm"]["lm_eye"] = True
    def v_linocc(c): c["opt"]["render"]["log_sdf_occ"] = False
    def v_noocc(c): c["opt"]["render"]["occlusion_on"] = False
    def v_gn(c): c["opt"]["lm"]["lm_on"] = False

    # wild_pepper.yaml sizes scaled down (F=4 frames, 100+100 rays, M=30, 512 points), 20 % "leaf" bg rays
    gen_fruit_case(dec, dec64, codes, "fruit_wild", wild, seed=1, index=3, n_pts=512, n_frames=4, n_fg=100,
                   n_bg=100, cube_radius=0.08, pose_known=False, leaf=0.2, n_trace=8, n_final=12,
                   with_shape_opt=True,
                   variants=(("se3", v_se3), ("lmeye", v_lmeye), ("linocc", v_linocc), ("noocc", v_noocc), ("gn", v_gn)))
    # shape_completion_challenge_pepper.yaml plumbing (BASELINE config 1): pose_known=True, linear
    # occupancy, no occlusion test, robust_iter=1
    gen_fruit_case(dec, dec64, codes, "fruit_challenge", chal, seed=2, index=5, n_pts=512, n_frames=5, n_fg=120,
                   n_bg=60, cube_radius=0.08, pose_known=True, leaf=0.0, n_trace=6, n_final=12)


if __name__ == "__main__":
    main()
