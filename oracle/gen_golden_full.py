"""Golden vectors at BASELINE.json's FULL joint sizes, from the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/gen_golden_full.py       # writes tests/golden/fruit_full.npz, strawberry_32.npz, fruit_berry.npz

`fruit_full`: one synthetic sweet pepper at the sizes north_star / SURVEY.md 8d name for the joint loop
(configs/wild_pepper.yaml: F = 10 frames x (200 fg + 200 bg) rays x M = 30 samples, 2048 surface points,
Sim(3) pose + latent, 20 % "leaf" background rays so the occlusion branch of loss.py:132-149 fires).  Kept
lean (no per-frame dumps): H / b / dx of the first iterations observed from outside (gen_golden.Capture),
the state after each of them so that every iteration can be replayed as ONE step from the reference's own
state, and fp32 / fp64 finals after 8 iterations.

`strawberry_32` / `fruit_berry`: the second shipped model (deepsdf/models/strawberry_32, ClampingDistance
0.05) with configs/lab_berry.yaml (M = 15, 400 + 200 rays, cube radius 0.04): folded weights, decoder rows
with Jacobians, one joint trace and the 80^3 mesher grid's SDF on a sub-sample.
"""
from __future__ import annotations

import copy
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gen_golden as G  # noqa: E402  (installs the import shim, imports the reference)

import torch  # noqa: E402
import yaml  # noqa: E402

from hortimapping_b200 import synth  # noqa: E402


def lean_case(dec, dec64, codes, name, cfg, seed, index, *, n_pts, n_frames, n_fg, n_bg, cube_radius, pose_known, leaf,
              n_trace, n_final, half_extent=0.07, max_radius=0.075):
    fruit = synth.make_fruit(G.make_sdf_jac(dec), codes.numpy(), seed, index, n_pts=n_pts, n_frames=n_frames, n_fg=n_fg,
                             n_bg=n_bg, leaf_fraction=leaf, noise_m=0.0005, half_extent=half_extent, max_radius=max_radius)
    out = G.pack_fruit(fruit)
    out["cfg_json"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    out["cube_radius"] = np.float32(cube_radius)
    out["pose_known"] = np.bool_(pose_known)
    # decoder rows per iteration, observed through a forward hook (the reference is not modified): (n,35) inputs are the no-grad
    # decode_sdf calls on the in-sphere ray samples (utils.py:165-166), (n,1,35) inputs the get_batch_sdf_jacobian calls on the
    # in-band samples and the surface points (utils.py:187-189).  The 39x39 torch.inverse is called once per iteration
    # (optimizer.py:234; the 4x4 inverses of optimizer.py:105 are ignored), so the counts are cut there.
    rows = {"fwd": 0, "jac": 0, "cuts": []}

    def hook(_m, inp, _o):
        x = inp[0]
        rows["jac" if x.dim() == 3 else "fwd"] += int(x.shape[0])

    h = dec.register_forward_hook(hook)
    inv0 = torch.inverse

    est = (7 if cfg["opt"]["scale_on"] else 6) + 32

    def inv(x):
        if x.shape[0] == est:
            rows["cuts"].append((rows["fwd"], rows["jac"]))
        return inv0(x)

    torch.inverse = inv
    try:
        lat, T, it, cap = G.run_joint(dec, cfg, fruit, cube_radius, pose_known, n_trace)
    finally:
        torch.inverse = inv0
        h.remove()
    cuts = np.array([(0, 0)] + rows["cuts"], np.int64)
    out["trace_rows_fwd"], out["trace_rows_jac"] = np.diff(cuts[:, 0]), np.diff(cuts[:, 1])
    out["trace_H"], out["trace_b"], out["trace_dx"] = np.stack(cap.H), np.stack(cap.b), np.stack(cap.dx)
    out["trace_iters"] = np.int32(it)
    out[f"after{n_trace}_latent"], out[f"after{n_trace}_T_ow"] = lat, T
    for k in range(1, n_trace):
        lat, T, it, _ = G.run_joint(dec, cfg, fruit, cube_radius, pose_known, k)
        out[f"after{k}_latent"], out[f"after{k}_T_ow"] = lat, T
    lat, T, it, _ = G.run_joint(dec, cfg, fruit, cube_radius, pose_known, n_final)
    out["final_latent"], out["final_T_ow"], out["final_iters"] = lat, T, np.int32(it)
    lat, T, it, _ = G.run_joint(dec64, cfg, fruit, cube_radius, pose_known, n_final, dtype=torch.float64)
    out["final64_latent"], out["final64_T_ow"] = lat, T
    np.savez_compressed(os.path.join(G.GOLD, f"{name}.npz"), **out)
    print(name, "done", {k: v.shape for k, v in out.items() if k.startswith("trace_")})


def berry_rows(dec, dec64, codes):
    g = torch.Generator().manual_seed(4321)
    n = 1500
    z = codes[torch.randint(0, codes.shape[0], (n,), generator=g)] + 0.02 * torch.randn(n, 32, generator=g)
    x = (torch.rand(n, 3, generator=g) - 0.5) * 0.08
    rows = torch.cat([z, x], 1)
    with torch.no_grad():
        sdf = dec(rows)
    inp = rows.unsqueeze(1).clone().requires_grad_(True)
    jac = G.ref_utils.get_gradient(inp, dec(inp)).detach()
    inp64 = rows.double().unsqueeze(1).clone().requires_grad_(True)
    y64 = dec64(inp64)
    jac64 = G.ref_utils.get_gradient(inp64, y64).detach()
    # mesher grid of configs/lab_berry.yaml: int(2 * 0.04 * 1e3 / 1.0) = 80, cube radius 0.04 (mesher.py:12-18); every 61st point
    lat = codes[11]
    grid = G.ref_utils.create_voxel_grid(80) * 0.04
    idx = np.arange(0, 80 ** 3, 61)
    sdf_g = G.ref_utils.decode_sdf(dec, lat, grid[idx])
    return {"rows": rows.numpy(), "sdf": sdf.numpy(), "jac": jac.numpy(), "sdf64": y64.detach().numpy(), "jac64": jac64.numpy(),
            "grid_lat": lat.numpy(), "grid_idx": idx, "grid_sdf": sdf_g.numpy()}


def main():
    which = set(sys.argv[1:]) or {"full", "berry"}
    if "full" in which:
        dec, codes = G.load_model("sweetpepper_32")
        dec64, _ = G.load_model("sweetpepper_32")
        dec64 = dec64.double()
        wild = yaml.safe_load(open(os.path.join(G.REF, "configs", "wild_pepper.yaml")))
        wild["device"] = "cpu"
        wild["vis"]["vis_on"] = False
        wild["vis"]["log_on"] = False
        lean_case(dec, dec64, codes, "fruit_full", wild, seed=3, index=11, n_pts=2048, n_frames=10, n_fg=200, n_bg=200,
                  cube_radius=0.08, pose_known=False, leaf=0.2, n_trace=4, n_final=8)
    if "berry" in which:
        dec, codes = G.load_model("strawberry_32")
        dec64, _ = G.load_model("strawberry_32")
        dec64 = dec64.double()
        G.export_weights(dec, codes, "strawberry_32")
        extra = berry_rows(dec, dec64, codes)
        path = os.path.join(G.GOLD, "strawberry_32.npz")
        with np.load(path) as z:
            merged = {k: z[k] for k in z.files}
        merged.update({f"rows_{k}": v for k, v in extra.items()})
        np.savez_compressed(path, **merged)
        berry = yaml.safe_load(open(os.path.join(G.REF, "configs", "lab_berry.yaml")))
        berry["device"] = "cpu"
        berry["vis"]["vis_on"] = False
        berry["vis"]["log_on"] = False
        r = float(berry["vis"]["object_radius_max_m"])
        lean_case(dec, dec64, codes, "fruit_berry", berry, seed=5, index=2, n_pts=512, n_frames=4, n_fg=120, n_bg=60,
                  cube_radius=r, pose_known=False, leaf=0.2, n_trace=4, n_final=8, half_extent=0.03, max_radius=0.035)


if __name__ == "__main__":
    main()
