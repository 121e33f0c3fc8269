"""Synthetic sequence in the on-disk layout of the reference's BUP20 example data (test_wild_completion.py:72-145):

    <root>/NNNNN_submap_id.png  NNNNN_depth.tiff  NNNNN_color.png  NNNNN_pose.txt      one set per frame
    <root>/submaps/00001_Background.ply  0000K_SweetPepper.ply                         one mesh per submap
    <root>/cam_info.yaml   <root>/config.yaml   <root>/gt.npz

TEST INFRASTRUCTURE (SURVEY.md 8c last row): the dataset itself cannot be downloaded here.  Fruits are DeepSDF shapes of the
shipped sweet-pepper model (training codes), posed with a small Sim(3) around the world axes (world: x right, y = viewing
direction, z up -- the convention get_pose_init's heuristics assume, utils.py:440-455); frames are rendered by sphere tracing
the SDF with the product decoder; a fruit's submap mesh is the camera-facing part of its surface; the background submap is a
wall behind the plants plus a small stem above each fruit (what the peduncle heuristic looks for).
"""
from __future__ import annotations

import os

import cv2
import numpy as np
import torch
import yaml


def _rodrigues(aa):
    th = float(np.linalg.norm(aa))
    if th < 1e-12:
        return np.eye(3)
    k = aa / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


R_WC = np.array([[1.0, 0, 0], [0, 0, 1.0], [0, -1.0, 0]])      # camera (x right, y down, z forward) -> world (x right, y forward, z up)


def make_sequence(root: str, dec, codes: np.ndarray, model_dir: str, n_fruits=3, n_frames=6, hw=(360, 480), seed=0, max_iter=30):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "standins"))
    import open3d as o3d                                   # the stand-in's PLY writer
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "submaps"), exist_ok=True)
    H, W = hw
    fx = 400.0
    K = np.array([[fx, 0, W / 2], [0, fx, H / 2], [0, 0, 1.0]])
    invK = np.linalg.inv(K)
    wall_y = 0.95
    # ---- fruits
    fruits = []
    for i in range(n_fruits):
        lat = codes[(97 * i + 13) % codes.shape[0]].astype(np.float32)
        scale = 0.85 + 0.25 * rng.random()
        R = _rodrigues((rng.random(3) * 2 - 1) * 0.15)
        t = np.array([-0.16 + 0.16 * i + 0.01 * rng.standard_normal(), 0.45 + 0.02 * rng.standard_normal(), 0.02 * rng.standard_normal()])
        T_wo = np.eye(4)
        T_wo[:3, :3], T_wo[:3, 3] = scale * R, t
        fruits.append({"id": i + 2, "latent": lat, "T_wo": T_wo, "T_ow": np.linalg.inv(T_wo), "scale": scale})
    # ---- cameras
    poses = []
    for k in range(n_frames):
        a = -1 + 2 * k / max(n_frames - 1, 1)
        c = np.array([0.10 * a, 0.0, 0.03 * np.cos(3 * a)])
        yaw = -0.12 * a                                     # turn towards the scene centre
        Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = Rz @ R_WC, c
        poses.append(T)
    # ---- render frames
    vv, uu = np.mgrid[0:H, 0:W]
    dirs_c = (np.stack([uu.ravel(), vv.ravel(), np.ones(H * W)], 1) @ invK.T)          # z = 1 camera rays
    for k, T_wc in enumerate(poses):
        Rw, cw = T_wc[:3, :3], T_wc[:3, 3]
        dw = dirs_c @ Rw.T
        depth = ((wall_y - cw[1]) / dw[:, 1]).astype(np.float64)                        # wall y = wall_y: z-depth along the z = 1 ray
        ids = np.ones(H * W, np.int32)
        for fr in fruits:
            T_oc = fr["T_ow"] @ T_wc
            # rays near the fruit's projection only
            pc = np.linalg.inv(T_wc) @ np.append(fr["T_wo"][:3, 3], 1.0)
            u0, v0, rad = fx * pc[0] / pc[2] + W / 2, fx * pc[1] / pc[2] + H / 2, fx * 0.075 / pc[2]
            sel = np.nonzero((np.abs(uu.ravel() - u0) < rad) & (np.abs(vv.ravel() - v0) < rad))[0]
            if not len(sel):
                continue
            hit, d = _sphere_trace(dec, fr["latent"], T_oc, dirs_c[sel], fr["scale"], start=float(pc[2]) - 0.1)
            closer = hit & (d < depth[sel])
            depth[sel[closer]] = d[closer]
            ids[sel[closer]] = fr["id"]
        stem = f"{k:05d}"
        cv2.imwrite(os.path.join(root, f"{stem}_submap_id.png"), ids.reshape(H, W).astype(np.uint16))
        cv2.imwrite(os.path.join(root, f"{stem}_depth.tiff"), depth.reshape(H, W).astype(np.float32))
        col = np.zeros((H, W, 3), np.uint8)
        col[ids.reshape(H, W) > 1] = (40, 40, 200)                                     # BGR red-ish fruits (visualisation only)
        cv2.imwrite(os.path.join(root, f"{stem}_color.png"), col)
        with open(os.path.join(root, f"{stem}_pose.txt"), "w") as fh:
            fh.write(" ".join(repr(float(x)) for x in T_wc.reshape(-1)))
    # ---- submaps
    cam_mid = poses[n_frames // 2][:3, 3]
    lat_dev = lambda l: torch.from_numpy(l).to(dec.device)
    bg_v, bg_f = [], []

    def add_box(lo, hi):
        lo, hi = np.asarray(lo, float), np.asarray(hi, float)
        c = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
        f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]])
        bg_f.append(f + sum(len(v) for v in bg_v))
        bg_v.append(c)

    add_box([-0.7, wall_y, -0.5], [0.7, wall_y + 0.01, 0.5])
    for fr in fruits:
        sdf = dec.sdf_grid(lat_dev(fr["latent"]), 64, 0.08)
        v, f = dec.isosurface(sdf, 0.0, 2.0 / 63, affine_radius=0.08)
        v, f = v.cpu().numpy().astype(np.float64), f.cpu().numpy()
        vw = v @ fr["T_wo"][:3, :3].T + fr["T_wo"][:3, 3]
        fr["surface_w"] = vw
        fn = np.cross(vw[f[:, 1]] - vw[f[:, 0]], vw[f[:, 2]] - vw[f[:, 0]])
        cen = vw[f].mean(1)
        to_cam = cam_mid - cen
        vis = (fn * to_cam).sum(1) > 0.15 * np.linalg.norm(fn, axis=1) * np.linalg.norm(to_cam, axis=1)
        fv = f[vis]
        used, inv = np.unique(fv.reshape(-1), return_inverse=True)
        m = o3d.geometry.TriangleMesh(vw[used] + rng.normal(0, 0.0004, (len(used), 3)), inv.reshape(-1, 3))
        m.paint_uniform_color([0.8, 0.15, 0.1])
        o3d.io.write_triangle_mesh(os.path.join(root, "submaps", f"{fr['id']:05d}_SweetPepper.ply"), m)
        top = fr["T_wo"][:3, 3] + np.array([0.004, 0.0, 0.075 * fr["scale"]])
        add_box(top - [0.004, 0.004, 0.0], top + [0.004, 0.004, 0.05])                  # a stem above the fruit
    bgm = o3d.geometry.TriangleMesh(np.concatenate(bg_v), np.concatenate(bg_f))
    bgm.paint_uniform_color([0.2, 0.5, 0.2])
    o3d.io.write_triangle_mesh(os.path.join(root, "submaps", "00001_Background.ply"), bgm)
    # ---- camera info, config (configs/wild_pepper.yaml with the paths / iteration count of this test, GUI off), ground truth
    with open(os.path.join(root, "cam_info.yaml"), "w") as fh:
        yaml.safe_dump({"intrinsics": K.tolist(), "extrinsics": np.eye(4).tolist(), "img_size": [H, W]}, fh)
    cfg = yaml.safe_load(open(os.path.join(os.path.dirname(os.path.dirname(model_dir.rstrip("/"))), "..", "configs", "wild_pepper.yaml")))
    cfg.update(deepsdf_dir=model_dir, data_dir=root, cam_info_path=os.path.join(root, "cam_info.yaml"), device="cuda")
    cfg["vis"].update(vis_on=False, log_on=False)
    cfg["opt"]["converge"]["max_iter"] = max_iter
    cfg_path = os.path.join(root, "config.yaml")
    with open(cfg_path, "w") as fh:
        yaml.safe_dump(cfg, fh)
    np.savez(os.path.join(root, "gt.npz"), ids=np.array([f["id"] for f in fruits]), T_wo=np.stack([f["T_wo"] for f in fruits]),
             latents=np.stack([f["latent"] for f in fruits]), **{f"surface_{f['id']}": f["surface_w"] for f in fruits})
    return cfg_path, fruits


def _sphere_trace(dec, latent, T_oc, dirs, obj_scale, start, steps=64):
    lat = torch.from_numpy(latent).to(dec.device)
    n = dirs.shape[0]
    dnorm = np.linalg.norm(dirs, axis=1)
    A, t = T_oc[:3, :3], T_oc[:3, 3]
    depth = np.full(n, start, np.float64)
    hit, alive = np.zeros(n, bool), np.ones(n, bool)
    for _ in range(steps):
        p = (dirs * depth[:, None]) @ A.T + t
        r = np.linalg.norm(p, axis=1)
        s = dec.sdf(lat, torch.from_numpy(p.astype(np.float32)).to(dec.device)).cpu().numpy().astype(np.float64)
        s = np.where(r > 0.078, r - 0.07, s)
        newly = alive & (np.abs(s) < 2e-4)
        hit |= newly
        alive &= ~newly
        alive &= depth < start + 0.4
        depth = np.where(alive, depth + 0.8 * s * obj_scale / dnorm, depth)
    return hit, depth
