"""Deterministic synthetic fruit generator.

Produces exactly the inputs the reference's host scripts hand to the optimiser
(test_wild_completion.py:154-226): a world-frame surface point cloud, the `render_data` dict of
wild_completion/utils.py:41,96-105 (per-frame T_wc, fg/bg ray directions with z = 1, observed
z-depths), the initial latent (mean of the training codes,
run_shape_completion_challenge.py:51-52) and the initial pose.

The generator is decoder-agnostic: it takes an `sdf_jac(latent(32,), pts(n,3)) -> (sdf(n,),
dsdf_dxyz(n,3))` callable, so the product path feeds it the CUDA decoder and the CPU tests feed it
the oracle -- this module itself imports neither.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import numpy as np

SdfJac = Callable[[np.ndarray, np.ndarray], tuple]


def _rodrigues(aa: np.ndarray) -> np.ndarray:
    th = float(np.linalg.norm(aa))
    if th < 1e-12:
        return np.eye(3)
    k = aa / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


@dataclass
class SynthFruit:
    gt_latent: np.ndarray        # (32,)
    T_wo_gt: np.ndarray          # (4,4) object -> world (Sim3)
    points_w: np.ndarray         # (Np,3) float32
    render_data: Dict[str, list]
    init_latent: np.ndarray      # (32,) float32
    init_T_ow: np.ndarray        # (4,4) float32


def project_to_surface(sdf_jac: SdfJac, latent: np.ndarray, pts: np.ndarray, steps: int = 8) -> np.ndarray:
    p = pts.astype(np.float32).copy()
    for _ in range(steps):
        s, g = sdf_jac(latent, p)
        g2 = np.maximum((g * g).sum(-1, keepdims=True), 1e-12)
        p = (p - s[:, None] * g / g2).astype(np.float32)
    return p


def make_surface_points(sdf_jac: SdfJac, latent: np.ndarray, n_pts: int, rng: np.random.Generator,
                        half_extent: float = 0.07, max_radius: float = 0.075) -> np.ndarray:
    out = []
    have = 0
    for _ in range(8):
        cand = ((rng.random((6 * n_pts, 3)) * 2 - 1) * half_extent).astype(np.float32)
        p = project_to_surface(sdf_jac, latent, cand)
        s, _ = sdf_jac(latent, p)
        keep = (np.abs(s) < 1e-4) & (np.linalg.norm(p, axis=1) < max_radius)
        out.append(p[keep])
        have += int(keep.sum())
        if have >= n_pts:
            break
    p = np.concatenate(out, 0)
    if p.shape[0] < n_pts:       # degenerate decoders (random weights): pad by repetition
        reps = int(np.ceil(n_pts / max(p.shape[0], 1)))
        p = np.tile(p, (reps, 1)) if p.shape[0] else np.zeros((n_pts, 3), np.float32)
    return p[:n_pts]


def _look_at(cam_pos: np.ndarray, target: np.ndarray) -> np.ndarray:
    """Camera-to-world pose with +z looking at the target, +y down (pinhole convention)."""
    z = target - cam_pos
    z = z / np.linalg.norm(z)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(z, up)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = x, y, z, cam_pos
    return T


def sphere_trace(sdf_jac: SdfJac, latent: np.ndarray, T_oc: np.ndarray, dirs: np.ndarray,
                 obj_scale: float, steps: int = 64, bound_radius: float = 0.07, start_depth: float = 0.2):
    """March z-depth along camera rays `dirs` (z = 1) against the GT SDF given in the object frame.
    `T_oc` is one (4,4) pose or a per-ray stack (n,4,4) -- all frames of a fruit are marched in ONE batch, so the
    decoder is called `steps` times per fruit instead of `steps` x frames.  Returns (hit mask, z-depth)."""
    n = dirs.shape[0]
    dnorm = np.linalg.norm(dirs, axis=1)
    T_oc = np.asarray(T_oc)
    depth = np.full(n, start_depth, np.float64)
    hit = np.zeros(n, bool)
    alive = np.ones(n, bool)
    for _ in range(steps):
        c = dirs * depth[:, None]
        if T_oc.ndim == 2:
            p = c @ T_oc[:3, :3].T + T_oc[:3, 3]
        else:
            p = np.einsum("nij,nj->ni", T_oc[:, :3, :3], c) + T_oc[:, :3, 3]
        r = np.linalg.norm(p, axis=1)
        s, _ = sdf_jac(latent, p.astype(np.float32))
        s = np.where(r > bound_radius + 0.008, r - bound_radius, s.astype(np.float64))
        newly = alive & (np.abs(s) < 2e-4)
        hit |= newly
        alive &= ~newly
        alive &= depth < start_depth + 0.5
        step = 0.8 * s * obj_scale / dnorm
        depth = np.where(alive, depth + step, depth)
    return hit, depth.astype(np.float32)


def make_fruit(sdf_jac: SdfJac, latent_codes: np.ndarray, seed: int, index: int, n_pts: int = 2048,
               n_frames: int = 10, n_fg: int = 200, n_bg: int = 200, with_rays: bool = True,
               leaf_fraction: float = 0.0, noise_m: float = 0.0, lattice: int = 48,
               half_extent: float = 0.07, max_radius: float = 0.075) -> SynthFruit:
    """`half_extent` / `max_radius` bound the object (defaults: sweet pepper, ~ +-5 cm; strawberry: 0.03 / 0.035)."""
    rng = np.random.default_rng(seed * 100003 + index)
    codes = np.asarray(latent_codes, np.float32)
    gt_latent = codes[(7919 * index) % codes.shape[0]].copy()
    aa = (rng.random(3) * 2 - 1) * 0.25
    scale = 0.9 + 0.2 * rng.random()
    trans = (rng.random(3) * 2 - 1) * 0.01
    T_wo = np.eye(4)
    T_wo[:3, :3] = scale * _rodrigues(aa)
    T_wo[:3, 3] = trans
    T_ow_gt = np.linalg.inv(T_wo)

    pts_o = make_surface_points(sdf_jac, gt_latent, n_pts, rng, half_extent, max_radius)
    pts_w = pts_o.astype(np.float64) @ T_wo[:3, :3].T + T_wo[:3, 3]
    if noise_m > 0:
        pts_w = pts_w + rng.normal(0, noise_m, pts_w.shape)

    rd: Dict[str, list] = {"frame_id": [], "T_wc": [], "rays_fg": [], "rays_bg": [], "depth_fg": [],
                           "depth_bg": [], "pix_fg": [], "pix_bg": [], "count": 0}
    if with_rays:
        fx = 600.0
        u = np.linspace(-64, 64, lattice)
        uu, vv = np.meshgrid(u, u, indexing="xy")
        dirs = np.stack([uu.ravel() / fx, vv.ravel() / fx, np.ones(uu.size)], -1)
        poses = []
        for k in range(n_frames):
            az = -0.6 + 1.2 * (k / max(n_frames - 1, 1))
            cam = np.array([0.4 * np.cos(az), 0.4 * np.sin(az), 0.05])
            poses.append(_look_at(cam, np.zeros(3)))
        # all frames marched in one batch (the per-ray results do not depend on the batching)
        T_oc_all = np.repeat(np.stack([T_ow_gt @ T for T in poses]), dirs.shape[0], axis=0)
        hit_all, depth_all = sphere_trace(sdf_jac, gt_latent, T_oc_all, np.tile(dirs, (n_frames, 1)), scale,
                                          bound_radius=max_radius - 0.005)
        for k in range(n_frames):
            T_wc = poses[k]
            hit, depth = hit_all[k * dirs.shape[0]:(k + 1) * dirs.shape[0]], depth_all[k * dirs.shape[0]:(k + 1) * dirs.shape[0]]
            fg_idx, bg_idx = np.nonzero(hit)[0], np.nonzero(~hit)[0]
            fg_sel = rng.choice(fg_idx, size=min(n_fg, fg_idx.size), replace=False) if fg_idx.size else fg_idx
            bg_sel = rng.choice(bg_idx, size=min(n_bg, bg_idx.size), replace=False) if bg_idx.size else bg_idx
            d_bg = np.zeros(bg_sel.size, np.float32)
            if leaf_fraction > 0 and bg_sel.size:
                leaf = rng.random(bg_sel.size) < leaf_fraction
                d_bg[leaf] = (0.4 - (0.05 + 0.05 * rng.random(int(leaf.sum())))).astype(np.float32)
            rd["frame_id"].append(k)
            rd["T_wc"].append(T_wc.astype(np.float32))
            rd["rays_fg"].append(dirs[fg_sel].astype(np.float32))
            rd["rays_bg"].append(dirs[bg_sel].astype(np.float32))
            rd["depth_fg"].append(depth[fg_sel].astype(np.float32))
            rd["depth_bg"].append(d_bg)
            rd["count"] += 1
    return SynthFruit(gt_latent=gt_latent, T_wo_gt=T_wo.astype(np.float32),
                      points_w=pts_w.astype(np.float32), render_data=rd,
                      init_latent=codes.mean(0).astype(np.float32),
                      init_T_ow=np.eye(4, dtype=np.float32))
