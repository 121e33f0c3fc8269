"""Host-side mirror of wild_completion/optimizer.py `Optimizer` on top of the C ABI.

Same constructor and call signatures as the reference (optimizer.py:17-25, :28, :306) so the
reference's host scripts (test_wild_completion.py:130,226; run_shape_completion_challenge.py:77,
216-218) keep working unchanged; the whole LM loop runs device-side in libhortimapping_b200.so.
Batched variants (`*_batch`) optimise many independent fruits in one call -- that is the new
surface the reference does not have (it loops over fruits in Python).
"""
from __future__ import annotations

import ctypes as C
import time
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import HM_LATENT, check
from .decoder import Decoder, _stream_ptr


def opt_params_from_cfg(opt_cfg: dict, iter_offset: int = 0, max_iter: Optional[int] = None) -> _lib.OptParams:
    """POD mirror of the reads at optimizer.py:31-53 (values are cast with float() like the reference,
    because YAML parses `5e-2` as a string)."""
    p = _lib.OptParams()
    cv, rd, lm, wt = opt_cfg["converge"], opt_cfg["render"], opt_cfg["lm"], opt_cfg["weight"]
    p.max_iter = int(cv["max_iter"] if max_iter is None else max_iter)
    p.epsilon_g, p.epsilon_c = float(cv["epsilon_g"]), float(cv["epsilon_c"])
    p.epsilon_t, p.epsilon_r, p.epsilon_s = float(cv.get("epsilon_t", 0)), float(cv.get("epsilon_r", 0)), float(cv.get("epsilon_s", 0))
    p.n_depth_samples = int(rd["n_sample_on_ray"])
    p.occ_cutoff_m = float(rd["occ_cutoff_m"])
    p.log_sdf_occ = int(bool(rd["log_sdf_occ"]))
    p.occlusion_on = int(bool(rd["occlusion_on"]))
    p.w_recon, p.w_depth, p.w_mask, p.w_codereg = (float(wt[k]) for k in ("w_recon", "w_depth", "w_mask", "w_codereg"))
    p.lm_on, p.lm_eye, p.lm_lambda_0 = int(bool(lm["lm_on"])), int(bool(lm["lm_eye"])), float(lm["lm_lambda_0"])
    p.robust_th_recon = float(opt_cfg["recon"]["robust_th_m"])
    p.robust_th_depth = float(rd["robust_th_m"])
    p.robust_iter = int(opt_cfg["robust_iter"])
    p.s_damp = float(lm["s_damp"])
    p.scale_on = int(bool(opt_cfg["scale_on"]))
    p.occlusion_th, p.min_valid_sample, p.min_grad_thre = 0.03, 100, 1e-6      # loss.py:11 defaults
    p.iter_offset = int(iter_offset)
    return p


def select_frames(n_frames: int, max_render_frame: int) -> np.ndarray:
    """optimizer.py:77-78."""
    return np.linspace(0, n_frames - 1, min(max_render_frame, n_frames)).astype(np.int32)


class PackedBatch:
    """Flat host-side arrays of a list of fruits, in the layout hm_fruit_batch wants."""

    def __init__(self, points: Sequence[np.ndarray], render_datas: Optional[Sequence[dict]], max_render_frame: int,
                 cube_radius: Sequence[float], pose_known: Sequence[bool]):
        nf = len(points)
        self.n_fruits = nf
        self.point_offsets = np.zeros(nf + 1, np.int64)
        self.point_offsets[1:] = np.cumsum([p.shape[0] for p in points])
        self.points = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float32).reshape(-1, 3) for p in points], 0))
        self.joint = render_datas is not None
        # own, contiguous copies: callers may pass scalars broadcast with stride 0, and the C side indexes [fruit]
        self.cube_radius = np.array(np.broadcast_to(np.asarray(cube_radius, np.float32), (nf,)), np.float32, order="C")
        self.pose_known = np.array(np.broadcast_to(np.asarray(pose_known, bool), (nf,)), np.uint8, order="C")
        if self.joint:
            T_wc, rays, dobs, n_fg, ray_counts = [], [], [], [], []
            self.frame_offsets = np.zeros(nf + 1, np.int32)
            for f, rd in enumerate(render_datas):
                idxs = select_frames(len(rd["T_wc"]), max_render_frame) if len(rd["T_wc"]) else []
                for idx in idxs:
                    a = lambda t: (t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)).astype(np.float32)
                    fg, bg = a(rd["rays_fg"][idx]).reshape(-1, 3), a(rd["rays_bg"][idx]).reshape(-1, 3)
                    T_wc.append(a(rd["T_wc"][idx]).reshape(16))
                    rays.append(np.concatenate([fg, bg], 0))                       # optimizer.py:113
                    dobs.append(np.concatenate([a(rd["depth_fg"][idx]).reshape(-1), a(rd["depth_bg"][idx]).reshape(-1)], 0))
                    n_fg.append(fg.shape[0])
                    ray_counts.append(fg.shape[0] + bg.shape[0])
                self.frame_offsets[f + 1] = len(T_wc)
            nfr = len(T_wc)
            self.T_wc = np.ascontiguousarray(np.stack(T_wc, 0)) if nfr else np.zeros((0, 16), np.float32)
            self.rays = np.ascontiguousarray(np.concatenate(rays, 0)) if nfr else np.zeros((0, 3), np.float32)
            self.depth_obs = np.ascontiguousarray(np.concatenate(dobs, 0)) if nfr else np.zeros((0,), np.float32)
            self.n_fg = np.asarray(n_fg, np.int32)
            self.ray_offsets = np.zeros(nfr + 1, np.int64)
            self.ray_offsets[1:] = np.cumsum(ray_counts)


class Optimizer(object):
    def __init__(self, cfg, decoder: Decoder, mesher, vis=None):
        self.dev = cfg['device']
        self.dtype = torch.float32
        self.opt_cfg = cfg['opt']
        self.decoder = decoder
        self.mesher = mesher
        self.vis = vis
        self.vis_pause_time = cfg['vis']['vis_pause_s']
        self.log_on = cfg['vis']['log_on']
        self.last_status: Optional[np.ndarray] = None
        self.last_iter_counts: Optional[np.ndarray] = None
        if not isinstance(decoder, Decoder):
            raise TypeError("hortimapping_b200.Optimizer needs a hortimapping_b200 Decoder (built by config_decoder)")

    # ------------------------------------------------------------------ device batch plumbing
    def _run(self, pk: PackedBatch, latents: torch.Tensor, T_ow: torch.Tensor, params: _lib.OptParams):
        dec = self.decoder
        dev = dec.device
        # the C ABI takes raw device pointers and updates them in place: anything but contiguous float32 tensors of the right
        # shape on the decoder's own GPU would be an illegal address or silently corrupted results
        for name, t, shape in (("latents", latents, (pk.n_fruits, HM_LATENT)), ("T_ow", T_ow, (pk.n_fruits, 4, 4))):
            if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device != dev:
                raise ValueError(f"{name} must be a CUDA tensor on {dev} (got {getattr(t, 'device', type(t))})")
            if t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
                raise ValueError(f"{name} must be a contiguous float32 tensor of shape {shape} (got {t.dtype}, {tuple(t.shape)}, "
                                 f"contiguous={t.is_contiguous()})")
        if getattr(pk, "_dev", None) is None:       # inputs are uploaded once per PackedBatch and stay resident
            t = lambda a: torch.from_numpy(a).to(dev)
            pk._dev = [t(pk.points)] + ([t(pk.T_wc), t(pk.rays), t(pk.depth_obs)] if pk.joint else [])
        keep = pk._dev
        b = _lib.FruitBatch()
        b.n_fruits = pk.n_fruits
        b.d_latents, b.d_T_ow = latents.data_ptr(), T_ow.data_ptr()
        b.d_points_w = keep[0].data_ptr()
        b.h_point_offsets = pk.point_offsets.ctypes.data
        iters = torch.zeros(pk.n_fruits, dtype=torch.int32, device=dev)
        status = torch.zeros(pk.n_fruits, dtype=torch.int32, device=dev)
        b.d_iter_count, b.d_status = iters.data_ptr(), status.data_ptr()
        if pk.joint:
            b.h_frame_offsets = pk.frame_offsets.ctypes.data
            b.d_T_wc, b.d_rays, b.d_depth_obs = keep[1].data_ptr(), keep[2].data_ptr(), keep[3].data_ptr()
            b.h_ray_offsets, b.h_n_fg = pk.ray_offsets.ctypes.data, pk.n_fg.ctypes.data
            b.h_cube_radius, b.h_pose_known = pk.cube_radius.ctypes.data, pk.pose_known.ctypes.data
            fn = dec._L.hm_optimize_joint
        else:
            fn = dec._L.hm_optimize_shape
        check(fn(dec.handle, C.byref(params), C.byref(b), _stream_ptr(dev)), fn.__name__)
        self._keep = keep
        return iters, status

    def last_system(self, n_fruits: int, joint: bool = True):
        """Test / diagnostics hook (hm_get_last_system): H (n,est,est), b (n,est), dx (n,est) of the LAST LM iteration run through
        this decoder's context, est = pose_dim + 32 for the joint loop and 32 for the latent-only loop."""
        dec = self.decoder
        est = ((7 if self.opt_cfg["scale_on"] else 6) if joint else 0) + HM_LATENT
        H = torch.empty(n_fruits, est, est, device=dec.device)
        b = torch.empty(n_fruits, est, device=dec.device)
        dx = torch.empty(n_fruits, est, device=dec.device)
        check(dec._L.hm_get_last_system(dec.handle, n_fruits, H.data_ptr(), b.data_ptr(), dx.data_ptr(), _stream_ptr(dec.device)), "hm_get_last_system")
        return H, b, dx

    def check_status(self, status: torch.Tensor) -> np.ndarray:
        """Opt-in synchronising check of a batched call's status words: warns on HM_STATUS_F16_SATURATED, prints the
        reference's messages for skipped frames / invalid submaps (optimizer.py:130-141).  Returns the words as numpy."""
        st = status.detach().cpu().numpy()
        self._report(st)
        return st

    def _report(self, status: np.ndarray):
        if any(int(s) & _lib.STATUS.get("F16_SATURATED", 0x40) for s in status):
            import warnings
            warnings.warn("hortimapping_b200: an operand left the calibrated fp16 range of the tensor-core decoder "
                          "(HM_STATUS_F16_SATURATED): results are not fp32-grade, call Decoder.calibrate() on representative rows")
        for f, s in enumerate(status):
            if s & _lib.STATUS["FRAME_SKIPPED"]:
                print("This frame is not valid")                       # optimizer.py:131
            if s & _lib.STATUS["SUBMAP_INVALID"]:
                print("This submap is not valid")                      # optimizer.py:140
            if self.log_on:
                for name, msg in (("CONV_GRADIENT", "gradient "), ("CONV_CODE", "Shape Latent Code"),
                                  ("CONV_POSE", "Pose Parameters"), ("MAX_ITER", "Maximum Iteration Numbers")):
                    if s & _lib.STATUS[name]:
                        print(f'**** Convergence in {msg} ****')

    # ------------------------------------------------------------------ batched API (new surface)
    def shape_pose_joint_opt_batch(self, latents: torch.Tensor, T_ow: torch.Tensor, render_datas: Sequence[dict],
                                   points_w: Sequence, cube_radius, pose_known=False, iter_offset: int = 0,
                                   max_iter: Optional[int] = None):
        """latents (n,32) and T_ow (n,4,4) are float32 CUDA tensors updated IN PLACE; returns
        (latents, T_ow, iter_counts int32 tensor, status int32 tensor) without synchronising.  The status words
        (HM_STATUS_* bits: fp16 saturation, invalid submap, skipped frame, stop reason) are the caller's to inspect --
        `check_status(status)` does it (synchronises) and prints / warns like the single-fruit calls."""
        n = latents.shape[0]
        if len(points_w) != n or len(render_datas) != n:
            raise ValueError(f"{n} latents but {len(points_w)} point clouds / {len(render_datas)} render_data dicts")
        cr = np.broadcast_to(np.asarray(cube_radius, np.float32), (n,))
        pkn = np.broadcast_to(np.asarray(pose_known, bool), (n,))
        pts = [p.detach().cpu().numpy() if isinstance(p, torch.Tensor) else np.asarray(p) for p in points_w]
        pk = PackedBatch(pts, render_datas, self.opt_cfg['render']['n_frame'], cr, pkn)
        params = opt_params_from_cfg(self.opt_cfg, iter_offset, max_iter)
        iters, status = self._run(pk, latents, T_ow, params)
        return latents, T_ow, iters, status

    def shape_opt_deepsdf_batch(self, latents: torch.Tensor, T_ow: torch.Tensor, points_w: Sequence, iter_offset: int = 0,
                                max_iter: Optional[int] = None):
        n = latents.shape[0]
        if len(points_w) != n:
            raise ValueError(f"{n} latents but {len(points_w)} point clouds")
        pts = [p.detach().cpu().numpy() if isinstance(p, torch.Tensor) else np.asarray(p) for p in points_w]
        pk = PackedBatch(pts, None, 0, np.zeros(n, np.float32), np.zeros(n, bool))
        params = opt_params_from_cfg(self.opt_cfg, iter_offset, max_iter)
        iters, status = self._run(pk, latents, T_ow, params)
        return latents, T_ow, iters, status

    # ------------------------------------------------------------------ reference signatures
    def _single(self, latent, T_ow_torch, render_data, points_w_torch, cube_radius, cur_color, pose_known, joint):
        dev = self.decoder.device
        lat = latent.detach().to(dev, torch.float32).reshape(1, HM_LATENT).contiguous().clone()
        T = T_ow_torch.detach().to(dev, torch.float32).reshape(1, 4, 4).contiguous().clone()
        max_iter = int(self.opt_cfg['converge']['max_iter'])
        if self.vis is None:
            if joint:
                _, _, iters, status = self.shape_pose_joint_opt_batch(lat, T, [render_data], [points_w_torch], cube_radius, pose_known)
            else:
                _, _, iters, status = self.shape_opt_deepsdf_batch(lat, T, [points_w_torch])
            iter_count = int(iters.item())
            st = status.cpu().numpy()
        else:
            # interactive path (optimizer.py:68-72,268-271): one device iteration per visualiser update
            T_wo = np.linalg.inv(T[0].cpu().numpy())
            self.vis.update_mesh_pose(self.mesher.complete_mesh(lat[0], np.eye(4), cur_color), T_wo, 0)
            time.sleep(self.vis_pause_time)
            iter_count, st = 0, np.zeros(1, np.int32)
            for i in range(max_iter):
                if joint:
                    _, _, iters, status = self.shape_pose_joint_opt_batch(lat, T, [render_data], [points_w_torch], cube_radius,
                                                                          pose_known, iter_offset=i, max_iter=1)
                else:
                    _, _, iters, status = self.shape_opt_deepsdf_batch(lat, T, [points_w_torch], iter_offset=i, max_iter=1)
                st = status.cpu().numpy()
                if int(iters.item()) == 0:
                    break
                iter_count = i + 1
                cur_T_wo = np.linalg.inv(T[0].cpu().numpy())
                self.vis.update_mesh_pose(self.mesher.complete_mesh(lat[0], np.eye(4), cur_color), cur_T_wo, i + 1)
                time.sleep(self.vis_pause_time)
                if st[0] & 0x07:
                    break
            if joint:
                self.vis.vis.remove_geometry(self.vis.txt, self.vis.reset_bounding_box)
                self.vis.stop()
        self.last_status, self.last_iter_counts = st, np.array([iter_count])
        self._report(st)
        with torch.no_grad():
            latent.copy_(lat[0].to(latent.device, latent.dtype))           # optimizer.py:248: in-place on the caller's tensor
        return latent, T[0].to(T_ow_torch.device, T_ow_torch.dtype), iter_count

    def shape_pose_joint_opt(self, latent, T_ow_torch, render_data, points_w_torch, cube_radius, cur_color, pose_known=False):
        """optimizer.py:28-302."""
        return self._single(latent, T_ow_torch, render_data, points_w_torch, cube_radius, cur_color, pose_known, True)

    def shape_opt_deepsdf(self, latent, T_ow_torch, points_w_torch, cur_color):
        """optimizer.py:306-429."""
        return self._single(latent, T_ow_torch, None, points_w_torch, 0.0, cur_color, False, False)


# ------------------------------------------------------------------------------------------------
# wild_completion/loss.py with the reference's signatures (single frame / single fruit)
# ------------------------------------------------------------------------------------------------
def compute_sdf_loss(decoder: Decoder, latent_vector, pts_surface_obj, scale_on=False):
    """loss.py:219-243 -> (res (N,1,1), J_pose (N,1,6|7), J_code (N,1,32))."""
    dev = decoder.device
    lat = latent_vector.detach().to(dev, torch.float32).contiguous()
    pts = pts_surface_obj.detach().to(dev, torch.float32).contiguous()
    n, pd = pts.shape[0], (7 if scale_on else 6)
    res = torch.empty(n, device=dev)
    jp = torch.empty(n, pd, device=dev)
    jc = torch.empty(n, HM_LATENT, device=dev)
    check(decoder._L.hm_sdf_loss(decoder.handle, lat.data_ptr(), pts.data_ptr(), n, int(bool(scale_on)), res.data_ptr(),
                                 jp.data_ptr(), jc.data_ptr(), _stream_ptr(dev)), "hm_sdf_loss")
    return res.view(n, 1, 1), jp.view(n, 1, pd), jc.view(n, 1, HM_LATENT)


def compute_render_loss(decoder: Decoder, latent_vector, ray_directions, depth_obs_fg, depth_obs_bg, t_obj_cam,
                        sampled_ray_depth, scale_on=False, log_occ_on=False, occupancy_th=0.01, object_bbx_radius=0.1,
                        occlusion_on=True, occlusion_th=0.03, min_valid_sample=100, min_grad_thre=1e-6):
    """loss.py:8-217: None when fewer than min_valid_sample samples fall in the object sphere, else the six
    per-ray tensors (res_d, J_d_pose, J_d_code, res_m, J_m_pose, J_m_code)."""
    dev = decoder.device
    f = lambda t: t.detach().to(dev, torch.float32).contiguous()
    lat, rays = f(latent_vector), f(ray_directions)
    dobs = torch.cat((f(depth_obs_fg).reshape(-1), f(depth_obs_bg).reshape(-1)), 0).contiguous()
    n_rays, n_fg = rays.shape[0], int(depth_obs_fg.shape[0])
    p = _lib.OptParams()
    p.n_depth_samples = int(sampled_ray_depth.shape[0])
    p.occ_cutoff_m, p.log_sdf_occ, p.occlusion_on, p.scale_on = float(occupancy_th), int(bool(log_occ_on)), int(bool(occlusion_on)), int(bool(scale_on))
    p.occlusion_th, p.min_valid_sample, p.min_grad_thre = float(occlusion_th), int(min_valid_sample), float(min_grad_thre)
    pd = 7 if scale_on else 6
    est = pd + HM_LATENT
    T = np.ascontiguousarray(t_obj_cam.detach().cpu().numpy().astype(np.float32).reshape(16))
    depths = np.ascontiguousarray(sampled_ray_depth.detach().cpu().numpy().astype(np.float32))
    valid = torch.empty(n_rays, dtype=torch.int32, device=dev)
    res_d, res_m = torch.empty(n_rays, device=dev), torch.empty(n_rays, device=dev)
    J_d, J_m = torch.empty(n_rays, est, device=dev), torch.empty(n_rays, est, device=dev)
    nv = C.c_int32(0)
    check(decoder._L.hm_render_loss(decoder.handle, C.byref(p), lat.data_ptr(), rays.data_ptr(), n_rays, n_fg, dobs.data_ptr(),
                                    T.ctypes.data_as(_lib.c_float_p), depths.ctypes.data_as(_lib.c_float_p),
                                    float(object_bbx_radius), valid.data_ptr(), res_d.data_ptr(), J_d.data_ptr(),
                                    res_m.data_ptr(), J_m.data_ptr(), C.byref(nv), _stream_ptr(dev)), "hm_render_loss")
    if nv.value < min_valid_sample:
        return None
    sel = valid > 0
    k = int(sel.sum().item())
    return (res_d[sel].view(k, 1, 1), J_d[sel][:, :pd].reshape(k, 1, pd), J_d[sel][:, pd:].reshape(k, 1, HM_LATENT),
            res_m[sel].view(k, 1, 1), J_m[sel][:, :pd].reshape(k, 1, pd), J_m[sel][:, pd:].reshape(k, 1, HM_LATENT))
