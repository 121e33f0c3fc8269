"""Runs the UNMODIFIED reference (PRBonn/HortiMapping) on given inputs -- the baseline arms of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE, never on the product path.  The reference's files are used where they lie: under
/root/reference in the build container, else under baseline/_ref/ (a byte-for-byte copy staged by
scripts/vendor_reference.py; git-ignored, travels to the GPU box).  They are imported through oracle/ref_shim.py (stub modules
for addict / plyfile / open3d / skimage; on the CPU arm the hard-coded `.cuda()` calls become no-ops and checkpoints load
with map_location='cpu').  Nothing of the reference is modified.

Timing sits where the reference's own t0 / t1 sit (run_shape_completion_challenge.py:213-220): around the
`opt.shape_opt_deepsdf(...)` / `opt.shape_pose_joint_opt(...)` call.
"""
from __future__ import annotations

import copy
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ["/root/reference", os.path.join(ROOT, "baseline", "_ref")]


def reference_root():
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "wild_completion", "optimizer.py")) and \
                os.path.isfile(os.path.join(c, "deepsdf", "models", "sweetpepper_32", "ModelParameters", "latest.pth")):
            return c
    return None


class ReferenceRunner:
    def __init__(self, device: str = "cpu", model: str = "sweetpepper_32", threads: int | None = None):
        root = reference_root()
        if root is None:
            raise RuntimeError("the unmodified reference is not available (neither /root/reference nor baseline/_ref)")
        from oracle import ref_shim
        ref_shim.install(root, force_cpu=(device == "cpu"))
        import torch
        if device == "cpu":
            torch.set_num_threads(threads or os.cpu_count())
        from deepsdf.deep_sdf.workspace import config_decoder, load_latent_vectors
        from wild_completion.optimizer import Optimizer
        self.torch, self.Optimizer = torch, Optimizer
        self.root, self.device = root, device
        d = os.path.join(root, "deepsdf", "models", model)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.decoder = config_decoder(d, "latest")
            self.codes = load_latent_vectors(d, "latest")
        if device != "cpu":
            self.decoder = self.decoder.cuda()
        self.threads = torch.get_num_threads()

    def _t(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32))).to(self.device)

    def _sync(self):
        if self.device != "cpu":
            self.torch.cuda.synchronize()

    def _opt(self, cfg, max_iter):
        cfg = copy.deepcopy(cfg)
        cfg["device"] = self.device
        cfg["vis"]["vis_on"] = False
        cfg["vis"]["log_on"] = False
        cfg["opt"]["converge"]["max_iter"] = int(max_iter)
        return self.Optimizer(cfg, self.decoder, None, None)

    def shape_opt(self, cfg, latent, T_ow, points_w, max_iter):
        """optimizer.py:306-429 -> (latent (32,), iter_count, seconds)."""
        opt = self._opt(cfg, max_iter)
        lat, T, pts = self._t(latent).clone(), self._t(T_ow), self._t(points_w)
        self._sync()
        t0 = time.perf_counter()
        lat, T, it = opt.shape_opt_deepsdf(lat, T, pts, [0.5, 0.5, 0.5])
        self._sync()
        dt = time.perf_counter() - t0
        return lat.detach().cpu().numpy().copy(), int(it), dt

    def observed_joint_iteration(self, cfg, latent, T_ow, render_data, points_w, cube_radius, pose_known):
        """ONE iteration of shape_pose_joint_opt observed from outside (the reference is not modified): H, b and dx of
        optimizer.py:234 by wrapping torch.inverse / torch.mv for the duration of the call, and the decoder rows evaluated
        (forward-only = the (n,35) no-grad calls of decode_sdf, utils.py:165-166; forward + gradient = the (n,1,35) calls of
        get_batch_sdf_jacobian, utils.py:187-189) through a forward hook on the decoder module."""
        torch = self.torch
        est = (7 if cfg["opt"]["scale_on"] else 6) + 32
        seen = {"H": None, "b": None, "dx": None, "rows_fwd": 0, "rows_jac": 0}
        inv0, mv0 = torch.inverse, torch.mv

        def inv(x):
            if x.shape[0] == est and seen["H"] is None:
                seen["H"] = x.detach().cpu().numpy().copy()
            return inv0(x)

        def mv(a, b):
            r = mv0(a, b)
            if a.shape[0] == est and seen["b"] is None:
                seen["b"], seen["dx"] = b.detach().cpu().numpy().copy(), r.detach().cpu().numpy().copy()
            return r

        def hook(_m, inp, _out):
            x = inp[0]
            if x.dim() == 3:
                seen["rows_jac"] += int(x.shape[0])
            else:
                seen["rows_fwd"] += int(x.shape[0])

        h = self.decoder.register_forward_hook(hook)
        torch.inverse, torch.mv = inv, mv
        try:
            lat, T, it, _ = self.joint_opt(cfg, latent, T_ow, render_data, points_w, cube_radius, pose_known, 1)
        finally:
            torch.inverse, torch.mv = inv0, mv0
            h.remove()
        seen.update(latent=lat, T_ow=T)
        return seen

    def joint_opt(self, cfg, latent, T_ow, render_data, points_w, cube_radius, pose_known, max_iter):
        """optimizer.py:28-302 -> (latent, T_ow, iter_count, seconds)."""
        opt = self._opt(cfg, max_iter)
        lat, T, pts = self._t(latent).clone(), self._t(T_ow), self._t(points_w)
        rd = dict(render_data)
        for k in ("T_wc", "rays_fg", "rays_bg", "depth_fg", "depth_bg"):
            rd[k] = [self._t(a) for a in render_data[k]]
        self._sync()
        t0 = time.perf_counter()
        lat, T, it = opt.shape_pose_joint_opt(lat, T, rd, pts, cube_radius, [0.5, 0.5, 0.5], pose_known)
        self._sync()
        dt = time.perf_counter() - t0
        return lat.detach().cpu().numpy().copy(), T.detach().cpu().numpy().copy(), int(it), dt
