"""CPU oracle: a numpy restatement of HortiMapping's per-fruit shape/pose inner loop.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this module, and there only as the checker or
as the timed CPU baseline.  The product path (hortimapping_b200/) never imports it and fails
loudly when its CUDA library is missing.

Parity status: PINNED.  Every function below is checked in tests/test_oracle_golden.py against
golden vectors produced by running the UNMODIFIED reference (imported from /root/reference through
oracle/ref_shim.py) in the build container; the generating script is oracle/gen_golden.py and the
vectors live in tests/golden/.  The reference itself ships no tests or golden vectors
(SURVEY.md section 4), so the reference run is the anchor.

Each function cites the reference file:line it restates (paths relative to the reference root).
All arithmetic runs in `dtype` (np.float32 mirrors the reference; np.float64 is the high-precision
oracle used for trajectory-level checks).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

LATENT = 32
N_IN = LATENT + 3


# --------------------------------------------------------------------------------------
# Decoder  (deepsdf/networks/deep_sdf_decoder.py:10-110, eval mode)
# --------------------------------------------------------------------------------------
def fold_weight_norm(v: np.ndarray, g: np.ndarray) -> np.ndarray:
    """W = g * v / ||v||_row : torch.nn.utils.weight_norm with dim=0, as applied to lin0..lin7
    (deep_sdf_decoder.py:49-54).  torch evaluates `v * (g / norm_except_dim(v, 2, 0))` in fp32."""
    v = np.asarray(v)
    g = np.asarray(g).reshape(-1, 1)
    norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(v.dtype)
    return (v * (g / norm)).astype(v.dtype)


# Matrix product used by the decoder restatement.  numpy's by default (deterministic across boxes: what the tests pin);
# bench.py's CPU arm swaps in torch's (the BLAS the reference itself runs on, several times faster than numpy's here).
_mm = np.matmul


def set_matmul(fn=None):
    """fn(a, b) -> a @ b for 2-D float arrays; None restores numpy's."""
    global _mm
    _mm = np.matmul if fn is None else fn


@dataclass
class DecoderOracle:
    """Weights are the folded (weight-norm applied) matrices W_l [out, in] and biases b_l.

    forward():  deep_sdf_decoder.py:75-110 -- x -> lin0..lin8 with ReLU after lin0..lin7, the raw
    input concatenated AFTER the activations before layer `latent_in` (:87-88), final tanh (:107).
    Dropout is inactive in eval mode (:104-105); weight_norm=True means no LayerNorm (:58-63).
    """
    weights: List[np.ndarray]
    biases: List[np.ndarray]
    latent_in: Tuple[int, ...] = (4,)
    dtype: type = np.float32

    def __post_init__(self):
        self.weights = [np.ascontiguousarray(w, dtype=self.dtype) for w in self.weights]
        self.biases = [np.ascontiguousarray(b, dtype=self.dtype) for b in self.biases]
        self._wt = [np.ascontiguousarray(w.T) for w in self.weights]

    @property
    def n_layers(self) -> int:
        return len(self.weights)

    def astype(self, dtype) -> "DecoderOracle":
        return DecoderOracle(self.weights, self.biases, self.latent_in, dtype)

    def forward(self, inp: np.ndarray, return_cache: bool = False):
        inp = np.asarray(inp, dtype=self.dtype)
        shape = inp.shape
        x0 = inp.reshape(-1, shape[-1])
        h = x0
        masks = []
        nl = self.n_layers
        for l in range(nl):
            if l in self.latent_in:
                h = np.concatenate([h, x0], axis=-1)
            h = _mm(h, self._wt[l]) + self.biases[l]
            if l < nl - 1:
                m = h > 0
                h = np.where(m, h, 0).astype(self.dtype)
                if return_cache:
                    masks.append(m)
        y = np.tanh(h)
        out = y.reshape(shape[:-1] + (1,))
        if return_cache:
            return out, (masks, y)
        return out

    def forward_jac(self, inp: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """SDF and d sdf / d input per row: what get_gradient (wild_completion/utils.py:112-122)
        obtains through torch.autograd.grad with grad_outputs = 1."""
        inp = np.asarray(inp, dtype=self.dtype)
        x0 = inp.reshape(-1, inp.shape[-1])
        y, (masks, yt) = self.forward(x0, return_cache=True)
        nl = self.n_layers
        d = (1 - yt * yt).astype(self.dtype)                    # tanh'
        g_skip = np.zeros_like(x0)
        for l in range(nl - 1, -1, -1):
            d = _mm(d, self.weights[l])                         # back through lin_l
            if l in self.latent_in:
                nin = x0.shape[1]
                g_skip = g_skip + d[:, -nin:]
                d = d[:, :-nin]
            if l > 0:
                d = np.where(masks[l - 1], d, 0).astype(self.dtype)
        g = d + g_skip
        return y.reshape(-1, 1), g


def decode_sdf(dec: DecoderOracle, lat: np.ndarray, x: np.ndarray, max_batch: int = 64 ** 3) -> np.ndarray:
    """wild_completion/utils.py:144-172: no-grad forward with the latent broadcast to every row,
    evaluated in chunks of `max_batch` rows; returns (N,)."""
    x = np.asarray(x, dtype=dec.dtype)
    lat = np.asarray(lat, dtype=dec.dtype)
    out = []
    for head in range(0, x.shape[0], max_batch):
        xs = x[head:head + max_batch, 0:3]
        inp = np.concatenate([np.broadcast_to(lat, (xs.shape[0], lat.shape[0])), xs], axis=-1)
        out.append(dec.forward(inp).reshape(-1))
    return np.concatenate(out, 0) if out else np.zeros((0,), dec.dtype)


def get_batch_sdf_jacobian(dec: DecoderOracle, lat: np.ndarray, x: np.ndarray):
    """wild_completion/utils.py:175-193: returns y (n,1,1) and g (n,1,code_len+3)."""
    x = np.asarray(x, dtype=dec.dtype)
    lat = np.asarray(lat, dtype=dec.dtype)
    n = x.shape[0]
    inp = np.concatenate([np.broadcast_to(lat, (n, lat.shape[0])), x], axis=1)
    y, g = dec.forward_jac(inp)
    return y.reshape(n, 1, 1), g.reshape(n, 1, -1)


# --------------------------------------------------------------------------------------
# Lie-algebra helpers (wild_completion/utils.py:197-324, 360-369)
# --------------------------------------------------------------------------------------
def points_to_pose_jacobian(points: np.ndarray, scale_on: bool) -> np.ndarray:
    """[I3 | -hat(p) | p] (sim3, utils.py:257-276) or [I3 | -hat(p)] (se3, utils.py:197-217).
    Left perturbation; parameter order translation, rotation, scale."""
    p = np.asarray(points)
    n = p.shape[0]
    dt = p.dtype
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    zero = np.zeros(n, dt)
    # torch.stack([...rows...], dim=-1) builds the matrix column by column (utils.py:208-213)
    col0 = np.stack([zero, -z, y], axis=-1)
    col1 = np.stack([z, zero, -x], axis=-1)
    col2 = np.stack([-y, x, zero], axis=-1)
    negate_hat = np.stack([col0, col1, col2], axis=-1)
    eye = np.broadcast_to(np.eye(3, dtype=dt), (n, 3, 3))
    parts = [eye, negate_hat]
    if scale_on:
        parts.append(p[..., None])
    return np.concatenate(parts, axis=-1)


def _hat(w):
    return np.array([[0., -w[2], w[1]], [w[2], 0., -w[0]], [-w[1], w[0], 0.]], dtype=w.dtype)


def exp_se3(x: np.ndarray) -> np.ndarray:
    """wild_completion/utils.py:220-254."""
    dt = x.dtype.type
    v, w = x[:3], x[3:6]
    w_hat = _hat(w)
    w_hat2 = w_hat @ w_hat
    theta = dt(np.sqrt((w * w).sum(dtype=x.dtype)))
    eye = np.eye(3, dtype=x.dtype)
    if theta <= 1e-8:
        e_w, j = eye, eye
    else:
        theta_2, theta_3 = theta ** 2, theta ** 3
        s, c = dt(np.sin(theta)), dt(np.cos(theta))
        e_w = eye + w_hat * s / theta + w_hat2 * (dt(1.) - c) / theta_2
        k1 = (dt(1) - c) / theta_2
        k2 = (theta - s) / theta_3
        j = eye + k1 * w_hat + k2 * w_hat2
    rst = np.eye(4, dtype=x.dtype)
    rst[:3, :3] = e_w
    rst[:3, 3] = j @ v
    return rst


def exp_sim3(x: np.ndarray) -> np.ndarray:
    """wild_completion/utils.py:279-324, quirks kept: `c = 0 if s <= eps` inside the theta > eps
    branch (:314) makes c vanish for every non-positive log-scale step."""
    dt = x.dtype.type
    v, w, s = x[:3], x[3:6], dt(x[6])
    w_hat = _hat(w)
    w_hat2 = w_hat @ w_hat
    theta = dt(np.sqrt((w * w).sum(dtype=x.dtype)))
    theta_2 = theta ** 2
    sin_t, cos_t = dt(np.sin(theta)), dt(np.cos(theta))
    e_s = dt(np.exp(s))
    s_2 = s ** 2
    eye = np.eye(3, dtype=x.dtype)
    eps = 1e-8
    if theta <= 1e-8:
        e_w = eye
        if s == 0:
            j = eye
        else:
            c = (e_s - dt(1.)) / s
            j = c * eye
    else:
        e_w = eye + w_hat * sin_t / theta + w_hat2 * (dt(1.) - cos_t) / theta_2
        a = e_s * sin_t
        b = e_s * cos_t
        c = dt(0.) if s <= eps else (e_s - dt(1.)) / s
        k_0 = c * eye
        k_1 = (a * s + (dt(1) - b) * theta) / (s_2 + theta_2)
        k_2 = c - ((b - dt(1)) * s + a * theta) / (s_2 + theta_2)
        j = k_0 + k_1 * w_hat / theta + k_2 * w_hat2 / theta_2
    rst = np.eye(4, dtype=x.dtype)
    rst[:3, :3] = e_s * e_w
    rst[:3, 3] = j @ v
    return rst


def rotation_matrix_to_axis_angle(R: np.ndarray):
    """wild_completion/utils.py:360-369: acos((trace-1)/2); NaN when the argument leaves [-1,1]."""
    dt = R.dtype.type
    with np.errstate(invalid="ignore"):
        return dt(np.arccos((np.trace(R) - dt(1)) / dt(2)))


# --------------------------------------------------------------------------------------
# Robust kernel (wild_completion/utils.py:327-358)
# --------------------------------------------------------------------------------------
def huber_norm_weights(x: np.ndarray, b: float) -> np.ndarray:
    """utils.py:327-340.  A residual norm of exactly 0 gets weight 0 (sqrt(0)/1)."""
    x = x.copy()
    dt = x.dtype.type
    b = dt(b)
    res_norm = np.zeros_like(x)
    m = x <= b
    res_norm[m] = x[m] ** 2
    res_norm[~m] = dt(2) * b * x[~m] - b ** 2
    x[x == 0] = dt(1.)
    return np.sqrt(res_norm) / x


def get_robust_res(res: np.ndarray, b: float):
    """utils.py:343-358: returns (w * res, w**2)."""
    res = res.reshape(-1, 1, 1)
    w = huber_norm_weights(np.abs(res), b)
    return w * res, w ** 2


# --------------------------------------------------------------------------------------
# Occupancy (wild_completion/utils.py:125-142)
# --------------------------------------------------------------------------------------
def sdf_to_occupancy(sdf: np.ndarray, th: float) -> np.ndarray:
    dt = sdf.dtype.type
    return dt(0.5) - np.clip(sdf, dt(-th), dt(th)) / dt(2 * th)


def sdf_to_occupancy_log(sdf: np.ndarray, sigma: float) -> np.ndarray:
    dt = sdf.dtype.type
    z = -sdf / dt(sigma)
    return (dt(1) / (dt(1) + np.exp(-z))).astype(sdf.dtype)


def torch_linspace(start, end, steps: int, dtype) -> np.ndarray:
    """torch.linspace on CPU/CUDA for floating types (optimizer.py:111): symmetric two-sided
    evaluation -- step = (end-start)/(steps-1); i < steps//2 ? start + i*step : end - (steps-1-i)*step."""
    dt = np.dtype(dtype).type
    start, end = dt(start), dt(end)
    if steps == 1:
        return np.array([start], dtype=dtype)
    step = (end - start) / dt(steps - 1)
    i = np.arange(steps)
    half = steps // 2
    lo = start + step * i.astype(dtype)
    hi = end - step * (steps - 1 - i).astype(dtype)
    return np.where(i < half, lo, hi).astype(dtype)


# --------------------------------------------------------------------------------------
# Loss terms (wild_completion/loss.py)
# --------------------------------------------------------------------------------------
def compute_sdf_loss(dec: DecoderOracle, latent: np.ndarray, pts_obj: np.ndarray, scale_on: bool):
    """loss.py:219-243: residual = SDF at the observed surface points; J_pose = dsdf/dxyz . D(p);
    J_code = dsdf/dlatent."""
    res, de_di = get_batch_sdf_jacobian(dec, latent, pts_obj)
    de_dxo = de_di[..., -3:]
    dxo = points_to_pose_jacobian(np.asarray(pts_obj, dtype=dec.dtype), scale_on)
    jac_tow = np.matmul(de_dxo, dxo)
    jac_code = de_di[..., :-3]
    return res, jac_tow, jac_code


def compute_render_loss(dec: DecoderOracle, latent, ray_directions, depth_obs_fg, depth_obs_bg,
                        t_obj_cam, sampled_ray_depth, scale_on=False, log_occ_on=False,
                        occupancy_th=0.01, object_bbx_radius=0.1, occlusion_on=True,
                        occlusion_th=0.03, min_valid_sample=100, min_grad_thre=1e-6,
                        return_debug: bool = False):
    """loss.py:8-217.  Returns None when fewer than `min_valid_sample` ray samples fall inside the
    object sphere (:43-45); otherwise (res_d, J_d_pose, J_d_code, res_m, J_m_pose, J_m_code) with one
    row per ray that keeps at least one in-band sample, rays in ascending index (fg first)."""
    dt = dec.dtype
    f = np.dtype(dt).type
    rays = np.asarray(ray_directions, dtype=dt)
    depth_obs = np.concatenate([np.asarray(depth_obs_fg, dt), np.asarray(depth_obs_bg, dt)], 0)
    n_fg = int(np.asarray(depth_obs_fg).shape[0])
    T = np.asarray(t_obj_cam, dtype=dt)
    d = np.asarray(sampled_ray_depth, dtype=dt)
    n_rays, n_depths = rays.shape[0], d.shape[0]

    pts_cam = rays[:, None, :] * d[:, None]                                          # :30
    pts_obj = (pts_cam[..., None, :] * T[:3, :3]).sum(-1) + T[:3, 3]                 # :32-33
    nrm = np.sqrt((pts_obj * pts_obj).sum(-1))
    valid = nrm < f(object_bbx_radius)                                               # :38
    vx, vy = np.nonzero(valid)                                                       # row-major
    query = pts_obj[vx, vy, :]
    if query.shape[0] < min_valid_sample:                                            # :43-45
        return None
    sdf = decode_sdf(dec, latent, query)                                             # :49

    occ = np.zeros((n_rays, n_depths), dt)                                           # :55
    if log_occ_on:
        sigma = f(occupancy_th) / f(3) * f(0.55)                                     # :59-60
        occ[vx, vy] = sdf_to_occupancy_log(sdf, sigma)
    else:
        sigma = None
        occ[vx, vy] = sdf_to_occupancy(sdf, occupancy_th)

    with_grad = (sdf > f(-occupancy_th)) & (sdf < f(occupancy_th))                   # :66
    gx, gy = vx[with_grad], vy[with_grad]

    occ_g = occ[gx, :]                                                               # :71
    k = occ_g.shape[0]
    d_min, d_max = d[0], d[-1]
    delta_d = (d_max - d_min) / f(n_depths - 1)                                      # :75
    d_term = d_max + delta_d                                                         # :78

    acc = np.cumprod(f(1) - occ_g, axis=-1, dtype=dt)                                # :81
    acc_aug = np.concatenate([np.ones((k, 1), dt), acc], -1)
    o_aug = np.concatenate([occ_g, np.ones((k, 1), dt)], -1)
    d_aug = np.concatenate([d, np.array([d_term], dt)], -1)
    term_prob = o_aug * acc_aug                                                      # :91
    occ_ray = term_prob[:, :-1].sum(-1, dtype=dt)                                    # :93
    d_u = (d_aug * term_prob).sum(-1, dtype=dt)                                      # :96

    o_k = occ[gx, gy]                                                                # :101
    dm_do = acc[:, -1] / (f(1.) - o_k)                                               # :102
    acc_after = acc.copy()
    acc_after[np.arange(n_depths)[None, :] < gy[:, None]] = 0.                       # :103-105
    de_do = acc_after.sum(-1, dtype=dt) * delta_d / (f(1.) - o_k)                    # :107

    nz = de_do > f(min_grad_thre)                                                    # :111
    de_do, dm_do, d_u, occ_ray, gx, gy, o_k = de_do[nz], dm_do[nz], d_u[nz], occ_ray[nz], gx[nz], gy[nz], o_k[nz]

    if log_occ_on:
        do_ds = -o_k * (f(1) - o_k) / sigma                                          # :121
    else:
        do_ds = f(-1.) / f(2 * occupancy_th)                                         # :123
    de_ds = (de_do * do_ds)
    dm_ds = (dm_do * do_ds)

    depth_obs = depth_obs.copy()
    if occlusion_on:                                                                 # :132-149
        dobs = depth_obs[gx]
        occluded = (gx >= n_fg) & (dobs < d_u - f(occlusion_th)) & (dobs > 0.)
        keep = ~occluded
        gx, gy = gx[keep], gy[keep]
        depth_obs[n_fg:] = d_term
        dobs = depth_obs[gx]
        d_u, de_ds, dm_ds, occ_ray = d_u[keep], de_ds[keep], dm_ds[keep], occ_ray[keep]
    else:
        depth_obs[n_fg:] = d_term
        dobs = depth_obs[gx]
    res_d = dobs - d_u                                                               # :155

    group_id, groups_idx, groups_count = np.unique(gx, return_inverse=True, return_counts=True)  # :164
    n_valid = group_id.shape[0]
    n_valid_fg = int((group_id < n_fg).sum())

    def _scatter(vals):
        out = np.zeros((n_valid,) + vals.shape[1:], dt)
        np.add.at(out, groups_idx, vals)
        return out

    occ_ray_r = _scatter(occ_ray) / groups_count.astype(dt)                          # :168
    res_d_ray = _scatter(res_d) / groups_count.astype(dt)                            # :169
    mask_ray = np.concatenate([np.ones(n_valid_fg, dt), np.zeros(n_valid - n_valid_fg, dt)])
    res_m_ray = occ_ray_r - mask_ray                                                 # :175

    pts_g = pts_obj[gx, gy]                                                          # :185
    _, ds_di = get_batch_sdf_jacobian(dec, latent, pts_g)                            # :186
    dm_di = dm_ds.reshape(-1, 1, 1) * ds_di
    de_di = de_ds.reshape(-1, 1, 1) * ds_di
    dxo = points_to_pose_jacobian(pts_g, scale_on)
    jac_d_tow = np.matmul(de_di[..., -3:], dxo)                                      # :202
    jac_m_tow = np.matmul(dm_di[..., -3:], dxo)                                      # :206
    jac_d_code, jac_m_code = de_di[..., :-3], dm_di[..., :-3]

    out = (res_d_ray.reshape(-1, 1, 1), _scatter(jac_d_tow), _scatter(jac_d_code),
           res_m_ray.reshape(-1, 1, 1), _scatter(jac_m_tow), _scatter(jac_m_code))
    if return_debug:
        dbg = dict(n_valid_samples=int(query.shape[0]), n_band=int(with_grad.sum()),
                   n_grad_rows=int(gx.shape[0]), ray_ids=group_id)
        return out, dbg
    return out


# --------------------------------------------------------------------------------------
# Optimiser (wild_completion/optimizer.py)
# --------------------------------------------------------------------------------------
@dataclass
class OptTrace:
    """Per-iteration quantities captured for step-level parity tests."""
    H: List[np.ndarray] = field(default_factory=list)
    b: List[np.ndarray] = field(default_factory=list)
    dx: List[np.ndarray] = field(default_factory=list)
    dx_raw: List[np.ndarray] = field(default_factory=list)     # before the pose_known zeroing (:237-238)
    latent: List[np.ndarray] = field(default_factory=list)
    T_ow: List[np.ndarray] = field(default_factory=list)
    rows_fwd: int = 0
    rows_grad: int = 0


def _normal_eq(J: np.ndarray, res: np.ndarray, w: np.ndarray, weight: float, dt):
    """w_t * sum_i rho_i J_i^T J_i / n  and  -w_t * sum_i rho_i J_i^T r_i / n
    (optimizer.py:152-159, 189-190): bmm(J^T, J) summed over the batch."""
    f = np.dtype(dt).type
    n = J.shape[0]
    Jm = J.reshape(n, -1)
    wv = w.reshape(n, 1)
    H = f(weight) * ((Jm * wv).T @ Jm) / f(n)
    b = -f(weight) * ((Jm * wv).T @ res.reshape(n)) / f(n)
    return H.astype(dt), b.astype(dt)


def shape_pose_joint_opt(dec: DecoderOracle, cfg: dict, latent: np.ndarray, T_ow: np.ndarray,
                         render_data: dict, points_w: np.ndarray, cube_radius: float,
                         pose_known: bool = False, trace: Optional[OptTrace] = None, iter_offset: int = 0):
    """wild_completion/optimizer.py:28-302 (vis is None).  `latent` is updated in place like the
    reference (:248) and also returned; returns (latent, T_ow, iter_count).

    `iter_offset` (test hook, 0 in the reference) makes the loop index start at that value so a single
    iteration of a longer run can be replayed from a stored state: it only shifts the `i` seen by the
    robust_iter switch (:145,:183) and the `i > 1` stop guards (:276-285)."""
    dt = dec.dtype
    f = np.dtype(dt).type
    opt = cfg['opt']
    iter_count_max = opt['converge']['max_iter']
    eps_g, eps_c = float(opt['converge']['epsilon_g']), float(opt['converge']['epsilon_c'])
    eps_t, eps_r, eps_s = (float(opt['converge'][k]) for k in ('epsilon_t', 'epsilon_r', 'epsilon_s'))
    max_render_frame = opt['render']['n_frame']
    n_depth = opt['render']['n_sample_on_ray']
    occ_cutoff = float(opt['render']['occ_cutoff_m'])
    log_occ = opt['render']['log_sdf_occ']
    occlusion_on = opt['render']['occlusion_on']
    w_recon, w_depth, w_mask, w_codereg = (float(opt['weight'][k]) for k in ('w_recon', 'w_depth', 'w_mask', 'w_codereg'))
    lm_on, lm_eye, lm_lambda_0 = opt['lm']['lm_on'], opt['lm']['lm_eye'], float(opt['lm']['lm_lambda_0'])
    t_recon, t_depth = float(opt['recon']['robust_th_m']), float(opt['render']['robust_th_m'])
    robust_iter = opt['robust_iter']
    s_damp = float(opt['lm']['s_damp'])
    scale_on = opt['scale_on']
    pose_dim = 7 if scale_on else 6
    code_len = latent.shape[0]
    est = pose_dim + code_len

    T_ow = np.asarray(T_ow, dtype=dt).copy()
    points_w = np.asarray(points_w, dtype=dt)
    cur_scale = f(np.linalg.det(T_ow[:3, :3])) ** f(-1 / 3)                          # :66

    n_frames = len(render_data["T_wc"])
    frame_ind = np.linspace(0, n_frames - 1, min(max_render_frame, n_frames)).astype(np.int32)  # :78

    iter_count = 0
    for i in range(iter_offset, iter_offset + iter_count_max):
        res_d_all, J_d_all, res_m_all, J_m_all = [], [], [], []
        for idx in frame_ind:                                                        # :102
            T_wc = np.asarray(render_data["T_wc"][idx], dtype=dt)
            T_oc = (T_ow @ T_wc).astype(dt)
            T_co = np.linalg.inv(T_oc).astype(dt)
            depth_range = f(cube_radius) * cur_scale                                 # :107
            d_min = T_co[2, 3] - f(1.0) * depth_range
            d_max = T_co[2, 3] + f(0.8) * depth_range
            depths = torch_linspace(d_min, d_max, n_depth, dt)                       # :111
            rays = np.concatenate([render_data["rays_fg"][idx], render_data["rays_bg"][idx]], 0)
            r = compute_render_loss(dec, latent, rays, render_data["depth_fg"][idx],
                                    render_data["depth_bg"][idx], T_oc, depths, scale_on, log_occ,
                                    occ_cutoff, depth_range, occlusion_on, return_debug=True)
            if r is None:                                                            # :130-132
                continue
            (rd, jdp, jdc, rm, jmp, jmc), dbg = r
            if trace is not None:
                trace.rows_fwd += dbg['n_valid_samples']
                trace.rows_grad += dbg['n_grad_rows']
            res_d_all.append(rd); J_d_all.append(np.concatenate([jdp, jdc], -1))
            res_m_all.append(rm); J_m_all.append(np.concatenate([jmp, jmc], -1))
        n_d = sum(r.shape[0] for r in res_d_all)
        if n_d == 0:                                                                 # :139-141
            break
        res_d, J_d = np.concatenate(res_d_all, 0), np.concatenate(J_d_all, 0)
        res_m, J_m = np.concatenate(res_m_all, 0), np.concatenate(J_m_all, 0)

        if i >= robust_iter:                                                         # :145-149
            _, rw = get_robust_res(res_d, t_depth)
        else:
            rw = np.ones_like(res_d)
        H_d, b_d = _normal_eq(J_d, res_d, rw, w_depth, dt)                           # :152-153
        H_m, b_m = _normal_eq(J_m, res_m, np.ones_like(res_m), w_mask, dt)           # :158-159

        pts_o = (points_w[..., None, :] * T_ow[:3, :3]).sum(-1) + T_ow[:3, 3]        # :168
        res_r, jrp, jrc = compute_sdf_loss(dec, latent, pts_o, scale_on)             # :170
        if trace is not None:
            trace.rows_grad += pts_o.shape[0]
        J_r = np.concatenate([jrp, jrc], -1)
        if i >= robust_iter:                                                         # :183-187
            _, rw = get_robust_res(res_r, t_recon)
        else:
            rw = np.ones_like(res_r)
        H_r, b_r = _normal_eq(J_r, res_r, rw, w_recon, dt)                           # :189-190

        H = np.zeros((est, est), dt)                                                 # :200-231
        H += H_d; H += H_m; H += H_r
        H[pose_dim:, pose_dim:] += f(w_codereg) * np.eye(code_len, dtype=dt)
        if scale_on:
            H[pose_dim - 1, pose_dim - 1] += f(s_damp)
        if lm_on:
            if lm_eye:
                H += f(lm_lambda_0) * np.max(np.diag(H)) * np.eye(est, dtype=dt)
            else:
                H += f(lm_lambda_0) * np.diag(np.diag(H))
        b = np.zeros(est, dt)
        b += b_d; b += b_m; b += b_r
        b[pose_dim:] += -f(w_codereg) * latent

        dx = (np.linalg.inv(H).astype(dt) @ b).astype(dt)                            # :234
        if trace is not None:
            trace.dx_raw.append(dx.copy())
        if pose_known:
            dx[:6] = 0                                                               # :237-238
        delta_c = dx[pose_dim:]
        delta_T = exp_sim3(dx[:pose_dim]) if scale_on else exp_se3(dx[:pose_dim])    # :242-245
        T_ow = (delta_T @ T_ow).astype(dt)                                           # :247
        latent += delta_c                                                            # :248 (in place)

        cur_scale = f(np.linalg.det(T_ow[:3, :3])) ** f(-1 / 3)                      # :250
        with np.errstate(invalid="ignore"):
            delta_scale = f(np.linalg.det(delta_T[:3, :3])) ** f(1 / 3)              # :251
            delta_tran = f(np.sqrt((delta_T[:3, 3] ** 2).sum())) * cur_scale         # :252
            delta_rot = abs(rotation_matrix_to_axis_angle(delta_T[:3, :3] * cur_scale)) * f(180.0 / math.pi)  # :253
        if trace is not None:
            trace.H.append(H.copy()); trace.b.append(b.copy()); trace.dx.append(dx.copy())
            trace.latent.append(latent.copy()); trace.T_ow.append(T_ow.copy())
        iter_count = i + 1 - iter_offset
        with np.errstate(invalid="ignore", divide="ignore"):
            if np.max(np.abs(b)) < eps_g and i > 1:                                  # :276
                break
            if np.max(np.abs(delta_c / (latent + f(1e-12)))) < eps_c and i > 1:      # :280
                break
            if (not pose_known) and (delta_tran < eps_t and delta_rot < eps_r and delta_scale < eps_s and i > 1):  # :285
                break
    return latent, T_ow, iter_count


def shape_opt_deepsdf(dec: DecoderOracle, cfg: dict, latent: np.ndarray, T_ow: np.ndarray,
                      points_w: np.ndarray, trace: Optional[OptTrace] = None, iter_offset: int = 0):
    """wild_completion/optimizer.py:306-429: latent-only LM ("DeepSDF baseline"), pose_dim = 0."""
    dt = dec.dtype
    f = np.dtype(dt).type
    opt = cfg['opt']
    iter_count_max = opt['converge']['max_iter']
    eps_g, eps_c = float(opt['converge']['epsilon_g']), float(opt['converge']['epsilon_c'])
    w_recon, w_codereg = float(opt['weight']['w_recon']), float(opt['weight']['w_codereg'])
    lm_on, lm_eye, lm_lambda_0 = opt['lm']['lm_on'], opt['lm']['lm_eye'], float(opt['lm']['lm_lambda_0'])
    t_recon = float(opt['recon']['robust_th_m'])
    robust_iter = opt['robust_iter']
    scale_on = opt['scale_on']
    code_len = latent.shape[0]
    T_ow = np.asarray(T_ow, dtype=dt)
    points_w = np.asarray(points_w, dtype=dt)
    iter_count = 0
    for i in range(iter_offset, iter_offset + iter_count_max):
        pts_o = (points_w[..., None, :] * T_ow[:3, :3]).sum(-1) + T_ow[:3, 3]        # :343
        res_r, _, J = compute_sdf_loss(dec, latent, pts_o, scale_on)                 # :345
        if trace is not None:
            trace.rows_grad += pts_o.shape[0]
        if i >= robust_iter:
            _, rw = get_robust_res(res_r, t_recon)
        else:
            rw = np.ones_like(res_r)
        H_r, b_r = _normal_eq(J, res_r, rw, w_recon, dt)                             # :362-363
        H = np.zeros((code_len, code_len), dt)
        H += H_r
        H += f(w_codereg) * np.eye(code_len, dtype=dt)
        if lm_on:                                                                    # :385-390
            if lm_eye:
                H += f(lm_lambda_0) * np.max(np.diag(H)) * np.eye(code_len, dtype=dt)
            else:
                H += f(lm_lambda_0) * np.diag(np.diag(H))
        b = np.zeros(code_len, dt)
        b += b_r
        b += -f(w_codereg) * latent
        dx = (np.linalg.inv(H).astype(dt) @ b).astype(dt)                            # :397
        latent += dx                                                                 # :401
        if trace is not None:
            trace.H.append(H.copy()); trace.b.append(b.copy()); trace.dx.append(dx.copy())
            trace.latent.append(latent.copy())
        iter_count = i + 1 - iter_offset
        with np.errstate(invalid="ignore", divide="ignore"):
            if np.max(np.abs(b)) < eps_g and i > 1:                                  # :417
                break
            if np.max(np.abs(dx / (latent + f(1e-12)))) < eps_c and i > 1:           # :421
                break
    return latent, T_ow, iter_count


# --------------------------------------------------------------------------------------
# Mesher grid (wild_completion/mesher.py:6-24, utils.py:542-562)
# --------------------------------------------------------------------------------------
def create_voxel_grid(vol_dim: int) -> np.ndarray:
    """utils.py:542-562, fp32.  `LongTensor / int` is TRUE division in torch >= 1.5, so the y and x
    columns are fractional ("sheared" grid); mirrored exactly, in the
    reference's fp32 op order (int64 -> float32 divide, fmod, then the assignment casts)."""
    n = vol_dim
    idx = np.arange(n ** 3, dtype=np.int64)
    voxel_size = 2.0 / (n - 1)
    vals = np.zeros((n ** 3, 3), np.float32)
    vals[:, 2] = (idx % n).astype(np.float32)
    q1 = idx.astype(np.float32) / np.float32(n)                      # long / int -> float32 true division
    vals[:, 1] = np.fmod(q1, np.float32(n))
    q2 = q1 / np.float32(n)
    vals[:, 0] = np.fmod(q2, np.float32(n))
    vals[:, 0] = vals[:, 0] * np.float32(voxel_size) + np.float32(-1)
    vals[:, 1] = vals[:, 1] * np.float32(voxel_size) + np.float32(-1)
    vals[:, 2] = vals[:, 2] * np.float32(voxel_size) + np.float32(-1)
    return vals


def decode_grid(dec: DecoderOracle, code: np.ndarray, vol_dim: int, cube_radius: float) -> np.ndarray:
    """mesher.py:12,18: SDF on create_voxel_grid(N) * cube_radius, reshaped (N,N,N) C-order."""
    pts = (create_voxel_grid(vol_dim) * np.float32(cube_radius)).astype(dec.dtype)
    return decode_sdf(dec, code, pts).reshape(vol_dim, vol_dim, vol_dim)


# --------------------------------------------------------------------------------------
# Emulation of the device decoder's split-fp16 arithmetic (used to pin the tolerance the CUDA
# tensor-core path is tested with; see DESIGN.md "numerics")
# --------------------------------------------------------------------------------------
def split_f16(x: np.ndarray, scale: float):
    xs = np.asarray(x, np.float32) * np.float32(scale)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)
