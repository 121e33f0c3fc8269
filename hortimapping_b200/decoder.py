"""Host-side mirror of the reference's DeepSDF decoder interface on top of the C ABI.

Mirrors deepsdf/networks/deep_sdf_decoder.py (`Decoder`, a callable nn.Module taking (N,35) or
(N,1,35) rows) and deepsdf/deep_sdf/workspace.py (`config_decoder`, `load_latent_vectors`), plus the
two helpers of wild_completion/utils.py the hot path goes through (`decode_sdf`,
`get_batch_sdf_jacobian`).  PyTorch is only the owner of device memory and streams here: every value
is computed by libhortimapping_b200.so.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import HM_IN, HM_LATENT, HM_LAYERS, check

_IN_DIM = [35, 512, 512, 512, 512, 512, 512, 512, 512]
_OUT_DIM = [512, 512, 512, 477, 512, 512, 512, 512, 1]


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class _SdfRows(torch.autograd.Function):
    """Decoder.forward on rows [n,35] with the input gradient the reference gets from autograd
    (wild_completion/utils.py:112-122): forward runs the fused forward+Jacobian kernel when a gradient
    may be requested, backward is grad_out * jac."""

    @staticmethod
    def forward(ctx, dec: "Decoder", rows: torch.Tensor):
        need = bool(ctx.needs_input_grad[1])
        sdf, jac = dec._eval_rows(rows, with_jac=need)
        if need:
            ctx.save_for_backward(jac)
        return sdf

    @staticmethod
    def backward(ctx, grad_out):
        (jac,) = ctx.saved_tensors
        return None, (grad_out.reshape(-1, 1) * jac.reshape(-1, HM_IN)).reshape(jac.shape)


class Decoder:
    """Drop-in for the reference `Decoder` instance returned by config_decoder: `decoder(x)` with
    x of shape (N,35) -> (N,1) or (N,1,35) -> (N,1,1), differentiable w.r.t. x.  `.cuda()`, `.eval()`,
    `.to()` are accepted and are no-ops (the context is bound to its device at construction)."""

    def __init__(self, weights: Sequence[np.ndarray], biases: Sequence[np.ndarray], device: Optional[int] = None,
                 latent_in: Sequence[int] = (4,), specs: Optional[dict] = None, _library=None):
        if not torch.cuda.is_available():
            raise RuntimeError("hortimapping_b200.Decoder needs a CUDA device (no CPU fallback)")
        if len(weights) != HM_LAYERS or tuple(latent_in) != (4,):
            raise ValueError("unsupported decoder architecture: need 9 linear layers and latent_in=[4]")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.specs = specs or {}
        self.training = False
        L = _library if _library is not None else _lib.lib()      # (_library: the test-only superset build, _testing.py)
        desc = _lib.DecoderDesc()
        desc.n_layers, desc.latent_size, desc.latent_in_layer = HM_LAYERS, HM_LATENT, 4
        self._keep = []
        for l in range(HM_LAYERS):
            w = np.ascontiguousarray(weights[l], dtype=np.float32)
            b = np.ascontiguousarray(biases[l], dtype=np.float32).reshape(-1)
            if w.shape != (_OUT_DIM[l], _IN_DIM[l]) or b.shape != (_OUT_DIM[l],):
                raise ValueError(f"layer {l}: weight {w.shape} / bias {b.shape}, expected {(_OUT_DIM[l], _IN_DIM[l])}")
            self._keep += [w, b]
            desc.in_dim[l], desc.out_dim[l] = _IN_DIM[l], _OUT_DIM[l]
            desc.weight[l] = w.ctypes.data_as(_lib.c_float_p)
            desc.bias[l] = b.ctypes.data_as(_lib.c_float_p)
        h = C.c_void_p()
        check(L.hm_create(C.byref(h), self.device_index, C.byref(desc)), "hm_create")
        self._h = h
        self._L = L

    # --- nn.Module-like surface the reference's host code touches (workspace.py:221-223)
    def cuda(self, *a, **k):
        return self

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def parameters(self):
        return iter(())

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._L.hm_destroy(h)
            except Exception:
                pass

    @property
    def handle(self):
        return self._h

    def set_engine(self, engine: str):
        code = {"tc": _lib.HM_ENGINE_TC, "simt": _lib.HM_ENGINE_SIMT}[engine]
        check(self._L.hm_set_engine(self._h, code), "hm_set_engine")

    def set_sparse_plan(self, on: bool):
        """Tensor-core engine: use the calibrated sparse plan (MMAs on all-zero activation chunks are dropped, checked per tile,
        violators re-evaluated with the full plan); results are bit-identical either way."""
        check(self._L.hm_set_sparse_plan(self._h, int(bool(on))), "hm_set_sparse_plan")

    def set_mask_reuse(self, on: bool):
        """Joint loop: take the gradient of the in-band ray samples from the forward pass's stored ReLU bits instead of evaluating
        those rows a second time (hm_set_mask_reuse); results are bit-identical either way."""
        check(self._L.hm_set_mask_reuse(self._h, int(bool(on))), "hm_set_mask_reuse")

    def plan_info(self) -> dict:
        """Tensor-core FLOP the engine issues per row under the current calibration (hm_plan_info): sparse / full plan, forward /
        forward + gradient, and the number of 64-wide chunks of each hidden layer the sparse plan keeps."""
        out = (C.c_double * 12)()
        check(self._L.hm_plan_info(self._h, out), "hm_plan_info")
        return {"issued_flop_per_row": {"sparse_forward": out[0], "sparse_jacobian": out[1], "full_forward": out[2], "full_jacobian": out[3]},
                "alive_chunks_per_layer": [int(out[4 + l]) for l in range(8)]}

    def calibrate(self, rows: torch.Tensor):
        rows = _f32c(rows.reshape(-1, HM_IN), self.device)
        check(self._L.hm_calibrate(self._h, rows.data_ptr(), rows.shape[0], _stream_ptr(self.device)), "hm_calibrate")

    def profile(self, on: bool):
        check(self._L.hm_profile_enable(self._h, int(bool(on))), "hm_profile_enable")

    def saturation_count(self) -> int:
        """Thread blocks of the tensor-core engine, since the last call, in which an operand left the calibrated fp16 range
        (hm_saturation_count; synchronises).  Non-zero means the results were not fp32-grade: calibrate() on representative rows."""
        n = C.c_int64(0)
        check(self._L.hm_saturation_count(self._h, C.byref(n)), "hm_saturation_count")
        return int(n.value)

    def counters(self) -> dict:
        c = _lib.Counters()
        check(self._L.hm_get_counters(self._h, C.byref(c)), "hm_get_counters")
        return {k: getattr(c, k) for k, _ in c._fields_}

    # --- evaluation
    def _eval_rows(self, rows: torch.Tensor, with_jac: bool):
        r = _f32c(rows.reshape(-1, HM_IN), self.device)
        n = r.shape[0]
        sdf = torch.empty(n, device=self.device, dtype=torch.float32)
        st = _stream_ptr(self.device)
        if with_jac:
            jac = torch.empty(n, HM_IN, device=self.device, dtype=torch.float32)
            check(self._L.hm_sdf_jacobian_rows(self._h, r.data_ptr(), n, sdf.data_ptr(), jac.data_ptr(), st), "hm_sdf_jacobian_rows")
        else:
            jac = None
            check(self._L.hm_sdf_forward_rows(self._h, r.data_ptr(), n, sdf.data_ptr(), st), "hm_sdf_forward_rows")
        return sdf.reshape(rows.shape[:-1] + (1,)), (jac.reshape(rows.shape) if jac is not None else None)

    def __call__(self, inp: torch.Tensor) -> torch.Tensor:
        if inp.shape[-1] != HM_IN:
            raise ValueError(f"decoder input must have {HM_IN} columns, got {tuple(inp.shape)}")
        return _SdfRows.apply(self, inp)

    forward = __call__

    def sdf(self, latent: torch.Tensor, xyz: torch.Tensor) -> torch.Tensor:
        """decode_sdf (wild_completion/utils.py:144-172): latent (32,), xyz (N,3) -> (N,)."""
        lat = _f32c(latent.reshape(HM_LATENT), self.device)
        x = _f32c(xyz[:, 0:3], self.device)
        out = torch.empty(x.shape[0], device=self.device, dtype=torch.float32)
        check(self._L.hm_sdf_forward(self._h, lat.data_ptr(), x.data_ptr(), x.shape[0], out.data_ptr(), _stream_ptr(self.device)), "hm_sdf_forward")
        return out

    def sdf_jacobian(self, latent: torch.Tensor, xyz: torch.Tensor):
        """get_batch_sdf_jacobian (wild_completion/utils.py:175-193): -> y (n,1,1), g (n,1,35)."""
        lat = _f32c(latent.reshape(HM_LATENT), self.device)
        x = _f32c(xyz, self.device)
        n = x.shape[0]
        y = torch.empty(n, device=self.device, dtype=torch.float32)
        g = torch.empty(n, HM_IN, device=self.device, dtype=torch.float32)
        check(self._L.hm_sdf_jacobian(self._h, lat.data_ptr(), x.data_ptr(), n, y.data_ptr(), g.data_ptr(), _stream_ptr(self.device)), "hm_sdf_jacobian")
        return y.reshape(n, 1, 1), g.reshape(n, 1, HM_IN)

    def sdf_grid(self, latent: torch.Tensor, vol_dim: int, cube_radius: float) -> torch.Tensor:
        """SDF on create_voxel_grid(vol_dim) * cube_radius (wild_completion/mesher.py:12-18) -> (N,N,N)."""
        lat = _f32c(latent.reshape(HM_LATENT), self.device)
        out = torch.empty(vol_dim ** 3, device=self.device, dtype=torch.float32)
        check(self._L.hm_sdf_grid(self._h, lat.data_ptr(), vol_dim, float(cube_radius), out.data_ptr(), _stream_ptr(self.device)), "hm_sdf_grid")
        return out.view(vol_dim, vol_dim, vol_dim)

    def isosurface(self, sdf_3d: torch.Tensor, level: float = 0.0, spacing: float = 1.0, affine_radius: Optional[float] = None):
        """Zero level set of an (N,N,N) SDF grid on the device (hm_isosurface: marching tetrahedra, the device twin of
        marching.marching_tetrahedra) -> vertices (V,3) float32, faces (F,3) int32 as CUDA tensors.  With `affine_radius`
        the vertices are mapped to (v - 1) * radius as wild_completion/utils.py:583-585 does."""
        vol = _f32c(sdf_3d, self.device)
        n = int(vol.shape[0])
        assert tuple(vol.shape) == (n, n, n), "isosurface expects a cubic (N,N,N) grid"
        nv, nf = C.c_int64(0), C.c_int64(0)
        st = _stream_ptr(self.device)
        check(self._L.hm_isosurface(self._h, vol.data_ptr(), n, float(level), float(spacing), C.byref(nv), C.byref(nf), st), "hm_isosurface")
        verts = torch.empty(nv.value, 3, device=self.device, dtype=torch.float32)
        faces = torch.empty(nf.value, 3, device=self.device, dtype=torch.int32)
        check(self._L.hm_isosurface_fetch(self._h, verts.data_ptr() if nv.value else None, faces.data_ptr() if nf.value else None,
                                          0 if affine_radius is None else 1, float(affine_radius or 0.0), st), "hm_isosurface_fetch")
        return verts, faces

    def voxel_grid(self, vol_dim: int, cube_radius: float = 1.0) -> torch.Tensor:
        out = torch.empty(vol_dim ** 3, 3, device=self.device, dtype=torch.float32)
        check(self._L.hm_voxel_grid(self._h, vol_dim, float(cube_radius), out.data_ptr(), _stream_ptr(self.device)), "hm_voxel_grid")
        return out


# ------------------------------------------------------------------------------------------------
# wild_completion/utils.py helpers with the reference's signatures
# ------------------------------------------------------------------------------------------------
def decode_sdf(decoder: Decoder, lat_vec: torch.Tensor, x: torch.Tensor, max_batch: int = 64 ** 3) -> torch.Tensor:
    return decoder.sdf(lat_vec, x)


def get_batch_sdf_jacobian(decoder: Decoder, lat_vec: torch.Tensor, x: torch.Tensor):
    return decoder.sdf_jacobian(lat_vec, x)


def create_voxel_grid(decoder: Decoder, vol_dim: int = 128) -> torch.Tensor:
    return decoder.voxel_grid(vol_dim, 1.0)


# ------------------------------------------------------------------------------------------------
# deepsdf/deep_sdf/workspace.py
# ------------------------------------------------------------------------------------------------
def fold_weight_norm(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """weight_norm(dim=0): W = v * (g / ||v||_row) (deep_sdf_decoder.py:49-54), evaluated by the very op torch's weight_norm hook
    calls (`torch._weight_norm`), so the folded matrix is bit-identical to the `lin.weight` the reference module computes."""
    v = v.float().contiguous()
    return torch._weight_norm(v, g.float().reshape(-1, *([1] * (v.dim() - 1))).contiguous(), 0)


def folded_weights_from_state_dict(state: dict):
    """(W [9], b [9]) as float32 numpy arrays from a DeepSDF `model_state_dict` (keys optionally prefixed `module.`): lin0..lin7
    carry weight-norm pairs (weight_g, weight_v) that are folded here, lin8 a plain weight (deep_sdf_decoder.py:49-56)."""
    sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in state.items()}
    W, b = [], []
    for l in range(HM_LAYERS):
        if f"lin{l}.weight_v" in sd:
            w = fold_weight_norm(sd[f"lin{l}.weight_v"].cpu(), sd[f"lin{l}.weight_g"].cpu())
        else:
            w = sd[f"lin{l}.weight"].cpu().float()
        W.append(w.numpy())
        b.append(sd[f"lin{l}.bias"].cpu().float().numpy())
    return W, b


def decoder_from_state_dict(state: dict, specs: Optional[dict] = None, device: Optional[int] = None) -> Decoder:
    W, b = folded_weights_from_state_dict(state)
    return Decoder(W, b, device=device, specs=specs)


def load_decoder_weights(experiment_directory: str, checkpoint: str = "latest"):
    """The host half of config_decoder (workspace.py:203-218), no GPU needed: specs.json + ModelParameters/<checkpoint>.pth ->
    (W [9], b [9], specs).  Raises like the reference when specs.json is missing, and for any architecture other than the
    shipped one."""
    specs_filename = os.path.join(experiment_directory, "specs.json")
    if not os.path.isfile(specs_filename):
        raise Exception('The experiment directory does not include specifications file "specs.json"')
    specs = json.load(open(specs_filename))
    ns = specs["NetworkSpecs"]
    if (specs["CodeLength"] != HM_LATENT or list(ns["dims"]) != [512] * 8 or list(ns["latent_in"]) != [4]
            or ns.get("xyz_in_all") or ns.get("use_tanh") or not ns.get("weight_norm")):
        raise ValueError("hortimapping_b200 supports the shipped DeepSDF architecture only (8x512, latent 32, latent_in=[4])")
    saved = torch.load(os.path.join(experiment_directory, "ModelParameters", checkpoint + ".pth"), map_location="cpu")
    W, b = folded_weights_from_state_dict(saved["model_state_dict"])
    return W, b, specs


def calibration_rows(codes, xyz_half_range: float, n: int = 262144, seed: int = 0) -> torch.Tensor:
    """Rows [n][35] on which `Decoder.calibrate` measures the tensor-core engine's operand ranges and the set of hidden units that
    are ever alive (the sparse plan).  They cover the region the optimisers work in: latents on the segments between
    the MEAN training code -- the initial latent of every fruit, run_shape_completion_challenge.py:51-52 -- and the training codes,
    with a little jitter (the LM steps leave those segments), and query points in a cube of +- xyz_half_range.  A hidden unit
    that fires outside this region is still handled exactly (its tile is re-evaluated with the full plan), only slower; measured
    on the bench's latent-only loop: 0.4 % of the tiles re-evaluated when calibrating on the raw training codes alone, none
    with these rows."""
    codes = torch.as_tensor(np.asarray(codes.detach().cpu() if isinstance(codes, torch.Tensor) else codes), dtype=torch.float32).reshape(-1, HM_LATENT)
    g = torch.Generator(device="cpu").manual_seed(seed)
    mean = codes.mean(0, keepdim=True)
    z = codes[torch.randint(0, codes.shape[0], (n,), generator=g)]
    t = torch.rand(n, 1, generator=g)
    t[: n // 4] = 1.0                                      # a quarter of the rows are the training codes themselves
    z = mean + t * (z - mean) + 0.03 * torch.randn(n, HM_LATENT, generator=g)
    x = (torch.rand(n, 3, generator=g) * 2 - 1) * float(xyz_half_range)
    return torch.cat([z, x], 1)


def config_decoder(experiment_directory: str, checkpoint: str = "latest", device: Optional[int] = None) -> Decoder:
    """deepsdf/deep_sdf/workspace.py:203-225: read specs.json + ModelParameters/<checkpoint>.pth, build the
    decoder on the GPU in eval mode.  The tensor-core engine's fp16 operand scales are calibrated on the model's own
    training codes when LatentCodes/<checkpoint>.pth is there (else the library's synthetic default stays; `calibration`
    says which)."""
    W, b, specs = load_decoder_weights(experiment_directory, checkpoint)
    dec = Decoder(W, b, device=device, specs=specs)
    dec.calibration = "default (synthetic latents ~ N(0, 0.1))"
    lat_file = os.path.join(experiment_directory, "LatentCodes", checkpoint + ".pth")
    if os.path.isfile(lat_file):
        try:
            codes = load_latent_vectors(experiment_directory, checkpoint)
            if isinstance(codes, (list, tuple)):          # legacy tensor-format checkpoints come back as a list (workspace.py:99-107)
                codes = torch.stack([c.reshape(-1) for c in codes])
            codes = codes.to(dec.device)
            clamp = float(specs.get("ClampingDistance", 0.1))
            dec.calibrate(calibration_rows(codes, 1.5 * clamp))
            dec.calibration = f"{codes.shape[0]} training codes of {lat_file}"
        except Exception as e:                            # a broken codes file must not take the decoder down
            import warnings
            warnings.warn(f"hortimapping_b200: calibration on {lat_file} failed ({e}); keeping the default operand scales")
    return dec


def load_latent_vectors(experiment_directory: str, checkpoint: str = "latest") -> torch.Tensor:
    """deepsdf/deep_sdf/workspace.py:82-114."""
    filename = os.path.join(experiment_directory, "LatentCodes", checkpoint + ".pth")
    if not os.path.isfile(filename):
        raise Exception(
            "The experiment directory ({}) does not include a latent code file".format(experiment_directory)
            + " for checkpoint '{}'".format(checkpoint))
    data = torch.load(filename, map_location="cpu")
    lc = data["latent_codes"]
    if isinstance(lc, torch.Tensor):
        return [lc[i].cuda() for i in range(lc.size()[0])]
    return lc["weight"].detach().float()
