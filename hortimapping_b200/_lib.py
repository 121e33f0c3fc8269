"""ctypes binding of libhortimapping_b200.so (the C ABI declared in include/hortimapping_b200.h).

The product path has NO fallback: if the shared library is missing or cannot be loaded this module
raises, it never routes to a CPU or PyTorch implementation.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhortimapping_b200.so")

HM_LATENT, HM_IN, HM_HIDDEN, HM_LAYERS = 32, 35, 512, 9
HM_ENGINE_TC, HM_ENGINE_SIMT = 0, 1
STATUS = dict(CONV_GRADIENT=0x01, CONV_CODE=0x02, CONV_POSE=0x04, MAX_ITER=0x08, FRAME_SKIPPED=0x10,
              SUBMAP_INVALID=0x20, F16_SATURATED=0x40, SOLVE_FAILED=0x80)

c_float_p = C.POINTER(C.c_float)


class DecoderDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("latent_size", C.c_int32), ("latent_in_layer", C.c_int32),
                ("in_dim", C.c_int32 * HM_LAYERS), ("out_dim", C.c_int32 * HM_LAYERS),
                ("weight", c_float_p * HM_LAYERS), ("bias", c_float_p * HM_LAYERS)]


class OptParams(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("n_depth_samples", C.c_int32), ("log_sdf_occ", C.c_int32),
                ("occlusion_on", C.c_int32), ("lm_on", C.c_int32), ("lm_eye", C.c_int32), ("robust_iter", C.c_int32),
                ("scale_on", C.c_int32), ("min_valid_sample", C.c_int32), ("iter_offset", C.c_int32),
                ("epsilon_g", C.c_double), ("epsilon_c", C.c_double), ("epsilon_t", C.c_double),
                ("epsilon_r", C.c_double), ("epsilon_s", C.c_double), ("occ_cutoff_m", C.c_double),
                ("w_recon", C.c_double), ("w_depth", C.c_double), ("w_mask", C.c_double), ("w_codereg", C.c_double),
                ("lm_lambda_0", C.c_double), ("robust_th_recon", C.c_double), ("robust_th_depth", C.c_double),
                ("s_damp", C.c_double), ("occlusion_th", C.c_double), ("min_grad_thre", C.c_double)]


class FruitBatch(C.Structure):
    _fields_ = [("n_fruits", C.c_int32), ("d_latents", C.c_void_p), ("d_T_ow", C.c_void_p), ("d_points_w", C.c_void_p),
                ("h_point_offsets", C.c_void_p), ("h_frame_offsets", C.c_void_p), ("d_T_wc", C.c_void_p),
                ("h_ray_offsets", C.c_void_p), ("h_n_fg", C.c_void_p), ("d_rays", C.c_void_p),
                ("d_depth_obs", C.c_void_p), ("h_cube_radius", C.c_void_p), ("h_pose_known", C.c_void_p),
                ("d_iter_count", C.c_void_p), ("d_status", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [("rows_forward", C.c_int64), ("rows_jacobian", C.c_int64), ("kernel_launches", C.c_int64),
                ("iterations", C.c_int64), ("decoder_launches", C.c_int64), ("decoder_ms", C.c_double),
                ("forward_launches", C.c_int64), ("forward_ms", C.c_double), ("jacobian_launches", C.c_int64),
                ("jacobian_ms", C.c_double), ("tiles_forward", C.c_int64), ("tiles_jacobian", C.c_int64),
                ("tiles_redone_forward", C.c_int64), ("tiles_redone_jacobian", C.c_int64),
                ("rows_backward", C.c_int64), ("tiles_backward", C.c_int64), ("tiles_redone_backward", C.c_int64),
                ("backward_launches", C.c_int64), ("backward_ms", C.c_double)]


_lib = None

# every symbol include/hortimapping_b200.h declares (checked by tests/test_abi.py)
EXPORTS = ["hm_last_error", "hm_version", "hm_create", "hm_destroy", "hm_set_engine", "hm_get_engine", "hm_set_sparse_plan", "hm_plan_info", "hm_set_mask_reuse", "hm_calibrate",
           "hm_get_counters", "hm_saturation_count", "hm_profile_enable", "hm_sdf_forward", "hm_sdf_forward_rows", "hm_sdf_jacobian", "hm_sdf_jacobian_rows",
           "hm_voxel_grid", "hm_sdf_grid", "hm_sdf_loss", "hm_render_loss", "hm_optimize_shape", "hm_optimize_joint",
           "hm_get_last_system", "hm_optimize_shape_host", "hm_optimize_joint_host", "hm_isosurface", "hm_isosurface_fetch", "hm_nn_distance", "hm_frame_id_bboxes", "hm_crop_candidates", "hm_gather_rays",
           "hm_dbscan", "hm_cloud_bounds", "hm_crop_mean_offset"]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(hortimapping_b200 has no CPU/PyTorch fallback)")
    _lib = bind(C.CDLL(LIB_PATH))
    return _lib


def bind(L: C.CDLL) -> C.CDLL:
    """Declares the argument types of every entry point of include/hortimapping_b200.h on a loaded library."""
    L.hm_last_error.restype = C.c_char_p
    L.hm_version.restype = C.c_int
    L.hm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(DecoderDesc)]
    L.hm_destroy.argtypes = [C.c_void_p]
    L.hm_destroy.restype = None
    L.hm_set_engine.argtypes = [C.c_void_p, C.c_int]
    L.hm_get_engine.argtypes = [C.c_void_p]
    L.hm_set_sparse_plan.argtypes = [C.c_void_p, C.c_int]
    L.hm_plan_info.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.hm_set_mask_reuse.argtypes = [C.c_void_p, C.c_int]
    L.hm_calibrate.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.hm_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    L.hm_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.hm_saturation_count.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.hm_sdf_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.hm_sdf_forward_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.hm_sdf_jacobian.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hm_sdf_jacobian_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hm_voxel_grid.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]
    L.hm_sdf_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]
    L.hm_sdf_loss.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]
    L.hm_render_loss.argtypes = [C.c_void_p, C.POINTER(OptParams), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                 C.c_void_p, c_float_p, c_float_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
    L.hm_optimize_shape.argtypes = [C.c_void_p, C.POINTER(OptParams), C.POINTER(FruitBatch), C.c_void_p]
    L.hm_optimize_joint.argtypes = [C.c_void_p, C.POINTER(OptParams), C.POINTER(FruitBatch), C.c_void_p]
    L.hm_get_last_system.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hm_optimize_shape_host.argtypes = [C.c_void_p, C.POINTER(OptParams), C.POINTER(FruitBatch)]
    L.hm_optimize_joint_host.argtypes = [C.c_void_p, C.POINTER(OptParams), C.POINTER(FruitBatch)]
    L.hm_isosurface.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                C.c_void_p]
    L.hm_frame_id_bboxes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.hm_crop_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                     C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hm_gather_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]
    L.hm_nn_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.hm_isosurface_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]
    c_double_p = C.POINTER(C.c_double)
    L.hm_dbscan.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int32, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
    L.hm_cloud_bounds.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, c_double_p, c_double_p, C.c_void_p]
    L.hm_crop_mean_offset.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, C.POINTER(C.c_int64), c_double_p, C.c_void_p]
    return L


class HmError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        raise HmError(f"{what}: {lib().hm_last_error().decode(errors='replace')} (code {rc})")
