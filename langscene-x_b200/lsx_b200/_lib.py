"""ctypes binding of liblsx_b200.so (C ABI: include/lsx_rasterizer.h).

This is the ONLY compute path: if the library is missing or fails to load, importing raises — there is
no CPU / eager fallback.  PyTorch is used purely for device memory and stream handles.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_size_t, c_uint64, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LSX_B200_LIB", os.path.join(_PKG_DIR, "liblsx_b200.so"))

ALLOC_FN = ctypes.CFUNCTYPE(c_void_p, c_void_p, c_size_t)

ABI_VERSION = 5
MAX_BLEND_CHANNELS = 40


class ForwardArgs(ctypes.Structure):
    _fields_ = [
        ("P", c_int32), ("D", c_int32), ("M", c_int32), ("W", c_int32), ("H", c_int32),
        ("F", c_int32), ("Fi", c_int32),
        ("tanfovx", c_float), ("tanfovy", c_float), ("scale_modifier", c_float),
        ("prefiltered", c_int32), ("render_geo", c_int32), ("debug", c_int32), ("include_feature", c_int32),
        ("background", c_void_p), ("means3D", c_void_p), ("shs", c_void_p), ("colors_precomp", c_void_p),
        ("language_feature", c_void_p), ("language_feature_instance", c_void_p), ("opacities", c_void_p),
        ("scales", c_void_p), ("rotations", c_void_p), ("cov3D_precomp", c_void_p), ("all_map", c_void_p),
        ("viewmatrix", c_void_p), ("projmatrix", c_void_p), ("campos", c_void_p),
        ("out_color", c_void_p), ("out_language_feature", c_void_p), ("out_language_feature_instance", c_void_p),
        ("radii", c_void_p), ("out_observe", c_void_p), ("out_all_map", c_void_p), ("out_plane_depth", c_void_p),
        ("geom_alloc", ALLOC_FN), ("geom_user", c_void_p),
        ("binning_alloc", ALLOC_FN), ("binning_user", c_void_p),
        ("image_alloc", ALLOC_FN), ("image_user", c_void_p),
        ("stream", c_void_p),
        ("binning_capacity_hint", c_int32),
        ("raw_params", c_int32), ("pose", c_void_p),
    ]


class BackwardArgs(ctypes.Structure):
    _fields_ = [
        ("P", c_int32), ("D", c_int32), ("M", c_int32), ("W", c_int32), ("H", c_int32),
        ("F", c_int32), ("Fi", c_int32), ("R", c_int32),
        ("tanfovx", c_float), ("tanfovy", c_float), ("scale_modifier", c_float),
        ("render_geo", c_int32), ("debug", c_int32), ("include_feature", c_int32),
        ("background", c_void_p), ("means3D", c_void_p), ("shs", c_void_p), ("colors_precomp", c_void_p),
        ("language_feature", c_void_p), ("language_feature_instance", c_void_p), ("all_map", c_void_p),
        ("scales", c_void_p), ("rotations", c_void_p), ("cov3D_precomp", c_void_p),
        ("viewmatrix", c_void_p), ("projmatrix", c_void_p), ("campos", c_void_p), ("radii", c_void_p),
        ("out_all_map", c_void_p), ("geom_buffer", c_void_p), ("binning_buffer", c_void_p), ("image_buffer", c_void_p),
        ("dL_dout_color", c_void_p), ("dL_dout_language_feature", c_void_p),
        ("dL_dout_language_feature_instance", c_void_p), ("dL_dout_all_map", c_void_p),
        ("dL_dout_plane_depth", c_void_p),
        ("dL_dmeans2D", c_void_p), ("dL_dmeans2D_abs", c_void_p), ("dL_dconic", c_void_p), ("dL_dopacity", c_void_p),
        ("dL_dcolors", c_void_p), ("dL_dlanguage_feature", c_void_p), ("dL_dlanguage_feature_instance", c_void_p),
        ("dL_dmeans3D", c_void_p), ("dL_dcov3D", c_void_p), ("dL_dsh", c_void_p), ("dL_dscales", c_void_p),
        ("dL_drotations", c_void_p), ("dL_dall_map", c_void_p),
        ("stream", c_void_p),
        ("accumulate_param_grads", c_int32),
        ("binning_bytes", c_uint64),
        ("raw_params", c_int32), ("pose", c_void_p), ("dL_dpose", c_void_p), ("accumulate_pose", c_int32),
    ]


class ScratchLayout(ctypes.Structure):
    _fields_ = [
        ("depths", c_size_t), ("clamped", c_size_t), ("means2D", c_size_t), ("cov3D", c_size_t),
        ("conic_opacity", c_size_t), ("rgb", c_size_t), ("tiles_touched", c_size_t), ("records", c_size_t),
        ("record_stride", c_int32), ("geom_bytes", c_size_t),
        ("final_T", c_size_t), ("n_contrib", c_size_t), ("ranges", c_size_t), ("image_bytes", c_size_t),
        ("point_list", c_size_t), ("binning_bytes", c_size_t), ("masks", c_size_t),
        ("blk_list", c_size_t), ("blk_cnt", c_size_t), ("k_contrib", c_size_t),
    ]


class DensifyPlanArgs(ctypes.Structure):
    _fields_ = [
        ("P", c_int32), ("grad_accum", c_void_p), ("grad_accum_abs", c_void_p), ("denom", c_void_p), ("max_radii2D", c_void_p),
        ("scaling_raw", c_void_p), ("opacity_raw", c_void_p),
        ("max_grad", c_float), ("abs_max_grad", c_float), ("min_opacity", c_float), ("extent", c_float),
        ("percent_dense", c_float), ("abs_split_radii2D_threshold", c_float),
        ("max_all_points", ctypes.c_int64), ("max_abs_split_points", ctypes.c_int64),
        ("prune_world_size", c_int32), ("workspace", c_void_p), ("workspace_bytes", c_size_t), ("stream", c_void_p),
    ]


class DensifyPlanResult(ctypes.Structure):
    _fields_ = [
        ("P_new", ctypes.c_int64), ("n_clone", ctypes.c_int64), ("n_split", ctypes.c_int64), ("n_split_abs", ctypes.c_int64),
        ("n_kept_original", ctypes.c_int64), ("n_kept_clone", ctypes.c_int64), ("n_kept_split", ctypes.c_int64),
        ("clone_capped", c_int32), ("split_capped", c_int32), ("abs_capped", c_int32),
        ("row_map", c_void_p), ("noise_index", c_void_p),
    ]


class DensifyApplyArgs(ctypes.Structure):
    _fields_ = [
        ("P_new", ctypes.c_int64), ("n_new_rows", ctypes.c_int64), ("n_groups", c_int32),
        ("old_begin", POINTER(ctypes.c_int64)), ("new_begin", POINTER(ctypes.c_int64)),
        ("width", POINTER(c_int32)), ("role", POINTER(c_int32)),
        ("row_map", c_void_p), ("noise_index", c_void_p), ("z_clone", c_void_p), ("z_split", c_void_p),
        ("old_params", c_void_p), ("old_exp_avg", c_void_p), ("old_exp_avg_sq", c_void_p),
        ("new_params", c_void_p), ("new_exp_avg", c_void_p), ("new_exp_avg_sq", c_void_p),
        ("stream", c_void_p), ("group_align", c_int32),
    ]


DENSIFY_ROLE = {"copy": 0, "xyz": 1, "scaling": 2, "rotation": 3}

# every symbol include/lsx_rasterizer.h declares: name -> (restype, argtypes)
EXPORTS = {
    "lsx_rasterize_forward": (c_int32, [POINTER(ForwardArgs), POINTER(c_int32)]),
    "lsx_rasterize_backward": (c_int32, [POINTER(BackwardArgs)]),
    "lsx_mark_visible": (c_int32, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lsx_knn_mean_dist2": (c_int32, [c_int32, c_void_p, c_void_p, ALLOC_FN, c_void_p, c_void_p]),
    "lsx_depth_normal_forward": (c_int32, [c_int32, c_int32, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                           c_void_p]),
    "lsx_depth_normal_backward": (c_int32, [c_int32, c_int32, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_void_p]),
    "lsx_image_loss_num_blocks": (ctypes.c_int64, [c_int32, c_int32, c_int32]),
    "lsx_image_loss_forward": (c_int32, [c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lsx_image_loss_backward": (c_int32, [c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p]),
    "lsx_arena_adam_step": (c_int32, [ctypes.c_int64, c_int32, POINTER(ctypes.c_int64), POINTER(c_float), c_int32, c_float, c_float,
                                      c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lsx_gaussian_head_forward": (c_int32, [c_int32] + [c_void_p] * 11),
    "lsx_gaussian_head_backward": (c_int32, [c_int32] + [c_void_p] * 16),
    "lsx_gaussian_head_backward_acc": (c_int32, [c_int32] + [c_void_p] * 15 + [c_int32, c_void_p]),
    "lsx_densify_stats_update": (c_int32, [c_int32] + [c_void_p] * 9),
    "lsx_densify_workspace_bytes": (c_size_t, [c_int32]),
    "lsx_densify_plan": (c_int32, [POINTER(DensifyPlanArgs), POINTER(DensifyPlanResult)]),
    "lsx_densify_apply": (c_int32, [POINTER(DensifyApplyArgs)]),
    "lsx_reset_opacity": (c_int32, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lsx_pose_num_partials": (c_int32, []),
    "lsx_pose_transform_forward": (c_int32, [c_int32] + [c_void_p] * 6),
    "lsx_pose_transform_backward": (c_int32, [c_int32] + [c_void_p] * 10),
    "lsx_masked_l1_num_blocks": (c_int32, [ctypes.c_int64]),
    "lsx_masked_l1_forward": (c_int32, [c_int32, c_int32, c_int32, c_int32] + [c_void_p] * 5),
    "lsx_masked_l1_backward": (c_int32, [c_int32, c_int32, c_int32, c_int32] + [c_void_p] * 6),
    "lsx_cls3d_scratch_bytes": (ctypes.c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "lsx_cls3d_forward": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_float] + [c_void_p] * 8),
    "lsx_cls3d_backward": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_float] + [c_void_p] * 8),
    "lsx_cls3d_forward_tree": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_float] + [c_void_p] * 9),
    "lsx_knn_tree_bytes": (ctypes.c_int64, [c_int32]),
    "lsx_knn_tree_build": (c_int32, [c_int32, c_void_p, c_void_p, c_void_p]),
    "lsx_rows_pack": (c_int32, [ctypes.c_int64, c_int32] + [c_void_p] * 5),
    "lsx_rows_unpack": (c_int32, [ctypes.c_int64, c_int32] + [c_void_p] * 5),
    "lsx_scratch_layout_query": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_int32, POINTER(ScratchLayout)]),
    "lsx_debug_sorted_keys": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_size_t, c_void_p,
                                        c_void_p, c_void_p]),
    "lsx_binning_capacity": (c_int32, [c_size_t, c_int32, c_int32]),
    "lsx_render_stats": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                                   c_void_p]),
    "lsx_kernel_launch_count": (c_uint64, []),
    "lsx_profile_enable": (None, [c_int32]),
    "lsx_profile_read": (c_int32, [POINTER(c_float), c_int32]),
    "lsx_microbench": (c_int32, [c_int32, c_void_p, c_size_t, POINTER(ctypes.c_double), c_void_p]),
    "lsx_last_error": (c_char_p, []),
    "lsx_abi_version": (c_int32, []),
}

# lsx_backward_args.accumulate_param_grads bits (LSX_ACC_* in the header), by the names ops.py uses for grad_buffers
ACC_BITS = {"means3D": 0x001, "sh": 0x002, "opacity": 0x004, "scales": 0x008, "rotations": 0x010, "colors": 0x020,
            "language_feature": 0x040, "instance_feature": 0x080, "all_map": 0x100, "cov3D": 0x200}

_lib = None


def load():
    """Load (once) and return the ctypes handle.  Raises if the CUDA library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the sm_100a CUDA library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root). "
            "There is no CPU fallback for this operator."
        )
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    ver = lib.lsx_abi_version()
    if ver != ABI_VERSION:
        raise ImportError(f"liblsx_b200.so ABI version {ver} != binding version {ABI_VERSION}")
    _lib = lib
    return lib


def last_error():
    msg = load().lsx_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status, what):
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}): {last_error()}")


STAGES = ["preprocess_fwd", "depth_sort", "offsets_scan", "emit", "tile_sort", "tile_ranges", "footprint_masks", "render_fwd", "bwd_zero",
          "render_bwd", "preprocess_bwd", "knn"]


def profile_enable(on=True):
    load().lsx_profile_enable(1 if on else 0)


def profile_read():
    """dict stage -> milliseconds accumulated since the last read (synchronises the recorded events)."""
    buf = (c_float * len(STAGES))()
    load().lsx_profile_read(buf, len(STAGES))
    return {name: float(buf[i]) for i, name in enumerate(STAGES)}


def kernel_launch_count():
    return int(load().lsx_kernel_launch_count())
