"""ctypes binding of liborv_b200.so (the C ABI declared in include/orv_b200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or the device is not a B200
(compute capability 10.x) the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
# ORVB_LIB_PATH: measurement builds only (e.g. the -DORVB_GEMM_TIMELINE variant tools/ build next to the product library)
LIB_PATH = Path(os.environ["ORVB_LIB_PATH"]).resolve() if os.environ.get("ORVB_LIB_PATH") else _HERE / "liborv_b200.so"

ORVB_OK = 0
EPI_BIAS, EPI_GELU, EPI_GATE_RESID, EPI_QKV = 0, 1, 2, 3

c_void_p = C.c_void_p
c_int = C.c_int32
c_float = C.c_float


class RowMap(C.Structure):
    _fields_ = [("seq_len", c_int), ("text_len", c_int), ("tokens_per_group", c_int), ("groups_per_batch", c_int)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", c_void_p), ("w", c_void_p), ("out", c_void_p), ("bias", c_void_p),
        ("m", c_int), ("n", c_int), ("k", c_int),
        ("lda", c_int), ("ldw", c_int), ("ldo", c_int),
        ("epilogue", c_int),
        ("src_rows", c_int), ("dst_rows", c_int), ("dst_offset", c_int),
        ("mv_tokens", c_int), ("mv_frames", c_int), ("mv_views", c_int),
        ("resid", c_void_p), ("ldr", c_int),
        ("resid_mod", c_int), ("resid_views", c_int), ("resid_view_stride", c_int),
        ("gate", c_void_p),
        ("gate_ld", c_int), ("gate_text_off", c_int), ("gate_video_off", c_int),
        ("rowmap", RowMap),
        ("qk_dim", c_int),
        ("q_norm_w", c_void_p), ("q_norm_b", c_void_p), ("k_norm_w", c_void_p), ("k_norm_b", c_void_p),
        ("qk_eps", c_float),
        ("rope_cos", c_void_p), ("rope_sin", c_void_p),
        ("out_f32", c_int),
        ("group_offset", c_void_p),
        ("k_wrap", c_int),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("qkv", c_void_p), ("out", c_void_p),
        ("batch", c_int), ("seq_len", c_int), ("heads", c_int), ("scale", c_float),
        ("q_row0", c_int), ("q_rows", c_int), ("out_f32", c_int),
    ]


class LnArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("y", c_void_p), ("ln_w", c_void_p), ("ln_b", c_void_p),
        ("rows", c_int), ("dim", c_int), ("eps", c_float),
        ("mod", c_void_p), ("mod_ld", c_int), ("text_off", c_int), ("video_off", c_int), ("scale_first", c_int),
        ("rowmap", RowMap),
        ("in_video_only", c_int),
        ("pre_w", c_void_p), ("pre_b", c_void_p), ("pre_eps", c_float),
        ("ab", c_void_p), ("ab_ld", c_int),
        ("y_f32", c_int),
        ("group_offset", c_void_p),
    ]


class Config(C.Structure):
    _fields_ = [
        ("dim", c_int), ("heads", c_int), ("head_dim", c_int), ("layers", c_int), ("ff_dim", c_int),
        ("time_embed_dim", c_int), ("text_embed_dim", c_int), ("in_channels", c_int), ("out_channels", c_int),
        ("patch_size", c_int), ("patch_size_t", c_int), ("use_rope", c_int), ("has_ofs", c_int),
        ("ofs_embed_dim", c_int), ("flip_sin_to_cos", c_int), ("freq_shift", c_float), ("norm_eps", c_float),
        ("visual_guidance", c_int), ("num_control_keys", c_int), ("multiview", c_int), ("max_n_view", c_int),
        ("action_state_dim", c_int), ("action_compress", c_int), ("action_hidden", c_int),
        ("modulate_text", c_int),
    ]


_BLOCK_FIELDS = [
    "norm1_lin_w", "norm1_lin_b", "norm1_ln_w", "norm1_ln_b", "qkv_w", "qkv_b", "q_norm_w", "q_norm_b",
    "k_norm_w", "k_norm_b", "out_w", "out_b", "norm2_lin_w", "norm2_lin_b", "norm2_ln_w", "norm2_ln_b",
    "ff1_w", "ff1_b", "ff2_w", "ff2_b", "proj_out_w", "proj_out_b",
]


class BlockWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in _BLOCK_FIELDS]


_WEIGHT_FIELDS = [
    "patch_w", "patch_b", "text_w", "text_b", "pos_embed", "pos_embed_plain", "time1_w", "time1_b", "time2_w", "time2_b",
    "ofs1_w", "ofs1_b", "ofs2_w", "ofs2_b", "act1_w", "act1_b", "act2_w", "act2_b", "act_mask_embed",
    "combine_w", "combine_b", "norm_final_w", "norm_final_b", "norm_out_lin_w", "norm_out_lin_b",
    "norm_out_ln_w", "norm_out_ln_b", "proj_out_w", "proj_out_b",
]


class Weights(C.Structure):
    _fields_ = [(n, c_void_p) for n in _WEIGHT_FIELDS] + [
        ("blocks_host", C.POINTER(BlockWeights)), ("mv_blocks_host", C.POINTER(BlockWeights))]


class Shape(C.Structure):
    _fields_ = [("batch", c_int), ("views", c_int), ("frames", c_int), ("height", c_int), ("width", c_int),
                ("text_len", c_int), ("action_frames", c_int)]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("shape", Shape),
        ("hidden_states", c_void_p), ("text", c_void_p), ("timesteps", c_void_p), ("ofs", c_float),
        ("actions", c_void_p), ("action_mask", c_void_p), ("depths", c_void_p), ("labels", c_void_p),
        ("rope_cos", c_void_p), ("rope_sin", c_void_p),
        ("out", c_void_p), ("workspace", c_void_p), ("workspace_bytes", C.c_size_t),
        ("tap_hidden", c_void_p), ("tap_layer", c_int),
        ("skip_modulation", c_int),
        ("static_mode", c_int),
        ("schedule", c_void_p), ("schedule_steps", c_int), ("schedule_row_offset", c_void_p),
    ]


STATIC_COMPUTE, STATIC_SAVE, STATIC_REUSE = 0, 1, 2


class SamplerStepArgs(C.Structure):
    _fields_ = [
        ("model_out", c_void_p), ("latents", c_void_p), ("old_x0", c_void_p), ("noise", c_void_p),
        ("n", C.c_int64), ("cfg_copies", c_int), ("guidance_scale", c_float),
        ("c_x", c_float), ("c_v", c_float), ("d_cur", c_float), ("d_old", c_float), ("k_x", c_float),
        ("k_d", c_float), ("k_noise", c_float),
        ("next_input", c_void_p), ("lat_channels", c_int), ("img_channels", c_int), ("hw", c_int),
    ]


class VoxelizeArgs(C.Structure):
    _fields_ = [
        ("points", c_void_p), ("n", c_int), ("c", c_int),
        ("voxel_size", c_float * 3), ("coors_range", c_float * 6),
        ("max_points", c_int), ("max_voxels", c_int),
        ("voxels", c_void_p), ("coors", c_void_p), ("num_points_per_voxel", c_void_p), ("voxel_num", c_void_p),
        ("voxel_labels", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class GsArgs(C.Structure):
    _fields_ = [
        ("p", c_int),
        ("means3d", c_void_p), ("colors", c_void_p), ("features", c_void_p), ("opacities", c_void_p),
        ("scales", c_void_p), ("rotations", c_void_p), ("cov3d", c_void_p), ("scale_modifier", c_float),
        ("viewmatrix", c_void_p), ("projmatrix", c_void_p), ("background", c_void_p),
        ("tan_fovx", c_float), ("tan_fovy", c_float), ("height", c_int), ("width", c_int),
        ("out_color", c_void_p), ("out_feature", c_void_p), ("out_depth", c_void_p), ("out_alpha", c_void_p),
        ("radii", c_void_p), ("num_rendered", c_void_p), ("max_instances", c_int),
        ("workspace", c_void_p), ("workspace_bytes", C.c_size_t),
    ]


# Every symbol include/orv_b200.h declares; tests check the .so exports all of them.
class ConvArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("cache", c_void_p), ("w", c_void_p), ("bias", c_void_p), ("resid", c_void_p), ("out", c_void_p),
        ("frames", c_int), ("height", c_int), ("width", c_int), ("c_in", c_int), ("c_out", c_int),
        ("kt", c_int), ("kh", c_int), ("kw", c_int), ("out_f32", c_int),
        ("gn_stats", c_void_p), ("gn_scratch", c_void_p), ("gn_groups", c_int), ("gn_eps", c_float),
    ]


class SpatialNormArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("y", c_void_p),
        ("frames", c_int), ("height", c_int), ("width", c_int), ("channels", c_int), ("groups", c_int),
        ("stats", c_void_p), ("gamma", c_void_p), ("beta", c_void_p), ("table", c_void_p),
        ("table_ld", c_int), ("y_off", c_int), ("b_off", c_int),
        ("t_src", c_void_p), ("lat_h", c_int), ("lat_w", c_int), ("shift", c_int), ("act", c_int), ("y_f32", c_int),
    ]


EXPORTED_SYMBOLS = [
    "orvb_version", "orvb_last_error", "orvb_check_device",
    "orvb_gemm_bf16", "orvb_gemm_bf16_bn", "orvb_gemm_tile_width", "orvb_gemm_tile_remainder", "orvb_gemm_tile_list", "orvb_attention_bf16", "orvb_attention", "orvb_attention_set_rescale_threshold",
    "orvb_attention_set_debug", "orvb_ln_modulate", "orvb_skinny_linear",
    "orvb_patchify", "orvb_unpatchify",
    "orvb_model_create", "orvb_model_destroy", "orvb_model_bind_weights", "orvb_workspace_bytes",
    "orvb_forward", "orvb_last_launch_count", "orvb_last_launch_classes", "orvb_model_set_profile", "orvb_model_get_profile",
    "orvb_modulation_bytes", "orvb_modulation_schedule", "orvb_modulation_select",
    "orvb_sampler_step",
    "orvb_dynamic_voxelize", "orvb_voxelize_workspace_bytes", "orvb_hard_voxelize",
    "orvb_gs_workspace_bytes", "orvb_gs_rasterize",
    "orvb_conv_cl", "orvb_conv_gn_scratch_bytes", "orvb_gn_scratch_bytes", "orvb_gn_stats_cl", "orvb_spatial_norm_cl", "orvb_upsample2x_cl",
    "orvb_cl_to_planar",
]

_lib = None


def load() -> C.CDLL:
    """Loads liborv_b200.so (building it first if the sources are newer and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if os.environ.get("ORVB_NO_BUILD"):
            raise RuntimeError(f"{LIB_PATH} is missing and ORVB_NO_BUILD is set")
        from . import build as _build
        _build.build()
    lib = C.CDLL(str(LIB_PATH))
    lib.orvb_version.restype = c_int
    lib.orvb_last_error.restype = C.c_char_p
    lib.orvb_check_device.restype = c_int
    lib.orvb_gemm_bf16.argtypes = [C.POINTER(GemmArgs), c_void_p]
    lib.orvb_gemm_bf16.restype = c_int
    if hasattr(lib, "orvb_gemm_tile_width"):
        lib.orvb_gemm_tile_width.argtypes = [c_int, c_int, c_int]
        lib.orvb_gemm_tile_width.restype = c_int
    if hasattr(lib, "orvb_gemm_tile_list"):
        lib.orvb_gemm_tile_list.argtypes = [c_int, c_int, c_int, c_int, c_int, c_void_p, c_int]
        lib.orvb_gemm_tile_list.restype = c_int
    if hasattr(lib, "orvb_gemm_tile_remainder"):
        lib.orvb_gemm_tile_remainder.argtypes = [c_int, c_int, c_int]
        lib.orvb_gemm_tile_remainder.restype = c_int
    if hasattr(lib, "orvb_gemm_bf16_bn"):
        lib.orvb_gemm_bf16_bn.argtypes = [C.POINTER(GemmArgs), c_int, c_void_p]
        lib.orvb_gemm_bf16_bn.restype = c_int
    if hasattr(lib, "orvb_attention_set_rescale_threshold"):
        lib.orvb_attention_set_rescale_threshold.argtypes = [c_float]
        lib.orvb_attention_set_rescale_threshold.restype = None
        lib.orvb_attention_set_debug.argtypes = [c_void_p]
        lib.orvb_attention_set_debug.restype = None
    if hasattr(lib, "orvb_attention_bf16"):
        lib.orvb_attention_bf16.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]
        lib.orvb_attention_bf16.restype = c_int
        lib.orvb_attention.argtypes = [C.POINTER(AttentionArgs), c_void_p]
        lib.orvb_attention.restype = c_int
    if hasattr(lib, "orvb_ln_modulate"):
        lib.orvb_ln_modulate.argtypes = [C.POINTER(LnArgs), c_void_p]
        lib.orvb_ln_modulate.restype = c_int
    if hasattr(lib, "orvb_skinny_linear"):
        lib.orvb_skinny_linear.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                           c_void_p]
        lib.orvb_skinny_linear.restype = c_int
    for name in ("orvb_patchify", "orvb_unpatchify"):
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
            fn.restype = c_int
    if hasattr(lib, "orvb_model_create"):
        lib.orvb_model_create.argtypes = [C.POINTER(Config), C.POINTER(c_void_p)]
        lib.orvb_model_create.restype = c_int
        lib.orvb_model_destroy.argtypes = [c_void_p]
        lib.orvb_model_destroy.restype = None
        lib.orvb_model_bind_weights.argtypes = [c_void_p, C.POINTER(Weights)]
        lib.orvb_model_bind_weights.restype = c_int
        lib.orvb_workspace_bytes.argtypes = [c_void_p, C.POINTER(Shape)]
        lib.orvb_workspace_bytes.restype = C.c_size_t
        lib.orvb_forward.argtypes = [c_void_p, C.POINTER(ForwardArgs), c_void_p]
        lib.orvb_forward.restype = c_int
        lib.orvb_last_launch_count.argtypes = [c_void_p]
        lib.orvb_last_launch_count.restype = c_int
        lib.orvb_last_launch_classes.argtypes = [c_void_p, C.POINTER(c_int), c_int]
        lib.orvb_last_launch_classes.restype = c_int
        lib.orvb_model_set_profile.argtypes = [c_void_p, c_int]
        lib.orvb_model_set_profile.restype = c_int
        lib.orvb_model_get_profile.argtypes = [c_void_p, C.POINTER(c_float), C.POINTER(c_int)]
        lib.orvb_model_get_profile.restype = c_int
        lib.orvb_modulation_bytes.argtypes = [c_void_p, C.POINTER(Shape), c_int]
        lib.orvb_modulation_bytes.restype = C.c_size_t
        lib.orvb_modulation_schedule.argtypes = [c_void_p, C.POINTER(Shape), c_int, c_void_p, c_float, c_void_p, c_void_p,
                                                 c_void_p, C.c_size_t, c_void_p]
        lib.orvb_modulation_schedule.restype = c_int
        lib.orvb_modulation_select.argtypes = [c_void_p, C.POINTER(Shape), c_int, c_int, c_void_p, c_void_p, c_void_p]
        lib.orvb_modulation_select.restype = c_int
    if hasattr(lib, "orvb_sampler_step"):
        lib.orvb_sampler_step.argtypes = [C.POINTER(SamplerStepArgs), c_void_p]
        lib.orvb_sampler_step.restype = c_int
    if hasattr(lib, "orvb_hard_voxelize"):
        lib.orvb_dynamic_voxelize.argtypes = [c_void_p, c_int, c_int, C.POINTER(c_float), C.POINTER(c_float), c_void_p,
                                              c_void_p]
        lib.orvb_dynamic_voxelize.restype = c_int
        lib.orvb_voxelize_workspace_bytes.argtypes = [c_int, c_int]
        lib.orvb_voxelize_workspace_bytes.restype = C.c_size_t
        lib.orvb_hard_voxelize.argtypes = [C.POINTER(VoxelizeArgs), c_void_p]
        lib.orvb_hard_voxelize.restype = c_int
    if hasattr(lib, "orvb_gs_rasterize"):
        lib.orvb_gs_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
        lib.orvb_gs_workspace_bytes.restype = C.c_size_t
        lib.orvb_gs_rasterize.argtypes = [C.POINTER(GsArgs), c_void_p]
        lib.orvb_gs_rasterize.restype = c_int
    if hasattr(lib, "orvb_conv_cl"):
        lib.orvb_conv_cl.argtypes = [C.POINTER(ConvArgs), c_void_p]
        lib.orvb_conv_cl.restype = c_int
        lib.orvb_conv_gn_scratch_bytes.argtypes = []
        lib.orvb_conv_gn_scratch_bytes.restype = C.c_size_t
        lib.orvb_gn_scratch_bytes.argtypes = [C.c_int64, c_int]
        lib.orvb_gn_scratch_bytes.restype = C.c_size_t
        lib.orvb_gn_stats_cl.argtypes = [c_void_p, C.c_int64, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]
        lib.orvb_gn_stats_cl.restype = c_int
        lib.orvb_spatial_norm_cl.argtypes = [C.POINTER(SpatialNormArgs), c_void_p]
        lib.orvb_spatial_norm_cl.restype = c_int
        lib.orvb_upsample2x_cl.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
        lib.orvb_upsample2x_cl.restype = c_int
        lib.orvb_cl_to_planar.argtypes = [c_void_p, c_void_p, C.c_int64, c_int, c_int, c_void_p]
        lib.orvb_cl_to_planar.restype = c_int
    _lib = lib
    return lib


def check(rc: int, what: str = "liborv_b200") -> None:
    if rc != ORVB_OK:
        msg = load().orvb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
