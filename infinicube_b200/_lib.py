"""ctypes binding of libinfinicube_b200.so (the C ABI declared in include/infinicube_b200.h).

There is no fallback: if the shared library is missing or the device is not a B200, calls raise.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(__import__("os").environ.get("ICB_LIB_PATH") or PKG_DIR / "libinfinicube_b200.so")  # override: developer A/B builds
HEADER_PATH = PKG_DIR.parent / "include" / "infinicube_b200.h"

IC_OK = 0
IC_DTYPE_F32 = 0
IC_DTYPE_BF16 = 1


class ICError(RuntimeError):
    pass


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p), ("bias_per_row", C.c_int), ("act", C.c_int),
        ("out_bf16", C.c_void_p), ("ld_out", C.c_int),
        ("rowss", C.c_void_p), ("rowss_ld", C.c_int),
        ("out_f32", C.c_void_p), ("ld_f32", C.c_int),
        ("addend", C.c_void_p), ("ld_add", C.c_int),
        ("resid", C.c_void_p), ("ld_res", C.c_int),
        ("gate", C.c_void_p),
    ]


class DitConfig(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("ffn_dim", C.c_int), ("num_heads", C.c_int), ("num_layers", C.c_int),
        ("in_dim", C.c_int), ("out_dim", C.c_int), ("text_dim", C.c_int), ("freq_dim", C.c_int),
        ("text_len", C.c_int), ("guide_channels", C.c_int), ("eps", C.c_float),
        ("lat_f", C.c_int), ("lat_h", C.c_int), ("lat_w", C.c_int),
        ("frame0", C.c_int), ("frames_local", C.c_int), ("world_size", C.c_int), ("rank", C.c_int),
    ]


_lib = None


def declared_symbols() -> list[str]:
    """Every function name include/infinicube_b200.h declares."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ic_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ICError(
            f"{LIB_PATH} is missing: build it with `python -m infinicube_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)"
        )
    L = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    vp, ci, cf, ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    L.ic_version.restype = ci
    L.ic_error_string.restype = C.c_char_p
    L.ic_error_string.argtypes = [ci]
    L.ic_device_check.restype = ci
    L.ic_gemm_block_n.argtypes = [ci]
    L.ic_gemm_bf16.argtypes = [vp, ci, vp, ci, ci, ci, ci, C.POINTER(GemmEpilogue), vp]
    L.ic_fmha_fwd.argtypes = [vp, ci, vp, ci, ll, vp, ci, ll, vp, ci, ci, ci, ci, ci, cf, vp]
    L.ic_ln_modulate.argtypes = [vp, ci, vp, vp, ci, vp, ci, ci, ci, cf, vp]
    L.ic_rmsnorm_rope.argtypes = [vp, ci, vp, ci, ci, ci, vp, vp, ci, ci, ci, cf, vp, vp, vp, ci, ci, ci, ci, vp]
    L.ic_patchify.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp]
    L.ic_unpatchify_cfg_step.argtypes = [vp, vp, vp, ci, ci, ci, ci, cf, cf, vp, vp]
    L.ic_dit_create.argtypes = [C.POINTER(DitConfig), C.POINTER(vp)]
    L.ic_dit_destroy.argtypes = [vp]
    L.ic_dit_workspace_bytes.argtypes = [vp]
    L.ic_dit_workspace_bytes.restype = ll
    L.ic_dit_load_tensor.argtypes = [vp, C.c_char_p, vp, ci, ll, vp]
    L.ic_nccl_unique_id.argtypes = [vp]
    L.ic_dit_init_comm.argtypes = [vp, vp]
    L.ic_dit_p2p_export.argtypes = [vp, vp]
    L.ic_dit_p2p_attach.argtypes = [vp, vp]
    L.ic_dit_p2p_enabled.argtypes = [vp]
    L.ic_dit_set_context.argtypes = [vp, ci, vp, ci, vp]
    L.ic_dit_set_guidance.argtypes = [vp, vp, vp]
    L.ic_dit_forward.argtypes = [vp, vp, cf, ci, vp, vp]
    L.ic_dit_embed.argtypes = [vp, vp, cf, vp]
    L.ic_dit_run_block.argtypes = [vp, ci, ci, vp]
    L.ic_dit_head.argtypes = [vp, vp, vp]
    L.ic_dit_run_block_phase.argtypes = [vp, ci, ci, ci, vp]
    L.ic_dit_kv_segment.argtypes = [vp, ci, C.POINTER(vp), C.POINTER(ll)]
    L.ic_dit_missing_tensors.argtypes = [vp, C.c_char_p, ci]
    L.ic_dit_tokens.argtypes = [vp]
    L.ic_dit_tokens.restype = vp
    L.ic_dit_flops_per_forward.argtypes = [vp]
    L.ic_dit_flops_per_forward.restype = ll
    L.ic_dit_launch_count.argtypes = [vp]
    L.ic_dit_set_profiling.argtypes = [vp, ci]
    L.ic_dit_profile_collect.argtypes = [vp, C.POINTER(cf), C.POINTER(ci)]
    L.ic_grid_build.argtypes = [vp, ll, C.POINTER(cf), C.POINTER(cf), vp, vp, C.POINTER(vp), vp]
    L.ic_grid_destroy.argtypes = [vp]
    L.ic_grid_num_voxels.argtypes = [vp]
    L.ic_grid_num_voxels.restype = ll
    L.ic_grid_num_bricks.argtypes = [vp]
    L.ic_grid_num_bricks.restype = ll
    L.ic_grid_info.argtypes = [vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]
    L.ic_grid_export.argtypes = [vp, vp, vp, vp, vp]
    L.ic_raster_render.argtypes = [vp, C.POINTER(cf), vp, ci, ci, ci, vp, vp, ci, ci, vp, vp, vp, vp]
    L.ic_semantic_rgb.argtypes = [vp, vp, vp, ll, vp, ci, vp, vp, ci, vp, vp]
    L.ic_lut_gather_f32.argtypes = [vp, ll, vp, ci, vp, vp]
    L.ic_instance_from_boxes.argtypes = [vp, ll, vp, vp, ci, C.c_uint, vp, vp]
    L.ic_coord_unproject.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp]
    L.ic_coord_normalize.argtypes = [vp, vp, ll, vp, vp, vp, vp, vp]
    L.ic_mesh_voxelize_mask.argtypes = [vp, ci, vp, ci, C.c_double, C.c_double, C.POINTER(ci), C.POINTER(ci), vp, vp]
    L.ic_conv_cl.argtypes = [vp, ci, ci, ci, ci, vp, vp, C.POINTER(ci), ci, vp, ci, ci, ci, ci, ci, vp, ci, vp]
    L.ic_rmsnorm_cl.argtypes = [vp, vp, vp, ll, ci, ci, vp]
    L.ic_upsample2x_cl.argtypes = [vp, vp, ci, ci, ci, ci, vp]
    L.ic_time_interleave_cl.argtypes = [vp, vp, vp, ci, ll, ci, vp]
    L.ic_time_gather3_cl.argtypes = [vp, vp, ci, ll, ci, vp]
    L.ic_space_to_depth_cl.argtypes = [vp, vp, ci, ci, ci, ci, vp]
    L.ic_softmax_rows.argtypes = [vp, ci, vp, ci, ci, ci, cf, vp]
    L.ic_frames_to_cl.argtypes = [vp, vp, ll, ci, vp]
    L.ic_latent_to_cl.argtypes = [vp, vp, vp, vp, ci, ll, ci, vp]
    L.ic_cl_to_cf.argtypes = [vp, ci, vp, vp, vp, ci, ll, vp]
    L.ic_blend_accumulate.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
    L.ic_blend_finalize.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp, vp]
    L.ic_t5_embed.argtypes = [vp, ci, vp, ci, ci, vp, ci, vp]
    L.ic_t5_rmsnorm.argtypes = [vp, ci, vp, vp, ci, ci, ci, cf, ci, vp]
    L.ic_t5_attention.argtypes = [vp, vp, vp, ci, vp, vp, vp, ci, ci, ci, vp]
    L.ic_mul_bf16.argtypes = [vp, vp, vp, ll, vp]
    L.ic_knn_build.argtypes = [vp, ll, ci, cf, C.POINTER(vp), vp]
    L.ic_knn_destroy.argtypes = [vp]
    L.ic_knn_info.argtypes = [vp, C.POINTER(ll), C.POINTER(ll), C.POINTER(cf), C.POINTER(ci)]
    L.ic_knn_query1.argtypes = [vp, vp, ll, ci, vp, vp, vp, vp, vp]
    L.ic_knn_query.argtypes = [vp, vp, ll, ci, ci, vp, vp, vp]
    _lib = L
    return L


def check(code: int, what: str = "") -> None:
    if code != IC_OK:
        msg = lib().ic_error_string(code).decode()
        raise ICError(f"{what or 'infinicube_b200 call'} failed: {msg} ({code})")


def require_device() -> None:
    check(lib().ic_device_check(), "device check")
