"""ctypes binding of libs2v.so (the C ABI declared in include/s2v.h).

The product path has NO CPU fallback: if the CUDA library is missing or the device is
not sm_100, loading raises.  Build the library with ``python __graft_entry__.py`` (or
``python speech-to-video-mpp_b200/build.py``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_i32, c_i64, c_f32, c_f64, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p


class View(C.Structure):
    _fields_ = [("ptr", c_vp), ("n", c_i32), ("h", c_i32), ("w", c_i32), ("c", c_i32),
                ("sn", c_i64), ("sh", c_i64), ("sw", c_i64)]


class Conv(C.Structure):
    _fields_ = [("x", View), ("y", View), ("w", c_vp), ("scale", c_vp), ("bias", c_vp),
                ("res1", View), ("res2", View),
                ("kh", c_i32), ("kw", c_i32), ("stride_h", c_i32), ("stride_w", c_i32),
                ("pad_h", c_i32), ("pad_w", c_i32), ("dil_h", c_i32), ("dil_w", c_i32),
                ("pad_mode", c_i32), ("up2", c_i32), ("act", c_i32), ("act_param", c_f32),
                ("out_mode", c_i32), ("y_f32", c_vp),
                ("x2", View), ("k2h", c_i32), ("k2w", c_i32), ("pad2_h", c_i32), ("pad2_w", c_i32),
                ("stats_partial", c_vp), ("stats_c_off", c_i32), ("stats_c_total", c_i32),
                ("stats_chunk_off", c_i32), ("stats_chunks_total", c_i32), ("stats_groups", c_i32), ("stats_gmax", c_i32),
                ("narrow_cin_from", c_i32), ("narrow_cout", c_i32)]


class LinGroup(C.Structure):
    _fields_ = [("wt", c_vp), ("bias", c_vp), ("in_off", c_i32), ("k", c_i32), ("out_off", c_i32), ("nout", c_i32)]


ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID, ACT_TANH, ACT_GELU = range(6)
OUT_F16_NHWC, OUT_F32_NCHW = 0, 1
PAD_ZERO, PAD_REFLECT = 0, 1

VP = C.POINTER(View)

_SIGNATURES = {
    "s2v_strerror": (C.c_char_p, [C.c_int]),
    "s2v_version": (C.c_int, []),
    "s2v_device_ok": (C.c_int, []),
    "s2v_last_cuda_error": (C.c_char_p, []),
    "s2v_mel_num_frames": (C.c_int, [c_i64]),
    "s2v_melspectrogram_f32": (C.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, C.c_int, c_vp]),
    "s2v_mel_init": (C.c_int, []),
    "s2v_resample_out_len": (c_i64, [c_i64, C.c_int, C.c_int]),
    "s2v_resample_f32": (C.c_int, [c_vp, c_i64, C.c_int, C.c_int, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_i64, c_vp]),
    "s2v_pcm_to_mono_f32": (C.c_int, [c_vp, C.c_int, C.c_int, c_i64, c_vp, c_vp]),
    "s2v_mel_window_count": (c_i64, [c_i64, c_f64]),
    "s2v_mel_window_starts_host": (C.c_int, [c_i64, c_f64, C.POINTER(c_i32), c_i64]),
    "s2v_mel_windows_f32": (C.c_int, [c_vp, c_i64, c_f64, c_i64, c_i64, c_vp, c_vp]),
    "s2v_resize_bilinear": (C.c_int, [VP, VP, c_vp, c_i64, c_vp]),
    "s2v_resize_planes_f32": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_i64, c_i64, C.c_int, C.c_int, c_vp]),
    "s2v_style_demod": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, C.c_int, C.c_int, c_f32, c_f32, c_vp, c_vp]),
    "s2v_style_epilogue": (C.c_int, [VP, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_i64, VP, c_vp]),
    "s2v_to_rgb": (C.c_int, [VP, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, C.c_int, c_vp]),
    "s2v_reflect_pad_nchw_f32": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "s2v_resize_linear_u8": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_i64, c_i64, c_vp, C.c_int, C.c_int, C.c_int, c_vp]),
    "s2v_resize_linear_f32": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp]),
    "s2v_fake_to_bgr_u8": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "s2v_face_batch": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "s2v_compose_pred_u8": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "s2v_semantic_windows": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, C.c_int, c_f64, C.c_int, c_vp, c_vp]),
    "s2v_pyrdown_u8": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "s2v_pyrdown_f32": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "s2v_lap_blend_level": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "s2v_flow_warp_f32": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, VP, C.c_int, c_vp]),
    "s2v_flow_to_deformation_f32": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp]),
    "s2v_warp_deformation_f32": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_vp]),
    "s2v_pack_nchw_f32": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_i64, VP, C.c_int, C.c_int, c_f32, c_f32, c_vp]),
    "s2v_unpack_to_nchw_f32": (C.c_int, [VP, C.c_int, C.c_int, c_vp, c_vp]),
    "s2v_glue_fake_to_face_f32": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_vp]),
    "s2v_conv_simt": (C.c_int, [C.POINTER(Conv), c_vp]),
    "s2v_conv_tc": (C.c_int, [C.POINTER(Conv), C.c_int, C.c_int, C.c_int, c_vp]),
    "s2v_conv_tc_tile_n": (C.c_int, [C.c_int]),
    "s2v_conv_head": (C.c_int, [C.POINTER(Conv), c_vp]),
    "s2v_grouped_linear": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_vp, C.c_int, c_vp, c_i64, c_vp]),
    "s2v_chan_stats": (C.c_int, [VP, C.c_int, c_vp, c_vp]),
    "s2v_ln2d_finalize": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, c_i64, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp]),
    "s2v_ln2d_finalize_totals": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, c_i64, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp]),
    "s2v_adain_finalize": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, c_i64, c_vp, c_vp, c_i64, c_f32, c_vp, c_vp, c_vp]),
    "s2v_affine_act": (C.c_int, [VP, c_vp, c_vp, C.c_int, c_f32, C.c_int, VP, VP, C.c_int, c_vp]),
    "s2v_affine_act2": (C.c_int, [VP, c_vp, c_vp, C.c_int, c_f32, VP, c_vp, c_vp, VP, C.c_int, c_vp]),
    "s2v_adain_fused_fits": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "s2v_adain_fused": (C.c_int, [VP, c_vp, c_vp, c_i64, c_f32, C.c_int, c_f32, VP, VP, C.c_int, c_vp]),
    "s2v_token_layernorm": (C.c_int, [VP, c_vp, c_vp, c_f32, VP, c_vp]),
    "s2v_add": (C.c_int, [VP, VP, VP, c_vp]),
    "s2v_reflect_border": (C.c_int, [VP, c_vp]),
    "s2v_fft_init": (C.c_int, []),
    "s2v_rfft2": (C.c_int, [VP, VP, c_vp]),
    "s2v_irfft2": (C.c_int, [VP, VP, VP, c_vp]),
    "s2v_attention": (C.c_int, [VP, VP, VP, VP, C.c_int, c_f32, c_vp]),
    "s2v_mean_over_w": (C.c_int, [VP, VP, c_vp]),
    "s2v_plan_load": (C.c_int, [C.c_char_p, C.POINTER(c_vp)]),
    "s2v_plan_load_memory": (C.c_int, [c_vp, c_i64, C.POINTER(c_vp)]),
    "s2v_plan_free": (None, [c_vp]),
    "s2v_plan_const_bytes": (c_i64, [c_vp]),
    "s2v_plan_workspace_bytes": (c_i64, [c_vp]),
    "s2v_plan_num_ops": (C.c_int, [c_vp]),
    "s2v_plan_num_io": (C.c_int, [c_vp]),
    "s2v_plan_io_info": (C.c_int, [c_vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(C.c_int)]),
    "s2v_plan_bind": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "s2v_plan_run": (C.c_int, [c_vp, c_vp]),
    "s2v_lnet_forward": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "s2v_dnet_forward": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
}

EXPORTS = tuple(_SIGNATURES)


def lib_path() -> str:
    return os.environ.get("S2V_LIB", os.path.join(_HERE, "libs2v.so"))


def load_library() -> C.CDLL:
    """Loads libs2v.so and sets the prototypes.  Raises if the library is missing."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                "libs2v.so not found at %s - the CUDA extension is required (no CPU fallback); "
                "build it with `python speech-to-video-mpp_b200/build.py`" % path)
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB


class S2VError(RuntimeError):
    pass


def check(code: int, what: str = "") -> None:
    if code != 0:
        lib = load_library()
        msg = lib.s2v_strerror(int(code)).decode()
        cuda = lib.s2v_last_cuda_error().decode()
        raise S2VError("%s failed: %s (code %d; last CUDA error: %s)" % (what or "s2v call", msg, code, cuda))


_DEVICE_READY = set()


def require_device(device_index: int) -> C.CDLL:
    """Library + one-time per-device constant uploads; raises when the GPU is not sm_100."""
    import torch
    lib = load_library()
    if not torch.cuda.is_available():
        raise S2VError("a CUDA device is required: this package has no CPU path")
    if device_index not in _DEVICE_READY:
        with torch.cuda.device(device_index):
            check(lib.s2v_device_ok(), "s2v_device_ok")
            check(lib.s2v_mel_init(), "s2v_mel_init")
            check(lib.s2v_fft_init(), "s2v_fft_init")
        _DEVICE_READY.add(device_index)
    return lib
