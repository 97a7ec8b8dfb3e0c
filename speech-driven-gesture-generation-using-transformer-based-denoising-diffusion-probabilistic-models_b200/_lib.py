"""ctypes binding of the C-ABI in include/gd_b200.h (libgd_b200.so, built in-tree by csrc/build.py).

There is deliberately no fallback: if the shared library is missing or a call fails, this module
raises. PyTorch is only used by callers for device memory and streams; the signatures here are raw
pointers and sizes.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GD_LIB points at another build of the same ABI (A/B runs of compile-time variants, profiles/chain_ab.py)
LIB_PATH = os.environ.get("GD_LIB") or os.path.join(_HERE, "libgd_b200.so")

c_i32, c_f32, c_vp = C.c_int32, C.c_float, C.c_void_p


class LinearDesc(C.Structure):
    _fields_ = [("A", c_vp), ("W", c_vp), ("M", c_i32), ("N", c_i32), ("K", c_i32), ("lda", c_i32), ("ldw", c_i32),
                ("bias", c_vp), ("rowbias", c_vp), ("rowbias_period", c_i32), ("rowbias_offset", c_i32),
                ("residual", c_vp), ("ldr", c_i32), ("act", c_i32), ("out_f32", c_vp), ("ldo_f32", c_i32),
                ("out_bf16", c_vp), ("ldo_bf16", c_i32), ("max_ctas", c_i32)]


class DdpmDesc(C.Structure):
    _fields_ = [("x", c_vp), ("noise_tape", c_vp), ("coef_A", c_vp), ("coef_B", c_vp), ("coef_C1", c_vp),
                ("coef_C2", c_vp), ("sigma", c_vp), ("step_ptr", c_vp), ("n_clips", c_i32), ("C", c_i32),
                ("T", c_i32), ("eps_out", c_vp), ("x0_out", c_vp), ("xa_bf16", c_vp), ("ld_xa", c_i32),
                ("inpaint_seed", c_vp), ("inpaint_mask", c_vp), ("inpaint_factor", c_vp), ("clip_x0", c_f32),
                ("xa_add", c_vp), ("mean_out", c_vp), ("raw_x0_out", c_vp), ("aux_step_ptr", c_vp)]


class LnDesc(C.Structure):
    _fields_ = [("gamma", c_vp), ("beta", c_vp), ("gamma2", c_vp), ("beta2", c_vp), ("split_row", c_i32),
                ("out_bf16", c_vp), ("ldo", c_i32), ("eps", c_f32)]


class AttnDesc(C.Structure):
    _fields_ = [("q", c_vp * 2), ("q_rows", c_i32 * 2), ("q_ld", c_i32 * 2), ("k", c_vp * 2), ("v", c_vp * 2),
                ("kv_rows", c_i32 * 2), ("kv_ld", c_i32 * 2), ("out", c_vp * 2), ("out_ld", c_i32 * 2),
                ("conv_wq", c_vp), ("conv_bq", c_vp), ("conv_wk", c_vp), ("conv_bk", c_vp), ("conv_wv", c_vp),
                ("conv_bv", c_vp), ("n_clips", c_i32), ("heads", c_i32), ("d_k", c_i32), ("scale", c_f32),
                ("q_clip_stride", c_i32 * 2), ("max_ctas_sms", c_i32)]


class ConvDesc(C.Structure):
    _fields_ = [("inp", c_vp), ("W", c_vp), ("n_images", c_i32), ("grid_h", c_i32), ("grid_w", c_i32),
                ("in_ld", c_i32), ("k_per_tap", c_i32), ("c_out", c_i32), ("n_taps", c_i32), ("tap_shift", c_i32 * 9),
                ("bias", c_vp), ("scale", c_vp), ("shift", c_vp), ("relu", c_i32), ("y0", c_i32), ("y1", c_i32),
                ("x0", c_i32), ("x1", c_i32), ("stride", c_i32), ("out", c_vp), ("out_ld", c_i32),
                ("out_img_stride", c_i32), ("out_y_stride", c_i32), ("out_x_stride", c_i32), ("out_offset", c_i32),
                ("c_store", c_i32), ("split_out", c_i32), ("walk", c_i32)]


ACT_NONE, ACT_RELU2, ACT_SILU = 0, 1, 2
ABI_VERSION = 4  # include/gd_b200.h GD_ABI_VERSION

# name -> (restype, argtypes); every symbol include/gd_b200.h declares
SYMBOLS = {
    "gd_abi_version": (c_i32, []),
    "gd_last_error": (C.c_char_p, []),
    "gd_launch_count": (C.c_uint64, []),
    "gd_linear_bf16": (c_i32, [C.POINTER(LinearDesc), c_vp]),
    "gd_ddpm_update": (c_i32, [C.POINTER(DdpmDesc), c_vp, c_vp]),
    "gd_linear_ddpm": (c_i32, [C.POINTER(LinearDesc), C.POINTER(DdpmDesc), c_vp]),
    "gd_linear_resid_ln": (c_i32, [C.POINTER(LinearDesc), C.POINTER(LnDesc), c_vp]),
    "gd_linear_ln_bf16": (c_i32, [C.POINTER(LinearDesc), c_vp, c_vp, c_f32, c_vp]),
    "gd_layernorm": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_f32, c_vp]),
    "gd_layernorm_split": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_f32, c_vp]),
    "gd_dconv_attention": (c_i32, [C.POINTER(AttnDesc), c_vp]),
    "gd_dconv_attention_f32in": (c_i32, [C.POINTER(AttnDesc), c_vp]),
    "gd_scatter_step_row_f32": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "gd_scatter_step_row_bf16": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "gd_pack_pose_rows": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "gd_pack_pose_rows_add": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "gd_cast_rows_bf16": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "gd_step_add": (c_i32, [c_vp, c_i32, c_vp]),
    "gd_conv_taps_bf16": (c_i32, [C.POINTER(ConvDesc), c_vp]),
    "gd_mel_power": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_f32, c_f32, c_vp, c_vp]),
    "gd_instance_norm_rows": (c_i32, [c_vp, c_i32, c_i32, c_f32, c_vp]),
    "gd_speech_stem": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "gd_se_gate_scratch_bytes": (C.c_int64, [c_i32, c_i32, c_i32, c_i32]),
    "gd_se_gate": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                           C.c_int64, c_vp]),
    "gd_se_residual_relu": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "gd_pixel_shuffle_rows": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
}


class GdError(RuntimeError):
    pass


_lib = None


def load():
    """Load libgd_b200.so; raises (never falls back) if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GdError(f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
                      "(nvcc, sm_100a). There is no CPU/PyTorch fallback for the sampling kernels.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.gd_abi_version() != ABI_VERSION:
        raise GdError("libgd_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().gd_last_error().decode("utf-8", "replace")
        raise GdError(f"{what} failed ({rc}): {msg}")


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
