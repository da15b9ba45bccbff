"""ctypes binding of libpapr_b200.so (the C ABI declared in include/papr_b200.h).

There is no fallback: if the library is missing or a call fails, a :class:`PaprError` is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpapr_b200.so")

_c = ctypes
_ptr, _i64, _i32, _f32 = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_float

# name -> argtypes; must list every symbol declared in include/papr_b200.h (tests/test_abi.py checks that)
SIGNATURES = {
    "papr_abi_version": [],
    "papr_check_device": [_i32],
    "papr_status_string": [_i32],
    "papr_last_cuda_error": [],
    "papr_select_topk": [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i32, _f32, _ptr, _ptr],
    "papr_select_topk_sorted": [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i32, _f32, _ptr, _ptr, _ptr],
    "papr_select_topk_grid": [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i32, _i32, _f32, _ptr, _ptr],
    "papr_select_grid_workspace_bytes": [_i64, _i64, _i32],
    "papr_select_grid_build": [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i32, _f32, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr],
    "papr_blocked_from_f32": [_ptr, _i64, _i32, _i64, _ptr, _i64, _i32, _ptr],
    "papr_blocked_to_f32": [_ptr, _i32, _ptr, _i64, _i32, _i64, _ptr],
    "papr_pack_weight": [_ptr, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _ptr, _ptr],
    "papr_linear_bf16": [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i32, _i32, _i32, _f32, _ptr],
    "papr_attn_prologue_fwd": [_ptr] * 7 + [_i64, _i64, _i32, _i32, _i32, _f32, _ptr, _i32, _ptr, _i32, _ptr, _ptr, _ptr],
    "papr_attn_prologue_bwd": [_ptr] * 5 + [_i64, _i64, _i32, _i32, _i32, _f32, _ptr, _i32, _ptr, _i32] + [_ptr] * 7,
    "papr_score_blend_fwd": [_ptr] * 7 + [_i64, _i64, _i32, _i32, _i32, _i32, _f32, _f32] + [_ptr] * 5,
    "papr_blend_bwd": [_ptr] * 7 + [_i64, _i64, _i32, _i32, _i32, _i32] + [_ptr] * 5,
    "papr_key_score_bwd": [_ptr] * 5 + [_i64, _i32, _f32] + [_ptr] * 6,
    "papr_stack_bf16": [_ptr, _i32, _ptr, _i32, _i64, _f32, _ptr],
    "papr_stack_bf16_ex": [_ptr, _i32, _ptr, _i32, _i64, _f32, _i32, _ptr],
    "papr_query_tail_fwd": [_ptr, _ptr, _f32, _f32, _i64, _ptr, _ptr, _ptr, _ptr],
    "papr_query_tail_bwd": [_ptr, _ptr, _ptr, _ptr, _i64, _ptr, _f32, _i64, _ptr, _ptr, _ptr, _ptr],
    "papr_query_prologue_fwd": [_ptr, _ptr, _ptr, _i64, _i32, _f32, _ptr, _ptr],
    "papr_query_prologue_bwd": [_ptr, _ptr, _i64, _i32, _f32, _ptr, _ptr, _ptr],
    "papr_wgrad_bf16": [_ptr, _i32, _ptr, _i32, _ptr, _i64, _i32, _i32, _i32, _i64, _ptr],
    "papr_stack_bwd_workspace_bytes": [],
    "papr_stack_bwd_fused": [_ptr, _i32, _ptr, _ptr, _i32, _i64, _i32, _ptr, _i64, _ptr],
    "papr_wgrad_bf16_ex": [_ptr, _i32, _ptr, _i32, _ptr, _i64, _i32, _i32, _i32, _i64, _i32, _ptr],
    "papr_wgrad_bias_bf16": [_ptr, _i32, _ptr, _i32, _ptr, _i64, _i32, _i32, _i32, _i64, _i32, _ptr, _ptr],
    "papr_adam_step": [_ptr, _ptr, _ptr, _i32, _i64, _ptr, _ptr, _ptr, _ptr, _i32, _f32, _ptr],
    "papr_pack_weight_batch": [_ptr, _i32, _ptr],
    "papr_knn": [_ptr, _i64, _ptr, _i64, _i32, _ptr, _ptr, _ptr],
    "papr_prune_compact": [_ptr, _ptr, _ptr, _i64, _i32, _f32, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "papr_conv_bf16": [_ptr, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _ptr, _ptr, _i32, _i32, _f32, _ptr, _i64, _i64, _ptr, _i64, _ptr, _i64, _i64, _ptr],
    "papr_conv_wgrad_bf16": [_ptr, _i64, _i32, _ptr, _i64, _i64, _i32, _i32, _i32, _ptr, _i64, _i64, _i64, _ptr],
    "papr_unet_pack_input": [_ptr, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _ptr],
    "papr_unet_unpack": [_ptr, _ptr, _i32, _ptr, _i64, _i32, _ptr],
    "papr_unet_spread": [_ptr, _ptr],
    "papr_unet_pool": [_ptr, _ptr, _i32, _ptr, _ptr, _i32, _i32, _ptr],
    "papr_unet_convt_scatter": [_ptr, _ptr, _i32, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr],
    "papr_unet_zero_border": [_ptr, _ptr, _i32, _i32, _ptr],
    "papr_unet_convt_gather": [_ptr, _ptr, _i32, _i32, _ptr, _ptr, _i32, _i32, _ptr, _ptr],
    "papr_generate_rays": [_ptr, _i64, _i32, _i32, _f32, _f32, _i32, _i32, _i32, _i32, _f32, _ptr, _ptr, _ptr],
}
_RESTYPE = {"papr_status_string": _c.c_char_p, "papr_last_cuda_error": _c.c_char_p, "papr_stack_bwd_workspace_bytes": _i64,
            "papr_select_grid_workspace_bytes": _i64}


class StackLayer(ctypes.Structure):
    """papr_stack_layer of include/papr_b200.h"""
    _fields_ = [("w_image", _ptr), ("bias", _ptr), ("out_blocked", _ptr), ("out_f32", _ptr), ("ld_f32", _i64),
                ("sign_bits_out", _ptr), ("sign_bits_in", _ptr), ("colsum", _ptr), ("N", ctypes.c_int32),
                ("act", ctypes.c_int32), ("w_replicas", ctypes.c_int32), ("_pad", ctypes.c_int32), ("w_replica_stride", _i64)]


class WgradLayer(ctypes.Structure):
    """papr_wgrad_layer of include/papr_b200.h"""
    _fields_ = [("x_blocked", _ptr), ("x_cols", ctypes.c_int32), ("gw", _ptr), ("ldw", _i64), ("n_out", ctypes.c_int32),
                ("n_in", ctypes.c_int32)]


class AdamGroup(ctypes.Structure):
    """papr_adam_group of include/papr_b200.h"""
    _fields_ = [("lr", _f32), ("beta1", _f32), ("beta2", _f32), ("eps", _f32), ("weight_decay", _f32),
                ("step", ctypes.c_int32), ("enabled", ctypes.c_int32)]


class PackDesc(ctypes.Structure):
    """papr_pack_desc of include/papr_b200.h"""
    _fields_ = [("w", _ptr), ("ld", _i64), ("rows", ctypes.c_int32), ("cols", ctypes.c_int32), ("transpose", ctypes.c_int32),
                ("N", ctypes.c_int32), ("K", ctypes.c_int32), ("replicas", ctypes.c_int32), ("scale", _f32),
                ("_pad", ctypes.c_int32), ("rep_stride", _i64), ("image", _ptr)]


class Raster(ctypes.Structure):
    """papr_raster of include/papr_b200.h"""
    _fields_ = [("H", ctypes.c_int32), ("W", ctypes.c_int32), ("Wp", ctypes.c_int32), ("_pad", ctypes.c_int32),
                ("row0", _i64), ("plane_bytes", _i64), ("copy_bytes", _i64)]


class SpreadArgs(ctypes.Structure):
    """papr_spread_args of include/papr_b200.h"""
    _fields_ = [("src", _ptr), ("add", _ptr), ("mask", _ptr), ("pool_grad", _ptr), ("pool_ref", _ptr), ("gamma", _ptr),
                ("beta", _ptr), ("dst", _ptr), ("colsum", _ptr), ("geom", Raster), ("pool_geom", Raster), ("dst_geom", Raster),
                ("src_cb0", ctypes.c_int32), ("add_cb0", ctypes.c_int32), ("mask_cb0", ctypes.c_int32),
                ("pool_ref_cb0", ctypes.c_int32), ("dst_cb0", ctypes.c_int32), ("ncopies", ctypes.c_int32),
                ("cbs", ctypes.c_int32), ("_pad", ctypes.c_int32)]


class PaprError(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded library; raises PaprError if it has not been built (python -m papr_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PaprError(f"{LIB_PATH} is missing: build it with `python -m papr_b200.build` "
                            "(there is no CPU or PyTorch fallback for the CUDA path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, _i32)
        _lib = handle
    return _lib


_checked_devices = set()


def check_device(index):
    """Raise PaprError unless CUDA device `index` is an sm_100 part (checked once per device)."""
    if index in _checked_devices:
        return
    status = lib().papr_check_device(int(index))
    if status != 0:
        raise PaprError(f"cuda:{index}: " + lib().papr_status_string(status).decode())
    _checked_devices.add(index)


def check(status, what):
    if status != 0:
        l = lib()
        msg = l.papr_status_string(status).decode()
        if status == -2:
            msg += ": " + l.papr_last_cuda_error().decode()
        raise PaprError(f"{what} failed: {msg}")
