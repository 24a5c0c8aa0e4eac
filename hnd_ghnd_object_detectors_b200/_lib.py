"""ctypes binding of libghnd_b200.so (the C ABI declared in include/ghnd_b200.h).

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked),
but every compute call goes to the CUDA library and raises if it is missing or fails.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_size_t,
                    c_void_p)

F16, BF16 = 0, 1
QSCALE_DIV, QSCALE_RECIP = 0, 1
CONV_FWD, CONV_DGRAD = 0, 1
SSE_MAX_LEVELS = 8

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# GHND_LIB_PATH: load another build of the same ABI (same-box A/B of kernel changes; scripts/gpu.sh ab)
LIB_PATH = os.environ.get("GHND_LIB_PATH") or os.path.join(_PKG_DIR, "libghnd_b200.so")


class GhndError(RuntimeError):
    pass


class SseLevel(Structure):
    _fields_ = [("teacher", c_void_p), ("student", c_void_p), ("grad", c_void_p), ("n", c_int64),
                ("factor", c_float), ("relu_mask", c_int)]


class ConvDesc(Structure):
    _fields_ = [("kind", c_int), ("N", c_int), ("H", c_int), ("W", c_int), ("C", c_int), ("K", c_int),
                ("R", c_int), ("S", c_int), ("stride", c_int), ("pad", c_int),
                ("src", c_void_p), ("src_fmt", c_int), ("weights", c_void_p), ("w_fmt", c_int),
                ("dst", c_void_p), ("dst_fmt", c_int), ("bias", c_void_p), ("residual", c_void_p),
                ("res_fmt", c_int), ("relu", c_int), ("mask", c_void_p), ("mask_fmt", c_int),
                ("accumulate", c_int), ("stats", c_void_p), ("stats_mode", c_int),
                ("mask_stats_only", c_int)]


class WgradDesc(Structure):
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("C", c_int), ("K", c_int), ("R", c_int),
                ("S", c_int), ("pad", c_int), ("x", c_void_p), ("x_fmt", c_int), ("dy", c_void_p),
                ("dy_fmt", c_int), ("dw", c_void_p)]


class PackWeightDesc(Structure):
    _fields_ = [("w_oihw", c_void_p), ("scale_o", c_void_p), ("O", c_int), ("I", c_int), ("R", c_int), ("S", c_int),
                ("transpose", c_int), ("dst", c_void_p), ("dst_fmt", c_int)]


PACK_MAX = 16  # GHND_PACK_MAX
SUMS_ZEROED = 0x100  # GHND_SUMS_ZEROED

_P = c_void_p
_I = c_int
_L = c_int64
_F = c_float
_Z = c_size_t

# name -> (restype, argtypes); mirrors include/ghnd_b200.h one to one
SIGNATURES = {
    "ghnd_last_error": (c_char_p, []),
    "ghnd_abi_version": (_I, []),
    "ghnd_device_check": (_I, []),
    "ghnd_quantize_u8_workspace_bytes": (_Z, [_L]),
    "ghnd_quantize_u8": (_I, [_P, _L, _I, _I, _P, _P, _P, _Z, _P]),
    "ghnd_quantize_u8_minmax": (_I, [_P, _L, _I, _I, _P, _I, _P, _P, _P]),
    "ghnd_dequantize_u8": (_I, [_P, _L, _P, _P, _P]),
    "ghnd_sse_workspace_bytes": (_Z, []),
    "ghnd_sse_fwd_bwd": (_I, [POINTER(SseLevel), _I, _I, _I, _P, _P, _Z, _P]),
    "ghnd_nchw_f32_to_nhwc16": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "ghnd_nhwc16_to_nchw_f32": (_I, [_P, _I, _P, _I, _I, _I, _I, _P]),
    "ghnd_stem_pack_image": (_I, [_P, _I, _I, POINTER(c_float), POINTER(c_float), _P, _I, _I, _I, _I, _P]),
    "ghnd_stem_pack_image_resized": (_I, [_P, _I, _I, _I, _I, c_float, c_float, POINTER(c_float),
                                          POINTER(c_float), _P, _I, _I, _I, _I, _P]),
    "ghnd_stem_pack_images": (_I, [POINTER(c_void_p), POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                   POINTER(c_float), _I, POINTER(c_float), POINTER(c_float), _P, _I, _I, _I, _I, _P]),
    "ghnd_conv_plan_create": (_I, [POINTER(ConvDesc), POINTER(c_void_p)]),
    "ghnd_conv_plan_run": (_I, [_P, _P]),
    "ghnd_conv_plan_run_range": (_I, [_P, _I, _I, _P]),
    "ghnd_conv_plan_destroy": (None, [_P]),
    "ghnd_conv_plan_launches": (_I, [_P]),
    "ghnd_wgrad_plan_create": (_I, [POINTER(WgradDesc), POINTER(c_void_p)]),
    "ghnd_wgrad_plan_run": (_I, [_P, _P]),
    "ghnd_wgrad_plan_destroy": (None, [_P]),
    "ghnd_pack_weight": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "ghnd_pack_weights": (_I, [POINTER(PackWeightDesc), _I, _P]),
    "ghnd_unpack_wgrad": (_I, [_P, _P, _I, _I, _I, _I, _F, _P]),
    "ghnd_conv_narrow_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "ghnd_conv_narrow_out": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "ghnd_conv_narrow_out_minmax": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _Z, _P, _I,
                                         POINTER(c_int), _P]),
    "ghnd_conv_narrow_in": (_I, [_P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "ghnd_conv_narrow_out_dgrad": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "ghnd_wgrad_narrow_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "ghnd_wgrad_narrow": (_I, [_P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "ghnd_stem_pack_weight": (_I, [_P, _P, _P, _I, _P]),
    "ghnd_stem_conv_plan_create": (_I, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _I, POINTER(c_void_p)]),
    "ghnd_stem_conv_plan_create_k": (_I, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _I, _I, POINTER(c_void_p)]),
    "ghnd_stem_conv_plan_run": (_I, [_P, _P]),
    "ghnd_stem_plan_destroy": (None, [_P]),
    "ghnd_stem_pool_plan_create": (_I, [_P, _I, _P, _I, _P, _I, POINTER(c_void_p), POINTER(c_void_p), _I, _I, _I, _I,
                                        POINTER(c_void_p)]),
    "ghnd_stem_pool_plan_run": (_I, [_P, _P]),
    "ghnd_stem_pool_plan_destroy": (None, [_P]),
    "ghnd_maxpool3x3s2": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ghnd_maxpool3x3s2_strided": (_I, [_P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ghnd_maxpool3x3s2_bwd": (_I, [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "ghnd_maxpool3x3s2_bwd_strided": (_I, [_P, _I, _I, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "ghnd_stem_wgrad_workspace_bytes": (_Z, []),
    "ghnd_stem_wgrad": (_I, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _P, _Z, _P]),
    "ghnd_stem_wgrad_plan_create": (_I, [_P, _P, _I, _P, _P, _I, _I, _I, _P, _Z, _P]),
    "ghnd_stem_wgrad_plan_run": (_I, [_P, _P]),
    "ghnd_stem_wgrad_plan_destroy": (None, [_P]),
    "ghnd_bn_stats": (_I, [_P, _I, _I, _I, _L, _I, _P, _P]),
    "ghnd_bn_stats_finalize": (_I, [_P, _I, _I, _I, _L, _I, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P]),
    "ghnd_bn_finalize": (_I, [_P, _L, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P]),
    "ghnd_bn_eval_params": (_I, [_I, _P, _P, _P, _P, _F, _P, _P]),
    "ghnd_bn_apply": (_I, [_P, _I, _P, _I, _P, _I, _L, _I, _P, _I, _P]),
    "ghnd_bn_finalize_apply": (_I, [_P, _I, _P, _I, _P, _I, _L, _I, _I, _P, _L, _P, _P, _F, _F, _P, _P, _P, _P,
                                    _P, _P]),
    "ghnd_convert16": (_I, [_P, _I, _P, _I, _L, _P]),
    "ghnd_upsample_add": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "ghnd_adaptive_avgpool_nhwc16": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    "ghnd_small_conv_f32": (_I, [_P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "ghnd_avgpool_linear": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    "ghnd_bn_bwd_reduce": (_I, [_P, _I, _P, _I, _I, _I, _L, _I, _P, _P, _I, _P, _P]),
    "ghnd_bn_bwd_apply": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _L, _I, _P, _P, _P, _I, _P, _P, _P, _P]),
    "ghnd_bn_bwd_apply_fused_sums": (_I, [_P, _I, _P, _I, _P, _I, _I, _L, _I, _P, _P, _P, _I, _P, _P, _P, _P]),
    "ghnd_comm_unique_id": (_I, [_P]),
    "ghnd_comm_init_from_unique_id": (_I, [_P, _I, _I, POINTER(c_void_p)]),
    "ghnd_comm_allreduce_flat": (_I, [_P, _P, _L, _P]),
    "ghnd_comm_broadcast_flat": (_I, [_P, _P, _L, _I, _P]),
    "ghnd_comm_destroy": (None, [_P]),
    "ghnd_adam_step": (_I, [_P, _P, _P, _P, _L, c_double, c_double, c_double, c_double, c_double, c_double, _I, _P]),
}

_lib = None


def load():
    """Load the shared library (built in-tree by __graft_entry__.build / build.py)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise GhndError("libghnd_b200.so not built: run `python -c 'import __graft_entry__ as g; "
                            "g.build()'` (no CPU fallback exists)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI and the header diverge
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().ghnd_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise GhndError("%s failed (code %d): %s" % (what or "ghnd call", rc, last_error()))


def call(name, *args):
    """Call an int-returning entry point and raise GhndError on a non-zero code."""
    check(getattr(load(), name)(*args), name)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)


def fmt_of(dtype):
    import torch
    if dtype == torch.float16:
        return F16
    if dtype == torch.bfloat16:
        return BF16
    raise GhndError("unsupported 16-bit dtype %s" % dtype)


def dtype_of(fmt):
    import torch
    return torch.float16 if fmt == F16 else torch.bfloat16
