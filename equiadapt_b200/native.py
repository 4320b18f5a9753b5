"""ctypes binding of the C ABI declared in include/equiadapt_b200.h.

The product path has no fallback: if the shared library is missing or a symbol cannot be
resolved, importing this module's `lib()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libequiadapt_b200.so")
HEADER_PATH = os.path.join(_PKG, "..", "include", "equiadapt_b200.h")

EQB_ERR_INVALID = -1
EQB_ERR_UNSUPPORTED = -2
REP_SCALAR, REP_REGULAR = 0, 1

_fp = C.c_void_p  # device pointers travel as integers
_i = C.c_int
_SIGNATURES = {
    "eqb_abi_version": (C.c_int, []),
    "eqb_last_error": (C.c_char_p, []),
    "eqb_crop_resize_aa": (C.c_int, [_fp, _fp] + [_i] * 10 + [_fp]),
    "eqb_crop_resize_aa_absmax": (C.c_int, [_fp, _fp, _fp] + [_i] * 10 + [_fp]),
    "eqb_lift_filter_orbit": (C.c_int, [_fp, _fp] + [_i] * 5 + [_fp]),
    "eqb_regular_filter_orbit": (C.c_int, [_fp, _fp] + [_i] * 5 + [_fp]),
    "eqb_gconv_stack_workspace_bytes": (C.c_int64, [_i] * 9),
    "eqb_gconv_stack_forward": (C.c_int, [_fp] + [_i] * 4 + [_fp, _fp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
                                + [_i] * 5 + [_fp, _fp, C.c_int64, _fp]),
    "eqb_gconv_stack_packed_bytes": (C.c_int64, [_i] * 6),
    "eqb_gconv_stack_pack": (C.c_int, [_fp, _fp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)] + [_i] * 6
                             + [_fp, C.c_int64, _fp]),
    "eqb_gconv_stack_run": (C.c_int, [_fp] + [_i] * 4 + [_fp, _fp] + [_i] * 5 + [_fp, _fp, C.c_int64, _fp]),
    "eqb_gconv_stack_run_scaled": (C.c_int, [_fp, _fp] + [_i] * 4 + [_fp, _fp] + [_i] * 5 + [_fp, _fp, C.c_int64, _fp]),
    "eqb_gconv_stack_run_select": (C.c_int, [_fp, _fp] + [_i] * 4 + [_fp, _fp] + [_i] * 5 + [_fp] * 6 + [_fp, C.c_int64, _fp]),
    "eqb_conv_stack_workspace_bytes": (C.c_int64, [_i] * 8),
    "eqb_conv_stack_forward": (C.c_int, [_fp] + [_i] * 4 + [C.POINTER(C.c_void_p)] * 4 + [_i] * 4
                               + [_fp, _fp, C.c_int64, _fp]),
    "eqb_debug_last_stall": (C.c_int, [C.POINTER(C.c_int)]),
    "eqb_debug_stack_trace": (C.c_int, [_fp, _i]),
    "eqb_group_pool_select": (C.c_int, [_fp, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp]),
    "eqb_warp_canonicalize": (C.c_int, [_fp, _fp, _fp] + [_i] * 6 + [_fp]),
    "eqb_warp_invert": (C.c_int, [_fp, _fp, _fp] + [_i] * 7 + [_fp]),
    "eqb_warp_adjoint": (C.c_int, [_fp, _fp, _fp] + [_i] * 7 + [_fp]),
    "eqb_warp_element_grad": (C.c_int, [_fp, _fp, _fp] + [_i] * 7 + [_fp, _fp, _fp]),
    "eqb_conv2d_forward": (C.c_int, [_fp] * 5 + [_i] * 7 + [_fp]),
    "eqb_conv2d_weight_grad": (C.c_int, [_fp] * 3 + [_i] * 6 + [_fp]),
    "eqb_conv2d_forward_scaled": (C.c_int, [_fp] * 5 + [_i] * 7 + [_fp] * 3),
    "eqb_conv2d_weight_grad_scaled": (C.c_int, [_fp] * 3 + [_i] * 6 + [_fp] * 4),
    "eqb_plane_sums": (C.c_int, [_fp, C.c_int64, C.c_int64, _fp, _fp]),
    "eqb_group_mean_backward": (C.c_int, [_fp, _fp, _i, _i, _i, C.c_int64, _fp]),
    "eqb_lift_filter_orbit_adjoint": (C.c_int, [_fp, _fp] + [_i] * 5 + [_fp]),
    "eqb_regular_filter_orbit_adjoint": (C.c_int, [_fp, _fp] + [_i] * 5 + [_fp]),
    "eqb_regular_roll_shift": (C.c_int, [_i, _i]),
    "eqb_orbit_expand": (C.c_int, [_fp, _fp] + [_i] * 8 + [_fp]),
    "eqb_orbit_rotate_nearest": (C.c_int, [_fp, _fp] + [_i] * 6 + [_fp]),
    "eqb_cosine_group_activations": (C.c_int, [_fp, _fp, _fp, _i, _i, _i, _fp]),
    "eqb_cosine_group_activations_backward": (C.c_int, [_fp] * 5 + [_i] * 3 + [_fp]),
    "eqb_gram_schmidt3": (C.c_int, [_fp, _fp, _i, _i, _fp]),
    "eqb_gram_schmidt3_backward": (C.c_int, [_fp, _fp, _fp, _i, _i, _fp]),
    "eqb_so3_apply_backward": (C.c_int, [_fp] * 5 + [_i, _i, _fp]),
    "eqb_e3_apply_backward": (C.c_int, [_fp] * 10 + [_i, _fp]),
    "eqb_e3_invert_backward": (C.c_int, [_fp] * 6 + [_i, _fp]),
    "eqb_so3_apply": (C.c_int, [_fp, _fp, _fp, _i, _i, _fp]),
    "eqb_e3_apply": (C.c_int, [_fp] * 6 + [_i, _fp]),
    "eqb_e3_invert": (C.c_int, [_fp] * 4 + [_i, _fp]),
    "eqb_prior_stats_continuous": (C.c_int, [_fp, _i, _i, _fp, _fp]),
    "eqb_warp_affine": (C.c_int, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, C.c_double, C.c_double, _fp]),
    "eqb_warp_affine_grad": (C.c_int, [_fp] * 4 + [_i] * 5 + [_fp, _fp, _fp]),
    "eqb_vnsmall_param_count": (C.c_int, []),
    "eqb_vnsmall_workspace_bytes": (C.c_int64, [_i, _i]),
    "eqb_vnsmall_forward": (C.c_int, [_fp, _i, _i, _fp, _i, C.c_float, _fp, _fp, C.c_int64, _fp]),
    "eqb_vndeepsets_param_count": (C.c_int, [_i, _i, _i]),
    "eqb_vndeepsets_workspace_bytes": (C.c_int64, [_i]),
    "eqb_vndeepsets_forward": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, _i, _fp] + [_i] * 10
                               + [_fp, _fp, _fp, C.c_int64, _fp, _fp]),
}

_lib = None


def declared_symbols():
    """Every function name declared in include/equiadapt_b200.h."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(eqb_[a-z0-9_]+)\s*\(", text)))


def lib():
    """The loaded native library (raises if it is absent: there is no CPU fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m equiadapt_b200.build` "
                "(or __graft_entry__.build()); equiadapt_b200 has no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if handle.eqb_abi_version() != 1:
            raise RuntimeError("libequiadapt_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    """Map a C-ABI return code to the exception the reference would have raised."""
    if rc == 0:
        return
    msg = lib().eqb_last_error().decode(errors="replace") or what
    if rc == EQB_ERR_INVALID:
        raise ValueError(msg)
    if rc == EQB_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"CUDA error {rc}: {msg}")
