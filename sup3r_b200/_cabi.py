"""ctypes binding of ``libsup3r_b200.so`` (the C ABI declared in ``include/sup3r_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C sup3r_b200/csrc`` into
``sup3r_b200/lib/``.  There is NO fallback: if the shared object is missing, or a call returns
a non-zero status, a Python exception is raised (``RuntimeError`` for invalid arguments / CUDA
failures, mirroring how the reference surfaces layer failures, abstract.py:1093-1098).
"""
from __future__ import annotations

import ctypes as C
import os

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libsup3r_b200.so")

S3_PAD_ZERO, S3_PAD_REFLECT, S3_PAD_SYMMETRIC = 0, 1, 2
S3_ACT_NONE, S3_ACT_RELU, S3_ACT_LEAKY, S3_ACT_SIGMOID, S3_ACT_TANH = 0, 1, 2, 3, 4
S3_FMT_BF16, S3_FMT_FP16, S3_FMT_FP16C = 0, 1, 2

c_i32x3 = C.c_int32 * 3
c_i32x5 = C.c_int32 * 5


class ConvDesc(C.Structure):
    """``s3_conv_desc``"""
    _fields_ = [
        ("ndim", C.c_int32), ("n", C.c_int32), ("in_dims", c_i32x3),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("ksize", c_i32x3), ("stride", c_i32x3), ("pad_lo", c_i32x3), ("pad_hi", c_i32x3),
        ("pad_mode", C.c_int32), ("act", C.c_int32), ("alpha", C.c_float),
        ("d2s", C.c_int32), ("d2t", C.c_int32), ("t_roll", C.c_int32),
        ("out_repeat", c_i32x3), ("out_cstride", C.c_int32), ("out_coffset", C.c_int32),
        ("cout_total", C.c_int32), ("cout_base", C.c_int32), ("res_pre_act", C.c_int32),
    ]


class UmmaTuning(C.Structure):
    """``s3_umma_tuning``"""
    _fields_ = [("tiles", C.c_int32), ("w_stages", C.c_int32), ("box_x", C.c_int32),
                ("box_y", C.c_int32), ("max_ctas", C.c_int32), ("fmt", C.c_int32), ("trace", C.c_void_p),
                ("scheme", C.c_int32), ("ring_slots", C.c_int32), ("acc_scale", C.c_float)]


_P = C.c_void_p
_SZ = C.c_size_t
_I = C.c_int
_F = C.c_float

# name -> (restype, argtypes); every symbol include/sup3r_b200.h declares
SIGNATURES = {
    "s3_init": (_I, [_I]),
    "s3_last_error": (C.c_char_p, []),
    "s3_version": (_I, []),
    "s3_sm_count": (_I, [_I]),
    "s3_conv_out_dims": (_I, [C.POINTER(ConvDesc), c_i32x3, c_i32x3, C.POINTER(C.c_int32)]),
    "s3_conv_fwd_f32": (_I, [C.POINTER(ConvDesc)] + [_P] * 10),
    "s3_conv_fwd_small_bf16": (_I, [C.POINTER(ConvDesc)] + [_P] * 9),
    "s3_conv_fwd_small_fp16": (_I, [C.POINTER(ConvDesc)] + [_P] * 9),
    "s3_conv_dgrad_f32": (_I, [C.POINTER(ConvDesc), _P, _P, _P, _P]),
    "s3_conv_wgrad_scratch_bytes": (_SZ, [C.POINTER(ConvDesc)]),
    "s3_conv_wgrad_f32": (_I, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P]),
    "s3_conv_wgrad_umma_ws_bytes": (_SZ, [_I, _I, _I, _I]),
    "s3_conv_wgrad_umma": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _SZ, _P]),
    "s3_conv_fwd_umma": (_I, [C.POINTER(ConvDesc)] + [_P] * 13 + [C.POINTER(UmmaTuning), _P]),
    "s3_umma_npad": (_I, [_I]),
    "s3_umma_weight_layout": (_I, [_I, _I, _I]),
    "s3_pack_weights_umma": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P]),
    "s3_pack_weights_umma_c": (_I, [_P, _I, _I, _I, _P, _P, _F, _I, _P]),
    "s3_pack_weights_umma_view": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _F, _I, _P]),
    "s3_pack_act_pad16": (_I, [_P, _I, _I, c_i32x3, _I, _P, _P, _I, _P]),
    "s3_pack_act_pad16_ex": (_I, [_P, _I, _I, c_i32x3, _I, _P, _P, _I, _I, _P]),
    "s3_pack_act_pad16_hw": (_I, [_P, _I, _I, c_i32x3, _I, _P, _P, _I, _I, _I, _P]),
    "s3_unpack_act_pad16": (_I, [_P, _P, _I, _I, c_i32x3, _I, _P, _I, _P]),
    "s3_pad_fwd": (_I, [_P, _P, c_i32x5, c_i32x5, c_i32x5, _I, _P]),
    "s3_pad_bwd": (_I, [_P, _P, c_i32x5, c_i32x5, c_i32x5, _I, _P]),
    "s3_crop_fwd": (_I, [_P, _P, c_i32x5, c_i32x5, c_i32x5, _P]),
    "s3_crop_bwd": (_I, [_P, _P, c_i32x5, c_i32x5, c_i32x5, _P]),
    "s3_act_fwd": (_I, [_P, _P, _SZ, _I, _F, _P]),
    "s3_act_bwd": (_I, [_P, _P, _P, _SZ, _I, _F, _P]),
    "s3_add": (_I, [_P, _P, _P, _SZ, _SZ, _P]),
    "s3_expand_fwd": (_I, [_P, _P, _I, _I, c_i32x3, _I, _I, _I, _I, _I, _P]),
    "s3_expand_bwd": (_I, [_P, _P, _I, _I, c_i32x3, _I, _I, _I, _I, _I, _P]),
    "s3_concat_fwd": (_I, [_P, _I, _P, _I, _P, _SZ, _P]),
    "s3_concat_bwd": (_I, [_P, _P, _I, _P, _I, _SZ, _P]),
    "s3_channel_affine": (_I, [_P, _P, _SZ, _I, _P, _P, _P]),
    "s3_dense_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "s3_dense_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "s3_content_loss": (_I, [_P, _P, _SZ, _I, _I, _I, _F, _P, _P, _P, _P]),
    "s3_loss_disc": (_I, [_P, _P, _I, _F, _P, _P, _P, _P]),
    "s3_adam_step": (_I, [_P, _P, _P, _P, _SZ, _F, _F, _F, _F, C.c_int64, _P]),
    "s3_cast_f16": (_I, [_P, _P, _SZ, _P]),
    "s3_stats": (_I, [_P, _SZ, _P, _P]),
    "s3_channel_check": (_I, [_P, _SZ, _I, _P, _P]),
    "s3_output_transform": (_I, [_P, _SZ, _I, _I, _P, _P, _P, _I, _P, _P, _I, _P, _P]),
    "s3_coarsen": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "s3_gauss_smooth2d": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, C.c_uint32, _P, _I, _P]),
    "s3_gather_samples": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P, _P]),
    "s3_qdm_bc": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "s3_peer_sum_f32": (_I, [_P, _I, _SZ, _P, _P]),
    "s3_peer_sum_adam": (_I, [_P, _I, _P, _I, C.c_uint64, _F, _F, _F, _F, C.c_int64, _P]),
}

_lib = None


class Sup3rB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def lib_path():
    return _LIB_PATH


def load():
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C sup3r_b200/csrc). "
            "sup3r_b200 has no CPU fallback.")
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().s3_last_error().decode("utf-8", "replace")
        raise Sup3rB200Error(f"{what}: {msg} (status {rc})")


def call(name, *args):
    """Call ``name`` and raise on a non-zero status."""
    rc = getattr(_lib or load(), name)(*args)
    if rc != 0:
        check(rc, name)
