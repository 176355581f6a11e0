"""The C-ABI library loads and exports every symbol include/sup3r_b200.h declares (no compute
calls: this runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sup3r_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(s3_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from sup3r_b200 import _cabi
    syms = declared_symbols()
    assert len(syms) >= 30
    assert os.path.exists(_cabi.lib_path()), "build the extension first (__graft_entry__.build())"
    lib = ctypes.CDLL(_cabi.lib_path())
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes table binds exactly the declared functions
    assert sorted(_cabi.SIGNATURES) == syms


def test_error_convention_without_gpu():
    """Invalid descriptors are rejected on the host with a message, no exception / exit."""
    from sup3r_b200 import _cabi
    lib = _cabi.load()
    assert lib.s3_version() >= 100
    d = _cabi.ConvDesc()
    d.ndim = 7
    rc = lib.s3_conv_out_dims(ctypes.byref(d), _cabi.c_i32x3(), _cabi.c_i32x3(),
                              ctypes.byref(ctypes.c_int32()))
    assert rc == -1 and b"ndim" in lib.s3_last_error()
    assert lib.s3_umma_npad(200) == 208 and lib.s3_umma_weight_layout(3, 64, 0) == 1
    assert lib.s3_umma_weight_layout(2, 64, 0) == 0 and lib.s3_umma_weight_layout(3, 200, 0) == 0


def test_ops_fail_loudly_on_cpu_tensors():
    import pytest
    import torch
    from sup3r_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.act_fwd(torch.zeros(4), 1)
