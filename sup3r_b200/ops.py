"""Tensor-level wrappers over the C ABI.  torch supplies device memory and streams only; every
kernel launched here is one of ours (``libsup3r_b200.so``).  All tensors are fp32, CUDA,
channels-last contiguous: ``(n, s1, s2, c)`` or ``(n, s1, s2, t, c)``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import torch

from . import _cabi
from ._cabi import (ConvDesc, UmmaTuning, c_i32x3, c_i32x5, S3_PAD_ZERO, S3_PAD_REFLECT,
                    S3_PAD_SYMMETRIC, S3_ACT_NONE, S3_ACT_RELU, S3_ACT_LEAKY, S3_ACT_SIGMOID,
                    S3_ACT_TANH, S3_FMT_BF16, S3_FMT_FP16, S3_FMT_FP16C)

ACT_CODES = {None: S3_ACT_NONE, "linear": S3_ACT_NONE, "relu": S3_ACT_RELU,
             "leaky_relu": S3_ACT_LEAKY, "sigmoid": S3_ACT_SIGMOID, "tanh": S3_ACT_TANH}
PAD_CODES = {"CONSTANT": S3_PAD_ZERO, "REFLECT": S3_PAD_REFLECT, "SYMMETRIC": S3_PAD_SYMMETRIC}

_launches = 0          # number of our kernels' launches requested through this module
_initialised = set()


def launch_count():
    return _launches


def _count(n=1):
    global _launches
    _launches += n


def ensure_device(t):
    """Fail loudly on anything but a CUDA tensor; initialise the library for its device."""
    if not t.is_cuda:
        raise RuntimeError("sup3r_b200 ops need CUDA tensors (there is no CPU fallback)")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _initialised:
        _cabi.call("s3_init", idx)
        _initialised.add(idx)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _s():
    """The current CUDA stream of the current device as a cudaStream_t (every launch takes it:
    the raw-handle query is ~20x cheaper than building a torch.cuda.Stream object)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t):
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.contiguous().float()
    return t


def dims3(shape):
    """(n, [z,] y, x, c) -> n, (z, y, x), c, ndim"""
    if len(shape) == 4:
        return shape[0], (1, shape[1], shape[2]), shape[3], 2
    if len(shape) == 5:
        return shape[0], (shape[1], shape[2], shape[3]), shape[4], 3
    raise RuntimeError(f"expected a 4-D or 5-D channels-last tensor, got shape {tuple(shape)}")


@dataclass
class ConvSpec:
    """Geometry + fused epilogue of one convolution (mirrors ``s3_conv_desc``)."""
    ndim: int
    cin: int
    cout: int
    ksize: tuple
    stride: tuple = (1, 1, 1)
    pad_lo: tuple = (0, 0, 0)
    pad_hi: tuple = (0, 0, 0)
    pad_mode: int = S3_PAD_ZERO
    act: int = S3_ACT_NONE
    alpha: float = 0.0
    d2s: int = 1
    d2t: int = 1
    t_roll: int = 0
    out_repeat: tuple = (1, 1, 1)
    out_cstride: int = 0
    out_coffset: int = 0
    cout_total: int = 0     # channel slice of a wider convolution (0 = whole convolution)
    cout_base: int = 0
    res_pre_act: int = 0    # f32 residual added before the activation (split-K partial sums)

    def desc(self, n, in_dims):
        d = ConvDesc()
        d.ndim, d.n = self.ndim, n
        d.in_dims = c_i32x3(*in_dims)
        d.cin, d.cout = self.cin, self.cout
        d.ksize = c_i32x3(*self.ksize)
        d.stride = c_i32x3(*self.stride)
        d.pad_lo = c_i32x3(*self.pad_lo)
        d.pad_hi = c_i32x3(*self.pad_hi)
        d.pad_mode, d.act, d.alpha = self.pad_mode, self.act, float(self.alpha)
        d.d2s, d.d2t, d.t_roll = self.d2s, self.d2t, self.t_roll
        d.out_repeat = c_i32x3(*self.out_repeat)
        d.out_cstride, d.out_coffset = self.out_cstride, self.out_coffset
        d.cout_total, d.cout_base = self.cout_total, self.cout_base
        d.res_pre_act = self.res_pre_act
        return d

    def out_dims(self, n, in_dims):
        """-> (conv_dims, out_dims, out_channels)"""
        cd, od, oc = c_i32x3(), c_i32x3(), C.c_int32()
        _cabi.call("s3_conv_out_dims", C.byref(self.desc(n, in_dims)), cd, od, C.byref(oc))
        return tuple(cd), tuple(od), oc.value


def _shape_from(n, dims, c, ndim):
    return (n, dims[1], dims[2], c) if ndim == 2 else (n, dims[0], dims[1], dims[2], c)


def pad16_shape(n, dims, c, ndim):
    pz = 1 if ndim == 3 else 0
    return (n, dims[0] + 2 * pz, dims[1] + 2, dims[2] + 2, c)


def conv_fwd(x, w, bias, spec: ConvSpec, residual=None, post_scale=None, post_shift=None,
             out=None, want_pad16=False, split=False, want_f32=True):
    """Generic fp32 direct convolution.  ``w`` is the keras kernel ``(*k, cin, cout)``.
    Returns ``y`` (f32) or ``(y, y_hi, y_lo)`` when ``want_pad16``."""
    x = _f32(x)
    ensure_device(x)
    n, dims, c, ndim = dims3(x.shape)
    assert c == spec.cin and ndim == spec.ndim, (x.shape, spec)
    _, od, oc = spec.out_dims(n, dims)
    cs = spec.out_cstride or oc
    y = out
    if y is None and want_f32:
        y = torch.empty(_shape_from(n, od, cs, ndim), device=x.device, dtype=torch.float32)
    y_hi = y_lo = None
    if want_pad16:
        y_hi = torch.empty(pad16_shape(n, od, cs, ndim), device=x.device, dtype=torch.bfloat16)
        if split:
            y_lo = torch.empty_like(y_hi)
    w, bias, residual = _f32(w), _f32(bias), _f32(residual)
    post_scale, post_shift = _f32(post_scale), _f32(post_shift)
    _cabi.call("s3_conv_fwd_f32", C.byref(spec.desc(n, dims)), _p(x), _p(w), _p(bias),
               _p(residual), _p(post_scale), _p(post_shift), _p(y), _p(y_hi), _p(y_lo), _s())
    _count()
    return (y, y_hi, y_lo) if want_pad16 else y


def small_bf16_ok(spec: ConvSpec):
    """Shapes s3_conv_fwd_small_bf16 covers (host mirror of the C side's check)."""
    return (spec.ndim == 3 and spec.cin <= 8 and spec.cout <= 8 and tuple(spec.ksize) == (3, 3, 3)
            and tuple(spec.stride) == (1, 1, 1) and tuple(spec.pad_lo) == (1, 1, 1)
            and tuple(spec.pad_hi) == (1, 1, 1) and spec.d2s == 1 and spec.d2t == 1
            and tuple(spec.out_repeat) == (1, 1, 1))


def conv_fwd_small_bf16(x, w, bias, spec: ConvSpec, residual=None, post_scale=None,
                        post_shift=None, out=None, fp16=False):
    """Narrow 3x3x3 convolution on mma.sync tensor cores (bf16 operands, or fp16 with ``fp16`` /
    an fp16 input tensor; fp32 accumulate)."""
    x16 = None
    if x.dtype == torch.float16:
        fp16 = True
    if x.dtype in (torch.bfloat16, torch.float16):
        x16, x = x.contiguous(), None
        ensure_device(x16)
        n, dims, c, ndim = dims3(x16.shape)
    else:
        x = _f32(x)
        ensure_device(x)
        n, dims, c, ndim = dims3(x.shape)
    assert c == spec.cin and ndim == spec.ndim, (c, ndim, spec)
    _, od, oc = spec.out_dims(n, dims)
    cs = spec.out_cstride or oc
    ref = x16 if x is None else x
    y = out if out is not None else torch.empty(_shape_from(n, od, cs, ndim), device=ref.device,
                                                dtype=torch.float32)
    w, bias, residual = _f32(w), _f32(bias), _f32(residual)
    post_scale, post_shift = _f32(post_scale), _f32(post_shift)
    _cabi.call("s3_conv_fwd_small_fp16" if fp16 else "s3_conv_fwd_small_bf16",
               C.byref(spec.desc(n, dims)), _p(x), _p(x16), _p(w),
               _p(bias), _p(residual), _p(post_scale), _p(post_shift), _p(y), _s())
    _count()
    return y


def conv_dgrad(dy, w, spec: ConvSpec, x_shape):
    dy = _f32(dy)
    ensure_device(dy)
    n, dims, c, ndim = dims3(x_shape)
    dx = torch.empty(tuple(x_shape), device=dy.device, dtype=torch.float32)
    w = _f32(w)
    _cabi.call("s3_conv_dgrad_f32", C.byref(spec.desc(n, dims)), _p(dy), _p(w), _p(dx), _s())
    _count()
    return dx


def conv_wgrad(x, dy, spec: ConvSpec, w_shape, want_bias=True):
    x, dy = _f32(x), _f32(dy)
    ensure_device(x)
    n, dims, c, ndim = dims3(x.shape)
    dw = torch.empty(tuple(w_shape), device=x.device, dtype=torch.float32)
    db = torch.empty(spec.cout, device=x.device, dtype=torch.float32) if want_bias else None
    desc = spec.desc(n, dims)
    need = _cabi.load().s3_conv_wgrad_scratch_bytes(C.byref(desc))
    scratch = _scratch(x.device, need) if need else None
    _cabi.call("s3_conv_wgrad_f32", C.byref(desc), _p(x), _p(dy), _p(dw), _p(db), _p(scratch),
               _s())
    _count(4 if want_bias else 2)
    return dw, db


_scratch_bufs = {}


def _scratch(device, nbytes):
    """Cached device scratch (stream-ordered reuse) for the split-K partial sums."""
    key = str(device)
    buf = _scratch_bufs.get(key)
    if buf is None or buf.numel() * 4 < nbytes:
        buf = _scratch_bufs[key] = torch.empty((nbytes + 3) // 4, device=device,
                                               dtype=torch.float32)
    return buf


def conv_bias_grad(dy, cout):
    """db[co] = sum over voxels of dy (the bias half of s3_conv_wgrad_f32)."""
    dy = _f32(dy)
    ensure_device(dy)
    n, dims, c, ndim = dims3(dy.shape)
    db = torch.empty(cout, device=dy.device, dtype=torch.float32)
    spec = ConvSpec(ndim, cout, cout, (1,) * 3)
    _cabi.call("s3_conv_wgrad_f32", C.byref(spec.desc(n, dims)), _p(dy), _p(dy), None, _p(db), None,
               _s())
    _count()
    return db


_wgrad_ws = {}


def conv_wgrad_umma(x_hi, g_hi, g_halo, n, dims, cin, scale=1.0):
    """Weight gradient (3, 3, 3, cin, 64) of a 3x3x3 stride-1 reflect-pad-1 convolution on tcgen05.
    ``x_hi``: fp16 padded (REFLECT halo) input of the forward pass, ``g_hi``: fp16 zero-halo
    padded output gradient with ``g_halo`` halo voxels per side (1, or 2 for the tensor the
    input-gradient kernel consumes); ``dims``: interior (z, y, x)."""
    ensure_device(x_hi)
    if x_hi.dtype != torch.float16 or g_hi.dtype != torch.float16:
        raise RuntimeError("conv_wgrad_umma needs fp16 padded operands")
    z, y, x = dims
    lib = _cabi.load()
    need = lib.s3_conv_wgrad_umma_ws_bytes(n, z, y, x)
    key = str(x_hi.device)
    ws = _wgrad_ws.get(key)
    if ws is None or ws.numel() * 4 < need:
        ws = _wgrad_ws[key] = torch.empty((need + 3) // 4, device=x_hi.device, dtype=torch.float32)
    dw = torch.empty((3, 3, 3, cin, 64), device=x_hi.device, dtype=torch.float32)
    _cabi.call("s3_conv_wgrad_umma", _p(x_hi), _p(g_hi), int(g_halo), n, z, y, x, int(cin),
               float(scale), _p(dw), _p(ws), ws.numel() * 4, _s())
    _count(2)
    return dw


def umma_npad(cout):
    return _cabi.load().s3_umma_npad(cout)


def _dt16(fmt):
    return torch.bfloat16 if fmt == S3_FMT_BF16 else torch.float16


def pack_weights_umma(w, split=False, fmt=0, ndim=None, wmax=None):
    """keras kernel ``(*k, 64, cout)`` f32 -> packed 16-bit (hi, lo|None) in the layout the
    tcgen05 kernel wants for this rank / cout (``s3_umma_weight_layout``).  ``fmt`` 2 (fp16c):
    -> (hi, corr, acc_scale): fp16 weights scaled by a power of two S with max|w| S in
    [2^13, 2^14), their e4m3 correction rows, and 1 / S for the kernel's epilogue."""
    w = _f32(w)
    ensure_device(w)
    cin, cout = w.shape[-2], w.shape[-1]
    taps = w.numel() // (cin * cout)
    npad = umma_npad(cout)
    if ndim is None:
        ndim = 3 if taps == 27 else 2
    if fmt == S3_FMT_FP16C:
        import math
        layout = _cabi.load().s3_umma_weight_layout(ndim, cout, 0)
        # (``wmax``: max |w| when the caller already knows it -- the reduction + host read is a
        #  device sync, which the training step pays once per network instead of once per layer)
        wmax = float(w.abs().max()) if wmax is None else float(wmax)
        scale = 2.0 ** math.floor(math.log2(16383.0 / wmax)) if wmax > 0 else 1.0
        scale = min(max(scale, 2.0 ** -24), 2.0 ** 24)
        hi = torch.empty((taps, npad, cin), device=w.device, dtype=torch.float16)
        corr = torch.empty_like(hi)
        _cabi.call("s3_pack_weights_umma_c", _p(w), taps, cin, cout, _p(hi), _p(corr),
                   float(scale), layout, _s())
        _count()
        return hi, corr, 1.0 / scale
    layout = _cabi.load().s3_umma_weight_layout(ndim, cout, 1 if split else 0)
    hi = torch.empty((taps, npad, cin), device=w.device, dtype=_dt16(fmt))
    lo = torch.empty_like(hi) if split else None
    _cabi.call("s3_pack_weights_umma", _p(w), taps, cin, cout, _p(hi), _p(lo), fmt, layout, _s())
    _count()
    return hi, lo


def pack_weights_umma_view(w, ci0, co0, cout, adjoint=False, ndim=None, wmax=None):
    """fp16c packing of a view of the keras kernel ``w (*k, cin, cout_w)`` without copies
    (``s3_pack_weights_umma_view``): 64 rows from input channel ``ci0`` and ``cout`` columns from
    output channel ``co0`` -- or, ``adjoint``, the flipped kernel with the channel roles swapped
    (rows = output channels from ``ci0``, columns = input channels from ``co0``), the operand of
    the input-gradient convolution.  Zero outside ``w``.  -> (hi, corr, acc_scale)."""
    import math
    w = _f32(w)
    ensure_device(w)
    src_cin, src_cout = w.shape[-2], w.shape[-1]
    taps = w.numel() // (src_cin * src_cout)
    if ndim is None:
        ndim = 3 if taps == 27 else 2
    npad = umma_npad(cout)
    layout = _cabi.load().s3_umma_weight_layout(ndim, cout, 0)
    wmax = float(w.abs().max()) if wmax is None else float(wmax)
    scale = 2.0 ** math.floor(math.log2(16383.0 / wmax)) if wmax > 0 else 1.0
    scale = min(max(scale, 2.0 ** -24), 2.0 ** 24)
    hi = torch.empty((taps, npad, 64), device=w.device, dtype=torch.float16)
    corr = torch.empty_like(hi)
    _cabi.call("s3_pack_weights_umma_view", _p(w), taps, src_cin, src_cout, int(ci0), int(co0),
               1 if adjoint else 0, int(cout), _p(hi), _p(corr), float(scale), layout, _s())
    _count()
    return hi, corr, 1.0 / scale


def pack_act_pad16(x, split=False, fmt=0, halo=S3_PAD_REFLECT, halo_width=1):
    """fp32 channels-last -> 16-bit operand tensor(s) with the halo written.  ``halo_width`` 2
    (zero halo only): the tensor of the extent + 2 with a one-voxel halo, all zeros outside."""
    x = _f32(x)
    ensure_device(x)
    n, dims, c, ndim = dims3(x.shape)
    grow = 2 * (halo_width - 1)
    pdims = tuple(d + grow if (ndim == 3 or i > 0) else d for i, d in enumerate(dims))
    hi = torch.empty(pad16_shape(n, pdims, c, ndim), device=x.device, dtype=_dt16(fmt))
    lo = torch.empty_like(hi) if split else None
    _cabi.call("s3_pack_act_pad16_hw", _p(x), ndim, n, c_i32x3(*dims), c, _p(hi), _p(lo), fmt,
               halo, halo_width, _s())
    _count()
    return hi, lo


def unpack_act_pad16(hi, lo, ndim, fmt=0):
    ensure_device(hi)
    n, pz = hi.shape[0], (1 if ndim == 3 else 0)
    dims = (hi.shape[1] - 2 * pz, hi.shape[2] - 2, hi.shape[3] - 2)
    c = hi.shape[4]
    x = torch.empty(_shape_from(n, dims, c, ndim), device=hi.device, dtype=torch.float32)
    _cabi.call("s3_unpack_act_pad16", _p(hi), _p(lo), ndim, n, c_i32x3(*dims), c, _p(x), fmt, _s())
    _count()
    return x


def conv_fwd_umma(x_hi, x_lo, w_hi, w_lo, bias, spec: ConvSpec, n, dims, residual=None,
                  post_scale=None, post_shift=None, out=None, want_f32=True, want_pad16=False,
                  tune=None, out_hi=None, out_lo=None, res_hi=None, res_lo=None, want_lo=False,
                  want_map16=False, fmt=None, acc_scale=0.0):
    """tcgen05 convolution on padded 16-bit activations.  ``dims`` = unpadded (z, y, x).
    ``fmt``: S3_FMT_* (default: from the dtype of ``x_hi``); ``acc_scale``: see fp16c weights."""
    ensure_device(x_hi)
    _, od, oc = spec.out_dims(n, dims)
    cs = spec.out_cstride or oc
    y = out
    if y is None and want_f32:
        y = torch.empty(_shape_from(n, od, cs, spec.ndim), device=x_hi.device, dtype=torch.float32)
    y_hi, y_lo = out_hi, out_lo
    if want_pad16 and y_hi is None:
        y_hi = torch.empty(pad16_shape(n, od, cs, spec.ndim), device=x_hi.device,
                           dtype=x_hi.dtype)
        if x_lo is not None or want_lo:
            y_lo = torch.empty_like(y_hi)
    if want_map16:   # unpadded 16-bit tensor of the mapped (depth_to_space) geometry
        y_hi = torch.empty(_shape_from(n, od, cs, spec.ndim), device=x_hi.device, dtype=x_hi.dtype)
    t = tune if tune is not None else UmmaTuning()
    if fmt is not None:
        t.fmt = fmt
    elif tune is None:
        t.fmt = S3_FMT_BF16 if x_hi.dtype == torch.bfloat16 else S3_FMT_FP16
    t.acc_scale = float(acc_scale)
    bias, residual = _f32(bias), _f32(residual)
    post_scale, post_shift = _f32(post_scale), _f32(post_shift)
    _cabi.call("s3_conv_fwd_umma", C.byref(spec.desc(n, dims)), _p(x_hi), _p(x_lo), _p(w_hi),
               _p(w_lo), _p(bias), _p(residual), _p(res_hi), _p(res_lo), _p(post_scale),
               _p(post_shift), _p(y), _p(y_hi),
               _p(y_lo), C.byref(t), _s())
    _count()
    return y, y_hi, y_lo


# --------------------------------------------------------------------------- eager layers
def _dims5(shape):
    """channels-last shape -> 5 extents (leading ones) for the pad / crop kernels"""
    s = list(shape)
    return [1] * (5 - len(s)) + s


def _pads5(paddings, rank):
    lo = [0] * (5 - rank) + [int(p[0]) for p in paddings]
    hi = [0] * (5 - rank) + [int(p[1]) for p in paddings]
    return lo, hi


def pad_fwd(x, paddings, mode):
    x = _f32(x)
    ensure_device(x)
    lo, hi = _pads5(paddings, x.dim())
    d = _dims5(x.shape)
    out_shape = tuple(s + int(p[0]) + int(p[1]) for s, p in zip(x.shape, paddings))
    y = torch.empty(out_shape, device=x.device, dtype=torch.float32)
    _cabi.call("s3_pad_fwd", _p(x), _p(y), c_i32x5(*d), c_i32x5(*lo), c_i32x5(*hi), mode, _s())
    _count()
    return y


def pad_bwd(dy, in_shape, paddings, mode):
    dy = _f32(dy)
    ensure_device(dy)
    lo, hi = _pads5(paddings, len(in_shape))
    dx = torch.empty(tuple(in_shape), device=dy.device, dtype=torch.float32)
    _cabi.call("s3_pad_bwd", _p(dy), _p(dx), c_i32x5(*_dims5(in_shape)), c_i32x5(*lo),
               c_i32x5(*hi), mode, _s())
    _count()
    return dx


def crop_fwd(x, cropping):
    """cropping: [(lo, hi)] per axis (all axes incl. batch and channel)."""
    x = _f32(x)
    ensure_device(x)
    lo, hi = _pads5(cropping, x.dim())
    out_shape = tuple(s - int(c[0]) - int(c[1]) for s, c in zip(x.shape, cropping))
    y = torch.empty(out_shape, device=x.device, dtype=torch.float32)
    _cabi.call("s3_crop_fwd", _p(x), _p(y), c_i32x5(*_dims5(x.shape)), c_i32x5(*lo), c_i32x5(*hi),
               _s())
    _count()
    return y


def crop_bwd(dy, in_shape, cropping):
    dy = _f32(dy)
    ensure_device(dy)
    lo, hi = _pads5(cropping, len(in_shape))
    dx = torch.empty(tuple(in_shape), device=dy.device, dtype=torch.float32)
    _cabi.call("s3_crop_bwd", _p(dy), _p(dx), c_i32x5(*_dims5(in_shape)), c_i32x5(*lo),
               c_i32x5(*hi), _s())
    _count()
    return dx


def act_fwd(x, act, alpha=0.0):
    x = _f32(x)
    ensure_device(x)
    y = torch.empty_like(x)
    _cabi.call("s3_act_fwd", _p(x), _p(y), x.numel(), act, float(alpha), _s())
    _count()
    return y


def act_bwd(y, dy, act, alpha=0.0):
    y, dy = _f32(y), _f32(dy)
    ensure_device(y)
    dx = torch.empty_like(y)
    _cabi.call("s3_act_bwd", _p(y), _p(dy), _p(dx), y.numel(), act, float(alpha), _s())
    _count()
    return dx


def add(a, b):
    a, b = _f32(a), _f32(b)
    ensure_device(a)
    if a.numel() % b.numel() != 0:
        raise RuntimeError(f"cannot add tensors of shapes {tuple(a.shape)} and {tuple(b.shape)}")
    y = torch.empty_like(a)
    _cabi.call("s3_add", _p(a), _p(b), _p(y), a.numel(), b.numel(), _s())
    _count()
    return y


def expand_fwd(x, spatial_mult=1, temporal_mult=1, method=0, t_roll=0):
    x = _f32(x)
    ensure_device(x)
    n, dims, c, ndim = dims3(x.shape)
    r, m = spatial_mult, temporal_mult
    cq = c // m if (m > 1 and method == 1) else c
    if (m > 1 and method == 1 and c % m) or cq % (r * r):
        raise RuntimeError(f"channels {c} not divisible by the expansion factors ({r}, {m})")
    oc = cq // (r * r)
    if ndim == 3:
        od = (dims[0] * r, dims[1] * r, dims[2] * m)
    else:
        od = (1, dims[1] * r, dims[2] * r)
    y = torch.empty(_shape_from(n, od, oc, ndim), device=x.device, dtype=torch.float32)
    _cabi.call("s3_expand_fwd", _p(x), _p(y), ndim, n, c_i32x3(*dims), c, r, m, method, t_roll,
               _s())
    _count()
    return y


def expand_bwd(dy, in_shape, spatial_mult=1, temporal_mult=1, method=0, t_roll=0):
    dy = _f32(dy)
    ensure_device(dy)
    n, dims, c, ndim = dims3(in_shape)
    dx = torch.empty(tuple(in_shape), device=dy.device, dtype=torch.float32)
    _cabi.call("s3_expand_bwd", _p(dy), _p(dx), ndim, n, c_i32x3(*dims), c, spatial_mult,
               temporal_mult, method, t_roll, _s())
    _count()
    return dx


def concat_fwd(a, b):
    a, b = _f32(a), _f32(b)
    ensure_device(a)
    if a.shape[:-1] != b.shape[:-1]:
        raise RuntimeError(f"cannot concat shapes {tuple(a.shape)} and {tuple(b.shape)}")
    ca, cb = a.shape[-1], b.shape[-1]
    y = torch.empty(a.shape[:-1] + (ca + cb,), device=a.device, dtype=torch.float32)
    _cabi.call("s3_concat_fwd", _p(a), ca, _p(b), cb, _p(y), a.numel() // ca, _s())
    _count()
    return y


def concat_bwd(dy, ca, cb, want_b=False):
    dy = _f32(dy)
    ensure_device(dy)
    da = torch.empty(dy.shape[:-1] + (ca,), device=dy.device, dtype=torch.float32)
    db = torch.empty(dy.shape[:-1] + (cb,), device=dy.device, dtype=torch.float32) \
        if want_b else None
    _cabi.call("s3_concat_bwd", _p(dy), _p(da), ca, _p(db), cb, dy.numel() // (ca + cb), _s())
    _count()
    return da, db


def channel_affine(x, scale, shift):
    x = _f32(x)
    ensure_device(x)
    c = x.shape[-1]
    y = torch.empty_like(x)
    scale, shift = _f32(scale), _f32(shift)
    _cabi.call("s3_channel_affine", _p(x), _p(y), x.numel() // c, c, _p(scale), _p(shift), _s())
    _count()
    return y


def dense_fwd(x, w, b, act=S3_ACT_NONE, alpha=0.0):
    x, w = _f32(x), _f32(w)
    ensure_device(x)
    m, k = x.shape
    assert w.shape[0] == k, (x.shape, w.shape)
    n = w.shape[1]
    y = torch.empty((m, n), device=x.device, dtype=torch.float32)
    b = _f32(b)
    _cabi.call("s3_dense_fwd", _p(x), _p(w), _p(b), _p(y), m, k, n, act, float(alpha), _s())
    _count(2)
    return y


def dense_bwd(x, w, dy, want_dx=True, want_dw=True, want_db=True):
    x, w, dy = _f32(x), _f32(w), _f32(dy)
    ensure_device(dy)
    m, k = x.shape
    n = w.shape[1]
    dx = torch.empty((m, k), device=dy.device, dtype=torch.float32) if want_dx else None
    dw = torch.empty((k, n), device=dy.device, dtype=torch.float32) if want_dw else None
    db = torch.empty((n,), device=dy.device, dtype=torch.float32) if want_db else None
    _cabi.call("s3_dense_bwd", _p(x), _p(w), _p(dy), _p(dx), _p(dw), _p(db), m, k, n, _s())
    _count(3)
    return dx, dw, db


def content_loss(gen, truth, c_use, kind, weight=1.0, want_grad=False):
    gen, truth = _f32(gen), _f32(truth)
    ensure_device(gen)
    c = gen.shape[-1]
    loss = torch.empty((), device=gen.device, dtype=torch.float32)
    dgen = torch.empty_like(gen) if want_grad else None
    scratch = torch.empty(1025, device=gen.device, dtype=torch.float32)   # S3_LOSS_SCRATCH_FLOATS
    _cabi.call("s3_content_loss", _p(gen), _p(truth), gen.numel() // c, c, c_use, kind,
               float(weight), _p(loss), _p(dgen), _p(scratch), _s())
    _count()
    return loss, dgen


def loss_disc(out_real, out_fake, weight=1.0, want_grad=False):
    out_real, out_fake = _f32(out_real), _f32(out_fake)
    ensure_device(out_real)
    b = out_real.numel()
    assert out_fake.numel() == b
    loss = torch.empty((), device=out_real.device, dtype=torch.float32)
    dr = torch.empty_like(out_real) if want_grad else None
    df = torch.empty_like(out_fake) if want_grad else None
    _cabi.call("s3_loss_disc", _p(out_real), _p(out_fake), b, float(weight), _p(loss), _p(dr),
               _p(df), _s())
    _count()
    return loss, dr, df


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step):
    ensure_device(p)
    g = _f32(g)
    _cabi.call("s3_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr),
               float(beta1), float(beta2), float(eps), int(step), _s())
    _count()


def cast_f16(x, out=None):
    """fp32 -> fp16 (saturating) copy for 16-bit result transfers."""
    x = _f32(x)
    ensure_device(x)
    y = out if out is not None else torch.empty(x.shape, device=x.device, dtype=torch.float16)
    _cabi.call("s3_cast_f16", _p(x), _p(y), x.numel(), _s())
    _count()
    return y


def stats(x):
    """-> tensor [sum, sum|x|, n_nonfinite, min, max] (device)"""
    x = _f32(x)
    ensure_device(x)
    out = torch.empty(5, device=x.device, dtype=torch.float32)
    _cabi.call("s3_stats", _p(x), x.numel(), _p(out), _s())
    _count(2)
    return out


def channel_check(x):
    """-> tensor (c, 3) = per-channel (min, max, n_nan)"""
    x = _f32(x)
    ensure_device(x)
    c = x.shape[-1]
    out = torch.empty((c, 3), device=x.device, dtype=torch.float32)
    _cabi.call("s3_channel_check", _p(x), x.numel() // c, c, _p(out), _s())
    _count(2)
    return out


# ------------------------------------------------------------------ chunk pipeline (8(f)2, 8(f)4)
def output_transform(data, cos_sin, pairs, lo, hi, clip=True):
    """In-place u/v -> (windspeed, winddirection) + limits on a device chunk (s1, s2, t, f).
    ``pairs``: [(u_idx, v_idx)]; ``lo`` / ``hi``: per-feature limits (python floats).  Returns the
    device tensor (f, 2) int64 of (below, above) counts taken before clipping."""
    ensure_device(data)
    if data.dtype != torch.float32 or not data.is_contiguous():
        raise RuntimeError("output_transform works in place on a contiguous float32 tensor")
    s1, s2, t, f = data.shape
    counts = torch.empty((f, 2), device=data.device, dtype=torch.int64)
    pu = (C.c_int * max(len(pairs), 1))(*[p[0] for p in pairs])
    pv = (C.c_int * max(len(pairs), 1))(*[p[1] for p in pairs])
    flo = (C.c_float * f)(*[float(v) for v in lo])
    fhi = (C.c_float * f)(*[float(v) for v in hi])
    _cabi.call("s3_output_transform", _p(data), s1 * s2, t, f, _p(cos_sin), pu, pv, len(pairs),
               flo, fhi, int(bool(clip)), _p(counts), _s())
    _count()
    return counts


COARSEN_METHODS = {"subsample": 0, "average": 1, "total": 2, "max": 3, "min": 4}


def coarsen(hr, s_enhance, t_enhance=1, method="subsample"):
    """(n, s1, s2, [t,] f) hi-res samples -> low-res (block mean + temporal method)."""
    hr = _f32(hr)
    ensure_device(hr)
    if method not in COARSEN_METHODS:
        raise ValueError(f"Did not recognize temporal coarsening method \"{method}\"")
    five = hr.dim() == 5
    n, s1, s2 = hr.shape[:3]
    t = hr.shape[3] if five else 1
    te = t_enhance if five else 1
    f = hr.shape[-1]
    shp = (n, s1 // s_enhance, s2 // s_enhance) + ((t // te,) if five else ()) + (f,)
    lr = torch.empty(shp, device=hr.device, dtype=torch.float32)
    _cabi.call("s3_coarsen", _p(hr), _p(lr), n, s1, s2, t, f, int(s_enhance), int(te),
               COARSEN_METHODS[method], _s())
    _count()
    return lr


def gauss_smooth2d(x, sigma, feature_mask, truncate=4.0):
    """scipy.ndimage.gaussian_filter(sigma, mode='nearest') over (s1, s2) of the masked features
    of x (n, s1, s2, [t,] f)."""
    import numpy as np
    x = _f32(x)
    ensure_device(x)
    radius = int(truncate * float(sigma) + 0.5)
    k = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (float(sigma) ** 2) * k ** 2)
    w = (w / w.sum()).astype(np.float32)
    wd = torch.from_numpy(w).to(x.device)
    n, s1, s2, f = x.shape[0], x.shape[1], x.shape[2], x.shape[-1]
    tf = x.numel() // (n * s1 * s2)
    tmp, y = torch.empty_like(x), torch.empty_like(x)
    _cabi.call("s3_gauss_smooth2d", _p(x), _p(tmp), _p(y), n, s1, s2, tf, f, int(feature_mask),
               _p(wd), radius, _s())
    _count(2)
    return y


def gather_samples(data, origins, sample_shape):
    """data (S1, S2, T, F) on the device, origins (n, 3) int32 device tensor -> (n, s1, s2, t, F)"""
    ensure_device(data)
    S1, S2, T, F = data.shape
    n = origins.shape[0]
    s1, s2, t = sample_shape
    out = torch.empty((n, s1, s2, t, F), device=data.device, dtype=torch.float32)
    _cabi.call("s3_gather_samples", _p(data), S1, S2, T, F, _p(origins), n, s1, s2, t, _p(out),
               _s())
    _count()
    return out


def qdm_bc(data, window, params_oh, params_mh, params_mf, quantiles, relative=True,
           delta_denom_zero=None, delta_denom_min=None, delta_range=None, out_range=None,
           tau_fut=None, k_factor=None):
    """Empirical quantile delta mapping on the device (``s3_qdm_bc``).  data (sites, times) f32,
    window (times,) int32, params_* (sites, windows, n_q) f32, quantiles (n_q,) f64, optional
    PresRat tau_fut (sites,) f32 / k_factor (sites, windows) f64 -- all device tensors.  Returns
    (corrected (sites, times) f32, device counters [non-finite results, NaN results])."""
    import ctypes as C
    ensure_device(data)
    n_sites, n_times = data.shape
    n_win, n_q = params_oh.shape[1], params_oh.shape[2]
    assert params_mh.shape == params_oh.shape == params_mf.shape and window.numel() == n_times
    assert quantiles.dtype == torch.float64 and quantiles.numel() == n_q
    assert window.dtype == torch.int32 and data.dtype == torch.float32

    def opt(vals):
        if vals is None:
            return None
        vals = list(vals) if isinstance(vals, (tuple, list)) else [vals]
        return (C.c_double * len(vals))(*[float(v) for v in vals])
    out = torch.empty_like(data)
    bad = torch.zeros(2, device=data.device, dtype=torch.int64)
    if tau_fut is not None:
        assert tau_fut.numel() == n_sites and tuple(k_factor.shape) == (n_sites, n_win)
        assert k_factor.dtype == torch.float64 and tau_fut.dtype == torch.float32
        tau_fut, k_factor = tau_fut.contiguous(), k_factor.contiguous()
    _cabi.call("s3_qdm_bc", _p(data), _p(window), _p(params_oh.contiguous()),
               _p(params_mh.contiguous()), _p(params_mf.contiguous()), _p(quantiles),
               _p(tau_fut), _p(k_factor), n_sites,
               n_times, n_win, n_q, 1 if relative else 0, opt(delta_denom_zero),
               opt(delta_denom_min), opt(delta_range), opt(out_range), _p(out), _p(bad), _s())
    _count()
    return out, bad

