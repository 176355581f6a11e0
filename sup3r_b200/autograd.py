"""Differentiable wrappers: every forward AND backward below is one of our CUDA kernels
(``ops``); ``torch.autograd`` is used only as the tape that orders them, standing in for
``tf.GradientTape`` in the reference (sup3r/models/abstract.py:1230-1238).
"""
from __future__ import annotations

import dataclasses
import os

import torch

from . import ops
from ._cabi import S3_ACT_NONE, S3_PAD_ZERO


class ConvFn(torch.autograd.Function):
    """conv (+ implicit zero / reflect / symmetric pad) + bias + activation."""

    @staticmethod
    def forward(ctx, x, w, b, spec):
        y = ops.conv_fwd(x, w, b, spec)
        ctx.spec = spec
        ctx.x_shape = tuple(x.shape)
        ctx.has_bias = b is not None
        ctx.save_for_backward(x, w, y if spec.act != S3_ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        spec = ctx.spec
        dy = dy.contiguous()
        if spec.act != S3_ACT_NONE:
            dy = ops.act_bwd(y, dy, spec.act, spec.alpha)
        lin = dataclasses.replace(spec, act=S3_ACT_NONE)
        dx = dw = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = ops.conv_wgrad(x, dy, lin, w.shape, want_bias=ctx.has_bias)
        if ctx.needs_input_grad[0]:
            if spec.pad_mode == S3_PAD_ZERO:
                dx = ops.conv_dgrad(dy, w, lin, ctx.x_shape)
            else:
                # adjoint of an implicit REFLECT / SYMMETRIC pad: valid-conv dgrad on the padded
                # extent, then fold the halo back with the pad adjoint kernel
                nd = spec.ndim
                lo, hi = spec.pad_lo[3 - nd:], spec.pad_hi[3 - nd:]
                pads = [(0, 0)] + list(zip(lo, hi)) + [(0, 0)]
                pshape = tuple(s + p[0] + p[1] for s, p in zip(ctx.x_shape, pads))
                valid = dataclasses.replace(lin, pad_lo=(0, 0, 0), pad_hi=(0, 0, 0),
                                            pad_mode=S3_PAD_ZERO)
                dxp = ops.conv_dgrad(dy, w, valid, pshape)
                dx = ops.pad_bwd(dxp, ctx.x_shape, pads, spec.pad_mode)
        return dx, dw, db, None


class ConvUmmaFn(torch.autograd.Function):
    """3x3[x3] stride-1 reflect-pad-1 convolution with cin <= 64 on the tcgen05 kernels in the
    fp16c operand format (fp16 + e4m3 correction rows, ~2^-15 relative: gradients stay inside
    the 2e-3 bound of the float64 autograd test) -- forward AND input gradient.  ``w_eff`` is
    the effective correlation kernel ``(*k, cin, cout)``; ``cache`` a dict owned by the layer
    for the packed weights (keyed by the weight tensor's version).
    Stands in for the generator convolutions under ``tf.GradientTape``
    (sup3r/models/abstract.py:1131-1173, 1230-1238)."""

    @staticmethod
    def _packed(cache, key, w, ver, nd):
        hit = cache.get(key)
        if hit is None or hit[0] != ver:
            with torch.no_grad():
                wk = w.detach()
                if wk.shape[-2] < 64:     # zero rows for the padded input channels
                    wk = torch.nn.functional.pad(wk, (0, 0, 0, 64 - wk.shape[-2]))
                hit = (ver, *ops.pack_weights_umma(wk, ndim=nd, fmt=ops.S3_FMT_FP16C))
            cache[key] = hit
        return hit[1:]

    @staticmethod
    def forward(ctx, x, w, b, spec, cache):
        n, dims, cin, nd = ops.dims3(x.shape)
        ver = (w._version, w.data_ptr())
        w_hi, w_c, acc = ConvUmmaFn._packed(cache, "fwd", w, ver, nd)
        xp = x if cin == 64 else torch.nn.functional.pad(x, (0, 64 - cin))
        x_hi, x_c = ops.pack_act_pad16(xp, split=True, fmt=ops.S3_FMT_FP16C)
        sp = dataclasses.replace(spec, cin=64)
        y, _, _ = ops.conv_fwd_umma(x_hi, x_c, w_hi, w_c, b, sp, n, dims, fmt=ops.S3_FMT_FP16C,
                                    acc_scale=acc)
        ctx.spec, ctx.cache, ctx.x_shape = spec, cache, tuple(x.shape)
        ctx.has_bias = b is not None
        # weight gradient on tcgen05: 3-D, 64 output channels -> keep the fp16 padded input (the
        # forward's own operand) instead of the f32 tensor
        ctx.wgrad_umma = (nd == 3 and spec.cout == 64 and min(dims) >= 2
                          and os.environ.get("SUP3R_B200_WGRAD_FP32", "0") != "1")
        ctx.save_for_backward(x_hi if ctx.wgrad_umma else x, w,
                              y if spec.act != S3_ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        spec = ctx.spec
        dy = dy.contiguous()
        if spec.act != S3_ACT_NONE:
            dy = ops.act_bwd(y, dy, spec.act, spec.alpha)
        lin = dataclasses.replace(spec, act=S3_ACT_NONE)
        dx = dw = db = None
        nd = spec.ndim
        want_w = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        if want_w and not ctx.wgrad_umma:
            dw, db = ops.conv_wgrad(x, dy, lin, w.shape, want_bias=ctx.has_bias)
        g_hi, g_halo = None, 0
        if ctx.needs_input_grad[0]:
            pads = [(0, 0)] + [(1, 1)] * nd + [(0, 0)]
            pshape = tuple(s + p[0] + p[1] for s, p in zip(ctx.x_shape, pads))
            if spec.cout == 64 and spec.cin <= 256 and spec.cin % 16 == 0:
                # zero-padded correlation of dy with the flipped / transposed kernel on the padded
                # extent (tensor cores), then the adjoint of the reflect pad folds the halo back
                ver = (w._version, w.data_ptr())
                wt = w.detach().flip(dims=tuple(range(nd))).transpose(-1, -2).contiguous()
                w_hi, w_c, acc = ConvUmmaFn._packed(ctx.cache, "dgrad", wt, ver, nd)
                dyz = ops.pad_fwd(dy, pads, S3_PAD_ZERO)
                n, dims, _, _ = ops.dims3(dyz.shape)
                g_hi, g_c = ops.pack_act_pad16(dyz, split=True, fmt=ops.S3_FMT_FP16C,
                                               halo=S3_PAD_ZERO)
                g_halo = 2
                sp = dataclasses.replace(lin, cin=64, cout=spec.cin)
                dxp, _, _ = ops.conv_fwd_umma(g_hi, g_c, w_hi, w_c, None, sp, n, dims,
                                              fmt=ops.S3_FMT_FP16C, acc_scale=acc)
            else:
                valid = dataclasses.replace(lin, pad_lo=(0, 0, 0), pad_hi=(0, 0, 0),
                                            pad_mode=S3_PAD_ZERO)
                dxp = ops.conv_dgrad(dy, w, valid, pshape)
            dx = ops.pad_bwd(dxp, ctx.x_shape, pads, spec.pad_mode)
        if want_w and ctx.wgrad_umma:
            # dW = sum_v x_pad[v + tap] (x) dy[v] on tcgen05 (voxels = the GEMM's K dimension);
            # the zero-halo fp16 gradient tensor of the input-gradient convolution is reused
            n, dims, _, _ = ops.dims3(dy.shape)
            if g_hi is None or spec.cin > 64:
                g_hi, _ = ops.pack_act_pad16(dy, split=False, fmt=ops.S3_FMT_FP16, halo=S3_PAD_ZERO)
                g_halo = 1
            dw = ops.conv_wgrad_umma(x, g_hi, g_halo, n, dims, min(spec.cin, 64))
            if tuple(dw.shape) != tuple(w.shape):
                dw = dw.reshape(w.shape)
            db = ops.conv_bias_grad(dy, spec.cout) if ctx.has_bias else None
        return dx, dw, db, None, None


class PadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, paddings, mode):
        ctx.meta = (tuple(x.shape), paddings, mode)
        return ops.pad_fwd(x, paddings, mode)

    @staticmethod
    def backward(ctx, dy):
        shape, paddings, mode = ctx.meta
        return ops.pad_bwd(dy.contiguous(), shape, paddings, mode), None, None


class CropFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cropping):
        ctx.meta = (tuple(x.shape), cropping)
        return ops.crop_fwd(x, cropping)

    @staticmethod
    def backward(ctx, dy):
        shape, cropping = ctx.meta
        return ops.crop_bwd(dy.contiguous(), shape, cropping), None


class ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act, alpha):
        y = ops.act_fwd(x, act, alpha)
        ctx.meta = (act, alpha)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        act, alpha = ctx.meta
        return ops.act_bwd(y, dy.contiguous(), act, alpha), None, None


class AddFn(torch.autograd.Function):
    """a + b with b broadcast over leading dims (SkipConnection, Sup3rAdder)."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.b_shape = tuple(b.shape)
        ctx.same = a.numel() == b.numel()
        return ops.add(a, b)

    @staticmethod
    def backward(ctx, dy):
        db = None
        if ctx.needs_input_grad[1]:
            if not ctx.same:
                raise RuntimeError("gradient w.r.t. a broadcast addend is not supported")
            db = dy
        return dy, db


class ExpandFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, r, m, method, roll):
        ctx.meta = (tuple(x.shape), r, m, method, roll)
        return ops.expand_fwd(x, r, m, method, roll)

    @staticmethod
    def backward(ctx, dy):
        shape, r, m, method, roll = ctx.meta
        return ops.expand_bwd(dy.contiguous(), shape, r, m, method, roll), None, None, None, None


class ConcatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.meta = (a.shape[-1], b.shape[-1])
        return ops.concat_fwd(a, b)

    @staticmethod
    def backward(ctx, dy):
        ca, cb = ctx.meta
        da, db = ops.concat_bwd(dy.contiguous(), ca, cb, want_b=ctx.needs_input_grad[1])
        return da, db


class DenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act, alpha):
        y = ops.dense_fwd(x, w, b, act, alpha)
        ctx.meta = (act, alpha, b is not None)
        ctx.save_for_backward(x, w, y if act != S3_ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        act, alpha, has_b = ctx.meta
        dy = dy.contiguous()
        if act != S3_ACT_NONE:
            dy = ops.act_bwd(y, dy, act, alpha)
        dx, dw, db = ops.dense_bwd(x, w, dy, want_dx=ctx.needs_input_grad[0],
                                   want_dw=ctx.needs_input_grad[1],
                                   want_db=has_b and ctx.needs_input_grad[2])
        return dx, dw, db, None, None


class AffineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, shift):
        ctx.save_for_backward(scale)
        return ops.channel_affine(x, scale, shift)

    @staticmethod
    def backward(ctx, dy):
        (scale,) = ctx.saved_tensors
        return ops.channel_affine(dy.contiguous(), scale, None), None, None


class ContentLossFn(torch.autograd.Function):
    """keras MeanSquaredError (kind 0) / MeanAbsoluteError (kind 1) over the first c_use
    channels (base.py:478-503: exo channels sliced off; argument order (gen, true))."""

    @staticmethod
    def forward(ctx, gen, truth, c_use, kind):
        loss, dgen = ops.content_loss(gen, truth, c_use, kind, 1.0, want_grad=True)
        ctx.save_for_backward(dgen)
        return loss

    @staticmethod
    def backward(ctx, dl):
        (dgen,) = ctx.saved_tensors
        return ScaleFn.apply(dgen, dl), None, None, None


class DiscLossFn(torch.autograd.Function):
    """Relativistic average discriminator loss (base.py:505-549)."""

    @staticmethod
    def forward(ctx, out_true, out_gen):
        loss, dt, dg = ops.loss_disc(out_true, out_gen, 1.0, want_grad=True)
        ctx.save_for_backward(dt, dg)
        return loss

    @staticmethod
    def backward(ctx, dl):
        dt, dg = ctx.saved_tensors
        return ScaleFn.apply(dt, dl), ScaleFn.apply(dg, dl)


class ScaleFn(torch.autograd.Function):
    """x * s for a 0-d device scalar s (chain rule through scalar loss weights)."""

    @staticmethod
    def forward(ctx, x, s):
        shp = x.shape
        flat = x.reshape(-1, 1)
        return ops.channel_affine(flat, s.reshape(1), None).reshape(shp)

    @staticmethod
    def backward(ctx, dy):  # pragma: no cover - second order not needed
        raise RuntimeError("double backward is not supported")
