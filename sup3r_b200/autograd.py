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
    def forward(ctx, x, w, b, spec, tc_wgrad=False):
        y = ops.conv_fwd(x, w, b, spec)
        ctx.spec = spec
        ctx.x_shape = tuple(x.shape)
        ctx.has_bias = b is not None
        # narrow 3x3x3 stride-1 pad-1 layers (e.g. the hi-res output convolution) keep the fp32
        # forward / input-gradient kernels, but their weight gradient -- a reduction over every
        # voxel -- runs on tcgen05 with the channels zero-padded to 64
        from ._cabi import S3_PAD_REFLECT
        ctx.tc_wgrad = bool(
            tc_wgrad and spec.ndim == 3 and tuple(spec.ksize) == (3, 3, 3)
            and tuple(spec.stride) == (1, 1, 1) and tuple(spec.pad_lo) == (1, 1, 1)
            and tuple(spec.pad_hi) == (1, 1, 1) and spec.pad_mode in (S3_PAD_REFLECT, S3_PAD_ZERO)
            and spec.cin <= 64 and spec.cout <= 64 and spec.d2s == 1 and spec.d2t == 1
            and min(x.shape[1:-1]) >= 2 and x.numel() // x.shape[-1] >= 4096)
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
            if ctx.tc_wgrad:
                n, dims, cin, _ = ops.dims3(x.shape)
                pad64 = torch.nn.functional.pad
                x_hi, _ = ops.pack_act_pad16(pad64(x, (0, 64 - cin)) if cin < 64 else x,
                                             split=False, fmt=ops.S3_FMT_FP16, halo=spec.pad_mode)
                g_hi, _ = ops.pack_act_pad16(
                    pad64(dy, (0, 64 - spec.cout)) if spec.cout < 64 else dy, split=False,
                    fmt=ops.S3_FMT_FP16, halo=S3_PAD_ZERO)
                dw = ops.conv_wgrad_umma(x_hi, g_hi, 1, n, dims, cin)
                if spec.cout < 64:
                    dw = dw[..., :spec.cout].contiguous()
                dw = dw.reshape(w.shape)
                db = ops.conv_bias_grad(dy, spec.cout) if ctx.has_bias else None
            else:
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
        return dx, dw, db, None, None


class ConvUmmaFn(torch.autograd.Function):
    """3x3[x3] stride-1 pad-1 convolution on the tcgen05 kernels -- forward, input gradient AND
    weight gradient.  ``halo`` selects the padding the fp16 operand tensors carry: REFLECT (the
    generator's FlexiblePadding -> Conv -> Cropping runs) or ZERO (keras ``padding='same'``, the
    discriminator).  Channels: cin <= 64 (zero-padded to 64) or a multiple of 64 -- the kernels
    contract 64 input channels per launch, wider inputs are summed over 64-channel groups through
    the f32 residual input; cout <= 256 per launch, wider outputs in slices.  Forward / input
    gradient use the fp16c operand format (fp16 + e4m3 correction rows, ~2^-15 relative); the
    weight gradient fp16 operands with the voxels as the GEMM's K dimension (its rounding errors
    average out over K).  ``w`` is the effective correlation kernel ``(*k, cin, cout)``;
    ``cache`` a dict owned by the layer for the packed weights (keyed by the weight version).
    Stands in for the convolutions under ``tf.GradientTape``
    (sup3r/models/abstract.py:1131-1173, 1230-1238; base.py:283-313)."""

    @staticmethod
    def _pack_w(cache, key, ver, w, ci0, co0, cout, adjoint, nd):
        """Packed fp16c operand of a 64-row view of ``w`` (see ``ops.pack_weights_umma_view``:
        slices, zero padding, the flip / transpose of the input-gradient kernel all happen inside
        the packing kernel), cached per weight version."""
        hit = cache.get(key)
        if hit is None or hit[0] != ver:
            # cache["wmax"]: max |w| of the whole kernel for this weight version (set by
            # Plan.forward_train for all layers with ONE host read per step); every view shares
            # it (its own maximum can only be smaller)
            wmax = cache.get("wmax")
            with torch.no_grad():
                hit = (ver, *ops.pack_weights_umma_view(
                    w.detach(), ci0, co0, cout, adjoint=adjoint, ndim=nd,
                    wmax=wmax[1] if (wmax is not None and wmax[0] == ver[0]) else None))
            cache[key] = hit
        return hit[1:]

    @staticmethod
    def _kernel_spec(spec, nd, **kw):
        """The tcgen05 kernel's view: 64 input channels, pad 1 per convolved dim, halo already in
        the operand tensor (pad_mode REFLECT is the kernel's name for "use the stored halo")."""
        from ._cabi import S3_PAD_REFLECT
        z = 3 - nd
        one = (0,) * z + (1,) * nd
        return dataclasses.replace(spec, cin=64, stride=(1, 1, 1), pad_lo=one, pad_hi=one,
                                   pad_mode=S3_PAD_REFLECT, **kw)

    @staticmethod
    def _groups(c):
        return 1 if c <= 64 else c // 64

    @staticmethod
    def supported(cin, cout):
        return (cin <= 64 or cin % 64 == 0) and (cout <= 256 or cout % 64 == 0)

    @staticmethod
    def forward(ctx, x, w, b, spec, cache, halo=None, wver=None):
        """``wver``: version stamp of the weight (``Variable.version``): the optimiser kernel
        updates weights through raw pointers, which torch's own ``_version`` does not see."""
        from ._cabi import S3_PAD_REFLECT
        halo = S3_PAD_REFLECT if halo is None else halo
        n, dims, cin, nd = ops.dims3(x.shape)
        cout = spec.cout
        ver = (wver, w._version, w.data_ptr())
        ctx.ver = ver
        gin = ConvUmmaFn._groups(cin)
        xs = []
        for gi in range(gin):
            xg = x if gin == 1 else x[..., 64 * gi:64 * gi + 64].contiguous()
            if xg.shape[-1] < 64:
                xg = torch.nn.functional.pad(xg, (0, 64 - xg.shape[-1]))
            xs.append(ops.pack_act_pad16(xg, split=True, fmt=ops.S3_FMT_FP16C, halo=halo))
        fuse_act = gin == 1
        outs = []
        for c0 in range(0, cout, 256):
            c1 = min(cout, c0 + 256)
            y = None
            for gi in range(gin):
                w_hi, w_c, acc = ConvUmmaFn._pack_w(cache, ("fwd", gi, c0), ver, w, 64 * gi, c0,
                                                    c1 - c0, False, nd)
                sp = ConvUmmaFn._kernel_spec(spec, nd, cout=c1 - c0)
                if not fuse_act:
                    sp = dataclasses.replace(sp, act=S3_ACT_NONE, alpha=0.0)
                bias = b[c0:c1].contiguous() if (b is not None and gi == 0) else None
                if b is not None and gi == 0 and c0 == 0 and c1 == cout:
                    bias = b
                y, _, _ = ops.conv_fwd_umma(xs[gi][0], xs[gi][1], w_hi, w_c, bias, sp, n, dims,
                                            residual=y, fmt=ops.S3_FMT_FP16C, acc_scale=acc)
            outs.append(y)
        y = outs[0] if len(outs) == 1 else torch.cat(outs, dim=-1)
        if not fuse_act and spec.act != S3_ACT_NONE:
            y = ops.act_fwd(y, spec.act, spec.alpha)
        ctx.spec, ctx.cache, ctx.x_shape, ctx.halo = spec, cache, tuple(x.shape), halo
        ctx.has_bias = b is not None
        # weight gradient on tcgen05: 3-D, output channels in blocks of 64 -> keep the fp16 padded
        # inputs (the forward's own operands) instead of the f32 tensor
        ctx.wgrad_umma = (nd == 3 and min(dims) >= 2
                          and os.environ.get("SUP3R_B200_WGRAD_FP32", "0") != "1")
        ctx.gin = gin
        keep = [t[0] for t in xs] if ctx.wgrad_umma else [x]
        ctx.save_for_backward(w, y if spec.act != S3_ACT_NONE else None, *keep)
        return y

    @staticmethod
    def backward(ctx, dy):
        from ._cabi import S3_PAD_REFLECT
        w, y, *xs = ctx.saved_tensors
        spec, halo, cache = ctx.spec, ctx.halo, ctx.cache
        dy = dy.contiguous()
        if spec.act != S3_ACT_NONE:
            dy = ops.act_bwd(y, dy, spec.act, spec.alpha)
        lin = dataclasses.replace(spec, act=S3_ACT_NONE, alpha=0.0)
        dx = dw = db = None
        nd, cin, cout = spec.ndim, spec.cin, spec.cout
        reflect = halo == S3_PAD_REFLECT
        fp32 = lin if reflect else dataclasses.replace(lin, pad_mode=S3_PAD_ZERO)
        want_w = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        if want_w and not ctx.wgrad_umma:
            dw, db = ops.conv_wgrad(xs[0], dy, fp32, w.shape, want_bias=ctx.has_bias)
        n, dims, _, _ = ops.dims3(dy.shape)
        # output channels in blocks of 64 (a ragged last block is zero-padded)
        gout = (cout + 63) // 64
        dy_p = dy if cout % 64 == 0 else torch.nn.functional.pad(dy, (0, 64 * gout - cout))
        g_parts, g_halo = None, 0       # fp16 zero-halo gradient tensors per 64-channel block
        pads = [(0, 0)] + [(1, 1)] * nd + [(0, 0)]
        if ctx.needs_input_grad[0]:
            if gout:
                # correlation of dy with the flipped / transposed kernel on tensor cores, summed
                # over the 64-channel blocks of dy.  REFLECT: on the padded extent, then the
                # adjoint of the reflect pad folds the halo back; ZERO ('same'): directly.
                ver = ctx.ver
                # (REFLECT: the gradient on the padded extent is dy zero-padded by one voxel; the
                #  operand packing writes that border itself -- a zero halo two voxels wide)
                g_parts, g_halo = [], (2 if reflect else 1)
                gn, gdims, _, _ = ops.dims3(dy_p.shape)
                if reflect:
                    gdims = tuple(d + 2 if (nd == 3 or i > 0) else d for i, d in enumerate(gdims))
                for go in range(gout):
                    blk = dy_p if gout == 1 else dy_p[..., 64 * go:64 * go + 64].contiguous()
                    g_parts.append(ops.pack_act_pad16(blk, split=True, fmt=ops.S3_FMT_FP16C,
                                                      halo=S3_PAD_ZERO, halo_width=g_halo))
                cin_p = (cin + 15) // 16 * 16
                slices = []
                for c0 in range(0, cin_p, 256):
                    c1 = min(cin_p, c0 + 256)
                    dxp = None
                    for go in range(gout):
                        # (flipped taps, rows = output channels of block go, columns = input
                        #  channels c0 .. c1 of w: a view packed without copies)
                        w_hi, w_c, acc = ConvUmmaFn._pack_w(cache, ("dgrad", go, c0), ver, w,
                                                            64 * go, c0, c1 - c0, True, nd)
                        sp = ConvUmmaFn._kernel_spec(lin, nd, cout=c1 - c0)
                        dxp, _, _ = ops.conv_fwd_umma(g_parts[go][0], g_parts[go][1], w_hi, w_c,
                                                      None, sp, gn, gdims, residual=dxp,
                                                      fmt=ops.S3_FMT_FP16C, acc_scale=acc)
                    slices.append(dxp)
                dxp = slices[0] if len(slices) == 1 else torch.cat(slices, dim=-1)
                if cin_p != cin:
                    dxp = dxp[..., :cin].contiguous()
                dx = ops.pad_bwd(dxp, ctx.x_shape, pads, spec.pad_mode) if reflect else dxp
            elif reflect:
                pshape = tuple(s + p[0] + p[1] for s, p in zip(ctx.x_shape, pads))
                valid = dataclasses.replace(lin, pad_lo=(0, 0, 0), pad_hi=(0, 0, 0),
                                            pad_mode=S3_PAD_ZERO)
                dxp = ops.conv_dgrad(dy, w, valid, pshape)
                dx = ops.pad_bwd(dxp, ctx.x_shape, pads, spec.pad_mode)
            else:
                dx = ops.conv_dgrad(dy, w, fp32, ctx.x_shape)
        if want_w and ctx.wgrad_umma:
            # dW = sum_v x_pad[v + tap] (x) dy[v] on tcgen05 (voxels = the GEMM's K dimension), one
            # launch per (input group, output block); the zero-halo fp16 gradient tensors of the
            # input-gradient convolution are reused
            if g_parts is None:
                g_parts, g_halo = [], 1
                for go in range(gout):
                    blk = dy_p if gout == 1 else dy_p[..., 64 * go:64 * go + 64].contiguous()
                    g_parts.append(ops.pack_act_pad16(blk, split=False, fmt=ops.S3_FMT_FP16,
                                                      halo=S3_PAD_ZERO))
            rows = []
            for gi in range(ctx.gin):
                cg = min(cin, 64)
                cols = [ops.conv_wgrad_umma(xs[gi], g_parts[go][0], g_halo, n, dims, cg)
                        for go in range(gout)]
                rows.append(cols[0] if gout == 1 else torch.cat(cols, dim=-1))
            dw = rows[0] if ctx.gin == 1 else torch.cat(rows, dim=-2)
            if dw.shape[-1] != cout:
                dw = dw[..., :cout].contiguous()
            if tuple(dw.shape) != tuple(w.shape):
                dw = dw.reshape(w.shape)
            db = ops.conv_bias_grad(dy, cout) if ctx.has_bias else None
        return dx, dw, db, None, None, None, None


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
