"""Fused execution plan for a ``CustomNetwork``.

The reference runs ~150 eager TF layers per generator call
(sup3r/models/abstract.py:1081-1092): ``FlexiblePadding -> Conv -> Cropping -> LeakyReLU ->
[Expansion] -> [SkipConnection]`` repeated.  Here the layer list is pattern-matched once into
fused convolution steps (implicit padding, bias, activation, expansion scatter, residual add in
the kernel epilogue) and executed either on the exact-fp32 direct kernel or on the tcgen05
kernel (16-bit operands, fp32 accumulate; optional 3-pass split precision).  Anything that
does not match a pattern stays an eager layer step, so every config still runs.

Precision modes (``SUP3R_B200_PRECISION`` or the ``precision`` argument):
  ``fp32``   every convolution on the fp32 CUDA-core kernel;
  ``bf16``   64-channel 3x3[x3] reflect convolutions on tcgen05 with bf16 operands;
  ``bf16x3`` same, with hi/lo split operands (hi*hi + lo*hi + hi*lo) ~ fp32-grade results;
  ``fp16c``  fp16 operands + e4m3 correction rows (one fp16 and one e4m3 MMA pass per layer,
             ~2^-15 relative operand precision): the fastest mode inside the 1e-3 tolerance of
             the fp32 reference path (sup3r/models/abstract.py:1037-1105 runs fp32 end to end).
"""
from __future__ import annotations

import dataclasses
import os
from dataclasses import dataclass, field

import torch

from . import ops
from ._cabi import S3_ACT_LEAKY, S3_ACT_NONE, S3_PAD_REFLECT, S3_PAD_ZERO
from .network import (Activation, LeakyReLU, SkipConnection, SpatialExpansion,
                      SpatioTemporalExpansion, Sup3rAdder, Sup3rConcat, FlexiblePadding, _Conv,
                      _Cropping, run_exo_layer, same_pads, SUP3R_LAYERS, to_device_tensor,
                      KERAS_LEAKY_RELU_SLOPE)

PRECISIONS = ("fp32", "bf16", "bf16x3", "fp16c")
_FMT = {"fp32": 0, "bf16": ops.S3_FMT_BF16, "bf16x3": ops.S3_FMT_BF16, "fp16c": ops.S3_FMT_FP16C}


def default_precision():
    p = os.environ.get("SUP3R_B200_PRECISION", "fp16c")
    if p not in PRECISIONS:
        raise ValueError(f"SUP3R_B200_PRECISION must be one of {PRECISIONS}, got {p!r}")
    return p


@dataclass
class FusedConv:
    """pad -> conv -> crop -> [act] -> [expansion] -> [skip add] -> [skip stores]"""
    conv: _Conv
    pads: list            # implicit [(lo, hi)] per conv dim
    pad_mode: int
    act: int = S3_ACT_NONE
    alpha: float = 0.0
    r: int = 1            # depth_to_space
    m: int = 1            # temporal multiplier
    method: int = 0       # 0 nearest repeat, 1 depth_to_time
    roll: int = 0
    skip_add: str | None = None
    skip_store: list = field(default_factory=list)
    n_layers: int = 1     # layers consumed (for error messages)
    first_layer: int = 0

    def spec(self, in_shape):
        s = self.conv.spec(in_shape, extra_pad=self.pads, act=self.act, alpha=self.alpha,
                           pad_mode=self.pad_mode)
        rep = (1, 1, 1)
        d2t = 1
        if self.m > 1:
            if self.method == 1:
                d2t = self.m
            else:
                rep = (1, 1, self.m)
        return dataclasses.replace(s, d2s=self.r, d2t=d2t, t_roll=self.roll, out_repeat=rep)


@dataclass
class EagerStep:
    layer: object
    index: int


@dataclass
class SkipStep:
    name: str
    store: bool
    index: int


def build_steps(layers):
    """Pattern-match the layer list into fused / eager steps (static; shape independent)."""
    steps = []
    open_skips = set()
    i, n = 0, len(layers)

    def skip_kind(name):
        if name in open_skips:
            open_skips.discard(name)
            return "add"
        open_skips.add(name)
        return "store"

    while i < n:
        lyr = layers[i]
        pad = None
        j = i
        if isinstance(lyr, FlexiblePadding) and j + 1 < n and isinstance(layers[j + 1], _Conv) \
                and lyr.rank == layers[j + 1].nd + 2 and lyr.paddings[0] == [0, 0] \
                and lyr.paddings[-1] == [0, 0]:
            pad = lyr
            j += 1
        if isinstance(layers[j], _Conv):
            conv = layers[j]
            nd = conv.nd
            ok = True
            k = j + 1
            crop = None
            if k < n and isinstance(layers[k], _Cropping) and layers[k].nd == nd:
                crop = layers[k]
                k += 1
            extra = [(kk - 1, kk - 1) if conv.transposed else (0, 0) for kk in conv.kernel_size]
            p = pad.paddings[1:-1] if pad is not None else [[0, 0]] * nd
            c = crop.cropping if crop is not None else [(0, 0)] * nd
            q = [(p[d][0] + extra[d][0] - c[d][0], p[d][1] + extra[d][1] - c[d][1])
                 for d in range(nd)]
            mode = ops.PAD_CODES[pad.mode] if pad is not None else S3_PAD_ZERO
            if pad is not None or crop is not None or conv.transposed:
                if any(s != 1 for s in conv.strides) or conv.padding != "valid":
                    ok = False
                if any(v < 0 for t in q for v in t):
                    ok = False
                if conv.transposed and any(c[d][0] < extra[d][0] or c[d][1] < extra[d][1]
                                           for d in range(nd)) and pad is not None \
                        and mode != S3_PAD_ZERO:
                    ok = False  # the transposed conv's zero ring would be visible
                if conv.transposed and pad is None and crop is None:
                    q = list(extra)
            else:
                q = None  # the conv's own 'valid' / 'same' padding (shape dependent)
            if ok:
                fc = FusedConv(conv, q, mode, first_layer=i)
                fc.act = ops.ACT_CODES[conv.activation]
                if fc.act == S3_ACT_LEAKY:
                    fc.alpha = KERAS_LEAKY_RELU_SLOPE
                # activation / expansion in either order
                for _ in range(2):
                    if k < n and fc.act == S3_ACT_NONE and isinstance(layers[k], LeakyReLU):
                        fc.act, fc.alpha = S3_ACT_LEAKY, layers[k].alpha
                        k += 1
                    elif k < n and fc.act == S3_ACT_NONE and isinstance(layers[k], Activation) \
                            and ops.ACT_CODES[layers[k].activation] != S3_ACT_NONE:
                        fc.act = ops.ACT_CODES[layers[k].activation]
                        fc.alpha = KERAS_LEAKY_RELU_SLOPE
                        k += 1
                    elif k < n and fc.r == 1 and fc.m == 1 and \
                            isinstance(layers[k], SpatialExpansion) and nd == 2:
                        fc.r = layers[k]._spatial_mult
                        k += 1
                    elif k < n and fc.r == 1 and fc.m == 1 and \
                            isinstance(layers[k], SpatioTemporalExpansion) and nd == 3:
                        e = layers[k]
                        fc.r, fc.m = e._spatial_mult, e._temporal_mult
                        fc.method, fc.roll = e.method_code, e._t_roll
                        k += 1
                # skip connections: at most one add (plain mapping only), then stores
                if k < n and isinstance(layers[k], SkipConnection) and fc.r == 1 and fc.m == 1 \
                        and layers[k].name in open_skips:
                    skip_kind(layers[k].name)
                    fc.skip_add = layers[k].name
                    k += 1
                while k < n and isinstance(layers[k], SkipConnection) \
                        and layers[k].name not in open_skips:
                    skip_kind(layers[k].name)
                    fc.skip_store.append(layers[k].name)
                    k += 1
                fc.n_layers = k - i
                steps.append(fc)
                i = k
                continue
        if isinstance(lyr, SkipConnection):
            steps.append(SkipStep(lyr.name, skip_kind(lyr.name) == "store", i))
        else:
            steps.append(EagerStep(lyr, i))
        i += 1
    return steps


class Act:
    """An activation tensor held in f32 and / or padded 16-bit (hi, lo) form."""

    __slots__ = ("f32", "hi", "lo", "shape", "b16", "fmt")

    def __init__(self, shape, f32=None, hi=None, lo=None, b16=None, fmt=0):
        self.shape = tuple(shape)
        self.f32, self.hi, self.lo = f32, hi, lo
        self.b16 = b16          # unpadded 16-bit tensor (depth_to_space destination)
        self.fmt = fmt          # S3_FMT_* of (hi, lo)

    def need_f32(self):
        if self.f32 is None:
            if self.b16 is not None:
                self.f32 = self.b16.float()
            else:
                self.f32 = ops.unpack_act_pad16(self.hi, self.lo, len(self.shape) - 2,
                                                fmt=self.fmt)
        return self.f32

    def need_pad16(self, split, fmt=0):
        if self.hi is None or (split and self.lo is None) or self.fmt != fmt:
            self.hi, self.lo = ops.pack_act_pad16(self.need_f32(), split=split, fmt=fmt)
            self.fmt = fmt
        return self.hi, (self.lo if split else None)


def _umma_ok(fc: FusedConv, in_shape, precision):
    if precision == "fp32" or fc.pads is None:
        return False
    conv = fc.conv
    nd = conv.nd
    # (narrow inputs, e.g. the 4-feature first layer, run on the tensor cores with the channel
    # axis zero-padded to 64: cheaper than the latency-bound generic kernel)
    if nd not in (2, 3):
        return False
    if in_shape[-1] > 64:
        # 64 feature channels + a few Sup3rConcat exo channels (2-D): tensor cores for the 64,
        # the fp32 kernel for the rest, summed before the activation
        if not (nd == 2 and in_shape[-1] <= 72 and conv.filters == 64
                and precision in ("bf16", "fp16c") and fc.r == 1 and fc.m == 1
                and fc.skip_add is None):
            return False
    if conv.filters > 256:
        # wide scatter heads (e.g. 64 -> 1600 with 5x depth_to_space, 64 -> 768 with 24x
        # depth_to_time) run as channel slices of <= 256 when the mapped voxel is 16-aligned
        d2t = fc.m if (fc.m > 1 and fc.method == 1) else 1
        groups = fc.r * fc.r * d2t
        if groups == 1 or conv.filters % groups or (conv.filters // groups) % 16:
            return False
        if fc.m > 1 and fc.method == 0:
            return False
    if in_shape[-1] < 64 and (precision not in ("bf16", "fp16c") or conv.filters < 32):
        return False
    if any(k != 3 for k in conv.kernel_size) or any(s != 1 for s in conv.strides):
        return False
    if fc.pad_mode != S3_PAD_REFLECT or any(tuple(p) != (1, 1) for p in fc.pads):
        return False
    if min(in_shape[1:-1]) < 2:
        return False
    return True


class Plan:
    """Executable plan of one network for one precision mode."""

    def __init__(self, net, precision=None):
        self.net = net
        self.precision = precision or default_precision()
        if self.precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        self.steps = build_steps(net.layers)
        self.fmt = _FMT[self.precision]
        self.head_single_pass = os.environ.get("SUP3R_B200_HEAD_2PASS", "0") != "1"
        self._wcache = {}
        self._graphs = {}

    # -- weights ---------------------------------------------------------------------
    def _packed(self, conv, split, fmt=None):
        fmt = self.fmt if fmt is None else fmt
        key = (id(conv), split, fmt)
        ver = (conv.kernel.version, conv.kernel.value.data_ptr())
        hit = self._wcache.get(key)
        if hit is None or hit[0] != ver:
            with torch.no_grad():
                w = conv.conv_kernel().detach()
                if w.shape[-2] < 64:     # zero rows for the padded input channels
                    w = torch.nn.functional.pad(w, (0, 0, 0, 64 - w.shape[-2]))
                packed = ops.pack_weights_umma(w, split=split, ndim=conv.nd, fmt=fmt)
            hit = (ver, *packed) if len(packed) == 3 else (ver, *packed, 0.0)
            self._wcache[key] = hit
        return hit[1], hit[2], hit[3]

    def invalidate(self):
        self._wcache.clear()
        self._graphs.clear()

    # -- execution -------------------------------------------------------------------
    def run(self, x, exo=None, post_scale=None, post_shift=None):
        """Inference forward (no autograd).  ``x``: fp32 device tensor; ``exo``: {layer name:
        tensor}.  ``post_scale/shift``: per-channel affine fused into the last convolution
        (un-normalisation) when the network ends with a fused conv."""
        net = self.net
        exo = exo or {}
        split = self.precision in ("bf16x3", "fp16c")
        if not net.built:
            net.build(tuple(x.shape), {k: v.shape[-1] for k, v in exo.items()})
        cur = Act(x.shape, f32=x)
        skips = {}
        steps = self.steps
        applied_post = False
        with torch.no_grad():
            for si, st in enumerate(steps):
                try:
                    if isinstance(st, FusedConv):
                        last = si == len(steps) - 1
                        # the fused affine indexes CONV channels: only for channel-preserving
                        # output maps (no depth_to_space / depth_to_time)
                        fuse = last and st.r == 1 and not (st.m > 1 and st.method == 1)
                        ps = post_scale if fuse else None
                        pf = post_shift if fuse else None
                        cur = self._run_conv(st, cur, skips, steps, si, split, ps, pf)
                        # (routes that cannot fuse the affine -- channel-split convs, wide
                        # scatter heads -- leave it to the channel_affine below)
                        applied_post = applied_post or (last and ps is not None
                                                        and self._post_fused)
                    elif isinstance(st, SkipStep):
                        if st.store:
                            skips[st.name] = cur
                        else:
                            t = cur.need_f32()
                            cache = skips.pop(st.name)
                            if tuple(cache.shape) != tuple(t.shape):
                                raise RuntimeError(f'SkipConnection "{st.name}" shape mismatch')
                            cur = Act(t.shape, f32=ops.add(t, cache.need_f32()))
                    else:
                        lyr = st.layer
                        t = cur.need_f32()
                        if isinstance(lyr, SUP3R_LAYERS):
                            y = run_exo_layer(lyr, t, exo)
                        else:
                            y = lyr.forward(t)
                        cur = Act(y.shape, f32=y)
                except Exception as e:
                    idx = st.first_layer if isinstance(st, FusedConv) else st.index
                    raise RuntimeError(
                        f'Could not run layer #{idx} "{net.layers[idx]}" on tensor of shape '
                        f"{cur.shape}") from e
            out = cur.need_f32()
            if post_scale is not None and not applied_post:
                out = ops.channel_affine(out, post_scale, post_shift)
        return out

    def _next_wants_pad16(self, steps, si, out_shape):
        """Does the consumer of step si's output run on the tcgen05 kernel?"""
        for nxt in steps[si + 1:]:
            if isinstance(nxt, FusedConv):
                return _umma_ok(nxt, out_shape, self.precision)
            return False
        return False

    def _run_conv(self, st, cur, skips, steps, si, split, post_scale, post_shift):
        conv = st.conv
        shp = cur.shape
        self._post_fused = True
        if not conv.built:
            raise RuntimeError("network weights are not built")
        if st.pads is None:  # the layer's own padding
            spec = conv.spec(shp, act=st.act, alpha=st.alpha)
            rep, d2t = (1, 1, 1), 1
            if st.m > 1:
                if st.method == 1:
                    d2t = st.m
                else:
                    rep = (1, 1, st.m)
            spec = dataclasses.replace(spec, d2s=st.r, d2t=d2t, t_roll=st.roll, out_repeat=rep)
        else:
            spec = st.spec(shp)
        n, dims, cin, ndim = ops.dims3(shp)
        _, od, oc = spec.out_dims(n, dims)
        out_shape = ops._shape_from(n, od, oc, ndim)
        res_act = None
        if st.skip_add is not None:
            res_act = skips.pop(st.skip_add)
            if tuple(res_act.shape) != tuple(out_shape):
                raise RuntimeError(f'SkipConnection "{st.skip_add}" shape mismatch: '
                                   f"{tuple(res_act.shape)} vs {tuple(out_shape)}")
        plain = st.r == 1 and st.m == 1 or (st.m > 1 and st.method == 0 and st.r == 1)
        want16 = plain and self._next_wants_pad16(steps, si, out_shape)
        last = si == len(steps) - 1
        fmt = self.fmt
        c_mode = self.precision == "fp16c"
        umma = _umma_ok(st, shp, self.precision)
        if c_mode:
            # (hi, corr) pairs are written by the tcgen05 kernels' plain 64-channel epilogues only
            want16 = (want16 and umma and out_shape[-1] == 64 and min(out_shape[1:-1]) >= 4
                      and post_scale is None and cin <= 64)
        # SkipConnection caches travel as a 16-bit (hi, lo) pair in the padded layout when the
        # producer writes 16-bit output anyway: hi is the next convolution's operand, hi + lo
        # (~16 mantissa bits) is the addend of the consuming convolution's epilogue
        # (phygnn SkipConnection semantics, call site sup3r/models/abstract.py:1081-1092)
        # (only the 3-D ring kernel's epilogue takes the pair back as a residual: elsewhere the
        # cache is written as f32 next to the 16-bit operand tensor)
        pair_ok = (len(out_shape) == 5 and out_shape[-1] == 64 and st.r == 1
                   and (st.m == 1 or st.method == 0) and min(out_shape[1:-1]) >= 4)
        pair_skip = want16 and bool(st.skip_store) and pair_ok and (c_mode or not split)
        want32 = (not want16) or last or (bool(st.skip_store) and not pair_skip)
        bias = conv.bias.value.detach() if conv.bias is not None else None
        if os.environ.get("SUP3R_B200_TRACE_PLAN"):
            route = ("umma" if umma else
                     ("small_bf16" if (self.precision == "bf16" and not want16
                                       and ops.small_bf16_ok(spec)) else "direct"))
            print(f"[plan] conv {tuple(shp)} -> cout {conv.filters} r={st.r} m={st.m} "
                  f"skip_add={st.skip_add} store={st.skip_store} want16={want16} route={route}")
        # depth_to_space head feeding the narrow output convolution: hand the high-resolution
        # tensor over as unpadded bf16 (8-channel voxels = 16 B) instead of f32
        map16 = (self.precision in ("bf16", "fp16c") and umma and st.r > 1
                 and st.m == 1 and oc == 8 and res_act is None and not st.skip_store and not last
                 and post_scale is None and self._next_is_small_bf16(steps, si, out_shape))
        if umma and cin > 64:
            # split over input-channel groups: direct kernel on the exo channels first (no bias,
            # no activation), then the tensor-core conv on the 64 features adds it pre-activation
            xf = cur.need_f32()
            w = conv.conv_kernel().detach()
            sp_rem = dataclasses.replace(spec, cin=cin - 64, act=S3_ACT_NONE, alpha=0.0)
            part = ops.conv_fwd(xf[..., 64:].contiguous(), w[..., 64:, :].contiguous(), None, sp_rem)
            x_hi, x_lo = ops.pack_act_pad16(xf[..., :64].contiguous(), split=split, fmt=fmt)
            key = (id(conv), split, "main64", fmt)
            ver = (conv.kernel.version, conv.kernel.value.data_ptr())
            hit = self._wcache.get(key)
            if hit is None or hit[0] != ver:
                hit = (ver, *ops.pack_weights_umma(w[..., :64, :].contiguous(), split=split,
                                                   ndim=conv.nd, fmt=fmt))
                self._wcache[key] = hit
            sp_main = dataclasses.replace(spec, cin=64, res_pre_act=1)
            self._post_fused = False
            y, y_hi, y_lo = ops.conv_fwd_umma(x_hi, x_lo, hit[1], hit[2], bias, sp_main, n, dims,
                                              residual=part, want_f32=want32, want_pad16=want16,
                                              want_lo=c_mode and want16, fmt=fmt,
                                              acc_scale=hit[3] if len(hit) > 3 else 0.0)
            return self._finish_conv(st, Act(out_shape, f32=y, hi=y_hi, lo=y_lo, fmt=fmt), skips)
        if umma:
            if cin < 64:
                x_hi, x_lo = ops.pack_act_pad16(
                    torch.nn.functional.pad(cur.need_f32(), (0, 64 - cin)), split=split, fmt=fmt)
                spec = dataclasses.replace(spec, cin=64)
            else:
                x_hi, x_lo = cur.need_pad16(split, fmt)
            if conv.filters > 256:
                self._post_fused = False
                y = self._run_wide_head(conv, spec, x_hi, x_lo, bias, n, dims, split, out_shape)
                return self._finish_conv(st, Act(out_shape, f32=y), skips)
            if c_mode and map16 and self.head_single_pass:
                # the depth_to_space head in front of the output convolution: one fp16 pass (its
                # operand rounding does not compound through further layers; emulation
                # tools/scheme_numerics.py: 4.3e-4 -> 5.2e-4 max-rel on the north-star generator)
                x_lo, fmt = None, ops.S3_FMT_FP16
                w_hi, w_lo, acc_scale = self._packed(conv, False, fmt)
            else:
                w_hi, w_lo, acc_scale = self._packed(conv, split)
            res16 = (res_act is not None and (c_mode or not split) and not want32 and want16
                     and self._ring16_ok(st, out_shape) and post_scale is None
                     and res_act.hi is not None and res_act.fmt == fmt
                     and (not c_mode or res_act.lo is not None))
            y, y_hi, y_lo = ops.conv_fwd_umma(
                x_hi, x_lo, w_hi, w_lo, bias, spec, n, dims,
                residual=None if (res16 or res_act is None) else res_act.need_f32(),
                res_hi=res_act.hi if res16 else None, res_lo=res_act.lo if res16 else None,
                post_scale=post_scale, post_shift=post_shift, want_f32=want32 and not map16,
                want_pad16=want16, want_lo=pair_skip or (c_mode and want16), want_map16=map16,
                fmt=fmt, acc_scale=acc_scale)
            if map16:
                out = Act(out_shape, b16=y_hi)
                return out
        elif (self.precision in ("bf16", "fp16c") and not want16 and ops.small_bf16_ok(spec)
              and spec.pad_mode == S3_PAD_REFLECT):
            # narrow high-resolution output convolution: warp-level tensor cores, 16-bit operands
            # (bf16 / fp16: the last layer's operand rounding does not compound)
            y = ops.conv_fwd_small_bf16(
                cur.b16 if (cur.b16 is not None and cur.f32 is None and shp[-1] == 8)
                else cur.need_f32(), conv.conv_kernel().detach(), bias, spec,
                residual=None if res_act is None else res_act.need_f32(),
                post_scale=post_scale, post_shift=post_shift, fp16=c_mode)
            y_hi = y_lo = None
        else:
            x = cur.need_f32()
            res = ops.conv_fwd(x, conv.conv_kernel().detach(), bias, spec,
                               residual=None if res_act is None else res_act.need_f32(),
                               post_scale=post_scale, post_shift=post_shift,
                               want_pad16=want16, split=split or pair_skip, want_f32=want32)
            fmt = 0   # (the direct kernel writes bf16 pairs; never asked for in fp16c mode)
            y, y_hi, y_lo = res if want16 else (res, None, None)
        return self._finish_conv(st, Act(out_shape, f32=y, hi=y_hi, lo=y_lo, fmt=fmt), skips)

    @staticmethod
    def _finish_conv(st, out, skips):
        for name in st.skip_store:
            skips[name] = out
        return out

    def _run_wide_head(self, conv, spec, x_hi, x_lo, bias, n, dims, split, out_shape):
        """cout > 256 with a depth_to_space / depth_to_time map: channel slices of <= 256 into
        one f32 destination (``cout_total`` / ``cout_base`` of ``s3_conv_desc``)."""
        key = (id(conv), split, "wide")
        ver = (conv.kernel.version, conv.kernel.value.data_ptr())
        hit = self._wcache.get(key)
        if hit is None or hit[0] != ver:
            with torch.no_grad():
                w = conv.conv_kernel().detach()
                if w.shape[-2] < 64:     # zero rows for the padded input channels
                    w = torch.nn.functional.pad(w, (0, 0, 0, 64 - w.shape[-2]))
                packs = []
                for cb in range(0, conv.filters, 256):
                    nc = min(256, conv.filters - cb)
                    pk = ops.pack_weights_umma(w[..., cb:cb + nc].contiguous(), split=split,
                                               ndim=conv.nd, fmt=self.fmt)
                    packs.append((cb, nc, *pk) if len(pk) == 3 else (cb, nc, *pk, 0.0))
            hit = (ver, packs)
            self._wcache[key] = hit
        y = torch.empty(out_shape, device=x_hi.device, dtype=torch.float32)
        for cb, nc, w_hi, w_lo, acc_scale in hit[1]:
            sp = dataclasses.replace(spec, cout=nc, cout_total=conv.filters, cout_base=cb)
            ops.conv_fwd_umma(x_hi, x_lo, w_hi, w_lo,
                              None if bias is None else bias[cb:cb + nc].contiguous(), sp, n, dims,
                              out=y, fmt=self.fmt, acc_scale=acc_scale)
        return y

    def _next_is_small_bf16(self, steps, si, out_shape):
        """Is the consumer of step si's output the narrow tensor-core convolution?"""
        for nxt in steps[si + 1:]:
            if not isinstance(nxt, FusedConv) or nxt.pads is None:
                return False
            if _umma_ok(nxt, out_shape, self.precision) or nxt.skip_add is not None:
                return False
            sp = nxt.spec(out_shape)
            return (ops.small_bf16_ok(sp) and sp.pad_mode == S3_PAD_REFLECT and sp.cin == 8
                    and nxt.r == 1 and nxt.m == 1)
        return False

    @staticmethod
    def _ring16_ok(st, out_shape):
        """Host mirror of the C side's conditions for the ring kernel's 16-bit tile epilogue
        (s3_conv_fwd_umma: 3-D, 64 -> 64 channels, plain output map, extents >= 4)."""
        return (len(out_shape) == 5 and out_shape[-1] == 64 and st.r == 1 and st.m == 1
                and min(out_shape[1:-1]) >= 4)

    @staticmethod
    def _train_umma_ok(st, in_shape):
        """Convolutions of the training forward that run on the tcgen05 kernels: 3x3[x3], stride
        1, reflect-pad-1, cin <= 64 (narrow inputs zero-padded), 32 <= cout <= 256."""
        conv = st.conv
        return (st.pads is not None and conv.nd in (2, 3) and in_shape[-1] <= 64
                and 32 <= conv.filters <= 256 and all(k == 3 for k in conv.kernel_size)
                and all(s == 1 for s in conv.strides) and st.pad_mode == S3_PAD_REFLECT
                and all(tuple(p) == (1, 1) for p in st.pads) and min(in_shape[1:-1]) >= 2)

    @staticmethod
    def _train_umma_same_ok(st, in_shape):
        """keras ``padding='same'`` 3x3[x3] convolutions (the discriminator) on tcgen05: stride 1,
        or stride 2 on every dim; cin <= 64 or a multiple of 64; 32 <= cout <= 256 or 64 | cout <= 512."""
        conv = st.conv
        if st.pads is not None or conv.transposed or conv.padding != "same":
            return False
        if st.r > 1 or st.m > 1:
            return False
        strides = tuple(conv.strides)
        if not (all(s == 1 for s in strides) or all(s == 2 for s in strides)):
            return False
        cin = in_shape[-1]
        filters_ok = 32 <= conv.filters <= 256 or (conv.filters % 64 == 0 and conv.filters <= 512)
        return (conv.nd in (2, 3) and (cin <= 64 or (cin % 64 == 0 and cin <= 512))
                and filters_ok
                and all(k == 3 for k in conv.kernel_size) and min(in_shape[1:-1]) >= 2)

    def _refresh_weight_maxima(self):
        """max |w| of every convolution kernel whose weights changed since the last call, with ONE
        device -> host read for the whole network (the fp16c weight packing needs it to pick its
        power-of-two scale; per layer it would be a sync per layer and step)."""
        stale = []
        for st in self.steps:
            if isinstance(st, FusedConv) and st.conv.built:
                cache = st.conv.__dict__.setdefault("_umma_train_cache", {})
                hit = cache.get("wmax")
                if hit is None or hit[0] != st.conv.kernel.version:
                    stale.append((st.conv, cache))
        if stale:
            with torch.no_grad():   # multi-tensor inf-norm: a couple of launches for all kernels
                norms = torch._foreach_norm([c.kernel.value.detach() for c, _ in stale],
                                            float("inf"))
                vals = torch.stack([n.reshape(()) for n in norms]).cpu()
            for (conv, cache), v in zip(stale, vals.tolist()):
                cache["wmax"] = (conv.kernel.version, float(v))

    # -- training forward (autograd) -------------------------------------------------
    def forward_train(self, x, exo=None):
        """Differentiable forward on the fp32 kernels: the same fused steps, each wrapped in
        an autograd Function (conv with implicit padding + bias + activation; expansion; skip
        add).  Stands in for ``_tf_generate`` / ``_tf_discriminate`` under a GradientTape
        (abstract.py:1131-1173, base.py:283-313)."""
        from .autograd import AddFn, ConvFn, ConvUmmaFn, ExpandFn
        net = self.net
        tensor_cores = self.precision != "fp32"
        exo = exo or {}
        if not net.built:
            net.build(tuple(x.shape), {k: v.shape[-1] for k, v in exo.items()})
        if tensor_cores:
            self._refresh_weight_maxima()
        skips = {}
        cur = x
        for st in self.steps:
            try:
                if isinstance(st, FusedConv):
                    conv = st.conv
                    if st.pads is None:
                        spec = conv.spec(cur.shape, act=st.act, alpha=st.alpha)
                    else:
                        spec = conv.spec(cur.shape, extra_pad=st.pads, act=st.act,
                                         alpha=st.alpha, pad_mode=st.pad_mode)
                    b = conv.bias.value if conv.bias is not None else None
                    if tensor_cores and self._train_umma_ok(st, tuple(cur.shape)):
                        # forward + input / weight gradients on tcgen05 (fp16c operands)
                        cache = conv.__dict__.setdefault("_umma_train_cache", {})
                        cur = ConvUmmaFn.apply(cur, conv.conv_kernel(), b, spec, cache, None,
                                               conv.kernel.version)
                    elif tensor_cores and self._train_umma_same_ok(st, tuple(cur.shape)):
                        # keras padding='same' (the discriminator): zero-halo operands.  Stride 2
                        # = the stride-1 convolution sampled at the positions TF's asymmetric
                        # 'same' padding selects (the tensor cores redo 8x the work of the strided
                        # convolution and are still far faster than the fp32 kernel)
                        cache = conv.__dict__.setdefault("_umma_train_cache", {})
                        nd = conv.nd
                        one = (0,) * (3 - nd) + (1,) * nd
                        sp1 = dataclasses.replace(spec, stride=(1, 1, 1), pad_lo=one, pad_hi=one,
                                                  pad_mode=S3_PAD_ZERO)
                        full = ConvUmmaFn.apply(cur, conv.conv_kernel(), b, sp1, cache, S3_PAD_ZERO,
                                                conv.kernel.version)
                        if any(s > 1 for s in conv.strides):
                            idx = (slice(None),) + tuple(
                                slice(1 - (n % 2), None, 2) for n in cur.shape[1:-1]) + (slice(None),)
                            full = full[idx].contiguous()
                        cur = full
                    else:
                        cur = ConvFn.apply(cur, conv.conv_kernel(), b, spec, tensor_cores)
                    if st.r > 1 or st.m > 1:
                        cur = ExpandFn.apply(cur, st.r, st.m, st.method, st.roll)
                    if st.skip_add is not None:
                        cache = skips.pop(st.skip_add)
                        if tuple(cache.shape) != tuple(cur.shape):
                            raise RuntimeError(f'SkipConnection "{st.skip_add}" shape mismatch')
                        cur = AddFn.apply(cur, cache)
                    for name in st.skip_store:
                        skips[name] = cur
                elif isinstance(st, SkipStep):
                    if st.store:
                        skips[st.name] = cur
                    else:
                        cache = skips.pop(st.name)
                        if tuple(cache.shape) != tuple(cur.shape):
                            raise RuntimeError(f'SkipConnection "{st.name}" shape mismatch')
                        cur = AddFn.apply(cur, cache)
                else:
                    lyr = st.layer
                    if isinstance(lyr, SUP3R_LAYERS):
                        cur = run_exo_layer(lyr, cur, exo)
                    else:
                        cur = lyr.forward(cur)
            except Exception as e:
                idx = st.first_layer if isinstance(st, FusedConv) else st.index
                raise RuntimeError(
                    f'Could not run layer #{idx} "{net.layers[idx]}" on tensor of shape '
                    f"{tuple(cur.shape)}") from e
        return cur

    # -- CUDA graph ------------------------------------------------------------------
    def run_graphed(self, x, exo=None, post_scale=None, post_shift=None):
        """Replay a captured CUDA graph of ``run`` for this input shape (launch-bound nets)."""
        exo = exo or {}
        if not self.net.built:
            self.net.build(tuple(x.shape), {k: v.shape[-1] for k, v in exo.items()})
        # packed tensor-core weights are baked into the captured graph: re-capture after any
        # weight update (optimizer step, set_weights, load)
        wver = tuple((st.conv.kernel.version, st.conv.kernel.value.data_ptr())
                     for st in self.steps if isinstance(st, FusedConv))
        if wver != getattr(self, "_graph_wver", None):
            self._graphs.clear()
            self._graph_wver = wver
        key = (tuple(x.shape), tuple(sorted((k, tuple(v.shape)) for k, v in exo.items())),
               post_scale is not None)
        g = self._graphs.get(key)
        if g is None:
            sx = x.clone()
            sexo = {k: v.clone() for k, v in exo.items()}
            sps = post_scale.clone() if post_scale is not None else None
            spf = post_shift.clone() if post_shift is not None else None
            self.run(sx, sexo, sps, spf)  # warm-up: builds weights, packs, sets attributes
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            launches0 = ops.launch_count()
            with torch.cuda.graph(graph):
                out = self.run(sx, sexo, sps, spf)
            # kernels recorded in the graph: every replay launches all of them again
            g = dict(graph=graph, x=sx, exo=sexo, ps=sps, pf=spf, out=out,
                     launches=ops.launch_count() - launches0)
            self._graphs[key] = g
        g["x"].copy_(x)
        for k, v in exo.items():
            g["exo"][k].copy_(v)
        if post_scale is not None:
            g["ps"].copy_(post_scale)
            g["pf"].copy_(post_shift)
        g["graph"].replay()
        ops._count(g["launches"])
        return g["out"]
