"""``CustomNetwork`` and the layer vocabulary of the sup3r generator / discriminator configs,
executing on our CUDA kernels.

This module mirrors the Python object protocol that ``sup3r.models`` consumes from
``phygnn.CustomNetwork`` / ``phygnn.layers.custom_layers`` / ``tf.keras.layers``
(sup3r/models/abstract.py:96-101, 1081-1092; sup3r/models/interface.py:69, 84, 105-121;
sup3r/models/utilities.py:9-27):

* ``CustomNetwork(hidden_layers=[...], name=...)``, ``.layers``, ``.weights``, iteration,
  ``CustomNetwork.load(path)``, ``.save(path)``, ``CustomNetwork.seed(s)``;
* callable layers ``layer(x)`` / exo layers ``layer(x, hi_res_exo)``, attributes ``name``,
  ``rank`` (FlexiblePadding), ``_spatial_mult`` / ``_temporal_mult`` (expansions);
* results expose ``.numpy()`` and ``.shape``.

The ``hidden_layers`` JSON dialect is phygnn's: a list of ``{"class": ..., **kwargs}`` dicts
with ``{"n": N, "repeat": [...]}`` blocks; ``SkipConnection`` instances are shared by name.
"""
from __future__ import annotations

import copy
import json
import logging
import pickle

import numpy as np
import torch

from . import ops
from .autograd import (ActFn, AddFn, ConcatFn, ConvFn, CropFn, DenseFn, ExpandFn, PadFn)
from ._cabi import S3_ACT_LEAKY, S3_ACT_NONE, S3_PAD_ZERO

logger = logging.getLogger(__name__)

_GLOBAL_SEED = [0]


# slope of the STRING activation 'leaky_relu' (keras.activations.leaky_relu: negative_slope 0.2);
# the LeakyReLU LAYER carries its own alpha (keras default 0.3)
KERAS_LEAKY_RELU_SLOPE = 0.2


def default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
        else torch.device("cpu")


class DeviceArray(torch.Tensor):
    """A CUDA tensor that also answers ``.numpy()`` like the TF eager tensors the reference's
    ``generate`` receives from layers (abstract.py:1100)."""

    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).cpu().numpy()


def to_device_tensor(x, device=None):
    """numpy / torch input -> contiguous fp32 tensor on the compute device."""
    device = device or default_device()
    if isinstance(x, torch.Tensor):
        t = x.as_subclass(torch.Tensor)
        if t.device != device or t.dtype != torch.float32:
            t = t.to(device=device, dtype=torch.float32)
        return t.contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(device)


class Variable:
    """Trainable weight: a named fp32 device tensor (stand-in for ``tf.Variable``)."""

    def __init__(self, name, value):
        self.name = name
        self.value = value.requires_grad_(True)
        self.version = 0

    @property
    def shape(self):
        return tuple(self.value.shape)

    def numpy(self):
        return self.value.detach().cpu().numpy()

    def assign(self, new):
        with torch.no_grad():
            self.value.copy_(torch.as_tensor(np.asarray(new), dtype=torch.float32)
                             .reshape(self.value.shape))
        self.version += 1

    def __repr__(self):
        return f"<Variable {self.name} {self.shape}>"


def _tuple(v, n):
    return (int(v),) * n if np.isscalar(v) else tuple(int(i) for i in v)


def same_pads(size, k, s):
    """TensorFlow SAME padding of one dim -> (lo, hi)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


# ------------------------------------------------------------------------------- layers
class Layer:
    """Base layer.  ``forward`` maps torch -> torch (autograd aware); ``__call__`` also accepts
    numpy and wraps the result so that ``.numpy()`` works."""

    has_weights = False

    def __init__(self, name=None):
        self.name = name or self._default_name()
        self.built = False

    _counters = {}

    def _default_name(self):
        # keras-style snake_case auto names: conv3d, conv3d_1, ...
        base = "".join("_" + c.lower() if c.isupper() else c for c in type(self).__name__)
        base = base.lstrip("_").replace("_d", "d")
        n = Layer._counters.get(base, 0)
        Layer._counters[base] = n + 1
        return base if n == 0 else f"{base}_{n}"

    @property
    def weights(self):
        return []

    def out_shape(self, shp):
        return tuple(shp)

    def build(self, shp, rng, device):
        self.built = True

    def forward(self, x):
        raise NotImplementedError

    def __call__(self, x, *extra):
        t = to_device_tensor(x)
        extra = tuple(to_device_tensor(e) for e in extra)
        if not self.built:
            self.build(tuple(t.shape), np.random.default_rng(_GLOBAL_SEED[0]), t.device)
        y = self.forward(t, *extra)
        return y.as_subclass(DeviceArray)

    def get_config(self):
        return {"class": type(self).__name__}

    def __repr__(self):
        return f"<{type(self).__name__} name={self.name!r}>"


class FlexiblePadding(Layer):
    """``tf.pad(x, paddings, mode)``; exposes ``rank`` (interface.py:84)."""

    def __init__(self, paddings, mode="REFLECT", option="tf", name=None):
        super().__init__(name)
        self.paddings = [[int(p[0]), int(p[1])] for p in paddings]
        self.rank = len(self.paddings)
        self.mode = mode.upper()
        if self.mode not in ops.PAD_CODES:
            raise ValueError(f"unknown padding mode {mode!r}")

    def out_shape(self, shp):
        if len(shp) != self.rank:
            raise RuntimeError(f"FlexiblePadding of rank {self.rank} got a tensor of rank {len(shp)}")
        lim = {"REFLECT": 1, "SYMMETRIC": 0}.get(self.mode)
        if lim is not None:
            for n, (lo, hi) in zip(shp, self.paddings):
                if max(lo, hi) > n - lim:
                    raise RuntimeError(f"{self.mode} padding ({lo}, {hi}) is too large for a "
                                       f"dimension of extent {n} (shape {tuple(shp)})")
        return tuple(n + lo + hi for n, (lo, hi) in zip(shp, self.paddings))

    def forward(self, x):
        self.out_shape(x.shape)
        return PadFn.apply(x, self.paddings, ops.PAD_CODES[self.mode])


class _Conv(Layer):
    nd = 2
    transposed = False
    has_weights = True

    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None,
                 use_bias=True, name=None, **_):
        super().__init__(name)
        self.filters = int(filters)
        self.kernel_size = _tuple(kernel_size, self.nd)
        self.strides = _tuple(strides, self.nd)
        self.padding = str(padding).lower()
        if self.padding not in ("valid", "same"):
            raise ValueError(f"unknown padding {padding!r}")
        self.activation = activation
        if activation not in ops.ACT_CODES:
            raise ValueError(f"unknown activation {activation!r}")
        self.use_bias = bool(use_bias)
        self.kernel = None
        self.bias = None
        if self.transposed and (any(s != 1 for s in self.strides) or self.padding != "valid"):
            raise NotImplementedError("transposed convolutions: only strides 1 / padding valid")

    @property
    def weights(self):
        return [v for v in (self.kernel, self.bias) if v is not None]

    def out_shape(self, shp):
        if len(shp) != self.nd + 2:
            raise RuntimeError(f"{type(self).__name__} expects a {self.nd + 2}-D tensor, got "
                               f"shape {tuple(shp)}")
        sp = []
        for n, k, s in zip(shp[1:-1], self.kernel_size, self.strides):
            if self.transposed:
                o = (n - 1) * s + k
            elif self.padding == "same":
                o = -(-n // s)
            else:
                o = (n - k) // s + 1
            if o <= 0:
                raise RuntimeError(f"{type(self).__name__}: input extent {n} too small for kernel "
                                   f"{k} (shape {tuple(shp)})")
            sp.append(o)
        return (shp[0], *sp, self.filters)

    def kernel_shape(self, cin):
        if self.transposed:
            return self.kernel_size + (self.filters, cin)
        return self.kernel_size + (cin, self.filters)

    def build(self, shp, rng, device):
        cin = int(shp[-1])
        if self.kernel is None:
            rf = int(np.prod(self.kernel_size))
            lim = np.sqrt(6.0 / (rf * cin + rf * self.filters))  # keras glorot_uniform
            k = rng.uniform(-lim, lim, size=self.kernel_shape(cin)).astype(np.float32)
            self.kernel = Variable(f"{self.name}/kernel:0", torch.from_numpy(k).to(device))
            if self.use_bias:
                self.bias = Variable(f"{self.name}/bias:0",
                                     torch.zeros(self.filters, dtype=torch.float32, device=device))
        self.built = True

    def conv_kernel(self):
        """Kernel as a plain cross-correlation kernel ``(*k, cin, cout)`` (tensor, autograd
        aware): transposed kernels are flipped and their channel axes swapped."""
        w = self.kernel.value
        if self.transposed:
            w = w.flip(dims=tuple(range(self.nd))).transpose(-1, -2).contiguous()
        return w

    def spec(self, shp, extra_pad=None, act=None, alpha=None, pad_mode=S3_PAD_ZERO):
        """ConvSpec for input shape ``shp``; ``extra_pad``: [(lo, hi)] per conv dim overriding
        the layer's own implicit padding."""
        nd = self.nd
        if self.kernel is not None:
            cin_built = self.kernel.shape[-1] if self.transposed else self.kernel.shape[-2]
            if int(shp[-1]) != int(cin_built):
                raise RuntimeError(f'{type(self).__name__} "{self.name}" was built for {cin_built} '
                                   f"input channels but got a tensor of shape {tuple(shp)}")
        if extra_pad is None:
            if self.transposed:
                extra_pad = [(k - 1, k - 1) for k in self.kernel_size]
            elif self.padding == "same":
                extra_pad = [same_pads(n, k, s) for n, k, s in
                             zip(shp[1:-1], self.kernel_size, self.strides)]
            else:
                extra_pad = [(0, 0)] * nd
        z = 3 - nd
        a = ops.ACT_CODES[self.activation] if act is None else act
        if alpha is None:
            alpha = KERAS_LEAKY_RELU_SLOPE if a == S3_ACT_LEAKY else 0.0
        return ops.ConvSpec(nd, int(shp[-1]), self.filters, (1,) * z + self.kernel_size,
                            stride=(1,) * z + self.strides,
                            pad_lo=(0,) * z + tuple(int(p[0]) for p in extra_pad),
                            pad_hi=(0,) * z + tuple(int(p[1]) for p in extra_pad),
                            pad_mode=pad_mode, act=a, alpha=alpha)

    def forward(self, x):
        self.out_shape(x.shape)
        b = self.bias.value if self.bias is not None else None
        return ConvFn.apply(x, self.conv_kernel(), b, self.spec(x.shape))


class Conv2D(_Conv):
    nd = 2


class Conv3D(_Conv):
    nd = 3


class Conv2DTranspose(_Conv):
    nd = 2
    transposed = True


class Conv3DTranspose(_Conv):
    nd = 3
    transposed = True


class _Cropping(Layer):
    nd = 2

    def __init__(self, cropping, name=None):
        super().__init__(name)
        if np.isscalar(cropping):
            c = [(int(cropping), int(cropping))] * self.nd
        else:
            c = [(int(v), int(v)) if np.isscalar(v) else (int(v[0]), int(v[1])) for v in cropping]
        self.cropping = c

    def _full(self):
        return [(0, 0)] + list(self.cropping) + [(0, 0)]

    def out_shape(self, shp):
        if len(shp) != self.nd + 2:
            raise RuntimeError(f"{type(self).__name__} expects a {self.nd + 2}-D tensor")
        out = tuple(n - lo - hi for n, (lo, hi) in zip(shp, self._full()))
        if min(out) <= 0:
            raise RuntimeError(f"cropping {self.cropping} removes a whole dim of {tuple(shp)}")
        return out

    def forward(self, x):
        self.out_shape(x.shape)
        return CropFn.apply(x, self._full())


class Cropping2D(_Cropping):
    nd = 2


class Cropping3D(_Cropping):
    nd = 3


class LeakyReLU(Layer):
    def __init__(self, alpha=0.3, name=None, **kw):
        super().__init__(name)
        self.alpha = float(kw.get("negative_slope", alpha))

    def forward(self, x):
        return ActFn.apply(x, S3_ACT_LEAKY, self.alpha)


class Activation(Layer):
    def __init__(self, activation, name=None):
        super().__init__(name)
        if activation not in ops.ACT_CODES:
            raise ValueError(f"unknown activation {activation!r}")
        self.activation = activation

    def forward(self, x):
        code = ops.ACT_CODES[self.activation]
        return x if code == S3_ACT_NONE else ActFn.apply(x, code, KERAS_LEAKY_RELU_SLOPE)


class SkipConnection(Layer):
    """First call caches the tensor and returns it, second call adds the cache and clears it."""

    def __init__(self, name):
        super().__init__(name)
        self._cache = None

    def forward(self, x):
        if self._cache is None:
            self._cache = x
            return x
        cache, self._cache = self._cache, None
        if tuple(cache.shape) != tuple(x.shape):
            raise RuntimeError(f'SkipConnection "{self.name}" shape mismatch: '
                               f"{tuple(x.shape)} vs cached {tuple(cache.shape)}")
        return AddFn.apply(x, cache)


class SpatialExpansion(Layer):
    """``tf.nn.depth_to_space`` (NHWC, DCR) on 4-D tensors."""

    def __init__(self, spatial_mult=1, name=None):
        super().__init__(name)
        self._spatial_mult = int(spatial_mult)

    def out_shape(self, shp):
        r = self._spatial_mult
        if len(shp) != 4:
            raise RuntimeError(f"SpatialExpansion expects a 4-D tensor, got {tuple(shp)}")
        if shp[3] % (r * r):
            raise RuntimeError(f"SpatialExpansion: {shp[3]} channels not divisible by "
                               f"spatial_mult^2 = {r * r}")
        return (shp[0], shp[1] * r, shp[2] * r, shp[3] // (r * r))

    def forward(self, x):
        self.out_shape(x.shape)
        return x if self._spatial_mult == 1 else ExpandFn.apply(x, self._spatial_mult, 1, 0, 0)


class SpatioTemporalExpansion(Layer):
    """Temporal expansion (nearest repeat | depth_to_time + roll) then per-time-slice
    depth_to_space, on 5-D tensors."""

    def __init__(self, spatial_mult=1, temporal_mult=1, temporal_method="nearest", t_roll=0,
                 name=None):
        super().__init__(name)
        self._spatial_mult = int(spatial_mult)
        self._temporal_mult = int(temporal_mult)
        self._temporal_meth = temporal_method
        self._t_roll = int(t_roll)
        if temporal_method not in ("nearest", "depth_to_time"):
            raise ValueError(f"unknown temporal_method {temporal_method!r}")

    @property
    def method_code(self):
        return 1 if self._temporal_meth == "depth_to_time" else 0

    def out_shape(self, shp):
        r, m = self._spatial_mult, self._temporal_mult
        if len(shp) != 5:
            raise RuntimeError(f"SpatioTemporalExpansion expects a 5-D tensor, got {tuple(shp)}")
        c = shp[4]
        if m > 1 and self.method_code == 1:
            if c % m:
                raise RuntimeError(f"depth_to_time: {c} channels not divisible by {m}")
            c //= m
        if c % (r * r):
            raise RuntimeError(f"SpatioTemporalExpansion: {c} channels not divisible by "
                               f"spatial_mult^2 = {r * r}")
        return (shp[0], shp[1] * r, shp[2] * r, shp[3] * m, c // (r * r))

    def forward(self, x):
        self.out_shape(x.shape)
        if self._spatial_mult == 1 and self._temporal_mult == 1:
            return x
        return ExpandFn.apply(x, self._spatial_mult, self._temporal_mult, self.method_code,
                              self._t_roll)


def _match_exo(x, exo, what):
    if exo is None:
        raise RuntimeError(f"{what} needs hi-res exogenous data but got None")
    if tuple(exo.shape[:-1]) != tuple(x.shape[:-1]):
        raise RuntimeError(f"{what}: exogenous data shape {tuple(exo.shape)} does not match the "
                           f"hi-res tensor {tuple(x.shape)}")


class Sup3rAdder(Layer):
    """``x + hi_res_adder``"""

    def forward(self, x, hi_res_adder=None):
        _match_exo(x, hi_res_adder, f'Sup3rAdder "{self.name}"')
        if hi_res_adder.shape[-1] not in (1, x.shape[-1]):
            raise RuntimeError("Sup3rAdder: channel mismatch")
        if hi_res_adder.shape[-1] != x.shape[-1]:
            hi_res_adder = hi_res_adder.expand(*x.shape).contiguous()
        return AddFn.apply(x, hi_res_adder)


class Sup3rConcat(Layer):
    """``tf.concat((x, hi_res_feature), axis=-1)``"""

    def out_shape(self, shp, n_exo=1):
        return (*shp[:-1], shp[-1] + n_exo)

    def forward(self, x, hi_res_feature=None):
        _match_exo(x, hi_res_feature, f'Sup3rConcat "{self.name}"')
        return ConcatFn.apply(x, hi_res_feature)



class _BatchNormFn(torch.autograd.Function):
    """y = x * inv + (beta - mean * inv),  inv = gamma / sqrt(var + eps)  (per channel).  The
    forward and dX are one fused per-channel affine kernel each; the two tiny parameter
    reductions (dgamma, dbeta: C values) are device tensor reductions."""

    @staticmethod
    def forward(ctx, x, gamma, beta, mean, var, eps):
        rstd = torch.rsqrt(var + eps)
        inv = gamma * rstd
        ctx.save_for_backward(x, inv, rstd, mean)
        return ops.channel_affine(x, inv.contiguous(), (beta - mean * inv).contiguous())

    @staticmethod
    def backward(ctx, dy):
        x, inv, rstd, mean = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.channel_affine(dy, inv.contiguous(), None) if ctx.needs_input_grad[0] else None
        dgamma = dbeta = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            c = x.shape[-1]
            d2, x2 = dy.reshape(-1, c), x.reshape(-1, c)
            dbeta = d2.sum(dim=0)
            dgamma = ((x2 - mean) * d2).sum(dim=0) * rstd
        return dx, dgamma, dbeta, None, None, None


class BatchNormalization(Layer):
    """keras ``BatchNormalization`` over the channel axis in INFERENCE mode: the reference calls
    every layer as ``layer(x)`` without a training flag (sup3r/models/abstract.py:1081-1092,
    1157-1165), so the moving statistics are used (and never updated) in ``generate`` and in
    ``_tf_generate`` alike.  gamma / beta are trainable; the moving mean / variance are state."""

    has_weights = True

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, name=None,
                 **_):
        super().__init__(name)
        if axis not in (-1,):
            raise ValueError("BatchNormalization: only the channel axis (-1) is supported")
        self.momentum, self.epsilon = float(momentum), float(epsilon)
        self.center, self.scale = bool(center), bool(scale)
        self.gamma = self.beta = self.moving_mean = self.moving_variance = None

    @property
    def weights(self):
        return [v for v, on in ((self.gamma, self.scale), (self.beta, self.center))
                if v is not None and on]

    def build(self, shp, rng, device):
        c = int(shp[-1])
        if self.gamma is None:
            mk = lambda nm, val: Variable(f"{self.name}/{nm}:0",
                                          torch.full((c,), val, dtype=torch.float32, device=device))
            self.gamma, self.beta = mk("gamma", 1.0), mk("beta", 0.0)
            self.moving_mean, self.moving_variance = mk("moving_mean", 0.0), mk("moving_variance", 1.0)
        elif self.gamma.shape[0] != c:
            raise RuntimeError(f'BatchNormalization "{self.name}" was built for '
                               f"{self.gamma.shape[0]} channels but got {c}")
        self.built = True

    def restore(self, it, device):
        vals = {}
        names = [n for n, on in (("gamma", self.scale), ("beta", self.center)) if on]
        for n in names:
            vals[n] = np.asarray(next(it), dtype=np.float32)
        c = len(next(iter(vals.values()))) if vals else None
        for n, dflt in (("gamma", 1.0), ("beta", 0.0)):
            if n in vals:
                setattr(self, n, Variable(f"{self.name}/{n}:0",
                                          torch.from_numpy(vals[n].copy()).to(device)))
        self._restore_c = c
        self.built = True

    def extra_state(self):
        if self.moving_mean is None:
            return {}
        return {"moving_mean": self.moving_mean.numpy(),
                "moving_variance": self.moving_variance.numpy()}

    def set_extra_state(self, state, device):
        for n in ("moving_mean", "moving_variance"):
            if n in state:
                setattr(self, n, Variable(f"{self.name}/{n}:0", torch.from_numpy(
                    np.asarray(state[n], np.float32).copy()).to(device)))

    def _param(self, v, c, val, device):
        if v is not None:
            return v.value
        return torch.full((c,), val, dtype=torch.float32, device=device)

    def forward(self, x):
        c, dev = x.shape[-1], x.device
        return _BatchNormFn.apply(x.contiguous(), self._param(self.gamma, c, 1.0, dev),
                                  self._param(self.beta, c, 0.0, dev),
                                  self._param(self.moving_mean, c, 0.0, dev),
                                  self._param(self.moving_variance, c, 1.0, dev), self.epsilon)


class _NanFillFn(torch.autograd.Function):
    """where(isnan(obs), fill, obs): gradient flows to ``fill`` at the unobserved locations."""

    @staticmethod
    def forward(ctx, obs, fill):
        m = torch.isnan(obs)
        ctx.save_for_backward(m)
        return torch.where(m, fill, obs)

    @staticmethod
    def backward(ctx, dy):
        (m,) = ctx.saved_tensors
        return None, torch.where(m, dy, torch.zeros_like(dy))


class Sup3rConcatObs(Layer):
    """Concatenate sparse observation data (NaN where unobserved) mid-network.  phygnn's source is
    not vendored in the reference; semantics restated from its call sites
    (sup3r/models/with_obs.py:18-27, abstract.py:1000-1035, tests/conftest.py:129-130): NaNs are
    replaced by the gridded estimate the network already carries -- channel ``fill_index`` of
    ``x`` (default: the obs layer's position among the consecutive obs layers, so ``u_10m_obs``
    after a (u, v) tensor is filled from u and ``v_10m_obs`` from v) -- and the filled field is
    appended as a new channel.  ``include_mask`` also appends the 0 / 1 observed mask.  With no
    observation data the layer is the identity (abstract.py:1004-1013)."""

    def __init__(self, name=None, fill_index=None, include_mask=False, features=None):
        super().__init__(name)
        self.fill_index = fill_index
        self.include_mask = bool(include_mask)
        if features is not None:
            self.features = list(features)

    def out_shape(self, shp, n_exo=1):
        return (*shp[:-1], shp[-1] + n_exo * (2 if self.include_mask else 1))

    def forward(self, x, hi_res_feature=None):
        if hi_res_feature is None:
            return x
        _match_exo(x, hi_res_feature, f'Sup3rConcatObs "{self.name}"')
        n = hi_res_feature.shape[-1]
        i0 = int(self.fill_index or 0)
        if i0 + n > x.shape[-1]:
            raise RuntimeError(f'Sup3rConcatObs "{self.name}": fill channels {i0}:{i0 + n} exceed '
                               f"the {x.shape[-1]} channels of the hi-res tensor")
        crop = [(0, 0)] * (x.dim() - 1) + [(i0, x.shape[-1] - i0 - n)]
        fill = CropFn.apply(x, crop)
        out = ConcatFn.apply(x, _NanFillFn.apply(hi_res_feature, fill).contiguous())
        if self.include_mask:
            out = ConcatFn.apply(out, (~torch.isnan(hi_res_feature)).to(x.dtype))
        return out


class Sup3rObsModel(Layer):
    """Observation-embedding sub-network ``layer(x, hi_res_obs[, extras])`` (call protocol of
    sup3r/models/abstract.py:1026-1035, 1118-1129; ``features`` = observation features,
    ``exo_features`` = gridded extras such as topography).  NaNs of the observations are filled
    from the leading channels of ``x``; [filled obs, observed mask, extras] run through the
    layer's own ``hidden_layers`` and the embedding is concatenated to ``x``.  (phygnn's source is
    not vendored: the internals are this library's restatement of that contract.)"""

    has_weights = True

    def __init__(self, name=None, features=None, exo_features=None, hidden_layers=None,
                 fill_index=0):
        super().__init__(name)
        self.features = list(features) if features is not None else [self.name]
        self.exo_features = list(exo_features or [])
        self.fill_index = int(fill_index)
        self._net = CustomNetwork(hidden_layers or [], name=f"{self.name}_embed")

    @property
    def weights(self):
        return self._net.weights

    def _embed_channels(self, n_obs, n_extra):
        return 2 * n_obs + n_extra

    def out_shape(self, shp, n_exo=None):
        n_obs = len(self.features) if n_exo is None else n_exo
        cin = self._embed_channels(n_obs, len(self.exo_features))
        return (*shp[:-1], shp[-1] + self._net.output_shape((*shp[:-1], cin))[-1])

    def build(self, shp, rng, device):
        cin = self._embed_channels(len(self.features), len(self.exo_features))
        self._net.device = torch.device(device)
        self._net.build((*shp[:-1], cin))
        self.built = True

    def restore(self, it, device):
        self._net.device = torch.device(device)
        self._net._restore_from(it)
        self.built = True

    def forward(self, x, hi_res_obs=None, extras=None):
        if hi_res_obs is None:
            return x
        _match_exo(x, hi_res_obs, f'Sup3rObsModel "{self.name}"')
        n = hi_res_obs.shape[-1]
        crop = [(0, 0)] * (x.dim() - 1) + [(self.fill_index, x.shape[-1] - self.fill_index - n)]
        filled = _NanFillFn.apply(hi_res_obs, CropFn.apply(x, crop)).contiguous()
        z = ConcatFn.apply(filled, (~torch.isnan(hi_res_obs)).to(x.dtype))
        if extras is not None:
            z = ConcatFn.apply(z, extras.contiguous())
        if not self._net.built:
            self._net.build(tuple(z.shape))
        for lyr in self._net.layers:
            z = lyr.forward(z)
        return ConcatFn.apply(x, z.contiguous())


class Flatten(Layer):
    def out_shape(self, shp):
        return (shp[0], int(np.prod(shp[1:])))

    def forward(self, x):
        return x.reshape(x.shape[0], -1)


class Dense(Layer):
    has_weights = True

    def __init__(self, units, activation=None, use_bias=True, name=None, **_):
        super().__init__(name)
        self.units = int(units)
        if activation not in ops.ACT_CODES:
            raise ValueError(f"unknown activation {activation!r}")
        self.activation = activation
        self.use_bias = bool(use_bias)
        self.kernel = None
        self.bias = None

    @property
    def weights(self):
        return [v for v in (self.kernel, self.bias) if v is not None]

    def out_shape(self, shp):
        return (*shp[:-1], self.units)

    def build(self, shp, rng, device):
        k = int(shp[-1])
        if self.kernel is None:
            lim = np.sqrt(6.0 / (k + self.units))
            w = rng.uniform(-lim, lim, size=(k, self.units)).astype(np.float32)
            self.kernel = Variable(f"{self.name}/kernel:0", torch.from_numpy(w).to(device))
            if self.use_bias:
                self.bias = Variable(f"{self.name}/bias:0",
                                     torch.zeros(self.units, dtype=torch.float32, device=device))
        elif self.kernel.shape[0] != k:
            raise RuntimeError(f'Dense "{self.name}" was built for {self.kernel.shape[0]} input '
                               f"features but got {k}")
        self.built = True

    def forward(self, x, act=None, alpha=None):
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        if self.kernel.shape[0] != x2.shape[1]:
            raise RuntimeError(f'Dense "{self.name}" expects {self.kernel.shape[0]} features, got '
                               f"{x2.shape[1]}")
        code = ops.ACT_CODES[self.activation] if act is None else act
        if alpha is None:
            alpha = KERAS_LEAKY_RELU_SLOPE if code == S3_ACT_LEAKY else 0.0
        y = DenseFn.apply(x2, self.kernel.value, self.bias.value if self.bias is not None else None,
                          code, alpha)
        return y.reshape(*lead, self.units)


class Dropout(Layer):
    """Identity (inference semantics; the sup3r configs never use Dropout)."""

    def __init__(self, rate=0.0, name=None, **_):
        super().__init__(name)
        self.rate = float(rate)

    def forward(self, x):
        return x


LAYER_CLASSES = {c.__name__: c for c in (
    FlexiblePadding, Conv2D, Conv3D, Conv2DTranspose, Conv3DTranspose, Cropping2D, Cropping3D,
    LeakyReLU, Activation, SkipConnection, SpatialExpansion, SpatioTemporalExpansion, Sup3rAdder,
    Sup3rConcat, Sup3rConcatObs, Sup3rObsModel, BatchNormalization, Flatten, Dense, Dropout)}

SUP3R_EXO_LAYERS = (Sup3rAdder, Sup3rConcat)
SUP3R_OBS_LAYERS = (Sup3rObsModel, Sup3rConcatObs)   # sup3r/models/utilities.py:23
SUP3R_LAYERS = (*SUP3R_EXO_LAYERS, *SUP3R_OBS_LAYERS)


def layer_features(lyr):
    """Exogenous feature names an exo / obs layer consumes (default: its own name;
    sup3r/models/abstract.py:1001-1002)."""
    return list(getattr(lyr, "features", [lyr.name]))


def layer_exo_channels(lyr, exo_channels):
    ec = exo_channels or {}
    return int(sum(ec.get(f, 1) for f in layer_features(lyr)))


def exo_out_shape(lyr, shp, exo_channels=None):
    """Shape after an exo / obs layer; an obs layer without data (0 channels) is the identity."""
    if isinstance(lyr, Sup3rAdder):
        return tuple(shp)
    n = layer_exo_channels(lyr, exo_channels)
    return tuple(shp) if n == 0 else lyr.out_shape(shp, n)


def run_exo_layer(lyr, x, exo):
    """Gather ``features`` (+ ``exo_features`` extras) of one exo / obs layer from ``exo``
    ({feature name: tensor}) and call it (sup3r/models/abstract.py:1107-1129; obs layers run
    without a missing observation feature, abstract.py:1004-1013)."""
    feats = layer_features(lyr)
    extra_feats = list(getattr(lyr, "exo_features", []))
    is_obs = isinstance(lyr, SUP3R_OBS_LAYERS)
    stack, extras = [], []
    for f in feats + extra_feats:
        if exo.get(f) is None:
            if is_obs and f in feats:
                logger.warning("%s does not match any features in exogenous_data (%s). Will run "
                               "without this observation feature.", f, list(exo))
                continue
            raise RuntimeError(f'exogenous data is missing required feature "{f}"')
        (stack if f in feats else extras).append(to_device_tensor(exo[f], x.device))
    hr = None if not stack else (stack[0] if len(stack) == 1 else torch.cat(stack, dim=-1))
    if extras:
        return lyr.forward(x, hr, extras[0] if len(extras) == 1 else torch.cat(extras, dim=-1))
    return lyr.forward(x, hr)


def expand_hidden_layers(hidden_layers):
    """Expand ``{"n": N, "repeat": [...]}`` blocks (recursively)."""
    out = []
    for cfg in hidden_layers:
        if not isinstance(cfg, dict):
            raise TypeError(f"hidden layer config must be a dict, got {type(cfg)}")
        if "repeat" in cfg:
            inner = expand_hidden_layers(cfg["repeat"])
            for _ in range(int(cfg.get("n", 1))):
                out.extend(copy.deepcopy(inner))
        else:
            out.append(dict(cfg))
    return out


class CustomNetwork:
    """Sequential network built from a phygnn ``hidden_layers`` list."""

    def __init__(self, hidden_layers=None, name=None, device=None):
        self.name = name
        self.hidden_layers = copy.deepcopy(list(hidden_layers or []))
        self.device = torch.device(device) if device is not None else default_device()
        self._layers = []
        self._skips = {}
        self._built_for = None
        self._plans = {}
        used = {}
        for cfg in expand_hidden_layers(self.hidden_layers):
            cfg = dict(cfg)
            cls = cfg.pop("class", None)
            if cls is None:
                if "units" in cfg:  # phygnn shorthand for a dense layer
                    cls = "Dense"
                else:
                    raise KeyError(f'hidden layer config needs a "class" key: {cfg}')
            if cls not in LAYER_CLASSES:
                raise KeyError(f'Could not retrieve layer class "{cls}"; supported: '
                               f"{sorted(LAYER_CLASSES)}")
            if cls == "SkipConnection":
                nm = cfg["name"]
                if nm not in self._skips:
                    self._skips[nm] = SkipConnection(nm)
                self._layers.append(self._skips[nm])
                continue
            if cfg.get("name") is None and cls not in ("Sup3rAdder", "Sup3rConcat", "Sup3rConcatObs",
                                                       "Sup3rObsModel"):
                base = cls.lower()
                i = used.get(base, 0)
                used[base] = i + 1
                prefix = f"{name}/" if name else ""
                cfg["name"] = f"{prefix}{base}" if i == 0 else f"{prefix}{base}_{i}"
            self._layers.append(LAYER_CLASSES[cls](**cfg))

    # ---- protocol ---------------------------------------------------------------
    @staticmethod
    def seed(s=0):
        """Seed weight initialisation (stand-in for ``CustomNetwork.seed``, interface.py:69)."""
        _GLOBAL_SEED[0] = int(s)
        np.random.seed(int(s))
        torch.manual_seed(int(s))

    @property
    def layers(self):
        return self._layers

    def __iter__(self):
        return iter(self._layers)

    def __len__(self):
        return len(self._layers)

    @property
    def weights(self):
        """Trainable variables in keras order (kernel, bias per layer)."""
        out, seen = [], set()
        for lyr in self._layers:
            if id(lyr) in seen:
                continue
            seen.add(id(lyr))
            out.extend(lyr.weights)
        return out

    @property
    def weight_tensors(self):
        return [v.value for v in self.weights]

    def reset_skips(self):
        for s in self._skips.values():
            s._cache = None

    # ---- shapes / build -----------------------------------------------------------
    def output_shape(self, in_shape, exo_channels=None):
        """Propagate a shape through the layer list (raises RuntimeError on a bad shape)."""
        shp = tuple(int(s) for s in in_shape)
        exo_channels = exo_channels or {}
        cache = {}
        for i, lyr in enumerate(self._layers):
            try:
                if isinstance(lyr, SkipConnection):
                    if lyr.name in cache:
                        if cache.pop(lyr.name) != shp:
                            raise RuntimeError(f'SkipConnection "{lyr.name}" shape mismatch')
                    else:
                        cache[lyr.name] = shp
                elif isinstance(lyr, SUP3R_LAYERS):
                    shp = exo_out_shape(lyr, shp, exo_channels)
                else:
                    shp = lyr.out_shape(shp)
            except Exception as e:
                raise RuntimeError(f'Could not run layer #{i} "{lyr}" on tensor of shape {shp}') \
                    from e
        return shp

    def build(self, in_shape, exo_channels=None):
        """Create the weights for an input of ``in_shape`` (idempotent)."""
        shp = tuple(int(s) for s in in_shape)
        exo_channels = exo_channels or {}
        for i, lyr in enumerate(self._layers):
            if lyr.has_weights and not lyr.built:
                rng = np.random.default_rng([_GLOBAL_SEED[0], i])
                lyr.build(shp, rng, self.device)
            elif not lyr.built:
                lyr.built = True
            if isinstance(lyr, SUP3R_LAYERS):
                shp = exo_out_shape(lyr, shp, exo_channels)
            else:
                shp = lyr.out_shape(shp)
        self._built_for = tuple(in_shape[1:])
        return shp

    @property
    def built(self):
        return all(lyr.built for lyr in self._layers if lyr.has_weights)

    # ---- execution ------------------------------------------------------------------
    def forward(self, x, exo=None):
        """Eager, literal layer loop (abstract.py:1081-1092).  ``exo``: {layer name: tensor}."""
        x = to_device_tensor(x, self.device)
        exo = exo or {}
        if not self.built:
            self.build(tuple(x.shape), {k: v.shape[-1] for k, v in exo.items()})
        self.reset_skips()
        try:
            for i, lyr in enumerate(self._layers):
                if isinstance(lyr, SUP3R_LAYERS):
                    x = run_exo_layer(lyr, x, exo)
                else:
                    x = lyr.forward(x)
        except Exception as e:
            self.reset_skips()
            raise RuntimeError(f'Could not run layer #{i} "{lyr}" on tensor of shape '
                               f"{tuple(x.shape)}") from e
        finally:
            self.reset_skips()
        return x

    def predict(self, x, exo=None):
        with torch.no_grad():
            return self.forward(x, exo).as_subclass(DeviceArray)

    # ---- persistence ------------------------------------------------------------------
    def get_weights(self):
        return [v.numpy() for v in self.weights]

    def set_weights(self, arrays):
        ws = self.weights
        if len(arrays) != len(ws):
            raise ValueError(f"expected {len(ws)} weight arrays, got {len(arrays)}")
        for v, a in zip(ws, arrays):
            if tuple(np.shape(a)) != v.shape:
                raise ValueError(f"weight {v.name}: shape {np.shape(a)} != {v.shape}")
            v.assign(a)
        self._plans.clear()

    def save(self, path):
        """Pickle of plain python / numpy objects: config, name and weights."""
        state = {"format": "sup3r_b200.CustomNetwork/1", "name": self.name,
                 "hidden_layers": self.hidden_layers, "built_for": self._built_for,
                 "weight_names": [v.name for v in self.weights], "weights": self.get_weights(),
                 "extra_state": {i: lyr.extra_state() for i, lyr in enumerate(self._layers)
                                 if hasattr(lyr, "extra_state")}}
        with open(path, "wb") as f:
            pickle.dump(state, f)

    @classmethod
    def load(cls, path, device=None):
        with open(path, "rb") as f:
            state = pickle.load(f)
        if not isinstance(state, dict) or "hidden_layers" not in state:
            raise TypeError(f"{path} is not a sup3r_b200 CustomNetwork file (phygnn pickles need "
                            "tools/export_phygnn_weights.py run in a TensorFlow environment)")
        net = cls(state["hidden_layers"], name=state.get("name"), device=device)
        net._restore(state["weights"])
        for i, st in state.get("extra_state", {}).items():
            net._layers[int(i)].set_extra_state(st, net.device)
        return net

    def _restore(self, arrays):
        """Assign saved weights without knowing input shapes (shapes come from the arrays).  A
        network saved before it was built has no weights: it stays unbuilt."""
        if len(arrays) == 0:
            return
        self._restore_from(iter(arrays))

    def _restore_from(self, it):
        seen = set()
        for lyr in self._layers:
            if not lyr.has_weights or id(lyr) in seen:
                continue
            seen.add(id(lyr))
            if hasattr(lyr, "restore"):
                lyr.restore(it, self.device)
                continue
            k = np.asarray(next(it), dtype=np.float32)
            lyr.kernel = Variable(f"{lyr.name}/kernel:0", torch.from_numpy(k.copy()).to(self.device))
            if lyr.use_bias:
                b = np.asarray(next(it), dtype=np.float32)
                lyr.bias = Variable(f"{lyr.name}/bias:0", torch.from_numpy(b.copy()).to(self.device))
            lyr.built = True
        for lyr in self._layers:
            lyr.built = True

    def to_json(self):
        return json.dumps({"hidden_layers": self.hidden_layers, "name": self.name})
