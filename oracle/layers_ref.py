"""Numpy restatement of the layer vocabulary the sup3r generator / discriminator
configs use.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``); parity
unpinned for conv arithmetic (phygnn 0.0.33 / TF 2.15.1 are not importable).

The reference drives these layers from ``sup3r/models/abstract.py:1081-1092``
(``generate``) and ``:1157-1165`` (``_tf_generate``); the layer classes
themselves live in ``phygnn.layers.custom_layers`` and ``tf.keras.layers``
(call sites: ``sup3r/models/abstract.py:16-19``, ``sup3r/models/utilities.py:9-17``).
Their published semantics are restated here so that a reviewer holding the
phygnn / keras sources can diff them function by function:

* keras ``Conv2D`` / ``Conv3D``: cross-correlation, kernel ``(*k, Cin, Cout)``,
  channels-last, ``padding='valid'`` default, ``'same'`` = TensorFlow SAME
  (``out = ceil(in / stride)``, total pad ``max((out-1)*stride + k - in, 0)``,
  the odd element goes AFTER), optional fused ``activation``.
* keras ``Conv2DTranspose``: kernel ``(kh, kw, Cout, Cin)``;
  ``out[n, h*s+i, w*s+j, co] += x[n, h, w, ci] * K[i, j, co, ci]``.
* ``tf.pad`` modes CONSTANT / REFLECT (no edge duplication) / SYMMETRIC.
* keras ``Cropping2D/3D``: symmetric int or per-dim (lo, hi) crops.
* keras ``LeakyReLU(alpha)``; ``Activation('relu'|'sigmoid'|'tanh'|...)``.
* ``tf.nn.depth_to_space`` NHWC "DCR":
  ``out[b, h*r+i, w*r+j, c] = in[b, h, w, (i*r + j)*C' + c]``.
* phygnn ``SpatioTemporalExpansion``: temporal first (``'nearest'`` =
  ``tf.repeat(x, m, axis=3)``; ``'depth_to_time'`` = row-major reshape
  ``(B,H,W,T,C) -> (B,H,W,T*m,C/m)`` then ``tf.roll(x, t_roll, axis=3)``), then
  per-time-slice ``depth_to_space``.
* phygnn ``SkipConnection``: first call caches ``x`` and returns it, second
  call returns ``x + cache`` and clears the cache; instances are shared by name.
* phygnn ``Sup3rAdder`` / ``Sup3rConcat``: ``x + exo`` / ``concat((x, exo), -1)``.
* keras ``Flatten`` row-major; ``Dense``: ``x @ W + b``.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------
# primitive ops
# ----------------------------------------------------------------------------
def same_pads(size, k, s):
    """TensorFlow SAME padding for one dimension -> (lo, hi)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def tf_pad(x, paddings, mode="CONSTANT"):
    """``tf.pad`` restatement.  ``paddings``: [[lo, hi], ...] per axis."""
    mode = mode.upper()
    np_mode = {"CONSTANT": "constant", "REFLECT": "reflect", "SYMMETRIC": "symmetric"}[mode]
    paddings = [tuple(int(v) for v in p) for p in paddings]
    if mode == "REFLECT":
        for (lo, hi), n in zip(paddings, x.shape):
            if max(lo, hi) > n - 1:
                raise ValueError(
                    f"REFLECT padding {lo, hi} needs a dimension larger than the pad, got {n}")
    return np.pad(x, paddings, mode=np_mode)


def conv_nd(x, w, b=None, strides=1, padding="valid"):
    """Channels-last N-d cross-correlation.  x: (N, *sp, Cin); w: (*k, Cin, Cout)."""
    nd = w.ndim - 2
    assert x.ndim == nd + 2, (x.shape, w.shape)
    k = w.shape[:nd]
    s = (strides,) * nd if np.isscalar(strides) else tuple(strides)
    if padding.lower() == "same":
        pads = [(0, 0)] + [same_pads(x.shape[1 + d], k[d], s[d]) for d in range(nd)] + [(0, 0)]
        x = np.pad(x, pads)
    elif padding.lower() != "valid":
        raise ValueError(padding)
    out_sp = tuple((x.shape[1 + d] - k[d]) // s[d] + 1 for d in range(nd))
    if min(out_sp) <= 0:
        raise ValueError(f"conv input {x.shape} too small for kernel {k}")
    out = np.zeros((x.shape[0],) + out_sp + (w.shape[-1],), dtype=np.result_type(x, w))
    for tap in np.ndindex(*k):
        sl = (slice(None),) + tuple(
            slice(tap[d], tap[d] + (out_sp[d] - 1) * s[d] + 1, s[d]) for d in range(nd))
        out += x[sl] @ w[tap]
    if b is not None:
        out += b
    return out


def conv_transpose_nd(x, w, b=None, strides=1, padding="valid"):
    """Keras ConvNDTranspose.  x: (N, *sp, Cin); w: (*k, Cout, Cin)."""
    nd = w.ndim - 2
    k = w.shape[:nd]
    s = (strides,) * nd if np.isscalar(strides) else tuple(strides)
    full = tuple((x.shape[1 + d] - 1) * s[d] + k[d] for d in range(nd))
    out = np.zeros((x.shape[0],) + full + (w.shape[-2],), dtype=np.result_type(x, w))
    for tap in np.ndindex(*k):
        sl = (slice(None),) + tuple(
            slice(tap[d], tap[d] + (x.shape[1 + d] - 1) * s[d] + 1, s[d]) for d in range(nd))
        out[sl] += x @ w[tap].T
    if padding.lower() == "same":
        sl = [slice(None)]
        for d in range(nd):
            want = x.shape[1 + d] * s[d]
            lo = (full[d] - want) // 2
            sl.append(slice(lo, lo + want))
        out = out[tuple(sl)]
    if b is not None:
        out = out + b
    return out


def crop_nd(x, cropping):
    nd = x.ndim - 2
    if np.isscalar(cropping):
        cropping = [(cropping, cropping)] * nd
    cropping = [(c, c) if np.isscalar(c) else tuple(c) for c in cropping]
    sl = (slice(None),) + tuple(slice(lo, x.shape[1 + d] - hi) for d, (lo, hi) in
                                enumerate(cropping))
    return x[sl]


def activation(x, name, alpha=None):
    if name is None or name == "linear":
        return x
    if name == "relu":
        return np.maximum(x, 0)
    if name == "leaky_relu":
        a = 0.2 if alpha is None else alpha   # keras.activations.leaky_relu: negative_slope 0.2
        return np.where(x >= 0, x, a * x)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-x))
    if name == "tanh":
        return np.tanh(x)
    if name == "elu":
        return np.where(x > 0, x, np.expm1(x))
    if name == "softplus":
        return np.logaddexp(x, 0)
    raise ValueError(f"unknown activation {name!r}")


def depth_to_space(x, r):
    """NHWC DCR pixel shuffle (``tf.nn.depth_to_space``)."""
    if r == 1:
        return x
    n, h, w, c = x.shape
    assert c % (r * r) == 0, "channels must be divisible by spatial_mult**2"
    cp = c // (r * r)
    y = x.reshape(n, h, w, r, r, cp)
    y = y.transpose(0, 1, 3, 2, 4, 5)
    return y.reshape(n, h * r, w * r, cp)


def spatiotemporal_expansion(x, spatial_mult=1, temporal_mult=1,
                             temporal_method="nearest", t_roll=0):
    if temporal_mult > 1:
        if temporal_method == "nearest":
            x = np.repeat(x, temporal_mult, axis=3)
        elif temporal_method == "depth_to_time":
            n, h, w, t, c = x.shape
            assert c % temporal_mult == 0
            x = x.reshape(n, h, w, t * temporal_mult, c // temporal_mult)
            x = np.roll(x, t_roll, axis=3)
        else:
            raise ValueError(temporal_method)
    if spatial_mult > 1:
        x = np.stack([depth_to_space(x[:, :, :, i], spatial_mult)
                      for i in range(x.shape[3])], axis=3)
    return x


# ----------------------------------------------------------------------------
# layer objects (the literal reference op order is executed one layer at a time)
# ----------------------------------------------------------------------------
class RefLayer:
    weights = ()

    def __init__(self, name=None):
        self.name = name or type(self).__name__

    def __repr__(self):
        return f"<ref {type(self).__name__} {self.name}>"


class FlexiblePadding(RefLayer):
    def __init__(self, paddings, mode="REFLECT", option="tf", name=None):
        super().__init__(name)
        self.paddings = [list(p) for p in paddings]
        self.rank = len(self.paddings)
        self.mode = mode.upper()

    def __call__(self, x):
        return tf_pad(x, self.paddings, self.mode)


class _Conv(RefLayer):
    nd = 2
    transposed = False

    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None,
                 use_bias=True, name=None, **_):
        super().__init__(name)
        self.filters = int(filters)
        ks = (kernel_size,) * self.nd if np.isscalar(kernel_size) else tuple(kernel_size)
        self.kernel_size = tuple(int(k) for k in ks)
        st = (strides,) * self.nd if np.isscalar(strides) else tuple(strides)
        self.strides = tuple(int(s) for s in st)
        self.padding = padding
        self.activation = activation
        self.use_bias = use_bias
        self.kernel = None
        self.bias = None

    def kernel_shape(self, cin):
        if self.transposed:
            return self.kernel_size + (self.filters, cin)
        return self.kernel_size + (cin, self.filters)

    @property
    def weights(self):
        return [w for w in (self.kernel, self.bias) if w is not None]

    def __call__(self, x):
        assert self.kernel is not None, "weights not set"
        f = conv_transpose_nd if self.transposed else conv_nd
        y = f(x, self.kernel.astype(x.dtype, copy=False),
              None if self.bias is None else self.bias.astype(x.dtype, copy=False),
              self.strides, self.padding)
        return activation(y, self.activation)


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


class Cropping2D(RefLayer):
    def __init__(self, cropping, name=None):
        super().__init__(name)
        self.cropping = cropping

    def __call__(self, x):
        return crop_nd(x, self.cropping)


class Cropping3D(Cropping2D):
    pass


class LeakyReLU(RefLayer):
    def __init__(self, alpha=0.3, name=None, **_):
        super().__init__(name)
        self.alpha = float(alpha)

    def __call__(self, x):
        return np.where(x >= 0, x, self.alpha * x).astype(x.dtype, copy=False)


class Activation(RefLayer):
    def __init__(self, activation, name=None):
        super().__init__(name)
        self.activation = activation

    def __call__(self, x):
        return activation(x, self.activation).astype(x.dtype, copy=False)


class SkipConnection(RefLayer):
    def __init__(self, name):
        super().__init__(name)
        self._cache = None

    def __call__(self, x):
        if self._cache is None:
            self._cache = x
            return x
        out = x + self._cache
        self._cache = None
        return out


class SpatialExpansion(RefLayer):
    def __init__(self, spatial_mult=1, name=None):
        super().__init__(name)
        self._spatial_mult = int(spatial_mult)

    def __call__(self, x):
        return depth_to_space(x, self._spatial_mult)


class SpatioTemporalExpansion(RefLayer):
    def __init__(self, spatial_mult=1, temporal_mult=1, temporal_method="nearest", t_roll=0,
                 name=None):
        super().__init__(name)
        self._spatial_mult = int(spatial_mult)
        self._temporal_mult = int(temporal_mult)
        self._temporal_meth = temporal_method
        self._t_roll = int(t_roll)

    def __call__(self, x):
        return spatiotemporal_expansion(x, self._spatial_mult, self._temporal_mult,
                                        self._temporal_meth, self._t_roll)


class Sup3rAdder(RefLayer):
    def __call__(self, x, hi_res_adder):
        return x + hi_res_adder.astype(x.dtype, copy=False)


class Sup3rConcat(RefLayer):
    def __call__(self, x, hi_res_feature):
        return np.concatenate((x, hi_res_feature.astype(x.dtype, copy=False)), axis=-1)


class Flatten(RefLayer):
    def __call__(self, x):
        return x.reshape(x.shape[0], -1)


class Dense(RefLayer):
    def __init__(self, units, activation=None, use_bias=True, name=None, **_):
        super().__init__(name)
        self.units = int(units)
        self.activation = activation
        self.use_bias = use_bias
        self.kernel = None
        self.bias = None

    @property
    def weights(self):
        return [w for w in (self.kernel, self.bias) if w is not None]

    def __call__(self, x):
        y = x @ self.kernel.astype(x.dtype, copy=False)
        if self.bias is not None:
            y = y + self.bias.astype(x.dtype, copy=False)
        return activation(y, self.activation)


class BatchNormalization(RefLayer):
    """Inference-mode keras BatchNormalization over the channel axis."""

    def __init__(self, epsilon=1e-3, name=None, **_):
        super().__init__(name)
        self.epsilon = float(epsilon)
        self.gamma = self.beta = self.moving_mean = self.moving_variance = None

    @property
    def weights(self):
        return [w for w in (self.gamma, self.beta) if w is not None]

    def __call__(self, x):
        inv = self.gamma / np.sqrt(self.moving_variance + self.epsilon)
        return (x * inv + (self.beta - self.moving_mean * inv)).astype(x.dtype, copy=False)


class Dropout(RefLayer):
    def __init__(self, rate=0.0, name=None, **_):
        super().__init__(name)

    def __call__(self, x):
        return x


LAYER_CLASSES = {c.__name__: c for c in (
    FlexiblePadding, Conv2D, Conv3D, Conv2DTranspose, Conv3DTranspose, Cropping2D, Cropping3D,
    LeakyReLU, Activation, SkipConnection, SpatialExpansion, SpatioTemporalExpansion,
    Sup3rAdder, Sup3rConcat, Flatten, Dense, BatchNormalization, Dropout)}
EXO_LAYERS = (Sup3rAdder, Sup3rConcat)


def expand_hidden_layers(hidden_layers):
    """phygnn ``hidden_layers`` dialect: expand ``{"n": N, "repeat": [...]}`` blocks."""
    out = []
    for cfg in hidden_layers:
        if "repeat" in cfg:
            inner = expand_hidden_layers(cfg["repeat"])
            for _ in range(int(cfg.get("n", 1))):
                out.extend(dict(c) for c in inner)
        else:
            out.append(dict(cfg))
    return out


def build_layers(hidden_layers):
    """Instantiate reference layers; SkipConnection instances are shared by name."""
    skips = {}
    layers = []
    for cfg in expand_hidden_layers(hidden_layers):
        cfg = dict(cfg)
        cls = cfg.pop("class")
        if cls == "SkipConnection":
            name = cfg["name"]
            if name not in skips:
                skips[name] = SkipConnection(name)
            layers.append(skips[name])
        else:
            layers.append(LAYER_CLASSES[cls](**cfg))
    return layers


def glorot_uniform(rng, shape, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def out_shape(lyr, shp):
    """Output shape of one layer for input shape ``shp`` (tuple, channels last)."""
    shp = tuple(shp)
    nd = len(shp) - 2
    if isinstance(lyr, FlexiblePadding):
        return tuple(n + lo + hi for n, (lo, hi) in zip(shp, lyr.paddings))
    if isinstance(lyr, _Conv):
        sp = []
        for d in range(nd):
            n, k, s = shp[1 + d], lyr.kernel_size[d], lyr.strides[d]
            if lyr.transposed:
                sp.append(n * s if lyr.padding == "same" else (n - 1) * s + k)
            else:
                sp.append(-(-n // s) if lyr.padding == "same" else (n - k) // s + 1)
        return (shp[0], *sp, lyr.filters)
    if isinstance(lyr, Cropping2D):
        c = lyr.cropping
        c = [(c, c)] * nd if np.isscalar(c) else [(v, v) if np.isscalar(v) else tuple(v) for v in c]
        return (shp[0], *[n - lo - hi for n, (lo, hi) in zip(shp[1:-1], c)], shp[-1])
    if isinstance(lyr, SpatialExpansion):
        r = lyr._spatial_mult
        return (shp[0], shp[1] * r, shp[2] * r, shp[3] // (r * r))
    if isinstance(lyr, SpatioTemporalExpansion):
        r, m = lyr._spatial_mult, lyr._temporal_mult
        c = shp[4] // m if (m > 1 and lyr._temporal_meth == "depth_to_time") else shp[4]
        return (shp[0], shp[1] * r, shp[2] * r, shp[3] * m, c // (r * r))
    if isinstance(lyr, Sup3rConcat):
        return (*shp[:-1], shp[-1] + 1)
    if isinstance(lyr, Flatten):
        return (shp[0], int(np.prod(shp[1:])))
    if isinstance(lyr, Dense):
        return (*shp[:-1], lyr.units)
    return shp


def build_weights(layers, in_shape, seed=0, bias_scale=0.05):
    """Seeded keras-style initialiser (glorot-uniform kernels; biases get a small normal
    draw instead of keras' zeros so that the bias path is exercised)."""
    rng = np.random.default_rng(seed)
    shp = tuple(in_shape)
    for lyr in layers:
        if isinstance(lyr, _Conv) and lyr.kernel is None:
            cin = shp[-1]
            rf = int(np.prod(lyr.kernel_size))
            lyr.kernel = glorot_uniform(rng, lyr.kernel_shape(cin), rf * cin, rf * lyr.filters)
            if lyr.use_bias:
                lyr.bias = (rng.standard_normal(lyr.filters) * bias_scale).astype(np.float32)
        elif isinstance(lyr, Dense) and lyr.kernel is None:
            lyr.kernel = glorot_uniform(rng, (shp[-1], lyr.units), shp[-1], lyr.units)
            if lyr.use_bias:
                lyr.bias = (rng.standard_normal(lyr.units) * bias_scale).astype(np.float32)
        shp = out_shape(lyr, shp)
    return shp


def get_weights(layers):
    """Keras order: kernel, bias per weighted layer (shared skips hold none)."""
    return [w for lyr in layers for w in lyr.weights]


def set_weights(layers, weights):
    it = iter(weights)
    for lyr in layers:
        if isinstance(lyr, (_Conv, Dense)):
            lyr.kernel = np.asarray(next(it))
            lyr.bias = np.asarray(next(it)) if lyr.use_bias else None
        elif isinstance(lyr, BatchNormalization):
            lyr.gamma, lyr.beta = np.asarray(next(it)), np.asarray(next(it))


def run_layers(layers, x, exo=None):
    """Literal layer loop of ``abstract.py:1081-1092``.  ``exo``: {layer name: array}."""
    reset_skips(layers)
    for lyr in layers:
        if isinstance(lyr, EXO_LAYERS):
            x = lyr(x, exo[lyr.name])
        else:
            x = lyr(x)
    return x


def reset_skips(layers):
    for lyr in layers:
        if isinstance(lyr, SkipConnection):
            lyr._cache = None
