"""torch-CPU restatement of the reference layer sequence.  TEST INFRASTRUCTURE ONLY (see
``oracle/__init__.py``): used (a) as an independent cross-check of ``oracle/layers_ref.py``,
(b) for float64 autograd gradients in the training parity tests, and (c) as the multi-threaded
CPU stand-in for the reference's TensorFlow path in ``bench.py`` (``cpu_baseline`` /
``--impl reference``): it executes the LITERAL op order of sup3r/models/abstract.py:1081-1092
(FlexiblePadding -> Conv -> Cropping -> LeakyReLU -> expansion -> skip, one op at a time,
channels-last tensors, oneDNN convolutions), including the padded-then-cropped work the
reference performs.  Parity unpinned for the conv arithmetic (TF / phygnn not importable)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .layers_ref import expand_hidden_layers, same_pads


def _to_cf(x):
    """channels-last (N, *sp, C) -> channels-first"""
    nd = x.dim() - 2
    return x.permute(0, nd + 1, *range(1, nd + 1))


def _to_cl(x):
    nd = x.dim() - 2
    return x.permute(0, *range(2, nd + 2), 1)


def _act(x, name, alpha=0.2):   # keras string activation: negative_slope 0.2
    if name is None or name == "linear":
        return x
    if name == "relu":
        return F.relu(x)
    if name == "leaky_relu":
        return F.leaky_relu(x, alpha)
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "tanh":
        return torch.tanh(x)
    raise ValueError(name)


def depth_to_space(x, r):
    """NHWC DCR"""
    if r == 1:
        return x
    n, h, w, c = x.shape
    cp = c // (r * r)
    return x.reshape(n, h, w, r, r, cp).permute(0, 1, 3, 2, 4, 5).reshape(n, h * r, w * r, cp)


class TorchRefNet:
    """Executes a ``hidden_layers`` config with torch CPU ops.  Weights are torch tensors in
    keras order (kernel, bias per weighted layer)."""

    def __init__(self, hidden_layers, weights, dtype=torch.float32, requires_grad=False):
        self.cfg = expand_hidden_layers(hidden_layers)
        self.dtype = dtype
        self.weights = [torch.as_tensor(np.asarray(w)).to(dtype).clone()
                        .requires_grad_(requires_grad) for w in weights]
        self.bn_state = []   # [(moving_mean, moving_variance)] per BatchNormalization layer

    def __call__(self, x, exo=None):
        x = torch.as_tensor(x).to(self.dtype) if not isinstance(x, torch.Tensor) else x
        wi = iter(self.weights)
        skips = {}
        exo = exo or {}
        self._bn_i = 0
        for cfg in self.cfg:
            cls = cfg["class"]
            nd = x.dim() - 2
            if cls == "FlexiblePadding":
                pads = cfg["paddings"]
                mode = cfg.get("mode", "REFLECT").upper()
                flat = []
                for lo, hi in reversed(pads[1:-1]):
                    flat += [lo, hi]
                xc = _to_cf(x)
                if mode == "CONSTANT":
                    xc = F.pad(xc, flat)
                elif mode == "REFLECT":
                    xc = F.pad(xc, flat, mode="reflect")
                else:
                    raise ValueError(mode)
                x = _to_cl(xc)
            elif cls in ("Conv2D", "Conv3D", "Conv2DTranspose", "Conv3DTranspose"):
                k = next(wi)
                b = next(wi) if cfg.get("use_bias", True) else None
                s = cfg.get("strides", 1)
                xc = _to_cf(x)
                if "Transpose" in cls:
                    # keras (k..., cout, cin) -> torch conv_transpose weight (cin, cout, k...)
                    w = k.permute(nd + 1, nd, *range(nd))
                    f = F.conv_transpose2d if nd == 2 else F.conv_transpose3d
                    yc = f(xc, w, b, stride=s)
                else:
                    if cfg.get("padding", "valid").lower() == "same":
                        ks = k.shape[:nd]
                        st = (s,) * nd if np.isscalar(s) else tuple(s)
                        flat = []
                        for d in reversed(range(nd)):
                            lo, hi = same_pads(x.shape[1 + d], ks[d], st[d])
                            flat += [lo, hi]
                        xc = F.pad(xc, flat)
                    w = k.permute(nd + 1, nd, *range(nd))
                    f = F.conv2d if nd == 2 else F.conv3d
                    yc = f(xc, w, b, stride=s)
                x = _act(_to_cl(yc), cfg.get("activation"))
            elif cls in ("Cropping2D", "Cropping3D"):
                c = cfg["cropping"]
                c = [(c, c)] * nd if np.isscalar(c) else [(v, v) if np.isscalar(v) else tuple(v)
                                                          for v in c]
                sl = (slice(None),) + tuple(slice(lo, x.shape[1 + d] - hi)
                                            for d, (lo, hi) in enumerate(c))
                x = x[sl]
            elif cls == "LeakyReLU":
                x = F.leaky_relu(x, cfg.get("alpha", 0.3))
            elif cls == "Activation":
                x = _act(x, cfg["activation"])
            elif cls == "SkipConnection":
                name = cfg["name"]
                if name in skips:
                    x = x + skips.pop(name)
                else:
                    skips[name] = x
            elif cls == "SpatialExpansion":
                x = depth_to_space(x, cfg.get("spatial_mult", 1))
            elif cls == "SpatioTemporalExpansion":
                m = cfg.get("temporal_mult", 1)
                r = cfg.get("spatial_mult", 1)
                if m > 1:
                    if cfg.get("temporal_method", "nearest") == "nearest":
                        x = torch.repeat_interleave(x, m, dim=3)
                    else:
                        n, h, w, t, c = x.shape
                        x = x.reshape(n, h, w, t * m, c // m)
                        x = torch.roll(x, cfg.get("t_roll", 0), dims=3)
                if r > 1:
                    x = torch.stack([depth_to_space(x[:, :, :, i], r)
                                     for i in range(x.shape[3])], dim=3)
            elif cls == "Sup3rAdder":
                x = x + torch.as_tensor(exo[cfg["name"]]).to(x.dtype)
            elif cls == "Sup3rConcat":
                x = torch.cat((x, torch.as_tensor(exo[cfg["name"]]).to(x.dtype)), dim=-1)
            elif cls == "Sup3rConcatObs":
                # restated contract (see sup3r_b200/network.py:Sup3rConcatObs): NaNs of the sparse
                # observation field are filled from channel ``fill_index`` of x, then concatenated
                obs = exo.get(cfg["name"])
                if obs is not None:
                    obs = torch.as_tensor(obs).to(x.dtype)
                    i0 = int(cfg.get("fill_index") or 0)
                    fill = x[..., i0:i0 + obs.shape[-1]]
                    nan = torch.isnan(obs)
                    x = torch.cat((x, torch.where(nan, fill, obs)), dim=-1)
                    if cfg.get("include_mask", False):
                        x = torch.cat((x, (~nan).to(x.dtype)), dim=-1)
            elif cls == "BatchNormalization":
                # keras inference mode (the reference never passes training=True); state =
                # (moving_mean, moving_variance) from ``self.bn_state[layer index]``
                g = next(wi) if cfg.get("scale", True) else 1.0
                b = next(wi) if cfg.get("center", True) else 0.0
                mean, var = self.bn_state[self._bn_i]
                self._bn_i += 1
                mean = torch.as_tensor(mean).to(x.dtype)
                var = torch.as_tensor(var).to(x.dtype)
                x = (x - mean) / torch.sqrt(var + cfg.get("epsilon", 1e-3)) * g + b
            elif cls == "Flatten":
                x = x.reshape(x.shape[0], -1)
            elif cls == "Dense":
                k = next(wi)
                b = next(wi) if cfg.get("use_bias", True) else None
                x = x @ k
                if b is not None:
                    x = x + b
                x = _act(x, cfg.get("activation"))
            elif cls == "Dropout":
                pass
            else:
                raise KeyError(cls)
        return x


def disc_loss(out_true, out_gen):
    """Relativistic average discriminator loss (sup3r/models/base.py:505-549)."""
    logits = torch.cat([out_true - out_gen.mean(), out_gen - out_true.mean()], dim=0)
    labels = torch.cat([torch.ones_like(out_true), torch.zeros_like(out_gen)], dim=0)
    return F.binary_cross_entropy_with_logits(logits, labels)
