"""Content losses of the Sup3rGan hot path on device tensors (sup3r/utilities/loss_metrics.py).

``MeanSquaredError`` / ``MeanAbsoluteError`` stand in for ``tf.keras.losses`` (default loss,
sup3r/models/base.py:30; lookup order of sup3r/models/abstract.py:520-541: this module first).
Each loss is a callable ``loss(x1, x2) -> 0-d device tensor``, argument order (generated, true)
as in sup3r/models/base.py:478-503, differentiable w.r.t. ``x1``.

The pointwise reductions (mean squared / absolute difference) run in this library's fused
value + gradient kernel (``s3_content_loss``).  The structured losses first map both tensors
through small differentiable tensor transforms (central differences, block means, min / max
over axes, FFT magnitude) expressed with device tensor ops, then call the same pointwise
kernels -- so they mirror the reference classes term by term:

=========================  ==============================================================
``ExpLoss``                loss_metrics.py:98-118
``MmdLoss``                loss_metrics.py:62-95, 121-147
``MaterialDerivativeLoss`` loss_metrics.py:150-225
``SpatialDerivativeLoss``  loss_metrics.py:228-260
``TemporalDerivativeLoss`` loss_metrics.py:263-294
``CoarseMseLoss``          loss_metrics.py:297-322
``SpatialExtremesLoss``    loss_metrics.py:325-357
``TemporalExtremesLoss``   loss_metrics.py:360-392
``SpatialFftLoss``         loss_metrics.py:395-437
``SpatiotemporalFftLoss``  loss_metrics.py:440-485
``LowResLoss``             loss_metrics.py:488-638
``SlicedWassersteinLoss``  loss_metrics.py:724-793
=========================  ==============================================================

Out of scope: ``PerceptualLoss`` (loss_metrics.py:641-721: feature maps of a VGG16 with ImageNet
weights, which cannot be obtained here).
"""
from __future__ import annotations

import numpy as np
import torch

from .autograd import ContentLossFn


def _as_device_tensor(x, like=None):
    if isinstance(x, torch.Tensor):
        return x
    dev = like.device if isinstance(like, torch.Tensor) else torch.device("cuda")
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device=dev)


class _ElementwiseMean:
    kind = 0

    def __init__(self, **kwargs):
        self.kwargs = kwargs

    def __call__(self, x1, x2):
        x1 = _as_device_tensor(x1, x2)
        x2 = _as_device_tensor(x2, x1)
        if tuple(x1.shape) != tuple(x2.shape):
            raise RuntimeError(f"loss inputs must have the same shape, got {tuple(x1.shape)} "
                               f"and {tuple(x2.shape)}")
        return ContentLossFn.apply(x1.contiguous(), x2.contiguous(), x1.shape[-1], self.kind)


class MeanSquaredError(_ElementwiseMean):
    """mean((x1 - x2)^2) over every element."""
    kind = 0


class MeanAbsoluteError(_ElementwiseMean):
    """mean(|x1 - x2|) over every element."""
    kind = 1


def _derivative(x, axis=1):
    """Central differences matching ``np.gradient`` (one-sided at the ends) along axis 1, 2
    (spatial) or 3 (temporal) of an ``(n_obs, s1, s2, t[, f])`` tensor (loss_metrics.py:12-59)."""
    if axis not in (1, 2, 3):
        raise ValueError(f"_derivative received axis={axis}. This is meant to compute only "
                         "temporal (axis=3) or spatial (axis=1/2) derivatives for tensors of "
                         "shape (n_obs, spatial_1, spatial_2, temporal)")
    n = x.shape[axis]
    first = x.narrow(axis, 1, 1) - x.narrow(axis, 0, 1)
    mid = (x.narrow(axis, 2, n - 2) - x.narrow(axis, 0, n - 2)) / 2
    last = x.narrow(axis, n - 1, 1) - x.narrow(axis, n - 2, 1)
    return torch.cat([first, mid, last], dim=axis)


def gaussian_kernel(x1, x2, sigma=1.0):
    """exp(-0.5 * sum_f (x1[i] - x2[j])^2 / sigma^2) for every pair of observations (i, j)."""
    d = x1.unsqueeze(1) - x2
    return torch.exp(-0.5 * (d * d).sum(dim=-1) / sigma ** 2)


class _Base:
    def __init__(self, **kwargs):
        self.kwargs = kwargs

    @staticmethod
    def _pair(x1, x2):
        x1 = _as_device_tensor(x1, x2)
        x2 = _as_device_tensor(x2, x1)
        return x1, x2


class ExpLoss(_Base):
    """mean(1 - exp(-(x1 - x2)^2))"""

    def __call__(self, x1, x2):
        x1, x2 = self._pair(x1, x2)
        return torch.mean(1 - torch.exp(-((x1 - x2) ** 2)))


class MmdLoss(_Base):
    """Maximum mean discrepancy with a gaussian kernel over the observation axis."""

    def __call__(self, x1, x2, sigma=1.0):
        x1, x2 = self._pair(x1, x2)
        mmd = torch.mean(gaussian_kernel(x1, x1, sigma))
        mmd = mmd + torch.mean(gaussian_kernel(x2, x2, sigma))
        mmd = mmd - torch.mean(2 * gaussian_kernel(x1, x2, sigma))
        return mmd


class MaterialDerivativeLoss(_Base):
    """MAE between the material derivatives Df/Dt = df/dt + u df/dx + v df/dy of the u / v wind
    pairs (features 2k, 2k + 1) of both tensors."""

    LOSS_METRIC = MeanAbsoluteError()

    def _compute_md(self, x, fidx):
        x = _as_device_tensor(x)
        uidx = 2 * (fidx // 2)
        vidx = 2 * (fidx // 2) + 1
        f = x[..., fidx]
        x_div = _derivative(f, axis=3)
        x_div = x_div + x[..., uidx] * _derivative(f, axis=1)
        x_div = x_div + x[..., vidx] * _derivative(f, axis=2)
        return x_div

    def __call__(self, x1, x2):
        x1, x2 = self._pair(x1, x2)
        hub_heights = x1.shape[-1] // 2
        msg = (f"The {self.__class__.__name__} is meant to be used on spatiotemporal data only. "
               "Received tensor(s) that are not 5D")
        assert len(x1.shape) == 5 and len(x2.shape) == 5, msg
        x1_div = torch.stack([self._compute_md(x1, fidx=i) for i in range(0, 2 * hub_heights, 2)])
        x2_div = torch.stack([self._compute_md(x2, fidx=i) for i in range(0, 2 * hub_heights, 2)])
        return self.LOSS_METRIC(x1_div, x2_div)


class SpatialDerivativeLoss(_Base):
    """MAE between d/ds1 + d/ds2 of both tensors."""

    LOSS_METRIC = MeanAbsoluteError()

    def __call__(self, x1, x2):
        x1, x2 = self._pair(x1, x2)
        msg = (f"The {self.__class__.__name__} is meant to be used on spatial or spatiotemporal "
               "data only. Received tensor(s) that are not at least 4D")
        assert len(x1.shape) >= 4 and len(x2.shape) >= 4, msg
        x1_div = _derivative(x1, axis=1) + _derivative(x1, axis=2)
        x2_div = _derivative(x2, axis=1) + _derivative(x2, axis=2)
        return self.LOSS_METRIC(x1_div, x2_div)


class TemporalDerivativeLoss(_Base):
    """MAE between d/dt of both tensors."""

    LOSS_METRIC = MeanAbsoluteError()

    def __call__(self, x1, x2):
        x1, x2 = self._pair(x1, x2)
        msg = (f"The {self.__class__.__name__} is meant to be used on spatiotemporal data only. "
               "Received tensor(s) that are not 5D")
        assert len(x1.shape) == 5 and len(x2.shape) == 5, msg
        return self.LOSS_METRIC(_derivative(x1, axis=3), _derivative(x2, axis=3))


class CoarseMseLoss(_Base):
    """MSE of the spatial means (axes 1, 2)."""

    MSE_LOSS = MeanSquaredError()

    def __call__(self, x1, x2):
        x1, x2 = self._pair(x1, x2)
        return self.MSE_LOSS(x1.mean(dim=(1, 2)), x2.mean(dim=(1, 2)))


class SpatialExtremesLoss(_Base):
    """(MAE of spatial minima + MAE of spatial maxima) / 2."""

    MAE_LOSS = MeanAbsoluteError()

    def __call__(self, x1, x2):
        x1, x2 = self._pair(x1, x2)
        mae_min = self.MAE_LOSS(x1.amin(dim=(1, 2)), x2.amin(dim=(1, 2)))
        mae_max = self.MAE_LOSS(x1.amax(dim=(1, 2)), x2.amax(dim=(1, 2)))
        return (mae_min + mae_max) / 2


class TemporalExtremesLoss(_Base):
    """(MAE of temporal minima + MAE of temporal maxima) / 2."""

    MAE_LOSS = MeanAbsoluteError()

    def __call__(self, x1, x2):
        x1, x2 = self._pair(x1, x2)
        mae_min = self.MAE_LOSS(x1.amin(dim=3), x2.amin(dim=3))
        mae_max = self.MAE_LOSS(x1.amax(dim=3), x2.amax(dim=3))
        return (mae_min + mae_max) / 2


class SpatialFftLoss(_Base):
    """MAE between log(1 + k0^2 k1^2 |FFT2(x)|) of both (n, s1, s2, f) tensors."""

    MAE_LOSS = MeanAbsoluteError()

    @staticmethod
    def _freq_weights(x):
        k0 = torch.arange(x.shape[1], device=x.device, dtype=x.dtype) ** 2
        k1 = torch.arange(x.shape[2], device=x.device, dtype=x.dtype) ** 2
        return (k0[:, None] * k1[None, :])[None, ..., None]

    def _fft(self, x):
        x_hat = torch.fft.fft2(x.to(torch.complex64), dim=(1, 2)).abs().to(x.dtype)
        return torch.log(1 + self._freq_weights(x) * x_hat)

    def __call__(self, x1, x2):
        x1, x2 = self._pair(x1, x2)
        return self.MAE_LOSS(self._fft(x1), self._fft(x2))


class SpatiotemporalFftLoss(_Base):
    """MAE between log(1 + k0^2 k1^2 f^2 |FFT3(x)|) of both (n, s1, s2, t, f) tensors."""

    MAE_LOSS = MeanAbsoluteError()

    @staticmethod
    def _freq_weights(x):
        k0 = torch.arange(x.shape[1], device=x.device, dtype=x.dtype) ** 2
        k1 = torch.arange(x.shape[2], device=x.device, dtype=x.dtype) ** 2
        f = torch.arange(x.shape[3], device=x.device, dtype=x.dtype) ** 2
        return (k0[:, None, None] * k1[None, :, None] * f[None, None, :])[None, ..., None]

    def _fft(self, x):
        x_hat = torch.fft.fftn(x.to(torch.complex64), dim=(1, 2, 3)).abs().to(x.dtype)
        return torch.log(1 + self._freq_weights(x) * x_hat)

    def __call__(self, x1, x2):
        x1, x2 = self._pair(x1, x2)
        return self.MAE_LOSS(self._fft(x1), self._fft(x2))


class LowResLoss(_Base):
    """Content loss on re-coarsened fields: spatial block means (``s_enhance``), temporal block
    means or subsampling (``t_enhance``, ``t_method``), then ``tf_loss`` on the low-res pair, plus
    an optional extremes term on the high-res pair."""

    EX_LOSS_METRICS = {"SpatialExtremesLoss": SpatialExtremesLoss,
                       "TemporalExtremesLoss": TemporalExtremesLoss}

    def __init__(self, s_enhance=1, t_enhance=1, t_method="average", tf_loss="MeanSquaredError",
                 ex_loss=None):
        super().__init__()
        self._s_enhance = s_enhance
        self._t_enhance = t_enhance
        self._t_method = str(t_method).casefold()
        if tf_loss not in ("MeanSquaredError", "MeanAbsoluteError"):
            raise AttributeError(f"module 'keras.losses' has no attribute {tf_loss!r} that this "
                                 "library implements (MeanSquaredError, MeanAbsoluteError)")
        self._tf_loss = LOSSES[tf_loss]()
        self._ex_loss = ex_loss
        if self._ex_loss is not None:
            self._ex_loss = self.EX_LOSS_METRICS[self._ex_loss]()

    def _s_coarsen_4d_tensor(self, t):
        s = self._s_enhance
        n, a, b, f = t.shape
        return t.reshape(n, a // s, s, b // s, s, f).sum(dim=(2, 4)) / s ** 2

    def _s_coarsen_5d_tensor(self, t):
        s = self._s_enhance
        n, a, b, tt, f = t.shape
        return t.reshape(n, a // s, s, b // s, s, tt, f).sum(dim=(2, 4)) / s ** 2

    def _t_coarsen_sample(self, t):
        assert len(t.shape) == 5
        return t[:, :, :, ::self._t_enhance, :]

    def _t_coarsen_avg(self, t):
        assert len(t.shape) == 5
        n, a, b, _, f = t.shape
        return t.reshape(n, a, b, -1, self._t_enhance, f).sum(dim=4) / self._t_enhance

    def __call__(self, x1, x2):
        x1, x2 = self._pair(x1, x2)
        assert x1.shape == x2.shape
        s_only = len(x1.shape) == 4
        ex_loss = 0.0
        if self._ex_loss is not None:
            ex_loss = self._ex_loss(x1, x2)
        if self._s_enhance > 1 and s_only:
            x1, x2 = self._s_coarsen_4d_tensor(x1), self._s_coarsen_4d_tensor(x2)
        elif self._s_enhance > 1 and not s_only:
            x1, x2 = self._s_coarsen_5d_tensor(x1), self._s_coarsen_5d_tensor(x2)
        if self._t_enhance > 1 and self._t_method == "average":
            x1, x2 = self._t_coarsen_avg(x1), self._t_coarsen_avg(x2)
        if self._t_enhance > 1 and self._t_method == "subsample":
            x1, x2 = self._t_coarsen_sample(x1), self._t_coarsen_sample(x2)
        return self._tf_loss(x1, x2) + ex_loss


class SlicedWassersteinLoss(_Base):
    """Sliced Wasserstein distance over random 1-D projections of the flattened space-time
    axes (loss_metrics.py:724-793): ``proj (P, HWT) @ x (B, HWT, C) -> (B, P, C)``, both sides
    sorted along axis 1, mean squared difference.  The projections are drawn on the device
    (unit rows of a standard normal matrix) at every call, as the reference does."""

    def __init__(self, n_projections=1024, **kwargs):
        super().__init__(**kwargs)
        self._n_projections = int(n_projections)

    def projections(self, n_points, device):
        proj = torch.randn((self._n_projections, n_points), device=device, dtype=torch.float32)
        return proj / proj.norm(dim=-1, keepdim=True).clamp_min(1e-12)

    def __call__(self, x1, x2):
        x1, x2 = self._pair(x1, x2)
        if x1.dim() not in (4, 5) or x2.dim() != x1.dim():
            raise AssertionError("The SlicedWassersteinLoss is meant to be used on spatial or "
                                 "spatiotemporal data only. Received tensor(s) that are not 4D "
                                 "or 5D")
        b, c = x1.shape[0], x1.shape[-1]
        x1_flat, x2_flat = x1.reshape(b, -1, c), x2.reshape(b, -1, c)
        proj = self.projections(x1_flat.shape[1], x1.device)
        x1_sorted = torch.sort(proj @ x1_flat, dim=1).values
        x2_sorted = torch.sort(proj @ x2_flat, dim=1).values
        return torch.mean((x1_sorted - x2_sorted) ** 2)


LOSSES = {"MeanSquaredError": MeanSquaredError, "MeanAbsoluteError": MeanAbsoluteError,
          "ExpLoss": ExpLoss, "MmdLoss": MmdLoss, "MaterialDerivativeLoss": MaterialDerivativeLoss,
          "SpatialDerivativeLoss": SpatialDerivativeLoss,
          "TemporalDerivativeLoss": TemporalDerivativeLoss, "CoarseMseLoss": CoarseMseLoss,
          "SpatialExtremesLoss": SpatialExtremesLoss, "TemporalExtremesLoss": TemporalExtremesLoss,
          "SpatialFftLoss": SpatialFftLoss, "SpatiotemporalFftLoss": SpatiotemporalFftLoss,
          "LowResLoss": LowResLoss, "SlicedWassersteinLoss": SlicedWassersteinLoss}


def get_loss_class(name):
    return LOSSES.get(name)
