"""Bias-correction transforms of the forward-pass input hook (sup3r/pipeline/strategy.py:502-517,
sup3r/bias/utilities.py:296-332): the linear family of sup3r/bias/bias_transforms.py --
``global_linear_bc`` (:224-248), ``local_linear_bc`` (:251-348), ``monthly_local_linear_bc``
(:351-487).  The arithmetic is the reference's; the correction factors come from a ``.npz`` file
or a dict of arrays (``"{feature}_scalar"``, ``"{feature}_adder"`` on the full low-res grid)
instead of the reference's h5 files read through rex (a file format that is out of scope here).
They act on the LOW-RES chunk (a few hundred KB) on the host before it is uploaded.
"""
from __future__ import annotations

import logging
from warnings import warn

import numpy as np

logger = logging.getLogger(__name__)


def _factors(feature_name, bias_fp):
    src = np.load(bias_fp) if isinstance(bias_fp, str) else bias_fp
    try:
        return (np.array(src[f"{feature_name}_scalar"], dtype=np.float32),
                np.array(src[f"{feature_name}_adder"], dtype=np.float32))
    except KeyError as e:
        raise RuntimeError(f'Bias correction factors for "{feature_name}" not found in '
                           f"{bias_fp if isinstance(bias_fp, str) else list(src)}") from e


def _clip(out, out_range):
    if out_range is not None:
        out = np.maximum(out, np.min(out_range))
        out = np.minimum(out, np.max(out_range))
    return out


def global_linear_bc(data, scalar, adder, out_range=None):
    """out = data * scalar + adder (bias_transforms.py:224-248)"""
    return _clip(data * scalar + adder, out_range)


def _smooth(arr, smoothing):
    if smoothing > 0:
        from scipy.ndimage import gaussian_filter
        for idt in range(arr.shape[-1]):
            arr[..., idt] = gaussian_filter(arr[..., idt], smoothing, mode="nearest")
    return arr


def local_linear_bc(data, lat_lon, feature_name, bias_fp, lr_padded_slice=None, out_range=None,
                    smoothing=0, threshold=0.1):
    """Site-by-site ``data * scalar + adder``; 3-D factors (monthly in the last axis) are averaged
    (bias_transforms.py:251-348).  ``data``: (s1, s2, t)."""
    scalar, adder = _factors(feature_name, bias_fp)
    if scalar.ndim == 3 and adder.ndim == 3:
        scalar, adder = scalar.mean(axis=-1), adder.mean(axis=-1)
    if lr_padded_slice is not None:
        sl = (lr_padded_slice[0], lr_padded_slice[1])
        scalar, adder = scalar[sl], adder[sl]
    if np.isnan(scalar).any() or np.isnan(adder).any():
        msg = f'Bias correction scalar/adder values had NaNs for "{feature_name}" from: {bias_fp}'
        logger.warning(msg)
        warn(msg)
    scalar = np.repeat(np.expand_dims(scalar, axis=-1), data.shape[-1], axis=-1)
    adder = np.repeat(np.expand_dims(adder, axis=-1), data.shape[-1], axis=-1)
    scalar, adder = _smooth(scalar, smoothing), _smooth(adder, smoothing)
    return _clip(data * scalar + adder, out_range)


def monthly_local_linear_bc(data, lat_lon, feature_name, bias_fp, months=None,
                            lr_padded_slice=None, temporal_avg=True, out_range=None, smoothing=0,
                            scalar_range=None, adder_range=None, threshold=0.1):
    """Site-by-site linear correction with one factor pair per calendar month
    (bias_transforms.py:351-487).  ``months``: 1-based month of every time step of ``data`` (the
    reference derives it from ``date_range_kwargs``)."""
    scalar, adder = _factors(feature_name, bias_fp)
    assert scalar.ndim == 3, "Monthly bias correct needs 3D scalars"
    assert adder.ndim == 3, "Monthly bias correct needs 3D adders"
    if lr_padded_slice is not None:
        sl = (lr_padded_slice[0], lr_padded_slice[1])
        scalar, adder = scalar[sl], adder[sl]
    imonths = np.asarray(months, dtype=int) - 1
    scalar, adder = scalar[..., imonths], adder[..., imonths]
    if temporal_avg:
        scalar = np.repeat(np.expand_dims(scalar.mean(axis=-1), -1), data.shape[-1], axis=-1)
        adder = np.repeat(np.expand_dims(adder.mean(axis=-1), -1), data.shape[-1], axis=-1)
        if len(np.unique(imonths)) > 2:
            msg = ('Bias correction method "monthly_local_linear_bc" was used with temporal '
                   "averaging over a time index with >2 months.")
            warn(msg)
            logger.warning(msg)
    if np.isnan(scalar).any() or np.isnan(adder).any():
        msg = f'Bias correction scalar/adder values had NaNs for "{feature_name}" from: {bias_fp}'
        logger.warning(msg)
        warn(msg)
    scalar, adder = _smooth(scalar, smoothing), _smooth(adder, smoothing)
    if scalar_range is not None:
        scalar = np.maximum(np.minimum(scalar, np.max(scalar_range)), np.min(scalar_range))
    if adder_range is not None:
        adder = np.maximum(np.minimum(adder, np.max(adder_range)), np.min(adder_range))
    return _clip(data * scalar + adder, out_range)


METHODS = {"global_linear_bc": global_linear_bc, "local_linear_bc": local_linear_bc,
           "monthly_local_linear_bc": monthly_local_linear_bc}


def bias_correct_features(data, features, lat_lon, bc_method, bc_kwargs, lr_padded_slice=None):
    """Correct the channels of ``data`` (s1, s2, t, f) named in ``bc_kwargs`` in place
    (bias/utilities.py:296-332).  ``bc_kwargs``: {feature: kwargs of the method}."""
    if bc_method not in METHODS:
        raise KeyError(f'Could not find bias correction method "{bc_method}"; available: '
                       f"{sorted(METHODS)} (the quantile-mapping methods are out of scope)")
    fun = METHODS[bc_method]
    for feat, kw in bc_kwargs.items():
        try:
            i = list(features).index(feat)
            kw = dict(kw)
            if bc_method == "global_linear_bc":
                data[..., i] = fun(data[..., i], **kw)
            else:
                kw.setdefault("lr_padded_slice", lr_padded_slice)
                data[..., i] = fun(data[..., i], lat_lon, feat, **kw)
        except Exception as e:
            msg = (f"Could not run bias correction method {bc_method} on feature {feat} with "
                   f"input of shape {data.shape}. Received error: {e}")
            logger.exception(msg)
            raise RuntimeError(msg) from e
    return data
