"""Bias-correction transforms of the forward-pass input hook (sup3r/pipeline/strategy.py:502-517,
sup3r/bias/utilities.py:296-332): the linear family of sup3r/bias/bias_transforms.py --
``global_linear_bc`` (:224-248), ``local_linear_bc`` (:251-348), ``monthly_local_linear_bc``
(:351-487).  The arithmetic is the reference's; the correction factors come from a ``.npz`` file
or a dict of arrays (``"{feature}_scalar"``, ``"{feature}_adder"`` on the full low-res grid)
instead of the reference's h5 files read through rex (a file format that is out of scope here).
They act on the LOW-RES chunk (a few hundred KB) on the host before it is uploaded.

``local_qdm_bc`` (:622-824) -- empirical quantile delta mapping -- runs on the device
(``s3_qdm_bc``: three table interpolations per value, HBM-bound).  Its core lives in a
third-party package the reference imports (``rex.utilities.bc_utils.QuantileDeltaMapping``,
NREL-rex >= 0.2.91, absent here): the kernel restates the published algorithm (Cannon et al.
2015, eq. 3-6) and is anchored on the known answers of the reference's own tests
(tests/bias/test_qdm_bias_correction.py: identity, +-10 offsets, no_trend); PARITY WITH rex IS
UNPINNED.  ``local_presrat_bc`` (:958-1137) adds its zero-rate / K-factor step to the same
kernel.  Parametric distributions (scipy.stats) and the CALCULATION of the tables / factors
(sup3r/bias/qdm.py, presrat.py) are out of scope.
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


def make_time_index_from_kws(date_range_kwargs):
    """``pd.date_range(**date_range_kwargs)``; the extra key ``drop_leap`` removes every
    29 February (sup3r/preprocessing/utilities.py:222-244; the caller's dict is left alone)."""
    import pandas as pd
    kw = dict(date_range_kwargs)
    drop_leap = kw.pop("drop_leap", False)
    time_index = pd.date_range(**kw)
    if drop_leap:
        time_index = time_index[~((time_index.month == 2) & (time_index.day == 29))]
    return time_index


def monthly_local_linear_bc(data, lat_lon, feature_name, bias_fp, date_range_kwargs=None,
                            lr_padded_slice=None, temporal_avg=True, out_range=None, smoothing=0,
                            scalar_range=None, adder_range=None, threshold=0.1, months=None):
    """Site-by-site linear correction with one factor pair per calendar month
    (bias_transforms.py:351-487).  The month of every time step of ``data`` comes from
    ``date_range_kwargs`` (-> ``pd.date_range``, like the reference) or, when the forward-pass
    hook knows the chunk's time index, from ``months`` (1-based)."""
    if months is None:
        assert date_range_kwargs is not None, (
            "monthly_local_linear_bc needs date_range_kwargs (or the months of the time steps)")
        months = make_time_index_from_kws(date_range_kwargs).month.values
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


def sample_q(n_samples, sampling="linear", log_base=10):
    """Quantile levels of an empirical CDF table (rex.utilities.bc_utils sample_q_linear /
    sample_q_log / sample_q_invlog): even spacing, or concentrated near 0 / near 1."""
    if sampling == "linear":
        return np.linspace(0, 1, n_samples)
    log_q = (np.logspace(0, 1, n_samples, base=log_base) - 1) / (log_base - 1)
    if sampling == "log":
        return log_q
    if sampling == "invlog":
        return 1 - log_q[::-1]
    raise KeyError(f'sampling option must be linear, log or invlog, got "{sampling}"')


def _qdm_params(base_dset, feature_name, bias_fp, presrat=False):
    src = np.load(bias_fp, allow_pickle=False) if isinstance(bias_fp, str) else bias_fp
    names = {"base": f"base_{base_dset}_params", "bias": f"bias_{feature_name}_params",
             "bias_fut": f"bias_fut_{feature_name}_params"}
    if presrat:     # bias_transforms.py:940-946
        names.update(bias_tau_fut=f"{feature_name}_tau_fut", k_factor=f"{feature_name}_k_factor")
    out = {}
    for k, n in names.items():
        if n in src:
            out[k] = np.asarray(src[n], dtype=np.float32)
        elif k != "bias_fut":
            raise RuntimeError(f'Bias correction parameters "{n}" not found in '
                               f"{bias_fp if isinstance(bias_fp, str) else list(src)}")
    cfg = {"time_window_center": np.asarray(src["time_window_center"], dtype=np.float64)}
    defaults = [("dist", "empirical"), ("sampling", "linear"), ("log_base", 10)]
    if presrat:
        defaults.append(("zero_rate_threshold", None))
    for k, default in defaults:
        v = src[k] if k in src else default
        cfg[k] = v.item() if isinstance(v, np.ndarray) else v
    if presrat and cfg["zero_rate_threshold"] is None:
        raise RuntimeError('PresRat parameters need the "zero_rate_threshold" attribute')
    out["cfg"] = cfg
    return out


def _run_qdm(data, params, day_of_year, date_range_kwargs, lr_padded_slice, relative, no_trend,
             delta_denom_min, delta_denom_zero, delta_range, out_range, k_range=None):
    """Shared driver of ``local_qdm_bc`` / ``local_presrat_bc``: window of every time step, tables
    of the chunk, one launch of ``s3_qdm_bc``.  Returns (out (s1, s2, t), [n_nonfinite, n_nan])."""
    import torch
    from . import ops
    data = np.asarray(data, dtype=np.float32)
    assert data.ndim == 3, f"data was expected to be a 3D array but got shape {data.shape}"
    if day_of_year is None:
        day_of_year = make_time_index_from_kws(date_range_kwargs).day_of_year
    day_of_year = np.asarray(day_of_year)
    assert data.shape[-1] == day_of_year.size, (
        f"Time should align with data 3rd dimension but got data {data.shape} and time_index "
        f"length {day_of_year.size}")
    cfg = params["cfg"]
    if cfg["dist"] != "empirical":
        raise NotImplementedError(f'QDM with dist="{cfg["dist"]}": only empirical CDFs run here')
    base, bias = params["base"], params["bias"]
    bias_fut = params.get("bias_fut")
    tau, k_factor = params.get("bias_tau_fut"), params.get("k_factor")
    if k_factor is not None and k_range is not None:
        k_factor = np.minimum(np.maximum(k_factor, np.min(k_range)), np.max(k_range))
    if lr_padded_slice is not None:
        sl = (lr_padded_slice[0], lr_padded_slice[1])
        base, bias = base[sl], bias[sl]
        bias_fut = None if bias_fut is None else bias_fut[sl]
        # (the reference leaves tau_fut / k_factor unsliced -- bias_transforms.py:1085-1090 --
        #  which only broadcasts when the slice is the whole grid; here they follow the chunk)
        tau = None if tau is None else tau[sl]
        k_factor = None if k_factor is None else k_factor[sl]
    if no_trend or bias_fut is None:
        bias_fut = bias             # (rex: params_mf defaults to params_mh)
    if no_trend:
        tau = k_factor = None       # QDM only (bias_transforms.py:1114-1116)
    window = np.array([np.argmin(abs(d - cfg["time_window_center"])) for d in day_of_year],
                      dtype=np.int32)
    q = sample_q(base.shape[-1], cfg["sampling"], cfg["log_base"])
    if not torch.cuda.is_available():
        raise RuntimeError("quantile delta mapping runs on the CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    shape = data.shape

    def t(a, dt=torch.float32):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)

    def flat(a, dt=torch.float32):
        return None if a is None else t(a.reshape(-1, *a.shape[2:]), dt)
    out, bad = ops.qdm_bc(flat(data), t(window, torch.int32), flat(base), flat(bias),
                          flat(bias_fut), t(q, torch.float64), relative=relative,
                          delta_denom_zero=delta_denom_zero, delta_denom_min=delta_denom_min,
                          delta_range=delta_range, out_range=out_range,
                          tau_fut=None if tau is None else flat(tau).reshape(-1),
                          k_factor=flat(k_factor, torch.float64))
    return out.cpu().numpy().reshape(shape), bad.tolist()


def local_qdm_bc(data, lat_lon, base_dset, feature_name, bias_fp, date_range_kwargs=None,
                 lr_padded_slice=None, threshold=0.1, relative=True, no_trend=False,
                 delta_denom_min=None, delta_denom_zero=None, delta_range=None, out_range=None,
                 max_workers=1, day_of_year=None):
    """Quantile delta mapping of one feature of a low-res chunk (bias_transforms.py:622-824).
    ``data``: (s1, s2, t); the distribution tables ``base_{base_dset}_params``,
    ``bias_{feature}_params``, ``bias_fut_{feature}_params`` (s1, s2, n_windows, n_quantiles) and
    ``time_window_center`` come from ``bias_fp`` (.npz or dict).  Every time step uses the window
    whose centre is closest to its day of year (``date_range_kwargs`` -> ``pd.date_range``, or
    ``day_of_year`` directly)."""
    params = _qdm_params(base_dset, feature_name, bias_fp)
    out, bad = _run_qdm(data, params, day_of_year, date_range_kwargs, lr_padded_slice, relative,
                        no_trend, delta_denom_min, delta_denom_zero, delta_range, out_range)
    if bad[0]:
        msg = ("QDM bias correction resulted in NaN / inf values! If this is a relative QDM, you "
               "may try setting ``delta_denom_min`` or ``delta_denom_zero``")
        logger.error(msg)
        raise RuntimeError(msg)
    return out


def local_presrat_bc(data, lat_lon, base_dset, feature_name, bias_fp, date_range_kwargs=None,
                     lr_padded_slice=None, threshold=0.1, relative=True, no_trend=False,
                     delta_denom_min=None, delta_denom_zero=None, delta_range=None,
                     k_range=None, out_range=None, max_workers=1, day_of_year=None):
    """PresRat (Pierce et al. 2015) of one feature of a low-res chunk
    (bias_transforms.py:958-1137): QDM with ``delta_denom_min`` defaulting to the file's
    ``zero_rate_threshold``, then results below ``{feature}_tau_fut`` become dry (0) and the
    others are scaled by the window's ``{feature}_k_factor`` (optionally clipped to ``k_range``)
    -- both steps inside the same kernel.  Raises on NaN results only (the reference's check)."""
    params = _qdm_params(base_dset, feature_name, bias_fp, presrat=True)
    delta_denom_min = delta_denom_min or params["cfg"]["zero_rate_threshold"]
    out, bad = _run_qdm(data, params, day_of_year, date_range_kwargs, lr_padded_slice, relative,
                        no_trend, delta_denom_min, delta_denom_zero, delta_range, out_range,
                        k_range=k_range)
    if bad[1]:
        msg = ("Presrat bias correction resulted in NaN values! If this is a relative QDM, you "
               "may try setting ``delta_denom_min`` or ``delta_denom_zero``")
        logger.error(msg)
        raise RuntimeError(msg)
    return out


METHODS = {"global_linear_bc": global_linear_bc, "local_linear_bc": local_linear_bc,
           "monthly_local_linear_bc": monthly_local_linear_bc, "local_qdm_bc": local_qdm_bc,
           "local_presrat_bc": local_presrat_bc}


def bias_correct_features(data, features, lat_lon, bc_method, bc_kwargs, lr_padded_slice=None,
                          time_index=None):
    """Correct the channels of ``data`` (s1, s2, t, f) named in ``bc_kwargs`` in place
    (bias/utilities.py:215-332).  ``bc_kwargs``: {feature: kwargs of the method}.  A datetime
    ``time_index`` of the chunk supplies what the reference passes as ``date_range_kwargs``:
    the day of year of every step (``local_qdm_bc``) / its month (``monthly_local_linear_bc``)."""
    from inspect import signature
    if bc_method not in METHODS:
        raise KeyError(f'Could not find bias correction method "{bc_method}"; available: '
                       f"{sorted(METHODS)}")
    fun = METHODS[bc_method]
    pars = signature(fun).parameters
    dates = None
    if time_index is not None and np.issubdtype(np.asarray(time_index).dtype, np.datetime64):
        import pandas as pd
        dates = pd.DatetimeIndex(np.asarray(time_index))
    for feat, kw in bc_kwargs.items():
        try:
            i = list(features).index(feat)
            kw = dict(kw)
            if bc_method == "global_linear_bc":
                data[..., i] = fun(data[..., i], **kw)
            else:
                kw.setdefault("lr_padded_slice", lr_padded_slice)
                kw.setdefault("feature_name", feat)
                if dates is not None:
                    if "day_of_year" in pars and "date_range_kwargs" not in kw:
                        kw.setdefault("day_of_year", np.asarray(dates.day_of_year))
                    if "months" in pars and "date_range_kwargs" not in kw:
                        kw.setdefault("months", np.asarray(dates.month))
                data[..., i] = fun(data[..., i], lat_lon, **kw)
        except Exception as e:
            msg = (f"Could not run bias correction method {bc_method} on feature {feat} with "
                   f"input of shape {data.shape}. Received error: {e}")
            logger.exception(msg)
            raise RuntimeError(msg) from e
    return data
