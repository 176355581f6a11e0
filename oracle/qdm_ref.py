"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

Numpy restatement of the empirical quantile delta mapping the reference applies to a low-res
chunk: sup3r/bias/bias_transforms.py:490-619 (``_apply_qdm``) and :622-824 (``local_qdm_bc``).
The core of ``_apply_qdm`` is ``rex.utilities.bc_utils.QuantileDeltaMapping`` -- a third-party
dependency (NREL-rex >= 0.2.91, pyproject.toml:29) that is NOT vendored in /root/reference and not
installed here, so its source could not be read or run: PARITY WITH rex IS UNPINNED.  What is
restated is the published algorithm (Cannon, Sobie & Murdock 2015, eq. 3-6) with empirical CDFs
evaluated by ``np.interp`` over the sampled quantile levels, and the options the reference
forwards (relative / absolute, no_trend, delta_denom_zero / delta_denom_min, delta_range).  It is
anchored on the known answers of the reference's own tests
(tests/bias/test_qdm_bias_correction.py:331-452: identical distributions -> identity; a -10 offset
of the observed distribution -> -10; a -10 offset of the modeled-historical one -> +10; no_trend ==
future := modeled-historical), see tests/test_qdm.py.
"""
import numpy as np


def sample_q(n_samples, sampling="linear", log_base=10):
    """rex.utilities.bc_utils sample_q_linear / sample_q_log / sample_q_invlog (published
    formulas: even spacing; (base^u - 1) / (base - 1) for u in linspace(0, 1); its mirror)."""
    if sampling == "linear":
        return np.linspace(0, 1, n_samples)
    log_q = (np.logspace(0, 1, n_samples, base=log_base) - 1) / (log_base - 1)
    return log_q if sampling == "log" else 1 - log_q[::-1]


def qdm_empirical(arr, params_oh, params_mh, params_mf=None, sampling="linear", log_base=10,
                  relative=True, delta_denom_min=None, delta_denom_zero=None, delta_range=None):
    """arr (time, space); params_* (space, N).  Cannon et al. 2015:
    tau = F_mf(x) (eq. 3), delta = x / F_mh^-1(tau) (eq. 4) or x - F_mh^-1(tau) (eq. 5),
    x_bc = F_oh^-1(tau) * delta (eq. 6) or + delta."""
    arr = np.asarray(arr, np.float64)
    if params_mf is None:
        params_mf = params_mh
    q = sample_q(params_oh.shape[-1], sampling, log_base)
    out = np.empty_like(arr)
    for s in range(arr.shape[1]):
        x = arr[:, s]
        tau = np.interp(x, np.asarray(params_mf[s], np.float64), q)
        x_oh = np.interp(tau, q, np.asarray(params_oh[s], np.float64))
        x_mh = np.interp(tau, q, np.asarray(params_mh[s], np.float64))
        if relative:
            if delta_denom_zero is not None:
                x_mh[x_mh == 0] = delta_denom_zero
            if delta_denom_min is not None:
                x_mh = np.maximum(x_mh, delta_denom_min)
            with np.errstate(divide="ignore", invalid="ignore"):
                delta = x / x_mh
            if delta_range is not None:
                delta = np.minimum(np.maximum(delta, np.min(delta_range)), np.max(delta_range))
            out[:, s] = x_oh * delta
        else:
            delta = x - x_mh
            if delta_range is not None:
                delta = np.minimum(np.maximum(delta, np.min(delta_range)), np.max(delta_range))
            out[:, s] = x_oh + delta
    return out


def apply_qdm(subset, base_params, bias_params, bias_fut_params, sampling="linear", log_base=10,
              relative=True, no_trend=False, **delta_kw):
    """bias_transforms.py:490-619: (s1, s2, t) subset, (s1, s2, N) tables -> (s1, s2, t)."""
    mf = None if no_trend else np.reshape(bias_fut_params, (-1, bias_fut_params.shape[-1]))
    tmp = np.reshape(subset, (-1, subset.shape[-1])).T
    tmp = qdm_empirical(tmp, np.reshape(base_params, (-1, base_params.shape[-1])),
                        np.reshape(bias_params, (-1, bias_params.shape[-1])), mf, sampling,
                        log_base, relative, **delta_kw)
    return np.reshape(tmp.T, subset.shape)


def local_qdm_bc(data, params, day_of_year, lr_padded_slice=None, relative=True, no_trend=False,
                 delta_denom_min=None, delta_denom_zero=None, delta_range=None, out_range=None,
                 presrat=False, k_range=None):
    """bias_transforms.py:752-824 on a dict of tables ``base`` / ``bias`` / ``bias_fut``
    (s1, s2, n_windows, N) + ``cfg`` (time_window_center, sampling, log_base).  ``presrat``:
    local_presrat_bc, bias_transforms.py:1063-1137 (+ ``bias_tau_fut`` (s1, s2, 1), ``k_factor``
    (s1, s2, n_windows), ``cfg['zero_rate_threshold']``); unlike the reference, tau_fut and
    k_factor follow ``lr_padded_slice`` (the reference leaves them on the full grid, which only
    broadcasts for the full slice)."""
    assert data.ndim == 3
    cfg = params["cfg"]
    base, bias, fut = params["base"], params["bias"], params.get("bias_fut")
    tau = k_factor = None
    if presrat:
        tau, k_factor = np.asarray(params["bias_tau_fut"]), params["k_factor"]
        delta_denom_min = delta_denom_min or cfg["zero_rate_threshold"]
        if k_range is not None:
            k_factor = np.minimum(np.maximum(k_factor, np.min(k_range)), np.max(k_range))
    if lr_padded_slice is not None:
        sl = (lr_padded_slice[0], lr_padded_slice[1])
        base, bias = base[sl], bias[sl]
        fut = None if fut is None else fut[sl]
        if presrat:
            tau, k_factor = tau[sl], k_factor[sl]
    out = np.full(data.shape, np.nan, dtype=data.dtype)
    closest = np.array([np.argmin(abs(d - cfg["time_window_center"])) for d in day_of_year])
    for nt in set(closest):
        idx = closest == nt
        mf = None if fut is None else fut[:, :, nt]
        subset = apply_qdm(
            data[:, :, idx], base[:, :, nt], bias[:, :, nt], mf,
            sampling=cfg.get("sampling", "linear"), log_base=cfg.get("log_base", 10),
            relative=relative, no_trend=no_trend or mf is None, delta_denom_min=delta_denom_min,
            delta_denom_zero=delta_denom_zero, delta_range=delta_range)
        if presrat and not no_trend:
            subset = np.where(subset < tau, 0, subset * k_factor[:, :, nt:nt + 1])
        out[:, :, idx] = subset
    if out_range is not None:
        out = np.maximum(out, np.min(out_range))
        out = np.minimum(out, np.max(out_range))
    if presrat and np.isnan(out).any():
        raise RuntimeError("Presrat bias correction resulted in NaN values!")
    if not presrat and not np.isfinite(out).all():
        raise RuntimeError("QDM bias correction resulted in NaN / inf values!")
    return out
