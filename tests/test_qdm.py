"""Quantile delta mapping bias correction (sup3r/bias/bias_transforms.py:490-824).

The oracle (oracle/qdm_ref.py) restates the published algorithm -- rex, which holds the
reference's implementation, is absent -- and is anchored here on the known answers of the
reference's own tests (tests/bias/test_qdm_bias_correction.py:331-452).  The CUDA kernel
(``s3_qdm_bc``) must equal the oracle bit for bit (double-precision np.interp restated)."""
import numpy as np
import pytest

from oracle import qdm_ref as R

RNG = np.random.default_rng(2024)


def make_params(s1=5, s2=4, n_win=3, n_q=51, zeros=False, sampling="linear", log_base=10):
    """Sorted quantile tables of three different gamma-like distributions per site / window."""
    q = R.sample_q(n_q, sampling, log_base)

    def table(scale, shift):
        raw = np.sort(RNG.gamma(2.0, scale, (s1, s2, n_win, 400)), axis=-1) + shift
        t = np.quantile(raw, q, axis=-1)
        t = np.moveaxis(t, 0, -1)
        if zeros:                      # a dry / night-time site: many repeated zeros
            t = np.maximum(t - np.median(t, axis=-1, keepdims=True), 0)
        return t.astype(np.float32)
    return {"base": table(90.0, 5.0), "bias": table(110.0, 0.0), "bias_fut": table(120.0, 2.0),
            "cfg": {"time_window_center": np.array([60.0, 182.0, 300.0]), "sampling": sampling,
                    "log_base": log_base, "dist": "empirical"}}


def make_data(params, n_t=40):
    lo, hi = params["bias_fut"].min(), params["bias_fut"].max()
    s1, s2 = params["base"].shape[:2]
    data = RNG.uniform(lo - 5, hi + 5, (s1, s2, n_t)).astype(np.float32)
    doy = RNG.integers(1, 366, n_t)
    return data, doy


# ---------------------------------------------------------------- oracle vs the reference's tests
def test_oracle_identity_when_distributions_are_identical():
    """test_bc_identity / test_bc_identity_absolute (:331-378)"""
    p = make_params()
    p["base"] = p["bias"] = p["bias_fut"]
    data, doy = make_data(p)
    inside = (data >= p["bias_fut"].min(axis=(2, 3))[..., None]) & \
        (data <= p["bias_fut"].max(axis=(2, 3))[..., None])
    for relative in (True, False):
        out = R.local_qdm_bc(data, p, doy, relative=relative)
        ok = inside & (np.abs(data) > 1e-3)
        assert np.allclose(out[ok], data[ok], rtol=1e-5)


@pytest.mark.parametrize("which,expected", [("base", -10.0), ("bias", 10.0), ("both", 0.0)])
def test_oracle_offsets_propagate(which, expected):
    """test_bc_model_constant, test_bc_trend, test_bc_trend_same_hist (:381-452): absolute QDM
    moves the data by the offset between the observed and the modeled-historical tables."""
    p = make_params()
    fut = p["bias_fut"]
    p["base"] = fut - 10 if which in ("base", "both") else fut
    p["bias"] = fut - 10 if which in ("bias", "both") else fut
    data, doy = make_data(p)
    # (inside the table range of every window the value may fall in)
    lo = fut.min(axis=-1).max(axis=-1)[..., None]
    hi = fut.max(axis=-1).min(axis=-1)[..., None]
    ok = (data > lo) & (data < hi)
    out = R.local_qdm_bc(data, p, doy, relative=False)
    assert ok.sum() > 100 and np.allclose(out[ok] - data[ok], expected, atol=2e-3)


def test_oracle_no_trend_equals_future_set_to_historical():
    """test_qdm_transform_notrend (:266-312)"""
    p = make_params()
    data, doy = make_data(p)
    a = R.local_qdm_bc(data, p, doy, no_trend=True)
    q = dict(p, bias_fut=p["bias"])
    b = R.local_qdm_bc(data, q, doy)
    c = R.local_qdm_bc(data, {k: v for k, v in p.items() if k != "bias_fut"}, doy)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert not np.allclose(a, R.local_qdm_bc(data, p, doy))


def test_oracle_quantile_sampling():
    for s in ("linear", "log", "invlog"):
        q = R.sample_q(21, s, 10)
        assert np.isclose(q[0], 0) and np.isclose(q[-1], 1) and (np.diff(q) > 0).all()
    assert np.median(R.sample_q(21, "log")) < 0.5 < np.median(R.sample_q(21, "invlog"))
    assert np.allclose(R.sample_q(21, "log") + R.sample_q(21, "invlog")[::-1], 1)


def test_host_side_rejects_what_does_not_run_here():
    from sup3r_b200 import bias
    assert "local_qdm_bc" in bias.METHODS and "local_presrat_bc" in bias.METHODS
    assert np.array_equal(bias.sample_q(11, "invlog", 7), R.sample_q(11, "invlog", 7))
    with pytest.raises(KeyError):
        bias.bias_correct_features(np.zeros((2, 2, 2, 1), np.float32), ["u"], None,
                                   "no_such_bc", {"u": {}})


def make_presrat(p, thr=0.5):
    s1, s2, n_win = p["base"].shape[:3]
    p = dict(p)
    p["bias_tau_fut"] = RNG.uniform(5, 60, (s1, s2, 1)).astype(np.float32)
    p["k_factor"] = RNG.uniform(0.7, 1.4, (s1, s2, n_win)).astype(np.float32)
    p["cfg"] = dict(p["cfg"], zero_rate_threshold=thr)
    return p


def test_oracle_presrat_identities():
    """tests/bias/test_presrat_bias_correction.py:633-736: identical CDFs + zero tau + K = 1 ->
    no change; zero tau -> no wet value turned dry; PresRat has at least as many values below
    tau as QDM; and PresRat == where(QDM < tau, 0, QDM * K) with the threshold as denominator
    floor (bias_transforms.py:1076, 1117-1120)."""
    p = make_presrat(make_params())
    data, doy = make_data(p)
    pr = R.local_qdm_bc(data, p, doy, presrat=True)
    qd = R.local_qdm_bc(data, p, doy, delta_denom_min=0.5)
    k = p["k_factor"][:, :, [int(np.argmin(abs(d - p["cfg"]["time_window_center"]))) for d in doy]]
    clear = np.abs(qd - p["bias_tau_fut"]) > 1e-3     # (qd is the fp32-rounded QDM result)
    want = np.where(qd < p["bias_tau_fut"], 0, qd.astype(np.float64) * k)
    assert np.allclose(pr[clear], want[clear], rtol=1e-6, atol=0)
    assert (pr < p["bias_tau_fut"]).sum() >= (qd < p["bias_tau_fut"]).sum() > 0
    nz = dict(p, bias_tau_fut=p["bias_tau_fut"] * 0)
    out = R.local_qdm_bc(data, nz, doy, presrat=True)
    assert not ((data > 0) & (out == 0)).any() and not np.allclose(out, data)
    same = dict(nz, base=p["bias_fut"], bias=p["bias_fut"], k_factor=p["k_factor"] * 0 + 1,
                cfg=dict(p["cfg"], zero_rate_threshold=0))
    out = R.local_qdm_bc(data, same, doy, presrat=True)
    inside = (data >= p["bias_fut"].min(axis=(2, 3))[..., None]) & \
        (data <= p["bias_fut"].max(axis=(2, 3))[..., None]) & (np.abs(data) > 1e-3)
    assert np.allclose(out[inside], data[inside], rtol=1e-5)
    # no_trend: QDM only (bias_transforms.py:1114-1116)
    assert np.array_equal(R.local_qdm_bc(data, p, doy, presrat=True, no_trend=True),
                          R.local_qdm_bc(data, p, doy, no_trend=True, delta_denom_min=0.5))


# ---------------------------------------------------------------------------- CUDA kernel (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("case", [dict(), dict(relative=False, k_range=(0.9, 1.1)),
                                  dict(no_trend=True), dict(out_range=(0.0, 400.0),
                                                            delta_range=(0.2, 3.0))],
                         ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()) or "default")
def test_cuda_presrat_equals_oracle_bit_for_bit(cuda, case):
    from sup3r_b200 import bias
    p = make_presrat(make_params())
    data, doy = make_data(p, n_t=45)
    src = {"base_ghi_params": p["base"], "bias_rsds_params": p["bias"],
           "bias_fut_rsds_params": p["bias_fut"], "rsds_tau_fut": p["bias_tau_fut"],
           "rsds_k_factor": p["k_factor"], **p["cfg"]}
    sl = (slice(0, 3), slice(1, 4))
    for lrps in (None, sl):
        d = data if lrps is None else data[sl]
        want = R.local_qdm_bc(d, p, doy, lr_padded_slice=lrps, presrat=True, **case)
        got = bias.local_presrat_bc(d, None, "ghi", "rsds", src, day_of_year=doy,
                                    lr_padded_slice=lrps, **case)
        assert np.array_equal(got, want), np.abs(got - want).max()
        assert (got == 0).any() or case.get("no_trend")
    bad = data.copy()
    bad[0, 0, 0] = np.nan
    with pytest.raises(RuntimeError, match="NaN values"):
        bias.local_presrat_bc(bad, None, "ghi", "rsds", src, day_of_year=doy)
    with pytest.raises(RuntimeError, match="zero_rate_threshold"):
        bias.local_presrat_bc(data, None, "ghi", "rsds",
                              {k: v for k, v in src.items() if k != "zero_rate_threshold"},
                              day_of_year=doy)

CASES = [
    dict(relative=True),
    dict(relative=False),
    dict(relative=True, no_trend=True),
    dict(relative=True, delta_denom_min=15.0, delta_range=(0.5, 1.8)),
    dict(relative=True, delta_denom_zero=1e-3, zeros=True, delta_range=(0.0, 4.0)),
    dict(relative=False, delta_range=(-20.0, 35.0), out_range=(0.0, 900.0)),
    dict(relative=True, sampling="log", log_base=10),
    dict(relative=False, sampling="invlog", log_base=4),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_cuda_qdm_equals_oracle_bit_for_bit(cuda, case):
    from sup3r_b200 import bias
    case = dict(case)
    p = make_params(zeros=case.pop("zeros", False), sampling=case.pop("sampling", "linear"),
                    log_base=case.pop("log_base", 10))
    data, doy = make_data(p, n_t=57)
    if case.get("delta_denom_zero") is None and case.get("relative") and \
            (p["bias"] == 0).any():
        pytest.skip("zeros in the denominator")
    src = {"base_ghi_params": p["base"], "bias_rsds_params": p["bias"],
           "bias_fut_rsds_params": p["bias_fut"], **p["cfg"]}
    sl = (slice(1, 4), slice(0, 3))
    for lrps in (None, sl):
        d = data if lrps is None else data[sl]
        want = R.local_qdm_bc(d, p, doy, lr_padded_slice=lrps, **case)
        got = bias.local_qdm_bc(d, None, "ghi", "rsds", src, day_of_year=doy,
                                lr_padded_slice=lrps, **case)
        assert got.dtype == np.float32 and got.shape == d.shape
        assert np.array_equal(got, want), np.abs(got - want).max()


@pytest.mark.gpu
def test_cuda_qdm_through_the_forward_pass_hook(cuda, tmp_path):
    """bias_correct_features: datetime time index -> day of year; .npz factor file; NaN / inf
    results raise as in the reference (bias_transforms.py:817-823)."""
    import pandas as pd
    from sup3r_b200 import bias
    p = make_params()
    data, _ = make_data(p, n_t=30)
    ti = pd.date_range("2015-02-20", periods=30, freq="5D")
    fp = str(tmp_path / "qdm.npz")
    np.savez(fp, base_ghi_params=p["base"], bias_rsds_params=p["bias"],
             bias_fut_rsds_params=p["bias_fut"], time_window_center=p["cfg"]["time_window_center"],
             sampling="linear", log_base=10, dist="empirical")
    chunk = np.stack([data, data * 2], axis=-1)
    kw = {"rsds": dict(base_dset="ghi", bias_fp=fp, relative=False)}
    out = bias.bias_correct_features(chunk.copy(), ["rsds", "other"], None, "local_qdm_bc", kw,
                                     time_index=ti.values)
    want = R.local_qdm_bc(data, p, np.asarray(ti.day_of_year), relative=False)
    assert np.array_equal(out[..., 0], want) and np.array_equal(out[..., 1], chunk[..., 1])
    # same through date_range_kwargs (what the reference passes)
    got = bias.local_qdm_bc(data, None, "ghi", "rsds", fp, relative=False,
                            date_range_kwargs=dict(start="2015-02-20", periods=30, freq="5D"))
    assert np.array_equal(got, want)
    bad = chunk.copy()
    bad[1, 2, 3, 0] = np.nan
    with pytest.raises(RuntimeError, match="NaN / inf"):
        bias.bias_correct_features(bad, ["rsds", "other"], None, "local_qdm_bc", kw,
                                   time_index=ti.values)
    with pytest.raises(RuntimeError, match="only empirical"):
        bias.local_qdm_bc(data, None, "ghi", "rsds",
                          {"base_ghi_params": p["base"], "bias_rsds_params": p["bias"],
                           "time_window_center": p["cfg"]["time_window_center"],
                           "dist": "weibull_min"}, day_of_year=np.ones(30))
