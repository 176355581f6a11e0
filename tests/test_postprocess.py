"""Chunk post-processing and batch production.

CPU part (``-m "not gpu"``): the numpy oracle ``oracle/postprocess_ref.py`` against golden vectors
produced by the REAL reference functions (``tools/make_golden_postprocess.py`` ->
``tests/golden/postprocess.npz``): bit-exact for limits / coarsening, exact float32 for u/v.
GPU part: the CUDA kernels (through the C ABI) against the same golden vectors."""
import os

import numpy as np
import pytest

from oracle import postprocess_ref as P

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "postprocess.npz"))
METHODS = ("subsample", "average", "total", "max", "min")


def test_oracle_invert_uv_matches_reference():
    for tag in ("desc", "asc"):
        ll = G["lat_lon"] if tag == "desc" else G["lat_lon"][::-1].copy()
        ws, wd = P.invert_uv(G["u"], G["v"], ll)
        assert np.array_equal(ws, G[f"ws_{tag}"]) and np.array_equal(wd, G[f"wd_{tag}"])


def test_oracle_limits_match_reference():
    feats = list(G["lim_features"])
    assert np.array_equal(P.enforce_limits(feats, G["lim_in"]), G["lim_clip"])
    assert np.array_equal(P.enforce_limits(feats, G["lim_in"], nn_fill=True), G["lim_nn"])
    with pytest.raises(KeyError):
        P.enforce_limits(["not_a_feature"], G["lim_in"][..., :1])
    assert P.get_renamed_features(["u_100m", "t_2m", "v_100m", "u_10m"]) == \
        ["windspeed_100m", "t_2m", "winddirection_100m", "u_10m"]


def test_oracle_batch_transform_matches_reference():
    for m in METHODS:
        assert np.array_equal(P.batch_transform(G["hr"], 4, 2, ["u", "v"], None, None, m),
                              G[f"low_{m}"])
        assert np.array_equal(P.batch_transform(G["hr"], 4, 2, ["u", "v"], 0.7, ["v"], m),
                              G[f"low_{m}_smooth"])
    assert np.array_equal(P.batch_transform(G["hr4"], 2, 1, ["u", "v"], 1.3, []), G["low4_smooth"])


def test_host_helpers_match_oracle():
    from sup3r_b200.pipeline import postprocess as pp
    assert pp.OUTPUT_LIMITS == P.OUTPUT_LIMITS
    feats = ["u_100m", "v_100m", "temperature_2m", "u_10m", "v_10m", "pressure_1000pa"]
    assert pp.get_renamed_features(feats) == P.get_renamed_features(feats)
    assert [p[:2] for p in pp.uv_pairs(feats)] == [(0, 1), (3, 4)]
    assert [pp.get_feature_basename(f) for f in feats] == [P.get_feature_basename(f) for f in feats]
    for ll in (G["lat_lon"], G["lat_lon"][::-1].copy()):
        cs = pp.grid_rotation(ll)
        flip = ll[-1, 0, 0] > ll[0, 0, 0]
        th = P.grid_angle(ll[::-1] if flip else ll)
        th = th[::-1] if flip else th
        assert np.allclose(cs[..., 0], np.cos(th)) and np.allclose(cs[..., 1], np.sin(th))


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["desc", "asc"])
def test_cuda_output_transform_matches_reference(cuda, tag):
    import torch
    from sup3r_b200.pipeline.postprocess import transform_output
    ll = G["lat_lon"] if tag == "desc" else G["lat_lon"][::-1].copy()
    feats = ["u_100m", "temperature_2m", "v_100m"]
    temp = np.random.default_rng(0).uniform(-250, 150, G["u"].shape).astype(np.float32)
    data = np.stack([G["u"], temp, G["v"]], axis=-1)
    dev = torch.as_tensor(data, device=cuda).contiguous()
    with pytest.warns(UserWarning):
        out, names = transform_output(dev, feats, ll, invert_uv=True, nn_fill=False)
    assert names == ["windspeed_100m", "temperature_2m", "winddirection_100m"]
    out = out.cpu().numpy()
    assert np.abs(out[..., 0] - G[f"ws_{tag}"]).max() < 1e-4
    dwd = np.abs(out[..., 2] - G[f"wd_{tag}"])
    assert np.minimum(dwd, 360 - dwd).max() < 1e-3
    assert np.array_equal(out[..., 1], np.clip(temp, -200, 100))
    # without invert_uv only the limits act
    dev = torch.as_tensor(data, device=cuda).contiguous()
    out, names = transform_output(dev, feats, ll, invert_uv=False)
    assert names == feats and np.array_equal(out.cpu().numpy()[..., 0], G["u"])


@pytest.mark.gpu
def test_cuda_limits_clip_and_nn_fill_match_reference(cuda):
    import torch
    from sup3r_b200.pipeline.postprocess import transform_output
    feats = list(G["lim_features"])
    for nn, key in ((False, "lim_clip"), (True, "lim_nn")):
        dev = torch.as_tensor(G["lim_in"], device=cuda).contiguous()
        with pytest.warns(UserWarning):
            out, _ = transform_output(dev, feats, None, invert_uv=True, nn_fill=nn)
        assert np.array_equal(out.cpu().numpy(), G[key]), key
    with pytest.raises(KeyError):
        transform_output(torch.zeros((2, 2, 2, 1), device=cuda), ["not_a_feature"], None)


@pytest.mark.gpu
def test_cuda_batch_production_matches_reference(cuda):
    import torch
    from sup3r_b200 import batch, ops
    hr = torch.as_tensor(G["hr"], device=cuda)
    for m in METHODS:
        low = batch.transform(hr, 4, 2, ["u", "v"], temporal_coarsening_method=m)
        assert np.allclose(low.cpu().numpy(), G[f"low_{m}"], rtol=1e-6, atol=1e-6), m
        lows = batch.transform(hr, 4, 2, ["u", "v"], smoothing=0.7, smoothing_ignore=["v"],
                               temporal_coarsening_method=m)
        assert np.allclose(lows.cpu().numpy(), G[f"low_{m}_smooth"], rtol=1e-5, atol=1e-6), m
    low4 = batch.transform(torch.as_tensor(G["hr4"], device=cuda), 2, 1, ["u", "v"], smoothing=1.3)
    assert np.allclose(low4.cpu().numpy(), G["low4_smooth"], rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        ops.coarsen(hr, 4, 2, "median")
    # sampler: gathered crops are exact copies
    data = torch.as_tensor(np.random.default_rng(1).standard_normal((9, 11, 13, 3)).astype(np.float32),
                           device=cuda)
    org = torch.tensor([[0, 0, 0], [3, 5, 7], [5, 6, 9]], dtype=torch.int32, device=cuda)
    got = ops.gather_samples(data, org, (4, 5, 4)).cpu().numpy()
    d = data.cpu().numpy()
    for b, (i, j, k) in enumerate(org.cpu().numpy()):
        assert np.array_equal(got[b], d[i:i + 4, j:j + 5, k:k + 4])


@pytest.mark.gpu
def test_device_batch_handler_trains_a_gan(cuda):
    """The on-device batch handler satisfies Sup3rGan.train's contract (base.py:728-733)."""
    import tempfile
    from sup3r_b200 import configs as C
    from sup3r_b200.batch import DeviceBatchHandler
    from sup3r_b200.models import Sup3rGan
    rng = np.random.default_rng(0)
    base = rng.standard_normal((8, 8, 10, 2))
    data = np.repeat(np.repeat(np.repeat(base, 4, 0), 4, 1), 4, 2).astype(np.float32) * 3 + 1
    bh = DeviceBatchHandler(data, ["u", "v"], sample_shape=(12, 12, 8), batch_size=4, n_batches=4,
                            s_enhance=2, t_enhance=2, smoothing=0.6,
                            temporal_coarsening_method="average")
    assert bh.shapes == ((4, 6, 6, 4, 2), (4, 12, 12, 8, 2))
    b = next(iter(bh))
    assert tuple(b.low_res.shape) == (4, 6, 6, 4, 2) and tuple(b.high_res.shape) == (4, 12, 12, 8, 2)
    assert b.low_res.is_cuda and abs(float(bh.data.mean())) < 1e-3
    Sup3rGan.seed(0)
    m = Sup3rGan(C.spatiotemporal_generator(2, 2, (2,), n_blocks=1, filters=16),
                 C.discriminator(3, "same", (8,)), learning_rate=2e-3)
    with tempfile.TemporaryDirectory() as td:
        m.train(bh, {"spatial": "8km", "temporal": "60min"}, n_epoch=3, weight_gen_advers=0.0,
                train_gen=True, train_disc=False, out_dir=os.path.join(td, "gan_{epoch}"))
    tl = m.history["train_loss_gen"].values
    assert bh.stopped and np.isfinite(tl).all() and tl[-1] < tl[0]
    assert m.meta["smoothing"] == 0.6 and m.meta["lr_features"] == ["u", "v"]
    assert abs(m.means["u"] - float(data[..., 0].mean())) < 1e-3


# ------------------------------------------------------------------------------ bias correction
def test_bias_transforms_match_reference(tmp_path):
    """sup3r_b200.bias vs golden outputs of the reference function bodies (tests/golden/bias.npz)."""
    import warnings
    from sup3r_b200 import bias
    g = np.load(os.path.join(HERE, "golden", "bias.npz"))
    fp = str(tmp_path / "bc.npz")
    np.savez(fp, u_scalar=g["scalar"], u_adder=g["adder"])
    d = g["data"]
    sl = (slice(1, 5), slice(0, 4), slice(None))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert np.array_equal(bias.global_linear_bc(d, 1.1, -0.3, out_range=(-1, 1)), g["global"])
        assert np.array_equal(bias.local_linear_bc(d, None, "u", fp, smoothing=0.8,
                                                   out_range=(-2, 2)), g["local"])
        assert np.array_equal(bias.local_linear_bc(d[sl[0], sl[1]], None, "u", fp,
                                                   lr_padded_slice=sl), g["local_slice"])
        assert np.array_equal(bias.monthly_local_linear_bc(
            d, None, "u", fp, months=g["months"], temporal_avg=True, scalar_range=(0.8, 1.2)),
            g["monthly_avg"])
        assert np.array_equal(bias.monthly_local_linear_bc(
            d, None, "u", fp, months=g["months"], temporal_avg=False, smoothing=0.5,
            adder_range=(-1, 1)), g["monthly"])
    with pytest.raises(RuntimeError):
        bias.bias_correct_features(np.zeros((6, 5, 8, 1), np.float32), ["v"], None,
                                   "local_linear_bc", {"v": {"bias_fp": fp}})
    with pytest.raises(KeyError):
        bias.bias_correct_features(np.zeros((6, 5, 8, 1), np.float32), ["u"], None, "no_such_bc",
                                   {"u": {}})


@pytest.mark.gpu
def test_forward_pass_bias_correction_hook(cuda, tmp_path):
    """strategy.py:502-517: the chunk is bias-corrected before the generator sees it."""
    from sup3r_b200 import configs as C
    from sup3r_b200.models import Sup3rGan
    from sup3r_b200.pipeline import ArrayInputHandler, ForwardPass, ForwardPassStrategy
    Sup3rGan.seed(0)
    m = Sup3rGan(C.spatiotemporal_generator(2, 2, (2,), n_blocks=1, filters=16),
                 C.discriminator(3, "same", (8,)),
                 meta={"lr_features": ["u", "v"], "hr_out_features": ["u", "v"], "s_enhance": 2,
                       "t_enhance": 2})
    rng = np.random.default_rng(0)
    data = rng.standard_normal((8, 8, 6, 2)).astype(np.float32)
    fp = str(tmp_path / "bc.npz")
    scalar = (1 + 0.3 * rng.standard_normal((8, 8))).astype(np.float32)
    adder = rng.standard_normal((8, 8)).astype(np.float32)
    np.savez(fp, u_scalar=scalar, u_adder=adder)
    mk = lambda d, **kw: ForwardPassStrategy(model=m, input_handler=ArrayInputHandler(d, ["u", "v"]),
                                             fwp_chunk_shape=(4, 4, 6), spatial_pad=1, **kw)
    got = ForwardPass.run(mk(data, bias_correct_method="local_linear_bc",
                             bias_correct_kwargs={"u": {"bias_fp": fp}}), 0)
    corrected = data.copy()
    corrected[..., 0] = data[..., 0] * scalar[..., None] + adder[..., None]
    want = ForwardPass.run(mk(corrected), 0)
    assert sorted(got) == sorted(want) == [0, 1, 2, 3]
    for k in got:
        assert np.array_equal(got[k], want[k])


def test_bias_hook_derives_months_from_a_datetime_time_index(tmp_path):
    """bias/utilities.py:263-265: the hook passes the chunk's time index to the transform
    (``date_range_kwargs`` in the reference); here a datetime index supplies the calendar month
    of every step to ``monthly_local_linear_bc`` unless the kwargs already name it."""
    import pandas as pd
    from sup3r_b200 import bias
    rng = np.random.default_rng(0)
    scalar = rng.uniform(0.5, 1.5, (4, 3, 12)).astype(np.float32)
    adder = rng.uniform(-1, 1, (4, 3, 12)).astype(np.float32)
    fp = str(tmp_path / "bc.npz")
    np.savez(fp, u_scalar=scalar, u_adder=adder)
    data = rng.standard_normal((4, 3, 6, 2)).astype(np.float32)
    ti = pd.date_range("2020-11-15", periods=6, freq="20D")
    months = np.asarray(ti.month)
    assert len(set(months)) > 2
    kw = {"u": {"bias_fp": fp, "temporal_avg": False}}
    out = bias.bias_correct_features(data.copy(), ["u", "v"], None, "monthly_local_linear_bc", kw,
                                     time_index=ti.values)
    want = bias.monthly_local_linear_bc(data[..., 0], None, "u", fp, months=months,
                                        temporal_avg=False)
    assert np.array_equal(out[..., 0], want) and np.array_equal(out[..., 1], data[..., 1])
    assert np.allclose(want, data[..., 0] * scalar[..., months - 1] + adder[..., months - 1])
    # explicit months win over the time index; an integer time index supplies nothing
    kw2 = {"u": {"bias_fp": fp, "temporal_avg": False, "months": [1] * 6}}
    out2 = bias.bias_correct_features(data.copy(), ["u", "v"], None, "monthly_local_linear_bc",
                                      kw2, time_index=ti.values)
    assert np.allclose(out2[..., 0], data[..., 0] * scalar[..., :1] + adder[..., :1])
    with pytest.raises(RuntimeError):
        bias.bias_correct_features(data.copy(), ["u", "v"], None, "monthly_local_linear_bc", kw,
                                   time_index=np.arange(6))


def test_monthly_bias_correction_with_the_reference_call_signature(tmp_path):
    """bias_transforms.py:351-487: ``date_range_kwargs`` is the fifth positional argument; the
    time index is ``pd.date_range(**kwargs)`` minus 29 February when ``drop_leap`` is set
    (preprocessing/utilities.py:222-244)."""
    import pandas as pd
    from sup3r_b200 import bias
    rng = np.random.default_rng(1)
    scalar = rng.uniform(0.5, 1.5, (4, 3, 12)).astype(np.float32)
    adder = rng.uniform(-1, 1, (4, 3, 12)).astype(np.float32)
    fp = str(tmp_path / "bc.npz")
    np.savez(fp, u_scalar=scalar, u_adder=adder)
    drk = {"start": "2020-02-27", "end": "2020-03-04", "freq": "D", "drop_leap": True}
    ti = bias.make_time_index_from_kws(drk)
    assert drk["drop_leap"] is True                      # the caller's dict is not consumed
    assert len(ti) == 6 and not ((ti.month == 2) & (ti.day == 29)).any()
    assert len(bias.make_time_index_from_kws({k: v for k, v in drk.items()
                                              if k != "drop_leap"})) == 7
    data = rng.standard_normal((4, 3, 6)).astype(np.float32)
    out = bias.monthly_local_linear_bc(data, None, "u", fp, drk, None, False)
    m = np.asarray(ti.month) - 1
    assert np.allclose(out, data * scalar[..., m] + adder[..., m])
    kw = {"u": {"bias_fp": fp, "temporal_avg": False, "date_range_kwargs": drk}}
    hooked = bias.bias_correct_features(
        np.stack([data, data], -1), ["u", "v"], None, "monthly_local_linear_bc", kw,
        time_index=pd.date_range("2021-07-01", periods=6).values)   # kwargs win over the index
    assert np.array_equal(hooked[..., 0], out)
    with pytest.raises(AssertionError):
        bias.monthly_local_linear_bc(data, None, "u", fp)
