"""Content losses: the numpy oracle against the reference's own known-answer
identities (CPU), and sup3r_b200.loss_metrics against the oracle + directional-derivative checks
of the gradients (GPU)."""
import zlib

import numpy as np
import pytest

from oracle import losses_ref as R

RNG = np.random.default_rng(42)


# ------------------------------------------------------------------ oracle pins (CPU)
def test_oracle_material_derivative_equals_np_gradient():
    # reference tests/utilities/test_loss_metrics.py:263-289
    x = RNG.random((6, 10, 10, 8, 3))
    u = np.gradient(x[..., 0], axis=3) + x[..., 0] * np.gradient(x[..., 0], axis=1) \
        + x[..., 1] * np.gradient(x[..., 0], axis=2)
    v = np.gradient(x[..., 1], axis=3) + x[..., 0] * np.gradient(x[..., 1], axis=1) \
        + x[..., 1] * np.gradient(x[..., 1], axis=2)
    assert np.allclose(R.compute_md(x, 0), u) and np.allclose(R.compute_md(x, 1), v)
    with pytest.raises(ValueError):
        R.derivative(x, 0)
    with pytest.raises(AssertionError):
        R.material_derivative_loss(x[..., 0], x[..., 0])
    assert R.material_derivative_loss(x, x.copy()) == 0


def test_oracle_low_res_loss_identities():
    # reference tests/utilities/test_loss_metrics.py:174-260
    x = RNG.uniform(-1, 1, (3, 10, 10, 48, 2))
    y = RNG.uniform(-1, 1, (3, 10, 10, 48, 2))
    assert np.isclose(R.low_res_loss(x, y), R.mse(x, y))
    s = 5
    xl = x.reshape(3, 2, s, 2, s, 48, 2).mean(axis=(2, 4))
    yl = y.reshape(3, 2, s, 2, s, 48, 2).mean(axis=(2, 4))
    assert np.isclose(R.low_res_loss(x, y, s_enhance=s), R.mse(xl, yl))
    xt, yt = xl[:, :, :, ::12], yl[:, :, :, ::12]
    assert np.isclose(R.low_res_loss(x, y, s_enhance=s, t_enhance=12, t_method="subsample"),
                      R.mse(xt, yt))
    xa = xl.reshape(3, 2, 2, 4, 12, 2).mean(axis=4)
    ya = yl.reshape(3, 2, 2, 4, 12, 2).mean(axis=4)
    assert np.isclose(R.low_res_loss(x, y, s_enhance=s, t_enhance=12, t_method="average",
                                     tf_loss="MeanAbsoluteError"), R.mae(xa, ya))
    x4, y4 = x[:, :, :, 0], y[:, :, :, 0]
    assert np.isclose(R.low_res_loss(x4, y4, s_enhance=s),
                      R.mse(x4.reshape(3, 2, s, 2, s, 2).mean(axis=(2, 4)),
                            y4.reshape(3, 2, s, 2, s, 2).mean(axis=(2, 4))))


def test_oracle_extremes_and_coarse_and_mmd():
    # reference tests/utilities/test_loss_metrics.py:26-141
    x = np.zeros((1, 1, 1, 72, 1)); y = np.zeros((1, 1, 1, 72, 1))
    x[..., 24, 0] = 20; y[..., 25, 0] = 25
    assert R.temporal_extremes_loss(x, y) > 1.5
    x[..., 24, 0] = -20; y[..., 25, 0] = -25
    assert R.temporal_extremes_loss(x, y) > 1.5
    x = np.zeros((1, 10, 10, 2, 1)); y = np.zeros((1, 10, 10, 2, 1))
    x[:, 5, 5, :, 0] = 20; y[:, 5, 5, :, 0] = 25
    assert R.spatial_extremes_loss(x, y) > 1.5
    x = RNG.uniform(0, 1, (6, 10, 10, 8, 3)); y = RNG.uniform(0, 1, (6, 10, 10, 8, 3))
    assert R.mse(x, y) > 10 * R.coarse_mse_loss(x, y)
    a = np.zeros((6, 10, 10, 8, 3)); b = np.zeros((6, 10, 10, 8, 3))
    a[:, 7:9, 7:9] = 1; b[:, 2:5, 2:5] = 1
    assert (R.mmd_loss(a, b) + R.mse(a, b)) / 2 > R.mse(a, b)
    assert abs(R.mmd_loss(a, a)) < 1e-12


# ------------------------------------------------------------------ device losses (GPU)
CASES = [
    ("MeanSquaredError", {}, (2, 6, 6, 8, 3), R.mse),
    ("MeanAbsoluteError", {}, (2, 6, 6, 8, 3), R.mae),
    ("ExpLoss", {}, (2, 6, 6, 8, 3), R.exp_loss),
    ("MmdLoss", {}, (4, 5, 5, 6, 2), R.mmd_loss),
    ("MaterialDerivativeLoss", {}, (2, 7, 6, 8, 5), R.material_derivative_loss),
    ("SpatialDerivativeLoss", {}, (2, 7, 6, 8, 3), R.spatial_derivative_loss),
    ("SpatialDerivativeLoss", {}, (3, 7, 6, 2), R.spatial_derivative_loss),
    ("TemporalDerivativeLoss", {}, (2, 5, 6, 9, 3), R.temporal_derivative_loss),
    ("CoarseMseLoss", {}, (2, 6, 6, 8, 3), R.coarse_mse_loss),
    ("SpatialExtremesLoss", {}, (2, 6, 6, 8, 3), R.spatial_extremes_loss),
    ("SpatialExtremesLoss", {}, (3, 6, 6, 2), R.spatial_extremes_loss),
    ("TemporalExtremesLoss", {}, (2, 6, 6, 8, 3), R.temporal_extremes_loss),
    ("SpatialFftLoss", {}, (2, 8, 6, 3), R.spatial_fft_loss),
    ("SpatiotemporalFftLoss", {}, (2, 6, 5, 8, 2), R.spatiotemporal_fft_loss),
    ("LowResLoss", dict(s_enhance=2, t_enhance=4, t_method="average"), (2, 6, 6, 8, 3), None),
    ("LowResLoss", dict(s_enhance=3, t_enhance=2, t_method="subsample",
                        tf_loss="MeanAbsoluteError", ex_loss="TemporalExtremesLoss"),
     (2, 6, 6, 8, 3), None),
    ("LowResLoss", dict(s_enhance=2, ex_loss="SpatialExtremesLoss"), (3, 6, 6, 2), None),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,kwargs,shape,ref", CASES)
def test_device_loss_matches_oracle_and_gradient(cuda, name, kwargs, shape, ref):
    import torch
    from sup3r_b200 import loss_metrics
    rng = np.random.default_rng(zlib.crc32(repr((name, shape, sorted(kwargs.items()))).encode()))
    x = rng.standard_normal(shape).astype(np.float32)
    y = rng.standard_normal(shape).astype(np.float32)
    fn = getattr(loss_metrics, name)(**kwargs)
    want = ref(x, y) if ref is not None else R.low_res_loss(x, y, **kwargs)
    xt = torch.tensor(x, device=cuda, requires_grad=True)
    yt = torch.tensor(y, device=cuda)
    val = fn(xt, yt)
    assert val.ndim == 0
    fval = float(val.detach())
    assert abs(fval - want) <= 2e-5 * max(1.0, abs(want)), (fval, want)
    # gradient w.r.t. the generated tensor: directional derivative in float64 on the oracle
    val.backward()
    g = xt.grad.double().cpu().numpy()
    d = rng.standard_normal(shape)
    eps = 1e-5
    oracle = (lambda a: ref(a, y.astype(np.float64))) if ref is not None else \
        (lambda a: R.low_res_loss(a, y.astype(np.float64), **kwargs))
    fd = (oracle(x.astype(np.float64) + eps * d) - oracle(x.astype(np.float64) - eps * d)) / (2 * eps)
    got = float(np.sum(g * d))
    assert abs(got - fd) <= 2e-3 * max(abs(fd), 1e-3) + 1e-6, (got, fd)


def test_oracle_sliced_wasserstein_identities():
    """Properties of the reference formula (loss_metrics.py:754-793): zero for equal inputs,
    and with unit projections onto single points it is the squared difference of the sorted
    point values."""
    x = RNG.standard_normal((2, 3, 3, 2, 2)); y = RNG.standard_normal((2, 3, 3, 2, 2))
    proj = RNG.standard_normal((16, 18)); proj /= np.linalg.norm(proj, axis=-1, keepdims=True)
    assert R.sliced_wasserstein_loss(x, x, proj) == 0.0
    assert R.sliced_wasserstein_loss(x, y, proj) > 0
    eye = np.eye(18)
    want = np.mean((np.sort(x.reshape(2, 18, 2), axis=1) - np.sort(y.reshape(2, 18, 2), axis=1)) ** 2)
    assert np.isclose(R.sliced_wasserstein_loss(x, y, eye), want)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 5, 4, 3, 2), (3, 6, 5, 2)])
def test_sliced_wasserstein_loss(cuda, shape, monkeypatch):
    """Device loss == numpy restatement on the same projection matrix; gradient against float64
    finite differences; fresh unit-norm projections at every call; zero for equal inputs."""
    import torch
    from sup3r_b200 import loss_metrics
    rng = np.random.default_rng(31)
    x = rng.standard_normal(shape).astype(np.float32)
    y = rng.standard_normal(shape).astype(np.float32)
    fn = loss_metrics.get_loss_class("SlicedWassersteinLoss")(n_projections=48)
    n_pts = int(np.prod(shape[1:-1]))
    p_a = fn.projections(n_pts, torch.device(cuda))
    p_b = fn.projections(n_pts, torch.device(cuda))
    assert p_a.shape == (48, n_pts) and not torch.equal(p_a, p_b)
    assert torch.allclose(p_a.norm(dim=-1), torch.ones(48, device=cuda), atol=1e-5)
    proj = p_a.double().cpu().numpy()
    monkeypatch.setattr(fn, "projections", lambda n, dev: p_a)
    xt = torch.tensor(x, device=cuda, requires_grad=True)
    yt = torch.tensor(y, device=cuda)
    val = fn(xt, yt)
    want = R.sliced_wasserstein_loss(x, y, proj)
    assert val.ndim == 0 and abs(float(val) - want) <= 2e-5 * max(1.0, want)
    assert float(fn(yt, yt)) == 0.0
    val.backward()
    d = rng.standard_normal(shape)
    eps = 1e-6
    fd = (R.sliced_wasserstein_loss(x + eps * d, y, proj)
          - R.sliced_wasserstein_loss(x - eps * d, y, proj)) / (2 * eps)
    got = float(np.sum(xt.grad.double().cpu().numpy() * d))
    assert abs(got - fd) <= 2e-3 * max(abs(fd), 1e-3) + 1e-6, (got, fd)


@pytest.mark.gpu
def test_multiterm_loss(cuda):
    # reference tests/utilities/test_loss_metrics.py:292-309
    import torch
    from sup3r_b200 import loss_metrics
    from sup3r_b200.models.abstract import AbstractSingleModel
    x = RNG.random((6, 10, 10, 8, 3)).astype(np.float32)
    y = (x + 0.1 * RNG.random(x.shape)).astype(np.float32)
    xt, yt = torch.tensor(x, device=cuda), torch.tensor(y, device=cuda)
    multi = AbstractSingleModel.get_loss_fun({"MaterialDerivativeLoss": {}, "MeanAbsoluteError": {},
                                              "term_weights": [0.2, 0.8]})
    loss, details = multi(xt, yt)
    want = 0.2 * R.material_derivative_loss(x, y) + 0.8 * R.mae(x, y)
    assert np.isclose(float(loss), want, rtol=1e-4)
    assert set(details) == {"material_derivative_loss", "mean_absolute_error"}
    with pytest.raises(KeyError):
        AbstractSingleModel.get_loss_fun("NoSuchLoss")
