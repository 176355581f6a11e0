"""SURVEY section 8(f)1 model family on the GPU: ``SolarCC``, ``SolarMultiStepGan``, ``Sup3rGanDC``,
``Sup3rGanWithObs`` (+ ``Sup3rConcatObs`` / ``Sup3rObsModel``), ``BatchNormalization``, and
training with ``hr_exo_features`` (row a9).  Checked against the torch float64 restatement in
``oracle/torch_ref.py`` and the reference's own property tests
(tests/training/test_train_solar.py, test_train_conditioned_obs.py, test_train_gan_dc.py,
tests/forward_pass/test_solar_module.py)."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle.torch_ref import TorchRefNet, disc_loss
from sup3r_b200 import configs as C
from test_models_gpu import SyntheticBatchHandler, _Batch, make_model, randomize_biases, rel_err

pytestmark = pytest.mark.gpu


def _close(got, want, name, tol=2e-3):
    got = got.detach().cpu().numpy().astype(np.float64) if isinstance(got, torch.Tensor) else got
    want = want.detach().numpy() if isinstance(want, torch.Tensor) else want
    d = np.abs(got - want).max()
    assert d < tol * np.abs(want).max() + 1e-7, (name, d, np.abs(want).max())


# ---------------------------------------------------------------------------- BatchNormalization
def _bn_config():
    return [*C._conv_block(3, 16, act=False), {"class": "BatchNormalization"},
            {"class": "LeakyReLU", "alpha": 0.2}, *C._conv_block(3, 2, act=False)]


def test_batch_normalization_forward_and_gradients(cuda):
    """keras inference-mode BatchNormalization (the reference never passes training=True):
    generate vs the float64 restatement, gradients w.r.t. gamma / beta / conv weights."""
    hl = _bn_config()
    shape = (2, 5, 6, 4, 3)
    m = make_model(hl, C.discriminator(3, "same", (8,)), shape, precision="fp32")
    rng = np.random.default_rng(3)
    bn = [lyr for lyr in m.generator.layers if type(lyr).__name__ == "BatchNormalization"][0]
    bn.gamma.assign(rng.uniform(0.5, 1.5, 16).astype(np.float32))
    bn.beta.assign(rng.standard_normal(16).astype(np.float32) * 0.1)
    bn.moving_mean.assign(rng.standard_normal(16).astype(np.float32) * 0.2)
    bn.moving_variance.assign(rng.uniform(0.5, 2.0, 16).astype(np.float32))
    assert [v.name.split("/")[-1] for v in bn.weights] == ["gamma:0", "beta:0"]
    x = rng.standard_normal(shape).astype(np.float32)
    ref = TorchRefNet(hl, m.generator.get_weights(), torch.float64, requires_grad=True)
    ref.bn_state = [(bn.moving_mean.numpy(), bn.moving_variance.numpy())]
    want = ref(torch.tensor(x, dtype=torch.float64))
    got = m.generate(x)
    assert rel_err(got, want.detach().numpy()) < 1e-4
    # gradients through the tape
    tgt = rng.standard_normal(tuple(want.shape)).astype(np.float32)
    ref_g = torch.autograd.grad(((want - torch.tensor(tgt, dtype=torch.float64)) ** 2).mean(),
                                ref.weights)
    with torch.enable_grad():
        out = m._tf_generate(x)
        loss, _ = m.calc_loss_gen_content(torch.as_tensor(tgt, device=cuda), out)
        grads = torch.autograd.grad(loss, [w.value for w in m.generator_weights])
    for g, w, v in zip(grads, ref_g, m.generator_weights):
        _close(g, w, v.name)
    # save / load keeps the moving statistics
    with tempfile.TemporaryDirectory() as td:
        m.save(td)
        loaded = type(m).load(td, precision="fp32")
        assert np.array_equal(loaded.generate(x), got)


# ---------------------------------------------------------------------------------------- SolarCC
def _solar_model(cls=None, **kw):
    from sup3r_b200.models import SolarCC
    cls = cls or SolarCC
    gen = C.sup3rcc_temporal_d2t_generator(1, 24, 12, n_blocks=1, filters=16)
    disc = C.discriminator(3, "same", (8,))
    cls.seed(0)
    m = cls(gen, disc, loss="MeanAbsoluteError", precision="fp32", **kw)
    m.generator.build((1, 4, 4, 4, 1))
    randomize_biases(m.generator, np.random.default_rng(1))
    return m, gen, disc


def test_solar_cc_loss_windows_and_padding(cuda):
    """SolarCC.calc_loss (solar_cc.py:94-264) against a float64 restatement with the same
    random windows; temporal_pad / generate shapes (tests/training/test_train_solar.py:62-110)."""
    m, gen_hl, disc_hl = _solar_model()
    rng = np.random.default_rng(5)
    hr_shape = (2, 4, 4, 48, 1)   # two days
    m.init_weights((2, 4, 4, 4, 1), hr_shape)
    assert m.discriminator._built_for[2] == 8   # the disc only ever sees 8 daylight hours
    randomize_biases(m.discriminator, rng)
    true = rng.standard_normal(hr_shape).astype(np.float32)
    gen = rng.standard_normal(hr_shape).astype(np.float32)
    windows = [5, 31]
    m._sample_gen_windows = lambda t_len, n_days: windows
    w_adv = 0.3
    tt, tg = torch.as_tensor(true, device=cuda), torch.as_tensor(gen, device=cuda)
    with torch.no_grad():
        loss, det = m.calc_loss(tt, tg, weight_gen_advers=w_adv, train_gen=True,
                                compute_disc=True)
    d = TorchRefNet(disc_hl, m.discriminator.get_weights(), torch.float64)
    t64, g64 = torch.tensor(true, dtype=torch.float64), torch.tensor(gen, dtype=torch.float64)
    dg = torch.cat([d(g64[:, :, :, t0:t0 + 8]).reshape(-1) for t0 in windows])
    dt = torch.cat([d(t64[:, :, :, 8 + 24 * i:16 + 24 * i]).reshape(-1) for i in range(2)])
    content = 0.0
    for i in range(2):
        sub = slice(8 + 24 * i, 16 + 24 * i)
        pl = slice(11 + 24 * i, 13 + 24 * i)
        day = slice(24 * i, 24 * i + 24)
        content = content + ((g64[:, :, :, pl] - t64[:, :, :, pl]).abs().mean()
                             + (g64[:, :, :, day].mean(3) - t64[:, :, :, sub].mean(3)).abs().mean()) / 2
    adv = disc_loss(dg, dt)
    assert abs(float(det["loss_gen_content"]) - content.item()) < 1e-5
    assert abs(float(det["loss_gen_advers"]) - adv.item()) < 1e-5
    assert abs(float(det["loss_disc"]) - disc_loss(dt, dg).item()) < 1e-5
    assert abs(float(loss) - (content + w_adv * adv).item()) < 1e-5
    for k in ("c_sub_mean_absolute_error", "c_24h_mean_absolute_error"):
        assert k in det
    with pytest.raises(AssertionError):
        m.calc_loss(tt[:, :, :, :30].contiguous(), tg[:, :, :, :30].contiguous())
    # generate pads the time axis to low_res_t * t_enhance
    m2, _, _ = _solar_model(t_enhance=26)
    x = rng.standard_normal((1, 4, 4, 4, 1)).astype(np.float32)
    y = m2.generate(x)
    assert y.shape == (1, 4, 4, 104, 1) and m2.meta["t_enhance"] == 26
    core = super(type(m2), m2).generate(x)
    assert core.shape == (1, 4, 4, 96, 1)
    assert np.array_equal(y[:, :, :, 4:-4], core)
    assert np.array_equal(y[:, :, :, :4], core[:, :, :, 4:0:-1])


def test_solar_cc_trains(cuda):
    """A SolarCC training run decreases the loss and reloads with t_enhance
    (tests/training/test_train_solar.py:24-60)."""
    from sup3r_b200.models import SolarCC
    m, _, _ = _solar_model(learning_rate=2e-3)

    class BH(SyntheticBatchHandler):
        def __init__(self):
            rng = np.random.default_rng(0)
            self.s_enhance, self.t_enhance = 1, 24
            self.lr_features = self.hr_out_features = ["clearsky_ratio"]
            self.hr_exo_features = []
            self.means, self.stds = {"clearsky_ratio": 0.0}, {"clearsky_ratio": 1.0}
            self.batches = []
            for _ in range(3):
                hr = rng.uniform(0, 1, (2, 4, 4, 96, 1)).astype(np.float32)
                lr = hr.reshape(2, 4, 4, 4, 24, 1).mean(axis=4)
                self.batches.append(_Batch(lr, hr))
            self.val_data = self.batches[-1:]
            self.batches = self.batches[:-1]
            self.lr_shape, self.hr_shape = (4, 4, 4, 1), (4, 4, 96, 1)
            self.shapes = ((2, 4, 4, 4, 1), (2, 4, 4, 96, 1))
            self.stopped = False

    with tempfile.TemporaryDirectory() as td:
        m.train(BH(), {"spatial": "4km", "temporal": "1440min"}, n_epoch=3,
                weight_gen_advers=1e-3, train_gen=True, train_disc=True,
                disc_loss_bounds=(-1.0, 100.0), out_dir=os.path.join(td, "solar_{epoch}"))
        tl = m.history["train_loss_gen"].values
        assert np.isfinite(tl).all() and tl[-1] < tl[0]
        assert "train_c_sub_mean_absolute_error" in m.history
        loaded = SolarCC.load(os.path.join(td, "solar_2"), t_enhance=24)
        assert loaded.meta["class"] == "SolarCC" and loaded._t_enhance == 24


# ----------------------------------------------------------------------------- SolarMultiStepGan
def test_solar_multi_step_gan(cuda):
    """tests/forward_pass/test_solar_module.py: spatial solar + spatial wind -> temporal solar."""
    from sup3r_b200.models import MultiStepGan, SolarCC, SolarMultiStepGan, Sup3rGan
    disc2, disc3 = C.discriminator(2, "same", (8,)), C.discriminator(3, "same", (8,))
    Sup3rGan.seed(0)
    s_solar = Sup3rGan(C.spatial_generator(1, (2,), n_blocks=1, filters=16), disc2,
                       meta={"lr_features": ["clearsky_ratio"],
                             "hr_out_features": ["clearsky_ratio"], "s_enhance": 2, "t_enhance": 1})
    s_wind = Sup3rGan(C.spatial_generator(2, (2,), n_blocks=1, filters=16), disc2,
                      meta={"lr_features": ["u", "v"], "hr_out_features": ["u", "v"],
                            "s_enhance": 2, "t_enhance": 1})
    t_solar = SolarCC(C.sup3rcc_temporal_d2t_generator(1, 24, 12, n_blocks=1, filters=16), disc3,
                      meta={"lr_features": ["clearsky_ratio", "u", "v"],
                            "hr_out_features": ["clearsky_ratio"], "s_enhance": 1,
                            "t_enhance": 24})
    rng = np.random.default_rng(2)
    s_solar.generator.build((1, 5, 5, 1)); s_wind.generator.build((1, 5, 5, 2))
    t_solar.generator.build((1, 10, 10, 4, 3))
    for mdl in (s_solar, s_wind, t_solar):
        randomize_biases(mdl.generator, rng)
    ms = SolarMultiStepGan(MultiStepGan([s_solar]), MultiStepGan([s_wind]),
                           MultiStepGan([t_solar]))
    assert ms.lr_features == ["clearsky_ratio", "u", "v"]
    assert ms.hr_out_features == ["clearsky_ratio"]
    assert list(ms.idf_wind) == [1, 2] and list(ms.idf_solar) == [0]
    x = rng.uniform(0, 1, (4, 5, 5, 3)).astype(np.float32)
    y = ms.generate(x)
    assert y.shape == (1, 10, 10, 96, 1) and y.dtype == np.float32
    hs = s_solar.generate(x[..., :1]); hw = s_wind.generate(x[..., 1:])
    mid = np.transpose(np.concatenate((hs, hw), axis=3), (1, 2, 0, 3))[None]
    want = t_solar.generate(np.ascontiguousarray(mid))
    assert np.array_equal(y, want)
    with pytest.raises(AssertionError):   # the solar chain must be clearsky_ratio only
        SolarMultiStepGan(MultiStepGan([s_wind]), MultiStepGan([s_wind]), MultiStepGan([t_solar]))


# ------------------------------------------------------------------------------------ Sup3rGanDC
def test_sup3r_gan_dc_val_loss_updates_sampling_weights(cuda):
    """dc.py:18-116: per-bin validation losses -> normalised spatial / temporal weights."""
    from sup3r_b200.models import Sup3rGanDC
    Sup3rGanDC.seed(0)
    m = Sup3rGanDC(C.spatiotemporal_generator(2, 2, (2,), n_blocks=1, filters=16),
                   C.discriminator(3, "same", (8,)), precision="fp32")
    bh = SyntheticBatchHandler(n_batches=6, batch=2)
    bh.n_space_bins, bh.n_time_bins = 2, 3
    bh.val_data = bh.batches
    bh.spatial_weights, bh.temporal_weights = np.ones(2) / 2, np.ones(3) / 3
    seen = {}
    bh.update_weights = lambda spatial_weights, temporal_weights: seen.update(
        s=spatial_weights, t=temporal_weights)
    m.init_weights(*bh.shapes)
    total, content = m.calc_val_loss_gen(bh, 1e-3)
    assert total.shape == (2, 3) and np.all(total > 0) and np.all(content > 0)
    # bin (row, col) is validation batch row * n_time_bins + col
    with torch.no_grad():
        loss, det, _, _ = m._get_hr_exo_and_loss(bh.batches[4].low_res, bh.batches[4].high_res,
                                                 weight_gen_advers=1e-3)
    assert abs(total[1, 1] - float(loss)) < 1e-6 * abs(float(loss)) + 1e-7
    det = m.calc_val_loss(bh, 1e-3)
    assert set(det) == {"mean_val_loss_gen", "mean_val_loss_gen_content"}
    np.testing.assert_allclose(seen["t"], total.mean(0) / total.mean(0).sum(), rtol=1e-5)
    np.testing.assert_allclose(seen["s"], total.mean(1) / total.mean(1).sum(), rtol=1e-5)
    assert abs(seen["t"].sum() - 1) < 1e-5 and abs(seen["s"].sum() - 1) < 1e-5


# -------------------------------------------------------------------------------- Sup3rGanWithObs
def _obs_gen_config():
    """tests/conftest.py:80-146 of the reference (gen_config_with_concat_masked): two
    Sup3rConcatObs layers after a 2-feature convolution."""
    def ct(filters):
        return [C._pad(2), {"class": "Conv2DTranspose", "filters": filters, "kernel_size": 3,
                            "strides": 1, "activation": "relu"},
                {"class": "Cropping2D", "cropping": 4}]
    return [*ct(16), *ct(16), {"class": "SpatialExpansion", "spatial_mult": 2},
            {"class": "Activation", "activation": "relu"}, *ct(2),
            {"class": "Sup3rConcatObs", "name": "u_10m_obs", "fill_index": 0},
            {"class": "Sup3rConcatObs", "name": "v_10m_obs", "fill_index": 1}, *ct(2)]


def test_sup3r_gan_with_obs(cuda):
    """tests/training/test_train_conditioned_obs.py:22-100."""
    from sup3r_b200.models import Sup3rGanWithObs
    gen_hl = _obs_gen_config()
    Sup3rGanWithObs.seed(0)
    m = Sup3rGanWithObs(gen_hl, C.discriminator(2, "same", (8,)),
                        onshore_obs_frac={"spatial": 0.1}, loss_obs_weight=0.1,
                        learning_rate=1e-3, precision="fp32")
    m.meta["hr_out_features"] = ["u_10m", "v_10m"]
    m.meta["lr_features"] = ["u_10m", "v_10m"]
    mask = m._get_full_obs_mask(np.zeros((1, 40, 40, 2)))
    assert mask.shape == (1, 40, 40, 2) and mask.dtype == bool
    assert abs(0.1 - (1 - mask.sum() / mask.size)) < 0.05
    assert m.obs_features == ["u_10m_obs", "v_10m_obs"]
    assert m.obs_training_inds == [0, 1]
    params = m.model_params
    assert params["onshore_obs_frac"] == {"spatial": 0.1} and params["loss_obs_weight"] == 0.1

    rng = np.random.default_rng(4)
    x = rng.uniform(0, 1, (4, 6, 6, 2)).astype(np.float32)
    m.generator.build(x.shape, {"u_10m_obs": 1, "v_10m_obs": 1})
    randomize_biases(m.generator, rng)
    u_obs = rng.uniform(0, 1, (4, 12, 12, 1)).astype(np.float32)
    v_obs = rng.uniform(0, 1, (4, 12, 12, 1)).astype(np.float32)
    nanmask = rng.choice([True, False], (12, 12, 1), p=[0.9, 0.1])
    u_obs[:, nanmask] = np.nan
    v_obs[:, nanmask] = np.nan
    with pytest.raises(RuntimeError):
        m.generate(x, exogenous_data=None)
    exo = {k: {"steps": [{"model": 0, "combine_type": "layer", "data": d}]}
           for k, d in (("u_10m_obs", u_obs), ("v_10m_obs", v_obs))}
    y = m.generate(x, exogenous_data=exo)
    assert y.dtype == np.float32 and y.shape == (4, 12, 12, 2) and np.isfinite(y).all()
    ref = TorchRefNet(gen_hl, m.generator.get_weights(), torch.float64)
    want = ref(torch.tensor(x, dtype=torch.float64),
               {"u_10m_obs": u_obs.astype(np.float64), "v_10m_obs": v_obs.astype(np.float64)})
    assert rel_err(y, want.numpy()) < 1e-4

    # one gradient step with masked truth as observations: loss terms of with_obs.py:259-291
    hr = rng.uniform(0, 1, (4, 12, 12, 2)).astype(np.float32)
    m.discriminator.build(hr.shape)
    grads, det = m.get_single_grad(x, hr, m.generator_weights, weight_gen_advers=0.0,
                                   train_gen=True, train_disc=False)
    for k in ("loss_obs", "loss_non_obs", "obs_frac", "loss_gen", "loss_gen_content"):
        assert k in det, k
    assert 0.0 < float(det["obs_frac"]) < 0.5
    assert all(torch.isfinite(g).all() for g in grads)
    assert any(float(g.abs().max()) > 0 for g in grads)


def test_obs_model_layer_runs_with_and_without_observations(cuda):
    """Sup3rObsModel call protocol: layer(x, obs, extras) and the identity without data."""
    from sup3r_b200.network import CustomNetwork
    hl = [*C._conv_block(2, 8, act=True),
          {"class": "Sup3rObsModel", "name": "obs", "features": ["u_obs"],
           "exo_features": ["topography"],
           "hidden_layers": [*C._conv_block(2, 4, act=True)]},
          *C._conv_block(2, 2, act=False)]
    CustomNetwork.seed(0)
    net = CustomNetwork(hl, name="generator")
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 8, 8, 3)).astype(np.float32)
    obs = rng.standard_normal((2, 8, 8, 1)).astype(np.float32)
    obs[:, ::2] = np.nan
    topo = rng.standard_normal((2, 8, 8, 1)).astype(np.float32)
    y = net.predict(x, {"u_obs": obs, "topography": topo})
    assert tuple(y.shape) == (2, 8, 8, 2) and np.isfinite(y.numpy()).all()
    assert len(net.weights) == 6   # two convs + the embedded conv
    with pytest.raises(RuntimeError):   # built with the observation channels
        net.predict(x, {"topography": topo})


# --------------------------------------------------------- training with hr_exo_features (row a9)
def test_training_with_hr_exo_features_matches_float64_autograd(cuda):
    """get_hr_exo_input / _combine_loss_input (abstract.py:415-459): the generator takes
    topography through a Sup3rConcat layer, the truth carries it as an extra channel, the
    discriminator sees [generated, topography]; gradients vs float64 autograd."""
    gen_hl = C.sup3rcc_spatial_generator(2, 2, 2, exo="topography", filters=16)
    disc_hl = C.discriminator(2, "same", (8,))
    lr_shape, hr_shape = (2, 5, 5, 2), (2, 10, 10, 3)
    m = make_model(gen_hl, disc_hl, lr_shape, hr_shape, exo={"topography": 1},
                   loss="MeanAbsoluteError", precision="fp32",
                   meta={"lr_features": ["u", "v"], "hr_out_features": ["u", "v"],
                         "hr_exo_features": ["topography"], "s_enhance": 2, "t_enhance": 1})
    assert m.hr_exo_features == ["topography"] and m.hr_features == ["u", "v", "topography"]
    rng = np.random.default_rng(8)
    lr = rng.standard_normal(lr_shape).astype(np.float32)
    hr = rng.standard_normal(hr_shape).astype(np.float32)
    w_adv = 0.1
    g = TorchRefNet(gen_hl, m.generator.get_weights(), torch.float64, requires_grad=True)
    d = TorchRefNet(disc_hl, m.discriminator.get_weights(), torch.float64, requires_grad=True)
    hr_t = torch.tensor(hr, dtype=torch.float64)
    topo = hr_t[..., 2:3]
    gen = g(torch.tensor(lr, dtype=torch.float64), {"topography": topo})
    gen_cat = torch.cat((gen, topo), dim=-1)
    dt, dg = d(hr_t), d(gen_cat)
    content = (gen - hr_t[..., :2]).abs().mean()
    loss_gen = content + w_adv * disc_loss(dg, dt)
    ref_g = torch.autograd.grad(loss_gen, g.weights, retain_graph=True)
    ref_d = torch.autograd.grad(disc_loss(dt, dg), d.weights)
    grads, det = m.get_single_grad(lr, hr, m.generator_weights, weight_gen_advers=w_adv,
                                   train_gen=True, train_disc=False, compute_disc=True)
    assert abs(float(det["loss_gen_content"]) - content.item()) < 1e-5
    assert abs(float(det["loss_gen"]) - loss_gen.item()) < 1e-4 * abs(loss_gen.item())
    for got, want, v in zip(grads, ref_g, m.generator_weights):
        _close(got, want, v.name)
    grads, _ = m.get_single_grad(lr, hr, m.discriminator_weights, weight_gen_advers=w_adv,
                                 train_gen=False, train_disc=True)
    for got, want, v in zip(grads, ref_d, m.discriminator_weights):
        _close(got, want, v.name)
