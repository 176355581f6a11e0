"""``ForwardPass`` on everything the stacked-batch path does not cover: 4-D models, exogenous data
(tests/forward_pass/test_forward_pass_exo.py of the reference), ``MultiStepGan`` chains -- the
streamed driver (device-side crop + output check, pinned D2H on a copy stream) must give exactly
what the serial reference-shaped driver gives -- plus BASELINE configs[2] shapes (sup3rcc wind:
20x20x72x6 chunk, 24x temporal then 5x spatial with topography) and ``Sup3rAdder``."""
import numpy as np
import pytest
import torch

from oracle.torch_ref import TorchRefNet
from sup3r_b200 import configs as C
from test_models_gpu import make_model, rel_err, rms_err

pytestmark = pytest.mark.gpu


def _spatial_model(f=2, exo=None, s=2, seed=0):
    hl = (C.sup3rcc_spatial_generator(f, s, 2, exo=exo, filters=16) if exo
          else C.spatial_generator(f, (s,), n_blocks=1, filters=16))
    feats = ["u", "v", "w", "x", "y", "z"][:f]
    return make_model(hl, C.discriminator(2, "same", (8,)), (3, 8, 8, f),
                      exo={exo: 1} if exo else None, seed=seed,
                      meta={"lr_features": feats, "hr_out_features": feats,
                            "s_enhance": s, "t_enhance": 1}), hl


def _run_both(make_strategy):
    from sup3r_b200.pipeline import ForwardPass
    a = ForwardPass.run(make_strategy(1), 0)
    b = ForwardPass.run(make_strategy(4), 0)
    assert sorted(a) == sorted(b) and len(a) > 1
    for k in a:
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
    return a


def test_streamed_driver_4d_model(cuda):
    from sup3r_b200.pipeline import ArrayInputHandler, ForwardPassStrategy
    m, _ = _spatial_model()
    data = np.random.default_rng(0).standard_normal((12, 12, 6, 2)).astype(np.float32)
    outs = _run_both(lambda w: ForwardPassStrategy(
        model=m, input_handler=ArrayInputHandler(data, ["u", "v"]), fwp_chunk_shape=(6, 6, 6),
        spatial_pad=2, temporal_pad=0, pass_workers=w))
    assert outs[0].shape == (12, 12, 6, 2)


def test_streamed_driver_with_exo_layer_data(cuda):
    """tests/forward_pass/test_forward_pass_exo.py: hi-res topography enters through a
    Sup3rConcat layer, sliced and edge-padded per chunk with the step's enhancement."""
    from sup3r_b200.pipeline import ArrayInputHandler, ForwardPass, ForwardPassStrategy
    m, hl = _spatial_model(exo="topography")
    rng = np.random.default_rng(1)
    data = rng.standard_normal((12, 12, 4, 2)).astype(np.float32)
    topo = rng.standard_normal((24, 24, 1)).astype(np.float32)

    def exo():
        return {"topography": {"steps": [{"model": 0, "combine_type": "layer",
                                          "data": topo.copy(), "s_enhance": 2, "t_enhance": 1}]}}

    outs = _run_both(lambda w: ForwardPassStrategy(
        model=m, input_handler=ArrayInputHandler(data, ["u", "v"]), fwp_chunk_shape=(6, 6, 4),
        spatial_pad=1, temporal_pad=0, pass_workers=w, exo_data=exo()))
    assert outs[0].shape == (12, 12, 4, 2)
    # one chunk over the whole domain == generate with the full exo field == the oracle
    strat = ForwardPassStrategy(model=m, input_handler=ArrayInputHandler(data, ["u", "v"]),
                                fwp_chunk_shape=(12, 12, 4), pass_workers=2, exo_data=exo())
    out = ForwardPass.run(strat, 0)[0]
    x = np.transpose(data, (2, 0, 1, 3))
    topo_t = np.repeat(topo[None], 4, axis=0)
    direct = m.generate(x, exogenous_data={"topography": {"steps": [
        {"model": 0, "combine_type": "layer", "data": topo_t}]}})
    assert np.array_equal(out, np.transpose(direct, (1, 2, 0, 3)))
    ref = TorchRefNet(hl, m.generator.get_weights(), torch.float64)(
        torch.tensor(x, dtype=torch.float64), {"topography": topo_t.astype(np.float64)})
    assert rel_err(direct, ref.numpy()) < 1e-3


def test_streamed_driver_multi_step_gan(cuda):
    """tests/forward_pass/test_multi_step.py through the tiler: 4-D spatial step then 5-D
    temporal step; device-resident intermediates."""
    from sup3r_b200.models import MultiStepGan
    from sup3r_b200.pipeline import ArrayInputHandler, ForwardPassStrategy
    m1, _ = _spatial_model()
    m2 = make_model(C.spatiotemporal_generator(2, 1, (2,), n_blocks=1, head_filters=16, filters=16),
                    C.discriminator(3, "same", (8,)), (1, 12, 12, 4, 2), seed=3,
                    meta={"lr_features": ["u", "v"], "hr_out_features": ["u", "v"],
                          "s_enhance": 1, "t_enhance": 2})
    ms = MultiStepGan([m1, m2])
    data = np.random.default_rng(2).standard_normal((12, 12, 8, 2)).astype(np.float32)
    outs = _run_both(lambda w: ForwardPassStrategy(
        model=ms, input_handler=ArrayInputHandler(data, ["u", "v"]), fwp_chunk_shape=(6, 6, 8),
        spatial_pad=1, temporal_pad=0, pass_workers=w))
    assert outs[0].shape == (12, 12, 16, 2)


def test_streamed_driver_output_check(cuda):
    """Constant output -> MemoryError from the device-side check (forward_pass.py:384-425)."""
    from sup3r_b200.pipeline import ArrayInputHandler, ForwardPass, ForwardPassStrategy
    m, _ = _spatial_model()
    m.generator.set_weights([np.zeros_like(w) for w in m.generator.get_weights()])
    data = np.random.default_rng(3).standard_normal((12, 12, 4, 2)).astype(np.float32)
    strat = ForwardPassStrategy(model=m, input_handler=ArrayInputHandler(data, ["u", "v"]),
                                fwp_chunk_shape=(6, 6, 4), pass_workers=2)
    with pytest.raises(MemoryError):
        ForwardPass.run(strat, 0)
    strat = ForwardPassStrategy(model=m, input_handler=ArrayInputHandler(data, ["u", "v"]),
                                fwp_chunk_shape=(6, 6, 4), pass_workers=2, allowed_const=[0])
    assert len(ForwardPass.run(strat, 0)) == 4


def test_sup3r_adder_generator(cuda):
    """Sup3rAdder (hi-res exo added mid-network) against the float64 restatement, all modes."""
    hl = [*C._conv_block(2, 64), *C._conv_block(2, 4, act=False),
          {"class": "SpatialExpansion", "spatial_mult": 2}, {"class": "Sup3rAdder", "name": "topo"},
          *C._conv_block(2, 64), *C._conv_block(2, 2, act=False)]
    shape = (2, 9, 7, 3)
    m = make_model(hl, C.discriminator(2, "same", (8,)), shape, exo={"topo": 1})
    rng = np.random.default_rng(6)
    x = rng.standard_normal(shape).astype(np.float32)
    topo = rng.standard_normal((2, 18, 14, 1)).astype(np.float32)
    ref = TorchRefNet(hl, m.generator.get_weights(), torch.float64)(
        torch.tensor(x, dtype=torch.float64), {"topo": topo.astype(np.float64)}).numpy()
    exo = {"topo": {"steps": [{"model": 0, "combine_type": "layer", "data": topo}]}}
    for precision, tol in (("fp32", 1e-4), ("fp16c", 1e-3), ("bf16x3", 1e-3)):
        y = m.generate(x, exogenous_data=exo, precision=precision)
        assert rel_err(y, ref) < tol and rms_err(y, ref) < tol, precision


def test_baseline_config2_shapes_sup3rcc_chain(cuda):
    """BASELINE configs[2]: sup3rcc wind, one (20, 20, 72, 6) chunk -> 24x temporal (5-D,
    depth_to_time) -> 5x spatial with topography (4-D, Sup3rConcat).  Full-size run in the default
    precision against the fp32 kernels (themselves oracle-checked at small shapes), and the
    spatial step against the float64 oracle on two of its 1728 time slices."""
    from sup3r_b200.models import MultiStepGan
    f = 6
    feats = ["u_10m", "v_10m", "u_100m", "v_100m", "t_2m", "rh_2m"]
    t_hl = C.sup3rcc_temporal_d2t_generator(f, 24, 12, n_blocks=4)
    s_hl = C.sup3rcc_spatial_generator(f, 5, 4, exo="topography")
    m1 = make_model(t_hl, C.discriminator(3, "same", (8,)), (1, 20, 20, 72, f),
                    meta={"lr_features": feats, "hr_out_features": feats, "s_enhance": 1,
                          "t_enhance": 24})
    m2 = make_model(s_hl, C.discriminator(2, "same", (8,)), (4, 20, 20, f),
                    exo={"topography": 1}, seed=5,
                    meta={"lr_features": feats, "hr_out_features": feats, "s_enhance": 5,
                          "t_enhance": 1})
    ms = MultiStepGan([m1, m2])
    rng = np.random.default_rng(9)
    x = rng.standard_normal((1, 20, 20, 72, f)).astype(np.float32)
    topo = rng.standard_normal((100, 100, 1)).astype(np.float32)

    def exo(n_t):
        return {"topography": {"steps": [{"model": 1, "combine_type": "layer",
                                          "data": np.repeat(topo[None], n_t, axis=0)}]}}

    y = ms.generate(x, exogenous_data=exo(1728))
    assert y.shape == (1728, 100, 100, f) and np.isfinite(y).all()
    for mdl in (m1, m2):
        mdl.precision = "fp32"
    y32 = ms.generate(x, exogenous_data=exo(1728))
    assert rel_err(y, y32) < 1e-3 and rms_err(y, y32) < 1e-3, (rel_err(y, y32), rms_err(y, y32))
    # spatial step vs float64 oracle on two time slices of the temporal step's output
    mid = m1.generate(x)                                     # (1, 20, 20, 1728, 6)
    sl = np.ascontiguousarray(np.transpose(mid[0][:, :, [0, 1000]], (2, 0, 1, 3)))
    e2 = {"topography": {"steps": [{"model": 0, "combine_type": "layer",
                                    "data": np.repeat(topo[None], 2, axis=0)}]}}
    got = m2.generate(sl, exogenous_data=e2)
    ref = TorchRefNet(s_hl, m2.generator.get_weights(), torch.float64)(
        torch.tensor(sl, dtype=torch.float64),
        {"topography": np.repeat(topo[None], 2, axis=0).astype(np.float64)}).numpy()
    assert rel_err(got, ref) < 1e-4


def test_forward_pass_postprocess_on_device(cuda):
    """strategy.postprocess: writers/base.py:297-346 (u/v -> ws/wd, limits) applied on the GPU
    before the chunk leaves; equals the numpy restatement applied to the raw chunk output."""
    import os
    import json
    import tempfile
    import warnings
    from oracle import postprocess_ref as P
    from sup3r_b200.pipeline import ArrayInputHandler, ForwardPass, ForwardPassStrategy
    m, _ = _spatial_model()
    m.meta["lr_features"] = m.meta["hr_out_features"] = ["u_100m", "v_100m"]
    data = (np.random.default_rng(5).standard_normal((12, 12, 4, 2)) * 3).astype(np.float32)
    raw = ForwardPass.run(ForwardPassStrategy(
        model=m, input_handler=ArrayInputHandler(data, ["u_100m", "v_100m"]),
        fwp_chunk_shape=(6, 6, 4), spatial_pad=1, pass_workers=2), 0)
    with tempfile.TemporaryDirectory() as td, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        strat = ForwardPassStrategy(
            model=m, input_handler=ArrayInputHandler(data, ["u_100m", "v_100m"]),
            fwp_chunk_shape=(6, 6, 4), spatial_pad=1, pass_workers=2, postprocess=True,
            invert_uv=True, nn_fill=False, out_pattern=os.path.join(td, "c_{file_id}.npy"))
        ForwardPass.run(strat, 0)
        for i in range(strat.n_chunks):
            got = np.load(strat.out_files[i])
            chunk = strat.init_chunk(i)
            ws, names = P.transform_output(raw[i], ["u_100m", "v_100m"], chunk.hr_lat_lon,
                                           invert=True)
            assert names == ["windspeed_100m", "winddirection_100m"]
            meta = json.load(open(strat.out_files[i] + ".meta.json"))
            assert meta["features"] == names
            assert np.abs(got[..., 0] - ws[..., 0]).max() < 1e-4
            dwd = np.abs(got[..., 1] - ws[..., 1])
            assert np.minimum(dwd, 360 - dwd).max() < 2e-2


def test_three_step_chain_sup3rwind_shape(cuda):
    """BASELINE configs[3](i) / tests/forward_pass/test_multi_step.py:61-127: 3x spatial (4-D) ->
    5x spatial (4-D) -> 12x temporal (5-D), 6 features, as one MultiStepGan."""
    from sup3r_b200.models import MultiStepGan
    feats = [f"f{i}" for i in range(6)]
    meta = lambda s, t: {"lr_features": feats, "hr_out_features": feats, "s_enhance": s,
                         "t_enhance": t}
    m1 = make_model(C.spatial_generator(6, (3,), n_blocks=1, filters=16),
                    C.discriminator(2, "same", (8,)), (4, 4, 4, 6), meta=meta(3, 1))
    m2 = make_model(C.sup3rcc_spatial_generator(6, 5, 2, filters=16),
                    C.discriminator(2, "same", (8,)), (4, 12, 12, 6), seed=2, meta=meta(5, 1))
    m3 = make_model(C.spatiotemporal_generator(6, 1, (2, 2, 3), n_blocks=1, head_filters=16,
                                               filters=16),
                    C.discriminator(3, "same", (8,)), (1, 60, 60, 4, 6), seed=4, meta=meta(1, 12))
    ms = MultiStepGan([m1, m2, m3])
    assert ms.s_enhance == 15 and ms.t_enhance == 12
    assert ms.s_enhancements == [3, 5, 1] and ms.t_enhancements == [1, 1, 12]
    x = np.random.default_rng(0).standard_normal((4, 4, 4, 6)).astype(np.float32)
    y = ms.generate(x)
    assert y.shape == (1, 60, 60, 48, 6) and y.dtype == np.float32 and np.isfinite(y).all()
    a = m2.generate(m1.generate(x))
    want = m3.generate(np.ascontiguousarray(np.transpose(a, (1, 2, 0, 3))[None]))
    assert np.array_equal(y, want)


def test_end_to_end_train_save_load_forward_pass_collect(cuda, tmp_path):
    """tests/training/test_end_to_end.py + tests/pipeline/test_pipeline.py in miniature, every
    stage on this library: on-device batch handler -> train 2 epochs -> save -> load ->
    ForwardPass over a chunked domain with device-side post-processing -> .nc chunk files ->
    collection; the collected field follows generate() on the whole (unchunked) domain."""
    import os
    import warnings
    from sup3r_b200.batch import DeviceBatchHandler
    from sup3r_b200.models import Sup3rGan
    from sup3r_b200.pipeline import ArrayInputHandler, ForwardPass, ForwardPassStrategy
    from sup3r_b200.pipeline.writers import CollectorNC, read_nc
    rng = np.random.default_rng(3)
    feats = ["temperature_2m", "relativehumidity_2m"]
    base = rng.standard_normal((6, 6, 8, 2))
    hr = (np.repeat(np.repeat(np.repeat(base, 4, 0), 4, 1), 4, 2) * [5, 10] + [15, 50]).astype(np.float32)
    bh = DeviceBatchHandler(hr, feats, sample_shape=(12, 12, 8), batch_size=4, n_batches=3,
                            s_enhance=2, t_enhance=2, temporal_coarsening_method="average")
    Sup3rGan.seed(0)
    m = Sup3rGan(C.spatiotemporal_generator(2, 2, (2,), n_blocks=1, filters=16),
                 C.discriminator(3, "same", (8,)), learning_rate=2e-3)
    out_dir = str(tmp_path / "gan_{epoch}")
    m.train(bh, {"spatial": "8km", "temporal": "60min"}, n_epoch=2, weight_gen_advers=1e-3,
            train_gen=True, train_disc=True, disc_loss_bounds=(-1.0, 100.0), out_dir=out_dir)
    loaded = Sup3rGan.load(out_dir.format(epoch=1))
    assert loaded.means["temperature_2m"] == pytest.approx(float(hr[..., 0].mean()), rel=1e-4)
    assert loaded.meta["s_enhance"] == 2 and loaded.meta["t_enhance"] == 2
    lr = (rng.standard_normal((12, 12, 16, 2)) * [5, 10] + [15, 50]).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        strat = ForwardPassStrategy(
            model=loaded, input_handler=ArrayInputHandler(lr, feats), fwp_chunk_shape=(6, 6, 16),
            spatial_pad=3, temporal_pad=0, pass_workers=2, postprocess=True, nn_fill=False,
            out_pattern=str(tmp_path / "fwp_{file_id}.nc"), pad_mode="reflect")
        ForwardPass.run(strat, 0)
    assert strat.n_chunks == 4 and all(os.path.exists(f) for f in strat.out_files)
    full = read_nc(CollectorNC.collect(str(tmp_path / "fwp_*.nc"), str(tmp_path / "full.nc")))
    t2m = full["features"]["temperature_2m"]
    assert t2m.shape == (32, 24, 24) and np.isfinite(t2m).all()
    whole = loaded.generate(lr[None])[0]                       # (24, 24, 32, 2), unchunked
    whole = np.clip(whole, [-200, 0], [100, 100])              # the writer-side limits
    got = np.transpose(t2m, (1, 2, 0))
    # (the 3-voxel halo is shorter than this generator's receptive field, so chunked and unchunked
    #  fields agree on average, not voxel by voxel: the exact statement is
    #  test_forward_pass_chunked_close_to_unchunked)
    assert np.abs(got - whole[..., 0]).mean() < 0.05 * np.abs(whole[..., 0]).mean()
