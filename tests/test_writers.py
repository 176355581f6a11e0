"""NETCDF chunk writer + collector (sup3r/writers/nc.py, postprocessing/collectors/nc.py)."""
import json
import os

import numpy as np
import pytest

from sup3r_b200.pipeline.writers import CollectorNC, OutputHandlerNC, read_nc


def test_chunk_names_and_grouping():
    files = [f"/x/out_{t:06d}_{s:06d}.nc" for t in range(2) for s in range(3)]
    c = CollectorNC(files)
    assert c.get_chunk_indices(files[4]) == ("000001", "000001")
    g = c.group_spatial_chunks()
    assert sorted(g) == ["000000", "000001", "000002"] and all(len(v) == 2 for v in g.values())


def test_write_read_collect_round_trip(tmp_path):
    """Chunks of a synthetic field written per (time, space) chunk, collected, equal the field."""
    rng = np.random.default_rng(0)
    S1, S2, T = 8, 12, 6
    feats = ["windspeed_100m", "temperature_2m"]
    field = rng.standard_normal((S1, S2, T, 2)).astype(np.float32)
    lat = np.repeat(np.linspace(40, 39, S1)[:, None], S2, 1).astype(np.float32)
    lon = np.repeat(np.linspace(-105, -104, S2)[None], S1, 0).astype(np.float32)
    lat_lon = np.stack([lat, lon], -1)
    gids = np.arange(S1 * S2).reshape(S1, S2)
    times = np.arange(T) * 3600.0
    s_chunks = [(slice(0, 8), slice(0, 6)), (slice(0, 8), slice(6, 12))]
    t_chunks = [slice(0, 3), slice(3, 6)]
    for ti, ts in enumerate(t_chunks):
        for si, (a, b) in enumerate(s_chunks):
            fp = str(tmp_path / f"chunk_{ti:06d}_{si:06d}.nc")
            OutputHandlerNC._write_output(
                field[a, b, ts], feats, lat_lon[a, b], times[ts], fp,
                meta_data={"model_meta": {"class": "Sup3rGan"}, "full_hr_shape": [S1, S2]},
                gids=gids[a, b], transform=False)
    one = read_nc(str(tmp_path / "chunk_000001_000001.nc"))
    assert one["features"]["temperature_2m"].shape == (3, 8, 6)
    assert np.array_equal(one["gids"], gids[:, 6:]) and np.array_equal(one["time"], times[3:])
    assert json.loads(one["attrs"]["model_meta"]) == {"class": "Sup3rGan"}
    assert "date_created" in one["attrs"]
    out = str(tmp_path / "collected" / "full.nc")
    CollectorNC.collect(str(tmp_path / "chunk_*.nc"), out)
    full = read_nc(out)
    for i, f in enumerate(feats):
        assert np.array_equal(full["features"][f], np.transpose(field[..., i], (2, 0, 1)))
    assert np.array_equal(full["latitude"], lat) and np.array_equal(full["time"], times)
    assert np.array_equal(full["gids"], gids)
    with pytest.raises(AssertionError):
        CollectorNC.get_chunk_indices("/x/no_id.nc")


@pytest.mark.gpu
def test_forward_pass_writes_nc_chunks_and_collects(cuda, tmp_path):
    """ForwardPass with an .nc out_pattern (+ device-side post-processing) and CollectorNC: the
    collected file equals the in-memory chunks stitched together."""
    import warnings
    from sup3r_b200 import configs as C
    from sup3r_b200.models import Sup3rGan
    from sup3r_b200.pipeline import ArrayInputHandler, ForwardPass, ForwardPassStrategy
    Sup3rGan.seed(0)
    m = Sup3rGan(C.spatiotemporal_generator(2, 2, (2,), n_blocks=1, filters=16),
                 C.discriminator(3, "same", (8,)),
                 meta={"lr_features": ["u_100m", "v_100m"], "hr_out_features": ["u_100m", "v_100m"],
                       "s_enhance": 2, "t_enhance": 2})
    data = (np.random.default_rng(1).standard_normal((8, 8, 6, 2)) * 4).astype(np.float32)
    mk = lambda **kw: ForwardPassStrategy(
        model=m, input_handler=ArrayInputHandler(data, ["u_100m", "v_100m"]),
        fwp_chunk_shape=(4, 4, 6), spatial_pad=1, pass_workers=2, postprocess=True,
        invert_uv=True, nn_fill=False, **kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mem = ForwardPass.run(mk(), 0)
        strat = mk(out_pattern=str(tmp_path / "fwp_{file_id}.nc"))
        ForwardPass.run(strat, 0)
    assert all(os.path.exists(f) for f in strat.out_files)
    assert strat.node_finished(0)
    out = str(tmp_path / "all.nc")
    CollectorNC.collect(str(tmp_path / "fwp_*.nc"), out)
    full = read_nc(out)
    assert sorted(full["features"]) == ["winddirection_100m", "windspeed_100m"]
    ws = full["features"]["windspeed_100m"]
    assert ws.shape == (12, 16, 16) and np.isfinite(ws).all() and ws.min() >= 0
    sl = strat.fwp_slicer
    for idx, chunk in mem.items():
        s_idx, t_idx = sl.get_chunk_indices(idx)
        hs = sl.s_hr_slices[s_idx]
        assert np.array_equal(ws[:, hs[0], hs[1]], np.transpose(chunk[..., 0], (2, 0, 1)))


def test_writer_transform_default_and_dispatch(monkeypatch, tmp_path):
    """``postprocess=None`` resolves to the reference's behaviour: chunk files in the reference's
    format (.nc) get the writer-side transforms (writers/nc.py -> ``_transform_output``),
    in-memory results and .npy files stay raw; runs that post-process go through the streamed
    driver for any number of pass workers."""
    from sup3r_b200 import configs as C
    from sup3r_b200.models import Sup3rGan
    from sup3r_b200.pipeline import ArrayInputHandler, ForwardPass, ForwardPassStrategy
    m = Sup3rGan(C.spatiotemporal_generator(2, 2, (2,), n_blocks=1, filters=16),
                 C.discriminator(3, "same", (8,)), default_device="/cpu:0",
                 meta={"lr_features": ["u_100m", "v_100m"], "hr_out_features": ["u_100m", "v_100m"],
                       "s_enhance": 2, "t_enhance": 2})
    data = np.zeros((8, 8, 6, 2), np.float32)
    mk = lambda **kw: ForwardPassStrategy(
        model=m, input_handler=ArrayInputHandler(data, ["u_100m", "v_100m"]),
        fwp_chunk_shape=(4, 4, 6), **kw)
    assert mk().postprocess is False
    assert mk(out_pattern=str(tmp_path / "a_{file_id}.npy")).postprocess is False
    assert mk(out_pattern=str(tmp_path / "a_{file_id}.nc")).postprocess is True
    assert mk(out_pattern=str(tmp_path / "a_{file_id}.nc"), postprocess=False).postprocess is False
    assert mk(postprocess=True).postprocess is True
    calls = []
    monkeypatch.setattr(ForwardPass, "_run_serial",
                        classmethod(lambda cls, s, n: calls.append("serial") or {}))
    monkeypatch.setattr(ForwardPass, "_run_streamed",
                        classmethod(lambda cls, s, n, fwp=None: calls.append("streamed") or {}))
    monkeypatch.setattr(ForwardPass, "__init__", lambda self, strategy, node_index=0: setattr(
        self, "model", strategy.model))
    ForwardPass.run(mk(), 0)
    ForwardPass.run(mk(postprocess=True), 0)
    ForwardPass.run(mk(postprocess=True, pass_workers=4), 0)
    ForwardPass.run(mk(out_pattern=str(tmp_path / "b_{file_id}.nc")), 0)
    assert calls == ["serial", "streamed", "streamed", "streamed"]
