"""Host logic vs golden fixtures generated from the REAL reference code by
tools/make_golden.py (slicer index math, ExoData protocol, coarsening utilities)."""
import copy
import json
import os
import warnings

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sl(s):
    def _i(v):
        return None if v is None else int(v)
    if isinstance(s, slice):
        return ["slice", _i(s.start), _i(s.stop), _i(s.step)]
    if isinstance(s, (tuple, list)):
        return [sl(v) for v in s]
    if isinstance(s, np.integer):
        return int(s)
    return s


SLICER_ATTRS = ["s1_lr_slices", "s2_lr_slices", "t_lr_slices", "s1_lr_pad_slices",
                "s2_lr_pad_slices", "t_lr_pad_slices", "s_lr_slices", "s_lr_pad_slices",
                "s_hr_slices", "s1_hr_crop_slices", "s2_hr_crop_slices", "t_hr_crop_slices",
                "t_lr_crop_slices", "s_lr_crop_slices", "hr_crop_slices", "extra_padding",
                "n_chunks", "n_spatial_chunks", "n_time_chunks"]


@pytest.mark.parametrize("case", range(5))
def test_slicer_matches_reference(case):
    from sup3r_b200.pipeline.slicer import ForwardPassSlicer
    rec = json.load(open(os.path.join(GOLD, "slicer.json")))["slicer"][case]
    kw = dict(rec["kwargs"])
    kw["time_slice"] = slice(*kw["time_slice"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = ForwardPassSlicer(**kw)
        for attr in SLICER_ATTRS:
            assert sl(getattr(s, attr)) == rec[attr], attr
    assert s.chunk_lookup.tolist() == rec["chunk_lookup"]
    assert [list(map(int, s.get_chunk_indices(i))) for i in range(s.n_chunks)] \
        == rec["chunk_indices"]


def test_get_chunk_slices_matches_reference():
    from sup3r_b200.pipeline.utilities import get_chunk_slices
    for rec in json.load(open(os.path.join(GOLD, "slicer.json")))["get_chunk_slices"]:
        n, c, s = rec["args"]
        assert sl(get_chunk_slices(n, c, slice(*s))) == rec["out"]


def _exo_steps():
    z = np.load(os.path.join(GOLD, "exodata_inputs.npz"))
    lr, hr2, hr3 = z["lr"], z["hr2"], z["hr3"]
    return {"topography": {"steps": [
        {"model": 0, "combine_type": "input", "data": lr, "s_enhance": 1, "t_enhance": 1},
        {"model": 0, "combine_type": "layer", "data": hr2, "s_enhance": 2, "t_enhance": 2},
        {"model": 1, "combine_type": "input", "data": hr3, "s_enhance": 2, "t_enhance": 2}]},
        "sza": {"steps": [
            {"model": 1, "combine_type": "output", "data": hr3 * 2, "s_enhance": 2,
             "t_enhance": 2}]}}


def test_exodata_matches_reference():
    from sup3r_b200.exo import ExoData
    gold = json.load(open(os.path.join(GOLD, "exodata.json")))
    exo = ExoData(_exo_steps())
    assert sorted(exo.get_model_step_exo(0)) == gold["model_step_0"]
    assert sorted(exo.get_model_step_exo(1)) == gold["model_step_1"]
    assert len(exo.get_model_step_exo(0)["topography"]["steps"]) == gold["n_steps_0_topo"]
    assert np.isclose(exo.get_combine_type_data("topography", "layer").sum(), gold["layer_sum"])
    chunk = exo.get_chunk([slice(1, 4), slice(2, 6), slice(3, 8)])
    assert {f: [list(s["data"].shape) for s in chunk[f]["steps"]] for f in chunk} \
        == gold["chunk_shapes"]
    for f in chunk:
        assert np.allclose([float(s["data"].sum()) for s in chunk[f]["steps"]],
                           gold["chunk_sums"][f], rtol=1e-5)
    parts = ExoData(copy.deepcopy(_exo_steps())).split([1])
    assert [{f: [s["model"] for s in p[f]["steps"]] for f in p} for p in parts] == gold["split"]
    with pytest.raises(ValueError):
        ExoData([1, 2])
    with pytest.raises(AssertionError):
        ExoData({"x": {"nosteps": []}})
    with pytest.raises(AssertionError):
        exo.get_combine_type_data("sza", "layer")


def test_utilities_match_reference():
    from sup3r_b200.utilities import (camel_to_underscore, spatial_coarsening,
                                      temporal_coarsening)
    z = np.load(os.path.join(GOLD, "utilities.npz"))
    assert np.allclose(spatial_coarsening(z["x5"], 4), z["sc5"], atol=1e-6)
    assert np.allclose(spatial_coarsening(z["x4"], 2), z["sc4"], atol=1e-6)
    assert np.allclose(spatial_coarsening(z["x4"][0], 2, obs_axis=False), z["sc3"], atol=1e-6)
    for m in ("subsample", "average", "total", "max", "min"):
        assert np.allclose(temporal_coarsening(z["x5"], 3, m), z[f"tc_{m}"], atol=1e-6), m
    with pytest.raises(KeyError):
        temporal_coarsening(z["x5"], 3, "bogus")
    with pytest.raises(ValueError):
        spatial_coarsening(z["x4"], 3)
    for name, want in json.load(open(os.path.join(GOLD, "names.json"))).items():
        assert camel_to_underscore(name) == want


def test_timer_records_calls():
    from sup3r_b200.utilities import Timer
    t = Timer()
    assert t(lambda a: a + 1)(1) == 2
    assert "<lambda>" in t.log

    def f():
        return 3
    t(f, call_id=7)()
    assert "f" in t.log[7]


def test_forward_pass_device_check_table():
    """_output_check (sup3r/pipeline/forward_pass.py:384-425) evaluated from the per-channel
    (min, max, n_nan) table the device-side check returns (host logic, no GPU needed)."""
    import numpy as np
    from sup3r_b200.pipeline import ForwardPass
    ok = np.array([[0.0, 1.0, 0.0], [-2.0, 3.0, 0.0]], dtype=np.float32)
    assert not ForwardPass._device_check_failed(ok, allowed_const=None)
    const = np.array([[0.0, 1.0, 0.0], [2.0, 2.0, 0.0]], dtype=np.float32)
    assert ForwardPass._device_check_failed(const, allowed_const=None)
    assert ForwardPass._device_check_failed(const, allowed_const=False)
    assert not ForwardPass._device_check_failed(const, allowed_const=[2.0])
    assert not ForwardPass._device_check_failed(const, allowed_const=2.0)
    assert not ForwardPass._device_check_failed(const, allowed_const=True)
    nan = ok.copy(); nan[1, 2] = 5.0
    assert ForwardPass._device_check_failed(nan, allowed_const=[2.0])
