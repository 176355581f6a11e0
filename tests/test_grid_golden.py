"""High-resolution output grid and time axis against vectors produced by
the REAL reference functions (tools/make_golden_grid.py execs ``OutputHandler.get_lat_lon /
pad_lat_lon / is_increasing_lons / get_times`` of sup3r/writers/base.py): regular, rotated
curvilinear and date-line crossing grids at 2x / 3x; hourly, 3-hourly, leap-day and single-step
time indices.  CPU only."""
import importlib.util
import json
import os

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "make_golden_grid", os.path.join(ROOT, "tools", "make_golden_grid.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)
G = np.load(os.path.join(ROOT, "tests", "golden", "grid.npz"))
REC = json.loads(str(G["record"]))


class Ours:
    """The reference's call signatures on this repo's functions."""

    @staticmethod
    def get_lat_lon(low_res_lat_lon, shape):
        from sup3r_b200.pipeline.strategy import _hr_lat_lon
        return _hr_lat_lon(low_res_lat_lon, shape=shape)

    @staticmethod
    def get_times(low_res_times, shape):
        from sup3r_b200.pipeline.strategy import _hr_times
        return _hr_times(low_res_times.values, shape)


def test_hr_grid_and_times_match_reference():
    rec, arrs = T.scenario(Ours)
    assert rec == REC
    for k, a in arrs.items():
        assert a.dtype == G[k].dtype and np.array_equal(a, G[k]), k


def test_numeric_time_index_continues_past_the_last_step():
    from sup3r_b200.pipeline.strategy import _hr_times
    assert np.allclose(_hr_times(np.arange(3), 6), [0, 0.5, 1, 1.5, 2, 2.5])
    assert np.allclose(_hr_times(np.array([10.0]), 4), [10, 10.25, 10.5, 10.75])
    assert isinstance(_hr_times(pd.date_range("2020-01-01", periods=2, freq="1h").values, 4),
                      pd.DatetimeIndex)


def test_golden_is_reproducible_from_the_reference_when_present():
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    rec, arrs = T.scenario(T.load_reference())
    assert rec == REC and all(np.array_equal(a, G[k]) for k, a in arrs.items())
