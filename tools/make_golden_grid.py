"""Golden vectors of the high-resolution output grid / time axis (SURVEY 8(a) row a20:
``ForwardPassStrategy.hr_lat_lon``, ``init_chunk`` hr_times) from the REAL reference functions:
``OutputHandler.get_lat_lon / pad_lat_lon / is_increasing_lons / get_times``
(sup3r/writers/base.py:347-549) and ``get_time_index_freqs`` (preprocessing/utilities.py:141-170)
exec'd from their source text.

    python tools/make_golden_grid.py   ->  tests/golden/grid.npz
"""
import json
import os
import textwrap
from unittest.mock import MagicMock

import numpy as np
import pandas as pd
from scipy.interpolate import griddata

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "grid.npz")


def grab(src, name):
    a = src.index(f"    def {name}(")
    b = a
    while True:
        b = src.find("\n    ", b + 1)
        if b < 0 or src[b + 5] not in (" ", "\n", ")"):
            break
    return textwrap.dedent(src[a:b if b > 0 else len(src)])


def load_reference():
    wsrc = open(os.path.join(REF, "sup3r/writers/base.py")).read()
    psrc = open(os.path.join(REF, "sup3r/preprocessing/utilities.py")).read()
    ns = {"np": np, "pd": pd, "griddata": griddata, "logger": MagicMock(),
          "pd_date_range": pd.date_range}
    a = psrc.index("def get_time_index_freqs(")
    exec(compile(psrc[a:psrc.index("\ndef ", a + 5)], "freqs", "exec"), ns)
    fns = {}
    for n in ("pad_lat_lon", "is_increasing_lons", "get_lat_lon", "get_times"):
        exec(compile(grab(wsrc, n), n, "exec"), ns)
        fns[n] = ns[n]

    class RefOutputHandler:
        pad_lat_lon = staticmethod(fns["pad_lat_lon"])
        is_increasing_lons = staticmethod(fns["is_increasing_lons"])
        get_lat_lon = classmethod(fns["get_lat_lon"])
        get_times = staticmethod(fns["get_times"])
    return RefOutputHandler


def grids():
    """regular, rotated curvilinear, and date-line crossing low-res grids (lat descending)."""
    out = {}
    jj, ii = np.meshgrid(np.arange(7), np.arange(5))
    out["regular"] = np.stack([45.0 - 0.25 * ii, -110.0 + 0.25 * jj], -1).astype(np.float32)
    out["rotated"] = np.stack([40.0 - 0.05 * ii + 0.012 * jj + 0.0007 * ii * jj,
                               -105.0 + 0.06 * jj + 0.015 * ii - 0.0005 * jj * jj],
                              -1).astype(np.float32)
    out["dateline"] = np.stack([10.0 - 0.5 * ii, (178.5 + 0.5 * jj + 180) % 360 - 180],
                               -1).astype(np.float32)
    return out


TIMES = {"hourly": (pd.date_range("2015-03-01", periods=6, freq="1h"), 4),
         "daily_no_leap": (pd.date_range("2024-02-26", periods=6, freq="1D").delete(3), 24),
         "3hourly": (pd.date_range("2016-12-31 12:00", periods=5, freq="3h"), 3),
         "single": (pd.date_range("2015-03-01", periods=1, freq="1h"), 12)}


def scenario(H):
    arrs, rec = {}, {}
    for name, ll in grids().items():
        for s in (2, 3):
            shape = (ll.shape[0] * s, ll.shape[1] * s)
            arrs[f"{name}_x{s}"] = np.asarray(H.get_lat_lon(ll.copy(), shape))
    for name, (ti, t_enh) in TIMES.items():
        try:
            t = H.get_times(ti, len(ti) * t_enh)
            rec[name] = [str(v) for v in pd.DatetimeIndex(t)]
        except Exception as e:      # noqa: BLE001
            rec[name] = type(e).__name__
    return rec, arrs


def main():
    rec, arrs = scenario(load_reference())
    np.savez_compressed(OUT, record=json.dumps(rec), **arrs)
    print("wrote", OUT, {k: (v.shape, str(v.dtype)) for k, v in arrs.items()})
    print({k: (v if isinstance(v, str) else (len(v), v[:2], v[-1])) for k, v in rec.items()})


if __name__ == "__main__":
    main()
