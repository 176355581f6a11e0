"""Golden vectors for the chunk post-processing / batch production rows
from the REAL reference functions: their source text is exec'd from /root/reference (the
modules themselves import dask / xarray / rex, which are not installed).

    python tools/make_golden_postprocess.py   ->  tests/golden/postprocess.npz
"""
import json
import os
import re
from unittest.mock import MagicMock

import numpy as np
from scipy import ndimage as nd
from scipy.ndimage import gaussian_filter

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def grab(src, fn, ns):
    a = src.index(f"def {fn}(")
    b = src.find("\ndef ", a + 10)
    exec(compile(src[a:(len(src) if b < 0 else b)], fn, "exec"), ns)


def main():
    usrc = open(os.path.join(REF, "sup3r/utilities/utilities.py")).read()
    dsrc = open(os.path.join(REF, "sup3r/preprocessing/derivers/utilities.py")).read()
    bsrc = open(os.path.join(REF, "sup3r/preprocessing/batch_queues/utilities.py")).read()
    attrs = json.load(open(os.path.join(REF, "sup3r/utilities/output_attrs.json")))
    ns = {"np": np, "nd": nd, "re": re, "logger": MagicMock(), "warn": lambda *a, **k: None,
          "OUTPUT_ATTRS": attrs, "gaussian_filter": gaussian_filter}
    for fn in ("nn_fill_array", "get_feature_basename", "enforce_limits", "spatial_coarsening",
               "temporal_coarsening"):
        grab(usrc, fn, ns)
    grab(dsrc, "invert_uv", ns)
    grab(bsrc, "smooth_data", ns)
    rng = np.random.default_rng(42)
    res = {}
    # ---- u / v inversion on a curvilinear (rotated) grid, both latitude orientations
    s1, s2, t = 12, 10, 5
    jj, ii = np.meshgrid(np.arange(s2), np.arange(s1))
    lat = 40.0 - 0.05 * ii + 0.01 * jj
    lon = -105.0 + 0.06 * jj + 0.012 * ii
    lat_lon = np.stack([lat, lon], axis=-1).astype(np.float32)
    u = rng.standard_normal((s1, s2, t)).astype(np.float32) * 8
    v = rng.standard_normal((s1, s2, t)).astype(np.float32) * 8
    res["lat_lon"], res["u"], res["v"] = lat_lon, u, v
    for tag, ll in (("desc", lat_lon), ("asc", lat_lon[::-1].copy())):
        ws, wd = ns["invert_uv"](u.copy(), v.copy(), ll.copy())
        res[f"ws_{tag}"], res[f"wd_{tag}"] = np.asarray(ws), np.asarray(wd)
    # ---- limits
    feats = ["windspeed_100m", "winddirection_100m", "temperature_2m", "relativehumidity_2m",
             "pressure_0m", "clearsky_ratio"]
    data = rng.standard_normal((s1, s2, t, len(feats))).astype(np.float32)
    data *= np.array([80, 300, 120, 90, 9e4, 1.2], np.float32)
    res["lim_in"] = data
    res["lim_clip"] = ns["enforce_limits"](feats, data.copy(), nn_fill=False)
    res["lim_nn"] = ns["enforce_limits"](feats, data.copy(), nn_fill=True)
    res["lim_features"] = np.array(feats)
    # ---- batch production: coarsening + smoothing
    hr = rng.standard_normal((3, 12, 16, 8, 2)).astype(np.float32)
    res["hr"] = hr
    for m in ("subsample", "average", "total", "max", "min"):
        low = ns["temporal_coarsening"](ns["spatial_coarsening"](hr, 4), 2, m)
        res[f"low_{m}"] = np.asarray(low)
        res[f"low_{m}_smooth"] = ns["smooth_data"](np.array(low, copy=True), ["u", "v"], ["v"],
                                                    smoothing=0.7)
    hr4 = rng.standard_normal((2, 12, 16, 2)).astype(np.float32)
    res["hr4"] = hr4
    res["low4_smooth"] = ns["smooth_data"](np.array(ns["spatial_coarsening"](hr4, 2), copy=True),
                                           ["u", "v"], [], smoothing=1.3)
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), **res)
    print("written", os.path.join(OUT, "postprocess.npz"))


if __name__ == "__main__":
    main()


def bias_golden():
    """Linear bias-correction transforms: the reference function bodies with the h5 factor reader
    (`_get_spatial_bc_factors`, rex) replaced by in-memory arrays -> tests/golden/bias.npz"""
    from unittest.mock import MagicMock as MM
    src = open(os.path.join(REF, "sup3r/bias/bias_transforms.py")).read()
    rng = np.random.default_rng(3)
    s1, s2, t = 6, 5, 8
    fac = {"scalar": (1 + 0.2 * rng.standard_normal((s1, s2, 12))).astype(np.float32),
           "adder": rng.standard_normal((s1, s2, 12)).astype(np.float32)}
    months = np.array([1, 1, 1, 2, 2, 2, 2, 2])

    class TI:
        class month:
            values = months

            @staticmethod
            def unique():
                return np.unique(months)

    ns = {"np": np, "logger": MM(), "warn": lambda *a, **k: None, "gaussian_filter": gaussian_filter,
          "_get_spatial_bc_factors": lambda *a, **k: {k2: v.copy() for k2, v in fac.items()},
          "make_time_index_from_kws": lambda kw: TI}
    for fn in ("global_linear_bc", "local_linear_bc", "monthly_local_linear_bc"):
        grab(src, fn, ns)
    data = rng.standard_normal((s1, s2, t)).astype(np.float32)
    sl = (slice(1, 5), slice(0, 4), slice(None))
    res = {"data": data, "scalar": fac["scalar"], "adder": fac["adder"], "months": months,
           "global": ns["global_linear_bc"](data, 1.1, -0.3, out_range=(-1, 1)),
           "local": ns["local_linear_bc"](data, None, "u", "fp", smoothing=0.8, out_range=(-2, 2)),
           "local_slice": ns["local_linear_bc"](data[sl[0], sl[1]], None, "u", "fp",
                                                lr_padded_slice=sl),
           "monthly_avg": ns["monthly_local_linear_bc"](data, None, "u", "fp", {}, temporal_avg=True,
                                                        scalar_range=(0.8, 1.2)),
           "monthly": ns["monthly_local_linear_bc"](data, None, "u", "fp", {}, temporal_avg=False,
                                                    smoothing=0.5, adder_range=(-1, 1))}
    np.savez_compressed(os.path.join(OUT, "bias.npz"), **res)
    print("written", os.path.join(OUT, "bias.npz"))


if __name__ == "__main__":
    bias_golden()
