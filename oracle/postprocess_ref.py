"""numpy restatement of the reference's chunk post-processing and batch production.  TEST
INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED: ``tests/golden/postprocess.npz`` holds
outputs of the REAL reference functions (``tools/make_golden_postprocess.py`` executes their
source from /root/reference), ``tests/test_postprocess.py`` checks this file against them.

  invert_uv            sup3r/preprocessing/derivers/utilities.py:204-255
  enforce_limits       sup3r/utilities/utilities.py:155-220 (+ nn_fill_array :55-75)
  get_renamed_features sup3r/writers/base.py:205-229
  smooth_data          sup3r/preprocessing/batch_queues/utilities.py:57-104
  transform            sup3r/preprocessing/batch_queues/base.py:32-87
"""
from __future__ import annotations

import re

import numpy as np

# sup3r/utilities/output_attrs.json: (min, max) per feature basename
OUTPUT_LIMITS = {
    "u": (-120, 120), "v": (-120, 120), "windspeed": (0, 120), "winddirection": (0, 360),
    "clearsky_ratio": (0, 1), "dhi": (0, 1350), "dni": (0, 1350), "ghi": (0, 1350),
    "rsds": (0, 1350), "temperature": (-200, 100), "temperature_min": (-200, 100),
    "temperature_max": (-200, 100), "relativehumidity": (0, 100),
    "relativehumidity_min": (0, 100), "relativehumidity_max": (0, 100),
    "pressure": (0, 150000), "pr": (0, np.inf), "srl": (0, np.inf)}


def get_feature_basename(feature):
    """utilities.py:78-92"""
    height = re.findall(r"_\d+m", feature)
    press = re.findall(r"_\d+pa", feature)
    if height:
        return feature.replace(height[0], "")
    if press:
        return feature.replace(press[0], "")
    if "_(.*)" in feature:
        return feature.split("_(.*)")[0]
    return feature


def grid_angle(lat_lon):
    """Angle of the grid's vertical from north (derivers/utilities.py:237-243), on the
    latitude-descending orientation the reference flips to."""
    dy = lat_lon[:, :, 0] - np.roll(lat_lon[:, :, 0], 1, axis=0)
    dx = lat_lon[:, :, 1] - np.roll(lat_lon[:, :, 1], 1, axis=0)
    dy = (dy + 90) % 180 - 90
    dx = (dx + 180) % 360 - 180
    theta = (np.pi / 2) - np.arctan2(dy, dx)
    if len(theta) > 1:
        theta[0] = theta[1]
    return theta


def invert_uv(u, v, lat_lon):
    """u, v (s1, s2, t) -> windspeed, winddirection [deg, clockwise from north]."""
    invert_lat = False
    if lat_lon[-1, 0, 0] > lat_lon[0, 0, 0]:
        invert_lat = True
        lat_lon, u, v = lat_lon[::-1], u[::-1], v[::-1]
    theta = grid_angle(lat_lon)
    u_rot = np.cos(theta)[:, :, np.newaxis] * u - np.sin(theta)[:, :, np.newaxis] * v
    v_rot = np.sin(theta)[:, :, np.newaxis] * u + np.cos(theta)[:, :, np.newaxis] * v
    ws = np.hypot(u_rot, v_rot)
    wd = (np.degrees(np.arctan2(u_rot, v_rot)) + 360) % 360
    if invert_lat:
        ws, wd = ws[::-1], wd[::-1]
    return ws, wd


def get_renamed_features(features):
    """u_{h}m / v_{h}m -> windspeed_{h}m / winddirection_{h}m (writers/base.py:205-229)."""
    out = list(features)
    for f in features:
        m = re.match(r"u_(\d+)m$", f.lower())
        if m and f"v_{m.group(1)}m" in features:
            out[features.index(f)] = f"windspeed_{m.group(1)}m"
            out[features.index(f"v_{m.group(1)}m")] = f"winddirection_{m.group(1)}m"
    return out


def nn_fill_array(array):
    from scipy import ndimage as nd
    nan_mask = np.isnan(array)
    idx = nd.distance_transform_edt(nan_mask, return_distances=False, return_indices=True)
    return array[tuple(idx)]


def enforce_limits(features, data, nn_fill=False):
    data = np.array(data, copy=True)
    for fidx, fn in enumerate(features):
        base = get_feature_basename(fn)
        if base not in OUTPUT_LIMITS:
            raise KeyError(f'Could not find "{base}" in OUTPUT_ATTRS dict!')
        lo, hi = OUTPUT_LIMITS[base]
        if nn_fill:
            d = data[..., fidx]
            d = np.where(d > hi, np.nan, d)
            d = np.where(d < lo, np.nan, d)
            data[..., fidx] = nn_fill_array(d)
        else:
            data[..., fidx] = np.minimum(np.maximum(data[..., fidx], lo), hi)
    return data.astype(np.float32)


def transform_output(data, features, lat_lon, invert=False, nn_fill=False):
    """writers/base.py:297-346"""
    data = np.array(data, copy=True)
    if invert:
        for f in features:
            m = re.match(r"u_(\d+)m$", f.lower())
            if m:
                iu, iv = features.index(f), features.index(f"v_{m.group(1)}m")
                ws, wd = invert_uv(data[..., iu], data[..., iv], lat_lon)
                data[..., iu], data[..., iv] = ws, wd
        features = get_renamed_features(features)
    return enforce_limits(features, data, nn_fill), features


def smooth_data(low_res, training_features, smoothing_ignore, smoothing=None):
    from scipy.ndimage import gaussian_filter
    low_res = np.array(low_res, copy=True)
    if smoothing is None:
        return low_res
    feats = [j for j in range(low_res.shape[-1]) if training_features[j] not in smoothing_ignore]
    for i in range(low_res.shape[0]):
        for j in feats:
            if low_res.ndim == 5:
                for t in range(low_res.shape[-2]):
                    low_res[i, ..., t, j] = gaussian_filter(low_res[i, ..., t, j], smoothing,
                                                            mode="nearest")
            else:
                low_res[i, ..., j] = gaussian_filter(low_res[i, ..., j], smoothing, mode="nearest")
    return low_res


def spatial_coarsening(data, s_enhance=2):
    """block mean over (s, s) (utilities.py:406-523, obs axis first)"""
    if s_enhance == 1:
        return data
    b, s1, s2 = data.shape[:3]
    rest = data.shape[3:]
    d = data.reshape(b, s1 // s_enhance, s_enhance, s2 // s_enhance, s_enhance, *rest)
    return d.sum(axis=(2, 4)) / (s_enhance * s_enhance)


def temporal_coarsening(data, t_enhance=4, method="subsample"):
    """utilities.py:345-403 (5-D)"""
    if t_enhance == 1 or data.ndim != 5:
        return data
    b, s1, s2, t, f = data.shape
    if method == "subsample":
        return data[:, :, :, ::t_enhance, :]
    d = data.reshape(b, s1, s2, t // t_enhance, t_enhance, f)
    if method == "average":
        return np.nansum(d, axis=4) / t_enhance
    if method == "max":
        return np.max(d, axis=4)
    if method == "min":
        return np.min(d, axis=4)
    if method == "total":
        return np.nansum(d, axis=4)
    raise ValueError(method)


def batch_transform(samples, s_enhance, t_enhance, features, smoothing=None, smoothing_ignore=None,
                    temporal_coarsening_method="subsample"):
    low = spatial_coarsening(samples, s_enhance)
    low = low if t_enhance == 1 else temporal_coarsening(low, t_enhance, temporal_coarsening_method)
    return smooth_data(low, features, smoothing_ignore or [], smoothing)
