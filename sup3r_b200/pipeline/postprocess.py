"""On-device post-processing of forward-pass chunks: the transforms the reference
applies on the host before writing a chunk -- ``OutputHandler._transform_output``
(sup3r/writers/base.py:297-346): u / v -> windspeed / winddirection on the rotated grid
(``invert_uv``, preprocessing/derivers/utilities.py:204-255), feature renaming
(writers/base.py:205-229) and physical limits (``enforce_limits``, utilities/utilities.py:155-220).
Here they run as one in-place HBM pass on the cropped chunk while it is still on the GPU
(``s3_output_transform``).  ``nn_fill`` keeps the reference's host algorithm (scipy EDT) and only
runs when the device-side counts say some value is out of range."""
from __future__ import annotations

import logging
import re
from warnings import warn

import numpy as np
import torch

from .. import ops

logger = logging.getLogger(__name__)

# sup3r/utilities/output_attrs.json: (min, max) per feature basename
OUTPUT_LIMITS = {
    "u": (-120, 120), "v": (-120, 120), "windspeed": (0, 120), "winddirection": (0, 360),
    "clearsky_ratio": (0, 1), "dhi": (0, 1350), "dni": (0, 1350), "ghi": (0, 1350),
    "rsds": (0, 1350), "temperature": (-200, 100), "temperature_min": (-200, 100),
    "temperature_max": (-200, 100), "relativehumidity": (0, 100),
    "relativehumidity_min": (0, 100), "relativehumidity_max": (0, 100),
    "pressure": (0, 150000), "pr": (0, np.inf), "srl": (0, np.inf)}


def get_feature_basename(feature):
    """Feature name without its height / pressure suffix (utilities/utilities.py:78-92)."""
    height = re.findall(r"_\d+m", feature)
    press = re.findall(r"_\d+pa", feature)
    if height:
        return feature.replace(height[0], "")
    if press:
        return feature.replace(press[0], "")
    if "_(.*)" in feature:
        return feature.split("_(.*)")[0]
    return feature


def uv_pairs(features):
    """[(u index, v index, height)] of the u_{h}m / v_{h}m pairs (writers/base.py:252-264)."""
    out = []
    low = [f.lower() for f in features]
    for i, f in enumerate(low):
        m = re.match(r"u_(\d+)m$", f)
        if m and f"v_{m.group(1)}m" in low:
            out.append((i, low.index(f"v_{m.group(1)}m"), int(m.group(1))))
    return out


def get_renamed_features(features):
    """u / v names -> windspeed / winddirection names (writers/base.py:205-229)."""
    out = list(features)
    for iu, iv, h in uv_pairs(features):
        out[iu], out[iv] = f"windspeed_{h}m", f"winddirection_{h}m"
    return out


def grid_rotation(lat_lon):
    """(s1, s2, 2) float32 (cos theta, sin theta): the angle of the grid's vertical from north,
    computed like ``invert_uv`` (derivers/utilities.py:228-243) incl. its latitude flip."""
    lat_lon = np.asarray(lat_lon)
    flip = lat_lon[-1, 0, 0] > lat_lon[0, 0, 0]
    ll = lat_lon[::-1] if flip else lat_lon
    dy = ll[:, :, 0] - np.roll(ll[:, :, 0], 1, axis=0)
    dx = ll[:, :, 1] - np.roll(ll[:, :, 1], 1, axis=0)
    dy = (dy + 90) % 180 - 90
    dx = (dx + 180) % 360 - 180
    theta = (np.pi / 2) - np.arctan2(dy, dx)
    if len(theta) > 1:
        theta[0] = theta[1]
    if flip:
        theta = theta[::-1]
    return np.ascontiguousarray(np.stack([np.cos(theta), np.sin(theta)], axis=-1), np.float32)


def feature_limits(features):
    lo, hi = [], []
    for fn in features:
        base = get_feature_basename(fn)
        if base not in OUTPUT_LIMITS:
            msg = f'Could not find "{base}" in OUTPUT_ATTRS dict!'
            logger.error(msg)
            raise KeyError(msg)
        a, b = OUTPUT_LIMITS[base]
        lo.append(float(a))
        hi.append(float(b))
    return lo, hi


def nn_fill_array(array):
    """NaNs <- nearest non-NaN value (utilities/utilities.py:55-75)."""
    from scipy import ndimage as nd
    idx = nd.distance_transform_edt(np.isnan(array), return_distances=False, return_indices=True)
    return array[tuple(idx)]


def transform_output(data, features, lat_lon, invert_uv=False, nn_fill=False):
    """``OutputHandler._transform_output`` on a device chunk ``data`` (s1, s2, t, f), in place.
    Returns ``(data, features)`` with the renamed features.  ``lat_lon``: (s1, s2, 2) hi-res grid
    (host array) -- only read when there are u / v pairs to invert."""
    if not isinstance(data, torch.Tensor) or not data.is_cuda:
        raise RuntimeError("transform_output works on a CUDA tensor (no CPU fallback)")
    features = list(features)
    pairs = uv_pairs(features) if invert_uv else []
    cs = None
    if pairs:
        logger.info("Converting u/v to ws/wd for %d heights", len(pairs))
        cs = torch.from_numpy(grid_rotation(lat_lon)).to(data.device)
        features = get_renamed_features(features)
    lo, hi = feature_limits(features)
    # nn_fill: first pass counts (and inverts u / v) without clipping; the fill itself only
    # runs when something is out of range
    counts = ops.output_transform(data, cs, [(p[0], p[1]) for p in pairs], lo, hi,
                                  clip=not nn_fill).cpu().numpy()
    n = data[..., 0].numel()
    for i, fn in enumerate(features):
        how = " with nearest neighbor interpolation." if nn_fill else " with clipping."
        if counts[i, 1]:
            msg = (f"{fn} has {counts[i, 1] / n:.4e} of points above the max of {hi[i]}. "
                   f'Enforcing range of ({lo[i]}, {hi[i]}) for "{fn}"' + how)
            logger.warning(msg)
            warn(msg)
        if counts[i, 0]:
            msg = (f"{fn} has {counts[i, 0] / n:.4e} of points below the min of {lo[i]}. "
                   f'Enforcing range of ({lo[i]}, {hi[i]}) for "{fn}"' + how)
            logger.warning(msg)
            warn(msg)
    if nn_fill and counts.any():
        host = data.cpu().numpy()
        for i in np.where(counts.any(axis=1))[0]:
            d = host[..., i]
            d = np.where(d > hi[i], np.nan, d)
            d = np.where(d < lo[i], np.nan, d)
            host[..., i] = nn_fill_array(d)
        data.copy_(torch.from_numpy(host))
    return data, features
