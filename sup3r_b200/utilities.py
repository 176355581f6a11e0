"""Small host utilities on the hot path (mirrors sup3r/utilities/utilities.py:140-152, 261-335,
345-523, 541-545): ``Timer``, block coarsening, name / json helpers."""
from __future__ import annotations

import logging
import re
import sys
import time

import numpy as np

logger = logging.getLogger(__name__)

RANDOM_GENERATOR = np.random.default_rng(42)  # utilities.py:24


def _versions():
    rec = {"sup3r_b200": "0.1.0", "python": sys.version.split()[0], "numpy": np.__version__}
    try:
        import torch
        rec["torch"] = torch.__version__
    except Exception:  # pragma: no cover
        pass
    try:
        import pandas
        rec["pandas"] = pandas.__version__
    except Exception:  # pragma: no cover
        pass
    return rec


VERSION_RECORD = _versions()


def camel_to_underscore(name):
    """``MeanSquaredError`` -> ``mean_squared_error`` (utilities.py:541-545)."""
    return re.sub(r"(?<!^)(?=[A-Z])", "_", name).lower()


def safe_cast(o):
    """Cast to a json-serialisable type (utilities.py:140-152)."""
    if hasattr(o, "detach"):
        o = o.detach().cpu().numpy()
    if isinstance(o, (float, np.floating)):
        return float(o)
    if isinstance(o, (bool, np.bool_)):
        return bool(o)
    if isinstance(o, (int, np.integer)):
        return int(o)
    if isinstance(o, np.ndarray):
        return o.tolist() if o.ndim else o.item()
    if isinstance(o, tuple):
        return list(o)
    if isinstance(o, (str, list)):
        return o
    return str(o)


class Timer:
    """Wall-clock timer keeping per-function (and per ``call_id``) elapsed times in ``.log``
    (utilities.py:261-335).  Device time is measured separately with CUDA events."""

    def __init__(self):
        self.log = {}
        self._start = None
        self._stop = None

    def start(self):
        self._start = time.time()
        self._stop = None

    def stop(self):
        self._stop = time.time()

    @property
    def elapsed(self):
        end = time.time() if self._stop is None else self._stop
        return end - self._start

    @property
    def elapsed_str(self):
        return f"{round(self.elapsed, 5)} seconds"

    def __call__(self, func, call_id=None, log=False):
        def wrapper(*args, **kwargs):
            self.start()
            out = func(*args, **kwargs)
            self.stop()
            if call_id is not None:
                self.log.setdefault(call_id, {})[func.__name__] = self.elapsed
            else:
                self.log[func.__name__] = self.elapsed
            if log:
                logger.debug("Call to %s finished in %s", func.__name__, self.elapsed_str)
            return out

        return wrapper


def _block_reduce(data, axis, factor, how):
    shp = list(data.shape)
    if shp[axis] % factor:
        raise ValueError(f"enhancement factor {factor} must evenly divide axis {axis} of "
                         f"{tuple(data.shape)}")
    new = shp[:axis] + [shp[axis] // factor, factor] + shp[axis + 1:]
    return how(np.reshape(data, new), axis=axis + 1)


def temporal_coarsening(data, t_enhance=4, method="subsample"):
    """Coarsen axis 3 of a 5-D array (utilities.py:345-403)."""
    if t_enhance is None or data.ndim != 5:
        return data
    if method == "subsample":
        return data[:, :, :, ::t_enhance, :]
    hows = {"average": lambda a, axis: np.nansum(a, axis=axis) / t_enhance,
            "total": np.nansum, "max": np.max, "min": np.min}
    if method not in hows:
        msg = (f'Did not recognize temporal_coarsening method "{method}", can only accept one '
               "of: [subsample, average, total, max, min]")
        logger.error(msg)
        raise KeyError(msg)
    return _block_reduce(data, 3, t_enhance, hows[method])


def spatial_coarsening(data, s_enhance=2, obs_axis=True):
    """Block-average the two spatial axes (utilities.py:406-523)."""
    if data.ndim < 2 or (obs_axis and data.ndim < 3):
        msg = f"Data has too few dims for spatial coarsening (obs_axis={obs_axis}): {data.shape}"
        logger.error(msg)
        raise ValueError(msg)
    if s_enhance is None or s_enhance <= 1:
        return data
    a0 = 1 if obs_axis else 0
    if data.shape[a0] % s_enhance or data.shape[a0 + 1] % s_enhance:
        msg = (f"s_enhance must evenly divide grid size. Received s_enhance: {s_enhance} with "
               f"data shape: {data.shape}")
        logger.error(msg)
        raise ValueError(msg)
    out = _block_reduce(data, a0, s_enhance, np.sum)
    out = _block_reduce(out, a0 + 1, s_enhance, np.sum)
    return out / s_enhance ** 2
