"""Helpers shared by the tiler (mirrors sup3r/pipeline/utilities.py:11-58)."""
from __future__ import annotations

import logging

logger = logging.getLogger(__name__)


def get_model(model_class, kwargs):
    """Look the class up in ``sup3r_b200.models`` and ``.load`` it.  A string ``kwargs`` is a
    model directory."""
    from .. import models
    cls = getattr(models, model_class, None)
    if isinstance(kwargs, str):
        kwargs = {"model_dir": kwargs}
    if cls is None:
        msg = (f'Could not load requested model class "{model_class}" from sup3r_b200.models, '
               "Make sure you typed in the model class name correctly.")
        logger.error(msg)
        raise KeyError(msg)
    return cls.load(**dict(kwargs), verbose=True)


def get_chunk_slices(arr_size, chunk_size, index_slice=slice(None)):
    """Consecutive slices of ``chunk_size`` (times the slice step) covering
    ``range(arr_size)[index_slice]``; the last one may be shorter."""
    lo, hi, step = index_slice.indices(arr_size)
    if index_slice.step is None:
        step = 1
    span = step * chunk_size
    out = []
    start = lo
    while start < hi:
        stop = min(start + span, hi)
        out.append(slice(start, stop, step))
        start = stop
    return out
