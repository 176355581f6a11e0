"""Chunk index math of the forward pass (mirrors sup3r/pipeline/slicer.py:19-716).

The low-res domain ``(s1, s2, t)`` is cut into ``chunk_shape`` tiles; every tile is widened by
``spatial_pad`` / ``temporal_pad`` (clipped at the domain edge, the missing part is REFLECT
padded by the caller), run through the generator, and cropped back by ``pad * enhance`` on the
high-res side.  Chunk index = ``t_idx * n_spatial + s_idx``.  Pure integer math (numpy only).
"""
from __future__ import annotations

import itertools as it
import logging
from dataclasses import dataclass
from typing import Optional, Union
from warnings import warn

import numpy as np

from .utilities import get_chunk_slices

logger = logging.getLogger(__name__)


def _parse_time_slice(value):
    if isinstance(value, slice):
        return value
    if isinstance(value, (tuple, list)):
        return slice(*value)
    return slice(None)


@dataclass
class ForwardPassSlicer:
    """Slices for chunking, padding and cropping forward-pass data."""

    coarse_shape: Union[tuple, list]
    time_steps: int
    s_enhance: int
    t_enhance: int
    time_slice: slice
    temporal_pad: int
    spatial_pad: int
    chunk_shape: Union[tuple, list]
    min_width: Optional[Union[tuple, list]] = None

    def __post_init__(self):
        self.dummy_time_index = np.arange(self.time_steps)
        self.time_slice = _parse_time_slice(self.time_slice)
        self.min_width = self.chunk_shape if self.min_width is None else self.min_width
        self._cache = {}

    def _memo(self, key, fn):
        if key not in self._cache:
            self._cache[key] = fn()
        return self._cache[key]

    # ---- low-res slices ----------------------------------------------------------------
    def _s_lr(self, dim):
        n = self.coarse_shape[dim]
        return get_chunk_slices(n, self.chunk_shape[dim], index_slice=slice(0, n))

    @property
    def s1_lr_slices(self):
        return self._s_lr(0)

    @property
    def s2_lr_slices(self):
        return self._s_lr(1)

    @property
    def t_lr_slices(self):
        """Even temporal split (``np.array_split``) of the sliced time index."""
        idx = self.dummy_time_index[self.time_slice]
        n_chunks = int(np.ceil(len(idx) / self.chunk_shape[2]))
        return [slice(int(c[0]), int(c[-1]) + 1, self.time_slice.step)
                for c in np.array_split(idx, n_chunks)]

    @staticmethod
    def get_padded_slices(slices, shape, enhancement, padding, step=None):
        """Widen each slice by ``padding`` (x enhancement x step), clipped to the domain."""
        step = step or 1
        pad = step * padding * enhancement
        return [slice(int(max(0, s.start * enhancement - pad)),
                      int(min(enhancement * shape, s.stop * enhancement + pad)), step)
                for s in slices]

    @property
    def s1_lr_pad_slices(self):
        return self._memo("s1p", lambda: self.get_padded_slices(
            self.s1_lr_slices, self.coarse_shape[0], 1, self.spatial_pad))

    @property
    def s2_lr_pad_slices(self):
        return self._memo("s2p", lambda: self.get_padded_slices(
            self.s2_lr_slices, self.coarse_shape[1], 1, self.spatial_pad))

    @property
    def t_lr_pad_slices(self):
        return self._memo("tp", lambda: self.get_padded_slices(
            self.t_lr_slices, self.time_steps, 1, self.temporal_pad, step=self.time_slice.step))

    @property
    def s_lr_slices(self):
        return self._memo("s", lambda: list(it.product(self.s1_lr_slices, self.s2_lr_slices)))

    @property
    def s_lr_pad_slices(self):
        return self._memo("sp", lambda: list(it.product(self.s1_lr_pad_slices,
                                                        self.s2_lr_pad_slices)))

    def get_spatial_slices(self):
        return self.s_lr_slices, self.s_lr_pad_slices, self.s_hr_slices

    def get_time_slices(self):
        return self.t_lr_slices, self.t_lr_pad_slices

    # ---- cropping ------------------------------------------------------------------------
    @staticmethod
    def get_cropped_slices(unpadded_slices, padded_slices, enhancement):
        """Slices that cut the padded tile back to the unpadded tile (relative indices)."""
        out = []
        for ps, us in zip(padded_slices, unpadded_slices):
            step = us.step or 1
            start = None if us.start is None else enhancement * (us.start - ps.start) // step
            stop = None if us.stop is None else enhancement * (us.stop - ps.stop) // step
            if start is not None and start <= 0:
                start = None
            if stop is not None and stop >= 0:
                stop = None
            out.append(slice(start, stop))
        return out

    def check_boundary_slice(self, unpadded_slices, cropped_slices, enhancement, padding, dim):
        """If the (padded) last tile of a dim is narrower than ``min_width`` it gets extra
        padding (see ``_get_pad_width``); crop that extra part off again."""
        last = unpadded_slices[-1]
        lo = last.start or 0
        hi = last.stop or self.coarse_shape[dim]
        if 2 * padding + hi - lo < self.min_width[dim]:
            half = self.min_width[dim] // 2 + 1
            msg = (f"The final slice for dimension #{dim + 1} is too small (slice=slice({lo}, "
                   f"{hi}), padding={padding}). The start of this slice will be reduced to try "
                   "to meet the minimum slice length.")
            logger.warning(msg)
            warn(msg)
            cropped_slices[-1] = slice(half * enhancement, -half * enhancement)
        return cropped_slices

    def _s_hr_crop(self, dim, lr_slices):
        start = self.s_enhance * self.spatial_pad or None
        stop = None if self.spatial_pad == 0 else -start
        crops = [slice(start, stop)] * len(lr_slices)
        return self.check_boundary_slice(lr_slices, crops, self.s_enhance, self.spatial_pad, dim)

    @property
    def s1_hr_crop_slices(self):
        return self._memo("s1c", lambda: self._s_hr_crop(0, self.s1_lr_slices))

    @property
    def s2_hr_crop_slices(self):
        return self._memo("s2c", lambda: self._s_hr_crop(1, self.s2_lr_slices))

    @property
    def s_hr_crop_slices(self):
        return self._memo("sc", lambda: list(it.product(self.s1_hr_crop_slices,
                                                        self.s2_hr_crop_slices)))

    @property
    def t_lr_crop_slices(self):
        return self._memo("tlc", lambda: self.get_cropped_slices(self.t_lr_slices,
                                                                 self.t_lr_pad_slices, 1))

    @property
    def t_hr_crop_slices(self):
        start = stop = None
        if self.temporal_pad > 0:
            start = self.t_enhance * self.temporal_pad
            stop = -start
        return self._memo("thc", lambda: [slice(start, stop) for _ in self.t_lr_slices])

    @property
    def s_lr_crop_slices(self):
        def build():
            c1 = self.get_cropped_slices(self.s1_lr_slices, self.s1_lr_pad_slices, 1)
            c1 = self.check_boundary_slice(self.s1_lr_slices, c1, self.s_enhance,
                                           self.spatial_pad, 0)
            c2 = self.get_cropped_slices(self.s2_lr_slices, self.s2_lr_pad_slices, 1)
            c2 = self.check_boundary_slice(self.s2_lr_slices, c2, self.s_enhance,
                                           self.spatial_pad, 1)
            return list(it.product(c1, c2))
        return self._memo("slc", build)

    @property
    def hr_crop_slices(self):
        """[time chunk][spatial chunk] -> (s1, s2, t, features) crop of the generator output."""
        return self._memo("hrc", lambda: [
            [(s[0], s[1], t, slice(None)) for s in self.s_hr_crop_slices]
            for t in self.t_hr_crop_slices])

    # ---- high-res slices -------------------------------------------------------------------
    @staticmethod
    def get_hr_slices(slices, enhancement, step=None):
        if step is not None:
            step *= enhancement
        return [slice(s.start * enhancement, s.stop * enhancement, step) for s in slices]

    @property
    def s1_hr_slices(self):
        return self.get_hr_slices(self.s1_lr_slices, self.s_enhance)

    @property
    def s2_hr_slices(self):
        return self.get_hr_slices(self.s2_lr_slices, self.s_enhance)

    @property
    def s_hr_slices(self):
        return self._memo("shr", lambda: list(it.product(self.s1_hr_slices, self.s2_hr_slices)))

    # ---- chunk bookkeeping -------------------------------------------------------------------
    @property
    def n_spatial_chunks(self):
        return len(self.hr_crop_slices[0])

    @property
    def n_time_chunks(self):
        return len(self.t_hr_crop_slices)

    @property
    def n_chunks(self):
        return self.n_spatial_chunks * self.n_time_chunks

    @property
    def chunk_lookup(self):
        """(n_s1, n_s2, n_t) array of chunk indices."""
        def build():
            n1, n2 = len(self.s1_lr_slices), len(self.s2_lr_slices)
            lookup = np.arange(self.n_chunks).reshape((self.n_time_chunks, n1, n2))
            return np.transpose(lookup, axes=(1, 2, 0))
        return self._memo("lookup", build)

    @property
    def spatial_chunk_lookup(self):
        return np.arange(self.n_spatial_chunks).reshape((len(self.s1_lr_slices),
                                                         len(self.s2_lr_slices)))

    def get_chunk_indices(self, chunk_index):
        """-> (spatial chunk index, temporal chunk index)"""
        return chunk_index % self.n_spatial_chunks, chunk_index // self.n_spatial_chunks

    # ---- reflect padding needed at domain edges -------------------------------------------------
    @staticmethod
    def _get_pad_width(window, max_steps, max_pad, min_width=None, check_boundary=False):
        lo = window.start or 0
        hi = window.stop or max_steps
        start = int(max(0, max_pad - lo))
        stop = int(max(0, max_pad + hi - max_steps))
        too_small = min_width is not None and (2 * max_pad + hi - lo) < min_width
        if check_boundary and hi == max_steps and too_small:
            half = min_width // 2 + 1
            start = stop = int(max(half, max_pad))
        return start, stop

    def get_pad_width(self, chunk_index):
        """((s1 lo, hi), (s2 lo, hi), (t lo, hi)) extra padding for one chunk."""
        s_idx, t_idx = self.get_chunk_indices(chunk_index)
        lr = self.s_lr_slices[s_idx]
        return (self._get_pad_width(lr[0], self.coarse_shape[0], self.spatial_pad,
                                    self.min_width[0], check_boundary=True),
                self._get_pad_width(lr[1], self.coarse_shape[1], self.spatial_pad,
                                    self.min_width[1], check_boundary=True),
                self._get_pad_width(self.t_lr_slices[t_idx], len(self.dummy_time_index),
                                    self.temporal_pad))

    @property
    def extra_padding(self):
        return self._memo("extra", lambda: [self.get_pad_width(i) for i in range(self.n_chunks)])
