"""``ForwardPassStrategy``: chunk bookkeeping for the forward pass (mirrors
sup3r/pipeline/strategy.py:37-700).

The reference reads its low-res source through xarray / dask data handlers and writes h5 / nc
chunk files; both are out of scope here.  The input side is an in-memory
``ArrayInputHandler`` exposing exactly the attributes the tiler consumes (``data``,
``features``, ``grid_shape``, ``time_index``, ``lat_lon``); ``file_paths`` may also name
``.npy`` / ``.npz`` files.  Chunk outputs are returned in memory or written as ``.npy`` files
named ``{file_id}`` = ``{t:06d}_{s:06d}`` so that the ``incremental`` restart logic is kept.
"""
from __future__ import annotations

import logging
import os
import pprint
from dataclasses import dataclass, field
from functools import cached_property
from typing import Optional, Union
from warnings import warn

import numpy as np

from ..exo import ExoData
from ..utilities import Timer
from .slicer import ForwardPassSlicer, _parse_time_slice
from .utilities import get_model

logger = logging.getLogger(__name__)


class ArrayInputHandler:
    """In-memory stand-in for the reference's input DataHandler."""

    def __init__(self, data, features, lat_lon=None, time_index=None, time_slice=slice(None)):
        data = np.asarray(data, dtype=np.float32)
        if data.ndim != 4:
            raise ValueError("ArrayInputHandler needs (spatial_1, spatial_2, temporal, features) "
                             f"data, got shape {data.shape}")
        if data.shape[-1] != len(features):
            raise ValueError(f"{len(features)} feature names for {data.shape[-1]} channels")
        time_slice = _parse_time_slice(time_slice)
        self.features = list(features)
        self.data = data[:, :, time_slice]
        n_t = self.data.shape[2]
        self.time_index = np.arange(n_t) if time_index is None \
            else np.asarray(time_index)[time_slice]
        if lat_lon is None:
            lat = np.linspace(40.0, 39.0, data.shape[0], dtype=np.float32)
            lon = np.linspace(-105.0, -104.0, data.shape[1], dtype=np.float32)
            lat_lon = np.stack(np.meshgrid(lat, lon, indexing="ij"), axis=-1)
        self.lat_lon = np.asarray(lat_lon, dtype=np.float32)

    @property
    def grid_shape(self):
        return tuple(self.data.shape[:2])

    @property
    def shape(self):
        return self.data.shape

    def get(self, features, s1, s2, t):
        idx = [self.features.index(f) for f in features]
        return np.ascontiguousarray(self.data[s1, s2, t][..., idx])

    @classmethod
    def from_files(cls, file_paths, features, **kwargs):
        if isinstance(file_paths, (str, os.PathLike)):
            file_paths = [file_paths]
        arrs = []
        for fp in file_paths:
            fp = str(fp)
            if fp.endswith(".npz"):
                with np.load(fp) as z:
                    arrs.append(np.stack([z[f] for f in features], axis=-1))
            elif fp.endswith(".npy"):
                arrs.append(np.load(fp))
            else:
                raise ValueError(f"ArrayInputHandler reads .npy / .npz files, got {fp} "
                                 "(h5 / netCDF sources are out of scope)")
        data = arrs[0] if len(arrs) == 1 else np.concatenate(arrs, axis=2)
        return cls(data, features, **kwargs)


@dataclass
class ForwardPassChunk:
    """Everything needed to run the generator on one chunk."""

    input_data: np.ndarray
    exo_data: Optional[dict]
    hr_crop_slice: tuple
    lr_pad_slice: tuple
    hr_lat_lon: np.ndarray
    hr_times: np.ndarray
    gids: np.ndarray
    out_file: Optional[str]
    pad_width: tuple
    index: int

    def __post_init__(self):
        self.shape = self.input_data.shape


def _extend_grid(lat_lon):
    """The low-res grid with one extra row / column on every side, continued with the spacing of
    the outermost cells (writers/base.py:347-421): new columns keep their row's latitude and
    step the longitude, new rows keep their column's longitude and step the latitude, corners
    take the latitude of their row neighbour and the longitude of their column neighbour."""
    n1, n2 = lat_lon.shape[:2]
    g = np.zeros((n1 + 2, n2 + 2, 2))
    g[1:-1, 1:-1] = lat_lon
    lat, lon = g[..., 0], g[..., 1]
    d_left, d_right = lon[:, 2] - lon[:, 1], lon[:, -2] - lon[:, -3]
    d_top, d_bottom = lat[1, :] - lat[2, :], lat[-3, :] - lat[-2, :]
    lon[:, 0], lat[:, 0] = lon[:, 1] - d_left, lat[:, 1]
    lon[:, -1], lat[:, -1] = lon[:, -2] + d_right, lat[:, -2]
    lat[0, :], lon[0, :] = lat[1, :] + d_top, lon[1, :]
    lat[-1, :], lon[-1, :] = lat[-2, :] - d_bottom, lon[-2, :]
    for r, rn in ((0, 1), (-1, -2)):
        for c, cn in ((0, 1), (-1, -2)):
            lat[r, c], lon[r, c] = lat[r, cn], lon[rn, c]
    return g


def _hr_lat_lon(lr_lat_lon, s_enhance=None, shape=None):
    """High-res (lat, lon) grid of the full output domain, the reference's
    ``OutputHandler.get_lat_lon`` (writers/base.py:434-508): longitudes wrapped to [-180, 180)
    (shifted to [0, 360) when a row crosses the date line), the grid extended by one cell
    (``_extend_grid``), both grids laid on cell centres of the same (0, 10) square and the
    coordinates interpolated linearly over its triangulation (``scipy.interpolate.griddata``, as
    the reference does -- host-side metadata, not a kernel), longitudes wrapped back.  float64
    (S1, S2, 2)."""
    from scipy.interpolate import griddata
    ll = np.array(lr_lat_lon, dtype=np.asarray(lr_lat_lon).dtype, copy=True)
    n1, n2 = ll.shape[:2]
    if shape is None:
        shape = (n1 * s_enhance, n2 * s_enhance)
    assert n1 > 1 and n2 > 1, "low res lat/lon must have at least 2 rows and 2 columns"
    ll[..., 1] = (ll[..., 1] + 180) % 360 - 180
    if any(ll[i, -1, 1] < ll[i, 0, 1] for i in range(n1)):
        ll[..., 1] = (ll[..., 1] + 360) % 360
    g = _extend_grid(ll)

    def centres(n):
        return np.arange(0, 10, 10 / n) + 5 / n
    y, x = centres(n1), centres(n2)
    y = np.concatenate([[y[0] - 10 / n1], y, [y[-1] + 10 / n1]])
    x = np.concatenate([[x[0] - 10 / n2], x, [x[-1] + 10 / n2]])
    xx, yy = np.meshgrid(x, y, copy=False)
    old = np.array([yy.flatten(), xx.flatten()], dtype=np.float32).T
    xx, yy = np.meshgrid(centres(shape[1]), centres(shape[0]), copy=False)
    new = np.array([yy.flatten(), xx.flatten()], dtype=np.float32).T
    lons = griddata(old, g[..., 1].flatten(), new)
    lats = griddata(old, g[..., 0].flatten(), new)
    lons = (lons + 180) % 360 - 180
    return np.dstack((lats.reshape(shape), lons.reshape(shape)))


def _hr_times(lr_times, n_hr):
    """High-res time axis of a chunk, the reference's ``OutputHandler.get_times``
    (writers/base.py:510-549): the smallest low-res step divided by the enhancement, continued
    one low-res step past the last time; 29 February dropped when the low-res index has none.
    Datetime input -> ``pd.DatetimeIndex``; a plain numeric index (the in-memory handler's
    default) -> evenly spaced floats."""
    lr = np.asarray(lr_times)
    t_enhance = int(n_hr / len(lr))
    if not np.issubdtype(lr.dtype, np.datetime64):
        step = float(np.min(np.diff(lr))) if len(lr) > 1 else 1.0
        return float(lr[0]) + np.arange(n_hr) * (step / t_enhance)
    import pandas as pd
    ti = pd.DatetimeIndex(lr)
    secs = min(set(np.diff(ti)) if len(ti) > 1 else [np.timedelta64(1, "D")]) \
        / np.timedelta64(1, "s")
    offset = pd.tseries.offsets.DateOffset(seconds=secs)
    freq = pd.tseries.offsets.DateOffset(seconds=int(offset.seconds / t_enhance))
    times = pd.date_range(ti[0], ti[-1] + offset, freq=freq)[:-1]
    if not any((ti.month == 2) & (ti.day == 29)):
        times = times[~((times.month == 2) & (times.day == 29))]
    assert len(times) == n_hr, (
        f"High res times length {len(times)} does not match expected shape {n_hr}")
    return times


@dataclass
class ForwardPassStrategy:
    """Chunking strategy + model / input handles for a forward pass."""

    file_paths: Union[str, list, None] = None
    model_kwargs: Optional[dict] = None
    fwp_chunk_shape: tuple = (None, None, None)
    spatial_pad: int = 0
    temporal_pad: int = 0
    min_width: tuple = (4, 4, 4)
    model_class: str = "Sup3rGan"
    out_pattern: Optional[str] = None
    input_handler_name: Optional[str] = None
    input_handler_kwargs: Optional[dict] = None
    exo_handler_kwargs: Optional[dict] = None
    bias_correct_method: Optional[str] = None
    bias_correct_kwargs: Optional[dict] = None
    allowed_const: Optional[Union[list, bool]] = None
    incremental: bool = True
    output_workers: int = 1
    invert_uv: Optional[bool] = True
    nn_fill: bool = True
    pass_workers: int = 1
    max_nodes: int = 1
    head_node: bool = False
    redistribute_chunks: bool = False
    use_cpu: bool = False
    # sup3r_b200 additions: live objects instead of files
    input_handler: Optional[ArrayInputHandler] = None
    model: Optional[object] = None
    exo_data: Optional[dict] = None
    pad_mode: str = "reflect"
    output_dtype: str = "float32"   # "float16": results are cast on the device and leave as fp16
    # apply the writer-side transforms (sup3r/writers/base.py:297-346: u/v -> ws/wd when
    # ``invert_uv``, physical limits with ``nn_fill`` or clipping) on the device before a chunk
    # leaves the GPU (pipeline/postprocess.py).  None: on for ``.nc`` chunk files -- what the
    # reference's writers always do (writers/nc.py:19-100 -> ``_transform_output``) -- and off for
    # in-memory results and ``.npy`` files, which hold the raw generator output like the
    # reference's ``run_chunk`` return value
    postprocess: Optional[bool] = None

    def __post_init__(self):
        self.bias_correct_kwargs = self.bias_correct_kwargs or {}
        if self.postprocess is None:
            self.postprocess = bool(self.out_pattern) and str(self.out_pattern).endswith(".nc")
        if self.bias_correct_kwargs:
            from ..bias import METHODS
            if self.bias_correct_method not in METHODS:
                raise KeyError(f'bias_correct_method "{self.bias_correct_method}" is not one of '
                               f"{sorted(METHODS)}")
        self.input_handler_kwargs = dict(self.input_handler_kwargs or {})
        self.timer = Timer()
        model = self.get_model()
        self.s_enhancements = model.s_enhancements
        self.t_enhancements = model.t_enhancements
        self.s_enhance, self.t_enhance = model.s_enhance, model.t_enhance
        self.input_features = model.lr_features
        self.output_features = model.hr_out_features
        self.features, self.exo_features = self._init_features(model)
        self.time_slice, self.padded_time_slice = self.get_time_slices()
        self.input_handler = self.timer(self.init_input_handler, log=True)()
        self.fwp_chunk_shape = self._get_fwp_chunk_shape()
        self.fwp_slicer = ForwardPassSlicer(
            coarse_shape=self.input_handler.grid_shape,
            time_steps=len(self.input_handler.time_index), time_slice=self.time_slice,
            chunk_shape=self.fwp_chunk_shape, s_enhance=self.s_enhance,
            t_enhance=self.t_enhance, spatial_pad=self.spatial_pad,
            temporal_pad=self.temporal_pad, min_width=self.min_width)
        self.n_chunks = self.fwp_slicer.n_chunks
        hr_shape = self.hr_lat_lon.shape[:-1]
        self.gids = np.arange(np.prod(hr_shape)).reshape(hr_shape)
        if self.exo_data is not None and not isinstance(self.exo_data, ExoData):
            self.exo_data = ExoData(self.exo_data)
        self.preflight()
        if self.use_cpu:
            raise RuntimeError("use_cpu=True: sup3r_b200 has no CPU compute path; the generator "
                               "runs on the GPU (the reference defaults to CPU, "
                               "strategy.py:201-204)")

    def get_model(self):
        if self.model is None:
            self.model = get_model(self.model_class, self.model_kwargs)
        return self.model

    @property
    def meta(self):
        return {"fwp_chunk_shape": self.fwp_chunk_shape, "spatial_pad": self.spatial_pad,
                "temporal_pad": self.temporal_pad, "model_kwargs": self.model_kwargs,
                "model_class": self.model_class, "spatial_enhance": int(self.s_enhance),
                "temporal_enhance": int(self.t_enhance), "input_files": self.file_paths,
                "input_features": self.features, "output_features": self.output_features,
                "input_shape": self.input_handler.grid_shape}

    def get_time_slices(self):
        """(slice of the padded time index that is unpadded, padded source slice)
        (strategy.py:302-333)."""
        ts = _parse_time_slice(self.input_handler_kwargs.get("time_slice", slice(None)))
        step = ts.step if ts.step else 1
        pstart = 0 if not ts.start else ts.start - self.temporal_pad * step
        pend = None if not ts.stop else ts.stop + self.temporal_pad * step
        padded = slice(pstart, pend, ts.step)
        start = 0 if not padded.start else self.temporal_pad
        stop = None if not padded.stop or not self.temporal_pad else -self.temporal_pad
        return slice(start, stop), padded

    def init_input_handler(self):
        if self.input_handler is not None:
            h = self.input_handler
            ps = self.padded_time_slice
            if (ps.start or 0) != 0 or ps.stop is not None or ps.step not in (None, 1):
                h = ArrayInputHandler(h.data, h.features, lat_lon=h.lat_lon,
                                      time_index=h.time_index, time_slice=ps)
            return h
        if self.file_paths is None:
            raise ValueError("ForwardPassStrategy needs an input_handler or file_paths")
        kwargs = {k: v for k, v in self.input_handler_kwargs.items()
                  if k in ("lat_lon", "time_index")}
        return ArrayInputHandler.from_files(self.file_paths, self.features,
                                            time_slice=self.padded_time_slice, **kwargs)

    def _init_features(self, model):
        self.exo_handler_kwargs = self.exo_handler_kwargs or {}
        exo_features = list(self.exo_handler_kwargs) or list(self.exo_data or {})
        features = [f for f in model.lr_features if f not in exo_features]
        return features, exo_features

    @property
    def node_chunks(self):
        """Chunk indices split over nodes / ranks with ``np.array_split``
        (strategy.py:363-372)."""
        chunks = self.unmasked_chunks
        if self.redistribute_chunks:
            chunks = [c for c in chunks if not self.chunk_finished(c)]
        n = int(min(self.max_nodes or np.inf, max(len(chunks), 1)))
        return np.array_split(np.asarray(chunks, dtype=int), n)

    @property
    def unmasked_chunks(self):
        return [i for i in range(self.n_chunks) if not self.chunk_masked(i, log=False)]

    def _get_fwp_chunk_shape(self):
        grid = self.input_handler.grid_shape
        tsteps = len(self.input_handler.time_index[self.time_slice])
        return tuple(fs or full for fs, full in zip(self.fwp_chunk_shape, (*grid, tsteps)))

    def preflight(self):
        self.ti_slices, self.ti_pad_slices = self.fwp_slicer.get_time_slices()
        s1 = self.fwp_chunk_shape[0] + 2 * self.spatial_pad
        s2 = self.fwp_chunk_shape[1] + 2 * self.spatial_pad
        if s1 < 4 or s2 < 4:
            msg = ("The padding layers in the generator typically require at least 4 elements "
                   f"per spatial dimension. The padded chunk shape ({s1}, {s2}) is smaller than "
                   "this.")
            logger.warning(msg)
            warn(msg)
        fwp_t = self.fwp_chunk_shape[2] + 2 * self.temporal_pad
        tsteps = len(self.input_handler.time_index[self.time_slice])
        if fwp_t > tsteps:
            msg = (f"Using a padded chunk size ({fwp_t}) larger than the full temporal domain "
                   f"({tsteps}). Should just run without temporal chunking. ")
            logger.warning(msg)
            warn(msg)
        self.lr_slices, self.lr_pad_slices, self.hr_slices = self.fwp_slicer.get_spatial_slices()
        info = {"n_nodes": len(self.node_chunks),
                "n_spatial_chunks": self.fwp_slicer.n_spatial_chunks,
                "n_time_chunks": self.fwp_slicer.n_time_chunks,
                "n_total_chunks": self.fwp_slicer.n_chunks}
        logger.info("Chunk strategy description:\n%s", pprint.pformat(info, indent=2))

    def get_chunk_indices(self, chunk_index):
        return self.fwp_slicer.get_chunk_indices(chunk_index)

    @cached_property
    def hr_lat_lon(self):
        return _hr_lat_lon(self.input_handler.lat_lon, self.s_enhance)

    @cached_property
    def out_files(self):
        ids = [f"{str(i).zfill(6)}_{str(j).zfill(6)}"
               for i in range(self.fwp_slicer.n_time_chunks)
               for j in range(self.fwp_slicer.n_spatial_chunks)]
        if self.out_pattern is None:
            return [None] * len(ids)
        assert "{file_id}" in self.out_pattern, "out_pattern must include a {file_id} format key"
        d = os.path.dirname(self.out_pattern)
        if d:
            os.makedirs(d, exist_ok=True)
        return [self.out_pattern.format(file_id=i) for i in ids]

    def prep_chunk_data(self, chunk_index=0):
        s_idx, t_idx = self.get_chunk_indices(chunk_index)
        lr_pad = self.lr_pad_slices[s_idx]
        ti_pad = self.ti_pad_slices[t_idx]
        exo = None
        if self.exo_data is not None:
            exo = self.timer(self.exo_data.get_chunk, log=True, call_id=chunk_index)(
                [lr_pad[0], lr_pad[1], ti_pad])
        data = self.input_handler.get(self.features, lr_pad[0], lr_pad[1], ti_pad)
        if self.bias_correct_kwargs:
            # strategy.py:502-517: bias-correct the low-res chunk before it goes to the model
            from ..bias import bias_correct_features
            logger.info("Bias correcting data for chunk_index=%s, with shape=%s", chunk_index,
                        data.shape)
            data = bias_correct_features(
                np.array(data, dtype=np.float32, copy=True), self.features,
                self.input_handler.lat_lon[lr_pad[0], lr_pad[1]], self.bias_correct_method,
                self.bias_correct_kwargs, lr_padded_slice=(lr_pad[0], lr_pad[1], ti_pad),
                time_index=self.input_handler.time_index[ti_pad])
        return data, exo

    def chunk_padded_shape(self, chunk_index):
        """(s1, s2, t, features) of the padded input chunk WITHOUT loading it (slice extents +
        the extra edge padding of forward_pass.py:122-186)."""
        s_idx, t_idx = self.get_chunk_indices(chunk_index)
        lr_pad, ti_pad = self.lr_pad_slices[s_idx], self.ti_pad_slices[t_idx]
        grid = self.input_handler.grid_shape
        n_t = len(self.input_handler.time_index)
        ext = [len(range(*sl.indices(n))) for sl, n in zip((lr_pad[0], lr_pad[1], ti_pad),
                                                           (grid[0], grid[1], n_t))]
        pw = self.fwp_slicer.extra_padding[chunk_index]
        return tuple(e + int(p[0]) + int(p[1]) for e, p in zip(ext, pw)) + (len(self.features),)

    def init_chunk(self, chunk_index=0):
        s_idx, t_idx = self.fwp_slicer.get_chunk_indices(chunk_index)
        assert chunk_index <= self.fwp_slicer.n_chunks, (
            f"Requested forward pass on chunk_index={chunk_index} > "
            f"n_chunks={self.fwp_slicer.n_chunks}")
        hr_slice = self.hr_slices[s_idx]
        ti_slice = self.ti_slices[t_idx]
        lr_times = self.input_handler.time_index[ti_slice]
        data, exo = self.timer(self.prep_chunk_data, log=True, call_id=chunk_index)(
            chunk_index=chunk_index)
        hr_times = _hr_times(lr_times, self.t_enhance * len(lr_times))
        return ForwardPassChunk(
            input_data=data, exo_data=exo, lr_pad_slice=self.lr_pad_slices[s_idx],
            hr_crop_slice=self.fwp_slicer.hr_crop_slices[t_idx][s_idx],
            hr_lat_lon=self.hr_lat_lon[hr_slice[:2]], hr_times=hr_times,
            gids=self.gids[hr_slice[:2]], out_file=self.out_files[chunk_index],
            pad_width=self.fwp_slicer.extra_padding[chunk_index], index=chunk_index)

    @cached_property
    def fwp_mask(self):
        """Spatial chunks whose ``mask`` feature is all ones are skipped (strategy.py:630-661);
        the in-memory handler carries the mask as an optional attribute."""
        mask = np.zeros(len(self.lr_pad_slices))
        vals = getattr(self.input_handler, "mask", None)
        if vals is not None:
            for i, sl in enumerate(self.lr_pad_slices):
                mask[i] = bool(np.prod(np.asarray(vals)[sl[0], sl[1]].flatten()))
        return mask

    def node_finished(self, node_idx):
        return all(self.chunk_finished(i) for i in self.node_chunks[node_idx])

    def chunk_finished(self, chunk_idx, log=True):
        out_file = self.out_files[chunk_idx]
        done = out_file is not None and os.path.exists(out_file) and self.incremental
        if done and log:
            logger.info("%s already exists and incremental = True. Skipping forward pass for "
                        "chunk index %s.", out_file, chunk_idx)
        return done

    def chunk_masked(self, chunk_idx, log=True):
        s_idx, _ = self.fwp_slicer.get_chunk_indices(chunk_idx)
        masked = bool(self.fwp_mask[s_idx])
        if masked and log:
            logger.info("Chunk %s has spatial chunk index %s, which is fully masked. Skipping.",
                        chunk_idx, s_idx)
        return masked
