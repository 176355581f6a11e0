"""Chunk writers and collectors: ``OutputHandlerNC`` and ``CollectorNC`` mirror
sup3r/writers/nc.py:16-100 and sup3r/postprocessing/collectors/nc.py:19-141 -- one NETCDF file
per forward-pass chunk with dimensions ``(time, south_north, west_east)``, 2-D ``latitude`` /
``longitude`` coordinates, ``gids`` (global hi-res indices used to stitch spatial chunks), one
float32 variable per (renamed) feature and the model / strategy meta data as global attributes;
the collector groups the files by their spatial chunk index (``{t:06d}_{s:06d}`` names,
collectors/base.py), concatenates each group along time and stitches the groups by ``gids``.

Files are NETCDF-3 64-bit-offset (``scipy.io.netcdf_file``; netCDF4 / h5py / xarray are not part of
this library's environment) -- xarray and netCDF4 read them.  The writer-side transforms
(u/v -> windspeed / winddirection, physical limits; writers/base.py:297-346) run on the GPU
(``pipeline/postprocess.py``).  H5 output (``OutputHandlerH5``, rex meta tables) is out of scope.
"""
from __future__ import annotations

import datetime
import glob
import json
import logging
import os
import re

import numpy as np

logger = logging.getLogger(__name__)

TIME, SOUTH_NORTH, WEST_EAST = "time", "south_north", "west_east"
LATITUDE, LONGITUDE = "latitude", "longitude"


def _attr(v):
    if isinstance(v, (str, bytes, int, float, np.integer, np.floating)):
        return v
    from ..utilities import safe_cast
    return json.dumps(v, default=safe_cast)


class OutputHandlerNC:
    """Forward-pass output handler for NETCDF chunk files."""

    @classmethod
    def _transform_output(cls, data, features, lat_lon, invert_uv=False, nn_fill=False):
        """writers/base.py:297-346 on the GPU; ``data``: numpy array or CUDA tensor
        (spatial_1, spatial_2, temporal, features)."""
        import torch
        from .postprocess import transform_output
        dev_data = data if isinstance(data, torch.Tensor) else \
            torch.as_tensor(np.ascontiguousarray(data, dtype=np.float32), device="cuda")
        dev_data, features = transform_output(dev_data.contiguous(), features, lat_lon,
                                              invert_uv=invert_uv, nn_fill=nn_fill)
        return dev_data.cpu().numpy(), features

    @classmethod
    def _write_output(cls, data, features, lat_lon, times, out_file, meta_data=None,
                      max_workers=None, invert_uv=False, nn_fill=False, gids=None,
                      transform=True):
        """Write one chunk (writers/nc.py:19-100).  ``times``: numeric time stamps of the hi-res
        steps; ``transform=False`` when the chunk was already post-processed on the device."""
        from scipy.io import netcdf_file
        if transform:
            data, features = cls._transform_output(data, list(features), lat_lon,
                                                   invert_uv=invert_uv, nn_fill=nn_fill)
        data = np.asarray(data)
        s1, s2, nt = data.shape[:3]
        tmp = out_file + ".tmp"
        with netcdf_file(tmp, "w", version=2) as f:
            f.createDimension(TIME, nt)
            f.createDimension(SOUTH_NORTH, s1)
            f.createDimension(WEST_EAST, s2)
            v = f.createVariable(TIME, "f8", (TIME,))
            v[:] = np.asarray(times, dtype=np.float64)
            for name, arr in ((LATITUDE, lat_lon[:, :, 0]), (LONGITUDE, lat_lon[:, :, 1])):
                v = f.createVariable(name, "f4", (SOUTH_NORTH, WEST_EAST))
                v[:] = np.asarray(arr, dtype=np.float32)
            if gids is not None:
                v = f.createVariable("gids", "i4", (SOUTH_NORTH, WEST_EAST))
                v[:] = np.asarray(gids, dtype=np.int32)
            for i, feat in enumerate(features):
                v = f.createVariable(feat, "f4", (TIME, SOUTH_NORTH, WEST_EAST))
                v[:] = np.transpose(data[..., i], axes=(2, 0, 1)).astype(np.float32)
            attrs = dict(meta_data or {})
            now = datetime.datetime.now(datetime.timezone.utc).isoformat()
            attrs["date_modified"] = now
            attrs.setdefault("date_created", now)
            for k, val in attrs.items():
                setattr(f, re.sub(r"[^0-9a-zA-Z_]", "_", str(k)), _attr(val))
        os.replace(tmp, out_file)    # a chunk file exists only when it is complete (incremental)
        return features


def read_nc(path):
    """-> dict(features={name: (time, s1, s2) array}, time, latitude, longitude, gids, attrs)"""
    from scipy.io import netcdf_file
    out = {"features": {}, "attrs": {}}
    with netcdf_file(path, "r", mmap=False) as f:
        for name, var in f.variables.items():
            arr = np.array(var[:])
            if name in (TIME, LATITUDE, LONGITUDE, "gids"):
                out[name] = arr
            else:
                out["features"][name] = arr
        for k, v in f._attributes.items():
            out["attrs"][k] = v.decode() if isinstance(v, bytes) else v
    return out


class CollectorNC:
    """Collect NETCDF chunk files into one file (collectors/nc.py:19-141)."""

    def __init__(self, file_paths):
        if isinstance(file_paths, str):
            file_paths = glob.glob(file_paths)
        self.flist = sorted(file_paths)
        assert self.flist, "no chunk files to collect"

    @staticmethod
    def get_chunk_indices(file):
        """(temporal, spatial) chunk index strings of a ``..._{t:06d}_{s:06d}.nc`` name
        (collectors/base.py:62-79)."""
        m = re.search(r"(\d{6})_(\d{6})(?=\.[a-zA-Z0-9]+$)", os.path.basename(file))
        assert m, f"chunk file name without a {{t:06d}}_{{s:06d}} id: {file}"
        return m.group(1), m.group(2)

    def group_spatial_chunks(self):
        """{spatial index: time-sorted files of the same footprint} (collectors/nc.py:131-141)"""
        chunks = {}
        for file in self.flist:
            _, s_idx = self.get_chunk_indices(file)
            chunks.setdefault(s_idx, []).append(file)
        return {k: sorted(v) for k, v in chunks.items()}

    @classmethod
    def collect(cls, file_paths, out_file, features="all", overwrite=True, full_shape=None):
        """Concatenate every spatial chunk's files along time, stitch the chunks by their ``gids``
        (row-major indices into the full hi-res grid of ``full_shape``; read from the chunk
        attributes ``full_hr_shape`` when not given) and write ``out_file``."""
        collector = cls(file_paths)
        logger.info("Collecting %d files to %s", len(collector.flist), out_file)
        d = os.path.dirname(out_file)
        if d:
            os.makedirs(d, exist_ok=True)
        if os.path.exists(out_file):
            if not overwrite:
                logger.info("%s exists and overwrite=False.", out_file)
                return out_file
            os.remove(out_file)
        groups = collector.group_spatial_chunks()
        full, lat, lon, times, attrs = None, None, None, None, {}
        for s_idx, files in groups.items():
            parts = [read_nc(fp) for fp in files]
            attrs = attrs or parts[0]["attrs"]
            if full_shape is None:
                full_shape = tuple(json.loads(parts[0]["attrs"]["full_hr_shape"]))
            t_all = np.concatenate([p[TIME] for p in parts])
            if times is None:
                times = t_all
            assert len(t_all) == len(times), "spatial chunks cover different time ranges"
            names = list(parts[0]["features"]) if features == "all" else list(features)
            if full is None:
                full = {n: np.full((len(times), *full_shape), np.nan, np.float32) for n in names}
                lat = np.full(full_shape, np.nan, np.float32)
                lon = np.full(full_shape, np.nan, np.float32)
            gids = parts[0]["gids"]
            rows, cols = gids // full_shape[1], gids % full_shape[1]
            lat[rows, cols], lon[rows, cols] = parts[0][LATITUDE], parts[0][LONGITUDE]
            for n in names:
                full[n][:, rows, cols] = np.concatenate([p["features"][n] for p in parts], axis=0)
        data = np.stack([np.transpose(full[n], (1, 2, 0)) for n in full], axis=-1)
        attrs = {k: v for k, v in attrs.items() if k not in ("date_modified",)}
        OutputHandlerNC._write_output(
            data, list(full), np.stack([lat, lon], axis=-1), times, out_file, meta_data=attrs,
            gids=np.arange(int(np.prod(full_shape))).reshape(full_shape), transform=False)
        logger.info("Finished file collection.")
        return out_file
