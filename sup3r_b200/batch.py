"""Batch production on the device: the hi-res training data stays resident in
HBM, random samples are gathered, coarsened (block mean + temporal method) and optionally
gaussian-smoothed there -- the on-GPU replacement for the reference's sampler + FIFO queue thread
(sup3r/preprocessing/samplers/base.py, batch_queues/abstract.py:135-296, base.py:32-87,
batch_queues/utilities.py:57-104).  ``DeviceBatchHandler`` offers the attributes
``Sup3rGan.train`` consumes (sup3r/models/base.py:728-733, 1138-1157)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .network import to_device_tensor


class Batch:
    """``low_res`` / ``high_res`` device tensors."""

    def __init__(self, low_res, high_res):
        self.low_res, self.high_res = low_res, high_res


def transform(samples, s_enhance, t_enhance, features, smoothing=None, smoothing_ignore=None,
              temporal_coarsening_method="subsample"):
    """hi-res samples (n, s1, s2, [t,] f) on the device -> low-res (base.py:32-87)."""
    low = ops.coarsen(samples, s_enhance, t_enhance, temporal_coarsening_method)
    if smoothing is not None:
        ignore = smoothing_ignore or []
        mask = 0
        for j, f in enumerate(features[: low.shape[-1]]):
            if f not in ignore:
                mask |= 1 << j
        if mask:
            low = ops.gauss_smooth2d(low, smoothing, mask)
    return low


class DeviceBatchHandler:
    """Random spatiotemporal samples of a device-resident hi-res dataset ``data``
    (s1, s2, t, features), normalised with its own per-feature means / stds.

    ``features``: names of the channels of ``data``; ``lr_features`` default to all of them,
    ``hr_out_features`` likewise; ``hr_exo_features`` are channels that only appear in the
    hi-res tensor (they follow the outputs, base.py:478-503)."""

    def __init__(self, data, features, sample_shape, batch_size=16, n_batches=64, s_enhance=1,
                 t_enhance=1, lr_features=None, hr_out_features=None, hr_exo_features=None,
                 smoothing=None, smoothing_ignore=None, temporal_coarsening_method="subsample",
                 val_frac=0.1, seed=42, device=None, means=None, stds=None):
        self.features = list(features)
        self.lr_features = list(lr_features or self.features)
        self.hr_out_features = list(hr_out_features or self.features)
        self.hr_exo_features = list(hr_exo_features or [])
        self.hr_features = self.hr_out_features + self.hr_exo_features
        self.s_enhance, self.t_enhance = int(s_enhance), int(t_enhance)
        self.sample_shape = tuple(int(v) for v in sample_shape)
        self.batch_size, self.n_batches = int(batch_size), int(n_batches)
        self.smoothing, self.smoothing_ignore = smoothing, list(smoothing_ignore or [])
        self.smoothed_features = [f for f in self.lr_features if f not in self.smoothing_ignore] \
            if smoothing is not None else []
        self.temporal_coarsening_method = temporal_coarsening_method
        data = np.asarray(data, dtype=np.float32)
        assert data.ndim == 4 and data.shape[-1] == len(self.features), data.shape
        s1, s2, t = self.sample_shape
        assert s1 % self.s_enhance == 0 and s2 % self.s_enhance == 0 and t % self.t_enhance == 0, (
            f"sample_shape {self.sample_shape} must be divisible by the enhancements")
        assert all(a >= b for a, b in zip(data.shape[:3], self.sample_shape)), (
            "sample_shape is larger than the data")
        flat = data.reshape(-1, data.shape[-1]).astype(np.float64)
        self.means = dict(means or {f: float(m) for f, m in zip(self.features, np.nanmean(flat, 0))})
        self.stds = dict(stds or {f: float(s) for f, s in zip(self.features, np.nanstd(flat, 0))})
        mean = np.array([self.means[f] for f in self.features], np.float32)
        std = np.array([self.stds[f] for f in self.features], np.float32)
        std = np.where(std == 0, 1, std)
        self.device = torch.device(device) if device is not None else torch.device("cuda", 0)
        dev = to_device_tensor(data, self.device)
        self.data = ops.channel_affine(dev, torch.from_numpy(1.0 / std).to(self.device),
                                       torch.from_numpy(-mean / std).to(self.device))
        # the last val_frac of the time axis is the validation range
        n_t = data.shape[2]
        self._t_split = n_t if not val_frac else max(t, int(round(n_t * (1 - val_frac))))
        if n_t - self._t_split < t:
            self._t_split = n_t
            self._val_t0 = max(0, n_t - t)
        else:
            self._val_t0 = self._t_split
        self.rng = np.random.default_rng(seed)
        self._lr_idx = [self.features.index(f) for f in self.lr_features]
        self._hr_idx = [self.features.index(f) for f in self.hr_features]
        self.lr_shape = (s1 // self.s_enhance, s2 // self.s_enhance, t // self.t_enhance,
                         len(self.lr_features))
        self.hr_shape = (s1, s2, t, len(self.hr_features))
        self.shapes = ((self.batch_size, *self.lr_shape), (self.batch_size, *self.hr_shape))
        self.stopped = False
        self._val = None

    def __len__(self):
        return self.n_batches

    def _origins(self, n, t_lo, t_hi):
        S1, S2 = self.data.shape[:2]
        s1, s2, t = self.sample_shape
        o = np.stack([self.rng.integers(0, S1 - s1 + 1, n), self.rng.integers(0, S2 - s2 + 1, n),
                      self.rng.integers(t_lo, t_hi - t + 1, n)], axis=1).astype(np.int32)
        return torch.from_numpy(o).to(self.device)

    def _select(self, x, idx):
        if idx == list(range(x.shape[-1])):
            return x
        return x[..., idx].contiguous()

    def make_batch(self, origins):
        """Gather + coarsen (+ smooth) one batch for the given (n, 3) int32 device origins."""
        samples = ops.gather_samples(self.data, origins, self.sample_shape)
        low_in = self._select(samples, self._lr_idx)
        low = transform(low_in, self.s_enhance, self.t_enhance, self.lr_features, self.smoothing,
                        self.smoothing_ignore, self.temporal_coarsening_method)
        return Batch(low, self._select(samples, self._hr_idx))

    def __iter__(self):
        for _ in range(self.n_batches):
            yield self.make_batch(self._origins(self.batch_size, 0, self._t_split))

    @property
    def val_data(self):
        if self._val is None:
            n_t = self.data.shape[2]
            self._val = [self.make_batch(self._origins(self.batch_size, self._val_t0, n_t))]
        return self._val

    def stop(self):
        self.stopped = True
