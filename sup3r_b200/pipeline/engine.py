"""Pipelined host <-> device execution of the generator over a stream of equal-shape low-res
batches: pinned host buffers, H2D / compute / D2H on three CUDA streams, two slots (each with
its own captured CUDA graph and static buffers) so that the D2H of batch i overlaps the
generator pass of batch i + 1.  This is the transfer pattern ``ForwardPass`` uses on the GPU in
place of the reference's per-chunk numpy round trips (forward_pass.py:188-272).
"""
from __future__ import annotations

import numpy as np
import torch

from ..network import to_device_tensor


class GeneratePipeline:
    """``push(lr_batch)`` enqueues one batch, ``pop()`` returns the oldest result (numpy view of
    a pinned buffer that stays valid until two more batches have been pushed)."""

    def __init__(self, model, lr_shape, precision=None, norm_in=True, un_norm_out=True,
                 slots=2, fresh_host=False, check=False, out_dtype="float32"):
        """``fresh_host``: every result gets its own pinned buffer (from torch's caching
        host allocator) that the caller may keep; ``check``: per-chunk device-side
        (min, max, n_nan) of every output channel (``ForwardPass._output_check``,
        forward_pass.py:384-425) travels with the result: ``pop()`` -> (array, checks);
        ``out_dtype``: "float32" (the reference's dtype) or "float16" -- the result is cast on
        the device and leaves as fp16 (half the D2H bytes and half the host memory traffic;
        for consumers that store 16-bit data anyway)."""
        if out_dtype not in ("float32", "float16"):
            raise ValueError(f"out_dtype must be float32 or float16, got {out_dtype!r}")
        self.out_dtype = out_dtype
        self._tdt = torch.float16 if out_dtype == "float16" else torch.float32
        self.fresh_host = fresh_host
        self.check = check
        if not torch.cuda.is_available():
            raise RuntimeError("GeneratePipeline needs a CUDA device (no CPU fallback)")
        self.model = model
        self.dev = model.torch_device()
        self.lr_shape = tuple(int(s) for s in lr_shape)
        self.precision = precision or model.precision
        gen = model.generator
        if not gen.built:
            gen.build(self.lr_shape)
        self.hr_shape = gen.output_shape(self.lr_shape)
        self.norm = None
        if norm_in and model.means is not None:
            means, stdevs = model._norm_arrays(model.lr_features, "low-res input")
            stdevs = np.where(stdevs == 0, 1, stdevs)
            self.norm = (torch.from_numpy((1.0 / stdevs).astype(np.float32)).to(self.dev),
                         torch.from_numpy((-means / stdevs).astype(np.float32)).to(self.dev))
        self.post = (None, None)
        if un_norm_out and model.means is not None:
            means, stdevs = model._norm_arrays(model.hr_out_features, "high-res output")
            self.post = (torch.from_numpy(stdevs).to(self.dev),
                         torch.from_numpy(means).to(self.dev))
        from ..plan import Plan
        self.slots = []
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_run = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        for _ in range(slots):
            plan = Plan(gen, self.precision)   # own graph + static buffers per slot
            self.slots.append(dict(
                plan=plan,
                x_host=torch.empty(self.lr_shape, dtype=torch.float32).pin_memory(),
                x_dev=torch.empty(self.lr_shape, dtype=torch.float32, device=self.dev),
                y_host=torch.empty(self.hr_shape, dtype=self._tdt).pin_memory(),
                y16=(torch.empty(self.hr_shape, dtype=torch.float16, device=self.dev)
                     if out_dtype == "float16" else None),
                h2d=torch.cuda.Event(), run=torch.cuda.Event(), d2h=torch.cuda.Event(),
                chk_host=torch.empty((self.hr_shape[0], self.hr_shape[-1], 3),
                                     dtype=torch.float32).pin_memory(),
                busy=False))
        self._next = 0
        self._queue = []
        self.h2d_bytes = int(np.prod(self.lr_shape)) * 4
        self.d2h_bytes = int(np.prod(self.hr_shape)) * (2 if out_dtype == "float16" else 4)
        # warm-up: capture the graphs before any timing
        for sl in self.slots:
            with torch.cuda.stream(self.s_run):
                x = sl["x_dev"].zero_()
                if self.norm is not None:
                    from .. import ops
                    x = ops.channel_affine(x, *self.norm)
                sl["plan"].run_graphed(x, None, *self.post)
        torch.cuda.synchronize(self.dev)

    def push(self, lr_batch, crops=None):
        """``crops``: optional per-chunk high-res slices the device-side check is restricted to
        (the reference checks the cropped chunk)."""
        sl = self.slots[self._next]
        if sl["busy"]:
            raise RuntimeError("pipeline slot still holds an un-popped result: call pop() first")
        sl["x_host"].numpy()[...] = lr_batch
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(sl["run"])          # previous pass on this slot has read x_dev
            sl["x_dev"].copy_(sl["x_host"], non_blocking=True)
            sl["h2d"].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(sl["h2d"])
            self.s_run.wait_event(sl["d2h"])         # previous result of this slot has left
            x = sl["x_dev"]
            if self.norm is not None:
                from .. import ops
                x = ops.channel_affine(x, *self.norm)
            out = sl["plan"].run_graphed(x, None, *self.post)
            chk = None
            if self.check:
                from .. import ops
                parts = []
                for k in range(out.shape[0]):
                    o = out[k] if crops is None or crops[k] is None else out[k][crops[k]]
                    parts.append(ops.channel_check(o.contiguous()))
                chk = torch.stack(parts)
            if sl["y16"] is not None:
                from .. import ops
                out = ops.cast_f16(out, out=sl["y16"])
            sl["run"].record(self.s_run)
        if self.fresh_host:
            sl["y_host"] = torch.empty(self.hr_shape, dtype=self._tdt, pin_memory=True)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(sl["run"])
            sl["y_host"].copy_(out, non_blocking=True)
            if chk is not None:
                chk.record_stream(self.s_out)
                sl["chk_host"].copy_(chk, non_blocking=True)
            sl["d2h"].record(self.s_out)
        sl["busy"] = True
        self._queue.append(self._next)
        self._next = (self._next + 1) % len(self.slots)

    def pop(self):
        i = self._queue.pop(0)
        sl = self.slots[i]
        sl["d2h"].synchronize()
        sl["busy"] = False
        if self.check:
            return sl["y_host"].numpy(), sl["chk_host"].numpy().copy()
        return sl["y_host"].numpy()

    def run(self, batches):
        """Generator over results for an iterable of low-res batches (keeps 1 batch in
        flight behind the one being computed)."""
        n_slots = len(self.slots)
        for b in batches:
            if len(self._queue) == n_slots:
                yield self.pop()
            self.push(b)
        while self._queue:
            yield self.pop()
