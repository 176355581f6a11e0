"""``ForwardPass``: run the generator over the chunks of a strategy (mirrors
sup3r/pipeline/forward_pass.py:32-673).

Differences that are deliberate ("where time goes"):
* the model stays resident on the GPU instead of being re-loaded from disk per chunk
  (forward_pass.py:638);
* chunks of equal padded shape can be stacked on the batch axis (``run_batched``);
* multi-node distribution is one rank per GPU (``torch.distributed``): rank r runs
  ``strategy.node_chunks[r]`` with no data-path collective.
Exception types are the reference's: NaN input -> RuntimeError, failed generator ->
RuntimeError, constant / NaN output -> MemoryError.
"""
from __future__ import annotations

import copy
import logging
import pprint
from datetime import datetime as dt

import numpy as np

from ..utilities import Timer
from .strategy import ForwardPassChunk, ForwardPassStrategy
from .utilities import get_model
from .writers import OutputHandlerNC

logger = logging.getLogger(__name__)


class ForwardPass:
    """Forward-pass driver for one node (= one GPU rank)."""

    # chunk file writers by file type (forward_pass.py:49-52; the h5 writer is out of scope)
    OUTPUT_HANDLER_CLASS = {"nc": OutputHandlerNC}

    def __init__(self, strategy, node_index=0):
        self.timer = Timer()
        self.strategy = strategy
        self.model = strategy.get_model()
        self.node_index = node_index
        out = strategy.out_pattern
        assert out is None or out.endswith((".npy", ".nc")), (
            f"Received bad output type {out}: sup3r_b200 writes .npy or .nc chunk files (the h5 "
            "writer is out of scope)")

    def get_input_chunk(self, chunk_index=0, mode="reflect"):
        """Chunk with its extra edge padding applied (forward_pass.py:67-74)."""
        chunk = self.strategy.init_chunk(chunk_index)
        chunk.input_data, chunk.exo_data = self.pad_source_data(
            chunk.input_data, chunk.pad_width, chunk.exo_data, mode=mode)
        return chunk

    @property
    def meta(self):
        return {"node_index": self.node_index,
                "creation_date": dt.now().strftime("%d/%m/%Y %H:%M:%S"),
                "model_meta": self.model.meta, "gan_params": self.model.model_params,
                "strategy_meta": self.strategy.meta,
                "model_meta_features": list(self.model.hr_out_features),
                "full_hr_shape": [int(v) for v in self.strategy.hr_lat_lon.shape[:2]]}

    def _get_step_enhance(self, step):
        """Enhancement of an exo step's data relative to the low-res input
        (forward_pass.py:87-120)."""
        combine_type = step["combine_type"]
        model_step = step["model"]
        assert combine_type in ("input", "output", "layer"), (
            f"Received weird combine_type {combine_type} for step: {step}")
        if combine_type.lower() == "input":
            n = model_step
        else:
            n = model_step + 1
        s = int(np.prod(self.model.s_enhancements[:n])) if n > 0 else 1
        t = int(np.prod(self.model.t_enhancements[:n])) if n > 0 else 1
        return s, t

    def pad_source_data(self, input_data, pad_width, exo_data, mode="reflect"):
        """Pad the chunk (and its exo data, scaled by the step enhancement) at domain edges
        (forward_pass.py:122-186)."""
        out = np.pad(input_data, (*pad_width, (0, 0)), mode=mode)
        if exo_data is not None:
            for feature in exo_data:
                for i, step in enumerate(exo_data[feature]["steps"]):
                    s_en, t_en = self._get_step_enhance(step)
                    widths = (*((en * pw[0], en * pw[1])
                                for en, pw in zip([s_en, s_en, t_en], pad_width)), (0, 0))
                    new = step["data"]
                    if new.ndim == 3:
                        new = np.repeat(np.expand_dims(new, axis=2),
                                        step["t_enhance"] * input_data.shape[2], axis=2)
                    exo_data[feature]["steps"][i]["data"] = np.pad(new, widths, mode=mode)
        return out, exo_data

    @classmethod
    def run_generator(cls, data_chunk, hr_crop_slices, model, s_enhance=None, t_enhance=None,
                      exo_data=None):
        """(s1, s2, t, f) chunk -> cropped (S1, S2, T, F) high-res output
        (forward_pass.py:188-272)."""
        data_chunk, exo_data, i_lr_t, i_lr_s = cls._reshape_data_chunk(model, data_chunk,
                                                                       exo_data)
        try:
            hi_res = Timer()(model.generate, log=True)(data_chunk, exogenous_data=exo_data)
        except Exception as e:
            msg = f"Forward pass failed on chunk with shape {data_chunk.shape}."
            logger.exception(msg)
            raise RuntimeError(msg) from e
        if hi_res.ndim == 4:
            hi_res = np.expand_dims(np.transpose(hi_res, (1, 2, 0, 3)), axis=0)
        cls._check_enhancement(hi_res, data_chunk, s_enhance, t_enhance, i_lr_s, i_lr_t)
        return hi_res[0][hr_crop_slices]

    @staticmethod
    def _check_enhancement(hi_res, data_chunk, s_enhance, t_enhance, i_lr_s, i_lr_t):
        if s_enhance is not None and hi_res.shape[1] != s_enhance * data_chunk.shape[i_lr_s]:
            msg = (f"The stated spatial enhancement of {s_enhance}x did not match the low res / "
                   f"high res shapes of {data_chunk.shape} -> {hi_res.shape}")
            logger.error(msg)
            raise RuntimeError(msg)
        if t_enhance is not None and hi_res.shape[3] != t_enhance * data_chunk.shape[i_lr_t]:
            msg = (f"The stated temporal enhancement of {t_enhance}x did not match the low res / "
                   f"high res shapes of {data_chunk.shape} -> {hi_res.shape}")
            logger.error(msg)
            raise RuntimeError(msg)

    @staticmethod
    def _reshape_data_chunk(model, data_chunk, exo_data):
        """5-D models get a leading obs axis; 4-D models get time as the obs axis
        (forward_pass.py:274-337)."""
        if exo_data is not None:
            models = getattr(model, "models", [model])
            for feature in exo_data:
                for i, entry in enumerate(exo_data[feature]["steps"]):
                    assert entry["model"] < len(models), (
                        f'model index ({entry["model"]}) for exo step {i} exceeds the number of '
                        "model steps")
                    if models[entry["model"]].is_4d:
                        out = np.transpose(entry["data"], axes=(2, 0, 1, 3))
                    else:
                        out = np.expand_dims(entry["data"], axis=0)
                    exo_data[feature]["steps"][i]["data"] = np.asarray(out)
        if model.is_4d:
            i_lr_t, i_lr_s = 0, 1
            data_chunk = np.transpose(data_chunk, axes=(2, 0, 1, 3))
        else:
            i_lr_t, i_lr_s = 3, 1
            data_chunk = np.expand_dims(data_chunk, axis=0)
        return np.ascontiguousarray(data_chunk), exo_data, i_lr_t, i_lr_s

    @classmethod
    def _output_check(cls, out_data, allowed_const):
        """True when the output has NaNs or a constant channel that is not explicitly allowed
        (forward_pass.py:384-425)."""
        if allowed_const is True:
            return False
        if allowed_const is False or allowed_const is None:
            allowed_const = []
        elif not isinstance(allowed_const, (list, tuple)):
            allowed_const = [allowed_const]
        if np.isnan(out_data).any():
            logger.error("Forward pass output contains NaN values!")
            return True
        for i in range(out_data.shape[-1]):
            value0 = out_data[0, 0, 0, i]
            if (value0 == out_data[..., i]).all() and value0 not in allowed_const:
                logger.error("All values are the same for feature channel %d!", i)
                return True
        return False

    # ---- drivers -----------------------------------------------------------------------------
    @classmethod
    def run(cls, strategy, node_index=0):
        """Run every unfinished chunk of ``strategy.node_chunks[node_index]``
        (forward_pass.py:427-449).  Returns {chunk_index: output} for in-memory runs."""
        out = {}
        if not strategy.node_finished(node_index):
            if strategy.pass_workers == 1 and not strategy.postprocess:
                out = cls._run_serial(strategy, node_index)
            else:
                # (device-side post-processing lives in the streamed driver, whatever the
                #  number of pass workers)
                out = cls._run_batched(strategy, node_index, batch_size=strategy.pass_workers)
            logger.debug("Timing report:\n%s", pprint.pformat(strategy.timer.log, indent=2))
        return out

    @classmethod
    def _run_serial(cls, strategy, node_index):
        start = dt.now()
        fwp = cls(strategy, node_index=node_index)
        outputs = {}
        chunks = strategy.node_chunks[node_index]
        for i, chunk_index in enumerate(chunks):
            chunk_index = int(chunk_index)
            now = dt.now()
            if strategy.chunk_finished(chunk_index):
                continue
            chunk = fwp.get_input_chunk(chunk_index=chunk_index, mode=strategy.pad_mode)
            failed, data = cls.run_chunk(
                chunk=chunk, model_kwargs=strategy.model_kwargs, model_class=strategy.model_class,
                allowed_const=strategy.allowed_const, output_workers=strategy.output_workers,
                invert_uv=strategy.invert_uv, nn_fill=strategy.nn_fill, meta=fwp.meta,
                model=fwp.model)
            logger.info("Finished forward pass on chunk_index=%s in %s. %d of %d complete.",
                        chunk_index, dt.now() - now, i + 1, len(chunks))
            if failed:
                raise MemoryError(f"Forward pass for chunk_index {chunk_index} failed with "
                                  "constant output or NaNs.")
            if chunk.out_file is None:
                outputs[chunk_index] = data
        logger.info("Finished forward passes on %d chunks in %s", len(chunks), dt.now() - start)
        return outputs

    @classmethod
    def _run_parallel(cls, strategy, node_index):
        """The reference's name for the multi-worker path (forward_pass.py:503-580: a pool of
        ``pass_workers`` processes); here the workers are the batch entries of one GPU."""
        return cls._run_batched(strategy, node_index, batch_size=strategy.pass_workers)

    @classmethod
    def _run_batched(cls, strategy, node_index, batch_size=8):
        """Stack chunks of equal padded shape on the obs axis (5-D models) -- the on-GPU
        replacement for the reference's ``SpawnProcessPool`` (forward_pass.py:503-580)."""
        fwp = cls(strategy, node_index=node_index)
        model = fwp.model
        if (hasattr(model, "models") or not model.is_5d or strategy.exo_data is not None
                or strategy.postprocess):
            # 4-D models, exogenous data, multi-step chains: one chunk per generator call,
            # software-pipelined on the device (see _run_streamed)
            return cls._run_streamed(strategy, node_index, fwp)
        # group the unfinished chunks by padded shape from the slicer alone; the chunks themselves
        # are sliced / padded lazily, one batch ahead of the GPU (host memory ~ 2 batches)
        groups = {}
        for chunk_index in strategy.node_chunks[node_index]:
            chunk_index = int(chunk_index)
            if strategy.chunk_finished(chunk_index):
                continue
            groups.setdefault(strategy.chunk_padded_shape(chunk_index), []).append(chunk_index)
        outputs = {}
        from .engine import GeneratePipeline
        cache = model.__dict__.setdefault("_fwp_pipelines", {})
        out_dtype = getattr(strategy, "output_dtype", "float32")
        for shape, ids in groups.items():
            lr_shape = (batch_size, *shape)
            in_memory = any(strategy.out_files[i] is None for i in ids)
            key = (lr_shape, model.precision, out_dtype, in_memory)
            pipe = cache.get(key)
            if pipe is None:
                # pinned buffers, H2D / generator / D2H on three streams, device-side output
                # check; results kept in memory land in their own pinned buffers (no second host
                # copy), chunk files are written straight from the reused slots
                pipe = cache[key] = GeneratePipeline(model, lr_shape, check=True,
                                                     out_dtype=out_dtype, fresh_host=in_memory)
            parts = [ids[i:i + batch_size] for i in range(0, len(ids), batch_size)]

            def load(part_ids):
                part = []
                for ci in part_ids:
                    chunk = fwp.get_input_chunk(chunk_index=ci, mode=strategy.pad_mode)
                    cls._check_nan_input(chunk, model)
                    assert chunk.input_data.shape == shape, (chunk.input_data.shape, shape)
                    part.append(chunk)
                batch = np.stack([c.input_data for c in part]
                                 + [part[-1].input_data] * (batch_size - len(part)), axis=0)
                crops = [c.hr_crop_slice for c in part] + [None] * (batch_size - len(part))
                for c in part:
                    c.input_data = None      # the batch holds the data now
                return part, batch, crops

            def finish(part, hi_res, checks):
                cls._check_enhancement(hi_res, np.empty(lr_shape), model.s_enhance,
                                       model.t_enhance, 1, 3)
                for k, c in enumerate(part):
                    if cls._device_check_failed(checks[k], strategy.allowed_const):
                        raise MemoryError(f"Forward pass for chunk_index {c.index} failed with "
                                          "constant output or NaNs.")
                    data = hi_res[k][c.hr_crop_slice]
                    if c.out_file is not None:
                        cls._write_output(data, c, fwp.meta)
                    else:
                        outputs[c.index] = data

            pending = []          # chunk lists of the batches in flight, oldest first
            for part_ids in parts:
                part, batch, crops = load(part_ids)
                if len(pipe._queue) == len(pipe.slots):
                    hi_res, checks = pipe.pop()
                    finish(pending.pop(0), hi_res, checks)
                pipe.push(batch, crops)
                pending.append(part)
            while pending:
                hi_res, checks = pipe.pop()
                finish(pending.pop(0), hi_res, checks)
        return outputs

    @classmethod
    def _run_streamed(cls, strategy, node_index, fwp=None):
        """Chunk-at-a-time driver for everything the stacked-batch path does not cover (4-D
        models, exogenous data, ``MultiStepGan`` chains; forward_pass.py:122-272, 582-673): the
        generator output stays on the GPU, the crop and ``_output_check`` run there
        (``s3_channel_check``), only the cropped interior leaves -- into its own pinned buffer on
        a copy stream -- while the next chunk is already being generated."""
        import torch
        from .. import ops
        start = dt.now()
        fwp = fwp or cls(strategy, node_index=node_index)
        model = fwp.model
        first = getattr(model, "models", [model])[0]
        dev = first.torch_device()
        s_out = torch.cuda.Stream(dev)
        outputs = {}
        pending = []

        def finish():
            chunk, host, chk_host, done = pending.pop(0)
            done.synchronize()
            if cls._device_check_failed(chk_host.numpy(), strategy.allowed_const):
                raise MemoryError(f"Forward pass for chunk_index {chunk.index} failed with "
                                  "constant output or NaNs.")
            data = host.numpy()
            if chunk.out_file is not None:
                logger.info("Saving forward pass output to %s.", chunk.out_file)
                cls._write_output(data, chunk, fwp.meta)
            else:
                outputs[chunk.index] = data

        chunks = strategy.node_chunks[node_index]
        for i, chunk_index in enumerate(chunks):
            chunk_index = int(chunk_index)
            if strategy.chunk_finished(chunk_index):
                continue
            chunk = fwp.get_input_chunk(chunk_index=chunk_index, mode=strategy.pad_mode)
            cls._check_nan_input(chunk, model)
            data_chunk, exo, i_lr_t, i_lr_s = cls._reshape_data_chunk(
                model, chunk.input_data, copy.deepcopy(chunk.exo_data))
            # exo channels appended to a step's OUTPUT are combined on the host by generate()
            host_combine = exo is not None and any(
                st["combine_type"] == "output" for f in exo for st in exo[f]["steps"])
            try:
                hi_res = fwp.timer(model.generate, log=True)(
                    data_chunk, exogenous_data=exo, **({} if host_combine else {"to_numpy": False}))
            except Exception as e:
                msg = f"Forward pass failed on chunk with shape {data_chunk.shape}."
                logger.exception(msg)
                raise RuntimeError(msg) from e
            if not isinstance(hi_res, torch.Tensor):
                # (models that post-process on the host, e.g. SolarCC's temporal padding)
                hi_res = torch.as_tensor(np.ascontiguousarray(hi_res), device=dev)
            if hi_res.ndim == 4:
                hi_res = hi_res.permute(1, 2, 0, 3)[None]
            cls._check_enhancement(hi_res, data_chunk, model.s_enhance, model.t_enhance, i_lr_s,
                                   i_lr_t)
            out = hi_res[0][chunk.hr_crop_slice].contiguous()
            if strategy.postprocess:
                from .postprocess import transform_output
                out, chunk.features = transform_output(
                    out, list(model.hr_out_features), chunk.hr_lat_lon,
                    invert_uv=bool(strategy.invert_uv), nn_fill=bool(strategy.nn_fill))
            chk = ops.channel_check(out)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(dev))
            host = torch.empty(tuple(out.shape), dtype=torch.float32, pin_memory=True)
            chk_host = torch.empty(tuple(chk.shape), dtype=torch.float32, pin_memory=True)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ready)
                host.copy_(out, non_blocking=True)
                chk_host.copy_(chk, non_blocking=True)
                out.record_stream(s_out)
                chk.record_stream(s_out)
                done = torch.cuda.Event()
                done.record(s_out)
            chunk.input_data = None
            pending.append((chunk, host, chk_host, done))
            if len(pending) > 1:
                finish()
            logger.info("Queued forward pass on chunk_index=%s. %d of %d.", chunk_index, i + 1,
                        len(chunks))
        while pending:
            finish()
        logger.info("Finished forward passes on %d chunks in %s", len(chunks), dt.now() - start)
        return outputs

    @staticmethod
    def _device_check_failed(chk, allowed_const):
        """``_output_check`` (forward_pass.py:384-425) from the device-side per-channel
        (min, max, n_nan) table of one chunk."""
        if allowed_const is True:
            return False
        if allowed_const is False or allowed_const is None:
            allowed_const = []
        elif not isinstance(allowed_const, (list, tuple)):
            allowed_const = [allowed_const]
        if (chk[:, 2] > 0).any() or np.isnan(chk[:, :2]).any():
            logger.error("Forward pass output contains NaN values!")
            return True
        for i in range(chk.shape[0]):
            if chk[i, 0] == chk[i, 1] and chk[i, 0] not in allowed_const:
                logger.error("All values are the same for feature channel %d!", i)
                return True
        return False

    @staticmethod
    def _check_nan_input(chunk, model):
        mask = np.isnan(chunk.input_data).any(axis=(0, 1, 2))
        if np.any(mask):
            feats = np.array(model.lr_features[: len(mask)])[mask] \
                if len(model.lr_features) >= len(mask) else np.where(mask)[0]
            msg = f"Input data for {feats} contains NaN values!"
            logger.error(msg)
            raise RuntimeError(msg)

    @staticmethod
    def _write_output(data, chunk, meta):
        if chunk.out_file.endswith(".nc"):
            # writers/nc.py:19-100 (the device-side transforms already ran when
            # strategy.postprocess is set: ``chunk.features`` then holds the renamed features)
            from .writers import OutputHandlerNC
            feats = getattr(chunk, "features", None) or meta["model_meta_features"]
            attrs = {k: v for k, v in meta.items() if k != "model_meta_features"}
            attrs["full_hr_shape"] = meta.get("full_hr_shape")
            OutputHandlerNC._write_output(np.asarray(data, np.float32), feats, chunk.hr_lat_lon,
                                          chunk.hr_times, chunk.out_file, meta_data=attrs,
                                          gids=chunk.gids, transform=False)
            return
        np.save(chunk.out_file, data)
        import json
        from ..utilities import safe_cast
        with open(chunk.out_file + ".meta.json", "w") as f:
            json.dump({"meta": meta, "hr_times": np.asarray(chunk.hr_times).tolist(),
                       "index": chunk.index, "features": getattr(chunk, "features", None)}, f,
                      default=safe_cast)

    @classmethod
    def run_chunk(cls, chunk: ForwardPassChunk, model_kwargs, model_class, allowed_const,
                  invert_uv=False, meta=None, nn_fill=True, output_workers=None, model=None):
        """NaN check -> generator -> output check -> optional write
        (forward_pass.py:582-673).  ``model``: resident model (else loaded from
        ``model_kwargs`` like the reference)."""
        logger.info("Running forward pass for chunk_index=%s.", chunk.index)
        if model is None:
            model = get_model(model_class, model_kwargs)
        cls._check_nan_input(chunk, model)
        output_data = cls.run_generator(
            data_chunk=chunk.input_data, hr_crop_slices=chunk.hr_crop_slice,
            s_enhance=model.s_enhance, t_enhance=model.t_enhance,
            exo_data=copy.deepcopy(chunk.exo_data), model=model)
        failed = cls._output_check(output_data, allowed_const=allowed_const)
        if chunk.out_file is not None and not failed:
            logger.info("Saving forward pass output to %s.", chunk.out_file)
            cls._write_output(output_data, chunk, meta)
        return failed, output_data
