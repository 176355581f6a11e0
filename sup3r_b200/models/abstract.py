"""Single-model logic shared by the GAN classes (mirrors sup3r/models/abstract.py:30-1251):
network loading, normalisation, loss / optimiser factories, gradient step, exo-layer
dispatch, ``generate`` / ``_tf_generate``, history bookkeeping and checkpoint triggers.
All tensor arithmetic runs on the CUDA kernels behind ``sup3r_b200.ops``."""
from __future__ import annotations

import copy
import json
import logging
import os
import pprint
import time
from inspect import signature
from warnings import warn

import numpy as np
import pandas as pd
import torch

from .. import loss_metrics, ops
from ..exo import ExoData
from ..network import (CustomNetwork, DeviceArray, SUP3R_EXO_LAYERS, SUP3R_LAYERS,
                       SUP3R_OBS_LAYERS, exo_out_shape, layer_features,
                       run_exo_layer as _run_exo_layer, default_device, to_device_tensor)
from ..optimizers import OPTIMIZERS, Adam
from ..plan import Plan, default_precision
from ..utilities import VERSION_RECORD, Timer, camel_to_underscore, safe_cast

logger = logging.getLogger(__name__)


def safe_json_load(fp):
    if not isinstance(fp, str) or not fp.endswith(".json"):
        raise ValueError(f"Filepath must be a .json file: {fp}")
    if not os.path.isfile(fp):
        raise FileNotFoundError(f"Could not find json file: {fp}")
    with open(fp) as f:
        return json.load(f)


def numpy_if_tensor(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    return v


class TensorboardMixIn:
    """Batch counters and the (optional) tensorboard scalar writer (models/utilities.py:77-147).
    Scalars are written with ``torch.utils.tensorboard`` when it is importable."""

    def __init__(self):
        self._tb_writer = None
        self._tb_log_dir = None
        self._write_tb_profile = False
        self._total_batches = None
        self._history = None
        self.timer = Timer()

    @property
    def total_batches(self):
        if self._total_batches is None:
            if self._history is not None and "total_batches" in self._history:
                self._total_batches = self._history["total_batches"].values[-1]
            else:
                self._total_batches = 0
        return self._total_batches

    @total_batches.setter
    def total_batches(self, value):
        self._total_batches = value

    def dict_to_tensorboard(self, entry):
        if self._tb_writer is None:
            return
        for name, value in entry.items():
            if isinstance(value, str):
                self._tb_writer.add_text(name, value, self.total_batches)
            elif isinstance(value, dict):
                continue
            else:
                self._tb_writer.add_scalar(name, float(value), self.total_batches)

    def profile_to_tensorboard(self, name):
        """Device profiles come from ncu (profiles/), not from the training loop."""

    def _init_tensorboard_writer(self, out_dir):
        tb_log_pardir = os.path.abspath(os.path.join(out_dir, os.pardir))
        self._tb_log_dir = os.path.join(tb_log_pardir, "logs")
        os.makedirs(self._tb_log_dir, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter
            self._tb_writer = SummaryWriter(self._tb_log_dir)
        except Exception as e:  # pragma: no cover
            logger.warning("tensorboard writer unavailable: %s", e)


class AbstractSingleModel(TensorboardMixIn):
    """Operations on one generator (+ optional discriminator) network."""

    def __init__(self):
        super().__init__()
        self.gpu_list = list(range(torch.cuda.device_count())) if torch.cuda.is_available() else []
        self.default_device = "/cpu:0" if len(self.gpu_list) == 0 else "/gpu:0"
        self.name = None
        self.loss_name = None
        self.loss_fun = None
        self.precision = default_precision()
        self._version_record = VERSION_RECORD
        self._meta = None
        self._optimizer = None
        self._gen = None
        self._means = None
        self._stdevs = None
        self._train_record = pd.DataFrame()
        self._val_record = pd.DataFrame()
        self._plans = {}
        from ..train_graph import GraphedSteps
        self._graphed_steps = GraphedSteps(self)

    # ---- device mapping ------------------------------------------------------------------
    def torch_device(self, device_name=None):
        """Map the reference's '/gpu:i' / '/cpu:0' strings (base.py:92-106, abstract.py:837)
        to a torch device.  There is no CPU compute path: '/cpu:0' only hosts shapes/weights."""
        name = device_name or self.default_device or "/gpu:0"
        if "gpu" in name.lower() and torch.cuda.is_available():
            idx = int(name.split(":")[-1]) if ":" in name else 0
            return torch.device("cuda", idx)
        return default_device()

    # ---- networks ------------------------------------------------------------------------
    def load_network(self, model, name):
        """CustomNetwork from a hidden-layers list, a ``.json`` config or a saved ``.pkl``
        (abstract.py:57-111)."""
        if isinstance(model, str) and model.endswith(".json"):
            model = safe_json_load(model)
            self._meta[f"config_{name}"] = model
            if "hidden_layers" in model:
                model = model["hidden_layers"]
            elif ("meta" in model and f"config_{name}" in model["meta"]
                  and "hidden_layers" in model["meta"][f"config_{name}"]):
                model = model["meta"][f"config_{name}"]["hidden_layers"]
            else:
                msg = ('Could not load model from json config, need "hidden_layers" key or '
                       f'"meta/config_{name}/hidden_layers"  at top level but only found: '
                       f"{model.keys()}")
                logger.error(msg)
                raise KeyError(msg)
        elif isinstance(model, str) and model.endswith(".pkl"):
            model = CustomNetwork.load(model, device=self.torch_device())
        if isinstance(model, list):
            model = CustomNetwork(hidden_layers=model, name=name, device=self.torch_device())
        if not isinstance(model, CustomNetwork):
            msg = ("Something went wrong. Tried to load a custom network but ended up with a "
                   f'model of type "{type(model)}"')
            logger.error(msg)
            raise TypeError(msg)
        return model

    def plan_for(self, net, precision=None):
        key = (id(net), precision or self.precision)
        if key not in self._plans:
            self._plans[key] = Plan(net, precision or self.precision)
        return self._plans[key]

    # ---- normalisation -------------------------------------------------------------------
    @property
    def means(self):
        return self._means

    @property
    def stdevs(self):
        return self._stdevs

    def set_norm_stats(self, new_means, new_stdevs):
        """Set normalisation statistics from a batch handler (abstract.py:133-195): the new
        values always REPLACE the model's (continued training / transfer learning normalises
        with the new handler's statistics), stored as float32 per feature."""
        if new_means is not None and new_stdevs is not None:
            logger.info("Setting new normalization statistics...")
            logger.info("Model's previous data mean values: %s", self._means)
            logger.info("Model's previous data stdev values: %s", self._stdevs)
            if not isinstance(new_means, dict) or not isinstance(new_stdevs, dict):
                msg = ("Means and stdevs need to be dictionaries with keys as feature names but "
                       f"received means of type {type(new_means)} and stdevs of type "
                       f"{type(new_stdevs)}")
                logger.error(msg)
                raise TypeError(msg)
            self._means = {k: np.float32(v) for k, v in new_means.items()}
            self._stdevs = {k: np.float32(v) for k, v in new_stdevs.items()}
            need = list(self.lr_features) + list(self.hr_exo_features) + list(self.hr_out_features)
            missing = [f for f in need if f not in self._means]
            if any(missing):
                logger.warning('Need means for features "%s" but did not find in new means %s',
                               missing, self._means)

    def _norm_arrays(self, features, what):
        missing = [fn for fn in features if fn not in self._means]
        if any(missing):
            msg = (f"Could not find {what} features {missing} in means/stdevs: "
                   f"{self._means}/{self._stdevs}")
            logger.error(msg)
            raise KeyError(msg)
        means = np.array([self._means[fn] for fn in features], dtype=np.float32)
        stdevs = np.array([self._stdevs[fn] for fn in features], dtype=np.float32)
        return means, stdevs

    def norm_input(self, low_res):
        """(x - mean) / stdev per low-res feature (abstract.py:197-238); zero stdevs -> 1."""
        if self._means is None:
            return low_res
        means, stdevs = self._norm_arrays(self.lr_features, "low-res input")
        if any(stdevs == 0):
            stdevs = np.where(stdevs == 0, 1, stdevs)
            msg = "Some standard deviations are zero."
            logger.warning(msg)
            warn(msg)
        if isinstance(low_res, torch.Tensor):
            dev = low_res.device
            return ops.channel_affine(low_res, torch.from_numpy(1.0 / stdevs).to(dev),
                                      torch.from_numpy(-means / stdevs).to(dev))
        return ((np.asarray(low_res) - means) / stdevs).astype(np.float32)

    def un_norm_output(self, output):
        """x * stdev + mean per output feature (abstract.py:240-275)."""
        if self._means is None:
            return output
        means, stdevs = self._norm_arrays(self.hr_out_features, "high-res output")
        if isinstance(output, torch.Tensor):
            dev = output.device
            return ops.channel_affine(output, torch.from_numpy(stdevs).to(dev),
                                      torch.from_numpy(means).to(dev))
        return (np.asarray(output) * stdevs + means).astype(np.float32)

    # ---- accessors -----------------------------------------------------------------------
    @property
    def optimizer(self):
        return self._optimizer

    @property
    def history(self):
        return self._history

    @property
    def generator(self):
        return self._gen

    @property
    def generator_weights(self):
        return self.generator.weights

    # ---- optimiser / loss factories ----------------------------------------------------------
    @staticmethod
    def init_optimizer(optimizer, learning_rate):
        """None -> Adam(learning_rate); dict(name=..., **kwargs) -> class.from_config; an
        optimiser instance is passed through (abstract.py:321-350)."""
        if isinstance(optimizer, dict):
            name = optimizer["name"]
            if name not in OPTIMIZERS:
                raise ValueError(f"{name} not found in sup3r_b200 optimizers.")
            cls = OPTIMIZERS[name]
            params = signature(cls.__init__).parameters
            return cls.from_config({k: v for k, v in optimizer.items() if k in params})
        if optimizer is None:
            return Adam(learning_rate=learning_rate)
        return optimizer

    @staticmethod
    def load_saved_params(out_dir, verbose=True):
        """``model_params.json`` -> constructor kwargs (abstract.py:352-402)."""
        with open(os.path.join(out_dir, "model_params.json")) as f:
            params = json.load(f)
        fp_history = os.path.join(out_dir, "history.csv")
        params["history"] = fp_history if os.path.exists(fp_history) else None
        if "version_record" in params:
            rec = params.pop("version_record")
            if verbose:
                logger.info("Loading model from disk that was created with the following "
                            "package versions: \n%s", pprint.pformat(rec, indent=2))
        means, stdevs = params.get("means", None), params.get("stdevs", None)
        if means is not None and stdevs is not None:
            params["means"] = {k: np.float32(v) for k, v in means.items()}
            params["stdevs"] = {k: np.float32(v) for k, v in stdevs.items()}
        return params

    def _init_records(self):
        """Seed the running loss records from a loaded history (abstract.py:404-413)."""
        if self._history is not None:
            tcols = [c for c in self._history.columns if "train_" in c]
            vcols = [c for c in self._history.columns if "val_" in c]
            self._train_record = self._history[tcols].iloc[-1:].reset_index(drop=True)
            self._val_record = self._history[vcols].iloc[-1:].reset_index(drop=True)

    def get_hr_exo_input(self, hi_res):
        """{exo feature: (..., 1) tensor} gathered from the true hi-res (abstract.py:415-436)."""
        if len(self.hr_exo_features) == 0:
            return {}
        out = {}
        c = hi_res.shape[-1]
        for f in self.hr_exo_features:
            i = self.hr_features.index(f)
            crop = [(0, 0)] * (hi_res.dim() - 1) + [(i, c - i - 1)]
            out[f] = ops.crop_fwd(hi_res, crop)
        return out

    def _combine_loss_input(self, hi_res_true, hi_res_gen):
        """Append the exo channels of the truth to the generated tensor before the
        discriminator (abstract.py:438-459)."""
        from ..autograd import ConcatFn
        if hi_res_true.shape[-1] > hi_res_gen.shape[-1]:
            exo = self.get_hr_exo_input(hi_res_true)
            for f in self.hr_exo_features:
                hi_res_gen = ConcatFn.apply(hi_res_gen, exo[f])
        return hi_res_gen

    @classmethod
    def get_loss_fun(cls, loss):
        """Loss name | {name: kwargs, ..., 'term_weights': [...]} -> callable returning
        ``(loss, {snake_case_name: value})`` (abstract.py:461-502)."""
        loss = {loss: {}} if isinstance(loss, str) else loss
        names = [ln for ln in loss if ln != "term_weights"]
        funcs = [cls._get_loss_fun({ln: loss[ln]}) for ln in names]
        weights = copy.deepcopy(loss).pop("term_weights", [1.0] * len(names))

        def loss_fun(x1, x2):
            details = {}
            total = None
            for w, ln, fn in zip(weights, names, funcs):
                val = fn(x1, x2)
                details[camel_to_underscore(ln)] = val
                term = val if w == 1.0 else val * w
                total = term if total is None else total + term
            return total, details

        return loss_fun

    @staticmethod
    def _get_loss_fun(loss):
        kwargs = {}
        if isinstance(loss, dict):
            loss, kwargs = next(iter(loss.items()))
        out = getattr(loss_metrics, loss, None)
        if out is None:
            msg = (f'Could not find requested loss function "{loss}" in '
                   "sup3r_b200.loss_metrics.")
            logger.error(msg)
            raise KeyError(msg)
        return out(**kwargs)

    @staticmethod
    def get_optimizer_config(optimizer):
        """JSON-safe optimiser config (abstract.py:543-564)."""
        conf = optimizer.get_config()
        for k, v in conf.items():
            if isinstance(v, np.floating):
                conf[k] = float(v)
            elif isinstance(v, np.integer):
                conf[k] = int(v)
        return conf

    @classmethod
    def get_optimizer_state(cls, optimizer):
        """learning rate + mean(|var|) of every optimiser variable (abstract.py:566-587)."""
        state = {"learning_rate": cls.get_optimizer_config(optimizer)["learning_rate"]}
        for var in optimizer.variables:
            state[var.name] = float(np.abs(np.asarray(var.numpy(), dtype=np.float64)).mean())
        return state

    @staticmethod
    def update_loss_details(record, new_data, max_batches, prefix=None):
        """Append one row of loss details to a running record, keep the last ``max_batches``
        rows (abstract.py:589-622)."""
        new_index = 0 if len(record) == 0 else record.index[-1] + 1
        for k, v in new_data.items():
            key = k if prefix is None or prefix in k else prefix + k
            record.loc[new_index, key] = float(np.asarray(numpy_if_tensor(v)))
        return record.iloc[-max_batches:]

    @staticmethod
    def log_loss_details(loss_details, level="INFO"):
        for k, v in sorted(loss_details.items()):
            fmt = "\t{}: {}" if isinstance(v, str) else "\t{}: {:.2e}"
            (logger.info if level.lower() == "info" else logger.debug)(fmt.format(k, v))

    @staticmethod
    def early_stop(history, column, threshold=0.005, n_epoch=5):
        """True when the last ``n_epoch`` absolute differences of ``column`` are all below
        ``threshold`` (abstract.py:643-685)."""
        if history is not None and len(history) > n_epoch + 1:
            diffs = np.abs(np.diff(history[column]))
            if all(diffs[-n_epoch:] < threshold):
                logger.info('Found early stop condition on "%s": %s', column, diffs[-n_epoch:])
                return True
        return False

    def save(self, out_dir):  # pragma: no cover - abstract
        raise NotImplementedError

    def finish_epoch(self, epoch, epochs, t0, loss_details, checkpoint_int, out_dir,
                     early_stop_on, early_stop_threshold, early_stop_n_epoch, extras=None):
        """History row, checkpoint trigger, early stop (abstract.py:698-783)."""
        self.log_loss_details(loss_details)
        self._history.at[epoch, "elapsed_time"] = time.time() - t0
        for k, v in loss_details.items():
            self._history.at[epoch, k] = float(v)
        last_epoch = epoch == epochs[-1]
        chp = checkpoint_int is not None and (epoch % checkpoint_int) == 0
        if last_epoch or chp:
            assert "{epoch}" in out_dir, (
                "Model output dir for checkpoint models should have {epoch} but did not: "
                f"{out_dir}")
            self._save_rank0(out_dir.format(epoch=epoch))
        stop = False
        if early_stop_on is not None and early_stop_on in self._history:
            stop = self.early_stop(self._history, early_stop_on, threshold=early_stop_threshold,
                                   n_epoch=early_stop_n_epoch)
            if stop:
                self._save_rank0(out_dir.format(epoch=epoch))
        if extras is not None:
            for k, v in extras.items():
                self._history.at[epoch, k] = safe_cast(v)
        return stop

    def _save_rank0(self, out_dir):
        """Checkpoint from rank 0 only (the weights are replicated across ranks)."""
        from .. import parallel
        if parallel.rank() == 0:
            self.save(out_dir)

    # ---- gradient step ---------------------------------------------------------------------
    def run_gradient_descent(self, low_res, hi_res_true, training_weights, optimizer=None,
                             multi_gpu=False, **calc_loss_kwargs):
        """One optimiser step (abstract.py:843-914).  ``multi_gpu`` with an initialised process
        group (one rank per GPU, every rank fed the SAME batch): the reference's
        ``_get_parallel_grad`` / ``_sum_parallel_grad`` (abstract.py:785-841) -- the batch is
        split along axis 0 into world_size equal shards, this rank takes shard ``rank``, the
        shard gradients are SUMMED with one all-reduce and every rank applies the same step;
        the returned loss details are the last shard's (see ``parallel``)."""
        from .. import parallel
        if optimizer is None:
            optimizer = self.optimizer
        t0 = time.time()
        multi_gpu = bool(multi_gpu) and parallel.world_size() > 1
        if multi_gpu:
            low_res = parallel.shard_batch(low_res)
            hi_res_true = parallel.shard_batch(hi_res_true)
            if calc_loss_kwargs.get("mask") is not None:
                calc_loss_kwargs["mask"] = parallel.shard_batch(calc_loss_kwargs["mask"])
        # the whole gradient computation replayed as one CUDA graph once the shapes have been
        # seen a few times (train_graph.py); None -> the eager path below
        loss_details = self._graphed_steps.run(low_res, hi_res_true, training_weights, optimizer,
                                               multi_gpu, calc_loss_kwargs)
        if loss_details is None:
            grad, loss_details = self.get_single_grad(low_res, hi_res_true, training_weights,
                                                      device_name=self.default_device,
                                                      **calc_loss_kwargs)
            # SUM of the shard gradients (multi-GPU) + one optimiser step: one fused kernel with
            # this library's Adam (over NVLink peer memory when multi_gpu)
            parallel.sum_grads_and_step(grad, training_weights, optimizer, multi_gpu)
        if multi_gpu:
            loss_details = parallel.broadcast_loss_details(loss_details)
        logger.debug("Finished single gradient descent step in %.4f seconds", time.time() - t0)
        return loss_details

    def _get_parallel_grad(self, low_res, hi_res_true, training_weights, **calc_loss_kwargs):
        """``(total_grad, loss_details)`` of a multi-GPU step (abstract.py:807-841) in a
        one-process-per-GPU world: the batch (and a ``mask`` keyword) is split along axis 0 into
        world_size equal shards, this rank differentiates shard ``rank`` (the reference's
        ``/gpu:<rank>`` thread), the shard gradients are SUMMED over the ranks and the loss
        details of the last shard reach every rank (``_sum_parallel_grad``)."""
        from .. import parallel
        kw = dict(calc_loss_kwargs)
        if kw.get("mask") is not None:
            kw["mask"] = parallel.shard_batch(kw["mask"])
        grad, loss_details = self.get_single_grad(
            parallel.shard_batch(low_res), parallel.shard_batch(hi_res_true), training_weights,
            device_name=self.default_device, **kw)
        return parallel.allreduce_sum_grads(grad), parallel.broadcast_loss_details(loss_details)

    def _sum_parallel_grad(self, futures, start_time):
        """SUM of the ``(grad, loss_details)`` results of ``futures`` (objects with
        ``.result()``), loss details of the last one (abstract.py:785-805).  The product's
        multi-GPU step sums over ranks instead (``_get_parallel_grad``); kept for callers that
        hold per-shard results in one process."""
        total_grad = loss_details = None
        for future in futures:
            grad, loss_details = future.result()
            total_grad = list(grad) if total_grad is None else \
                [t + g for t, g in zip(total_grad, grad)]
        logger.info("Finished %d gradient descent steps in %.4f seconds", len(futures),
                    time.time() - start_time)
        return total_grad, loss_details

    # ---- exo layers --------------------------------------------------------------------------
    def _run_exo_layer(self, layer, input_array, hi_res_exo):
        """Run an exo / observation layer on a device tensor with the ``{feature: tensor}``
        dictionary of ``_tf_generate`` (abstract.py:1107-1129)."""
        return _run_exo_layer(layer, input_array, hi_res_exo)

    def _reshape_norm_exo(self, hi_res, hi_res_exo, exo_name, norm_in=True):
        """Normalise a hi-res exo array and tile it to the rank of ``hi_res``
        (abstract.py:916-979)."""
        if hi_res_exo is None:
            return hi_res_exo
        hi_res_exo = np.asarray(hi_res_exo, dtype=np.float32)
        if norm_in and self._means is not None:
            nm = exo_name.replace("_obs", "") if exo_name not in self._means else exo_name
            hi_res_exo = (hi_res_exo - self._means[nm]) / self._stdevs[nm]
        if hi_res_exo.ndim == 3:
            hi_res_exo = np.repeat(hi_res_exo[None], hi_res.shape[0], axis=0)
        if hi_res_exo.ndim == 4 and len(hi_res.shape) == 5:
            hi_res_exo = np.repeat(np.expand_dims(hi_res_exo, 3), hi_res.shape[3], axis=3)
        if hi_res_exo.ndim != len(hi_res.shape):
            msg = ("hi_res and hi_res_exo arrays are not of the same rank: "
                   f"{tuple(hi_res.shape)} and {hi_res_exo.shape}")
            logger.error(msg)
            raise RuntimeError(msg)
        return hi_res_exo

    def _exo_for_layer(self, layer, shape_like, exogenous_data, norm_in):
        """{feature: normalised, rank-matched array} for the ``features`` (+ ``exo_features``
        extras) of one exo / obs layer (abstract.py:981-1035).  Observation features that are
        not in ``exogenous_data`` are skipped with a warning (the layer then runs without them);
        any other missing feature is an ``AssertionError``."""
        out = {}
        feats = layer_features(layer)
        is_obs = isinstance(layer, SUP3R_OBS_LAYERS)
        for feat in feats + list(getattr(layer, "exo_features", [])):
            missing = exogenous_data is None or feat not in exogenous_data
            if is_obs and feat in feats and missing:
                logger.warning("%s does not match any features in exogenous_data (%s). Will run "
                               "without this observation feature.", feat,
                               list(exogenous_data or {}))
                continue
            assert not missing, f'exogenous_data is missing required feature "{feat}"'
            exo = exogenous_data.get_combine_type_data(feat, "layer")
            out[feat] = self._reshape_norm_exo(shape_like, exo, feat, norm_in=norm_in)
        return out

    def run_exo_layer(self, layer, input_array, exogenous_data, norm_in=True):
        """Run one Sup3rAdder / Sup3rConcat / observation layer from public ``generate`` inputs
        (abstract.py:981-1035)."""
        exo = self._exo_for_layer(layer, input_array, exogenous_data, norm_in)
        x = to_device_tensor(input_array, self.torch_device())
        return _run_exo_layer(layer, x, exo).as_subclass(DeviceArray)

    def _hr_shapes_at_exo_layers(self, in_shape, exo_channels=None):
        """Tensor shape entering each exo / obs layer for an input of ``in_shape``."""
        shp = tuple(in_shape)
        out = {}
        for lyr in self.generator.layers:
            if isinstance(lyr, SUP3R_LAYERS):
                out[lyr.name] = shp
                shp = exo_out_shape(lyr, shp, exo_channels)
            else:
                shp = lyr.out_shape(shp)
        return out

    # ---- generate ------------------------------------------------------------------------------
    def generate(self, low_res, norm_in=True, un_norm_out=True, exogenous_data=None,
                 precision=None, use_graph=True, to_numpy=True):
        """Public generate (abstract.py:1037-1105): exo-combine -> normalise -> generator ->
        un-normalise -> exo-output concat.  numpy in, numpy out (float32).  The generator runs
        as a fused plan on the GPU; ``precision`` overrides the model's precision mode."""
        if exogenous_data is not None and not isinstance(exogenous_data, ExoData):
            exogenous_data = ExoData(exogenous_data)
        if isinstance(low_res, torch.Tensor):
            # device-resident input (MultiStepGan keeps intermediates on the GPU); exo channels
            # that must be appended to the input are combined on the host like the reference
            if exogenous_data is not None and len(self.lr_features) > low_res.shape[-1]:
                low_res = self._combine_fwp_input(low_res.detach().cpu().numpy(), exogenous_data)
        else:
            low_res = self._combine_fwp_input(np.asarray(low_res), exogenous_data)
        gen = self.generator
        rank = getattr(gen.layers[0], "rank", None)
        if rank is not None and low_res.ndim != rank:
            raise RuntimeError(f"generator expects {rank}-D input but received shape "
                               f"{tuple(low_res.shape)}")
        dev = self.torch_device()
        if dev.type != "cuda":
            raise RuntimeError("sup3r_b200 needs a CUDA device to run the generator "
                               "(no CPU fallback)")
        exo_layers = [lyr for lyr in gen.layers if isinstance(lyr, SUP3R_LAYERS)]
        exo_dev = {}
        if exo_layers:
            # observation features that are absent make their layer the identity
            absent = {f: 0 for lyr in exo_layers if isinstance(lyr, SUP3R_OBS_LAYERS)
                      for f in layer_features(lyr)
                      if exogenous_data is None or f not in exogenous_data}
            shapes = self._hr_shapes_at_exo_layers(low_res.shape, absent)
            for lyr in exo_layers:
                try:
                    arrs = self._exo_for_layer(lyr, np.empty(shapes[lyr.name][:-1] + (0,)),
                                               exogenous_data, norm_in)
                except AssertionError as e:
                    raise RuntimeError(f'Could not run layer "{lyr}": {e}') from e
                for f, arr in arrs.items():
                    exo_dev[f] = to_device_tensor(arr, dev)
        x = to_device_tensor(low_res, dev)
        if norm_in and self._means is not None:
            x = self.norm_input(x)
        ps = pf = None
        if un_norm_out and self._means is not None:
            means, stdevs = self._norm_arrays(self.hr_out_features, "high-res output")
            ps, pf = torch.from_numpy(stdevs).to(dev), torch.from_numpy(means).to(dev)
        plan = self.plan_for(gen, precision)
        if not gen.built:
            gen.build(tuple(x.shape), {k: v.shape[-1] for k, v in exo_dev.items()})
        run = plan.run_graphed if use_graph else plan.run
        hi_res = run(x, exo_dev, ps, pf)
        if not to_numpy:
            return hi_res
        # D2H through the caching pinned-host allocator: DMA at PCIe speed, and the returned
        # ndarray owns its (pinned) buffer, which goes back to the cache when it is released
        host = torch.empty(tuple(hi_res.shape), dtype=torch.float32, pin_memory=True)
        host.copy_(hi_res, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return self._combine_fwp_output(host.numpy(), exogenous_data)

    def _tf_generate(self, low_res, hi_res_exo=None):
        """Differentiable generator forward on device tensors (abstract.py:1131-1173).
        ``hi_res_exo``: {feature: tensor} for the exo layers."""
        x = to_device_tensor(low_res, self.torch_device())
        # (generator convolutions: tcgen05 forward + input gradient unless precision == "fp32")
        plan = self.plan_for(self.generator, self.precision)
        if not getattr(self, "_gen_on_tape", True):
            # a step that trains no generator weight (the discriminator step): nothing of the
            # generator is recorded -- no saved activations, no input gradient of the
            # discriminator's first layer on the synthetic branch
            with torch.no_grad():
                return plan.forward_train(x, hi_res_exo or {})
        return plan.forward_train(x, hi_res_exo or {})

    def _get_hr_exo_and_loss(self, low_res, hi_res_true, **calc_loss_kwargs):
        """Generator forward + loss (abstract.py:1175-1188)."""
        hi_res_true = to_device_tensor(hi_res_true, self.torch_device())
        hi_res_exo = self.get_hr_exo_input(hi_res_true)
        hi_res_gen = self._tf_generate(low_res, hi_res_exo)
        loss, loss_details = self.calc_loss(hi_res_true, hi_res_gen, **calc_loss_kwargs)
        return loss, loss_details, hi_res_gen, hi_res_exo

    def get_single_grad(self, low_res, hi_res_true, training_weights, device_name=None,
                        **calc_loss_kwargs):
        """Gradients of the loss w.r.t. ``training_weights`` (abstract.py:1190-1238): the tape
        is torch.autograd over Functions whose forward / backward are all our kernels.

        The backward pass runs on the loss times a power of two (``grad_loss_scale``) and the
        gradients are divided by it afterwards: the losses are means over ~1e6 hi-res values, so
        the raw activation gradients are ~1e-6 -- below the normal range of the fp16 operands the
        tensor-core input-gradient kernels use.  Exact for the fp32 kernels (power of two)."""
        gen_ids = {id(w) for w in self.generator_weights}
        self._gen_on_tape = any(id(w) in gen_ids for w in training_weights)
        try:
            with torch.enable_grad():
                loss, loss_details, hi_res_gen, _ = self._get_hr_exo_and_loss(
                    low_res, hi_res_true, **calc_loss_kwargs)
        finally:
            self._gen_on_tape = True
        with torch.enable_grad():
            tensors = [w.value for w in training_weights]
            scale = self.grad_loss_scale(hi_res_gen, calc_loss_kwargs.get("train_gen", True))
            if scale != 1.0:
                from .base import ScaleFnScalar
                grad = torch.autograd.grad(ScaleFnScalar.apply(loss, scale), tensors,
                                           allow_unused=True)
                have = [g for g in grad if g is not None]
                if have:
                    torch._foreach_mul_(have, 1.0 / scale)     # one multi-tensor launch
            else:
                grad = torch.autograd.grad(loss, tensors, allow_unused=True)
        grad = [g if g is not None else torch.zeros_like(t) for g, t in zip(grad, tensors)]
        loss_details = {k: (v.detach() if isinstance(v, torch.Tensor) else v)
                        for k, v in loss_details.items()}
        return grad, loss_details

    def grad_loss_scale(self, hi_res_gen, train_gen=True):
        """Power-of-two loss scale of the backward pass (1 in ``fp32`` mode) that brings the
        top-level gradient to order one: the generator loss is a mean over the generated hi-res
        values (gradients ~ 1 / their number); the discriminator loss a mean over the 2 x batch
        logits."""
        if self.precision == "fp32":
            return 1.0
        n = hi_res_gen.numel() if train_gen else 2 * hi_res_gen.shape[0]
        return float(2.0 ** int(np.floor(np.log2(max(n, 1)))))

    def calc_loss(self, hi_res_true, hi_res_gen, weight_gen_advers=0.001, train_gen=True,
                  train_disc=False, compute_disc=False):  # pragma: no cover - abstract
        raise NotImplementedError
