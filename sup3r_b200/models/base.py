"""``Sup3rGan``: generator + discriminator, relativistic average GAN loss, training loop,
checkpointing (mirrors sup3r/models/base.py:20-1191) on the sup3r_b200 CUDA kernels."""
from __future__ import annotations

import copy
import logging
import os
import pprint
import time
from warnings import warn

import numpy as np
import pandas as pd
import torch

from .. import ops
from ..autograd import DiscLossFn
from ..network import to_device_tensor
from ..optimizers import get_optimizer_class
from ..utilities import VERSION_RECORD
from .abstract import AbstractSingleModel
from .interface import AbstractInterface

logger = logging.getLogger(__name__)


class Sup3rGan(AbstractSingleModel, AbstractInterface):
    """Spatial (4-D) or spatiotemporal (5-D) super-resolution GAN."""

    _graph_safe = True    # gradient step may be captured as a CUDA graph (train_graph.py)

    def __init__(self, gen_layers, disc_layers, loss="MeanSquaredError", optimizer=None,
                 learning_rate=1e-4, optimizer_disc=None, learning_rate_disc=None, history=None,
                 meta=None, means=None, stdevs=None, default_device=None, name=None,
                 precision=None):
        super().__init__()
        self.default_device = default_device
        if self.default_device is None:
            self.default_device = "/gpu:0" if len(self.gpu_list) >= 1 else "/cpu:0"
        if precision is not None:
            self.precision = precision
        self.name = name if name is not None else self.__class__.__name__
        self._meta = meta if meta is not None else {}
        self.loss_name = loss
        self.loss_fun = self.get_loss_fun(loss)
        self._history = history
        if isinstance(self._history, str):
            self._history = pd.read_csv(self._history, index_col=0)
        self._init_records()
        optimizer_disc = optimizer_disc or copy.deepcopy(optimizer)
        learning_rate_disc = learning_rate_disc or learning_rate
        self._optimizer = self.init_optimizer(optimizer, learning_rate)
        self._optimizer_disc = self.init_optimizer(optimizer_disc, learning_rate_disc)
        self._gen = self.load_network(gen_layers, "generator")
        self._disc = self.load_network(disc_layers, "discriminator")
        self._means = means
        self._stdevs = stdevs

    # ---- persistence -------------------------------------------------------------------------
    def save(self, out_dir):
        """``model_gen.pkl``, ``model_disc.pkl``, ``history.csv``, ``model_params.json``
        (base.py:133-159)."""
        os.makedirs(out_dir, exist_ok=True)
        self.generator.save(os.path.join(out_dir, "model_gen.pkl"))
        self.discriminator.save(os.path.join(out_dir, "model_disc.pkl"))
        if isinstance(self.history, pd.DataFrame):
            self.history.to_csv(os.path.join(out_dir, "history.csv"))
        self.save_params(out_dir)
        logger.info("Saved GAN to disk in directory: %s", out_dir)

    @classmethod
    def _load(cls, model_dir, verbose=True):
        if verbose:
            logger.info("Loading GAN from disk in directory: %s", model_dir)
            logger.info("Active python environment versions: \n%s",
                        pprint.pformat(VERSION_RECORD, indent=4))
        fp_gen = os.path.join(model_dir, "model_gen.pkl")
        fp_disc = os.path.join(model_dir, "model_disc.pkl")
        params = cls.load_saved_params(model_dir, verbose=verbose)
        return fp_gen, fp_disc, params

    @classmethod
    def load(cls, model_dir, verbose=True, **kwargs):
        fp_gen, fp_disc, params = cls._load(model_dir, verbose=verbose)
        params.update(kwargs)
        return cls(fp_gen, fp_disc, **params)

    # ---- discriminator -----------------------------------------------------------------------
    @property
    def discriminator(self):
        return self._disc

    @property
    def discriminator_weights(self):
        return self.discriminator.weights

    def discriminate(self, hi_res, norm_in=False):
        """Discriminator logits for ``hi_res`` (numpy in / out) (base.py:237-281)."""
        if isinstance(hi_res, torch.Tensor):
            hi_res = hi_res.detach().cpu().numpy()
        hi_res = np.asarray(hi_res, dtype=np.float32)
        if norm_in and self._means is not None:
            mean_arr = np.array([self._means[fn] for fn in self.hr_out_features], np.float32)
            std_arr = np.array([self._stdevs[fn] for fn in self.hr_out_features], np.float32)
            hi_res = (hi_res - mean_arr) / std_arr
        with torch.no_grad():
            out = self._tf_discriminate(hi_res)
        return out.detach().cpu().numpy()

    def _tf_discriminate(self, hi_res):
        """Differentiable discriminator forward on device tensors (base.py:283-313)."""
        x = to_device_tensor(hi_res, self.torch_device())
        # (64-channel-block 'same' convolutions run on tcgen05 unless precision == "fp32")
        prec = "fp32" if self.precision == "fp32" else "fp16c"
        return self.plan_for(self.discriminator, prec).forward_train(x)

    # ---- optimisers --------------------------------------------------------------------------
    @property
    def optimizer_disc(self):
        return self._optimizer_disc

    def update_optimizer(self, option="generator", **kwargs):
        """Re-create optimisers with updated config values (base.py:326-348)."""
        if "gen" in option.lower() or "all" in option.lower():
            conf = self.get_optimizer_config(self.optimizer)
            conf.update(**kwargs)
            self._optimizer = get_optimizer_class(conf).from_config(conf)
        if "disc" in option.lower() or "all" in option.lower():
            conf = self.get_optimizer_config(self.optimizer_disc)
            conf.update(**kwargs)
            self._optimizer_disc = get_optimizer_class(conf).from_config(conf)

    @property
    def meta(self):
        if "class" not in self._meta:
            self._meta["class"] = self.__class__.__name__
        return self._meta

    @property
    def model_params(self):
        means, stdevs = self._means, self._stdevs
        if means is not None and stdevs is not None:
            means = {k: float(v) for k, v in means.items()}
            stdevs = {k: float(v) for k, v in stdevs.items()}
        return {"name": self.name, "loss": self.loss_name, "version_record": self.version_record,
                "optimizer": self.get_optimizer_config(self.optimizer),
                "optimizer_disc": self.get_optimizer_config(self.optimizer_disc),
                "means": means, "stdevs": stdevs, "meta": self.meta,
                "default_device": self.default_device}

    @property
    def weights(self):
        return self.generator_weights + self.discriminator_weights

    def init_weights(self, lr_shape, hr_shape, device=None):
        """Build the generator / discriminator variables for these batch shapes and check the
        number of generator outputs (base.py:394-437)."""
        if not self.generator_weights or not self.discriminator_weights:
            exo = {f: 1 for f in self.hr_exo_features + self.obs_features}
            out_shape = self.generator.build(tuple(lr_shape), exo)
            if self.hr_out_features:
                assert out_shape[-1] == len(self.hr_out_features), (
                    f"Number of model outputs {out_shape[-1]} does not match the number of "
                    f"computed hr_out_features {len(self.hr_out_features)}")
            self.discriminator.build(tuple(hr_shape))
            self._plans.clear()

    @staticmethod
    def get_weight_update_fraction(history, comparison_key, update_bounds=(0.5, 0.95),
                                   update_frac=0.0):
        """Multiplier for the adversarial weight from the disc training fraction
        (base.py:439-476)."""
        val = history[comparison_key]
        if isinstance(val, (list, tuple, np.ndarray)):
            val = val[-1]
        if val < update_bounds[0]:
            return 1 + update_frac
        if val > update_bounds[1]:
            return 1 / (1 + update_frac)
        return 1

    # ---- losses ------------------------------------------------------------------------------
    def calc_loss_gen_content(self, hi_res_true, hi_res_gen):
        """Content loss on the output channels only (exo channels sliced off); argument order
        of the loss callable is (gen, true) (base.py:478-503)."""
        n_exo = len(self.hr_exo_features)
        if n_exo:
            c = hi_res_gen.shape[-1]
            crop = [(0, 0)] * (hi_res_gen.dim() - 1) + [(0, n_exo)]
            from ..autograd import CropFn
            hi_res_gen = CropFn.apply(hi_res_gen, crop)
            hi_res_true = ops.crop_fwd(hi_res_true, crop)
            assert hi_res_gen.shape[-1] == c - n_exo
        return self.loss_fun(hi_res_gen, hi_res_true)

    @staticmethod
    def calc_loss_disc(disc_out_true, disc_out_gen):
        """Relativistic average discriminator loss (ESRGAN) (base.py:505-549): mean sigmoid
        cross entropy of [true - mean(gen), gen - mean(true)] against labels [1, 0]."""
        return DiscLossFn.apply(disc_out_true.reshape(-1).contiguous(),
                                disc_out_gen.reshape(-1).contiguous())

    def update_adversarial_weights(self, history, adaptive_update_fraction,
                                   adaptive_update_bounds, weight_gen_advers, train_disc):
        """Adaptive adversarial weight update (base.py:551-606)."""
        if adaptive_update_fraction > 0:
            update_frac = 1
            if train_disc:
                update_frac = self.get_weight_update_fraction(
                    history, "disc_train_frac", update_frac=adaptive_update_fraction,
                    update_bounds=adaptive_update_bounds)
                weight_gen_advers *= update_frac
            if update_frac != 1:
                logger.debug("New discriminator weight: %.4e", weight_gen_advers)
        return weight_gen_advers

    @staticmethod
    def check_batch_handler_attrs(batch_handler):
        keys = ["smoothing", "lr_features", "hr_exo_features", "hr_out_features",
                "smoothed_features"]
        return {k: getattr(batch_handler, k, None) for k in keys if hasattr(batch_handler, k)}

    def calc_loss(self, hi_res_true, hi_res_gen, weight_gen_advers=0.001, train_gen=True,
                  train_disc=False, compute_disc=False):
        """GAN loss (base.py:830-911).  Returns ``(loss, loss_details)`` with keys
        ``loss_disc, loss_gen, loss_gen_content, loss_gen_advers`` + per-term content names."""
        hi_res_gen = self._combine_loss_input(hi_res_true, hi_res_gen)
        if tuple(hi_res_gen.shape) != tuple(hi_res_true.shape):
            msg = ("The tensor shapes of the synthetic output {} and true high res {} did not "
                   "have matching shape! Check the spatiotemporal enhancement multipliers in "
                   "your your model config and data handlers.".format(
                       tuple(hi_res_gen.shape), tuple(hi_res_true.shape)))
            logger.error(msg)
            raise RuntimeError(msg)
        disc_out_true = self._tf_discriminate(hi_res_true)
        disc_out_gen = self._tf_discriminate(hi_res_gen)
        loss_details = {}
        loss = None
        if compute_disc or train_disc:
            loss_details["loss_disc"] = self.calc_loss_disc(disc_out_true, disc_out_gen)
        if train_gen:
            loss_gen_content, content_details = self.calc_loss_gen_content(hi_res_true,
                                                                            hi_res_gen)
            loss_gen_advers = self.calc_loss_disc(disc_out_gen, disc_out_true)
            loss = loss_gen_content + ScaleFnScalar.apply(loss_gen_advers, float(weight_gen_advers))
            loss_details["loss_gen"] = loss
            loss_details["loss_gen_content"] = loss_gen_content
            loss_details["loss_gen_advers"] = loss_gen_advers
            loss_details.update(content_details)
        elif train_disc:
            loss = loss_details["loss_disc"]
        return loss, loss_details

    def calc_val_loss(self, batch_handler, weight_gen_advers):
        """End-of-epoch validation loss (forward only) (base.py:913-942)."""
        logger.debug("Starting end-of-epoch validation loss calculation...")
        for batch in batch_handler.val_data:
            with torch.no_grad():
                _, v_loss_details, _, _ = self._get_hr_exo_and_loss(
                    batch.low_res, batch.high_res, weight_gen_advers=weight_gen_advers)
            self._val_record = self.update_loss_details(
                self._val_record, v_loss_details, len(batch_handler.val_data), prefix="val_")
        return self._val_record.mean(axis=0)

    # ---- training ----------------------------------------------------------------------------
    def _train_batch(self, batch, train_gen, only_gen, gen_too_good, train_disc, only_disc,
                     disc_too_good, weight_gen_advers, multi_gpu=False):
        """One batch of the GAN schedule (base.py:944-1031)."""
        trained_gen = trained_disc = False
        loss_details = {}
        if only_gen or (train_gen and not gen_too_good):
            trained_gen = True
            b = self.timer(self.run_gradient_descent)(
                batch.low_res, batch.high_res, self.generator_weights,
                weight_gen_advers=weight_gen_advers, optimizer=self.optimizer, train_gen=True,
                train_disc=False, compute_disc=train_disc, multi_gpu=multi_gpu)
            loss_details.update(b)
        if only_disc or (train_disc and not disc_too_good):
            trained_disc = True
            b = self.timer(self.run_gradient_descent)(
                batch.low_res, batch.high_res, self.discriminator_weights,
                weight_gen_advers=weight_gen_advers, optimizer=self.optimizer_disc,
                train_gen=False, train_disc=True, multi_gpu=multi_gpu)
            loss_details.update(b)
        loss_details = {k: float(v) for k, v in loss_details.items()}
        loss_details["gen_train_frac"] = float(trained_gen)
        loss_details["disc_train_frac"] = float(trained_disc)
        return loss_details

    def _post_batch(self, ib, b_loss_details, n_batches, previous_means):
        """Running means over the last ``n_batches`` (base.py:1033-1095)."""
        for key, val in previous_means.items():
            if key.startswith("train_"):
                b_loss_details.setdefault(key.replace("train_", ""), val)
        self._train_record = self.update_loss_details(self._train_record, b_loss_details,
                                                      n_batches, prefix="train_")
        if self._tb_writer is not None:
            self.dict_to_tensorboard(b_loss_details)
            self.dict_to_tensorboard(self.timer.log)
        trained_gen = bool(self._train_record["gen_train_frac"].values[-1])
        trained_disc = bool(self._train_record["disc_train_frac"].values[-1])
        disc_loss = self._train_record["train_loss_disc"].values.mean()
        gen_loss = self._train_record["train_loss_gen"].values.mean()
        logger.debug("Batch %d out of %d has (gen / disc) loss of: (%.2e / %.2e). Running mean "
                     "(gen / disc): (%.2e / %.2e). Trained (gen / disc): (%s / %s)", ib + 1,
                     n_batches, b_loss_details["loss_gen"], b_loss_details["loss_disc"], gen_loss,
                     disc_loss, trained_gen, trained_disc)
        if not trained_gen and not trained_disc:
            msg = (f"For some reason none of the GAN networks trained during batch {ib} out of "
                   f"{n_batches}!")
            logger.warning(msg)
            warn(msg)
        return self._train_record.mean(axis=0).to_dict()

    def _train_epoch(self, batch_handler, weight_gen_advers, train_gen, train_disc,
                     disc_loss_bounds, multi_gpu=False):
        """One epoch over the batch handler (base.py:1097-1191)."""
        lr_shape, hr_shape = batch_handler.shapes
        self.init_weights(lr_shape, hr_shape)
        self.init_weights((1, *batch_handler.lr_shape), (1, *batch_handler.hr_shape))
        disc_th_low, disc_th_high = np.min(disc_loss_bounds), np.max(disc_loss_bounds)
        loss_means = self._train_record.mean().to_dict()
        loss_means.setdefault("train_loss_disc", 0)
        loss_means.setdefault("train_loss_gen", 0)
        only_gen = train_gen and not train_disc
        only_disc = train_disc and not train_gen
        for ib, batch in enumerate(batch_handler):
            start = time.time()
            loss_disc = loss_means["train_loss_disc"]
            disc_too_good = loss_disc <= disc_th_low
            disc_too_bad = (loss_disc > disc_th_high) and train_disc
            gen_too_good = disc_too_bad
            b_loss_details = self.timer(self._train_batch, log=True)(
                batch, train_gen, only_gen, gen_too_good, train_disc, only_disc, disc_too_good,
                weight_gen_advers, multi_gpu)
            loss_means = self.timer(self._post_batch, log=True)(ib, b_loss_details,
                                                                len(batch_handler), loss_means)
            logger.info("Finished batch step %d / %d in %.4f seconds", ib + 1,
                        len(batch_handler), time.time() - start)
        self.total_batches += len(batch_handler)
        loss_details = self._train_record.mean().to_dict()
        loss_details["total_batches"] = int(self.total_batches)
        self.profile_to_tensorboard("training_epoch")
        return loss_details

    def train(self, batch_handler, input_resolution, n_epoch, weight_gen_advers=0.001,
              train_gen=True, train_disc=True, disc_loss_bounds=(0.45, 0.6), checkpoint_int=None,
              out_dir="./gan_{epoch}", early_stop_on=None, early_stop_threshold=0.005,
              early_stop_n_epoch=5, adaptive_update_bounds=(0.9, 0.99),
              adaptive_update_fraction=0.0, multi_gpu=False, tensorboard_log=False,
              tensorboard_profile=False):
        """Train the GAN (base.py:624-828).  ``batch_handler`` must provide ``means/stds``,
        ``s_enhance/t_enhance``, ``shapes``, ``lr_shape/hr_shape``, ``__len__/__iter__``
        yielding ``.low_res/.high_res``, ``val_data`` and ``stop()``."""
        if tensorboard_log:
            self._init_tensorboard_writer(out_dir)
        if tensorboard_profile:
            self._write_tb_profile = True
        self.set_norm_stats(batch_handler.means, batch_handler.stds)
        params = self.check_batch_handler_attrs(batch_handler)
        self.set_model_params(input_resolution=input_resolution,
                              s_enhance=batch_handler.s_enhance,
                              t_enhance=batch_handler.t_enhance, **params)
        epochs = list(range(n_epoch))
        if self._history is None:
            self._history = pd.DataFrame(columns=["elapsed_time"])
            self._history.index.name = "epoch"
        else:
            epochs = [e + int(self._history.index.values[-1]) + 1 for e in epochs]
        t0 = time.time()
        logger.info("Training model with adversarial weight: %s for %d epochs starting at "
                    "epoch %d", weight_gen_advers, n_epoch, epochs[0])
        for epoch in epochs:
            t_epoch = time.time()
            loss_details = self._train_epoch(batch_handler, weight_gen_advers, train_gen,
                                             train_disc, disc_loss_bounds, multi_gpu=multi_gpu)
            loss_details.update(self.calc_val_loss(batch_handler, weight_gen_advers))
            msg = "Epoch {} of {} gen/disc train loss: {:.2e}/{:.2e} ".format(
                epoch, epochs[-1], loss_details["train_loss_gen"],
                loss_details["train_loss_disc"])
            if "val_loss_gen" in loss_details and "val_loss_disc" in loss_details:
                msg += "gen/disc val loss: {:.2e}/{:.2e} ".format(
                    loss_details["val_loss_gen"], loss_details["val_loss_disc"])
            logger.info(msg)
            extras = {"weight_gen_advers": weight_gen_advers,
                      "disc_loss_bound_0": disc_loss_bounds[0],
                      "disc_loss_bound_1": disc_loss_bounds[1]}
            opt_g = self.get_optimizer_state(self.optimizer)
            opt_d = self.get_optimizer_state(self.optimizer_disc)
            extras.update({f"OptmGen/{k}": v for k, v in opt_g.items()})
            extras.update({f"OptmDisc/{k}": v for k, v in opt_d.items()})
            weight_gen_advers = self.update_adversarial_weights(
                loss_details, adaptive_update_fraction, adaptive_update_bounds,
                weight_gen_advers, train_disc)
            stop = self.finish_epoch(epoch, epochs, t0, loss_details, checkpoint_int, out_dir,
                                     early_stop_on, early_stop_threshold, early_stop_n_epoch,
                                     extras=extras)
            logger.info("Finished training epoch in %.4f seconds", time.time() - t_epoch)
            if stop:
                break
        logger.info("Finished training %d epochs in %.4f seconds", n_epoch, time.time() - t0)
        batch_handler.stop()


class ScaleFnScalar(torch.autograd.Function):
    """loss * w for a python float w (keeps the scaling on our kernels)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.w = w
        return ops.channel_affine(x.reshape(1, 1), torch.full((1,), w, device=x.device),
                                  None).reshape(())

    @staticmethod
    def backward(ctx, dy):
        return ops.channel_affine(dy.reshape(1, 1).contiguous(),
                                  torch.full((1,), ctx.w, device=dy.device),
                                  None).reshape(()), None
