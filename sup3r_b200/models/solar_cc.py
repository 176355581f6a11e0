"""``SolarCC``: solar climate-change GAN (mirrors sup3r/models/solar_cc.py:13-324).

Differences to ``Sup3rGan`` (solar_cc.py:16-30):
  * the pointwise content loss looks at the centre ``POINT_LOSS_HOURS`` of every true /
    synthetic day plus the 24-hour temporal mean of the synthetic day against the daylight mean
    of the true day;
  * the discriminator only sees ``DAYLIGHT_HOURS``-long windows: the fixed daylight window of
    every true day, and randomly placed windows of the synthetic sample;
  * ``generate`` reflect-pads the time axis so that the output is ``low_res_t * t_enhance`` long.

All discriminator / loss arithmetic runs in this library's kernels; the time-window slicing and
the temporal means are device tensor views recorded on the autograd tape.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from ..utilities import RANDOM_GENERATOR
from .base import ScaleFnScalar, Sup3rGan

logger = logging.getLogger(__name__)


class SolarCC(Sup3rGan):
    """Solar climate change model."""

    STARTING_HOUR = 8      # hour at which the daylight window starts (solar_cc.py:32-34)
    DAYLIGHT_HOURS = 8     # length of the window the discriminator sees (solar_cc.py:36-39)
    POINT_LOSS_HOURS = 2   # centre hours of the day used for the pointwise loss (solar_cc.py:41-44)

    def __init__(self, *args, t_enhance=None, **kwargs):
        """``t_enhance``: optional override of the temporal enhancement; when it differs from
        the layers' value ``generate`` pads the output (solar_cc.py:46-66)."""
        super().__init__(*args, **kwargs)
        self._t_enhance = t_enhance or self.t_enhance
        self.meta["t_enhance"] = self._t_enhance

    def init_weights(self, lr_shape, hr_shape, device=None):
        """The discriminator is built for daylight windows only (solar_cc.py:68-92)."""
        if hr_shape[3] != self.DAYLIGHT_HOURS:
            hr_shape = tuple(hr_shape[0:3]) + (self.DAYLIGHT_HOURS,) + tuple(hr_shape[-1:])
        super().init_weights(lr_shape, hr_shape, device=device)

    def _sample_gen_windows(self, t_len, n_days):
        """Start hours of the random synthetic daylight windows (``tf.random.categorical`` over
        uniform logits in the reference, solar_cc.py:189-196)."""
        return [int(t) for t in RANDOM_GENERATOR.integers(0, t_len - self.DAYLIGHT_HOURS + 1,
                                                          size=n_days)]

    def calc_loss(self, hi_res_true, hi_res_gen, weight_gen_advers=0.001, train_gen=True,
                  train_disc=False, compute_disc=False):
        """GAN loss on daylight windows (solar_cc.py:94-264)."""
        if tuple(hi_res_gen.shape) != tuple(hi_res_true.shape):
            msg = ("The tensor shapes of the synthetic output {} and true high res {} did not "
                   "have matching shape! Check the spatiotemporal enhancement multipliers in "
                   "your your model config and data handlers.".format(
                       tuple(hi_res_gen.shape), tuple(hi_res_true.shape)))
            logger.error(msg)
            raise RuntimeError(msg)
        msg = ("Special SolarCC model can only accept multi-day hourly (multiple of 24) true / "
               "synthetic high res data in the axis=3 position but received shape {}".format(
                   tuple(hi_res_true.shape)))
        assert hi_res_true.shape[3] % 24 == 0, msg
        t_len = int(hi_res_true.shape[3])
        n_days = t_len // 24
        day_24h = [slice(x, x + 24) for x in range(0, 24 * n_days, 24)]
        sub_day = [slice(self.STARTING_HOUR + x, self.STARTING_HOUR + x + self.DAYLIGHT_HOURS)
                   for x in range(0, 24 * n_days, 24)]
        p0 = (24 - self.POINT_LOSS_HOURS) // 2
        point = [slice(p0 + x, p0 + x + self.POINT_LOSS_HOURS) for x in range(0, 24 * n_days, 24)]

        disc_out_gen = []
        for t0 in self._sample_gen_windows(t_len, n_days):
            win = hi_res_gen[:, :, :, t0:t0 + self.DAYLIGHT_HOURS, :].contiguous()
            disc_out_gen.append(self._tf_discriminate(win))
        disc_out_true = [self._tf_discriminate(hi_res_true[:, :, :, ts, :].contiguous())
                         for ts in sub_day]
        disc_out_true = torch.cat([d.reshape(-1) for d in disc_out_true])
        disc_out_gen = torch.cat([d.reshape(-1) for d in disc_out_gen])

        loss_details = {}
        loss = None
        if compute_disc or train_disc:
            loss_details["loss_disc"] = self.calc_loss_disc(disc_out_true, disc_out_gen)
        if train_gen:
            loss_gen_content = None
            nd = len(sub_day)
            for ts_sub, ts_p, ts_24 in zip(sub_day, point, day_24h):
                hr_true_mean = hi_res_true[:, :, :, ts_sub, :].mean(dim=3)
                hr_gen_mean = hi_res_gen[:, :, :, ts_24, :].mean(dim=3)
                c_sub, c_sub_d = self.calc_loss_gen_content(
                    hi_res_true[:, :, :, ts_p, :].contiguous(),
                    hi_res_gen[:, :, :, ts_p, :].contiguous())
                c_24h, c_24h_d = self.calc_loss_gen_content(hr_true_mean.contiguous(),
                                                            hr_gen_mean.contiguous())
                term = (c_sub + c_24h) / nd
                loss_gen_content = term if loss_gen_content is None else loss_gen_content + term
                for k, v in c_sub_d.items():
                    loss_details[f"c_sub_{k}"] = loss_details.get(f"c_sub_{k}", 0) + v / nd
                for k, v in c_24h_d.items():
                    loss_details[f"c_24h_{k}"] = loss_details.get(f"c_24h_{k}", 0) + v / nd
            loss_gen_advers = self.calc_loss_disc(disc_out_gen, disc_out_true)
            loss = loss_gen_content + ScaleFnScalar.apply(loss_gen_advers,
                                                          float(weight_gen_advers))
            loss_details["loss_gen"] = loss
            loss_details["loss_gen_content"] = loss_gen_content
            loss_details["loss_gen_advers"] = loss_gen_advers
        elif train_disc:
            loss = loss_details["loss_disc"]
        return loss, loss_details

    def temporal_pad(self, low_res, hi_res, mode="reflect"):
        """Pad the time axis of the generated array to ``low_res_t * t_enhance``
        (solar_cc.py:266-296)."""
        t_shape = low_res.shape[-2] * self._t_enhance
        t_pad = int((t_shape - hi_res.shape[-2]) / 2)
        pad_width = ((0, 0), (0, 0), (0, 0), (t_pad, t_pad), (0, 0))
        prepad = hi_res.shape
        hi_res = np.pad(hi_res, pad_width, mode=mode)
        logger.debug("Padded hi_res output from %s to %s", prepad, hi_res.shape)
        return hi_res

    def generate(self, low_res, **kwargs):
        """Parent ``generate`` + temporal padding (solar_cc.py:298-308)."""
        hi_res = super().generate(low_res=low_res, **kwargs)
        if isinstance(hi_res, torch.Tensor):
            hi_res = hi_res.detach().cpu().numpy()
        hi_res = self.temporal_pad(low_res, hi_res)
        logger.debug("Final SolarCC output has shape: %s", hi_res.shape)
        return hi_res

    @classmethod
    def load(cls, model_dir, t_enhance=None, verbose=True, **kwargs):
        """(solar_cc.py:299-324)"""
        fp_gen, fp_disc, params = cls._load(model_dir, verbose=verbose)
        params.update(kwargs)
        return cls(fp_gen, fp_disc, t_enhance=t_enhance, **params)
