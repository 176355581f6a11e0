"""``MultiStepGan``: a serial chain of generators (mirrors sup3r/models/multi_step.py:23-337):
4-D <-> 5-D transposes between spatial and spatiotemporal steps, feature matching by name,
per-step exogenous data, un-normalise / re-normalise between steps."""
from __future__ import annotations

import json
import logging
import os

import numpy as np
import torch

from ..exo import ExoData
from .base import Sup3rGan
from .interface import AbstractInterface

logger = logging.getLogger(__name__)


def _model_class(name):
    from .. import models as pkg
    cls = getattr(pkg, name, None)
    if cls is None:
        raise KeyError(f'Could not find requested model class "{name}" in sup3r_b200.models')
    return cls


class MultiStepGan(AbstractInterface):
    """Multi-step (spatial -> ... -> spatiotemporal) generator chain."""

    def __init__(self, models):
        self._models = tuple(models)

    def __len__(self):
        return len(self._models)

    @classmethod
    def load(cls, model_dirs, model_kwargs=None, verbose=True):
        """Load every step; the class of each comes from its ``model_params.json`` meta
        (multi_step.py:45-85)."""
        if isinstance(model_dirs, str):
            model_dirs = [model_dirs]
        model_kwargs = model_kwargs or [{}] * len(model_dirs)
        if isinstance(model_kwargs, dict):
            model_kwargs = [model_kwargs]
        models = []
        for model_dir, kwargs in zip(model_dirs, model_kwargs):
            fp_params = os.path.join(model_dir, "model_params.json")
            assert os.path.exists(fp_params), f"Could not find: {fp_params}"
            with open(fp_params) as f:
                params = json.load(f)
            class_name = params.get("meta", {"class": "Sup3rGan"}).get("class", "Sup3rGan")
            models.append(_model_class(class_name).load(model_dir, verbose=verbose, **kwargs))
        return cls(models)

    @property
    def models(self):
        return self._models

    @property
    def means(self):
        return tuple(m.means for m in self.models)

    @property
    def stdevs(self):
        return tuple(m.stdevs for m in self.models)

    @staticmethod
    def seed(s=0):
        Sup3rGan.seed(s=s)

    def _transpose_model_input(self, model, hi_res):
        """(t, s1, s2, c) <-> (1, s1, s2, t, c) between 4-D and 5-D steps
        (multi_step.py:128-170)."""
        on_dev = isinstance(hi_res, torch.Tensor)
        if model.is_5d and hi_res.ndim == 4:
            hi_res = (hi_res.permute(1, 2, 0, 3)[None] if on_dev
                      else np.transpose(hi_res, axes=(1, 2, 0, 3))[np.newaxis])
        elif model.is_4d and hi_res.ndim == 5:
            assert hi_res.shape[0] == 1, (
                f"Recieved 5D input data with shape ({tuple(hi_res.shape)}) to a 4D model.")
            hi_res = (hi_res[0].permute(2, 0, 1, 3) if on_dev
                      else np.transpose(hi_res[0], axes=(2, 0, 1, 3)))
        else:
            assert model.input_dims == hi_res.ndim, (
                f"Recieved input data with shape {tuple(hi_res.shape)} to a {model.input_dims}D "
                "model.")
        return hi_res

    def _match_model_input(self, model_step, hi_res, exo_data):
        """Select the previous step's outputs the current step consumes (by feature name)
        (multi_step.py:172-194)."""
        if model_step > 0:
            out_feats = self.models[model_step - 1].hr_out_features
            in_feats = [f for f in self.models[model_step].lr_features
                        if f not in (exo_data or {})]
            if not set(in_feats).issubset(set(out_feats)):
                msg = (f"Model step {model_step} input features {in_feats} do not match previous "
                       f"model step {model_step - 1} output features {out_feats}")
                logger.error(msg)
                raise ValueError(msg)
            hi_res = hi_res[..., [out_feats.index(fn) for fn in in_feats]]
        return hi_res

    def generate(self, low_res, norm_in=True, un_norm_out=True, exogenous_data=None, **kwargs):
        """Chain ``model.generate`` over the steps (multi_step.py:196-275)."""
        if isinstance(exogenous_data, dict) and not isinstance(exogenous_data, ExoData):
            exogenous_data = ExoData(exogenous_data)
        hi_res = np.array(low_res, copy=True)
        for i, model in enumerate(self.models):
            i_norm_in = not (i == 0 and not norm_in)
            last = i + 1 == len(self.models)
            i_un_norm_out = not (last and not un_norm_out)
            i_exo = None if exogenous_data is None else exogenous_data.get_model_step_exo(i)
            # intermediates stay on the GPU between steps (the reference round-trips numpy
            # arrays and transposes on the host)
            keep_dev = not last and "to_numpy" not in kwargs
            try:
                hi_res = self._transpose_model_input(model, hi_res)
                hi_res = self._match_model_input(i, hi_res, i_exo)
                if not isinstance(hi_res, torch.Tensor):
                    hi_res = np.ascontiguousarray(hi_res)
                hi_res = model.generate(hi_res, norm_in=i_norm_in, un_norm_out=i_un_norm_out,
                                        exogenous_data=i_exo,
                                        **({"to_numpy": False} if keep_dev else {}), **kwargs)
                if keep_dev and i_exo is not None and \
                        len(model.hr_out_features) > hi_res.shape[-1]:
                    # exo channels appended to this step's output: host-side like the reference
                    hi_res = model._combine_fwp_output(hi_res.cpu().numpy(), i_exo)
            except Exception as e:
                msg = (f'Could not run model #{i + 1} of {len(self.models)} "{model}" on tensor '
                       f"of shape {hi_res.shape}")
                logger.exception(msg)
                raise RuntimeError(msg) from e
        return hi_res

    @property
    def version_record(self):
        return tuple(m.version_record for m in self.models)

    @property
    def meta(self):
        return tuple(m.meta for m in self.models)

    @property
    def lr_features(self):
        return self.models[0].lr_features

    @property
    def hr_out_features(self):
        return self.models[-1].hr_out_features

    @property
    def hr_exo_features(self):
        return [m.hr_exo_features for m in self.models]

    @property
    def obs_features(self):
        return [m.obs_features for m in self.models]

    @property
    def model_params(self):
        return tuple(m.model_params for m in self.models)


class SolarMultiStepGan(MultiStepGan):
    """Solar multi-step chain (mirrors sup3r/models/multi_step.py:484-911): a spatial solar
    chain (clearsky_ratio only) and a spatial wind chain run on the same low-res input, their
    outputs are concatenated, transposed to 5-D and fed to the temporal solar chain
    (typically a ``SolarCC``); the time axis is reflect-padded to ``n_t * t_enhance``."""

    def __init__(self, spatial_solar_models, spatial_wind_models, temporal_solar_models,
                 t_enhance=None):
        super().__init__(models=[*spatial_wind_models.models, *temporal_solar_models.models])
        self._spatial_solar_models = spatial_solar_models
        self._spatial_wind_models = spatial_wind_models
        self._temporal_solar_models = temporal_solar_models
        self._t_enhance = t_enhance
        self.preflight()
        if self._t_enhance is not None:
            msg = "Can only update t_enhance for a single temporal solar model."
            assert len(self.temporal_solar_models) == 1, msg
            self.temporal_solar_models.models[0].meta["t_enhance"] = self._t_enhance

    def preflight(self):
        """Consistency checks between the three chains (multi_step.py:561-611)."""
        s_enh = self.spatial_solar_models.s_enhancements
        w_enh = self.spatial_wind_models.s_enhancements
        msg = ("Solar and wind spatial enhancements must be equivalent but received models that "
               "do spatial enhancements of {} (solar) and {} (wind)".format(s_enh, w_enh))
        assert np.prod(s_enh) == np.prod(w_enh), msg
        s_t_feat = self.spatial_solar_models.lr_features
        s_o_feat = self.spatial_solar_models.hr_out_features
        msg = ('Solar spatial enhancement models need to take "clearsky_ratio" as the only input '
               "and output feature but received models that need {} and output {}".format(
                   s_t_feat, s_o_feat))
        assert s_t_feat == ["clearsky_ratio"], msg
        assert s_o_feat == ["clearsky_ratio"], msg
        temp_solar_feats = self.temporal_solar_models.lr_features
        msg = ('Input feature 0 for the temporal_solar_models should be "clearsky_ratio" but '
               "received: {}".format(temp_solar_feats))
        assert temp_solar_feats[0] == "clearsky_ratio", msg
        spatial_out = (self.spatial_wind_models.hr_out_features
                       + self.spatial_solar_models.hr_out_features)
        missing = [fn for fn in temp_solar_feats if fn not in spatial_out]
        msg = ("Solar temporal model needs features {} that were not found in the solar + wind "
               "model output feature list {}".format(missing, spatial_out))
        assert not any(missing), msg

    @property
    def spatial_solar_models(self):
        return self._spatial_solar_models

    @property
    def spatial_wind_models(self):
        return self._spatial_wind_models

    @property
    def temporal_solar_models(self):
        return self._temporal_solar_models

    @property
    def meta(self):
        return (self.spatial_solar_models.meta + self.spatial_wind_models.meta
                + self.temporal_solar_models.meta)

    @property
    def lr_features(self):
        return self.spatial_solar_models.lr_features + self.spatial_wind_models.lr_features

    @property
    def hr_out_features(self):
        return self.temporal_solar_models.hr_out_features

    @property
    def idf_wind(self):
        """Indices of the wind-chain input features in the low-res input."""
        return np.array([self.lr_features.index(fn)
                         for fn in self.spatial_wind_models.lr_features if fn != "topography"])

    @property
    def idf_wind_out(self):
        """Indices of the wind-chain outputs the temporal solar chain consumes."""
        return np.array([self.spatial_wind_models.hr_out_features.index(fn)
                         for fn in self.temporal_solar_models.lr_features[1:]])

    @property
    def idf_solar(self):
        return np.array([self.lr_features.index(fn)
                         for fn in self.spatial_solar_models.lr_features if fn != "topography"])

    def generate(self, low_res, norm_in=True, un_norm_out=True, exogenous_data=None):
        """(multi_step.py:679-790) ``low_res``: (n_t, s1, s2, features) with the solar feature
        first; returns (1, S1, S2, n_t * t_enhance, features)."""
        logger.debug("Data input to the SolarMultiStepGan has shape %s which will be split up "
                     "for solar- and wind-only features.", low_res.shape)
        if isinstance(exogenous_data, dict) and not isinstance(exogenous_data, ExoData):
            exogenous_data = ExoData(exogenous_data)
        if exogenous_data is not None:
            s_exo, t_exo = exogenous_data.split(split_steps=[len(self.spatial_wind_models)])
        else:
            s_exo = t_exo = None
        low_res = np.asarray(low_res)
        try:
            hi_res_wind = self.spatial_wind_models.generate(
                low_res[..., self.idf_wind], norm_in=norm_in, un_norm_out=True,
                exogenous_data=s_exo)
        except Exception as e:
            msg = ("Could not run the 1st step spatial-wind-only GAN on input shape "
                   "{}".format(low_res.shape))
            logger.exception(msg)
            raise RuntimeError(msg) from e
        try:
            hi_res_solar = self.spatial_solar_models.generate(
                low_res[..., self.idf_solar], norm_in=norm_in, un_norm_out=True)
        except Exception as e:
            msg = ("Could not run the 1st step spatial-solar-only GAN on input shape "
                   "{}".format(low_res.shape))
            logger.exception(msg)
            raise RuntimeError(msg) from e
        logger.debug("Data output from the 1st step spatial enhancement has shape %s (solar) and "
                     "shape %s (wind)", hi_res_solar.shape, hi_res_wind.shape)
        hi_res = np.concatenate((hi_res_solar, hi_res_wind[..., self.idf_wind_out]), axis=3)
        hi_res = np.expand_dims(np.transpose(hi_res, axes=(1, 2, 0, 3)), axis=0)
        try:
            hi_res = self.temporal_solar_models.generate(
                np.ascontiguousarray(hi_res), norm_in=True, un_norm_out=un_norm_out,
                exogenous_data=t_exo)
        except Exception as e:
            msg = ("Could not run the 2nd step (spatio)temporal solar GAN on input shape "
                   "{}".format(low_res.shape))
            logger.exception(msg)
            raise RuntimeError(msg) from e
        hi_res = self.temporal_pad(low_res, hi_res)
        logger.debug("Final SolarMultiStepGan output has shape: %s", hi_res.shape)
        return hi_res

    def temporal_pad(self, low_res, hi_res, mode="reflect"):
        """(multi_step.py:792-823)"""
        t_shape = low_res.shape[0] * self.t_enhance
        t_pad = int((t_shape - hi_res.shape[-2]) / 2)
        return np.pad(hi_res, ((0, 0), (0, 0), (0, 0), (t_pad, t_pad), (0, 0)), mode=mode)

    @classmethod
    def load(cls, spatial_solar_model_dirs, spatial_wind_model_dirs, temporal_solar_model_dirs,
             t_enhance=None, verbose=True):
        """(multi_step.py:825-911)"""
        ssm = MultiStepGan.load(spatial_solar_model_dirs, verbose=verbose)
        swm = MultiStepGan.load(spatial_wind_model_dirs, verbose=verbose)
        tsm = MultiStepGan.load(temporal_solar_model_dirs, verbose=verbose)
        return cls(ssm, swm, tsm, t_enhance=t_enhance)
