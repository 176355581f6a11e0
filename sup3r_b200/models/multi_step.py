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
