"""Model meta-data interface (mirrors sup3r/models/interface.py:22-517): enhancement factors
from layer attributes, feature lists, resolutions, exo combination at input / output,
``set_model_params`` validation, ``save_params``."""
from __future__ import annotations

import json
import logging
import os
import re
from warnings import warn

import numpy as np

from ..exo import ExoData
from ..network import CustomNetwork, SUP3R_EXO_LAYERS, SUP3R_OBS_LAYERS
from ..utilities import VERSION_RECORD, safe_cast

logger = logging.getLogger(__name__)


class AbstractInterface:
    """Interface shared by single and multi-step models."""

    # ---- abstract in the reference (interface.py:32-57, 358-361): every model class defines them
    @classmethod
    def load(cls, model_dir, verbose=True):
        """Load the model from a previously saved-to output directory."""
        raise NotImplementedError(f"{cls.__name__} must define load()")

    def generate(self, low_res, norm_in=True, un_norm_out=True, exogenous_data=None):
        """High-res data from low-res input: the public generate function."""
        raise NotImplementedError(f"{type(self).__name__} must define generate()")

    @property
    def meta(self):
        """Meta data dictionary that defines how the model was created."""
        raise NotImplementedError(f"{type(self).__name__} must define meta")

    @staticmethod
    def seed(s=0):
        """Seed weight initialisation for reproducible results (interface.py:59-69)."""
        CustomNetwork.seed(s)

    # ---- dimensions -------------------------------------------------------------------
    @property
    def input_dims(self):
        if hasattr(self, "_gen"):
            return self._gen.layers[0].rank
        if hasattr(self, "models"):
            return self.models[0].input_dims
        return 5

    @property
    def is_5d(self):
        return self.input_dims == 5

    @property
    def is_4d(self):
        return self.input_dims == 4

    def _enhance_from_layers(self, attr):
        if not hasattr(self, "_gen"):
            return None
        return int(np.prod([getattr(lyr, attr, 1) for lyr in self._gen.layers]))

    def get_s_enhance_from_layers(self):
        return self._enhance_from_layers("_spatial_mult")

    def get_t_enhance_from_layers(self):
        return self._enhance_from_layers("_temporal_mult")

    def _enhance(self, key, from_layers):
        models = getattr(self, "models", [self])
        vals = [m.meta.get(key, None) for m in models]
        val = from_layers() if any(v is None for v in vals) else int(np.prod(vals))
        if len(models) == 1 and isinstance(self.meta, dict):
            self.meta[key] = val
        return val

    @property
    def s_enhance(self):
        return self._enhance("s_enhance", self.get_s_enhance_from_layers)

    @property
    def t_enhance(self):
        return self._enhance("t_enhance", self.get_t_enhance_from_layers)

    @property
    def s_enhancements(self):
        if hasattr(self, "models"):
            return [m.s_enhance for m in self.models]
        return [self.s_enhance]

    @property
    def t_enhancements(self):
        if hasattr(self, "models"):
            return [m.t_enhance for m in self.models]
        return [self.t_enhance]

    # ---- resolutions ------------------------------------------------------------------
    @property
    def input_resolution(self):
        res = self.meta.get("input_resolution", None)
        assert res is not None, "model.input_resolution is None. This needs to be set."
        return res

    def _get_numerical_resolutions(self):
        ires = {k: int(re.search(r"\d+", v).group(0)) for k, v in self.input_resolution.items()}
        enh = {"spatial": self.s_enhance, "temporal": self.t_enhance}
        return ires, {k: v // enh[k] for k, v in ires.items()}

    def _ensure_valid_input_resolution(self):
        if self.meta.get("input_resolution", None) is None:
            return
        ires, ores = self._get_numerical_resolutions()
        s, t = self.meta["s_enhance"], self.meta["t_enhance"]
        ok = ires["temporal"] / ores["temporal"] == t and ires["spatial"] / ores["spatial"] == s
        if not ok:
            msg = (f"Enhancement factors (s_enhance={s}, t_enhance={t}) do not evenly divide "
                   f"input resolution ({self.input_resolution})")
            logger.error(msg)
            raise RuntimeError(msg)

    def _ensure_valid_enhancement_factors(self):
        t, s = self.meta.get("t_enhance", None), self.meta.get("s_enhance", None)
        if s is None or t is None:
            return
        ls, lt = self.get_s_enhance_from_layers(), self.get_t_enhance_from_layers()
        ls = ls if ls is not None else s
        lt = lt if lt is not None else t
        if not (ls == s or lt == t):
            msg = (f"Enhancement factors computed from layer attributes (s_enhance={ls}, "
                   f"t_enhance={lt}) conflict with user provided values (s_enhance={s}, "
                   f"t_enhance={t})")
            logger.error(msg)
            raise RuntimeError(msg)

    @property
    def output_resolution(self):
        out = self.meta.get("output_resolution", None)
        if out is None and self.meta.get("input_resolution", None) is not None:
            ires, ores = self._get_numerical_resolutions()
            out = {k: v.replace(str(ires[k]), str(ores[k]))
                   for k, v in self.input_resolution.items()}
            self.meta["output_resolution"] = out
        return out

    # ---- exogenous data at input / output resolution ------------------------------------
    def _combine_exo(self, arr, exogenous_data, features, combine_type):
        if exogenous_data is None:
            return arr
        if not isinstance(exogenous_data, ExoData):
            exogenous_data = ExoData(exogenous_data)
        n_missing = len(features) - arr.shape[-1]
        exo_feats = [] if n_missing <= 0 else features[-n_missing:]
        assert all(f in exogenous_data for f in exo_feats), (
            f"Provided exogenous_data: {exogenous_data} is missing some required features "
            f"({exo_feats})")
        for f in exo_feats:
            data = exogenous_data.get_combine_type_data(f, combine_type)
            if data is not None:
                arr = np.concatenate((arr, data), axis=-1)
        return arr

    def _combine_fwp_input(self, low_res, exogenous_data=None):
        """Append input-resolution exo channels to ``low_res`` (interface.py:259-307)."""
        return self._combine_exo(low_res, exogenous_data, self.lr_features, "input")

    def _combine_fwp_output(self, hi_res, exogenous_data=None):
        """Append output-resolution exo channels to ``hi_res`` (interface.py:309-358)."""
        return self._combine_exo(hi_res, exogenous_data, self.hr_out_features, "output")

    # ---- features -----------------------------------------------------------------------
    @property
    def lr_features(self):
        return self.meta.get("lr_features", [])

    @property
    def hr_out_features(self):
        return self.meta.get("hr_out_features", [])

    @property
    def obs_features(self):
        feats = []
        if hasattr(self, "_gen") and SUP3R_OBS_LAYERS:
            for lyr in self._gen.layers:
                if isinstance(lyr, SUP3R_OBS_LAYERS):
                    feats += [f for f in getattr(lyr, "features", [lyr.name]) if f not in feats]
        return feats

    @property
    def hr_exo_features(self):
        feats = []
        if hasattr(self, "_gen"):
            feats = [lyr.name for lyr in self._gen.layers if isinstance(lyr, SUP3R_EXO_LAYERS)]
        obs = [f.replace("_obs", "") for f in self.obs_features]
        return feats + [f for f in obs if f not in self.hr_out_features]

    @property
    def hr_features(self):
        return self.hr_out_features + self.hr_exo_features

    @property
    def smoothing(self):
        return self.meta.get("smoothing", None)

    @property
    def smoothed_features(self):
        return self.meta.get("smoothed_features", [])

    @property
    def model_params(self):
        return {"meta": self.meta}

    @property
    def version_record(self):
        return VERSION_RECORD

    def set_model_params(self, **kwargs):
        """Record training parameters in ``meta`` and validate them (interface.py:453-499)."""
        keys = ("input_resolution", "lr_features", "hr_exo_features", "hr_out_features",
                "smoothed_features", "s_enhance", "t_enhance", "smoothing")
        keys = [k for k in keys if k in kwargs]
        if "hr_out_features" in kwargs:
            self.meta["hr_out_features"] = kwargs["hr_out_features"]
        hr_exo = kwargs.get("hr_exo_features", [])
        assert list(self.hr_exo_features) == list(hr_exo), (
            f"Expected high-res exo features {self.hr_exo_features} based on model architecture "
            f'but received "hr_exo_features" from data handler: {hr_exo}')
        for var in keys:
            val = self.meta.get(var, None)
            if val is None:
                self.meta[var] = kwargs[var]
            elif val != kwargs[var]:
                msg = (f"Model was previously trained with {var}={val} but received new "
                       f"{var}={kwargs[var]}")
                logger.warning(msg)
                warn(msg)
        self._ensure_valid_enhancement_factors()
        self._ensure_valid_input_resolution()

    def save_params(self, out_dir):
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "model_params.json"), "w") as f:
            json.dump(self.model_params, f, sort_keys=True, indent=2, default=safe_cast)
