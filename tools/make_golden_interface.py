"""Golden records of the model meta API from the REAL reference class:
``AbstractInterface`` (sup3r/models/interface.py:23-517) and ``ExoData`` are exec'd from their
source text (phygnn / tensorflow stubbed) and driven through a stand-in subclass: enhancement
factors from layers / meta, feature lists incl. exo and observation layers, resolutions,
``set_model_params`` (first values, conflicting values -> warning, invalid resolution /
enhancement -> exception), ``_combine_fwp_input`` / ``_combine_fwp_output``.

    python tools/make_golden_interface.py   ->  tests/golden/interface.npz
"""
import json
import os
import re
from abc import ABC, abstractmethod
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "interface.npz")


class Layer:
    def __init__(self, name="layer", **attrs):
        self.name = name
        self.__dict__.update(attrs)


class ExoLayer(Layer):
    pass


class ObsLayer(Layer):
    pass


def load_reference(warn_log):
    ns = {"np": np, "json": json, "os": os, "re": re, "logger": MagicMock(), "ABC": ABC,
          "abstractmethod": abstractmethod, "warn": lambda m, *a, **k: warn_log.append(str(m)),
          "CustomNetwork": MagicMock(), "VERSION_RECORD": {"sup3r": "ref"},
          "safe_cast": lambda v: v, "SUP3R_EXO_LAYERS": (ExoLayer,),
          "SUP3R_OBS_LAYERS": (ObsLayer,), "locale": MagicMock()}
    src = open(os.path.join(REF, "sup3r/preprocessing/data_handlers/exo.py")).read()
    exec(compile(src[src.index("class SingleExoDataStep"):src.index("class ExoDataHandler")],
                 "exo.py", "exec"), ns)
    src = open(os.path.join(REF, "sup3r/models/interface.py")).read()
    exec(compile(src[src.index("class AbstractInterface"):], "interface.py", "exec"), ns)
    return ns["AbstractInterface"], ns["ExoData"]


def make_model(Base, meta, layers=None):
    class M(Base):
        def __init__(self):
            self._meta = meta
            if layers is not None:
                self._gen = SimpleNamespace(layers=layers)

        @property
        def meta(self):
            return self._meta

        def generate(self, *a, **k):
            raise NotImplementedError

        @classmethod
        def load(cls, *a, **k):
            raise NotImplementedError
    return M()


def attempt(fn):
    try:
        return ["ok", fn()]
    except Exception as e:      # noqa: BLE001 - the exception TYPE is what is pinned
        return ["raises", type(e).__name__]


def scenario(Base, warn_log):
    """-> (json-able record, {name: array})"""
    rec, arrs = {}, {}
    layers = [Layer("pad", rank=5), Layer("conv"), Layer("st", _spatial_mult=1, _temporal_mult=2),
              Layer("st2", _spatial_mult=3, _temporal_mult=3), ExoLayer("topography"),
              ObsLayer("u_10m_obs"), ObsLayer("obs2", features=["v_10m_obs", "temperature_2m_obs"]),
              ExoLayer("sza")]
    m = make_model(Base, {}, layers)
    rec["from_layers"] = [m.s_enhance, m.t_enhance, m.s_enhancements, m.t_enhancements,
                          dict(m.meta)]
    m2 = make_model(Base, {"s_enhance": 5, "t_enhance": 24, "lr_features": ["u_10m", "v_10m"],
                           "hr_out_features": ["u_10m", "v_10m"]}, layers)
    rec["from_meta"] = [m2.s_enhance, m2.t_enhance, m2.lr_features, m2.hr_out_features,
                        m2.obs_features, m2.hr_exo_features, m2.hr_features, m2.smoothing,
                        m2.smoothed_features]
    m3 = make_model(Base, {})
    rec["no_gen"] = [m3.get_s_enhance_from_layers(), m3.get_t_enhance_from_layers(),
                     m3.obs_features, m3.hr_exo_features]
    rec["input_resolution_missing"] = attempt(lambda: make_model(Base, {}, layers).input_resolution)
    # resolutions
    m4 = make_model(Base, {"input_resolution": {"spatial": "30km", "temporal": "60min"},
                           "s_enhance": 3, "t_enhance": 4})
    rec["resolutions"] = [m4.input_resolution, list(m4._get_numerical_resolutions()),
                          m4.output_resolution]
    # set_model_params
    exo_feats = ["topography", "sza", "u_10m", "temperature_2m"]
    m5 = make_model(Base, {}, layers)
    kw = dict(input_resolution={"spatial": "12km", "temporal": "60min"},
              lr_features=["u_10m", "v_10m", "topography"], hr_exo_features=exo_feats,
              hr_out_features=["v_10m"], smoothed_features=["u_10m"], s_enhance=3, t_enhance=6,
              smoothing=None, not_a_param=1)
    n0 = len(warn_log)
    rec["set_first"] = [attempt(lambda: m5.set_model_params(**kw)), dict(m5.meta),
                        len(warn_log) - n0]
    n0 = len(warn_log)
    kw2 = dict(kw, lr_features=["u_10m"], s_enhance=3)
    rec["set_conflict"] = [attempt(lambda: m5.set_model_params(**kw2)), dict(m5.meta),
                           len(warn_log) - n0]
    rec["set_bad_exo"] = attempt(lambda: make_model(Base, {}, layers).set_model_params(
        **dict(kw, hr_exo_features=["topography"])))
    rec["set_bad_enhance"] = attempt(lambda: make_model(Base, {}, layers).set_model_params(
        **dict(kw, s_enhance=2)))
    # both factors differ from the layers' (one matching factor is enough for the reference's
    # "or" check) while the resolutions still divide evenly: the enhancement check itself
    rec["set_bad_both_enhance"] = attempt(lambda: make_model(Base, {}, layers).set_model_params(
        **dict(kw, s_enhance=2, t_enhance=3)))
    rec["set_bad_resolution"] = attempt(lambda: make_model(Base, {}, layers).set_model_params(
        **dict(kw, input_resolution={"spatial": "10km", "temporal": "60min"})))
    rec["set_bad_t_resolution"] = attempt(lambda: make_model(Base, {}, layers).set_model_params(
        **dict(kw, input_resolution={"spatial": "12km", "temporal": "50min"})))
    # exo combination
    rng = np.random.default_rng(3)
    low = rng.standard_normal((2, 4, 5, 3, 2)).astype(np.float32)
    topo_lr = rng.standard_normal((2, 4, 5, 3, 1)).astype(np.float32)
    sza_hr = rng.standard_normal((2, 8, 10, 6, 1)).astype(np.float32)
    exo = {"topography": {"steps": [{"combine_type": "input", "data": topo_lr},
                                    {"combine_type": "layer", "data": sza_hr}]},
           "sza": {"steps": [{"combine_type": "output", "data": sza_hr}]}}
    m6 = make_model(Base, {"lr_features": ["u", "v", "topography"],
                           "hr_out_features": ["u", "v", "sza"]})
    arrs["combine_in"] = np.asarray(m6._combine_fwp_input(low, exo))
    arrs["combine_in_none"] = np.asarray(m6._combine_fwp_input(low, None))
    hi = rng.standard_normal((2, 8, 10, 6, 2)).astype(np.float32)
    arrs["combine_out"] = np.asarray(m6._combine_fwp_output(hi, exo))
    arrs["combine_out_full"] = np.asarray(m6._combine_fwp_output(
        np.concatenate([hi, sza_hr], -1), exo))
    rec["combine_missing"] = attempt(lambda: m6._combine_fwp_input(low, {"sza": exo["sza"]}))
    return rec, arrs


def main():
    warn_log = []
    Base, _ = load_reference(warn_log)
    rec, arrs = scenario(Base, warn_log)
    np.savez_compressed(OUT, record=json.dumps(rec), **arrs)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    print(json.dumps(rec, indent=1)[:1800])


if __name__ == "__main__":
    main()
