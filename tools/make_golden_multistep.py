"""Golden records of the multi-step chains from the REAL reference
classes: ``AbstractInterface`` (sup3r/models/interface.py), ``MultiStepGan`` and
``SolarMultiStepGan`` (sup3r/models/multi_step.py) and ``ExoData`` (data_handlers/exo.py) are
exec'd from their source text with tensorflow / phygnn stubbed out; the chain steps are stand-in
models whose ``generate`` is a small deterministic numpy map that also encodes the flags and exo
data it was called with, so the final arrays pin transposes, feature matching, flag routing, exo
routing, the solar / wind split and the temporal pad.

    python tools/make_golden_multistep.py   ->  tests/golden/multistep.npz
"""
import json
import os
import re
from abc import ABC, abstractmethod
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "multistep.npz")


class StepModel:
    """Stand-in for one trained GAN of a chain: nearest-neighbour enhancement, a fixed feature
    mixing matrix, and additive markers for every argument ``generate`` receives."""

    def __init__(self, ndim, lr_features, hr_out_features, s_enhance=1, t_enhance=1, seed=0):
        self.input_dims = ndim
        self.is_4d, self.is_5d = ndim == 4, ndim == 5
        self.lr_features, self.hr_out_features = list(lr_features), list(hr_out_features)
        self.hr_exo_features, self.obs_features = [], []
        self.s_enhance, self.t_enhance = s_enhance, t_enhance
        self.s_enhancements, self.t_enhancements = [s_enhance], [t_enhance]
        self.meta = {"lr_features": self.lr_features, "hr_out_features": self.hr_out_features,
                     "s_enhance": s_enhance, "t_enhance": t_enhance}
        rng = np.random.default_rng(seed)
        self.mix = rng.uniform(-1, 1, (len(self.lr_features), len(self.hr_out_features)))
        self.calls = []

    def __repr__(self):
        return f"StepModel({self.input_dims}D x{self.s_enhance}/{self.t_enhance})"

    def generate(self, low_res, norm_in=True, un_norm_out=True, exogenous_data=None, **kwargs):
        x = np.asarray(low_res, dtype=np.float64)
        exo_keys = sorted(exogenous_data) if exogenous_data is not None else []
        # (exo features are declared in lr_features but arrive through exogenous_data)
        assert x.ndim == self.input_dims
        assert x.shape[-1] == len([f for f in self.lr_features if f not in exo_keys])
        self.calls.append([list(x.shape), bool(norm_in), bool(un_norm_out), exo_keys])
        y = x @ self.mix[:x.shape[-1]]
        y = np.repeat(np.repeat(y, self.s_enhance, axis=1), self.s_enhance, axis=2)
        if self.is_5d:
            y = np.repeat(y, self.t_enhance, axis=3)
        y = y + 0.5 * norm_in + 0.25 * un_norm_out
        for k in exo_keys:
            for step in exogenous_data[k]["steps"]:
                y = y + 1e-3 * float(np.mean(step["data"]))
        return y


def load_reference_classes():
    ns = {"np": np, "json": json, "os": os, "re": re, "logger": MagicMock(), "ABC": ABC,
          "abstractmethod": abstractmethod, "warn": lambda *a, **k: None,
          "CustomNetwork": MagicMock(), "VERSION_RECORD": {}, "safe_cast": lambda v: v,
          "SUP3R_EXO_LAYERS": (), "SUP3R_OBS_LAYERS": (), "locale": MagicMock(),
          "Sup3rGan": MagicMock(), "sup3r": MagicMock()}
    src = open(os.path.join(REF, "sup3r/preprocessing/data_handlers/exo.py")).read()
    exec(compile(src[src.index("class SingleExoDataStep"):src.index("class ExoDataHandler")],
                 "exo.py", "exec"), ns)
    src = open(os.path.join(REF, "sup3r/models/interface.py")).read()
    exec(compile(src[src.index("class AbstractInterface"):], "interface.py", "exec"), ns)
    src = open(os.path.join(REF, "sup3r/models/multi_step.py")).read()
    a, b = src.index("class MultiStepGan"), src.index("class MultiStepSurfaceMetGan")
    c = src.index("class SolarMultiStepGan")
    exec(compile(src[a:b] + "\n\n" + src[c:], "multi_step.py", "exec"), ns)
    return ns["MultiStepGan"], ns["SolarMultiStepGan"], ns["ExoData"]


def chain_models(topo=False):
    """sup3rwind-shaped chain: 3x spatial (4-D) -> 2x spatial (4-D, re-ordered subset of the
    features) -> 4x temporal (5-D); ``topo``: the spatial steps also take topography (exo)."""
    f5 = ["u_10m", "v_10m", "u_100m", "v_100m", "temperature_2m"]
    extra = ["topography"] if topo else []
    return [StepModel(4, f5 + extra, f5, s_enhance=3, seed=1),
            StepModel(4, ["u_100m", "v_100m", "u_10m"] + extra,
                      ["u_100m", "v_100m", "u_10m"], s_enhance=2, seed=2),
            StepModel(5, ["u_10m", "u_100m"], ["u_10m", "u_100m", "extra"], s_enhance=1,
                      t_enhance=4, seed=3)]


def chain_inputs():
    rng = np.random.default_rng(5)
    low = rng.standard_normal((6, 4, 5, 5))
    exo = {"topography": {"steps": [
        {"model": 0, "combine_type": "input", "data": rng.standard_normal((4, 5, 1))},
        {"model": 1, "combine_type": "input", "data": rng.standard_normal((12, 15, 1))},
        {"model": 1, "combine_type": "layer", "data": rng.standard_normal((24, 30, 1))}]}}
    return low, exo


def solar_models(MultiStepGan, topo=False):
    wind_feats = ["u_200m", "v_200m", "temperature_2m", "topography"] if topo else \
        ["u_200m", "v_200m", "temperature_2m"]
    s_solar = MultiStepGan([StepModel(4, ["clearsky_ratio"], ["clearsky_ratio"], s_enhance=2,
                                      seed=11),
                            StepModel(4, ["clearsky_ratio"], ["clearsky_ratio"], s_enhance=2,
                                      seed=12)])
    s_wind = MultiStepGan([StepModel(4, wind_feats, wind_feats[:3], s_enhance=4, seed=13)])
    t_solar = MultiStepGan([StepModel(5, ["clearsky_ratio", "temperature_2m", "u_200m"],
                                      ["clearsky_ratio", "ghi", "dni"], s_enhance=1,
                                      t_enhance=8, seed=14)])
    # (the real temporal solar model crops the ends of the day: emulate with a shorter output)
    inner = t_solar.models[0].generate

    def cropped(*a, **k):
        return inner(*a, **k)[:, :, :, 4:-4]
    t_solar.models[0].generate = cropped
    return s_solar, s_wind, t_solar


def solar_inputs():
    rng = np.random.default_rng(9)
    low = rng.standard_normal((3, 4, 5, 4))     # clearsky_ratio + 3 wind features
    exo = {"topography": {"steps": [
        {"model": 0, "combine_type": "input", "data": rng.standard_normal((4, 5, 1))},
        {"model": 1, "combine_type": "layer", "data": rng.standard_normal((16, 20, 1))}]}}
    return low, exo


def run_chain(MultiStepGan, variant):
    low, exo = chain_inputs()
    models = chain_models(topo=variant == 3)
    msg = MultiStepGan(models)
    kw = [dict(), dict(norm_in=False), dict(un_norm_out=False), dict(exogenous_data=exo)][variant]
    out = msg.generate(low, **kw)
    return np.asarray(out, dtype=np.float64), [m.calls for m in models]


def run_solar(MultiStepGan, SolarMultiStepGan, variant):
    low, exo = solar_inputs()
    s_solar, s_wind, t_solar = solar_models(MultiStepGan, topo=variant == 1)
    model = SolarMultiStepGan(s_solar, s_wind, t_solar)
    kw = [dict(), dict(exogenous_data=exo, un_norm_out=False)][variant]
    out = model.generate(low, **kw)
    calls = [m.calls for grp in (s_solar, s_wind, t_solar) for m in grp.models]
    props = {"lr_features": list(model.lr_features), "hr_out_features": list(model.hr_out_features),
             "idf_wind": [int(i) for i in model.idf_wind],
             "idf_solar": [int(i) for i in model.idf_solar],
             "idf_wind_out": [int(i) for i in model.idf_wind_out]}
    return np.asarray(out, dtype=np.float64), calls, props


def main():
    MultiStepGan, SolarMultiStepGan, _ = load_reference_classes()
    res, meta = {}, {}
    for v in range(4):
        out, calls = run_chain(MultiStepGan, v)
        res[f"chain_{v}"] = out
        meta[f"chain_{v}_calls"] = calls
    for v in range(2):
        out, calls, props = run_solar(MultiStepGan, SolarMultiStepGan, v)
        res[f"solar_{v}"] = out
        meta[f"solar_{v}_calls"] = calls
        meta[f"solar_{v}_props"] = props
    # feature mismatch between steps raises (multi_step.py:177-186, wrapped :262-273)
    models = chain_models()
    models[1].lr_features = ["u_100m", "no_such_feature"]
    try:
        MultiStepGan(models).generate(chain_inputs()[0])
        meta["mismatch_raises"] = None
    except RuntimeError as e:
        meta["mismatch_raises"] = type(e.__cause__).__name__
    np.savez_compressed(OUT, meta=json.dumps(meta), **res)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()
