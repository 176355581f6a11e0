"""Golden record of the GAN loss assembly from the REAL reference
methods: the source text of ``calc_loss``, ``calc_loss_disc``, ``calc_loss_gen_content``
(sup3r/models/base.py), ``get_loss_fun``, ``_get_loss_fun``, ``get_hr_exo_input``,
``_combine_loss_input``, ``_reshape_norm_exo`` and ``run_exo_layer`` (sup3r/models/abstract.py)
is exec'd from /root/reference with a numpy-backed ``tf`` stub (float64), the reference's own
``loss_metrics`` module (tools/make_golden_losses.py) and its own ``ExoData`` class, and bound
to a stand-in object whose discriminator is a small deterministic map that is sensitive to
every input value and channel.  Pinned: the relativistic average discriminator loss, which
tensor (with which exo channels appended) reaches the discriminator, the content loss on the
output channels only with the (generated, true) argument order, multi-term losses with
``term_weights`` and their snake-case detail names, the loss / detail keys of every flag
combination, exception types; the arrays an exo / observation layer receives (normalisation
incl. the ``_obs`` name rule, 3-D / 4-D -> 5-D tiling, ``features`` / ``exo_features`` stacking,
missing observation features).

    python tools/make_golden_gan_loss.py   ->  tests/golden/gan_loss.json
"""
import copy
import importlib.util
import json
import os
import re
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "gan_loss.json")


def _tool(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


LT = _tool("make_golden_losses")
TT = _tool("make_golden_training")

HR_FEATURES = ["u_10m", "v_10m", "topography", "sza"]


class Np:
    """Array operations of the stand-in discriminator on numpy (reference run)."""
    mean = staticmethod(lambda x, axis: np.mean(x, axis=axis))
    arange = staticmethod(lambda n, like: np.arange(1, n + 1, dtype=np.float64))


def discriminate_for(xp):
    """Stand-in ``_tf_discriminate``: (n, 1) logits, a different weight on every channel and on
    the first spatial row, so that channel order, the appended exo channels and the sample order
    all show in the loss."""
    def _tf_discriminate(self, hi_res):
        w = xp.arange(hi_res.shape[-1], hi_res) / 4.0
        axes = tuple(range(1, len(hi_res.shape) - 1))
        per_channel = xp.mean(hi_res, axes)                     # (n, c)
        first_row = xp.mean(hi_res[:, 0], tuple(range(1, len(hi_res.shape) - 2)))
        return ((per_channel * w).sum(-1) + 0.3 * (first_row * w).sum(-1)).reshape(-1, 1)
    return _tf_discriminate


def order_probe(mean):
    """A deliberately ASYMMETRIC loss class (every loss the reference ships is symmetric in its
    two arguments): put into the loss module on both sides, it shows which tensor is passed
    first (base.py:478-503: the generated one) and that constructor kwargs arrive."""
    class OrderProbe:
        def __init__(self, scale=1.0):
            self.scale = scale

        def __call__(self, x1, x2):
            return mean(x1) - self.scale * mean(x2 * x2)
    return OrderProbe


def load_reference():
    """Stand-in class carrying the reference's methods."""
    def bce(logits, labels):
        # tf.nn.sigmoid_cross_entropy_with_logits: max(x, 0) - x * z + log(1 + exp(-|x|))
        return np.maximum(logits, 0) - logits * labels + np.log1p(np.exp(-np.abs(logits)))
    tf = LT.tf_stub()
    tf.ones_like, tf.zeros_like = np.ones_like, np.zeros_like
    tf.nn = SimpleNamespace(
        sigmoid_cross_entropy_with_logits=lambda logits, labels: bce(logits, labels))
    tf.gather = lambda x, inds, axis: np.take(x, inds, axis=axis)
    tf.unstack = lambda x, axis: [np.take(x, i, axis=axis) for i in range(x.shape[axis])]
    losses_ns = LT.load_reference()
    sup3r = SimpleNamespace(utilities=SimpleNamespace(
        loss_metrics=SimpleNamespace(OrderProbe=order_probe(np.mean),
                                     **{k: v for k, v in losses_ns.items()
                                        if isinstance(v, type)})))
    ns = {"np": np, "tf": tf, "copy": copy, "re": re, "logger": MagicMock(), "sup3r": sup3r,
          "SUP3R_OBS_LAYERS": (ObsLayer,), "SUP3R_LAYERS": (ObsLayer, ExoLayer)}
    u_src = open(os.path.join(REF, "sup3r/utilities/utilities.py")).read()
    a = u_src.index("def camel_to_underscore")
    b = u_src.find("\ndef ", a + 1)
    exec(compile(u_src[a:b if b > 0 else len(u_src)], "utilities.py", "exec"), ns)
    a_src = open(os.path.join(REF, "sup3r/models/abstract.py")).read()
    b_src = open(os.path.join(REF, "sup3r/models/base.py")).read()
    m = {n: TT.grab_method(a_src, n, ns) for n in
         ("get_loss_fun", "_get_loss_fun", "get_hr_exo_input", "_combine_loss_input",
          "_reshape_norm_exo", "run_exo_layer")}
    m.update({n: TT.grab_method(b_src, n, ns) for n in
              ("calc_loss", "calc_loss_disc", "calc_loss_gen_content")})
    body = dict(m)
    body["get_loss_fun"] = classmethod(m["get_loss_fun"])
    body["_get_loss_fun"] = staticmethod(m["_get_loss_fun"])
    body["calc_loss_disc"] = staticmethod(m["calc_loss_disc"])
    body["_tf_discriminate"] = discriminate_for(Np)
    cls = type("RefGan", (), body)
    ms = _tool("make_golden_multistep")
    return cls, ms.load_reference_classes()[2]


class ExoLayer:
    """Stand-in Sup3rAdder / Sup3rConcat: records what it is called with."""
    def __init__(self, name, features=None, exo_features=None):
        self.name = name
        if features is not None:
            self.features = features
        if exo_features is not None:
            self.exo_features = exo_features

    def __call__(self, x, hr_exo, extras=None):
        return ("called", x, hr_exo, extras)


class ObsLayer(ExoLayer):
    """Stand-in observation layer (Sup3rObsModel / Sup3rConcatObs)."""


LOSSES = {
    "mae": "MeanAbsoluteError",
    "mse_dict": {"MeanSquaredError": {}},
    "two_terms": {"MeanAbsoluteError": {}, "SpatialExtremesLoss": {}, "term_weights": [0.7, 0.3]},
    "three_terms": {"MaterialDerivativeLoss": {}, "MmdLoss": {},
                    "LowResLoss": {"s_enhance": 2, "t_enhance": 2, "tf_loss": "MeanAbsoluteError"},
                    "term_weights": [0.2, 1.0, 3.0]},
    "no_weights": {"ExpLoss": {}, "TemporalExtremesLoss": {}},
    "order_probe": {"OrderProbe": {"scale": 2.0}, "MeanAbsoluteError": {},
                    "term_weights": [0.5, 1.0]},
}
FLAGS = [dict(train_gen=True, train_disc=False), dict(train_gen=True, train_disc=False,
                                                      compute_disc=True),
         dict(train_gen=False, train_disc=True), dict(train_gen=False, train_disc=False),
         dict(train_gen=False, train_disc=False, compute_disc=True)]


def tensors(n_true, n_gen, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((3, 4, 6, 4, n_true)),
            rng.standard_normal((3, 4, 6, 4, n_gen)) * 0.8 + 0.1)


def loss_scenario(make_obj, convert=lambda a: a, tofloat=float):
    """``make_obj(loss, hr_exo_features)`` -> object; returns the record."""
    rec = {}
    for lname, loss in LOSSES.items():
        for n_exo in (0, 2):
            obj = make_obj(loss, HR_FEATURES[2:2 + n_exo])
            true, gen = (convert(a) for a in tensors(2 + n_exo, 2, 50 + n_exo))
            for flags in FLAGS:
                out, details = obj.calc_loss(true, gen, weight_gen_advers=0.05, **flags)
                key = f"{lname}_exo{n_exo}_" + "".join(
                    str(int(flags.get(k, False))) for k in ("train_gen", "train_disc",
                                                            "compute_disc"))
                rec[key] = {"loss": None if out is None else tofloat(out),
                            "details": {k: tofloat(v) for k, v in details.items()}}
    obj = make_obj("MeanAbsoluteError", [])
    d_true, d_gen = (convert(np.random.default_rng(s).standard_normal((5, 1))) for s in (60, 61))
    rec["calc_loss_disc"] = tofloat(obj.calc_loss_disc(d_true, d_gen))
    rec["calc_loss_disc_swapped"] = tofloat(obj.calc_loss_disc(disc_out_true=d_gen,
                                                               disc_out_gen=d_true))
    exo = obj.__class__.get_hr_exo_input(make_obj("MeanAbsoluteError", HR_FEATURES[2:]),
                                         convert(tensors(4, 2, 52)[0]))
    rec["hr_exo_input"] = {k: [[int(s) for s in v.shape], tofloat(v.sum())]
                           for k, v in exo.items()}
    for name, fn in (
            ("shape_mismatch", lambda: obj.calc_loss(convert(tensors(2, 2, 53)[0]),
                                                     convert(tensors(2, 2, 53)[1][:, :2]))),
            ("unknown_loss", lambda: make_obj("NoSuchLoss", [])),
            ("bad_kwargs", lambda: make_obj({"LowResLoss": {"nope": 1}}, []))):
        try:
            fn()
            rec[name] = "ok"
        except Exception as e:      # noqa: BLE001
            rec[name] = type(e).__name__
    return rec


MEANS = {"topography": 120.0, "u_10m": 1.5, "sza": 40.0}
STDEVS = {"topography": 30.0, "u_10m": 4.0, "sza": 20.0}


def exo_scenario(make_obj, exo_data_cls, gather):
    """What an exo / observation layer receives from ``run_exo_layer``.  ``gather(obj, layer,
    x, exo, norm_in)`` -> (hr_exo, extras) arrays (or None)."""
    rng = np.random.default_rng(70)

    def exo_of(**arrays):
        return exo_data_cls({k: {"steps": [{"model": 0, "combine_type": "layer", "data": v}]}
                             for k, v in arrays.items()})
    topo3, topo4, topo5 = (rng.standard_normal(s) * 30 + 120
                           for s in ((4, 6, 1), (2, 4, 6, 1), (2, 4, 6, 3, 1)))
    sza5 = rng.standard_normal((2, 4, 6, 3, 1)) * 20 + 40
    u_obs = rng.standard_normal((2, 4, 6, 3, 1)) * 4 + 1.5
    x5, x4 = np.zeros((2, 4, 6, 3, 8)), np.zeros((2, 4, 6, 8))
    cases = {
        "topo3_into_5d": (ExoLayer("topography"), x5, exo_of(topography=topo3), True),
        "topo4_into_5d": (ExoLayer("topography"), x5, exo_of(topography=topo4), True),
        "topo5_no_norm": (ExoLayer("topography"), x5, exo_of(topography=topo5), False),
        "topo3_into_4d": (ExoLayer("topography"), x4, exo_of(topography=topo3), True),
        "two_features": (ExoLayer("both", features=["topography", "sza"]), x5,
                         exo_of(topography=topo5, sza=sza5), True),
        "obs_with_extras": (ObsLayer("obs", features=["u_10m_obs"], exo_features=["topography"]),
                            x5, exo_of(u_10m_obs=u_obs, topography=topo4), True),
        "obs_missing": (ObsLayer("obs", features=["u_10m_obs"]), x5, exo_of(topography=topo4),
                        True),
        "obs_missing_with_extras": (ObsLayer("obs", features=["u_10m_obs"],
                                             exo_features=["topography"]), x5,
                                    exo_of(topography=topo4), True),
    }
    rec = {}
    for key, (layer, x, exo, norm_in) in cases.items():
        hr_exo, extras = gather(make_obj("MeanAbsoluteError", []), layer, x, exo, norm_in)
        rec[key] = [None if a is None else
                    {"shape": [int(s) for s in np.shape(a)], "sum": float(np.sum(a)),
                     "abs_sum": float(np.sum(np.abs(a))),
                     "first": float(np.ravel(a)[0]), "last": float(np.ravel(a)[-1])}
                    for a in (hr_exo, extras)]
    for key, (layer, x, exo) in {
            "missing_feature": (ExoLayer("sza"), x5, exo_of(topography=topo4)),
            "rank_mismatch": (ExoLayer("topography"), x4, exo_of(topography=topo5))}.items():
        try:
            gather(make_obj("MeanAbsoluteError", []), layer, x, exo, True)
            rec[key] = "ok"
        except Exception as e:      # noqa: BLE001
            rec[key] = type(e).__name__
    return rec


def reference_maker(cls):
    def make_obj(loss, hr_exo_features):
        obj = cls()
        obj.hr_features = HR_FEATURES[:2 + len(hr_exo_features)]
        obj.hr_exo_features = list(hr_exo_features)
        obj.loss_fun = cls.get_loss_fun(loss)
        obj._means, obj._stdevs = MEANS, STDEVS
        return obj
    return make_obj


def reference_gather(obj, layer, x, exo, norm_in):
    _, _, hr_exo, extras = obj.run_exo_layer(layer, x, exo, norm_in=norm_in)
    return hr_exo, extras


def main():
    cls, exo_cls = load_reference()
    rec = {"loss": loss_scenario(reference_maker(cls)),
           "exo": exo_scenario(reference_maker(cls), exo_cls, reference_gather)}
    json.dump(rec, open(OUT, "w"), indent=1)
    print("wrote", OUT)
    print({k: v["loss"] for k, v in rec["loss"].items() if isinstance(v, dict) and "loss" in v})
    print({k: v for k, v in rec["loss"].items() if not isinstance(v, dict)})
    print(rec["exo"])


if __name__ == "__main__":
    main()
