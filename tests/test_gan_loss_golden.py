"""GAN loss assembly and exo-layer inputs against a record made
from the REAL reference methods (tools/make_golden_gan_loss.py execs ``calc_loss``,
``calc_loss_disc``, ``calc_loss_gen_content``, ``get_loss_fun``, ``get_hr_exo_input``,
``_combine_loss_input``, ``_reshape_norm_exo``, ``run_exo_layer`` from sup3r/models with a
numpy-backed ``tf`` stub, the reference's own loss classes and ``ExoData``).  This repo's
``Sup3rGan`` methods run here on torch float64 CPU tensors: host logic only -- the device
kernels they call (channel crop / concat, pointwise mean losses, the discriminator loss, the
scalar scale) are replaced by their formulas; the kernels themselves are GPU tests."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

from oracle import torch_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "make_golden_gan_loss", os.path.join(ROOT, "tools", "make_golden_gan_loss.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)
G = json.load(open(os.path.join(ROOT, "tests", "golden", "gan_loss.json")))


class Tc:
    mean = staticmethod(lambda x, axis: x.mean(dim=axis))
    arange = staticmethod(lambda n, like: torch.arange(1, n + 1, dtype=like.dtype))


def _crop(x, crop):
    idx = tuple(slice(lo, x.shape[i] - hi) for i, (lo, hi) in enumerate(crop))
    return x[idx]


def _fn(f):
    return type("Fn", (), {"apply": staticmethod(f)})


class _MeanFormula:
    @staticmethod
    def apply(x1, x2, n_feat, kind):
        d = x1 - x2
        return (d * d).mean() if kind == 0 else d.abs().mean()


@pytest.fixture
def host_only(monkeypatch):
    from sup3r_b200 import autograd, loss_metrics, ops
    from sup3r_b200.models import base
    monkeypatch.setattr(loss_metrics, "ContentLossFn", _MeanFormula)
    monkeypatch.setattr(loss_metrics, "OrderProbe", T.order_probe(torch.mean), raising=False)
    monkeypatch.setattr(ops, "crop_fwd", _crop)
    monkeypatch.setattr(autograd, "CropFn", _fn(_crop))
    monkeypatch.setattr(autograd, "ConcatFn", _fn(lambda a, b: torch.cat([a, b], dim=-1)))
    monkeypatch.setattr(base, "DiscLossFn", _fn(
        lambda t, g: torch_ref.disc_loss(t.reshape(-1, 1), g.reshape(-1, 1))))
    monkeypatch.setattr(base, "ScaleFnScalar", _fn(lambda x, w: x * w))


def _scripted_class():
    from sup3r_b200.models import Sup3rGan

    class Scripted(Sup3rGan):
        hr_features = hr_exo_features = None
        _tf_discriminate = T.discriminate_for(Tc)

        def __init__(self):
            pass

    def make_obj(loss, hr_exo_features):
        obj = Scripted()
        obj.hr_features = T.HR_FEATURES[:2 + len(hr_exo_features)]
        obj.hr_exo_features = list(hr_exo_features)
        obj.loss_fun = Scripted.get_loss_fun(loss)
        obj._means, obj._stdevs = T.MEANS, T.STDEVS
        return obj
    return make_obj


def test_disc_loss_oracle_matches_reference():
    t, g = (torch.tensor(np.random.default_rng(s).standard_normal((5, 1))) for s in (60, 61))
    assert float(torch_ref.disc_loss(t, g)) == pytest.approx(G["loss"]["calc_loss_disc"],
                                                             rel=1e-12)
    assert float(torch_ref.disc_loss(g, t)) == pytest.approx(G["loss"]["calc_loss_disc_swapped"],
                                                             rel=1e-12)


def test_gan_loss_assembly_matches_reference(host_only):
    rec = T.loss_scenario(_scripted_class(),
                          convert=lambda a: torch.tensor(a, dtype=torch.float64))
    want_all = G["loss"]
    assert rec.keys() == want_all.keys()
    for k, want in want_all.items():
        got = rec[k]
        if isinstance(want, dict) and "details" in want:
            assert list(got["details"]) == list(want["details"]), k
            assert (got["loss"] is None) == (want["loss"] is None), k
            if want["loss"] is not None:
                assert got["loss"] == pytest.approx(want["loss"], rel=1e-11), k
            for d, v in want["details"].items():
                assert got["details"][d] == pytest.approx(v, rel=1e-11), (k, d)
        elif isinstance(want, dict):        # hr_exo_input: {feature: [shape, sum]}
            assert list(got) == list(want), k
            for f, (shape, total) in want.items():
                assert got[f][0] == shape and got[f][1] == pytest.approx(total, rel=1e-12), (k, f)
        elif isinstance(want, float):
            assert got == pytest.approx(want, rel=1e-11), k
        else:
            assert got == want, k


def test_exo_layer_inputs_match_reference(monkeypatch):
    from sup3r_b200 import network
    from sup3r_b200.exo import ExoData
    from sup3r_b200.models import abstract
    for mod in (network, abstract):
        monkeypatch.setattr(mod, "SUP3R_OBS_LAYERS", (T.ObsLayer,))
    monkeypatch.setattr(T.ExoLayer, "forward", T.ExoLayer.__call__, raising=False)

    def gather(obj, layer, x, exo, norm_in):
        arrays = obj._exo_for_layer(layer, x, exo, norm_in)
        _, _, hr_exo, extras = network.run_exo_layer(layer, torch.zeros(x.shape), arrays)
        return tuple(None if a is None else a.numpy() for a in (hr_exo, extras))
    rec = T.exo_scenario(_scripted_class(), ExoData, gather)
    want_all = G["exo"]
    assert rec.keys() == want_all.keys()
    for k, want in want_all.items():
        if isinstance(want, str):
            assert rec[k] == want, k
            continue
        for got_a, want_a in zip(rec[k], want):
            assert (got_a is None) == (want_a is None), k
            if want_a is None:
                continue
            assert got_a["shape"] == want_a["shape"], k
            tol = 2e-6 * want_a["abs_sum"]      # float32 exo tensors against the float64 record
            assert abs(got_a["sum"] - want_a["sum"]) <= tol, k
            assert got_a["abs_sum"] == pytest.approx(want_a["abs_sum"], rel=2e-6), k
            for e in ("first", "last"):
                assert got_a[e] == pytest.approx(want_a[e], rel=1e-5, abs=1e-6), (k, e)


def test_golden_is_reproducible_from_the_reference_when_present():
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    cls, exo_cls = T.load_reference()
    rec = {"loss": T.loss_scenario(T.reference_maker(cls)),
           "exo": T.exo_scenario(T.reference_maker(cls), exo_cls, T.reference_gather)}
    assert json.loads(json.dumps(rec)) == G
