"""``Sup3rGanWithObs`` host logic against arrays produced by the REAL reference
class (tools/make_golden_obs.py execs sup3r/models/with_obs.py with a numpy-backed ``tf`` stub
and a stand-in parent): with the same seeded generator the random observation masks, the sparse
observation tensors for the exo layers (NaN where unobserved) and the observation loss terms
must be identical.  torch CPU tensors here: host logic only."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "make_golden_obs", os.path.join(ROOT, "tools", "make_golden_obs.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)
G = np.load(os.path.join(ROOT, "tests", "golden", "with_obs.npz"))
REC = json.loads(str(G["record"]))


class Tc:
    numel = staticmethod(lambda a: int(a.numel()))
    mean = staticmethod(torch.mean)
    abs = staticmethod(torch.abs)
    nan = staticmethod(lambda: torch.tensor(float("nan"), dtype=torch.float64))


def test_obs_model_matches_reference(monkeypatch):
    from sup3r_b200.models import Sup3rGanWithObs, with_obs
    get_exo, get_loss = T.parent_methods(0.05)
    for name, fn in (("get_hr_exo_input", get_exo), ("_get_hr_exo_and_loss", get_loss)):
        owner = next(c for c in Sup3rGanWithObs.__mro__[1:] if name in c.__dict__)
        monkeypatch.setattr(owner, name, fn)

    class Scripted(Sup3rGanWithObs):
        hr_out_features = obs_features = hr_features = is_5d = is_4d = None

        def __init__(self):
            pass

    def make_obj(is_5d, weight, offshore):
        monkeypatch.setattr(with_obs, "RANDOM_GENERATOR", np.random.default_rng(T.SEED))
        obj = T.configure(Scripted(), is_5d, weight, offshore)
        obj.loss_obs_fun = T.loss_fun_for(Tc)
        return obj
    rec, arrs = T.scenario(make_obj, to_backend=lambda a: torch.tensor(a, dtype=torch.float64),
                           to_np=lambda t: t.detach().numpy() if isinstance(t, torch.Tensor)
                           else np.asarray(t))
    assert rec.keys() == REC.keys()
    for k, want in REC.items():
        if isinstance(want, float):
            assert rec[k] == pytest.approx(want, rel=1e-12), k
        elif isinstance(want, dict):
            assert rec[k].keys() == want.keys(), k
            for d, v in want.items():
                assert rec[k][d] == pytest.approx(v, rel=1e-12, nan_ok=True), (k, d)
        else:
            assert rec[k] == want, k
    for k in G.files:
        if k != "record":
            assert arrs[k].shape == G[k].shape, k
            assert np.array_equal(arrs[k], G[k], equal_nan=G[k].dtype != bool), k


def test_golden_is_reproducible_from_the_reference_when_present():
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    cls, ns = T.load_reference()

    def make_obj(is_5d, weight, offshore):
        cls._get_single_obs_mask.__globals__["RANDOM_GENERATOR"] = np.random.default_rng(T.SEED)
        obj = T.configure(cls.__new__(cls), is_5d, weight, offshore)
        obj.loss_obs_fun = T.loss_fun_for(T.Np)
        return obj
    rec, arrs = T.scenario(make_obj)
    assert json.loads(json.dumps(rec)) == REC
    assert all(np.array_equal(arrs[k], G[k], equal_nan=G[k].dtype != bool)
               for k in G.files if k != "record")
