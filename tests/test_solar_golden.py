"""``SolarCC`` loss windows against a record produced by the REAL reference class
(tools/make_golden_solar.py execs sup3r/models/solar_cc.py with a numpy-backed ``tf`` stub):
which hours go to the discriminator / the point loss / the daily-mean loss, argument order, the
per-day averaging, detail keys, exception types, ``temporal_pad``.  The stand-in discriminator
and losses run on torch CPU tensors here: host logic only."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "make_golden_solar", os.path.join(ROOT, "tools", "make_golden_solar.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)
G = json.load(open(os.path.join(ROOT, "tests", "golden", "solar_cc.json")))


class TorchBackend:
    mean = staticmethod(lambda x, axis=None: x.mean() if axis is None else x.mean(dim=axis))
    abs = staticmethod(torch.abs)
    numel = staticmethod(lambda x: int(x.numel()))
    first = staticmethod(lambda x: x[:, 0, 0, 0, 0])


def test_solar_cc_loss_windows_match_reference(monkeypatch):
    from sup3r_b200.models import SolarCC, solar_cc
    # (the scalar scaling of the adversarial term is a CUDA kernel in the product path)
    monkeypatch.setattr(solar_cc, "ScaleFnScalar",
                        type("Scale", (), {"apply": staticmethod(lambda x, w: x * w)}))
    log = []
    disc, ldisc, lcont = T.stand_ins(TorchBackend, log)

    class Scripted(SolarCC):
        _tf_discriminate = disc
        calc_loss_gen_content = lcont
        calc_loss_disc = staticmethod(ldisc)

        def __init__(self):
            self._t_enhance = 8

        def _sample_gen_windows(self, t_len, n_days):
            return list(T.WINDOWS[n_days])
    rec = T.scenario(Scripted(), log, to_backend=lambda a: torch.tensor(a, dtype=torch.float64),
                     tofloat=float)
    assert rec.keys() == G.keys()
    for k, want in G.items():
        got = rec[k]
        if not isinstance(want, dict):
            assert got == want, k
            continue
        assert got["calls"] == want["calls"], k
        assert got["details"].keys() == want["details"].keys(), k
        assert got["loss"] == pytest.approx(want["loss"], rel=1e-12, abs=1e-14), k
        for d, v in want["details"].items():
            assert got["details"][d] == pytest.approx(v, rel=1e-12, abs=1e-14), (k, d)


def test_golden_is_reproducible_from_the_reference_when_present():
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    log = []
    assert T.scenario(T.load_reference(log), log) == G
