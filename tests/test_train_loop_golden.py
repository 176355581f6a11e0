"""Top-level training loop against a record made from the REAL reference
method (tools/make_golden_train_loop.py execs ``Sup3rGan.train`` of sup3r/models/base.py:624-828
with the reference's own adaptive-weight methods on a scripted stand-in).  This repo's
``Sup3rGan.train`` driven through the same stand-ins must make the same calls with the same
arguments in the same order: epoch numbering of fresh / continued runs, the ``extras`` of
``finish_epoch``, the adversarial weight from epoch to epoch, early break, ``stop()``."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "make_golden_train_loop", os.path.join(ROOT, "tools", "make_golden_train_loop.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)
G = json.load(open(os.path.join(ROOT, "tests", "golden", "train_loop.json")))


def _close(a, b, path=""):
    if isinstance(b, float) and isinstance(a, (int, float)) and not isinstance(a, bool):
        assert a == pytest.approx(b, rel=1e-12), path
    elif isinstance(b, dict):
        assert isinstance(a, dict) and list(a) == list(b), path
        for k in b:
            _close(a[k], b[k], f"{path}/{k}")
    elif isinstance(b, list):
        assert isinstance(a, list) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _close(x, y, f"{path}[{i}]")
    else:
        assert a == b, path


def test_train_loop_matches_reference():
    from sup3r_b200.models import Sup3rGan

    def make_obj(log, stop_at):
        body = T.stand_ins(log, stop_at)
        body["__init__"] = lambda self: None
        obj = type("Scripted", (Sup3rGan,), body)()
        obj._optimizer, obj._optimizer_disc = {"lr": 1e-4, "it": 10}, {"lr": 4e-4, "it": 20}
        return obj
    got = json.loads(json.dumps(T.scenario(make_obj)))
    assert got.keys() == G.keys()
    for name in G:
        _close(got[name], G[name], name)


def test_golden_is_reproducible_from_the_reference_when_present():
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    assert json.loads(json.dumps(T.scenario(T.make_reference_object))) == G
