"""``ForwardPassStrategy`` chunk bookkeeping (node splits, masks, output
file names, incremental restart) against records produced by the REAL reference methods
(tools/make_golden_strategy.py execs them from sup3r/pipeline/strategy.py).  CPU only."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "make_golden_strategy", os.path.join(ROOT, "tools", "make_golden_strategy.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)
G = json.load(open(os.path.join(ROOT, "tests", "golden", "strategy.json")))


def test_strategy_bookkeeping_matches_reference():
    from sup3r_b200.pipeline.strategy import ForwardPassStrategy
    got = T.scenario(ForwardPassStrategy)
    assert len(got) == len(G)
    for g, w in zip(got, G):
        assert g.keys() == w.keys()
        for k in w:
            assert g[k] == w[k], k


def test_time_slices_chunk_shape_and_features_match_reference():
    from sup3r_b200.pipeline.strategy import ForwardPassStrategy
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "strategy_slices.json")))
    got = T.slices_scenario(ForwardPassStrategy)
    assert got.keys() == want.keys()
    for k in want:
        assert len(got[k]) == len(want[k]), k
        for i, (g, w) in enumerate(zip(got[k], want[k])):
            assert g == w, (k, i)


def test_golden_is_reproducible_from_the_reference_when_present():
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    assert T.scenario(T.load_reference()) == G
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "strategy_slices.json")))
    assert json.loads(json.dumps(T.slices_scenario(T.load_reference_slices()))) == want
