"""``ForwardPass`` host functions around the generator call against
vectors produced by the REAL reference methods (tools/make_golden_forward_pass.py execs
``_get_step_enhance / pad_source_data / _reshape_data_chunk / run_generator / _output_check`` of
sup3r/pipeline/forward_pass.py with a stand-in model): same arrays, same exo routing, same
exception types, same output-check table.  CPU only."""
import importlib.util
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "make_golden_forward_pass", os.path.join(ROOT, "tools", "make_golden_forward_pass.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)
G = np.load(os.path.join(ROOT, "tests", "golden", "forward_pass.npz"))
REC = json.loads(str(G["record"]))


def test_forward_pass_host_functions_match_reference():
    from sup3r_b200.pipeline.forward_pass import ForwardPass
    rec, arrs = T.scenario(ForwardPass)
    rec = json.loads(json.dumps(rec))
    assert rec.keys() == REC.keys()
    for k in REC:
        assert rec[k] == REC[k], k
    assert set(arrs) == set(G.files) - {"record"}
    for k, a in arrs.items():
        assert a.shape == G[k].shape and a.dtype == G[k].dtype and np.array_equal(a, G[k]), k


def test_golden_is_reproducible_from_the_reference_when_present():
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    rec, arrs = T.scenario(T.load_reference())
    assert json.loads(json.dumps(rec)) == REC
    assert all(np.array_equal(a, G[k]) for k, a in arrs.items())
