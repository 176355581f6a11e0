"""``MultiStepGan`` / ``SolarMultiStepGan`` chain logic against arrays
produced by the REAL reference classes (tools/make_golden_multistep.py execs
sup3r/models/interface.py, multi_step.py and data_handlers/exo.py with tensorflow / phygnn
stubbed): the same stand-in step models are chained by this repo's classes and must give the
same arrays bit for bit -- transposes 4-D <-> 5-D, feature matching between steps, norm flag
routing, exo routing and splitting, the solar / wind split and concat, the temporal pad.
Host logic only: runs on CPU."""
import importlib.util
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "make_golden_multistep", os.path.join(ROOT, "tools", "make_golden_multistep.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)
G = np.load(os.path.join(ROOT, "tests", "golden", "multistep.npz"))
META = json.loads(str(G["meta"]))


@pytest.mark.parametrize("variant", range(4))
def test_multi_step_chain_matches_reference(variant):
    from sup3r_b200.models import MultiStepGan
    out, calls = T.run_chain(MultiStepGan, variant)
    assert calls == META[f"chain_{variant}_calls"]
    assert out.shape == G[f"chain_{variant}"].shape and np.array_equal(out, G[f"chain_{variant}"])


@pytest.mark.parametrize("variant", range(2))
def test_solar_multi_step_matches_reference(variant):
    from sup3r_b200.models import MultiStepGan, SolarMultiStepGan
    out, calls, props = T.run_solar(MultiStepGan, SolarMultiStepGan, variant)
    assert props == META[f"solar_{variant}_props"]
    assert calls == META[f"solar_{variant}_calls"]
    assert np.array_equal(out, G[f"solar_{variant}"])


def test_feature_mismatch_between_steps_raises_like_the_reference():
    from sup3r_b200.models import MultiStepGan
    models = T.chain_models()
    models[1].lr_features = ["u_100m", "no_such_feature"]
    with pytest.raises(RuntimeError) as e:
        MultiStepGan(models).generate(T.chain_inputs()[0])
    assert type(e.value.__cause__).__name__ == META["mismatch_raises"] == "ValueError"


def test_golden_is_reproducible_from_the_reference_when_present():
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    MultiStepGan, SolarMultiStepGan, _ = T.load_reference_classes()
    out, calls = T.run_chain(MultiStepGan, 3)
    assert np.array_equal(out, G["chain_3"]) and calls == META["chain_3_calls"]
