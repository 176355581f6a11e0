"""Drop-in boundary: every method, property and dataclass field the reference
defines on the classes this repo mirrors exists here with the same parameter names, positional
order and default values (extra trailing keywords are this repo's additions).  The reference's
surface is a committed record parsed with ``ast`` from its source (tools/make_golden_api.py ->
tests/golden/api_surface.json); nothing of the reference is imported."""
import dataclasses
import inspect
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "api_surface.json")))

# members of the reference that are out of this repo's scope, by class
OUT_OF_SCOPE = {
    # CLI command strings for SLURM / gaps job submission
    "ForwardPass": {"get_node_cmd"},
    # file-based ExoDataHandler construction (rex / xarray rasterisers): exo data is passed live
    "ForwardPassStrategy": {"get_exo_kwargs", "get_exo_cache_files", "load_exo_data"},
}
# dataclass field defaults that differ on purpose: {(class, field): (reference, here)}
DIFFERENT_DEFAULTS = {
    # the reference runs the forward pass on the CPU unless told otherwise; this repo has no CPU
    # compute path (use_cpu=True raises)
    ("ForwardPassStrategy", "use_cpu"): ("True", "False"),
}


def _classes():
    import sup3r_b200.exo as E
    import sup3r_b200.models as M
    import sup3r_b200.pipeline as P
    from sup3r_b200.models.abstract import AbstractSingleModel
    from sup3r_b200.models.interface import AbstractInterface
    from sup3r_b200.pipeline.slicer import ForwardPassSlicer
    from sup3r_b200.pipeline.strategy import ForwardPassChunk
    return {"AbstractInterface": AbstractInterface, "AbstractSingleModel": AbstractSingleModel,
            "Sup3rGan": M.Sup3rGan, "MultiStepGan": M.MultiStepGan,
            "SolarMultiStepGan": M.SolarMultiStepGan, "SolarCC": M.SolarCC,
            "Sup3rGanDC": M.Sup3rGanDC, "Sup3rGanWithObs": M.Sup3rGanWithObs,
            "ForwardPassSlicer": ForwardPassSlicer, "ForwardPassChunk": ForwardPassChunk,
            "ForwardPassStrategy": P.ForwardPassStrategy, "ForwardPass": P.ForwardPass,
            "SingleExoDataStep": E.SingleExoDataStep, "ExoData": E.ExoData}


FUNCTION_HOMES = {"sup3r/bias/bias_transforms.py": "sup3r_b200.bias",
                  "sup3r/utilities/utilities.py": "sup3r_b200.utilities",
                  "sup3r/pipeline/utilities.py": "sup3r_b200.pipeline.utilities",
                  "sup3r/models/utilities.py": "sup3r_b200.models.utilities",
                  "sup3r/preprocessing/utilities.py": "sup3r_b200.bias"}


def test_function_signatures_match_reference():
    import importlib
    problems = []
    for name, d in G["__functions__"].items():
        fn = getattr(importlib.import_module(FUNCTION_HOMES[d["file"]]), name)
        params = list(inspect.signature(fn).parameters.values())
        ref_names = [p for p, _ in d["params"]]
        if [p.name for p in params[:len(ref_names)]] != ref_names:
            problems.append(f"{name}: {ref_names} in the reference, {[p.name for p in params]}")
            continue
        for (pn, pdef), p in zip(d["params"], params):
            if pdef is not None and pdef != "slice(None)" and repr(p.default) != pdef:
                problems.append(f"{name}({pn}): default {pdef} in the reference, {p.default!r}")
        problems += [f"{name}({p.name}): extra REQUIRED parameter"
                     for p in params[len(ref_names):] if p.default is inspect.Parameter.empty]
    assert not problems, "\n  ".join(problems)


@pytest.mark.parametrize("cname", sorted(k for k in G if not k.startswith("__")))
def test_class_surface_matches_reference(cname):
    cls = _classes()[cname]
    fields = {f.name: f for f in dataclasses.fields(cls)} if dataclasses.is_dataclass(cls) else {}
    problems = []
    for name, d in G[cname]["members"].items():
        if name in OUT_OF_SCOPE.get(cname, ()):
            assert not hasattr(cls, name), f"{cname}.{name} exists: drop it from OUT_OF_SCOPE"
            continue
        if d["kind"] == "field":
            if name in fields:
                f = fields[name]
                if (cname, name) in DIFFERENT_DEFAULTS:
                    assert (d["default"], repr(f.default)) == DIFFERENT_DEFAULTS[(cname, name)]
                elif d["default"] is not None and f.default is not dataclasses.MISSING \
                        and repr(f.default) != d["default"]:
                    problems.append(f"field default {name}: reference {d['default']}, "
                                    f"here {f.default!r}")
            elif not hasattr(cls, name):
                problems.append(f"missing field {name}")
            continue
        if not hasattr(cls, name):
            problems.append(f"missing {d['kind']} {name}")
            continue
        attr = inspect.getattr_static(cls, name)
        if d["kind"] == "property":
            if not isinstance(attr, (property, type(None))) and not hasattr(attr, "__get__"):
                problems.append(f"{name}: a property in the reference")
            continue
        if isinstance(attr, property):
            problems.append(f"{name}: a method in the reference, a property here")
            continue
        if d["kind"] in ("staticmethod", "classmethod") and type(attr).__name__ != d["kind"]:
            problems.append(f"{name}: {d['kind']} in the reference, {type(attr).__name__} here")
        fn = attr.__func__ if isinstance(attr, (staticmethod, classmethod)) else attr
        params = [p for p in inspect.signature(fn).parameters.values()
                  if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        ref_names = [p for p, _ in d["params"]]
        if [p.name for p in params[:len(ref_names)]] != ref_names:
            problems.append(f"{name}: parameters {ref_names} in the reference, "
                            f"{[p.name for p in params]} here")
            continue
        for (pn, pdef), p in zip(d["params"], params):
            if pdef is None:
                continue        # (a default where the reference requires a value is compatible)
            if p.default is inspect.Parameter.empty:
                problems.append(f"{name}({pn}): default {pdef} in the reference, required here")
            elif repr(p.default) != pdef:
                problems.append(f"{name}({pn}): default {pdef} in the reference, "
                                f"{p.default!r} here")
        for p in params[len(ref_names):]:
            if p.default is inspect.Parameter.empty:
                problems.append(f"{name}({p.name}): extra REQUIRED parameter")
    assert not problems, f"{cname} ({G[cname]['file']}):\n  " + "\n  ".join(problems)


def test_golden_is_reproducible_from_the_reference_when_present(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "make_golden_api", os.path.join(ROOT, "tools", "make_golden_api.py"))
    T = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(T)
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    T.OUT = str(tmp_path / "api.json")
    T.main()
    assert json.load(open(T.OUT)) == G
