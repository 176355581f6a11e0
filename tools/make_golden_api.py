"""Golden record of the reference's public API surface on the hot path: for every
class this repo mirrors, the names, positional order and default values of the parameters of
every method / property defined in the reference class body (parsed with ``ast`` from
/root/reference; nothing is imported).

    python tools/make_golden_api.py   ->  tests/golden/api_surface.json
"""
import ast
import json
import os

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "api_surface.json")

# reference file -> classes mirrored here
CLASSES = {
    "sup3r/models/interface.py": ["AbstractInterface"],
    "sup3r/models/abstract.py": ["AbstractSingleModel"],
    "sup3r/models/base.py": ["Sup3rGan"],
    "sup3r/models/multi_step.py": ["MultiStepGan", "SolarMultiStepGan"],
    "sup3r/models/solar_cc.py": ["SolarCC"],
    "sup3r/models/dc.py": ["Sup3rGanDC"],
    "sup3r/models/with_obs.py": ["Sup3rGanWithObs"],
    "sup3r/pipeline/slicer.py": ["ForwardPassSlicer"],
    "sup3r/pipeline/strategy.py": ["ForwardPassChunk", "ForwardPassStrategy"],
    "sup3r/pipeline/forward_pass.py": ["ForwardPass"],
    "sup3r/preprocessing/data_handlers/exo.py": ["SingleExoDataStep", "ExoData"],
}


# reference file -> module-level functions mirrored here
FUNCTIONS = {
    "sup3r/bias/bias_transforms.py": ["global_linear_bc", "local_linear_bc",
                                      "monthly_local_linear_bc", "local_qdm_bc",
                                      "local_presrat_bc"],
    "sup3r/utilities/utilities.py": ["camel_to_underscore", "safe_cast", "temporal_coarsening",
                                     "spatial_coarsening"],
    "sup3r/pipeline/utilities.py": ["get_model", "get_chunk_slices"],
    "sup3r/models/utilities.py": ["get_optimizer_class"],
    "sup3r/preprocessing/utilities.py": ["make_time_index_from_kws"],
}


def _default(node):
    try:
        return repr(ast.literal_eval(node))
    except Exception:      # noqa: BLE001
        return ast.unparse(node)


def describe(fn):
    a = fn.args
    pos = [x.arg for x in a.posonlyargs + a.args]
    defaults = [None] * (len(pos) - len(a.defaults)) + [_default(d) for d in a.defaults]
    kind = "method"
    for d in fn.decorator_list:
        name = ast.unparse(d)
        if name in ("property", "cached_property", "functools.cached_property"):
            kind = "property"
        elif name in ("staticmethod", "classmethod"):
            kind = name
        elif name.endswith(".setter"):
            kind = "setter"
    return {"kind": kind, "params": [[p, d] for p, d in zip(pos, defaults)],
            "kwonly": [[x.arg, None if d is None else _default(d)]
                       for x, d in zip(a.kwonlyargs, a.kw_defaults)],
            "varargs": a.vararg is not None, "varkw": a.kwarg is not None}


def main():
    rec = {}
    for path, classes in CLASSES.items():
        tree = ast.parse(open(os.path.join(REF, path)).read())
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name in classes:
                members = {}
                for item in node.body:
                    if isinstance(item, ast.FunctionDef):
                        d = describe(item)
                        if d["kind"] == "setter":
                            continue
                        members[item.name] = d
                    elif isinstance(item, ast.AnnAssign) and isinstance(item.target, ast.Name):
                        # dataclass field
                        members[item.target.id] = {
                            "kind": "field",
                            "default": None if item.value is None else _default(item.value)}
                rec[node.name] = {"file": path, "members": members}
    funcs = {}
    for path, names in FUNCTIONS.items():
        tree = ast.parse(open(os.path.join(REF, path)).read())
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                funcs[node.name] = dict(describe(node), file=path)
    rec["__functions__"] = funcs
    json.dump(rec, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, {k: len(v.get("members", v)) for k, v in rec.items()})


if __name__ == "__main__":
    main()
