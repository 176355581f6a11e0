"""Generate golden fixtures from the REAL reference code (pure-python / numpy parts only).

Run in the build container (needs /root/reference):  python tools/make_golden.py
TensorFlow / phygnn / rex / xarray / dask are not installed, so they are replaced by inert stub
modules; only reference code that never touches them is exercised:
  * sup3r.pipeline.slicer.ForwardPassSlicer            -> tests/golden/slicer.json
  * sup3r.pipeline.utilities.get_chunk_slices           (inside slicer.json)
  * sup3r.preprocessing.data_handlers.exo.ExoData       -> tests/golden/exodata.json
  * sup3r.utilities.utilities.{spatial,temporal}_coarsening, camel_to_underscore
                                                        -> tests/golden/utilities.npz
The fixtures are committed; the GPU box never reads /root/reference.
"""
import importlib
import importlib.util
import json
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


def stub(*names):
    for n in names:
        parts = n.split(".")
        for i in range(1, len(parts) + 1):
            k = ".".join(parts[:i])
            if k not in sys.modules:
                sys.modules[k] = _Stub(k)
                sys.modules[k].__path__ = []


def load_ref_module(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _i(v):
    return None if v is None else int(v)


def sl(s):
    if isinstance(s, slice):
        return ["slice", _i(s.start), _i(s.stop), _i(s.step)]
    if isinstance(s, (tuple, list)):
        return [sl(v) for v in s]
    if isinstance(s, (np.integer,)):
        return int(s)
    return s


def main():
    os.makedirs(OUT, exist_ok=True)
    stub("tensorflow", "tensorflow.keras", "tensorflow.keras.layers", "tensorflow.keras.losses",
         "tensorflow.keras.optimizers", "phygnn", "phygnn.layers", "phygnn.layers.custom_layers",
         "rex", "rex.utilities", "rex.utilities.bc_utils", "gaps", "xarray", "dask",
         "dask.array", "h5py", "netCDF4", "cftime", "pandas.api.extensions")
    # package shells so that relative imports resolve without running sup3r/__init__.py
    for pkg in ("sup3r", "sup3r.pipeline", "sup3r.preprocessing", "sup3r.utilities",
                "sup3r.preprocessing.data_handlers"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split("."))]
        sys.modules[pkg] = m
    sys.modules["sup3r.models"] = _Stub("sup3r.models")
    # sup3r.preprocessing.utilities pulls xarray etc. in; the slicer only needs two helpers
    pu = types.ModuleType("sup3r.preprocessing.utilities")

    def _parse_time_slice(value):
        return value if isinstance(value, slice) else slice(*value) \
            if isinstance(value, (tuple, list)) else slice(None)

    pu._parse_time_slice = _parse_time_slice
    pu.log_args = lambda f: f
    sys.modules["sup3r.preprocessing.utilities"] = pu
    load_ref_module("sup3r.pipeline.utilities", "sup3r/pipeline/utilities.py")
    slicer_mod = load_ref_module("sup3r.pipeline.slicer", "sup3r/pipeline/slicer.py")

    cases = [
        dict(coarse_shape=(20, 20), time_steps=100, s_enhance=3, t_enhance=4,
             time_slice=[None, None, None], temporal_pad=4, spatial_pad=2, chunk_shape=(8, 8, 32)),
        dict(coarse_shape=(8, 8), time_steps=20, s_enhance=2, t_enhance=1,
             time_slice=[None, None, None], temporal_pad=1, spatial_pad=1, chunk_shape=(4, 4, 150)),
        dict(coarse_shape=(17, 23), time_steps=61, s_enhance=5, t_enhance=12,
             time_slice=[3, 57, None], temporal_pad=6, spatial_pad=3, chunk_shape=(10, 10, 24),
             min_width=(4, 4, 4)),
        dict(coarse_shape=(9, 9), time_steps=30, s_enhance=2, t_enhance=2,
             time_slice=[None, None, None], temporal_pad=0, spatial_pad=0, chunk_shape=(8, 8, 7),
             min_width=(4, 4, 4)),
        dict(coarse_shape=(21, 11), time_steps=48, s_enhance=1, t_enhance=24,
             time_slice=[0, 48, 2], temporal_pad=2, spatial_pad=1, chunk_shape=(10, 10, 12),
             min_width=(4, 4, 4)),
    ]
    golden = []
    import warnings
    warnings.simplefilter("ignore")
    for kw in cases:
        k = dict(kw)
        k["time_slice"] = slice(*kw["time_slice"])
        s = slicer_mod.ForwardPassSlicer(**k)
        rec = {"kwargs": kw}
        for attr in ["s1_lr_slices", "s2_lr_slices", "t_lr_slices", "s1_lr_pad_slices",
                     "s2_lr_pad_slices", "t_lr_pad_slices", "s_lr_slices", "s_lr_pad_slices",
                     "s_hr_slices", "s1_hr_crop_slices", "s2_hr_crop_slices", "t_hr_crop_slices",
                     "t_lr_crop_slices", "s_lr_crop_slices", "hr_crop_slices", "extra_padding",
                     "n_chunks", "n_spatial_chunks", "n_time_chunks"]:
            rec[attr] = sl(getattr(s, attr))
        rec["chunk_lookup"] = s.chunk_lookup.tolist()
        rec["chunk_indices"] = [list(map(int, s.get_chunk_indices(i))) for i in range(s.n_chunks)]
        golden.append(rec)
    gcs = sys.modules["sup3r.pipeline.utilities"].get_chunk_slices
    chunk_cases = [(10, 3, [None, None, None]), (17, 5, [2, 15, None]), (48, 12, [0, 48, 2]),
                   (7, 10, [None, None, None])]
    json.dump({"slicer": golden,
               "get_chunk_slices": [{"args": c, "out": sl(gcs(c[0], c[1], slice(*c[2])))}
                                    for c in chunk_cases]},
              open(os.path.join(OUT, "slicer.json"), "w"), indent=1)

    # ---- ExoData: exec only the two classes from the source text (the module imports xarray)
    src = open(os.path.join(REF, "sup3r/preprocessing/data_handlers/exo.py")).read()
    start = src.index("class SingleExoDataStep")
    end = src.index("class ExoDataHandler")
    ns = {"np": np, "logger": MagicMock()}
    exec(compile(src[start:end], "exo.py", "exec"), ns)
    ExoData = ns["ExoData"]
    rng = np.random.default_rng(0)
    lr = rng.standard_normal((6, 8, 10, 1)).astype(np.float32)
    hr2 = rng.standard_normal((12, 16, 1)).astype(np.float32)
    hr3 = rng.standard_normal((12, 16, 20, 1)).astype(np.float32)
    steps = {"topography": {"steps": [
        {"model": 0, "combine_type": "input", "data": lr, "s_enhance": 1, "t_enhance": 1},
        {"model": 0, "combine_type": "layer", "data": hr2, "s_enhance": 2, "t_enhance": 2},
        {"model": 1, "combine_type": "input", "data": hr3, "s_enhance": 2, "t_enhance": 2}]},
        "sza": {"steps": [
            {"model": 1, "combine_type": "output", "data": hr3 * 2, "s_enhance": 2,
             "t_enhance": 2}]}}
    exo = ExoData(steps)
    out = {"model_step_0": sorted(exo.get_model_step_exo(0)),
           "model_step_1": sorted(exo.get_model_step_exo(1)),
           "n_steps_0_topo": len(exo.get_model_step_exo(0)["topography"]["steps"]),
           "layer_sum": float(exo.get_combine_type_data("topography", "layer").sum())}
    chunk = exo.get_chunk([slice(1, 4), slice(2, 6), slice(3, 8)])
    out["chunk_shapes"] = {f: [list(s["data"].shape) for s in chunk[f]["steps"]] for f in chunk}
    out["chunk_sums"] = {f: [float(s["data"].sum()) for s in chunk[f]["steps"]] for f in chunk}
    import copy
    parts = ExoData(copy.deepcopy(steps)).split([1])
    out["split"] = [{f: [s["model"] for s in p[f]["steps"]] for f in p} for p in parts]
    json.dump(out, open(os.path.join(OUT, "exodata.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(OUT, "exodata_inputs.npz"), lr=lr, hr2=hr2, hr3=hr3)

    # ---- utilities: coarsening + names (functions exec'd from source text)
    usrc = open(os.path.join(REF, "sup3r/utilities/utilities.py")).read()
    ns = {"np": np, "logger": MagicMock(), "re": __import__("re")}
    for fn in ("temporal_coarsening", "spatial_coarsening", "camel_to_underscore"):
        a = usrc.index(f"def {fn}(")
        b = usrc.find("\ndef ", a + 10)
        b = len(usrc) if b < 0 else b
        exec(compile(usrc[a:b], fn, "exec"), ns)
    x5 = rng.standard_normal((2, 8, 12, 12, 3)).astype(np.float32)
    x4 = rng.standard_normal((3, 8, 12, 2)).astype(np.float32)
    res = {"x5": x5, "x4": x4,
           "sc5": ns["spatial_coarsening"](x5, 4), "sc4": ns["spatial_coarsening"](x4, 2),
           "sc3": ns["spatial_coarsening"](x4[0], 2, obs_axis=False)}
    for m in ("subsample", "average", "total", "max", "min"):
        res[f"tc_{m}"] = ns["temporal_coarsening"](x5, 3, m)
    np.savez_compressed(os.path.join(OUT, "utilities.npz"), **res)
    names = ["MeanSquaredError", "MeanAbsoluteError", "LowResLoss", "SpatialExtremesLoss"]
    json.dump({n: ns["camel_to_underscore"](n) for n in names},
              open(os.path.join(OUT, "names.json"), "w"), indent=1)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
