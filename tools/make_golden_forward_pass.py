"""Golden vectors of the ``ForwardPass`` host functions around the generator call (SURVEY 8(a)
row a21) from the REAL reference methods: ``_get_step_enhance``, ``pad_source_data``,
``_reshape_data_chunk``, ``run_generator`` and ``_output_check``
(sup3r/pipeline/forward_pass.py:87-272, 384-425) are exec'd from their source text and bound to a
stand-in class; the model is a stand-in whose ``generate`` is a small deterministic numpy map.

    python tools/make_golden_forward_pass.py   ->  tests/golden/forward_pass.npz
"""
import copy
import json
import os
import textwrap
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "forward_pass.npz")


class Model:
    """Stand-in generator: nearest-neighbour enhancement + markers for the exo data it gets."""

    def __init__(self, ndim, s_enhance, t_enhance, n_steps=1):
        self.is_4d, self.is_5d, self.input_dims = ndim == 4, ndim == 5, ndim
        self.s_enhance, self.t_enhance = s_enhance, t_enhance
        self.s_enhancements = [s_enhance] + [1] * (n_steps - 1)
        self.t_enhancements = [1] * (n_steps - 1) + [t_enhance]
        if n_steps > 1:
            self.models = [self] * n_steps
        self.seen = []

    def generate(self, low_res, exogenous_data=None, **kwargs):
        x = np.asarray(low_res, dtype=np.float32)
        self.seen.append([list(x.shape), sorted(
            (k, i, list(np.shape(s["data"]))) for k, v in (exogenous_data or {}).items()
            for i, s in enumerate(v["steps"]))])
        y = np.repeat(np.repeat(x, self.s_enhance, axis=1), self.s_enhance, axis=2)
        if self.is_5d:
            y = np.repeat(y, self.t_enhance, axis=3)
        return y * 2 + 1


def grab(src, name):
    a = src.index(f"    def {name}(")
    start = src.rfind("\n", 0, src.rfind("\n", 0, a)) + 1     # include a decorator line
    if "@" not in src[start:a]:
        start = a
    b = a
    while True:
        b = src.find("\n    ", b + 1)
        if b < 0 or src[b + 5] not in (" ", "\n", ")"):
            break
    return textwrap.dedent(src[start:b if b > 0 else len(src)])


def load_reference():
    src = open(os.path.join(REF, "sup3r/pipeline/forward_pass.py")).read()
    body = "\n".join(textwrap.indent(grab(src, n), "    ") for n in (
        "_get_step_enhance", "pad_source_data", "run_generator", "_reshape_data_chunk",
        "_output_check"))
    ns = {"np": np, "logger": MagicMock(),
          "Timer": lambda: (lambda f, log=False, **k: f)}
    exec(compile("class RefForwardPass:\n" + body, "forward_pass.py", "exec"), ns)
    return ns["RefForwardPass"]


def scenario(FP):
    """-> (record, arrays) for a ForwardPass-like class ``FP`` (instances only need ``.model``)."""
    rng = np.random.default_rng(21)
    rec, arrs = {}, {}
    fp = FP.__new__(FP)
    fp.model = Model(5, 2, 3, n_steps=3)
    steps = [{"model": m, "combine_type": c} for m in range(3) for c in ("input", "layer", "output")]
    rec["step_enhance"] = [[int(v) for v in fp._get_step_enhance(s)] for s in steps]
    try:
        fp._get_step_enhance({"model": 0, "combine_type": "weird"})
        rec["weird_combine_type"] = "ok"
    except Exception as e:      # noqa: BLE001
        rec["weird_combine_type"] = type(e).__name__
    # pad_source_data: chunk + exo (a 3-D exo array gets its time axis from the chunk)
    data = rng.standard_normal((5, 6, 4, 2)).astype(np.float32)
    pad_width = ((2, 0), (1, 3), (0, 2))
    exo = {"topography": {"steps": [
        {"model": 0, "combine_type": "input", "data": rng.standard_normal((5, 6, 4, 1)),
         "s_enhance": 1, "t_enhance": 1},
        {"model": 0, "combine_type": "layer", "data": rng.standard_normal((10, 12, 1)),
         "s_enhance": 2, "t_enhance": 1},
        {"model": 2, "combine_type": "output", "data": rng.standard_normal((10, 12, 12, 1)),
         "s_enhance": 2, "t_enhance": 3}]}}
    for mode in ("reflect", "edge"):
        out, e = fp.pad_source_data(data.copy(), pad_width, copy.deepcopy(exo), mode=mode)
        arrs[f"pad_{mode}"] = np.asarray(out)
        for i, s in enumerate(e["topography"]["steps"]):
            arrs[f"pad_{mode}_exo{i}"] = np.asarray(s["data"])
    out, e = fp.pad_source_data(data.copy(), pad_width, None)
    rec["pad_no_exo"] = [list(out.shape), e is None]
    # run_generator: 5-D and 4-D models, crop slices, exo reshaping, enhancement check
    chunk = rng.standard_normal((4, 5, 6, 2)).astype(np.float32)
    crop = (slice(2, -2), slice(None), slice(3, 15), slice(None))
    m5 = Model(5, 2, 3)
    exo5 = {"topography": {"steps": [{"model": 0, "combine_type": "layer",
                                      "data": rng.standard_normal((8, 10, 18, 1))}]}}
    arrs["gen5"] = np.asarray(FP.run_generator(chunk.copy(), crop, m5, s_enhance=2, t_enhance=3,
                                                exo_data=copy.deepcopy(exo5)))
    rec["gen5_seen"] = m5.seen
    m4 = Model(4, 3, 1)
    exo4 = {"topography": {"steps": [{"model": 0, "combine_type": "input",
                                      "data": rng.standard_normal((4, 5, 6, 1))}]}}
    crop4 = (slice(None), slice(3, -3), slice(None), slice(None))
    arrs["gen4"] = np.asarray(FP.run_generator(chunk.copy(), crop4, m4, s_enhance=3, t_enhance=1,
                                                exo_data=copy.deepcopy(exo4)))
    rec["gen4_seen"] = m4.seen
    for name, kw in (("bad_s", dict(s_enhance=4, t_enhance=3)), ("bad_t", dict(s_enhance=2, t_enhance=2))):
        try:
            FP.run_generator(chunk.copy(), crop, Model(5, 2, 3), **kw)
            rec[name] = "ok"
        except Exception as e:      # noqa: BLE001
            rec[name] = type(e).__name__
    bad_exo = {"topography": {"steps": [{"model": 1, "combine_type": "input",
                                         "data": rng.standard_normal((4, 5, 6, 1))}]}}
    try:
        FP.run_generator(chunk.copy(), crop, Model(5, 2, 3), exo_data=bad_exo)
        rec["bad_exo_model"] = "ok"
    except Exception as e:      # noqa: BLE001
        rec["bad_exo_model"] = type(e).__name__

    class Failing(Model):
        def generate(self, *a, **k):
            raise ValueError("boom")
    try:
        FP.run_generator(chunk.copy(), crop, Failing(5, 2, 3))
        rec["failing_model"] = "ok"
    except Exception as e:      # noqa: BLE001
        rec["failing_model"] = [type(e).__name__, type(e.__cause__).__name__]
    # _output_check table
    good = rng.standard_normal((3, 4, 5, 3)).astype(np.float32)
    const = good.copy()
    const[..., 1] = 0.0
    nan = good.copy()
    nan[1, 2, 3, 0] = np.nan
    table = []
    for name, arr in (("good", good), ("const0", const), ("nan", nan)):
        for allowed in (True, False, 0, [0.0, 1.0], 5.0, (2,)):
            table.append([name, allowed if not isinstance(allowed, tuple) else list(allowed),
                          bool(FP._output_check(arr, allowed))])
    rec["output_check"] = table
    # a 3-D exo array of a step WITH temporal enhancement: its time axis is the chunk's length
    # times the step's own t_enhance (forward_pass.py:166-172) -- own generator: the arrays above
    # keep their values
    rng2 = np.random.default_rng(22)
    exo3 = {"sza": {"steps": [{"model": 0, "combine_type": "layer", "s_enhance": 2, "t_enhance": 3,
                               "data": rng2.standard_normal((10, 12, 1))}]}}
    out, e = fp.pad_source_data(data.copy(), pad_width, copy.deepcopy(exo3), mode="reflect")
    arrs["pad_exo3d_t3"] = np.asarray(e["sza"]["steps"][0]["data"])
    return rec, arrs


def main():
    rec, arrs = scenario(load_reference())
    np.savez_compressed(OUT, record=json.dumps(rec), **arrs)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    print(json.dumps(rec)[:1500])


if __name__ == "__main__":
    main()
