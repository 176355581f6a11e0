"""Golden records of the ``ForwardPassStrategy`` chunk bookkeeping (the
incremental-restart logic of SURVEY §5(d)) from the REAL reference methods: ``node_chunks``,
``unmasked_chunks``, ``get_chunk_indices``, ``out_files``, ``node_finished``, ``chunk_finished``
and ``chunk_masked`` (sup3r/pipeline/strategy.py:363-383, 438-472, 663-700) are exec'd from
their source text onto a stand-in object.

Second record: ``get_time_slices`` (with the reference's ``_parse_time_slice``),
``_get_fwp_chunk_shape`` and ``_init_features`` (strategy.py:306-333, 356-362, 385-391) over
time slices given as slice / list / tuple / None, with and without a temporal pad, default
(``None``) chunk-shape entries, exo features from ``exo_handler_kwargs``.

    python tools/make_golden_strategy.py   ->  tests/golden/strategy.json, strategy_slices.json
"""
import json
import os
import tempfile
import textwrap
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "strategy.json")

N_T, N_S = 3, 4
MASK = [0.0, 1.0, 0.0, 0.0]                     # spatial chunk 1 is fully masked
FINISHED = [0, 2, 3, 6, 11]                     # chunks whose output file already exists
CASES = [dict(incremental=True, max_nodes=None, redistribute_chunks=False),
         dict(incremental=True, max_nodes=5, redistribute_chunks=False),
         dict(incremental=True, max_nodes=2, redistribute_chunks=True),
         dict(incremental=False, max_nodes=4, redistribute_chunks=True),
         dict(incremental=True, max_nodes=None, redistribute_chunks=True, no_pattern=True)]


def grab_body(src, name):
    a = src.index(f"    def {name}(")
    b = a
    while True:
        b = src.find("\n    ", b + 1)
        if b < 0 or src[b + 5] not in (" ", "\n", ")"):
            break
    return textwrap.dedent(src[a:b if b > 0 else len(src)])


def load_reference():
    src = open(os.path.join(REF, "sup3r/pipeline/strategy.py")).read()
    ns = {"np": np, "os": os, "logger": MagicMock()}
    fns = {}
    for n in ("node_chunks", "unmasked_chunks", "get_chunk_indices", "out_files", "node_finished",
              "chunk_finished", "chunk_masked"):
        exec(compile(grab_body(src, n), n, "exec"), ns)
        fns[n] = ns[n]

    class RefStrategy:
        node_chunks = property(fns["node_chunks"])
        unmasked_chunks = property(fns["unmasked_chunks"])
        out_files = property(fns["out_files"])
        get_chunk_indices = fns["get_chunk_indices"]
        node_finished = fns["node_finished"]
        chunk_finished = fns["chunk_finished"]
        chunk_masked = fns["chunk_masked"]
    return RefStrategy


def slicer():
    return SimpleNamespace(
        n_time_chunks=N_T, n_spatial_chunks=N_S, n_chunks=N_T * N_S,
        get_chunk_indices=lambda i: (i % N_S, i // N_S))


def scenario(Strategy):
    rec = []
    for case in CASES:
        case = dict(case)
        with tempfile.TemporaryDirectory() as td:
            st = Strategy.__new__(Strategy)
            st.fwp_slicer = slicer()
            st.n_chunks = N_T * N_S
            st.__dict__["fwp_mask"] = np.array(MASK)
            st.out_pattern = None if case.pop("no_pattern", False) else \
                os.path.join(td, "chunks", "fwp_{file_id}.nc")
            st.incremental = case["incremental"]
            st.max_nodes = case["max_nodes"]
            st.redistribute_chunks = case["redistribute_chunks"]
            files = list(st.out_files)
            if st.out_pattern is not None:
                for i in FINISHED:
                    open(files[i], "w").close()
            out = {"out_files": [None if f is None else os.path.relpath(f, td) for f in files],
                   "dir_created": os.path.isdir(os.path.join(td, "chunks")),
                   "unmasked": [int(i) for i in st.unmasked_chunks],
                   "node_chunks": [[int(i) for i in c] for c in st.node_chunks],
                   "chunk_indices": [[int(v) for v in st.get_chunk_indices(i)]
                                     for i in range(N_T * N_S)],
                   "chunk_finished": [bool(st.chunk_finished(i)) for i in range(N_T * N_S)],
                   "chunk_masked": [bool(st.chunk_masked(i)) for i in range(N_T * N_S)]}
            out["node_finished"] = [bool(st.node_finished(n))
                                    for n in range(len(out["node_chunks"]))]
            rec.append(out)
    return rec


OUT_SLICES = os.path.join(ROOT, "tests", "golden", "strategy_slices.json")
TIME_SLICES = [None, slice(None), slice(0, 20), slice(5, 40, 2), [3, 30], (10, None), [None, 25],
               slice(4, None, 3), [0, 12, 1]]
CHUNK_SHAPES = [(4, 4, 10), (None, 6, None), (8, None, 5), (None, None, None)]


def load_reference_slices():
    src = open(os.path.join(REF, "sup3r/pipeline/strategy.py")).read()
    usrc = open(os.path.join(REF, "sup3r/preprocessing/utilities.py")).read()
    a = usrc.index("def _parse_time_slice")
    ns = {"np": np, "logger": MagicMock()}
    exec(compile(usrc[a:usrc.index("\ndef ", a + 1)], "utilities.py", "exec"), ns)
    body = {}
    for n in ("get_time_slices", "_get_fwp_chunk_shape", "_init_features"):
        exec(compile(grab_body(src, n), n, "exec"), ns)
        body[n] = ns[n]
    return type("RefStrategy", (), body)


def _sl(s):
    return [s.start, s.stop, s.step]


def slices_scenario(Strategy):
    rec = {"time_slices": [], "chunk_shapes": [], "features": []}
    for ts in TIME_SLICES:
        for pad in (0, 3):
            st = Strategy.__new__(Strategy)
            st.input_handler_kwargs = {} if ts is None else {"time_slice": ts}
            st.temporal_pad = pad
            unpadded, padded = st.get_time_slices()
            # (what the pair selects from a 60-step source, the statement that matters)
            src_steps = np.arange(60)[padded]
            rec["time_slices"].append({"unpadded": _sl(unpadded), "padded": _sl(padded),
                                       "source_steps": [int(v) for v in src_steps],
                                       "kept_steps": [int(v) for v in src_steps[unpadded]]})
    for shape in CHUNK_SHAPES:
        for tslice in (slice(None), slice(2, -2)):
            st = Strategy.__new__(Strategy)
            st.fwp_chunk_shape = shape
            st.time_slice = tslice
            st.input_handler = SimpleNamespace(grid_shape=(20, 30), time_index=np.arange(48))
            rec["chunk_shapes"].append([int(v) for v in st._get_fwp_chunk_shape()])
    model = SimpleNamespace(lr_features=["u_10m", "v_10m", "topography", "sza"])
    for exo_kwargs in (None, {}, {"topography": {"a": 1}}, {"sza": {}, "topography": {}}):
        st = Strategy.__new__(Strategy)
        st.exo_handler_kwargs = exo_kwargs
        st.exo_data = None
        feats, exo = st._init_features(model)
        rec["features"].append([list(feats), list(exo), st.exo_handler_kwargs])
    return rec


def main():
    rec = scenario(load_reference())
    json.dump(rec, open(OUT, "w"), indent=1)
    print("wrote", OUT)
    srec = slices_scenario(load_reference_slices())
    json.dump(srec, open(OUT_SLICES, "w"), indent=1)
    print("wrote", OUT_SLICES)
    for r in rec:
        print(r["node_chunks"], r["node_finished"])


if __name__ == "__main__":
    main()
