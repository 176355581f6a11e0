"""Genuine-TensorFlow parity slot: every export directory under
``tests/golden/phygnn/`` (written by ``tools/export_phygnn_weights.py`` in a real sup3r
environment) is checked against the oracle (CPU) and against the CUDA path (GPU).  With no
directory present the parametrised tests are skipped; the export FORMAT and the loader are
always exercised through a synthetic export written here."""
import json
import os

import numpy as np
import pytest

from oracle import layers_ref as L
from sup3r_b200 import configs as C
from sup3r_b200.interop import find_golden_dirs, load_golden

HERE = os.path.dirname(os.path.abspath(__file__))
DIRS = find_golden_dirs(os.path.join(HERE, "golden", "phygnn"))


def _hidden_layers(d):
    with open(os.path.join(d, "gen_hidden_layers.json")) as f:
        return json.load(f)["hidden_layers"]


def _weights(d):
    with np.load(os.path.join(d, "gen_weights.npz")) as z:
        return [z[k] for k in sorted(z.files)]


def write_synthetic_export(d, hl, shape, seed=0):
    """An export directory in the exporter's format whose golden output comes from the oracle."""
    layers = L.build_layers(hl)
    L.build_weights(layers, shape, seed=seed)
    ws = L.get_weights(layers)
    x = np.random.default_rng(42).standard_normal(shape).astype(np.float32)
    y = L.run_layers(layers, x.astype(np.float64)).astype(np.float32)
    os.makedirs(d, exist_ok=True)
    np.savez_compressed(os.path.join(d, "gen_weights.npz"),
                        **{f"w{i:03d}": w for i, w in enumerate(ws)})
    with open(os.path.join(d, "gen_hidden_layers.json"), "w") as f:
        json.dump({"hidden_layers": hl}, f)
    np.savez_compressed(os.path.join(d, "golden.npz"), low_res=x, hi_res=y)
    return x, y


@pytest.mark.skipif(not DIRS, reason="no TF / phygnn export under tests/golden/phygnn "
                                     "(see its README.md): conv parity stays unpinned")
@pytest.mark.parametrize("d", DIRS, ids=[os.path.basename(d) for d in DIRS])
def test_oracle_matches_tensorflow_golden(d):
    x, y = load_golden(d)
    layers = L.build_layers(_hidden_layers(d))
    L.set_weights(layers, _weights(d))
    got = L.run_layers(layers, x.astype(np.float64))
    assert got.shape == y.shape
    assert np.abs(got - y).max() <= 1e-5 * max(np.abs(y).max(), 1e-30)


@pytest.mark.gpu
@pytest.mark.skipif(not DIRS, reason="no TF / phygnn export under tests/golden/phygnn")
@pytest.mark.parametrize("d", DIRS, ids=[os.path.basename(d) for d in DIRS])
def test_cuda_path_matches_tensorflow_golden(cuda, d):
    from sup3r_b200.interop import load_exported_model
    x, y = load_golden(d)
    model = load_exported_model(d)
    for precision, tol in (("fp32", 1e-4), ("fp16c", 1e-3), ("bf16x3", 1e-3)):
        got = model.generate(x, norm_in=False, un_norm_out=False, precision=precision)
        err = np.abs(got - y).max() / max(np.abs(y).max(), 1e-30)
        assert got.shape == y.shape and err < tol, (precision, err)


def test_export_format_round_trip_oracle(tmp_path):
    """The exporter's file format -> oracle (no GPU): golden reproduced bit-for-bit."""
    d = str(tmp_path / "synthetic")
    hl = C.spatial_generator(2, (2,), n_blocks=1, filters=8)
    x, y = write_synthetic_export(d, hl, (2, 6, 6, 2))
    assert find_golden_dirs(str(tmp_path)) == [d]
    layers = L.build_layers(_hidden_layers(d))
    L.set_weights(layers, _weights(d))
    x2, y2 = load_golden(d)
    assert np.array_equal(x, x2)
    assert np.array_equal(L.run_layers(layers, x2.astype(np.float64)).astype(np.float32), y2)


@pytest.mark.gpu
def test_export_format_round_trip_cuda(cuda, tmp_path):
    """Loader -> model -> generate against the golden of a synthetic export."""
    from sup3r_b200.interop import load_exported_model
    d = str(tmp_path / "synthetic")
    hl = C.spatiotemporal_generator(2, 2, (2,), n_blocks=1)
    x, y = write_synthetic_export(d, hl, (1, 6, 6, 4, 2))
    model = load_exported_model(d)
    assert model.generator.built and len(model.generator.weights) == len(_weights(d))
    for precision, tol in (("fp32", 1e-4), ("fp16c", 1e-3)):
        got = model.generate(x, norm_in=False, un_norm_out=False, precision=precision)
        assert np.abs(got - y).max() / np.abs(y).max() < tol, precision
