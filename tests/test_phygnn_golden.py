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


def test_exporter_tool_against_a_stand_in_sup3r(tmp_path, monkeypatch):
    """tools/export_phygnn_weights.py run end to end (``--config`` mode) against a stand-in
    ``sup3r.models.Sup3rGan`` that exposes what the tool touches of the real class (lazily built
    ``generator.weights`` with ``name / numpy() / assign() / shape``, ``generate``, ``is_5d``,
    feature lists): the directory it writes is what the slot tests above consume."""
    import importlib.util
    import sys
    import types

    class Var:
        def __init__(self, name, arr):
            self.name, self._a, self.shape = name, arr, arr.shape

        def numpy(self):
            return self._a

        def assign(self, v):
            assert v.shape == self._a.shape
            self._a = np.asarray(v, self._a.dtype)

    class Net:
        def __init__(self, fp):
            self.hl = json.load(open(fp))["hidden_layers"]
            self.layers, self.weights = None, []

        def run(self, x):
            if self.layers is None:
                self.layers = L.build_layers(self.hl)
                L.build_weights(self.layers, x.shape, seed=5)
                names = [f"generator/layer_{i // 2}/{'bias' if i % 2 else 'kernel'}:0"
                         for i in range(len(L.get_weights(self.layers)))]
                self.weights = [Var(n, w) for n, w in zip(names, L.get_weights(self.layers))]
            L.set_weights(self.layers, [w.numpy() for w in self.weights])
            return L.run_layers(self.layers, x.astype(np.float64)).astype(np.float32)

    class FakeGan:
        lr_features, hr_exo_features, obs_features, is_5d = [], [], [], True

        def __init__(self, fp_gen, fp_disc):
            self.generator, self.discriminator = Net(fp_gen), Net(fp_disc)

        @staticmethod
        def seed(s=0):
            pass

        def generate(self, x, norm_in=True, un_norm_out=True):
            assert not norm_in and not un_norm_out
            return self.generator.run(x)
    pkg, mod = types.ModuleType("sup3r"), types.ModuleType("sup3r.models")
    mod.Sup3rGan = FakeGan
    pkg.models = mod
    monkeypatch.setitem(sys.modules, "sup3r", pkg)
    monkeypatch.setitem(sys.modules, "sup3r.models", mod)
    fp_gen, fp_disc = str(tmp_path / "gen.json"), str(tmp_path / "disc.json")
    json.dump({"hidden_layers": C.spatiotemporal_generator(2, 2, (2,), n_blocks=1, filters=8)},
              open(fp_gen, "w"))
    json.dump({"hidden_layers": C.discriminator(3, "same", (8,))}, open(fp_disc, "w"))
    spec = importlib.util.spec_from_file_location(
        "export_phygnn_weights", os.path.join(os.path.dirname(HERE), "tools",
                                              "export_phygnn_weights.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    out = str(tmp_path / "export" / "random_st")
    tool.main(["--config", fp_gen, fp_disc, out])
    assert find_golden_dirs(str(tmp_path / "export")) == [out]
    assert not os.path.exists(os.path.join(out, "disc_weights.npz"))   # never built: not exported
    x, y = load_golden(out)
    assert x.shape == (1, 10, 10, 6, 2) and y.shape == (1, 20, 20, 12, 2)
    ws = _weights(out)
    assert any(w.ndim == 1 and np.abs(w).max() > 0 for w in ws)          # biases were randomised
    layers = L.build_layers(_hidden_layers(out))
    L.set_weights(layers, ws)
    got = L.run_layers(layers, x.astype(np.float64))
    assert np.abs(got - y).max() <= 1e-5 * np.abs(y).max()
