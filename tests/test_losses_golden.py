"""Content losses against values produced by the REAL reference source:
tools/make_golden_losses.py execs sup3r/utilities/loss_metrics.py with a numpy-backed ``tf`` stub
(float64) on seeded inputs -- every class but ``PerceptualLoss``, every ``LowResLoss`` option,
``_derivative``, ``gaussian_kernel`` and the exception types.  Checked here: the numpy oracle
(``oracle/losses_ref.py``, the checker of the GPU tests in tests/test_losses.py) and this repo's
loss classes on torch float64 CPU tensors (host logic only: the pointwise mean kernel
``s3_content_loss`` is replaced by its formula; the kernel itself is a GPU test)."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

from oracle import losses_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "make_golden_losses", os.path.join(ROOT, "tools", "make_golden_losses.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)
G = json.load(open(os.path.join(ROOT, "tests", "golden", "losses.json")))


def _unit_rows(shape):
    p = T.proj_normal(shape)
    return p / np.sqrt(np.maximum((p * p).sum(-1, keepdims=True), 1e-12))


ORACLE = {
    "ExpLoss": R.exp_loss, "MmdLoss": R.mmd_loss,
    "MaterialDerivativeLoss": R.material_derivative_loss,
    "SpatialDerivativeLoss": R.spatial_derivative_loss,
    "TemporalDerivativeLoss": R.temporal_derivative_loss, "CoarseMseLoss": R.coarse_mse_loss,
    "SpatialExtremesLoss": R.spatial_extremes_loss,
    "TemporalExtremesLoss": R.temporal_extremes_loss, "SpatialFftLoss": R.spatial_fft_loss,
    "SpatiotemporalFftLoss": R.spatiotemporal_fft_loss, "LowResLoss": R.low_res_loss,
}


def test_oracle_reproduces_the_reference_losses():
    for key, name, ckw, kw, shape, seed in T.CASES:
        x1, x2 = T.inputs(shape, seed)
        if name == "SlicedWassersteinLoss":
            n_pts = int(np.prod(shape[1:-1]))
            got = R.sliced_wasserstein_loss(x1, x2, _unit_rows((ckw["n_projections"], n_pts)))
        else:
            got = ORACLE[name](x1, x2, **ckw, **kw)
        assert float(got) == pytest.approx(G["losses"][key], rel=1e-12, abs=1e-14), key
    for k, want in G["derivative"].items():
        if isinstance(want, str):
            continue
        nd, axis = k.split("_axis")
        x = np.random.default_rng(31 if nd == "4d" else 32).standard_normal(
            (2, 4, 5, 6) if nd == "4d" else (2, 5, 4, 3, 2))
        np.testing.assert_allclose(R.derivative(x, int(axis)), np.array(want), rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        R.derivative(np.zeros((2, 4, 5, 6)), 0)
    assert G["derivative"]["axis0"] == "ValueError"
    np.testing.assert_allclose(R.gaussian_kernel(*T.inputs((3, 2, 2, 2), 33), sigma=1.5),
                               np.array(G["gaussian_kernel"]), rtol=1e-13)


class _MeanFormula:
    """s3_content_loss's value (kind 0: mean squared, 1: mean absolute difference)."""
    @staticmethod
    def apply(x1, x2, n_feat, kind):
        d = x1 - x2
        return (d * d).mean() if kind == 0 else d.abs().mean()


def test_loss_classes_reproduce_the_reference_losses(monkeypatch):
    from sup3r_b200 import loss_metrics as LM
    monkeypatch.setattr(LM, "ContentLossFn", _MeanFormula)
    monkeypatch.setattr(
        LM.SlicedWassersteinLoss, "projections",
        lambda self, n_points, device: torch.tensor(_unit_rows((self._n_projections, n_points))))

    def t64(a):
        return torch.tensor(a, dtype=torch.float64)
    rec = T.scenario(LM.get_loss_class, convert=t64, tofloat=float)
    assert rec.keys() == G["losses"].keys()
    for k, want in G["losses"].items():
        if isinstance(want, str):
            assert rec[k] == want, k
        else:
            # (the FFT losses transform in complex64 like the reference, loss_metrics.py:412,
            #  459: single-precision agreement with the float64 record)
            rel = 1e-6 if k.endswith("_fft") else 1e-11
            assert rec[k] == pytest.approx(want, rel=rel, abs=1e-13), k
    der = T.derivative_record(LM._derivative, convert=t64, toarray=lambda t: t.numpy())
    assert der.keys() == G["derivative"].keys()
    for k, want in G["derivative"].items():
        if isinstance(want, str):
            assert der[k] == want, k
        else:
            np.testing.assert_allclose(np.array(der[k]), np.array(want), rtol=0, atol=1e-15)
    gk = LM.gaussian_kernel(*(t64(a) for a in T.inputs((3, 2, 2, 2), 33)), sigma=1.5)
    np.testing.assert_allclose(gk.numpy(), np.array(G["gaussian_kernel"]), rtol=1e-13)


def test_golden_is_reproducible_from_the_reference_when_present():
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    ns = T.load_reference()
    assert T.scenario(lambda n: ns[n]) == G["losses"]
    assert T.derivative_record(ns["_derivative"]) == G["derivative"]
