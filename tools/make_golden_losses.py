"""Golden values of the content losses from the REAL reference source:
sup3r/utilities/loss_metrics.py is exec'd as it stands with a numpy-backed ``tf`` stub (float64;
``tf.complex64`` widened to complex128 so that the record is a float64 statement of the
algorithm; ``tf.random.normal`` scripted from a seeded numpy generator; the keras
``MeanSquaredError / MeanAbsoluteError`` objects are the mean over every element, which is what
keras' ``sum_over_batch_size`` reduction gives for equal shapes).  Every loss class except
``PerceptualLoss`` (VGG16 ImageNet weights) is called on seeded inputs, incl. every
``LowResLoss`` option and the exception types.

    python tools/make_golden_losses.py   ->  tests/golden/losses.json
"""
import json
import os
import re
from types import SimpleNamespace

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "losses.json")

PROJ_SEED = 1234


def proj_normal(shape):
    """The scripted ``tf.random.normal`` draw of SlicedWassersteinLoss."""
    return np.random.default_rng(PROJ_SEED).standard_normal(shape)


class _KerasLoss:
    def __init__(self, *a, **k):
        pass


def _elementwise(fn):
    class _L(_KerasLoss):
        def __call__(self, y_true, y_pred):
            return np.mean(fn(np.asarray(y_pred) - np.asarray(y_true)))
    return _L


MeanSquaredError = _elementwise(np.square)
MeanAbsoluteError = _elementwise(np.abs)


def tf_stub():
    def cast(x, dtype):
        return np.asarray(x).astype(dtype)

    def l2_normalize(x, axis=-1):
        # tf.math.l2_normalize: x * rsqrt(max(sum(x^2), epsilon)), epsilon 1e-12
        return x / np.sqrt(np.maximum(np.sum(x * x, axis=axis, keepdims=True), 1e-12))
    math = SimpleNamespace(multiply=np.multiply, log=np.log, l2_normalize=l2_normalize,
                           reduce_sum=lambda x, axis=None: np.sum(x, axis=axis))
    losses = SimpleNamespace(Loss=_KerasLoss, MeanSquaredError=MeanSquaredError,
                             MeanAbsoluteError=MeanAbsoluteError)
    return SimpleNamespace(
        concat=lambda values, axis=0: np.concatenate(values, axis=axis),
        stack=lambda values, axis=0: np.stack(values, axis=axis),
        exp=np.exp, abs=np.abs, square=np.square, math=math,
        reduce_sum=lambda x, axis=None: np.sum(x, axis=axis),
        reduce_mean=lambda x, axis=None: np.mean(x, axis=axis),
        reduce_min=lambda x, axis=None: np.min(x, axis=axis),
        reduce_max=lambda x, axis=None: np.max(x, axis=axis),
        expand_dims=lambda x, axis: np.expand_dims(x, axis),
        reshape=lambda x, shape: np.reshape(x, shape),
        transpose=lambda x, perm=None: np.transpose(x, perm),
        sort=lambda x, axis=-1: np.sort(x, axis=axis),
        cast=cast, complex64=np.complex128, constant=lambda v, dtype=None: np.asarray(v, dtype),
        convert_to_tensor=np.asarray, Tensor=np.ndarray,
        signal=SimpleNamespace(fft2d=lambda x: np.fft.fftn(x, axes=(-2, -1)),
                               fft3d=lambda x: np.fft.fftn(x, axes=(-3, -2, -1))),
        random=SimpleNamespace(normal=proj_normal),
        keras=SimpleNamespace(losses=losses, Model=None))


def load_reference():
    """Namespace of the reference module: its source minus the tensorflow import lines."""
    src = open(os.path.join(REF, "sup3r/utilities/loss_metrics.py")).read()
    src = re.sub(r"^(import tensorflow.*|from tensorflow.*)$", "", src, flags=re.M)
    ns = {"tf": tf_stub(), "MeanSquaredError": MeanSquaredError,
          "MeanAbsoluteError": MeanAbsoluteError, "VGG16": None, "preprocess_input": None}
    exec(compile(src, "loss_metrics.py", "exec"), ns)
    return ns


def inputs(shape, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(shape), rng.standard_normal(shape) * 1.3 + 0.2


# (record key, class, constructor kwargs, call kwargs, input shape, seed)
CASES = [
    ("exp_5d", "ExpLoss", {}, {}, (3, 6, 5, 4, 2), 1),
    ("exp_4d", "ExpLoss", {}, {}, (3, 6, 5, 2), 2),
    ("mmd_5d", "MmdLoss", {}, {}, (4, 5, 4, 3, 2), 3),
    ("mmd_4d_sigma2", "MmdLoss", {}, {"sigma": 2.0}, (4, 5, 4, 3), 4),
    ("material_derivative_2f", "MaterialDerivativeLoss", {}, {}, (2, 7, 6, 5, 2), 5),
    ("material_derivative_4f", "MaterialDerivativeLoss", {}, {}, (2, 5, 6, 7, 4), 6),
    ("spatial_derivative", "SpatialDerivativeLoss", {}, {}, (2, 7, 6, 5, 3), 7),
    ("temporal_derivative", "TemporalDerivativeLoss", {}, {}, (2, 4, 6, 9, 3), 8),
    ("coarse_mse_5d", "CoarseMseLoss", {}, {}, (3, 6, 5, 4, 2), 9),
    ("coarse_mse_4d", "CoarseMseLoss", {}, {}, (3, 6, 5, 2), 10),
    ("spatial_extremes_5d", "SpatialExtremesLoss", {}, {}, (3, 6, 5, 4, 2), 11),
    ("spatial_extremes_4d", "SpatialExtremesLoss", {}, {}, (3, 6, 5, 2), 12),
    ("temporal_extremes", "TemporalExtremesLoss", {}, {}, (3, 6, 5, 8, 2), 13),
    ("spatial_fft", "SpatialFftLoss", {}, {}, (2, 8, 6, 3), 14),
    ("spatiotemporal_fft", "SpatiotemporalFftLoss", {}, {}, (2, 6, 4, 8, 2), 15),
    ("low_res_identity", "LowResLoss", {}, {}, (2, 6, 6, 8, 2), 16),
    ("low_res_s3_4d", "LowResLoss", {"s_enhance": 3}, {}, (2, 6, 9, 2), 17),
    ("low_res_s2_t4_average", "LowResLoss", {"s_enhance": 2, "t_enhance": 4}, {},
     (2, 6, 4, 8, 2), 18),
    ("low_res_s2_t4_subsample", "LowResLoss",
     {"s_enhance": 2, "t_enhance": 4, "t_method": "Subsample"}, {}, (2, 6, 4, 8, 2), 19),
    ("low_res_t2_mae", "LowResLoss", {"t_enhance": 2, "tf_loss": "MeanAbsoluteError"}, {},
     (2, 3, 4, 8, 2), 20),
    ("low_res_s2_spatial_extremes", "LowResLoss",
     {"s_enhance": 2, "ex_loss": "SpatialExtremesLoss"}, {}, (2, 6, 4, 8, 2), 21),
    ("low_res_s2_t2_temporal_extremes", "LowResLoss",
     {"s_enhance": 2, "t_enhance": 2, "ex_loss": "TemporalExtremesLoss"}, {}, (2, 6, 4, 8, 2), 22),
    ("sliced_wasserstein_5d", "SlicedWassersteinLoss", {"n_projections": 16}, {},
     (3, 4, 3, 5, 2), 23),
    ("sliced_wasserstein_4d", "SlicedWassersteinLoss", {"n_projections": 8}, {}, (3, 4, 5, 2), 24),
]

# (record key, class, constructor kwargs, shapes of x1 / x2): exception types
FAILURES = [
    ("material_derivative_4d_input", "MaterialDerivativeLoss", {}, (2, 5, 5, 2), (2, 5, 5, 2)),
    ("spatial_derivative_4d_input", "SpatialDerivativeLoss", {}, (2, 5, 5, 2), (2, 5, 5, 2)),
    ("spatial_derivative_3d_input", "SpatialDerivativeLoss", {}, (2, 5, 5), (2, 5, 5)),
    ("temporal_derivative_4d_input", "TemporalDerivativeLoss", {}, (2, 5, 5, 2), (2, 5, 5, 2)),
    ("low_res_shape_mismatch", "LowResLoss", {"s_enhance": 2}, (2, 4, 4, 2), (2, 4, 6, 2)),
    ("low_res_unknown_ex_loss", "LowResLoss", {"ex_loss": "NoSuchLoss"}, (2, 4, 4, 2), (2, 4, 4, 2)),
    ("sliced_wasserstein_3d_input", "SlicedWassersteinLoss", {}, (2, 4, 4), (2, 4, 4)),
]


def scenario(get_class, convert=lambda a: a, tofloat=float):
    """Run every case with ``get_class(name)``; ``convert`` maps the numpy inputs to the
    backend's tensors."""
    rec = {}
    for key, name, ckw, kw, shape, seed in CASES:
        x1, x2 = (convert(a) for a in inputs(shape, seed))
        rec[key] = tofloat(get_class(name)(**ckw)(x1, x2, **kw))
    for key, name, ckw, s1, s2 in FAILURES:
        try:
            get_class(name)(**ckw)(convert(np.ones(s1)), convert(np.zeros(s2)))
            rec[key] = "ok"
        except Exception as e:      # noqa: BLE001
            rec[key] = type(e).__name__
    return rec


def derivative_record(fn, convert=lambda a: a, toarray=np.asarray):
    """``_derivative`` itself (axes 1 - 3, 4-D and 5-D input, the bad axis)."""
    rec = {}
    x4 = np.random.default_rng(31).standard_normal((2, 4, 5, 6))
    x5 = np.random.default_rng(32).standard_normal((2, 5, 4, 3, 2))
    for name, x in (("4d", x4), ("5d", x5)):
        for axis in (1, 2, 3):
            rec[f"{name}_axis{axis}"] = toarray(fn(convert(x), axis=axis)).tolist()
    try:
        fn(convert(x4), axis=0)
        rec["axis0"] = "ok"
    except Exception as e:      # noqa: BLE001
        rec["axis0"] = type(e).__name__
    return rec


def main():
    ns = load_reference()
    rec = {"losses": scenario(lambda n: ns[n]),
           "derivative": derivative_record(ns["_derivative"]),
           "gaussian_kernel": np.asarray(ns["gaussian_kernel"](
               *inputs((3, 2, 2, 2), 33), sigma=1.5)).tolist()}
    json.dump(rec, open(OUT, "w"), indent=1)
    print("wrote", OUT)
    print(rec["losses"])


if __name__ == "__main__":
    main()
