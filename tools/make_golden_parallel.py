"""Golden record of the multi-GPU gradient step from the REAL
reference methods: the source text of ``run_gradient_descent``, ``_get_parallel_grad`` and
``_sum_parallel_grad`` (sup3r/models/abstract.py:785-914) is exec'd from /root/reference
(``tf.split`` on numpy arrays, the real ``ThreadPoolExecutor``) and bound to a stand-in object
whose ``get_single_grad`` is a deterministic function of exactly the shard it is given.  Pinned:
equal batch shards in rank order (also of the ``mask`` keyword), the SUM (not the mean) of the
shard gradients, ONE ``apply_gradients`` call with the summed gradients, the loss details of the
LAST shard, and the single-device path (``multi_gpu=False`` or one GPU).

    python tools/make_golden_parallel.py   ->  tests/golden/parallel_grad.json
"""
import importlib.util
import json
import os
import time
from concurrent.futures import ThreadPoolExecutor
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "parallel_grad.json")

WEIGHT_SHAPES = [(3, 2), (4,), (2, 2, 2)]


def _tool(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def batch(n=4):
    rng = np.random.default_rng(90)
    return (rng.standard_normal((n, 3, 3, 2)), rng.standard_normal((n, 6, 6, 2)),
            (rng.random((n, 6, 6, 1)) > 0.5).astype(np.float64))


def shard_gradient(low_res, hi_res_true, mask=None):
    """Stand-in for the tape: gradients (one array per weight) and loss details that depend on
    every value of the shard (numpy in, numpy out; float64)."""
    s, h = float(np.sum(low_res * low_res)), float(np.mean(hi_res_true))
    m = 0.0 if mask is None else float(np.sum(mask))
    grads = [np.arange(1, int(np.prod(shp)) + 1, dtype=np.float64).reshape(shp)
             * (s + (i + 1) * h + 0.01 * m) for i, shp in enumerate(WEIGHT_SHAPES)]
    details = {"loss_gen": s / low_res.shape[0], "loss_gen_content": h, "mask_sum": m,
               "n_obs": float(low_res.shape[0]), "first": float(np.ravel(low_res)[0])}
    return grads, details


class Optimizer:
    """Records what ``apply_gradients`` receives."""
    def __init__(self):
        self.applied = []

    def apply_gradients(self, grads_and_vars):
        self.applied.append([(np.asarray(g, dtype=np.float64).tolist(), v)
                             for g, v in grads_and_vars])


def load_reference():
    ns = {"np": np, "time": time, "logger": MagicMock(), "ThreadPoolExecutor": ThreadPoolExecutor,
          "tf": SimpleNamespace(split=lambda x, n, axis=0: np.split(np.asarray(x), n, axis=axis))}
    grab = _tool("make_golden_training").grab_method
    src = open(os.path.join(REF, "sup3r/models/abstract.py")).read()
    body = {n: grab(src, n, ns) for n in
            ("run_gradient_descent", "_get_parallel_grad", "_sum_parallel_grad")}

    def get_single_grad(self, low_res, hi_res_true, training_weights, device_name=None,
                        **calc_loss_kwargs):
        self.calls.append([device_name, [int(v) for v in low_res.shape],
                           sorted(k for k in calc_loss_kwargs if k != "mask")])
        return shard_gradient(low_res, hi_res_true, calc_loss_kwargs.get("mask"))
    body["get_single_grad"] = get_single_grad
    return type("RefModel", (), body)


CASES = {
    # name: (number of GPUs, multi_gpu flag, with mask)
    "two_gpus": (2, True, False),
    "two_gpus_mask": (2, True, True),
    "four_gpus": (4, True, False),
    "two_gpus_flag_off": (2, False, False),
    "one_gpu_flag_on": (1, True, False),
}


def scenario(cls):
    rec = {}
    weights = [f"w{i}" for i in range(len(WEIGHT_SHAPES))]
    for name, (n_gpus, multi_gpu, with_mask) in CASES.items():
        obj = cls()
        obj.gpu_list = [f"/gpu:{i}" for i in range(n_gpus)]
        obj.default_device = "/gpu:0"
        obj.calls = []
        obj.optimizer = opt = Optimizer()
        lr, hr, mask = batch()
        kw = {"weight_gen_advers": 0.01, "train_gen": True}
        if with_mask:
            kw["mask"] = mask
        details = obj.run_gradient_descent(lr, hr, weights, multi_gpu=multi_gpu, **kw)
        assert len(opt.applied) == 1
        rec[name] = {"calls": obj.calls, "details": {k: float(v) for k, v in details.items()},
                     "applied": [g for g, _ in opt.applied[0]],
                     "applied_to": [v for _, v in opt.applied[0]]}
    return rec


def main():
    rec = scenario(load_reference())
    json.dump(rec, open(OUT, "w"), indent=1)
    print("wrote", OUT)
    for k, v in rec.items():
        print(k, v["calls"], v["details"], v["applied"][1])


if __name__ == "__main__":
    main()
