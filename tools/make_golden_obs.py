"""Golden record of the ``Sup3rGanWithObs`` host logic from the REAL reference
class: ``Sup3rGanWithObs`` (sup3r/models/with_obs.py:15-291) is exec'd from its source with a
numpy-backed ``tf`` stub on top of a stand-in parent.  With the same seeded generator the random
observation masks (onshore / offshore composite over topography, per-sample fractions drawn
from ranges, time fractions) must come out identical, and so must the sparse observation
tensors handed to the exo layers and the observation terms added to the loss.

    python tools/make_golden_obs.py   ->  tests/golden/with_obs.npz
"""
import json
import os
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "with_obs.npz")
SEED = 1234

HR_OUT = ["u_10m", "v_10m", "temperature_2m"]
OBS = ["u_10m_obs", "temperature_2m_obs"]
HR_FEATURES = HR_OUT + ["topography"]             # hr_out + hr_exo (obs features map into hr_out)


def loss_fun_for(xp):
    def loss_obs_fun(a, b):
        if xp.numel(a) == 0:
            return xp.nan(), {}
        return xp.mean(xp.abs(a - b)) + 0.5 * xp.mean(a), {}
    return loss_obs_fun


class Np:
    numel = staticmethod(lambda a: int(np.size(a)))
    mean = staticmethod(np.mean)
    abs = staticmethod(np.abs)
    nan = staticmethod(lambda: np.float64("nan"))


def parent_methods(gen_offset):
    """Stand-ins for what ``super()`` provides: the plain exo input (topography) and the plain
    forward + loss."""
    def get_hr_exo_input(self, hi_res_true):
        return {"topography": hi_res_true[..., 3:4]}

    def _get_hr_exo_and_loss(self, low_res, hi_res_true, **kw):
        exo = self.get_hr_exo_input(hi_res_true)
        hi_res_gen = hi_res_true[..., :len(HR_OUT)] * 0.9 + gen_offset
        loss = 2.0
        details = {"loss_gen": 2.0, "loss_gen_content": 1.5, "loss_gen_advers": 0.5}
        return loss, details, hi_res_gen, exo
    return get_hr_exo_input, _get_hr_exo_and_loss


def load_reference():
    src = open(os.path.join(REF, "sup3r/models/with_obs.py")).read()
    tf = SimpleNamespace(
        function=lambda f=None, **k: f if f is not None else (lambda g: g),
        stack=lambda vals, axis=0: np.stack(vals, axis=axis),
        where=np.where, constant=lambda v, dtype=None: np.asarray(v, dtype=dtype),
        gather=lambda x, inds, axis=-1: np.take(x, inds, axis=axis),
        expand_dims=np.expand_dims,
        unstack=lambda x, axis=-1: [np.take(x, i, axis=axis) for i in range(x.shape[axis])],
        reduce_sum=np.sum, cast=lambda x, dt: np.asarray(x, dtype=dt),
        size=lambda x: np.size(x), float32=np.float32)
    get_exo, get_loss = parent_methods(0.05)

    class Parent:
        get_hr_exo_input = get_exo
        _get_hr_exo_and_loss = get_loss
    ns = {"np": np, "tf": tf, "logger": MagicMock(), "Sup3rGan": Parent,
          "RANDOM_GENERATOR": np.random.default_rng(SEED)}
    exec(compile(src[src.index("class Sup3rGanWithObs"):], "with_obs.py", "exec"), ns)
    return ns["Sup3rGanWithObs"], ns


def configure(obj, is_5d, weight, offshore=True):
    obj.hr_out_features, obj.obs_features, obj.hr_features = HR_OUT, OBS, HR_FEATURES
    obj.is_5d, obj.is_4d = is_5d, not is_5d
    obj.onshore_obs_frac = {"spatial": [0.2, 0.6], "time": 0.7}
    obj.offshore_obs_frac = {"spatial": 0.05} if offshore else {}
    obj.loss_obs_weight = weight
    return obj


def hi_res(is_5d):
    rng = np.random.default_rng(77)
    shape = (3, 6, 7, 5, 4) if is_5d else (3, 6, 7, 4)
    x = rng.standard_normal(shape)
    x[..., 3] = rng.uniform(-50, 200, shape[:-1])          # topography: some cells offshore
    return x


def scenario(make_obj, to_backend=lambda a: a, to_np=np.asarray, tofloat=float):
    """``make_obj(is_5d, weight, offshore)`` -> configured model with a freshly seeded RNG."""
    rec, arrs = {}, {}
    for tag, is_5d, weight, offshore in (("5d", True, 0.3, True), ("4d", False, 0.0, True),
                                         ("5d_onshore_only", True, 0.3, False)):
        obj = make_obj(is_5d, weight, offshore)
        true = to_backend(hi_res(is_5d))
        rec[f"{tag}_obs_training_inds"] = [int(i) for i in obj.obs_training_inds]
        exo = obj.get_hr_exo_input(true)
        rec[f"{tag}_exo_keys"] = sorted(exo)
        arrs[f"{tag}_mask"] = to_np(exo["mask"]).astype(bool)
        for f in OBS:
            arrs[f"{tag}_{f}"] = to_np(exo[f]).astype(np.float64)
        loss, details, gen, exo2 = obj._get_hr_exo_and_loss(None, true, train_gen=True)
        rec[f"{tag}_loss"] = tofloat(loss)
        rec[f"{tag}_details"] = {k: tofloat(v) for k, v in details.items()}
        arrs[f"{tag}_mask2"] = to_np(exo2["mask"]).astype(bool)
        loss, details, _, _ = obj._get_hr_exo_and_loss(None, true, train_gen=False)
        rec[f"{tag}_disc_step_keys"] = sorted(details)
    return rec, arrs


def main():
    cls, ns = load_reference()

    def make_obj(is_5d, weight, offshore):
        ns["RANDOM_GENERATOR"] = np.random.default_rng(SEED)
        cls._get_single_obs_mask.__globals__["RANDOM_GENERATOR"] = ns["RANDOM_GENERATOR"]
        obj = configure(cls.__new__(cls), is_5d, weight, offshore)
        obj.loss_obs_fun = loss_fun_for(Np)
        return obj
    rec, arrs = scenario(make_obj)
    np.savez_compressed(OUT, record=json.dumps(rec), **arrs)
    print("wrote", OUT)
    print({k: v for k, v in rec.items() if "details" not in k})
    print({k: (v.shape, float(np.mean(v)) if v.dtype == bool else int(np.isnan(v).sum()))
           for k, v in arrs.items()})


if __name__ == "__main__":
    main()
