"""Golden records of the GAN training schedule and history bookkeeping (SURVEY 8(a) rows a13,
a14) from the REAL reference methods: the source text of ``update_loss_details``, ``early_stop``
(sup3r/models/abstract.py), ``get_weight_update_fraction``, ``update_adversarial_weights``,
``_train_batch``, ``_post_batch`` and ``_train_epoch`` (sup3r/models/base.py) is exec'd from
/root/reference and bound to a stand-in object whose ``run_gradient_descent`` returns scripted
loss values (the modules themselves import tensorflow / phygnn, which are not installed).

    python tools/make_golden_training.py   ->  tests/golden/training_schedule.json
"""
import json
import os
import textwrap
import time
from unittest.mock import MagicMock

import numpy as np
import pandas as pd

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "training_schedule.json")


def grab_method(src, name, ns):
    """exec one method of a class body (found by its 4-space indentation)."""
    a = src.index(f"    def {name}(")
    deco = src.rfind("\n", 0, a - 1)
    if src[deco:a].strip() == "@staticmethod":
        pass
    b = a
    while True:
        b = src.find("\n    ", b + 1)
        nxt = src[b + 5:b + 9]
        if b < 0 or (src[b + 5] != " " and src[b + 5] != "\n" and src[b + 5] != ")"):
            break
    body = textwrap.dedent(src[a:b if b > 0 else len(src)])
    exec(compile(body, name, "exec"), ns)
    return ns[name]


# the scripted losses: what run_gradient_descent reports for call number n
def scripted_losses(n, train_gen, train_disc, compute_disc):
    out = {}
    if train_gen:
        out["loss_gen"] = 1.0 / (1 + 0.1 * n)
        out["loss_gen_content"] = 0.9 / (1 + 0.1 * n)
        out["loss_gen_advers"] = 0.7 + 0.01 * n
    if train_disc or compute_disc:
        out["loss_disc"] = 0.72 - 0.045 * (n % 11) + 0.02 * (n // 11)
    return out


SCENARIOS = [
    # (train_gen, train_disc, disc_loss_bounds, n_batches) per epoch
    [(True, True, (0.45, 0.6), 6), (True, True, (0.45, 0.6), 6), (True, True, (0.3, 0.5), 5)],
    [(True, False, (0.45, 0.6), 4), (False, True, (0.45, 0.6), 4), (True, True, (0.6, 0.7), 7)],
    # the running discriminator loss EQUAL to the lower bound (the 0 a fresh record starts from):
    # "too good" includes equality
    [(True, True, (0.0, 0.6), 3), (True, True, (0.5, 0.9), 4)],
]


class Batch:
    def __init__(self, i):
        self.low_res, self.high_res = ("lr", i), ("hr", i)


class Handler:
    shapes = ((2, 4, 4, 4, 2), (2, 8, 8, 8, 2))
    lr_shape, hr_shape = (4, 4, 4, 2), (8, 8, 8, 2)

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __iter__(self):
        return iter([Batch(i) for i in range(self.n)])


def run_scenario(obj, scenario):
    """Drive ``obj._train_epoch`` through the scenario; returns the golden record."""
    rec = {"calls": [], "epochs": []}
    counter = [0]

    def run_gradient_descent(low_res, hi_res, weights, weight_gen_advers=None, optimizer=None,
                             train_gen=True, train_disc=False, compute_disc=False,
                             multi_gpu=False):
        counter[0] += 1
        rec["calls"].append([weights, optimizer, bool(train_gen), bool(train_disc),
                             bool(compute_disc), low_res[1]])
        return scripted_losses(counter[0], train_gen, train_disc, compute_disc)
    obj.run_gradient_descent = run_gradient_descent
    for train_gen, train_disc, bounds, n in scenario:
        details = obj._train_epoch(Handler(n), 1e-3, train_gen, train_disc, bounds)
        rec["epochs"].append({k: float(v) for k, v in details.items()})
    tr = obj._train_record
    rec["record_columns"] = list(tr.columns)
    rec["record_index"] = [int(i) for i in tr.index]
    rec["record"] = [[None if pd.isna(v) else float(v) for v in row] for row in tr.values]
    return rec


def make_reference_object():
    asrc = open(os.path.join(REF, "sup3r/models/abstract.py")).read()
    bsrc = open(os.path.join(REF, "sup3r/models/base.py")).read()
    ns = {"np": np, "pd": pd, "time": time, "logger": MagicMock(), "warn": lambda *a, **k: None,
          "tf": MagicMock(), "numpy_if_tensor": lambda v: v}

    class Ref:
        pass
    for name in ("update_loss_details", "early_stop"):
        setattr(Ref, name, staticmethod(grab_method(asrc, name, ns)))
    setattr(Ref, "get_weight_update_fraction",
            staticmethod(grab_method(bsrc, "get_weight_update_fraction", ns)))
    for name in ("update_adversarial_weights", "_train_batch", "_post_batch", "_train_epoch"):
        setattr(Ref, name, grab_method(bsrc, name, ns))
    return Ref


def prime(obj):
    obj._train_record = pd.DataFrame()
    obj._tb_writer = None
    obj._write_tb_profile = False
    obj.total_batches = 0
    obj.generator_weights, obj.discriminator_weights = "gen_weights", "disc_weights"
    obj.timer = lambda f, log=False, **k: f
    obj.init_weights = lambda *a, **k: None
    obj.profile_to_tensorboard = lambda *a, **k: None
    return obj


# ---------------------------------------------------------------- finish_epoch (row a14)
def finish_epoch_scenario(obj):
    """History rows, checkpoint triggers and early stop over 9 epochs; -> record."""
    saves = []
    obj.save = lambda out_dir: saves.append(out_dir)
    obj._history = pd.DataFrame(columns=["elapsed_time"])
    obj._history.index.name = "epoch"
    epochs = list(range(3, 12))
    vals = [1.0, 0.8, 0.7, 0.699, 0.6985, 0.6981, 0.698, 0.6979, 0.6979]
    stops = []
    for e, v in zip(epochs, vals):
        details = {"train_loss_gen": v * 1.1, "val_loss_gen": v, "train_loss_disc": 0.5 + 0.01 * e}
        extras = {"weight_gen_advers": 1e-3 * e, "OptmGen/learning_rate": np.float32(1e-4),
                  "flag": True} if e % 2 else None
        stops.append(bool(obj.finish_epoch(e, epochs, time.time(), details, 4, "ckpt_{epoch}",
                                           "val_loss_gen", 0.002, 3, extras=extras)))
    h = obj._history
    rec = {"saves": saves, "stops": stops, "columns": list(h.columns),
           "index": [int(i) for i in h.index], "index_name": h.index.name,
           "elapsed_positive": bool((h["elapsed_time"] >= 0).all()),
           "values": [[None if pd.isna(v) else float(v) for v in row]
                      for row in h.drop(columns=["elapsed_time"]).values]}
    try:
        obj.finish_epoch(12, [12], time.time(), {"val_loss_gen": 0.5}, None, "no_key", None,
                         0.002, 3)
        rec["bad_out_dir"] = "ok"
    except Exception as e:      # noqa: BLE001
        rec["bad_out_dir"] = type(e).__name__
    return rec


def make_reference_finish_object():
    asrc = open(os.path.join(REF, "sup3r/models/abstract.py")).read()
    from types import SimpleNamespace
    ns = {"np": np, "pd": pd, "time": time, "logger": MagicMock(),
          "tf": SimpleNamespace(Tensor=type("Tensor", (), {}))}
    usrc = open(os.path.join(REF, "sup3r/utilities/utilities.py")).read()
    a = usrc.index("def safe_cast(")
    exec(compile(usrc[a:usrc.index("\ndef ", a + 5)], "safe_cast", "exec"), ns)

    class RefFinish:
        pass
    for name in ("early_stop", "log_loss_details"):
        setattr(RefFinish, name, staticmethod(grab_method(asrc, name, ns)))
    RefFinish.finish_epoch = grab_method(asrc, "finish_epoch", ns)
    return RefFinish()


# ---------------------------------------------------------------- persistence (rows a12, a16)
class StubVar:
    def __init__(self, name, arr):
        self.name, self._arr = name, np.asarray(arr)

    def numpy(self):
        return self._arr


class StubOptimizer:
    variables = [StubVar("Adam/iteration:0", 7), StubVar("Adam/m/conv/kernel:0", [[-1.0, 3.0]]),
                 StubVar("Adam/v/conv/kernel:0", [0.5, 0.25, 0.0])]

    def get_config(self):
        return {"name": "Adam", "learning_rate": np.float32(1e-4), "beta_1": np.float64(0.9),
                "steps": np.int64(3), "amsgrad": False, "epsilon": 1e-7}


class StubHandler:
    smoothing = 0.5
    lr_features = ["u", "v"]
    hr_out_features = ["u"]
    unrelated = 1


PARAMS = {"name": "gan", "loss": {"MeanAbsoluteError": {}}, "learning_rate": np.float32(1e-4),
          "means": {"u": np.float32(1.5), "v": 2}, "stdevs": {"u": 0.5, "v": np.float64(3)},
          "meta": {"s_enhance": np.int64(3), "lr_features": ("u", "v"), "arr": np.arange(3),
                   "class": "Sup3rGan", "obj": complex(1, 2)},
          "version_record": {"sup3r": "0.2"}, "default_device": "/gpu:0"}


def persistence_scenario(cls, td):
    """save_params / load_saved_params / optimiser config + state / batch handler attributes."""
    rec = {}
    obj = cls.__new__(cls)
    type(obj).model_params = property(lambda self: PARAMS)
    out_dir = os.path.join(td, "model", "sub")
    obj.save_params(out_dir)
    rec["files"] = sorted(os.listdir(out_dir))
    rec["params_json"] = open(os.path.join(out_dir, "model_params.json")).read()
    for with_history in (False, True):
        if with_history:
            open(os.path.join(out_dir, "history.csv"), "w").write("epoch,x\n0,1\n")
        p = cls.load_saved_params(out_dir, verbose=with_history)
        rec[f"loaded_{int(with_history)}"] = {
            "keys": sorted(p), "history": None if p["history"] is None
            else os.path.relpath(p["history"], td),
            "means": {k: [float(v), type(v).__name__] for k, v in p["means"].items()},
            "stdevs": {k: [float(v), type(v).__name__] for k, v in p["stdevs"].items()},
            "meta": p["meta"]}
    conf = cls.get_optimizer_config(StubOptimizer())
    rec["optimizer_config"] = {k: [v, type(v).__name__] for k, v in conf.items()}
    state = cls.get_optimizer_state(StubOptimizer())
    rec["optimizer_state"] = {k: [float(v), type(v).__name__] for k, v in state.items()}
    rec["handler_attrs"] = cls.check_batch_handler_attrs(StubHandler())
    return rec


def make_reference_persistence_class():
    import locale
    import pprint
    from types import SimpleNamespace
    asrc = open(os.path.join(REF, "sup3r/models/abstract.py")).read()
    bsrc = open(os.path.join(REF, "sup3r/models/base.py")).read()
    isrc = open(os.path.join(REF, "sup3r/models/interface.py")).read()
    usrc = open(os.path.join(REF, "sup3r/utilities/utilities.py")).read()
    ns = {"np": np, "os": os, "json": json, "locale": locale, "pprint": pprint,
          "logger": MagicMock(), "tf": SimpleNamespace(Tensor=type("Tensor", (), {}))}
    a = usrc.index("def safe_cast(")
    exec(compile(usrc[a:usrc.index("\ndef ", a + 5)], "safe_cast", "exec"), ns)

    class RefPersist:
        pass
    for name in ("load_saved_params", "get_optimizer_config"):
        setattr(RefPersist, name, staticmethod(grab_method(asrc, name, ns)))
    RefPersist.get_optimizer_state = classmethod(grab_method(asrc, "get_optimizer_state", ns))
    RefPersist.check_batch_handler_attrs = staticmethod(
        grab_method(bsrc, "check_batch_handler_attrs", ns))
    RefPersist.save_params = grab_method(isrc, "save_params", ns)
    return RefPersist


# ---------------------------------------------------------------- Sup3rGanDC (8(f)1)
class DCHandler:
    n_space_bins, n_time_bins = 3, 4
    spatial_weights, temporal_weights = [1 / 3] * 3, [0.25] * 4

    def __init__(self):
        self.val_data = [Batch(i) for i in range(12)]
        self.updates = []

    def update_weights(self, spatial_weights, temporal_weights):
        self.updates.append([np.asarray(spatial_weights), np.asarray(temporal_weights)])


def dc_scenario(obj):
    """Per-bin validation losses -> sampler weights (sup3r/models/dc.py:18-116)."""
    def scripted(low_res, hi_res_true, weight_gen_advers):
        i = low_res[1]
        loss = 0.3 + 0.07 * ((i * 5) % 12) + 1e-3 * weight_gen_advers
        return loss, {"loss_gen_content": 0.9 * loss, "loss_gen": loss}, None, None
    obj._get_hr_exo_and_loss = scripted
    bh = DCHandler()
    total, content = obj.calc_val_loss_gen(bh, 0.5)
    details = obj.calc_val_loss(bh, 0.5)
    rec = {"details": {k: [float(v), type(v).__name__] for k, v in details.items()},
           "n_updates": len(bh.updates),
           "weight_dtypes": [str(a.dtype) for a in bh.updates[0]]}
    arrs = {"dc_total": np.asarray(total), "dc_content": np.asarray(content),
            "dc_spatial_weights": bh.updates[0][0], "dc_temporal_weights": bh.updates[0][1]}
    return rec, arrs


def make_reference_dc_object():
    src = open(os.path.join(REF, "sup3r/models/dc.py")).read()
    ns = {"np": np, "logger": MagicMock(), "Sup3rGan": object}
    exec(compile(src[src.index("class Sup3rGanDC"):], "dc.py", "exec"), ns)
    return ns["Sup3rGanDC"]()


# ---------------------------------------------------------------- normalisation (row a15)
NORM_META = {"lr_features": ["u_10m", "v_10m", "topography"],
             "hr_out_features": ["v_10m", "u_10m"], "hr_exo_features": []}
NORM_STATS = ({"u_10m": 1.5, "v_10m": -0.25, "topography": 812.0, "unused": 3.0},
              {"u_10m": 4.0, "v_10m": 0.0, "topography": 330.5, "unused": 1.0})


def norm_scenario(obj, warn_log):
    """set_norm_stats / norm_input / un_norm_output on float32 arrays; -> (record, arrays)."""
    rng = np.random.default_rng(8)
    low = rng.standard_normal((2, 3, 4, 5, 3)).astype(np.float32) * 7
    out = rng.standard_normal((2, 6, 8, 5, 2)).astype(np.float32)
    rec, arrs = {}, {}
    arrs["norm_none"] = np.asarray(obj.norm_input(low))          # no stats yet: unchanged
    obj.set_norm_stats({"u_10m": 9.0, "v_10m": 9.0, "topography": 9.0},
                       {"u_10m": 9.0, "v_10m": 9.0, "topography": 9.0})
    obj.set_norm_stats(*NORM_STATS)                               # new stats REPLACE the old ones
    rec["means"] = {k: [float(v), type(v).__name__] for k, v in obj._means.items()}
    rec["stdevs"] = {k: [float(v), type(v).__name__] for k, v in obj._stdevs.items()}
    n0 = len(warn_log)
    arrs["norm"] = np.asarray(obj.norm_input(low))
    rec["zero_std_warnings"] = len(warn_log) - n0
    arrs["unnorm"] = np.asarray(obj.un_norm_output(out))
    rec["dtypes"] = [str(arrs["norm"].dtype), str(arrs["unnorm"].dtype)]
    obj.set_norm_stats(None, NORM_STATS[1])                       # ignored
    rec["means_after_none"] = {k: float(v) for k, v in obj._means.items()}
    obj._means = {"u_10m": np.float32(1)}
    for name, fn in (("norm_missing", lambda: obj.norm_input(low)),
                     ("unnorm_missing", lambda: obj.un_norm_output(out))):
        try:
            fn()
            rec[name] = "ok"
        except Exception as e:      # noqa: BLE001
            rec[name] = type(e).__name__
    return rec, arrs


def make_reference_norm_object(warn_log):
    import pprint
    from types import SimpleNamespace
    asrc = open(os.path.join(REF, "sup3r/models/abstract.py")).read()
    ns = {"np": np, "logger": MagicMock(), "pprint": pprint,
          "warn": lambda m, *a, **k: warn_log.append(str(m)),
          "tf": SimpleNamespace(Tensor=type("Tensor", (), {}))}

    class RefNorm:
        _means = _stdevs = None
        lr_features = NORM_META["lr_features"]
        hr_out_features = NORM_META["hr_out_features"]
        hr_exo_features = NORM_META["hr_exo_features"]
    for name in ("set_norm_stats", "norm_input", "un_norm_output"):
        setattr(RefNorm, name, grab_method(asrc, name, ns))
    return RefNorm()


def main():
    Ref = make_reference_object()
    out = {"scenarios": []}
    for sc in SCENARIOS:
        obj = prime(Ref())
        obj.optimizer, obj.optimizer_disc = "opt_gen", "opt_disc"
        out["scenarios"].append(run_scenario(obj, sc))
    # history bookkeeping
    rec = pd.DataFrame()
    rows = []
    for i in range(7):
        new = {"loss_gen": 1.0 - 0.1 * i, "disc_train_frac": float(i % 2),
               "train_extra": 3.0 * i}
        if i % 3 == 0:
            new["loss_disc"] = 0.5 + 0.01 * i
        rec = Ref.update_loss_details(rec, new, 4, prefix="train_")
        rows.append({"columns": list(rec.columns), "index": [int(j) for j in rec.index],
                     "values": [[None if pd.isna(v) else float(v) for v in r]
                                for r in rec.values]})
    out["update_loss_details"] = rows
    hist = pd.DataFrame({"val_loss_gen": [1.0, 0.8, 0.7, 0.699, 0.6985, 0.6981, 0.698, 0.6979,
                                          0.6979]})
    out["early_stop"] = [[n, thr, m, bool(Ref.early_stop(hist.iloc[:n], "val_loss_gen", thr, m))]
                         for n in (3, 6, 7, 8, 9) for thr in (0.005, 0.0005) for m in (3, 5)]
    out["early_stop_none"] = bool(Ref.early_stop(None, "val_loss_gen"))
    out["weight_update_fraction"] = [
        [v, b, f, float(Ref.get_weight_update_fraction({"disc_train_frac": v}, "disc_train_frac",
                                                       update_bounds=b, update_frac=f))]
        for v in (0.2, 0.5, 0.93, 0.97, [0.1, 0.99], (0.99, 0.2))
        for b in ((0.5, 0.95), (0.9, 0.99)) for f in (0.0, 0.05)]
    obj = prime(Ref())
    out["adversarial_weights"] = [
        [frac, td, v, float(obj.update_adversarial_weights({"disc_train_frac": v}, frac,
                                                           (0.9, 0.99), 1e-3, td))]
        for frac in (0.0, 0.1) for td in (True, False) for v in (0.5, 0.95, 1.0)]
    out["finish_epoch"] = finish_epoch_scenario(make_reference_finish_object())
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        out["persistence"] = persistence_scenario(make_reference_persistence_class(), td)
    warn_log = []
    rec, arrs = norm_scenario(make_reference_norm_object(warn_log), warn_log)
    out["norm"] = rec
    out["dc"], dc_arrs = dc_scenario(make_reference_dc_object())
    arrs.update(dc_arrs)
    np.savez_compressed(OUT.replace("training_schedule.json", "norm.npz"), **arrs)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
