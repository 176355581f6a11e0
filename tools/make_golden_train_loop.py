"""Golden record of the top-level training loop from the REAL reference
method: the source text of ``Sup3rGan.train`` (sup3r/models/base.py:624-828), with the
reference's own ``update_adversarial_weights`` / ``get_weight_update_fraction``, is exec'd from
/root/reference and bound to a stand-in object whose per-epoch work (``_train_epoch``,
``calc_val_loss``, ``finish_epoch``, norm stats, model params, optimiser state, tensorboard)
is scripted and logged.  Pinned: the order and arguments of every call, the epoch numbering of a
fresh and of a continued run, the ``extras`` dictionary handed to ``finish_epoch`` (incl. the
``OptmGen/`` / ``OptmDisc/`` keys), the adaptive adversarial weight from epoch to epoch, the
early break and ``batch_handler.stop()``.

    python tools/make_golden_train_loop.py   ->  tests/golden/train_loop.json
"""
import importlib.util
import json
import os
import time
from unittest.mock import MagicMock

import numpy as np
import pandas as pd

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "train_loop.json")

spec = importlib.util.spec_from_file_location(
    "make_golden_training", os.path.join(ROOT, "tools", "make_golden_training.py"))
TT = importlib.util.module_from_spec(spec)
spec.loader.exec_module(TT)


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    if isinstance(v, (np.floating, float)):
        return float(v)
    if isinstance(v, (np.integer, int)) and not isinstance(v, bool):
        return int(v)
    return v


class Handler:
    means, stds = {"u": 1.0, "v": 2.0}, {"u": 3.0, "v": 4.0}
    s_enhance, t_enhance = 3, 4

    def __init__(self, log):
        self.log = log

    def stop(self):
        self.log.append(["stop"])


def stand_ins(log, stop_at):
    """Methods of the scripted object: every call is logged with its arguments."""
    def _train_epoch(self, batch_handler, weight_gen_advers, train_gen, train_disc,
                     disc_loss_bounds, multi_gpu=False):
        n = sum(1 for c in log if c[0] == "_train_epoch")
        log.append(["_train_epoch", float(weight_gen_advers), bool(train_gen), bool(train_disc),
                    _plain(disc_loss_bounds), bool(multi_gpu)])
        return {"train_loss_gen": 1.0 / (n + 1), "train_loss_disc": 0.6 - 0.05 * n,
                "disc_train_frac": [0.995, 0.95, 0.5, 0.85, 1.0, 0.2][n % 6]}

    def calc_val_loss(self, batch_handler, weight_gen_advers):
        log.append(["calc_val_loss", float(weight_gen_advers)])
        n = sum(1 for c in log if c[0] == "calc_val_loss")
        return {"val_loss_gen": 2.0 / n, "val_loss_disc": 0.7} if n % 2 else {"val_loss_gen": 2.0 / n}

    def finish_epoch(self, epoch, epochs, t0, loss_details, checkpoint_int, out_dir,
                     early_stop_on, early_stop_threshold, early_stop_n_epoch, extras=None):
        log.append(["finish_epoch", int(epoch), [int(e) for e in epochs], _plain(loss_details),
                    checkpoint_int, out_dir, early_stop_on, early_stop_threshold,
                    early_stop_n_epoch, _plain(extras)])
        assert t0 <= time.time()
        return stop_at is not None and int(epoch) == stop_at

    def get_optimizer_state(self, optimizer):
        return {"learning_rate": optimizer["lr"], "iterations": optimizer["it"]}
    body = {"_train_epoch": _train_epoch, "calc_val_loss": calc_val_loss,
            "finish_epoch": finish_epoch, "get_optimizer_state": get_optimizer_state,
            "set_norm_stats": lambda self, m, s: log.append(["set_norm_stats", m, s]),
            "check_batch_handler_attrs": lambda self, bh: (
                log.append(["check_batch_handler_attrs"]) or {"lr_features": ["u", "v"]}),
            "set_model_params": lambda self, **kw: log.append(["set_model_params", _plain(kw)]),
            "_init_tensorboard_writer": lambda self, out_dir: log.append(["tensorboard", out_dir])}
    return body


SCENARIOS = {
    # name: (previous history index or None, stop_at epoch, train kwargs)
    "fresh": (None, None, dict(n_epoch=3, weight_gen_advers=0.01, adaptive_update_fraction=0.05,
                               adaptive_update_bounds=(0.9, 0.99), checkpoint_int=2,
                               out_dir="./gan_{epoch}", early_stop_on="val_loss_gen")),
    "continued_no_disc": ([0, 1, 2, 3, 4], None,
                          dict(n_epoch=2, weight_gen_advers=0.002, train_disc=False,
                               adaptive_update_fraction=0.1, disc_loss_bounds=(0.3, 0.5),
                               multi_gpu=True, tensorboard_log=True, tensorboard_profile=True)),
    "early_stop": (None, 1, dict(n_epoch=5, adaptive_update_fraction=0.2,
                                 adaptive_update_bounds=(0.6, 0.9), early_stop_threshold=0.1,
                                 early_stop_n_epoch=2)),
    "defaults": ([7], None, dict(n_epoch=2)),
}


def scenario(make_obj):
    """``make_obj(log, stop_at)`` -> object with ``train`` and the stand-ins."""
    rec = {}
    for name, (hist, stop_at, kw) in SCENARIOS.items():
        log = []
        obj = make_obj(log, stop_at)
        obj._history = None if hist is None else pd.DataFrame(
            {"elapsed_time": np.arange(len(hist), dtype=float)}, index=pd.Index(hist, name="epoch"))
        obj._write_tb_profile = False
        out = obj.train(Handler(log), {"spatial": "12km", "temporal": "60min"}, **kw)
        rec[name] = {"log": _plain(log), "returned": out,
                     "write_tb_profile": bool(obj._write_tb_profile),
                     "history_columns": list(obj._history.columns),
                     "history_index_name": obj._history.index.name}
    return rec


def make_reference_object(log, stop_at):
    bsrc = open(os.path.join(REF, "sup3r/models/base.py")).read()
    ns = {"np": np, "pd": pd, "time": time, "logger": MagicMock()}
    body = stand_ins(log, stop_at)
    body["train"] = TT.grab_method(bsrc, "train", ns)
    body["update_adversarial_weights"] = TT.grab_method(bsrc, "update_adversarial_weights", ns)
    body["get_weight_update_fraction"] = staticmethod(
        TT.grab_method(bsrc, "get_weight_update_fraction", ns))
    obj = type("RefGan", (), body)()
    obj.optimizer, obj.optimizer_disc = {"lr": 1e-4, "it": 10}, {"lr": 4e-4, "it": 20}
    return obj


def main():
    rec = scenario(make_reference_object)
    json.dump(rec, open(OUT, "w"), indent=1)
    print("wrote", OUT)
    for k, v in rec.items():
        print(k, [c[:2] for c in v["log"]])


if __name__ == "__main__":
    main()
