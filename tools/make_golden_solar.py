"""Golden record of the ``SolarCC`` loss windows from the REAL reference class:
``SolarCC`` (sup3r/models/solar_cc.py:13-324) is exec'd from its source with a numpy-backed
``tf`` stub (``tf.random.categorical`` scripted, ``tf.concat`` / ``tf.math.reduce_mean`` on
arrays) and stand-in ``_tf_discriminate / calc_loss_disc / calc_loss_gen_content`` that are
sensitive to argument order and to exactly which hours they are given: the record pins the
daylight / point-loss / 24-hour window arithmetic, which tensor goes to which call, the per-day
averaging and the detail keys.  Also ``temporal_pad``.

    python tools/make_golden_solar.py   ->  tests/golden/solar_cc.json
"""
import json
import os
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "solar_cc.json")

WINDOWS = {2: [5, 30], 3: [0, 41, 17]}            # scripted "random" window starts per n_days


class NumpyBackend:
    """The array operations the stand-ins need, on numpy (reference run)."""
    mean = staticmethod(lambda x, axis=None: np.mean(x, axis=axis))
    abs = staticmethod(np.abs)
    numel = staticmethod(lambda x: int(np.size(x)))
    first = staticmethod(lambda x: x[:, 0, 0, 0, 0])
    tofloat = staticmethod(float)


def stand_ins(xp, log):
    """(_tf_discriminate, calc_loss_disc, calc_loss_gen_content) sensitive to their inputs."""
    def discriminate(self, hi_res):
        log.append(["disc", [int(v) for v in hi_res.shape]])
        per_sample = xp.mean(hi_res, axis=(1, 2, 3, 4)) + 0.1 * xp.first(hi_res)
        return per_sample.reshape(-1, 1)

    def loss_disc(disc_out_true, disc_out_gen):
        return xp.mean(disc_out_true) - 2.0 * xp.mean(disc_out_gen) + \
            0.01 * xp.numel(disc_out_true)

    def loss_content(self, hi_res_true, hi_res_gen):
        log.append(["content", [int(v) for v in hi_res_true.shape]])
        val = xp.mean(xp.abs(hi_res_true - hi_res_gen)) + 0.5 * xp.mean(hi_res_true)
        return val, {"mean_absolute_error": val, "other": xp.mean(hi_res_gen)}
    return discriminate, loss_disc, loss_content


def load_reference(log):
    src = open(os.path.join(REF, "sup3r/models/solar_cc.py")).read()
    tf = SimpleNamespace(
        random=SimpleNamespace(categorical=lambda logits, n: np.array([WINDOWS[n]])),
        concat=lambda values, axis=0: np.concatenate([np.asarray(v) for v in values], axis=axis),
        math=SimpleNamespace(reduce_mean=lambda x, axis=None: np.mean(x, axis=axis)),
        function=lambda f=None, **k: f if f is not None else (lambda g: g))
    ns = {"np": np, "tf": tf, "logger": MagicMock(), "Sup3rGan": object}
    exec(compile(src[src.index("class SolarCC"):], "solar_cc.py", "exec"), ns)
    cls = ns["SolarCC"]
    disc, ldisc, lcont = stand_ins(NumpyBackend, log)
    cls._tf_discriminate, cls.calc_loss_gen_content = disc, lcont
    cls.calc_loss_disc = staticmethod(ldisc)
    obj = cls.__new__(cls)
    obj._t_enhance = 8
    return obj


def inputs(n_days):
    rng = np.random.default_rng(40 + n_days)
    shape = (2, 3, 4, 24 * n_days, 2)
    return rng.standard_normal(shape), rng.standard_normal(shape)


def scenario(obj, log, to_backend=lambda a: a, tofloat=float):
    rec = {}
    for n_days in (2, 3):
        true, gen = (to_backend(a) for a in inputs(n_days))
        for flags in (dict(train_gen=True, train_disc=False, compute_disc=True),
                      dict(train_gen=True, train_disc=False),
                      dict(train_gen=False, train_disc=True)):
            del log[:]
            loss, details = obj.calc_loss(true, gen, weight_gen_advers=0.05, **flags)
            key = f"{n_days}d_" + "".join(str(int(v)) for v in flags.values())
            rec[key] = {"loss": tofloat(loss),
                        "details": {k: tofloat(v) for k, v in details.items()},
                        "calls": [list(c) for c in log]}
    for name, fn in (("bad_shape", lambda: obj.calc_loss(to_backend(inputs(2)[0]),
                                                        to_backend(inputs(3)[0]))),
                     ("not_daily", lambda: obj.calc_loss(to_backend(inputs(2)[0][:, :, :, :30]),
                                                        to_backend(inputs(2)[1][:, :, :, :30])))):
        try:
            fn()
            rec[name] = "ok"
        except Exception as e:      # noqa: BLE001
            rec[name] = type(e).__name__
    low = np.zeros((1, 4, 5, 3, 1))
    hi = np.arange(1 * 8 * 10 * 20 * 1, dtype=np.float64).reshape(1, 8, 10, 20, 1)
    padded = obj.temporal_pad(low, hi)
    rec["temporal_pad"] = [list(padded.shape), float(padded.sum()), float(padded[0, 0, 0, 0, 0]),
                           float(padded[0, 0, 0, -1, 0])]
    return rec


def main():
    log = []
    rec = scenario(load_reference(log), log)
    json.dump(rec, open(OUT, "w"), indent=1)
    print("wrote", OUT)
    print({k: (round(v["loss"], 6), len(v["calls"])) for k, v in rec.items() if isinstance(v, dict)})
    print(rec["bad_shape"], rec["not_daily"], rec["temporal_pad"])


if __name__ == "__main__":
    main()
