"""GAN training schedule and history bookkeeping against golden
records produced by the REAL reference methods (tools/make_golden_training.py execs the source
of sup3r/models/base.py ``_train_epoch / _train_batch / _post_batch / update_adversarial_weights
/ get_weight_update_fraction`` and abstract.py ``update_loss_details / early_stop`` on a stand-in
object with scripted loss values).  Host logic only: runs on CPU."""
import importlib.util
import json
import os

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "training_schedule.json")))
spec = importlib.util.spec_from_file_location(
    "make_golden_training", os.path.join(ROOT, "tools", "make_golden_training.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)


def make_model():
    from sup3r_b200.models import Sup3rGan

    class Scripted(Sup3rGan):
        generator_weights = "gen_weights"
        discriminator_weights = "disc_weights"
        optimizer = "opt_gen"
        optimizer_disc = "opt_disc"
        total_batches = 0

        def __init__(self):
            pass
    return T.prime(Scripted())


def _same_table(values, want):
    assert len(values) == len(want)
    for row, wrow in zip(values, want):
        for v, w in zip(row, wrow):
            assert (w is None and pd.isna(v)) or (w is not None and v == pytest.approx(w, rel=1e-12, abs=0))


@pytest.mark.parametrize("i", range(len(T.SCENARIOS)))
def test_gan_schedule_matches_reference(i):
    got = T.run_scenario(make_model(), T.SCENARIOS[i])
    want = G["scenarios"][i]
    # which network was trained on which batch, with which optimiser and flags
    assert got["calls"] == want["calls"]
    assert got["record_columns"] == want["record_columns"]
    assert got["record_index"] == want["record_index"]
    _same_table(got["record"], want["record"])
    for e, w in zip(got["epochs"], want["epochs"]):
        assert e.keys() == w.keys()
        assert all(e[k] == pytest.approx(w[k], rel=1e-12, abs=0) for k in w)


def test_history_bookkeeping_matches_reference():
    from sup3r_b200.models import Sup3rGan
    rec = pd.DataFrame()
    for i, want in enumerate(G["update_loss_details"]):
        new = {"loss_gen": 1.0 - 0.1 * i, "disc_train_frac": float(i % 2), "train_extra": 3.0 * i}
        if i % 3 == 0:
            new["loss_disc"] = 0.5 + 0.01 * i
        rec = Sup3rGan.update_loss_details(rec, new, 4, prefix="train_")
        assert list(rec.columns) == want["columns"] and list(rec.index) == want["index"]
        _same_table(rec.values, want["values"])
    hist = pd.DataFrame({"val_loss_gen": [1.0, 0.8, 0.7, 0.699, 0.6985, 0.6981, 0.698, 0.6979,
                                          0.6979]})
    for n, thr, m, want in G["early_stop"]:
        assert Sup3rGan.early_stop(hist.iloc[:n], "val_loss_gen", thr, m) is want
    assert Sup3rGan.early_stop(None, "val_loss_gen") is G["early_stop_none"]
    for v, b, f, want in G["weight_update_fraction"]:
        got = Sup3rGan.get_weight_update_fraction({"disc_train_frac": v}, "disc_train_frac",
                                                  update_bounds=tuple(b), update_frac=f)
        assert float(got) == want
    m = make_model()
    for frac, td, v, want in G["adversarial_weights"]:
        assert float(m.update_adversarial_weights({"disc_train_frac": v}, frac, (0.9, 0.99), 1e-3,
                                                  td)) == want


def test_golden_file_is_reproducible_from_the_reference_when_present():
    """Where /root/reference exists (the build container), regenerating gives the committed file."""
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    Ref = T.make_reference_object()
    obj = T.prime(Ref())
    obj.optimizer, obj.optimizer_disc = "opt_gen", "opt_disc"
    assert T.run_scenario(obj, T.SCENARIOS[0]) == G["scenarios"][0]


def test_normalisation_matches_reference(monkeypatch):
    """a15: ``set_norm_stats / norm_input / un_norm_output`` (abstract.py:133-275) on float32
    host arrays -- stored statistics (float32, replaced by new ones, ignored when None), the
    zero-stdev warning, KeyError on a missing feature and the arrays themselves, bit for bit."""
    from sup3r_b200.models import Sup3rGan, abstract
    log = []
    monkeypatch.setattr(abstract, "warn", lambda m, *a, **k: log.append(str(m)))

    class Scripted(Sup3rGan):
        def __init__(self):
            self._means = self._stdevs = None
            self._meta = dict(T.NORM_META)
    rec, arrs = T.norm_scenario(Scripted(), log)
    want = G["norm"]
    assert json.loads(json.dumps(rec)) == want
    gold = np.load(os.path.join(ROOT, "tests", "golden", "norm.npz"))
    for k, a in arrs.items():
        assert a.dtype == gold[k].dtype and np.array_equal(a, gold[k]), k


def test_finish_epoch_matches_reference():
    """a14: ``finish_epoch`` (abstract.py:698-783) -- history rows incl. the ``extras`` columns,
    checkpoint triggers (every ``checkpoint_int`` epochs and the last epoch, ``{epoch}`` in the
    directory name), early stop and its extra checkpoint."""
    got = T.finish_epoch_scenario(make_model())
    want = G["finish_epoch"]
    assert got.keys() == want.keys()
    for k in want:
        if k != "values":
            assert got[k] == want[k], k
    _same_table(got["values"], want["values"])


def test_data_centric_validation_matches_reference():
    """8(f)1 ``Sup3rGanDC`` (dc.py:18-116): per-bin validation losses, the sampler weights handed
    to ``batch_handler.update_weights`` and the reported means (value AND type), bit for bit."""
    from sup3r_b200.models import Sup3rGanDC

    class Scripted(Sup3rGanDC):
        def __init__(self):
            pass
    rec, arrs = T.dc_scenario(Scripted())
    assert json.loads(json.dumps(rec)) == G["dc"]
    gold = np.load(os.path.join(ROOT, "tests", "golden", "norm.npz"))
    for k, a in arrs.items():
        assert a.dtype == gold[k].dtype and np.array_equal(a, gold[k]), k


def test_persistence_helpers_match_reference(tmp_path):
    """a12 / a16: ``save_params`` (the bytes of model_params.json: key order, indent, casting of
    numpy / tuple / arbitrary values), ``load_saved_params`` (history path, version record
    dropped, float32 statistics), ``get_optimizer_config`` / ``get_optimizer_state``,
    ``check_batch_handler_attrs``."""
    from sup3r_b200.models import Sup3rGan
    got = T.persistence_scenario(type("Scripted", (Sup3rGan,), {}), str(tmp_path))
    want = G["persistence"]
    assert got.keys() == want.keys()
    assert got["params_json"] == want["params_json"]
    for k in want:
        assert json.loads(json.dumps(got[k])) == want[k], k
