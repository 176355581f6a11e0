"""Host-only cases of the reference's own test-suite, restated against this repo's classes
(they need no device: optimiser bookkeeping and ``set_model_params`` validation).

* tests/training/test_train_gan.py::test_optimizer_update   (reference :357-387)
* tests/training/test_train_gan.py::test_input_res_check    (reference :390-403)
* tests/training/test_train_gan.py::test_enhancement_check  (reference :406-421)
"""
import warnings

import pytest

from sup3r_b200 import configs as C
from sup3r_b200.models import Sup3rGan


def _model():
    # pytest.ST_FP_GEN / ST_FP_DISC of the reference: spatiotemporal/gen_3x_4x_2f.json, disc.json
    Sup3rGan.seed()
    return Sup3rGan(C.spatiotemporal_generator(2, 3, (2, 2)), C.discriminator(3),
                    learning_rate=1e-4, learning_rate_disc=4e-4, default_device="/cpu:0")


def test_optimizer_update():
    model = _model()
    assert model.optimizer.learning_rate == 1e-4
    assert model.optimizer_disc.learning_rate == 4e-4
    model.update_optimizer(option="generator", learning_rate=2)
    assert model.optimizer.learning_rate == 2
    assert model.optimizer_disc.learning_rate == 4e-4
    model.update_optimizer(option="discriminator", learning_rate=0.4)
    assert model.optimizer.learning_rate == 2
    assert model.optimizer_disc.learning_rate == 0.4
    model.update_optimizer(option="all", learning_rate=0.1)
    assert model.optimizer.learning_rate == 0.1
    assert model.optimizer_disc.learning_rate == 0.1


def test_input_res_check():
    model = _model()
    with pytest.raises(RuntimeError):
        model.set_model_params(input_resolution={"spatial": "22km", "temporal": "9min"})


def test_enhancement_check():
    model = _model()     # (fresh: nothing has read model.s_enhance, which would record it in meta)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(RuntimeError):
            model.set_model_params(input_resolution={"spatial": "12km", "temporal": "60min"},
                                   s_enhance=7, t_enhance=3)


def test_training_session_runs_train_in_a_thread_and_stops_the_handler_on_failure():
    """sup3r/models/utilities.py:30-74."""
    import threading
    from sup3r_b200.models.utilities import (SUP3R_EXO_LAYERS, SUP3R_LAYERS, SUP3R_OBS_LAYERS,
                                             TrainingSession, get_optimizer_class)
    assert set(SUP3R_LAYERS) == set(SUP3R_EXO_LAYERS) | set(SUP3R_OBS_LAYERS)
    assert get_optimizer_class({"name": "Adam"}).__name__ == "Adam"
    seen = {}

    class Handler:
        stopped = 0

        def stop(self):
            Handler.stopped += 1

    class Model:
        def train(self, batch_handler, **kw):
            seen.update(handler=batch_handler, kw=kw, thread=threading.current_thread())
    bh = Handler()
    TrainingSession(bh, Model(), n_epoch=3, input_resolution={"spatial": "4km"}).run()
    assert seen["handler"] is bh and seen["kw"]["n_epoch"] == 3
    assert seen["thread"] is not threading.main_thread() and Handler.stopped == 0
    with pytest.raises(SystemExit):          # no n_epoch: the session cannot start
        TrainingSession(bh, Model()).run()
    assert Handler.stopped == 1
