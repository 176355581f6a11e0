"""Module-level helpers of ``sup3r.models.utilities`` that belong to the GAN path
(sup3r/models/utilities.py:9-158): the exo / observation layer groups, the optimiser class
lookup, the tensorboard mix-in and ``TrainingSession``.  (``st_interp`` belongs to the
``LinearInterp`` model, which is out of scope.)"""
from __future__ import annotations

import logging
import sys
import threading

from ..network import SUP3R_EXO_LAYERS, SUP3R_LAYERS, SUP3R_OBS_LAYERS
from ..optimizers import get_optimizer_class
from .abstract import TensorboardMixIn

logger = logging.getLogger(__name__)

__all__ = ["SUP3R_EXO_LAYERS", "SUP3R_LAYERS", "SUP3R_OBS_LAYERS", "TensorboardMixIn",
           "TrainingSession", "get_optimizer_class"]


class TrainingSession:
    """Run ``model.train(batch_handler, **kwargs)`` in its own thread and stop the batch
    handler's producer when the session is interrupted or fails to start
    (sup3r/models/utilities.py:30-74)."""

    def __init__(self, batch_handler, model, **kwargs):
        self.batch_handler = batch_handler
        self.model = model
        self.kwargs = kwargs

    def run(self):
        """Wrap ``model.train()``."""
        worker = threading.Thread(target=self.model.train, args=(self.batch_handler,),
                                  kwargs=self.kwargs)
        try:
            logger.info("Starting training session. Training for %s epochs",
                        self.kwargs["n_epoch"])
            worker.start()
        except KeyboardInterrupt:
            self._abort(worker, "Ending training session.")
        except Exception as e:      # noqa: BLE001
            self._abort(worker, f"Ending training session. {e}")
        worker.join()
        logger.info("Finished training")

    def _abort(self, worker, msg):
        logger.info(msg)
        self.batch_handler.stop()
        if worker.ident is not None:
            worker.join()
        sys.exit()
