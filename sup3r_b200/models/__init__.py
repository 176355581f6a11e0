"""sup3r_b200.models: the ``sup3r.models`` API surface of the GAN hot path."""
from .base import Sup3rGan
from .dc import Sup3rGanDC
from .multi_step import MultiStepGan, SolarMultiStepGan
from .solar_cc import SolarCC
from .with_obs import Sup3rGanWithObs

__all__ = ["Sup3rGan", "Sup3rGanDC", "Sup3rGanWithObs", "SolarCC", "MultiStepGan",
           "SolarMultiStepGan"]
