"""sup3r_b200.models: the ``sup3r.models`` API surface of the GAN hot path."""
from .base import Sup3rGan
from .multi_step import MultiStepGan

__all__ = ["Sup3rGan", "MultiStepGan"]
