"""Content losses of the Sup3rGan hot path on device tensors.

``MeanSquaredError`` / ``MeanAbsoluteError`` stand in for ``tf.keras.losses`` (default loss,
sup3r/models/base.py:30; lookup order of sup3r/models/abstract.py:520-541: this module first).
Each loss is a callable ``loss(x1, x2) -> 0-d device tensor`` that is differentiable w.r.t.
``x1`` through our kernels (fused value + gradient in one pass over both tensors).
``LowResLoss`` mirrors sup3r/utilities/loss_metrics.py (coarsen both tensors, then MSE).
"""
from __future__ import annotations

import torch

from .autograd import ContentLossFn


class _ElementwiseMean:
    kind = 0

    def __init__(self, **kwargs):
        self.kwargs = kwargs

    def __call__(self, x1, x2):
        if tuple(x1.shape) != tuple(x2.shape):
            raise RuntimeError(f"loss inputs must have the same shape, got {tuple(x1.shape)} "
                               f"and {tuple(x2.shape)}")
        return ContentLossFn.apply(x1.contiguous(), x2.contiguous(), x1.shape[-1], self.kind)


class MeanSquaredError(_ElementwiseMean):
    """mean((x1 - x2)^2) over every element."""
    kind = 0


class MeanAbsoluteError(_ElementwiseMean):
    """mean(|x1 - x2|) over every element."""
    kind = 1


LOSSES = {"MeanSquaredError": MeanSquaredError, "MeanAbsoluteError": MeanAbsoluteError}


def get_loss_class(name):
    return LOSSES.get(name)
