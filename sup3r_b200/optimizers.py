"""Keras-compatible optimiser objects backed by the fused Adam kernel (``s3_adam_step``).

Stands in for ``tf.keras.optimizers`` as consumed by sup3r/models/abstract.py:321-350, 543-587,
899-912 and sup3r/models/base.py:326-348: ``Adam(learning_rate)``, ``get_config()`` /
``from_config()``, ``apply_gradients(zip(grads, weights))``, ``.variables`` (objects with
``.name`` / ``.numpy()``), ``.learning_rate`` comparable to a float.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class _OptVar:
    def __init__(self, name, tensor):
        self.name = name
        self.value = tensor

    def numpy(self):
        return self.value.detach().cpu().numpy()


class Adam:
    """keras Adam: ``m, v`` moments, ``lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)``,
    ``p -= lr_t * m / (sqrt(v) + eps)``; defaults b1 0.9, b2 0.999, eps 1e-7."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
                 amsgrad=False, name="Adam", **_):
        if amsgrad:
            raise NotImplementedError("amsgrad is not supported")
        self.learning_rate = float(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.name = name
        self.iterations = 0
        self._slots = {}   # id(weight) -> (m, v, weight name)

    def get_config(self):
        return {"name": self.name, "learning_rate": self.learning_rate, "beta_1": self.beta_1,
                "beta_2": self.beta_2, "epsilon": self.epsilon, "amsgrad": False}

    @classmethod
    def from_config(cls, config):
        return cls(**config)

    @property
    def variables(self):
        out = [_OptVar(f"{self.name}/iteration:0", torch.tensor(self.iterations))]
        for m, v, wname in self._slots.values():
            base = wname.replace(":0", "")
            out.append(_OptVar(f"{self.name}/m/{base}:0", m))
            out.append(_OptVar(f"{self.name}/v/{base}:0", v))
        return out

    def slots_for(self, var):
        """(m, v) moment tensors of a weight (created on first use)."""
        slot = self._slots.get(id(var))
        if slot is None:
            w = var.value
            slot = (torch.zeros_like(w), torch.zeros_like(w), var.name)
            self._slots[id(var)] = slot
        return slot[0], slot[1]

    def apply_gradients(self, grads_and_vars):
        """``grads_and_vars``: iterable of (grad tensor, Variable)."""
        self.iterations += 1
        for g, var in grads_and_vars:
            if g is None:
                continue
            w = var.value
            slot = self._slots.get(id(var))
            if slot is None:
                slot = (torch.zeros_like(w), torch.zeros_like(w), var.name)
                self._slots[id(var)] = slot
            with torch.no_grad():
                ops.adam_step(w.detach(), g.detach().contiguous(), slot[0], slot[1],
                              self.learning_rate, self.beta_1, self.beta_2, self.epsilon,
                              self.iterations)
            var.version += 1


OPTIMIZERS = {"Adam": Adam}


def get_optimizer_class(conf):
    """Optimiser class lookup by config name (models/utilities.py:150-158)."""
    name = conf["name"] if isinstance(conf, dict) else conf
    if name not in OPTIMIZERS:
        raise ValueError(f"{name} not found in sup3r_b200 optimizers ({sorted(OPTIMIZERS)}).")
    return OPTIMIZERS[name]
