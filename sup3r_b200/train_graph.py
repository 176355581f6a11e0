"""The gradient step as ONE CUDA graph.

A generator + discriminator step of the north-star GAN is ~1 600 kernel launches for ~29 ms of
device work: issued one by one from Python the step is bound by the host (38 - 40 ms).
``GraphedSteps`` captures ``get_single_grad`` (generator forward, discriminator forwards, losses,
the whole backward pass, the copy of the gradients into the optimiser arena) into a CUDA graph
the third time a (weights, shapes, loss arguments) combination is seen and replays it afterwards
with one ``cudaGraphLaunch``; the optimiser step stays outside (its bias-corrected learning rate
changes every step; with this library's Adam it is ONE fused kernel over the arena, see
``parallel.GradArena``).  Stands in for the ``@tf.function`` compilation of
``get_single_grad`` in the reference (sup3r/models/abstract.py:1190-1238).

What a graph bakes in, and how each is kept honest:

* tensor addresses -- the inputs are copied into static buffers; everything else is allocated
  from the graph's private pool during capture;
* the power-of-two scales of the fp16c weight packing (they follow max |w| of each kernel):
  checked before every replay with the one host read per step the eager path also does; a
  change re-captures;
* python scalars of the loss (``weight_gen_advers`` ...) and the schedule flags: part of the key
  (at most ``MAX_GRAPHS`` graphs stay alive per model, the oldest are dropped);
* the packed tensor-core weights: every weight version is bumped before a capture, so each graph
  re-packs all the weights it reads at every replay and never depends on another graph's buffers.

Only networks built from deterministic layers are captured (no noise / dropout layers), only
for the model classes that declare ``_graph_safe`` themselves (``Sup3rGan``, ``Sup3rGanDC``: their
loss code has no per-batch host-side state; ``SolarCC`` draws random windows and
``Sup3rGanWithObs`` random masks on the host per batch and stay eager), and
``SUP3R_B200_TRAIN_GRAPH=0`` turns the whole thing off.  A failed capture falls back to the eager
step for that key (logged once).
"""
from __future__ import annotations

import logging
import math
import os

import torch

logger = logging.getLogger(__name__)

WARMUP_CALLS = 2          # eager steps of a key before it is captured (lazy scratch, caches)
MAX_GRAPHS = 4            # live graphs per model (each owns a memory pool the size of one step)
_STOCHASTIC = ("noise", "dropout", "random")


def _shape(x):
    s = getattr(x, "shape", None)
    if s is None:
        import numpy as np
        s = np.shape(x)
    return tuple(int(d) for d in s)


class _Step:
    __slots__ = ("graph", "lr", "hr", "grads", "details", "scales", "arena", "replays")


class GraphedSteps:
    """Per-model cache of captured gradient steps (see module docstring)."""

    def __init__(self, model):
        self.model = model
        self._steps = {}       # key -> _Step | False (capture failed: stay eager)
        self._seen = {}
        self.stats = {"captures": 0, "replays": 0, "recaptures": 0}

    # -- eligibility ------------------------------------------------------------------------
    @staticmethod
    def enabled():
        return os.environ.get("SUP3R_B200_TRAIN_GRAPH", "1") != "0"

    def _nets(self):
        m = self.model
        return [n for n in (getattr(m, "generator", None), getattr(m, "discriminator", None))
                if n is not None]

    def _deterministic(self):
        for net in self._nets():
            for layer in getattr(net, "layers", []):
                if any(s in type(layer).__name__.lower() for s in _STOCHASTIC):
                    return False
        return True

    def _eligible(self, low_res, kwargs, multi_gpu):
        m = self.model
        if not (self.enabled() and type(m).__dict__.get("_graph_safe", False)):
            return False
        if not torch.cuda.is_available() or m.torch_device().type != "cuda":
            return False
        if any(not isinstance(v, (bool, int, float, str, type(None))) for v in kwargs.values()):
            return False      # masks / tensors in the loss arguments: eager
        return self._deterministic()

    # -- the baked weight-packing scales ----------------------------------------------------------
    def _plans(self):
        m = self.model
        plans = [m.plan_for(m.generator, m.precision)]
        if getattr(m, "discriminator", None) is not None:
            plans.append(m.plan_for(m.discriminator,
                                    "fp32" if m.precision == "fp32" else "fp16c"))
        return plans

    def _scales(self):
        """Exponent of the fp16c weight scale of every convolution (one host read for all whose
        weights changed): what ``ops.pack_weights_umma`` would pick."""
        if self.model.precision == "fp32":
            return ()
        out = []
        for plan in self._plans():
            if not plan.net.built:
                return None
            plan._refresh_weight_maxima()
            for st in plan.steps:
                conv = getattr(st, "conv", None)
                hit = conv.__dict__.get("_umma_train_cache", {}).get("wmax") if conv is not None \
                    else None
                if hit is not None:
                    out.append(math.floor(math.log2(16383.0 / hit[1])) if hit[1] > 0 else 0)
        return tuple(out)

    def _bump_versions(self):
        for net in self._nets():
            for var in net.weights:
                var.version += 1

    # -- capture / replay -----------------------------------------------------------------------
    def _capture(self, lr, hr, weights, optimizer, multi_gpu, kwargs):
        from . import parallel
        m = self.model
        st = _Step()
        st.lr, st.hr = lr.clone(), hr.clone()
        st.replays = 0
        self._bump_versions()          # every pack this step needs happens inside the graph
        st.scales = self._scales()     # (also refreshes the maxima: no host read while capturing)
        st.graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(st.graph):     # (own private pool: graphs never alias)
            grads, details = m.get_single_grad(st.lr, st.hr, weights,
                                               device_name=m.default_device, **kwargs)
            st.arena = parallel.step_arena(grads, optimizer, multi_gpu)
            if st.arena is not None:
                st.arena.stage(grads)
        st.grads = grads
        st.details = {k: v for k, v in details.items()}
        return st

    def _evict(self, keep):
        """Least-recently-captured graphs beyond MAX_GRAPHS are dropped (e.g. the adversarial
        weight changes every epoch with adaptive updates: each value is its own graph)."""
        live = [k for k, v in self._steps.items() if v is not False and k != keep]
        while len(live) + 1 > MAX_GRAPHS:
            k = live.pop(0)
            del self._steps[k]
            self._seen.pop(k, None)
        if len(self._seen) > 256:
            self._seen.clear()

    def run(self, low_res, hi_res_true, weights, optimizer, multi_gpu, kwargs):
        """One gradient step through a captured graph.  Returns the loss details, or None when
        this call has to take the eager path."""
        if not self._eligible(low_res, kwargs, multi_gpu):
            return None
        from . import parallel
        from .network import to_device_tensor
        m = self.model
        dev = m.torch_device()
        # (the storage of every weight of both networks is part of the key: a graph reads the
        #  tensors it was captured on, re-loaded weights live somewhere else)
        ptrs = tuple(var.value.data_ptr() for net in self._nets() for var in net.weights)
        key = (tuple(id(w) for w in weights), _shape(low_res), _shape(hi_res_true),
               tuple(sorted(kwargs.items())), bool(multi_gpu), id(optimizer), ptrs, m.precision)
        st = self._steps.get(key)
        if st is False:
            return None
        if st is None:
            n = self._seen.get(key, 0)
            if n < WARMUP_CALLS:
                self._seen[key] = n + 1
                return None
        lr = to_device_tensor(low_res, dev)
        hr = to_device_tensor(hi_res_true, dev)
        if st is not None and st.scales != self._scales():
            st = None                   # a kernel's max |w| crossed a power of two
            self.stats["recaptures"] += 1
        fresh = st is None
        if fresh:
            try:
                st = self._capture(lr, hr, weights, optimizer, multi_gpu, kwargs)
            except Exception as e:      # noqa: BLE001 - anything: stay on the eager path
                logger.warning("CUDA-graph capture of the gradient step failed (%s: %s); this "
                               "configuration stays on the eager path", type(e).__name__, e)
                self._steps[key] = False
                self._bump_versions()   # buffers stamped during the aborted capture hold nothing
                torch.cuda.synchronize()
                return None
            self._steps[key] = st
            self.stats["captures"] += 1
            self._evict(key)
        else:
            st.lr.copy_(lr)
            st.hr.copy_(hr)
        st.graph.replay()
        st.replays += 1
        self.stats["replays"] += 1
        parallel.sum_grads_and_step(st.grads, weights, optimizer, multi_gpu, arena=st.arena,
                                    staged=st.arena is not None)
        # (the graph's outputs are overwritten by the next replay: hand out copies)
        return {k: (v.clone() if isinstance(v, torch.Tensor) else v)
                for k, v in st.details.items()}
