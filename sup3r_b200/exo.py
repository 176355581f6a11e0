"""``ExoData``: the dict protocol that carries exogenous features into ``generate`` and the
forward-pass tiler (mirrors sup3r/preprocessing/data_handlers/exo.py:20-274).

``{feature: {'steps': [{'model': i, 'combine_type': 'input'|'layer'|'output', 'data': arr,
                        ['s_enhance': s, 't_enhance': t]}, ...]}}``
"""
from __future__ import annotations

import logging

logger = logging.getLogger(__name__)


class SingleExoDataStep(dict):
    """One model step of one exogenous feature."""

    def __init__(self, feature, combine_type, model, data):
        super().__init__(model=model, combine_type=combine_type, data=data)
        self.feature = feature

    @property
    def shape(self):
        return self["data"].shape


class ExoData(dict):
    """Validated dictionary of exogenous features and their per-step data."""

    def __init__(self, steps):
        super().__init__()
        if not isinstance(steps, dict):
            msg = "ExoData must be initialized with a dictionary of features."
            logger.error(msg)
            raise ValueError(msg)
        for feat, entry in steps.items():
            assert "steps" in entry, f'ExoData entry for {feat} must have a "steps" key.'
            for i, step in enumerate(entry["steps"]):
                assert "data" in step and "combine_type" in step, (
                    f"ExoData entry for {feat}, step #{i + 1}, must have a "
                    '"data" and "combine_type" key.')
        self.update(steps)

    def get_model_step_exo(self, model_step):
        """Entries whose ``model`` index equals ``model_step``."""
        out = {}
        for feature, entry in self.items():
            steps = [s for s in entry["steps"] if s["model"] == model_step]
            if steps:
                out[feature] = {"steps": steps}
        return ExoData(out)

    @staticmethod
    def _get_bounded_steps(steps, min_step, max_step=None):
        """Steps whose model index lies in [min_step, max_step) (exo.py:132-142)."""
        return [s for s in steps
                if min_step <= s["model"] and (max_step is None or s["model"] < max_step)]

    def split(self, split_steps):
        """Split into consecutive ExoData objects at the given model-step indices; the step
        indices of each part are re-based to start at zero."""
        split_steps = list(split_steps)
        if split_steps[0] != 0:
            split_steps = [0, *split_steps]
        parts = [{} for _ in split_steps]
        for feature, entry in self.items():
            for i, lo in enumerate(split_steps):
                hi = split_steps[i + 1] if i + 1 < len(split_steps) else None
                chosen = self._get_bounded_steps(entry["steps"], lo, hi)
                for s in chosen:
                    s.update({"model": s["model"] - lo})
                if chosen:
                    parts[i][feature] = {"steps": chosen}
        return [ExoData(p) for p in parts]

    def get_combine_type_data(self, feature, combine_type, model_step=None):
        """Data of the first step of ``feature`` with the requested ``combine_type``."""
        steps = self[feature]["steps"]
        if model_step is not None:
            steps = [s for s in steps if s["model"] == model_step]
        kinds = [s["combine_type"] for s in steps]
        assert combine_type in kinds, (
            f'Received exogenous_data without any combine_type = "{combine_type}" steps.')
        return steps[kinds.index(combine_type)]["data"]

    @staticmethod
    def _get_enhanced_slices(lr_slices, step):
        """Low-res slices scaled by the step's enhancement factors (exo.py:226-238)."""
        factors = [step["s_enhance"], step["s_enhance"], step["t_enhance"]]
        return [slice(s.start * f, s.stop * f) for f, s in zip(factors, lr_slices)]

    def get_chunk(self, lr_slices):
        """Exo data for one forward-pass chunk: every step's data sliced by the low-res
        slices scaled with that step's enhancement (2-D data only takes the spatial slices)."""
        chunk = {f: {"steps": []} for f in self}
        for feature in self:
            for step in self[feature]["steps"]:
                sl = self._get_enhanced_slices(lr_slices, step)
                new = {}
                for k, v in step.items():
                    new[k] = v[tuple(sl)[: len(v.shape) - 1]] if k == "data" else v
                chunk[feature]["steps"].append(new)
        return chunk
