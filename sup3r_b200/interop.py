"""Weight interchange with a real sup3r / phygnn installation.  phygnn's ``.pkl`` files cannot be
read without TensorFlow, so ``tools/export_phygnn_weights.py`` (run in a sup3r environment) dumps
neutral files; this module loads them:

    out_dir/gen_hidden_layers.json, disc_hidden_layers.json   {"hidden_layers": [...]}
    out_dir/gen_weights.npz, disc_weights.npz                  w000, w001, ... (keras order)
    out_dir/model_params.json                                  the model's own params file
    out_dir/golden.npz                                         low_res, hi_res (reference output
                                                               with norm_in = un_norm_out = False)
"""
from __future__ import annotations

import json
import os

import numpy as np


def _read_weights(path):
    with np.load(path) as z:
        return [z[k] for k in sorted(z.files)]


def load_exported_model(export_dir, model_class="Sup3rGan", **kwargs):
    """Build a model from an export directory (see module docstring).  Returns the model; its
    generator (and discriminator, when exported) carry the exported weights."""
    from . import models
    with open(os.path.join(export_dir, "gen_hidden_layers.json")) as f:
        gen_hl = json.load(f)["hidden_layers"]
    fp_disc = os.path.join(export_dir, "disc_hidden_layers.json")
    disc_hl = []
    if os.path.exists(fp_disc):
        with open(fp_disc) as f:
            disc_hl = json.load(f)["hidden_layers"] or []
    params = {}
    fp_params = os.path.join(export_dir, "model_params.json")
    if os.path.exists(fp_params):
        cls = getattr(models, model_class)
        params = cls.load_saved_params(export_dir, verbose=False)
        params.pop("history", None)
        cname = (params.get("meta") or {}).get("class")
        if cname and hasattr(models, cname):
            model_class = cname
    params.update(kwargs)
    model = getattr(models, model_class)(gen_hl, disc_hl, **params)
    model.generator._restore(_read_weights(os.path.join(export_dir, "gen_weights.npz")))
    fp_dw = os.path.join(export_dir, "disc_weights.npz")
    if disc_hl and os.path.exists(fp_dw):
        model.discriminator._restore(_read_weights(fp_dw))
    return model


def load_golden(export_dir):
    """-> (low_res, hi_res) of ``golden.npz``."""
    with np.load(os.path.join(export_dir, "golden.npz")) as z:
        return z["low_res"], z["hi_res"]


def find_golden_dirs(root):
    """Export directories (those holding a ``golden.npz``) below ``root``."""
    out = []
    if os.path.isdir(root):
        for name in sorted(os.listdir(root)):
            d = os.path.join(root, name)
            if os.path.exists(os.path.join(d, "golden.npz")):
                out.append(d)
    return out
