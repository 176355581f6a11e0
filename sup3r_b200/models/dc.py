"""``Sup3rGanDC``: data-centric GAN (mirrors sup3r/models/dc.py:14-116): the validation loss is
evaluated per (space bin, time bin) of the batch handler and fed back as sampling weights."""
from __future__ import annotations

import logging

import numpy as np
import torch

from .base import Sup3rGan

logger = logging.getLogger(__name__)


class Sup3rGanDC(Sup3rGan):
    """Data-centric model using loss across space / time bins to select training samples."""

    _graph_safe = True    # gradient step may be captured as a CUDA graph (train_graph.py)

    def calc_val_loss_gen(self, batch_handler, weight_gen_advers):
        """Total and content loss of every validation bin, two ``(n_space_bins, n_time_bins)``
        float32 arrays; validation batch ``i`` belongs to bin ``divmod(i, n_time_bins)``
        (dc.py:18-62).  Forward only: nothing is taped."""
        n_t = batch_handler.n_time_bins
        per_bin = np.zeros((2, batch_handler.n_space_bins, n_t), dtype=np.float32)
        n_val = len(batch_handler.val_data)
        for i, batch in enumerate(batch_handler.val_data):
            logger.info("Calculating validation loss for batch %d / %d...", i, n_val)
            with torch.no_grad():
                loss, details, _, _ = self._get_hr_exo_and_loss(
                    low_res=batch.low_res, hi_res_true=batch.high_res,
                    weight_gen_advers=weight_gen_advers)
            per_bin[(slice(None), *divmod(i, n_t))] = (float(loss),
                                                       float(details["loss_gen_content"]))
        return per_bin[0], per_bin[1]

    def calc_val_loss(self, batch_handler, weight_gen_advers):
        """End-of-epoch validation for the data-centric sampler (dc.py:64-116): the mean loss of
        every time bin / space bin, normalised to sum 1, becomes the sampler's new temporal /
        spatial weight (bins the generator does badly on are drawn more often)."""
        logger.debug("Starting end-of-epoch validation loss calculation...")
        total, content = self.calc_val_loss_gen(batch_handler, weight_gen_advers)

        def share(v):
            return v / v.sum()
        new = {"spatial_weights": share(total.mean(axis=1)),
               "temporal_weights": share(total.mean(axis=0))}
        logger.debug("Sampler weights (space, time) before: %s, %s; after: %s, %s; per bin:\n%s",
                     batch_handler.spatial_weights, batch_handler.temporal_weights,
                     new["spatial_weights"], new["temporal_weights"], share(total))
        batch_handler.update_weights(**new)
        return {"mean_val_loss_gen": round(np.mean(total), 3),
                "mean_val_loss_gen_content": round(np.mean(content), 3)}
