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
        """Total and content loss of every validation bin, shape (n_space_bins, n_time_bins)
        (dc.py:18-62)."""
        shape = (batch_handler.n_space_bins, batch_handler.n_time_bins)
        total_losses = np.zeros(shape, dtype=np.float32)
        content_losses = np.zeros(shape, dtype=np.float32)
        for i, batch in enumerate(batch_handler.val_data):
            logger.info("Calculating validation loss for batch %d / %d...", i,
                        len(batch_handler.val_data))
            with torch.no_grad():
                loss, loss_details, _, _ = self._get_hr_exo_and_loss(
                    low_res=batch.low_res, hi_res_true=batch.high_res,
                    weight_gen_advers=weight_gen_advers)
            row, col = i // batch_handler.n_time_bins, i % batch_handler.n_time_bins
            total_losses[row, col] = float(loss)
            content_losses[row, col] = float(loss_details["loss_gen_content"])
        return total_losses, content_losses

    def calc_val_loss(self, batch_handler, weight_gen_advers):
        """Update the batch handler's spatial / temporal sampling weights from the per-bin
        validation losses (dc.py:64-116)."""
        logger.debug("Starting end-of-epoch validation loss calculation...")
        loss_details = {}
        total_losses, content_losses = self.calc_val_loss_gen(batch_handler, weight_gen_advers)
        t_weights = total_losses.mean(axis=0)
        t_weights /= t_weights.sum()
        s_weights = total_losses.mean(axis=1)
        s_weights /= s_weights.sum()
        logger.debug("Previous spatial weights: %s", batch_handler.spatial_weights)
        logger.debug("Previous temporal weights: %s", batch_handler.temporal_weights)
        batch_handler.update_weights(spatial_weights=s_weights, temporal_weights=t_weights)
        logger.debug("New spatiotemporal weights (space, time):\n%s",
                     total_losses / total_losses.sum())
        logger.debug("New spatial weights: %s", s_weights)
        logger.debug("New temporal weights: %s", t_weights)
        loss_details["mean_val_loss_gen"] = round(float(np.mean(total_losses)), 3)
        loss_details["mean_val_loss_gen_content"] = round(float(np.mean(content_losses)), 3)
        return loss_details
