"""``Sup3rGanWithObs``: GAN with mid-network observation fusion (mirrors
sup3r/models/with_obs.py:15-291).  During training sparse "observations" are simulated by
masking the true hi-res data (NaN where unobserved) and fed to the ``Sup3rConcatObs`` /
``Sup3rObsModel`` layers; an extra content-loss term compares observed and unobserved
locations.  At inference real observations come in as exogenous data with NaNs."""
from __future__ import annotations

import logging

import numpy as np
import torch

from ..utilities import RANDOM_GENERATOR
from .base import Sup3rGan

logger = logging.getLogger(__name__)


class Sup3rGanWithObs(Sup3rGan):
    """Sup3r GAN with observation layers."""

    def __init__(self, *args, onshore_obs_frac=None, offshore_obs_frac=None, loss_obs_weight=0.0,
                 loss_obs=None, **kwargs):
        """``onshore_obs_frac / offshore_obs_frac``: {'spatial': f | [lo, hi], 'time': ...};
        ``loss_obs``: loss of the extra observation term (defaults to ``loss``);
        ``loss_obs_weight``: its weight (with_obs.py:33-89)."""
        super().__init__(*args, **kwargs)
        self.onshore_obs_frac = {} if onshore_obs_frac is None else onshore_obs_frac
        self.offshore_obs_frac = {} if offshore_obs_frac is None else offshore_obs_frac
        loss_obs = self.loss_name if loss_obs is None else loss_obs
        self.loss_obs_name = loss_obs
        self.loss_obs_fun = self.get_loss_fun(loss_obs)
        self.loss_obs_weight = loss_obs_weight

    def _get_loss_obs_comparison(self, hi_res_true, hi_res_gen, obs_mask):
        """Loss at observed (~mask) and unobserved (mask) locations (with_obs.py:91-103).  The
        boolean gathers of the reference become masked means: for the pointwise losses
        mean(f(a[m], b[m])) == sum(f(a, b) * m) / sum(m)."""
        n_out = len(self.hr_out_features)
        hr_true = hi_res_true[..., :n_out]
        gen = hi_res_gen[..., :n_out]
        m = obs_mask[..., :n_out]

        def masked(sel):
            idx = sel.reshape(-1)
            a = gen.reshape(-1)[idx].reshape(1, -1, 1)
            b = hr_true.reshape(-1)[idx].reshape(1, -1, 1)
            if a.numel() == 0:
                return torch.full((), float("nan"), device=gen.device)
            return self.loss_obs_fun(a.contiguous(), b.contiguous())[0]

        return masked(~m), masked(m)

    @property
    def obs_training_inds(self):
        """Indices of the observation features in the true hi-res data (``_obs`` suffix
        stripped) (with_obs.py:105-117)."""
        hr_feats = [f.replace("_obs", "") for f in self.hr_features]
        return [hr_feats.index(f.replace("_obs", "")) for f in self.obs_features]

    def _get_single_obs_mask(self, hi_res, spatial_frac, time_frac=1.0):
        """Mask of one batch entry: True = not observed (with_obs.py:119-151)."""
        mask_shape = [*hi_res.shape[:3], 1, len(self.hr_out_features)]
        mask_shape[3] = hi_res.shape[3] if self.is_5d else 1
        s_mask = RANDOM_GENERATOR.uniform(size=mask_shape[1:3]) <= spatial_frac
        s_mask = s_mask[..., None, None]
        t_mask = RANDOM_GENERATOR.uniform(size=mask_shape[-2]) <= time_frac
        t_mask = t_mask[None, None, ..., None]
        mask = ~(s_mask & t_mask)
        mask = np.repeat(mask, mask_shape[-1], axis=-1)
        return mask if self.is_5d else np.squeeze(mask, axis=-2)

    def _get_obs_mask(self, hi_res, spatial_frac, time_frac=1.0):
        """Mask for a whole batch, fractions drawn per entry (with_obs.py:153-203)."""
        s_range = (spatial_frac if isinstance(spatial_frac, (list, tuple))
                   else [spatial_frac, spatial_frac])
        t_range = time_frac if isinstance(time_frac, (list, tuple)) else [time_frac, time_frac]
        s_fracs = np.clip(RANDOM_GENERATOR.uniform(*s_range, size=hi_res.shape[0]), 0, 1)
        t_fracs = np.clip(RANDOM_GENERATOR.uniform(*t_range, size=hi_res.shape[0]), 0, 1)
        return np.stack([self._get_single_obs_mask(hi_res, s, t)
                         for s, t in zip(s_fracs, t_fracs)], axis=0)

    def _get_full_obs_mask(self, hi_res):
        """Composite of an onshore and an offshore mask, selected by topography > 0
        (with_obs.py:205-222).  Returns a numpy bool array."""
        on_sf = self.onshore_obs_frac["spatial"]
        on_tf = self.onshore_obs_frac.get("time", 1.0)
        obs_mask = self._get_obs_mask(hi_res, on_sf, on_tf)
        if "topography" in self.hr_features and self.offshore_obs_frac:
            topo_idx = self.hr_features.index("topography")
            topo = hi_res[..., topo_idx]
            if isinstance(topo, torch.Tensor):
                topo = topo.detach().cpu().numpy()
            off_sf = self.offshore_obs_frac["spatial"]
            off_tf = self.offshore_obs_frac.get("time", 1.0)
            offshore_mask = self._get_obs_mask(hi_res, off_sf, off_tf)
            obs_mask = np.where(np.asarray(topo)[..., None] > 0, obs_mask, offshore_mask)
        return obs_mask

    @property
    def model_params(self):
        params = super().model_params
        params["onshore_obs_frac"] = self.onshore_obs_frac
        params["offshore_obs_frac"] = self.offshore_obs_frac
        params["loss_obs_weight"] = self.loss_obs_weight
        params["loss_obs"] = self.loss_obs_name
        return params

    def get_hr_exo_input(self, hi_res_true):
        """Standard hi-res exo input + the true data masked to sparse observations (NaN where
        unobserved) + the mask itself under ``'mask'`` (with_obs.py:240-257)."""
        exo_data = super().get_hr_exo_input(hi_res_true)
        if len(self.obs_features) == 0:
            return exo_data
        mask_np = self._get_full_obs_mask(hi_res_true)
        obs_mask = torch.from_numpy(np.ascontiguousarray(mask_np)).to(hi_res_true.device)
        inds = self.obs_training_inds
        obs = hi_res_true[..., inds]
        nan = torch.full((), float("nan"), dtype=obs.dtype, device=obs.device)
        obs = torch.where(obs_mask[..., :obs.shape[-1]], nan, obs)
        for i, f in enumerate(self.obs_features):
            exo_data[f] = obs[..., i:i + 1].contiguous()
        exo_data["mask"] = obs_mask
        return exo_data

    def _get_hr_exo_and_loss(self, low_res, hi_res_true, **calc_loss_kwargs):
        """Parent forward + loss, plus the observation loss terms (with_obs.py:259-291)."""
        out = super()._get_hr_exo_and_loss(low_res, hi_res_true, **calc_loss_kwargs)
        loss, loss_details, hi_res_gen, hi_res_exo = out
        if calc_loss_kwargs.get("train_gen", True) and "mask" in hi_res_exo:
            hi_true = hi_res_true if isinstance(hi_res_true, torch.Tensor) else \
                torch.as_tensor(np.asarray(hi_res_true, np.float32), device=hi_res_gen.device)
            mask = hi_res_exo["mask"]
            loss_obs, loss_non_obs = self._get_loss_obs_comparison(hi_true, hi_res_gen, mask)
            # (float32 division, like the reference's tf.cast(..., tf.float32) operands)
            obs_frac = float(np.float32(int((~mask).sum())) / np.float32(mask.numel()))
            loss_update = {"loss_obs": loss_obs, "loss_non_obs": loss_non_obs,
                           "obs_frac": obs_frac}
            if self.loss_obs_weight and obs_frac > 0:
                loss_obs = loss_obs * self.loss_obs_weight
                loss = loss + loss_obs
                loss_details["loss_gen"] = loss_details["loss_gen"] + loss_obs
                loss_details["loss_gen_content"] = loss_details["loss_gen_content"] + loss_obs
            loss_details.update(loss_update)
        return loss, loss_details, hi_res_gen, hi_res_exo

    def _post_batch(self, ib, b_loss_details, n_batches, previous_means):
        if "obs_frac" in b_loss_details:
            logger.debug("Batch %d out of %d has obs_frac: %.4e", ib + 1, n_batches,
                         b_loss_details["obs_frac"])
        return super()._post_batch(ib, b_loss_details, n_batches, previous_means)
