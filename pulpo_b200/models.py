"""Hot-path pieces of the reference's ``src/models.py`` (PULPo LightningModule) as plain
functions / a small nn.Module, so a maintainer can delegate to them from ``PULPo``:

    combine_dfs(individual_dfs, ...)        <- PULPo.combine_dfs            src/models.py:349-368
    transform_segmentation(decoders, ...)   <- PULPo.transform_segmentation src/models.py:370-388
    loss_config(latent_levels, ...)         <- PULPo.__init__ loss constants src/models.py:104-123
    RegistrationHotPath                     <- the starred rows of PULPo.training_step
                                               (src/models.py:134-164)
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as PF
from .components.pulpo import PULPoPrior, SVFDecoder, moving_pyramid
from .losses import (HierarchicalKLLoss, HierarchicalReconstructionLoss, HierarchicalRegularization, JDetStd,
                     KL_two_gauss_with_diag_cov, L2_reg)
from .synthetic import level_sizes


def combine_dfs(individual_dfs, input_size, df_resolution="level_res", nsteps=7):
    """Combine per-level fields coarse-to-fine and integrate them (PULPo.combine_dfs).
    Unlike the reference no VecInt/grid objects are rebuilt (and copied host->device) per call."""
    combined, final = {}, {}
    levels = sorted(individual_dfs.keys(), reverse=True)
    for l in levels:
        if l + 1 in combined:
            f = individual_dfs[l].shape[2] / individual_dfs[l + 1].shape[2]
            if f != int(f) or int(f) < 1:
                raise NotImplementedError("pulpo_b200.combine_dfs: non-integer level ratio %r" % f)
            combined[l] = individual_dfs[l] + combined[l + 1] if int(f) == 1 else \
                PF.resize_up(combined[l + 1], int(f), f, addend=individual_dfs[l])
        else:
            combined[l] = individual_dfs[l]
    for l in levels:
        v = PF.vecint(combined[l], nsteps)
        target = input_size[0] if (l == 0 or df_resolution == "full_res") else combined[l].shape[2]
        f = target / v.shape[2]
        if f != int(f) or int(f) < 1:
            raise NotImplementedError("pulpo_b200.combine_dfs: non-integer output ratio %r" % f)
        final[l] = v if int(f) == 1 else PF.resize_up(v, int(f), f)
    return combined, final


def transform_segmentation(decoders, dfs, seg, latent_levels, lk_offset, df_resolution="level_res"):
    """Warp (multi-channel) segmentation maps by each level's final field."""
    level_seg = moving_pyramid(seg, latent_levels, lk_offset, df_resolution)
    return {key: decoders[key].spatial_transform(dfs[key], level_seg[key]) for key in dfs}


def loss_config(latent_levels, lk_offset, ndims=3, df_resolution="level_res"):
    """NCC window sizes and per-level loss weights exactly as PULPo.__init__ derives them."""
    window_size = {l: 1 + 2 * (latent_levels - l) for l in range(latent_levels)}
    if latent_levels == 1:
        window_size = {0: 9}
    scale = {l: (2.0 ** ndims) ** l for l in range(latent_levels)}
    kl_w = scale.copy()
    if df_resolution == "full_res":
        rec_w = {l: 1.0 for l in range(latent_levels)}
        reg_w = {l: 1.0 for l in range(latent_levels)}
    else:
        rec_w, reg_w = scale.copy(), scale.copy()
        rec_w[0] = scale[0] / (2 ** (ndims * lk_offset))
        reg_w[0] = scale[0] / (2 ** (ndims * lk_offset))
    rec_w[0] *= 4
    return window_size, kl_w, rec_w, reg_w


class RegistrationHotPath(nn.Module):
    """Everything PULPo.training_step does between the encoder outputs and the scalar loss,
    minus the convolutions: per level (coarse to fine) combine -> integrate -> output resize ->
    warp, then hierarchical NCC + beta*KL (+ L2) losses.  ``velocity_fields[l]`` stands for the
    VelocityField conv output, ``mus`` / ``sigmas`` for the encoder's posterior."""

    def __init__(self, input_size, total_levels, latent_levels, beta=0.1, gamma=0.05, lamb=0.025,
                 df_resolution="level_res", similarity_pyramid=False, with_reg=True, regularizer="l2"):
        super().__init__()
        self.input_size = [int(s) for s in input_size]
        self.total_levels, self.latent_levels = total_levels, latent_levels
        self.lk_offset = total_levels - latent_levels
        self.beta, self.gamma, self.lamb = beta, gamma, lamb
        self.df_resolution, self.with_reg = df_resolution, with_reg
        sizes = level_sizes(self.input_size, total_levels)
        self.decoders = nn.ModuleDict()
        for l in range(latent_levels):
            insize = sizes[self.lk_offset + l]
            outsize = self.input_size if (df_resolution == "full_res" or l == 0) else insize
            self.decoders[str(l)] = SVFDecoder(3, insize, outsize, df_resolution, cp_depth=0)
        win, kl_w, rec_w, reg_w = loss_config(latent_levels, self.lk_offset, 3, df_resolution)
        self.window_size = win
        self.prior = PULPoPrior()
        self.hierarchical_kl_loss = HierarchicalKLLoss(KL_two_gauss_with_diag_cov, kl_w, similarity_pyramid)
        self.hierarchical_recon_loss = HierarchicalReconstructionLoss(["ncc"], rec_w, similarity_pyramid, 3, win)
        if regularizer not in ("l2", "jdet"):      # PULPo.__init__, src/models.py:94-99
            raise ValueError("regularizer must be 'l2' or 'jdet', got %r" % (regularizer,))
        self.hierarchical_regularization = HierarchicalRegularization(L2_reg if regularizer == "l2" else JDetStd, reg_w,
                                                                      similarity_pyramid)

    def decode(self, x, velocity_fields):
        level_x = moving_pyramid(x, self.latent_levels, self.lk_offset, self.df_resolution)
        combined, final, moved = {}, {}, {}
        for l in reversed(range(self.latent_levels)):
            _, _, combined[l], final[l], moved[l] = self.decoders[str(l)](
                velocity_fields[l], level_x[l], combined_df=combined.get(l + 1))
        return combined, final, moved

    def forward(self, x, y, velocity_fields, mus, sigmas):
        combined, final, moved = self.decode(x, velocity_fields)
        prior_mus, prior_sigmas = self.prior(mus, sigmas)
        kl, kl_levels = self.hierarchical_kl_loss(prior_mus, prior_sigmas, mus, sigmas)
        kl = kl * self.beta
        rec, rec_levels = self.hierarchical_recon_loss(moved, y, None, None, gamma=self.gamma)
        if self.with_reg:
            reg, reg_levels = self.hierarchical_regularization(final, lamb=self.lamb)
        else:
            reg, reg_levels = torch.zeros((), device=x.device), {}
        total = kl + rec + reg
        parts = {"kl": kl, "recon": rec, "reg": reg, "kl_levels": kl_levels, "recon_levels": rec_levels,
                 "reg_levels": reg_levels}
        return total, parts, {"combined": combined, "final": final, "moved": moved}
