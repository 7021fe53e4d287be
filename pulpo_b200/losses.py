"""Drop-in replacements for the hot-path losses of the reference's ``src/losses.py``.

    KL_two_gauss_with_diag_cov      <- src/losses.py:47-76
    NCC_loss                        <- src/losses.py:85-135
    L2_reg                          <- src/losses.py:208-222
    HierarchicalKLLoss              <- src/losses.py:225-276
    HierarchicalReconstructionLoss  <- src/losses.py:279-325
    HierarchicalRegularization      <- src/losses.py:327-355
    jacobian_det / JDetStd          <- src/losses.py:147-204

Same names, arguments, return types ((total, {level: loss}) tuples) and quirks (the KL
argument-order swap at :271-273, weight_dict mutated in place by similarity_pyramid, division
by len(recon_loss)).  "mse" / "dice" / KL_nondiagonal are outside this path
 and raise NotImplementedError.
"""
from __future__ import annotations

from typing import Union

import torch
import torch.nn as nn

from . import functional as PF


def KL_two_gauss_with_diag_cov(mu0, sigma0, mu1, sigma1, eps: float = 1e-10) -> torch.Tensor:
    """Returns KL[p0 || p1] assuming diagonal covariance matrices (mean over batch)."""
    return PF.kl_diag(mu0, sigma0, mu1, sigma1, eps)


def NCC_loss(y_pred, y_true, win_size=9, gamma=0.05):
    """Windowed local normalized cross-correlation loss."""
    if y_pred.dim() != 5:
        raise NotImplementedError("pulpo_b200.NCC_loss: only 3-D volumes [B,C,D0,D1,D2] are implemented")
    return PF.ncc_loss(y_pred, y_true, win_size, gamma)


def L2_reg(deformation_field: torch.Tensor, lamb=0) -> torch.Tensor:
    """L2 norm of the forward-difference gradient of the field."""
    if deformation_field.dim() != 5:
        raise NotImplementedError("pulpo_b200.L2_reg: only 3-D fields are implemented")
    return PF.l2_reg(deformation_field, lamb)


def jacobian_det(deformation_field: torch.Tensor, lamb=None, normalize=True) -> torch.Tensor:
    """Jacobian determinant map of a displacement field, [B,3,D,H,W] -> [B,D,H,W] (reference :147-199)."""
    if deformation_field.dim() != 5:
        raise NotImplementedError("pulpo_b200.jacobian_det: only 3-D fields are implemented")
    return PF.jacobian_det(deformation_field, normalize)


def JDetStd(deformation_field: torch.Tensor, lamb=0, normalize=True) -> torch.Tensor:
    """The standard deviation of the Jacobian determinant as a regularization loss (reference :202-204)."""
    if deformation_field.dim() != 5:
        raise NotImplementedError("pulpo_b200.JDetStd: only 3-D fields are implemented")
    return PF.jdet_std(deformation_field, lamb, normalize)


class HierarchicalKLLoss(nn.Module):
    def __init__(self, KL_divergence, weight_dict, similarity_pyramid: bool, level_sizes=None) -> None:
        super().__init__()
        self.weight_dict = weight_dict
        if similarity_pyramid:
            for l in self.weight_dict.keys():
                self.weight_dict[l] = self.weight_dict[l] / 2 ** l
        if getattr(KL_divergence, "__name__", "") == "KL_nondiagonal":
            raise NotImplementedError("pulpo_b200: KL_nondiagonal is outside the accelerated path")
        self.KL_divergence = KL_divergence

    def forward(self, prior_mus, prior_sigmas, posterior_mus, posterior_sigmas):
        assert self.weight_dict.keys() == prior_mus.keys()
        assert prior_mus.keys() == prior_sigmas.keys() == posterior_mus.keys() == posterior_sigmas.keys()
        kl_loss = 0.0
        all_levels = {}
        for l, w in self.weight_dict.items():
            # posterior is p0, prior is p1 (reference :271-273)
            all_levels[l] = w * self.KL_divergence(posterior_mus[l], posterior_sigmas[l], prior_mus[l], prior_sigmas[l])
            kl_loss += all_levels[l]
        return kl_loss, all_levels


class HierarchicalReconstructionLoss(nn.Module):
    def __init__(self, recon_loss, weight_dict, similarity_pyramid: bool, ndims: int, window_size) -> None:
        super().__init__()
        self.recon_loss = recon_loss
        self.weight_dict = weight_dict
        if similarity_pyramid:
            for l in self.weight_dict.keys():
                self.weight_dict[l] = self.weight_dict[l] / 2 ** l
        self.window_size = window_size
        self.ndims = ndims
        if ndims != 3:
            raise NotImplementedError("pulpo_b200: only ndims == 3 is implemented")
        self.mode = "trilinear"
        for name in recon_loss:
            if name != "ncc":
                raise NotImplementedError("pulpo_b200: recon_loss %r is outside the accelerated path (only 'ncc')" % name)

    def forward(self, y_hat, y, y_hat_seg=None, seg_y=None, gamma: float = 0.05, dice_factor: int = 1):
        loss = 0.0
        all_levels = {}
        for l, w in self.weight_dict.items():
            y_target = PF.interp_to_size(y, y_hat[l].shape[2:])
            all_levels[l] = 0.0
            if "ncc" in self.recon_loss:
                all_levels[l] += w * NCC_loss(y_hat[l], y_target, gamma=gamma, win_size=self.window_size[l])
            all_levels[l] = all_levels[l] / len(self.recon_loss)
            loss += all_levels[l]
        return loss, all_levels


class HierarchicalRegularization(nn.Module):
    def __init__(self, regularizer, weight_dict, similarity_pyramid: bool) -> None:
        super().__init__()
        self.regularizer = regularizer
        self.weight_dict = weight_dict
        if similarity_pyramid:
            for l in self.weight_dict.keys():
                self.weight_dict[l] = self.weight_dict[l] / 2 ** l

    def forward(self, dfs, lamb: float = 0) -> Union[torch.Tensor, tuple]:
        assert self.weight_dict.keys() == dfs.keys()
        total_loss = 0.0
        all_levels = {}
        for l, w in self.weight_dict.items():
            all_levels[l] = w * self.regularizer(dfs[l], lamb)
            total_loss += all_levels[l]
        return total_loss, all_levels
