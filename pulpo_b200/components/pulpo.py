"""Drop-in for the hot-path half of the reference's ``src/components/pulpo.py``.

    SVFDecoder   <- src/components/pulpo.py:265-319  (forward :301-319)
    PULPoPrior   <- src/components/pulpo.py:323-341
    moving_pyramid (function) <- Autoencoder.forward :168-179
    feedback_resample (function) <- Autoencoder.forward :195-206

``SVFDecoder`` keeps the attributes callers reach into (``spatial_transform``, ``integrate``,
``resizer_level``, ``resizer_output``, ``velocity_field``; reference models.py:330,387 and
evaluate.py:198,240).  The VelocityField convolutions stay PyTorch (north_star): pass the
reference's own module via ``velocity_field=`` or let the small torch stand-in be built.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import functional as PF
from ..network_blocks import DFAdder, ResizeTransform, SpatialTransformer, VecInt


class _ConvUnit(nn.Module):
    """Conv3d(3, pad 1) + BatchNorm3d + LeakyReLU(0.2) -- parameter names match the reference's ConvUnit."""

    def __init__(self, cin, cout):
        super().__init__()
        self._op = nn.Sequential(nn.Conv3d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm3d(cout),
                                 nn.LeakyReLU(negative_slope=0.2, inplace=True))

    def forward(self, x):
        return self._op(x)


class _TorchVelocityField(nn.Module):
    """PyTorch stand-in for the reference's VelocityField (src/network_blocks.py:63-85); the
    convolutions are deliberately NOT part of the accelerated path."""

    def __init__(self, zdim, max_channels, depth):
        super().__init__()
        if depth == 1:
            convs = [nn.Conv3d(zdim, 3, kernel_size=3)]
        elif depth == 0:
            convs = [nn.Identity()]
        else:
            convs = [_ConvUnit(zdim, max_channels)] + [_ConvUnit(max_channels, max_channels) for _ in range(depth - 2)]
            convs = convs + [nn.Conv3d(max_channels, 3, kernel_size=1)]
        self._op = nn.Sequential(*convs)

    def forward(self, x):
        return self._op(x)


class SVFDecoder(nn.Module):
    def __init__(self, zdim, insize, outsize, df_resolution, n0: int = 32, cp_depth: int = 3,
                 velocity_field: Optional[nn.Module] = None) -> None:
        super().__init__()
        self.zdim, self.insize, self.outsize, self.cp_depth = zdim, list(insize), list(outsize), cp_depth
        if len(self.insize) != 3:
            raise NotImplementedError("pulpo_b200.SVFDecoder: only 3-D volumes are implemented")
        self.velocity_field = velocity_field if velocity_field is not None else _TorchVelocityField(zdim, n0, cp_depth)
        self.vel_resize_level = 1 / 2
        self.resizer_level = ResizeTransform(self.vel_resize_level, ndims=3)
        self.vel_resize_output = 1 / (self.outsize[0] / self.insize[0])
        self.resizer_output = ResizeTransform(self.vel_resize_output, ndims=3)
        self.combine_deformation_field = DFAdder()
        self.integrate = VecInt(self.insize, nsteps=7)
        self.spatial_transform = SpatialTransformer(self.outsize)

    def forward(self, z, input_image, combined_df=None):
        individual_df = self.velocity_field(z)
        if combined_df is None:
            combined_df = individual_df
        else:
            # resizer_level + DFAdder in one kernel: up2(2 * lower) + individual
            combined_df = PF.resize_up(combined_df, 2, self.resizer_level.factor, addend=individual_df)
        integrated_df = self.integrate(combined_df)
        integrated_df = self.resizer_output(integrated_df)
        transformed_image = self.spatial_transform(integrated_df, input_image)
        return individual_df, individual_df, combined_df, integrated_df, transformed_image


class PULPoPrior(nn.Module):
    """N(0,1) prior.  Returns *expanded* constants (stride 0) so the KL kernel can take its
    closed-form fast path and never reads prior tensors from HBM; values equal the reference's
    zeros_like / ones_like."""

    def forward(self, posterior_mus, posterior_sigmas):
        prior_mus, prior_sigmas = {}, {}
        for l in posterior_mus.keys():
            m, s = posterior_mus[l], posterior_sigmas[l]
            prior_mus[l] = torch.zeros((), dtype=torch.float32, device=m.device).expand(m.shape)
            prior_sigmas[l] = torch.ones((), dtype=torch.float32, device=s.device).expand(s.shape)
            prior_mus[l]._pulpo_const, prior_sigmas[l]._pulpo_const = 0.0, 1.0   # host-side tag
        return prior_mus, prior_sigmas


def feedback_resample(feedback_items, down_size):
    """``torch.cat([F.interpolate(item, size=down_size, mode='trilinear', align_corners=False) ...], dim=1)`` of
    Autoencoder.forward (src/components/pulpo.py:195-206): the coarser level's samples / velocity fields /
    displacement fields / warped image, resampled straight into one concatenated tensor that feeds
    ``up_blocks[k]`` (PyTorch convolutions)."""
    return PF.resample_cat(list(feedback_items), down_size)


def moving_pyramid(x, latent_levels, lk_offset, df_resolution="level_res"):
    """level_x of Autoencoder.forward (src/components/pulpo.py:168-179): repeated 2x average
    pooling (ceil_mode) down to each latent level; level 0 keeps the full-resolution image."""
    if df_resolution == "full_res":
        return {l: x for l in range(latent_levels)}
    level_x = {0: x}
    for _ in range(lk_offset):
        level_x[0] = PF.avgpool2(level_x[0])
    for l in range(1, latent_levels):
        level_x[l] = PF.avgpool2(level_x[l - 1])
    level_x[0] = x
    return level_x
