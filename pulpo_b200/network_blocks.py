"""Drop-in replacements for the hot-path blocks of the reference's ``src/network_blocks.py``.

Same class names, constructor arguments, ``forward`` signatures and attribute names as the
reference (so ``src/components/pulpo.py`` / ``src/models.py`` can import them unchanged, see
INTEGRATION.md), but every forward/backward runs in libpulpo_b200's sm_100a kernels:

    SpatialTransformer  <- src/network_blocks.py:88-121
    ResizeTransform     <- src/network_blocks.py:124-150
    DFAdder             <- src/network_blocks.py:152-158
    VecInt              <- src/network_blocks.py:160-177
    gauss_sampler       <- src/network_blocks.py:7-8   (gauss_sampler_kl: fused with the level's KL, f-4)

Only ndims == 3 and CUDA fp32 tensors are implemented; anything else raises (no fallback).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import functional as PF
from ._lib import CPU_EXACT, FAST


def gauss_sampler(mu: torch.Tensor, sigma: torch.Tensor, var: Optional[int] = 1,
                  generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """mu + sigma * (var * N(0,1)).  ``generator`` (extension) makes MC-sample sharding
    reproducible: sample i always uses seed0 + i whichever rank draws it."""
    noise = torch.randn(sigma.shape, dtype=torch.float32, device=sigma.device, generator=generator)
    return mu + sigma * (var * noise)


def gauss_sampler_kl(mu: torch.Tensor, sigma: torch.Tensor, var: Optional[int] = 1,
                     generator: Optional[torch.Generator] = None, prior_mu: Optional[torch.Tensor] = None,
                     prior_sigma: Optional[torch.Tensor] = None, eps: float = 1e-10):
    """``gauss_sampler`` plus the level's KL[N(mu, sigma) || prior] (prior None = N(0,1), the reference's
    PULPoPrior) from one read of mu and sigma (SURVEY f-4).  Draws the same noise as ``gauss_sampler`` for
    the same generator state.  Returns ``(z, kl)``; ``kl`` is what
    ``KL_two_gauss_with_diag_cov(mu, sigma, prior_mu, prior_sigma)`` returns."""
    noise = torch.randn(sigma.shape, dtype=torch.float32, device=sigma.device, generator=generator)
    return PF.gauss_sample_kl(mu, sigma, noise, prior_mu, prior_sigma, var, eps)


def _drop_grid_key(module, state_dict, prefix, *args):
    # reference checkpoints hold a persistent identity-grid buffer (network_blocks.py:99, 82.6 MB
    # at full resolution); coordinates are computed in-kernel here, so the key is accepted and dropped
    state_dict.pop(prefix + "grid", None)


class SpatialTransformer(nn.Module):
    """Trilinear warp of ``moving_image`` [B,C,*size] by the voxel displacement ``df`` [B,3,*size]."""

    def __init__(self, size, mode="bilinear", coord_mode=CPU_EXACT):
        super().__init__()
        self.size = tuple(int(s) for s in size)
        self.mode = mode  # ignored, like the reference (:120 hard-codes "bilinear")
        self.coord_mode = coord_mode
        if len(self.size) != 3:
            raise NotImplementedError("pulpo_b200.SpatialTransformer: only 3-D volumes are implemented")
        if min(self.size) < 2:
            raise ValueError("pulpo_b200.SpatialTransformer: every axis needs size >= 2 (reference divides by S-1)")
        self._register_load_state_dict_pre_hook(_drop_grid_key, with_module=True)

    def forward(self, df, moving_image):
        if tuple(df.shape[2:]) != self.size:
            raise RuntimeError("SpatialTransformer(size=%s) got a field of size %s" % (self.size, tuple(df.shape[2:])))
        return PF.warp(df, moving_image, self.coord_mode)


class ResizeTransform(nn.Module):
    """Resize a field: rescale its values by ``factor = 1/vel_resize`` and trilinearly resize it."""

    def __init__(self, vel_resize, ndims):
        super().__init__()
        self.factor = 1.0 / vel_resize
        self.mode = "linear"
        if ndims == 2:
            self.mode = "bi" + self.mode
        elif ndims == 3:
            self.mode = "tri" + self.mode
        self.ndims = ndims

    def _int_factor(self):
        f = int(round(self.factor))
        if abs(self.factor - f) > 1e-9 or f < 2:
            raise NotImplementedError("pulpo_b200.ResizeTransform: only integer up-sampling factors are "
                                      "implemented (got %r); the hot path never down-samples fields" % self.factor)
        return f

    def forward(self, x):
        if self.factor == 1:      # strict no-op, like the reference
            return x
        if self.ndims != 3:
            raise NotImplementedError("pulpo_b200.ResizeTransform: only 3-D fields are implemented")
        return PF.resize_up(x, self._int_factor(), self.factor)


class DFAdder(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, df1, df2):
        return df1 + df2


class VecInt(nn.Module):
    """Scaling-and-squaring integration: all ``nsteps`` steps in one cooperative kernel."""

    def __init__(self, inshape, nsteps, coord_mode=CPU_EXACT):
        super().__init__()
        assert nsteps >= 0, "nsteps should be >= 0, found: %d" % nsteps
        self.nsteps = nsteps
        self.scale = 1.0 / (2 ** self.nsteps)
        # kept for attribute / state-dict compatibility (reference: self.transformer.grid)
        self.transformer = SpatialTransformer(inshape)
        self.coord_mode = coord_mode

    def forward(self, vec):
        return PF.vecint(vec, self.nsteps, self.coord_mode)
