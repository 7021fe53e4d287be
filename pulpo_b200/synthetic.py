"""Deterministic synthetic volumes and fields (host-side, CPU torch).

There is no dataset in this environment, so every test and benchmark runs on seeded
synthetic data shaped like the reference's inputs: a textured ellipsoidal
"head" that is exactly zero outside its mask (local NCC is only well conditioned on
zero-or-textured windows), a moving image derived from it, and smooth velocity fields.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _smooth_noise(shape, sigma, gen):
    """Gaussian-filtered N(0,1) noise, renormalised to unit std (separable, zero padded)."""
    t = torch.randn(1, 1, *shape, generator=gen)
    r = max(1, int(3 * sigma))
    k = torch.exp(-0.5 * (torch.arange(-r, r + 1, dtype=torch.float32) / sigma) ** 2)
    k = k / k.sum()
    for ax in range(3):
        view = [1, 1, 1, 1, 1]
        view[2 + ax] = -1
        pad = [0, 0, 0]
        pad[ax] = r
        t = F.conv3d(t, k.view(view), padding=tuple(pad))
    return (t / t.std().clamp_min(1e-12))[0, 0]


def _ellipsoid_mask(shape, frac=0.42):
    axes = [torch.arange(s, dtype=torch.float32) for s in shape]
    z, y, x = torch.meshgrid(*axes, indexing="ij")
    c = [(s - 1) / 2.0 for s in shape]
    d = ((z - c[0]) / (frac * shape[0])) ** 2 + ((y - c[1]) / (frac * shape[1])) ** 2 + \
        ((x - c[2]) / (frac * shape[2])) ** 2
    return (d <= 1.0).float()


def make_pair(shape, seed=0, batch=1):
    """Returns (moving x, fixed y), each [batch,1,*shape] fp32 in [0,1], zero background."""
    gen = torch.Generator().manual_seed(seed)
    mask = _ellipsoid_mask(shape)
    xs, ys = [], []
    sig = max(1.5, 6.0 * min(shape) / 160.0)
    for _ in range(batch):
        base = 0.35 + 0.25 * _smooth_noise(shape, sig, gen).clamp(-2, 2) / 2
        fixed = mask * (base + 0.15 * torch.rand(shape, generator=gen))
        shift = 0.08 * _smooth_noise(shape, sig * 1.3, gen).clamp(-2, 2)
        moving = mask * (base + shift + 0.15 * torch.rand(shape, generator=gen))
        xs.append(moving.clamp(0, 1))
        ys.append(fixed.clamp(0, 1))
    return torch.stack(xs)[:, None].contiguous(), torch.stack(ys)[:, None].contiguous()


def make_field(shape, seed=0, batch=1, max_abs=3.0, channels=3, sigma_vox=None):
    """Smooth random field [batch,channels,*shape], scaled so max|v| == max_abs voxels.  Default smoothing:
    sigma = 8 * min(shape) / 160 voxels (>= 1), i.e. the same PHYSICAL length scale at every pyramid level;
    ``sigma_vox`` fixes it in this grid's own voxels instead (coarser levels then carry smoother components,
    as the levels of a real Laplacian pyramid do)."""
    gen = torch.Generator().manual_seed(1000 + seed)
    sig = float(sigma_vox) if sigma_vox is not None else max(1.0, 8.0 * min(shape) / 160.0)
    f = torch.stack([torch.stack([_smooth_noise(shape, sig, gen) for _ in range(channels)])
                     for _ in range(batch)])
    return (f * (max_abs / f.abs().max().clamp_min(1e-12))).contiguous()


def level_sizes(input_size, total_levels):
    """src/components/pulpo.py:93-96 -- repeated ceil(/2)."""
    sizes = {0: [int(s) for s in input_size]}
    for k in range(total_levels - 1):
        sizes[k + 1] = [int(math.ceil(s / 2)) for s in sizes[k]]
    return sizes


def make_hot_path_inputs(input_size, total_levels, latent_levels, seed=0, batch=1, max_abs=3.0, field_sigma_vox=None):
    """Everything the hot path consumes for one step: x, y, and per latent level the
    velocity field (stand-in for the VelocityField conv output), mu and sigma."""
    x, y = make_pair(tuple(input_size), seed, batch)
    sizes = level_sizes(input_size, total_levels)
    lk = total_levels - latent_levels
    dfs, mus, sigmas = {}, {}, {}
    for l in range(latent_levels):
        s = tuple(sizes[l + lk])
        dfs[l] = make_field(s, seed * 17 + l, batch, max_abs=max_abs, sigma_vox=field_sigma_vox)
        mus[l] = make_field(s, seed * 17 + 100 + l, batch, max_abs=1.5)
        g = torch.Generator().manual_seed(7000 + seed * 17 + l)
        sigmas[l] = (0.2 + 0.8 * torch.rand(batch, 3, *s, generator=g)).contiguous()
    return x, y, dfs, mus, sigmas
