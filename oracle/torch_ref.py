"""Torch-CPU restatement of PULPo's registration hot path.  TEST INFRASTRUCTURE ONLY.

The reference's arithmetic for this path lives in a third-party dependency -- PyTorch ATen
(reference pins pytorch=1.12.1, package-list.txt:134; this image has torch 2.11) -- reached
through ``grid_sample``, ``interpolate``, ``conv3d`` and ``avg_pool3d``.  This module restates
the reference's *use* of those ops as plain functions (no nn.Module state), written from
SURVEY.md section 9.  It is what ``bench.py`` times as the "reference PyTorch CPU path"
(``cpu_baseline.kind == "port"``) on the GPU box, where /root/reference does not exist, and
it is validated against the live reference by ``oracle/gen_golden.py`` + tests/golden.

Citations are path:line in the upstream repo.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- warp
def identity_grid(size, device=None):
    """src/network_blocks.py:94-99 -- [1,3,*size] voxel coordinates, fp32."""
    axes = [torch.arange(0, int(s), dtype=torch.float32, device=device) for s in size]
    return torch.stack(torch.meshgrid(*axes, indexing="ij")).unsqueeze(0)


def warp(df, img, grid=None):
    """src/network_blocks.py:101-121 -- normalise with (S-1), sample with align_corners=False,
    border padding, channel order reversed to (x, y, z) for grid_sample."""
    size = df.shape[2:]          # the transformer's `size` is the field's grid; the image may be larger (grid_sample)
    if grid is None:
        grid = identity_grid(size, df.device)
    loc = grid + df
    norm = [2 * (loc[:, a] / (int(size[a]) - 1) - 0.5) for a in range(3)]
    sample_at = torch.stack([norm[2], norm[1], norm[0]], dim=-1)
    return F.grid_sample(img, sample_at, mode="bilinear", padding_mode="border", align_corners=False)


def vecint(vec, nsteps=7, grid=None):
    """src/network_blocks.py:173-177 -- scaling and squaring."""
    if grid is None:
        grid = identity_grid(vec.shape[2:], vec.device)
    v = vec * (1.0 / (2 ** nsteps))
    for _ in range(nsteps):
        v = v + warp(v, v, grid)
    return v


# --------------------------------------------------------------------------- resize
def resize_field(x, factor):
    """src/network_blocks.py:138-150 with self.factor == factor."""
    if factor > 1:
        return F.interpolate(factor * x, scale_factor=factor, mode="trilinear", align_corners=False)
    if factor < 1:
        return factor * F.interpolate(x, scale_factor=factor, mode="trilinear", align_corners=False)
    return x


def combine_level(lower_combined, individual):
    """src/components/pulpo.py:308 -- up2(2*lower) + individual."""
    return resize_field(lower_combined, 2.0) + individual


def moving_pyramid(x, latent_levels, lk_offset, full_res=False):
    """src/components/pulpo.py:168-179."""
    if full_res:
        return {l: x for l in range(latent_levels)}
    lv = {0: x}
    for _ in range(lk_offset):
        lv[0] = F.avg_pool3d(lv[0], kernel_size=2, stride=2, padding=0, ceil_mode=True)
    for l in range(1, latent_levels):
        lv[l] = F.avg_pool3d(lv[l - 1], kernel_size=2, stride=2, padding=0, ceil_mode=True)
    lv[0] = x
    return lv


def target_to_size(y, size):
    """src/losses.py:313."""
    return F.interpolate(y, size=tuple(int(s) for s in size), mode="trilinear", align_corners=False)


# --------------------------------------------------------------------------- decoder level
def decoder_level(individual_df, image, lower_combined, out_factor, grid_in=None, grid_out=None):
    """src/components/pulpo.py:305-319 without the VelocityField conv (cp_depth=0)."""
    combined = individual_df if lower_combined is None else combine_level(lower_combined, individual_df)
    integrated = vecint(combined, 7, grid_in)
    final = resize_field(integrated, float(out_factor)) if out_factor != 1 else integrated
    moved = warp(final, image, grid_out)
    return combined, final, moved


def combine_dfs(individual_dfs, input_size, full_res=False):
    """src/models.py:349-368."""
    levels = sorted(individual_dfs.keys(), reverse=True)
    combined, final = {}, {}
    for l in levels:
        if l + 1 in combined:
            f = individual_dfs[l].shape[2] / individual_dfs[l + 1].shape[2]
            combined[l] = individual_dfs[l] + resize_field(combined[l + 1], f)
        else:
            combined[l] = individual_dfs[l]
    for l in levels:
        v = vecint(combined[l], 7)
        tgt = input_size[0] if (l == 0 or full_res) else combined[l].shape[2]
        f = tgt / v.shape[2]
        final[l] = resize_field(v, f)
    return combined, final


# --------------------------------------------------------------------------- losses
def ncc_loss(y_pred, y_true, win_size=9, gamma=0.05):
    """src/losses.py:85-135."""
    I, J = y_true, y_pred
    k = torch.ones(1, 1, win_size, win_size, win_size, dtype=I.dtype, device=I.device)
    pad = win_size // 2

    def box(t):
        return F.conv3d(t, k, stride=1, padding=pad)

    sI, sJ, sII, sJJ, sIJ = box(I), box(J), box(I * I), box(J * J), box(I * J)
    W = win_size ** 3
    uI, uJ = sI / W, sJ / W
    cross = sIJ - uJ * sI - uI * sJ + uI * uJ * W
    vI = sII - 2 * uI * sI + uI * uI * W
    vJ = sJJ - 2 * uJ * sJ + uJ * uJ * W
    cc = cross * cross / (vI * vJ + 1e-8)
    return -torch.sum(torch.mean(cc, dim=0)) * gamma


def kl_diag(mu0, sigma0, mu1, sigma1, eps=1e-10):
    """src/losses.py:47-76 -- KL[p0 || p1]."""
    v0 = sigma0.flatten(1) ** 2
    v1 = sigma1.flatten(1) ** 2
    dm = mu1.flatten(1) - mu0.flatten(1)
    t = (v0 + dm * dm) / (v1 + eps) + torch.log(v1 + eps) - torch.log(v0 + eps) - 1
    return torch.mean(0.5 * torch.sum(t, dim=1))


def l2_reg(df, lamb):
    """src/losses.py:208-222 (3-D branch)."""
    D0, D1, D2 = df.shape[-3:]
    c = df[:, :, 1:, 1:, 1:]
    d = (c - df[:, :, :-1, 1:, 1:]) ** 2 + (c - df[:, :, 1:, :-1, 1:]) ** 2 + (c - df[:, :, 1:, 1:, :-1]) ** 2
    return d.mean() * lamb * D0 * D1 * D2


def loss_weights(latent_levels, lk_offset, ndims=3, full_res=False):
    """src/models.py:104-123 -- NCC windows and per-level weights."""
    win = {l: 1 + 2 * (latent_levels - l) for l in range(latent_levels)}
    if latent_levels == 1:
        win = {0: 9}
    scale = {l: (2.0 ** ndims) ** l for l in range(latent_levels)}
    kl_w = dict(scale)
    if full_res:
        rec_w = {l: 1.0 for l in range(latent_levels)}
        reg_w = {l: 1.0 for l in range(latent_levels)}
    else:
        rec_w, reg_w = dict(scale), dict(scale)
        rec_w[0] = scale[0] / (2 ** (ndims * lk_offset))
        reg_w[0] = scale[0] / (2 ** (ndims * lk_offset))
    rec_w[0] *= 4
    return win, kl_w, rec_w, reg_w


# --------------------------------------------------------------------------- whole hot path
def hot_path_losses(x, y, individual_dfs, mus, sigmas, total_levels, beta=0.1, gamma=0.05,
                    lamb=0.025, with_reg=True, full_res=False):
    """One pass of the hot path (the starred rows): per level
    combine -> integrate -> output resize -> warp, then hierarchical NCC + KL (+ L2) losses.
    ``individual_dfs[l]`` stands for the VelocityField conv output (pulpo.py:303), ``mus`` /
    ``sigmas`` for the encoder outputs.  Returns (total, parts, outputs)."""
    L = len(individual_dfs)
    lk = total_levels - L
    win, kl_w, rec_w, reg_w = loss_weights(L, lk, 3, full_res)
    lx = moving_pyramid(x, L, lk, full_res)
    combined, final, moved = {}, {}, {}
    for l in reversed(range(L)):
        out_factor = x.shape[2] // individual_dfs[l].shape[2] if (l == 0 or full_res) else 1   # pulpo.py:146
        combined[l], final[l], moved[l] = decoder_level(
            individual_dfs[l], lx[l], combined.get(l + 1), out_factor)
    kl = sum(kl_w[l] * kl_diag(mus[l], sigmas[l], torch.zeros_like(mus[l]), torch.ones_like(sigmas[l]))
             for l in range(L)) * beta
    rec = sum(rec_w[l] * ncc_loss(moved[l], target_to_size(y, moved[l].shape[2:]), win[l], gamma)
              for l in range(L))
    reg = sum(reg_w[l] * l2_reg(final[l], lamb) for l in range(L)) if with_reg else torch.zeros(())
    total = kl + rec + reg
    return total, {"kl": kl, "recon": rec, "reg": reg}, {"combined": combined, "final": final, "moved": moved}
