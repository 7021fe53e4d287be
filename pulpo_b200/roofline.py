"""Algorithmic HBM bytes per C-ABI call: every tensor that must exist in HBM
is read or written once, fp32.  ``algo_bytes(name, args)`` takes the ctypes argument tuple of
the call as recorded by ``_lib.profiler``."""
from __future__ import annotations

import json
import os


def _v(a):
    return getattr(a, "value", a)


def algo_bytes(name, args):
    a = [_v(x) for x in args]
    if name == "pulpo_warp3d_fwd":          # img, df, out, idx, B, C, D0, D1, D2
        B, C, n = a[4], a[5], a[6] * a[7] * a[8]
        return B * n * (12 + 8 * C)          # df 12 + img 4C + out 4C
    if name == "pulpo_warp3d_bwd":          # gout, img, df, gimg, gdf, B, C, D0..
        B, C, n = a[5], a[6], a[7] * a[8] * a[9]
        return B * n * (4 * C + 12 + 4 * C + (12 if a[4] else 0) + (4 * C if a[3] else 0))
    if name == "pulpo_warp3d_l2reg_fwd":    # img, df, out, lamb, reg_out, ws, ws_bytes, B, C, D0, D1, D2
        B, C, n = a[7], a[8], a[9] * a[10] * a[11]
        return B * n * (12 + 8 * C)          # the regulariser re-uses the field the warp reads
    if name == "pulpo_warp3d_l2reg_bwd":    # gout, img, df, gdf, lamb, reg_gloss, B, C, D0, D1, D2
        B, C, n = a[6], a[7], a[8] * a[9] * a[10]
        return B * n * (4 * C + 12 + 4 * C + 12)
    if name == "pulpo_warp3d_fwd_dpos":     # img, df, out, dpos, B, D0, D1, D2  (dpos is scratch, not algorithmic)
        return a[4] * a[5] * a[6] * a[7] * (12 + 8)
    if name == "pulpo_warp3d_bwd_dpos":     # gout, dpos, gdf, accumulate, B, D0, D1, D2: the warp backward's bytes
        return a[4] * a[5] * a[6] * a[7] * (4 + 12 + 4 + 12)
    if name == "pulpo_l2reg_fwd_bwd":       # f, lamb, out, gout, dpos, gf, accumulate, ws, bytes, B, C, D0, D1, D2
        n = a[9] * a[11] * a[12] * a[13]
        return n * (4 + 12 + 4 + 12) if a[3] else n * a[10] * 8   # with the product: same bytes as warp3d_l2reg_bwd
    if name == "pulpo_vecint_fwd":          # vec, out, ws, ws_bytes, nsteps, save, B, D0, D1, D2
        return a[4] * 24 * a[6] * a[7] * a[8] * a[9]
    if name == "pulpo_vecint_bwd":          # gout, saved, gvec, scratch, bytes, nsteps, B, D0..
        return a[5] * 36 * a[6] * a[7] * a[8] * a[9]
    if name == "pulpo_vecint_multi_fwd":    # levels*, nlevels, nsteps, save, B, mode
        return sum(a[2] * 24 * a[4] * lv.D0 * lv.D1 * lv.D2 for lv in args[0][:a[1]])
    if name == "pulpo_vecint_multi_bwd":    # levels*, nlevels, nsteps, B, mode
        return sum(a[2] * 36 * a[3] * lv.D0 * lv.D1 * lv.D2 for lv in args[0][:a[1]])
    if name == "pulpo_combine_vecint_multi_fwd":   # levels*, indiv*, nlevels, nsteps, save, B, mode
        lv = args[0][:a[2]]
        n = [v.D0 * v.D1 * v.D2 for v in lv]
        comb = sum(12 * n[l + 1] + 24 * n[l] for l in range(len(n) - 1))     # combine 12 n_{k+1} + 24 n_k
        return a[5] * (sum(a[3] * 24 * v for v in n) + comb)
    if name == "pulpo_combine_vecint_multi_bwd":   # levels*, nlevels, nsteps, B, mode
        lv = args[0][:a[1]]
        n = [v.D0 * v.D1 * v.D2 for v in lv]
        comb = sum(12 * n[l] + 12 * n[l + 1] for l in range(len(n) - 1))     # combine bwd 12 n_k + 12 n_{k+1}
        return a[3] * (sum(a[2] * 36 * v for v in n) + comb)
    if name == "pulpo_resize_up_fwd":       # x, addend, out, factor, scale, B, C, d0, d1, d2
        f, B, C, n = a[3], a[5], a[6], a[7] * a[8] * a[9]
        return B * C * 4 * (n + n * f ** 3 * (2 if a[1] else 1))
    if name == "pulpo_resize_up_bwd":       # gout, gx, factor, scale, accumulate, B, C, d0, d1, d2
        f, B, C, n = a[2], a[5], a[6], a[7] * a[8] * a[9]
        return B * C * 4 * (n * (2 if a[4] else 1) + n * f ** 3)
    if name == "pulpo_resize_up2_bwd_dpos":  # gout, dpos, gx, scale, accumulate, B, d0, d1, d2: warp bwd (32 B / voxel) + x2 adjoint
        n = a[5] * a[6] * a[7] * a[8]
        return n * 8 * 32 + n * 12 * (2 if a[4] else 1)
    if name == "pulpo_interp_size_fwd":     # x, out, B, C, i0, i1, i2, o0, o1, o2
        return a[2] * a[3] * 4 * (a[4] * a[5] * a[6] + a[7] * a[8] * a[9])
    if name == "pulpo_avgpool2_fwd":        # x, out, B, C, D0, D1, D2
        n = a[4] * a[5] * a[6]
        return a[2] * a[3] * 4 * (n + ((a[4] + 1) // 2) * ((a[5] + 1) // 2) * ((a[6] + 1) // 2))
    if name == "pulpo_ncc_fwd":             # pred, target, loss, abc, ws, ws_bytes, win, gamma, B, C, D0..
        return a[8] * a[9] * a[10] * a[11] * a[12] * 8
    if name == "pulpo_ncc_bwd":             # abc, pred, target, gloss, gpred, win, gamma, B, C, D0..
        return a[7] * a[8] * a[9] * a[10] * a[11] * 12
    if name == "pulpo_avgpool2_pyramid_fwd":   # x, outs*, nlevels, B, C, D0, D1, D2
        n = a[5] * a[6] * a[7]
        return a[3] * a[4] * 4 * (n + sum(n >> (3 * (l + 1)) for l in range(a[2])))
    if name == "pulpo_kl_diag_fwd":         # mu0, s0, mu1, s1, eps, weight, out, ws, bytes, B, n
        return a[9] * a[10] * 4 * (2 + (1 if a[2] else 0) + (1 if a[3] else 0))
    if name == "pulpo_kl_diag_bwd":         # gloss, mu0, s0, mu1, s1, eps, weight, gmu, gsg, B, n
        return a[9] * a[10] * 4 * (4 + (1 if a[3] else 0) + (1 if a[4] else 0))
    if name == "pulpo_kl_n01_multi":        # levels*, nlevels, eps, B, ws, bytes
        return sum(a[3] * lv.n * 16 for lv in args[0][:a[1]])     # mu, sigma read once; gmu, gsigma written
    if name == "pulpo_l2reg_fwd":           # f, lamb, out, ws, bytes, B, C, D0..
        return a[5] * a[6] * a[7] * a[8] * a[9] * 4
    if name == "pulpo_l2reg_bwd":           # gloss, f, lamb, gf, accumulate, B, C, D0..
        return a[5] * a[6] * a[7] * a[8] * a[9] * (12 if a[4] else 8)
    if name == "pulpo_moments_update":
        return a[4] * 20
    if name == "pulpo_moments_merge":
        return a[6] * 24
    if name == "pulpo_moments_std":
        return a[3] * 8
    return 0


def measured_peaks(root):
    """HBM GB/s from the driver-written MEASURED_PEAKS.json, else the profiling recipe's fallback."""
    path = os.path.join(root, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"
