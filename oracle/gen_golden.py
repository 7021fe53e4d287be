"""Generate tests/golden/*.npz from the REAL reference (build container only).

TEST INFRASTRUCTURE ONLY.  The reference ships no tests or golden vectors,
so parity is pinned on outputs of the reference itself: this script imports the unmodified
modules from /root/reference (torch CPU, fp32), runs them on small seeded synthetic inputs
and stores inputs + outputs + gradients.  The committed .npz files are what travels to the
GPU box; /root/reference never does.

    python -m oracle.gen_golden          # rewrites tests/golden/
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from pulpo_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32) if t.dtype.is_floating_point else t.detach().cpu().numpy()


def _save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (_np(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def _randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def adversarial_field(shape):
    """Exact integers, half-integers, signed zeros and displacements far outside the volume."""
    D0, D1, D2 = shape
    g = torch.Generator().manual_seed(77)
    f = torch.zeros(1, 3, *shape)
    pick = torch.randint(0, 6, (1, 3, *shape), generator=g)
    ints = torch.randint(-4, 5, (1, 3, *shape), generator=g).float()
    f = torch.where(pick == 0, ints, f)
    f = torch.where(pick == 1, ints + 0.5, f)
    f = torch.where(pick == 2, torch.full_like(f, -0.0), f)
    f = torch.where(pick == 3, (torch.rand(f.shape, generator=g) - 0.5) * 4 * max(shape), f)
    f = torch.where(pick == 4, (torch.rand(f.shape, generator=g) - 0.5) * 1e-3, f)
    f = torch.where(pick == 5, (torch.rand(f.shape, generator=g) - 0.5) * 2.0, f)
    return f.contiguous()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)           # serial reductions -> reproducible fixtures
    nb, ls, cp, md = ref_import.load()
    torch.autograd.set_detect_anomaly(False)

    # ------------------------------------------------------------------ warp (a2)
    shape = (10, 12, 14)
    for tag, B, C, df in [
        ("warp_c1", 2, 1, syn.make_field(shape, 0, batch=2, max_abs=3.0)),
        ("warp_c3", 1, 3, syn.make_field(shape, 1, batch=1, max_abs=3.0)),
        ("warp_c5_big", 1, 5, syn.make_field(shape, 2, batch=1, max_abs=25.0)),
        ("warp_adversarial", 1, 2, adversarial_field(shape)),
    ]:
        img = syn.make_field(shape, 10 + C, batch=B, max_abs=1.0, channels=C)
        d = df.clone().requires_grad_(True)
        im = img.clone().requires_grad_(True)
        out = nb.SpatialTransformer(shape)(d, im)
        gout = _randn(out.shape, 5)
        out.backward(gout)
        _save(tag, df=df, img=img, out=out, gout=gout, gdf=d.grad, gimg=im.grad)

    # ------------------------------------------------------------------ VecInt (a3)
    for tag, shp, B, amp in [("vecint_small", (10, 12, 14), 2, 3.0), ("vecint_large_disp", (8, 10, 12), 1, 12.0)]:
        vec = syn.make_field(shp, 3, batch=B, max_abs=amp)
        v = vec.clone().requires_grad_(True)
        out = nb.VecInt(shp, 7)(v)
        gout = _randn(out.shape, 6)
        out.backward(gout)
        _save(tag, vec=vec, out=out, gout=gout, gvec=v.grad)

    # ------------------------------------------------------------------ resize / combine (a4, a5)
    lo_shape, hi_shape = (5, 6, 7), (10, 12, 14)
    lower = syn.make_field(lo_shape, 4, batch=2, max_abs=2.0)
    indiv = syn.make_field(hi_shape, 5, batch=2, max_abs=2.0)
    lo = lower.clone().requires_grad_(True)
    ind = indiv.clone().requires_grad_(True)
    comb = nb.DFAdder()(nb.ResizeTransform(1 / 2, 3)(lo), ind)
    gout = _randn(comb.shape, 7)
    comb.backward(gout)
    _save("combine_up2", lower=lower, indiv=indiv, out=comb, gout=gout, glower=lo.grad, gindiv=ind.grad)
    for f in (2, 4, 8):
        xin = syn.make_field((3, 4, 5), 6 + f, batch=1, max_abs=2.0)
        xi = xin.clone().requires_grad_(True)
        out = nb.ResizeTransform(1 / f, 3)(xi)
        gout = _randn(out.shape, 8)
        out.backward(gout)
        _save("resize_up%d" % f, x=xin, out=out, gout=gout, gx=xi.grad)

    # ------------------------------------------------------------------ pyramids (a8, a10)
    _, yv = syn.make_pair((16, 24, 32), 1)
    arrs = {"y": yv}
    for i, size in enumerate([(8, 12, 16), (4, 6, 8), (2, 3, 4), (16, 24, 32), (5, 7, 9)]):
        arrs["size%d" % i] = np.asarray(size)
        arrs["out%d" % i] = F.interpolate(yv, size=size, mode="trilinear", align_corners=False)
    _save("target_pyramid", **arrs)
    xe, _ = syn.make_pair((16, 24, 32), 2)
    xo = torch.rand(2, 2, 5, 7, 9, generator=torch.Generator().manual_seed(3))
    _save("avgpool2", x_even=xe, out_even=F.avg_pool3d(xe, 2, 2, 0, ceil_mode=True),
          x_odd=xo, out_odd=F.avg_pool3d(xo, 2, 2, 0, ceil_mode=True))

    # ------------------------------------------------------------------ NCC (a9)
    xs, ys = syn.make_pair((18, 20, 22), 3, batch=2)
    arrs = {"pred": xs, "target": ys}
    for w in (9, 7, 5, 3):
        p = xs.clone().requires_grad_(True)
        loss = ls.NCC_loss(p, ys, win_size=w, gamma=0.05)
        loss.backward()
        arrs["loss_w%d" % w] = loss
        arrs["gpred_w%d" % w] = p.grad
    _save("ncc", **arrs)

    # ------------------------------------------------------------------ KL (a11)
    shp = (6, 8, 10)
    mu0 = syn.make_field(shp, 20, batch=2, max_abs=1.5)
    sg0 = 0.2 + 0.8 * torch.rand(2, 3, *shp, generator=torch.Generator().manual_seed(21))
    mu1 = syn.make_field(shp, 22, batch=2, max_abs=0.5)
    sg1 = 0.5 + torch.rand(2, 3, *shp, generator=torch.Generator().manual_seed(23))
    arrs = dict(mu0=mu0, sigma0=sg0, mu1=mu1, sigma1=sg1)
    for tag, m1, s1 in [("std", torch.zeros_like(mu0), torch.ones_like(sg0)), ("gen", mu1, sg1)]:
        m = mu0.clone().requires_grad_(True)
        s = sg0.clone().requires_grad_(True)
        k = ls.KL_two_gauss_with_diag_cov(m, s, m1, s1)
        k.backward()
        arrs.update({"kl_" + tag: k, "gmu0_" + tag: m.grad, "gsigma0_" + tag: s.grad})
    _save("kl_diag", **arrs)

    # ------------------------------------------------------------------ L2 reg (f-1)
    f = syn.make_field((10, 12, 14), 30, batch=2, max_abs=3.0)
    fr = f.clone().requires_grad_(True)
    r = ls.L2_reg(fr, 0.025)
    r.backward()
    _save("l2reg", f=f, loss=r, gf=fr.grad)

    # ------------------------------------------------------------------ decoder chain + losses (a6, a7, a10, a12)
    input_size, total, latent = [16, 24, 32], 3, 2
    feedback = ["samples", "velocity_fields", "individual_dfs", "combined_dfs", "final_dfs", "transformed"]
    model = md.PULPo(total_levels=total, latent_levels=latent, beta=0.1, input_size=input_size,
                     feedback=feedback, n0=2, cp_depth=0)
    torch.autograd.set_detect_anomaly(False)
    model.eval()
    x, y, dfs, mus, sigmas = syn.make_hot_path_inputs(input_size, total, latent, seed=1, batch=2)
    d = {l: dfs[l].clone().requires_grad_(True) for l in dfs}
    m = {l: mus[l].clone().requires_grad_(True) for l in dfs}
    s = {l: sigmas[l].clone().requires_grad_(True) for l in dfs}
    # src/components/pulpo.py:168-214 with the encoder outputs replaced by fixed tensors
    ae = model.autoencoder
    level_x = {0: x}
    for _ in range(ae.lk_offset):
        level_x[0] = F.avg_pool3d(level_x[0], 2, 2, 0, ceil_mode=True)
    for l in range(1, latent):
        level_x[l] = F.avg_pool3d(level_x[l - 1], 2, 2, 0, ceil_mode=True)
    level_x[0] = x
    comb, fin, moved = {}, {}, {}
    for l in reversed(range(latent)):
        dec = ae.decoders[l]
        _, _, comb[l], fin[l], moved[l] = dec(d[l], level_x[l], combined_df=comb.get(l + 1))
    pm, ps = model.prior(m, s)
    kl, kl_lv = model.hierarchical_kl_loss(pm, ps, m, s)
    kl = kl * model.beta
    rec, rec_lv = model.hierarchical_recon_loss(moved, y, {k: None for k in fin}, None, gamma=0.05)
    reg, reg_lv = model.hierarchical_regularization(fin, lamb=0.025)
    total_loss = kl + rec + reg
    total_loss.backward()
    arrs = dict(x=x, y=y, total_levels=total, latent_levels=latent, kl=kl, recon=rec, reg=reg, total=total_loss)
    for l in range(latent):
        arrs.update({
            "df%d" % l: dfs[l], "mu%d" % l: mus[l], "sigma%d" % l: sigmas[l],
            "combined%d" % l: comb[l], "final%d" % l: fin[l], "moved%d" % l: moved[l],
            "gdf%d" % l: d[l].grad, "gmu%d" % l: m[l].grad, "gsigma%d" % l: s[l].grad,
            "kl_level%d" % l: kl_lv[l], "recon_level%d" % l: rec_lv[l], "reg_level%d" % l: reg_lv[l],
        })
    _save("hot_path_3lvl", **arrs)

    # PULPo.combine_dfs (src/models.py:349-368) on the same individual fields
    with torch.no_grad():
        c2, f2 = model.combine_dfs({l: dfs[l] for l in dfs})
    arrs = {}
    for l in range(latent):
        arrs.update({"df%d" % l: dfs[l], "combined%d" % l: c2[l], "final%d" % l: f2[l]})
    _save("combine_dfs", input_size=np.asarray(input_size), **arrs)


def gen_jacdet():
    """jacobian_det / JDetStd (src/losses.py:147-204): determinant maps with and without normalisation,
    the std regulariser and its autograd gradient; a large-displacement case folds (det <= 0 somewhere)."""
    torch.set_num_threads(1)
    nb, ls, cp, md = ref_import.load()
    arrs = {}
    for tag, shape, B, amp in [("a", (9, 11, 13), 2, 3.0), ("b", (6, 7, 5), 1, 9.0)]:
        df = syn.make_field(shape, 40 + len(tag) + B, batch=B, max_abs=amp)
        d = df.clone().requires_grad_(True)
        det = ls.jacobian_det(d, normalize=True)
        loss = ls.JDetStd(d, lamb=0.7, normalize=True)
        loss.backward()
        gout = _randn(det.shape, 8)
        d2 = df.clone().requires_grad_(True)
        ls.jacobian_det(d2, normalize=False).backward(gout)
        arrs.update({"df_" + tag: df, "det_" + tag: det, "det_nonorm_" + tag: ls.jacobian_det(df, normalize=False),
                     "jdetstd_" + tag: loss, "gdf_std_" + tag: d.grad, "gout_" + tag: gout, "gdf_nonorm_" + tag: d2.grad})
    _save("jacdet", **arrs)


def _reference_method_source(path, cls, name):
    """Source of one method of the reference's evaluate.py, which cannot be imported here (matplotlib, seaborn,
    h5py are absent): parsed with ast and compiled on its own -- the method body is the unmodified reference."""
    import ast, textwrap
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == name:
                    return textwrap.dedent(ast.get_source_segment(src, fn))
    raise KeyError(name)


def gen_uncertainty():
    """f-3: the per-pair numbers of Evaluate.uncertainty (evaluate.py:1534-1545) from explicit sample stacks:
    moved_std = mean_c(std_n(all_moved)) (evaluate.py:243), mse = mean_n((all_moved - y)^2), var = moved_std^2,
    ncc = Evaluate.ncc(var, mse) (the reference's own method, evaluate.py:334-353), var.mean()."""
    torch.set_num_threads(1)
    ns = {"np": np}
    exec(_reference_method_source(os.path.join(ref_import.REF_ROOT, "evaluate.py"), "Evaluate", "ncc"), ns)
    shape, N = (6, 10, 12), 9
    g = torch.Generator().manual_seed(5)
    y = torch.rand((1, 1) + shape, generator=g)
    base = y + 0.1 * torch.randn((1, 1) + shape, generator=g)
    amp = torch.rand((1, 1) + shape, generator=g) * 0.2
    all_moved = base + amp * torch.randn((N, 1) + shape, generator=g)          # [N, 1, *S]
    moved_std = torch.mean(torch.std(all_moved, axis=0), axis=0)                # evaluate.py:243
    mse = np.array(torch.mean((all_moved - y) ** 2, axis=0).squeeze(0))        # evaluate.py:1538
    var = np.array(moved_std ** 2)
    ncc = ns["ncc"](None, var, mse)
    _save("uncertainty", all_moved=all_moved, y=y, moved_std=moved_std, mse=mse, var=var,
          ncc=np.float64(ncc), var_mean=np.float64(var.mean()))


def gen_warp_img_size():
    """a2 with an image larger than the field: a level-sized SpatialTransformer resampling a full-resolution image /
    segmentation, as Evaluate.predict does in level_res mode (evaluate.py:198, 240, 246)."""
    torch.set_num_threads(1)
    nb, ls, cp, md = ref_import.load()
    fshape, ishape, B, C = (5, 6, 7), (10, 12, 14), 2, 3
    df = syn.make_field(fshape, 21, batch=B, max_abs=2.5)
    img = syn.make_field(ishape, 22, batch=B, max_abs=1.0, channels=C)
    d, im = df.clone().requires_grad_(True), img.clone().requires_grad_(True)
    out = nb.SpatialTransformer(fshape)(d, im)
    gout = _randn(out.shape, 23)
    out.backward(gout)
    _save("warp_img_size", df=df, img=img, out=out, gout=gout, gdf=d.grad, gimg=im.grad)


def gen_resize16():
    """a4 at x16: the output resize of the coarsest latent level in df_resolution="full_res" at config 2
    (10x12x14 -> 160x192x224, src/components/pulpo.py:146,297); small volume, same factor."""
    torch.set_num_threads(1)
    nb, ls, cp, md = ref_import.load()
    xin = syn.make_field((2, 3, 2), 22, batch=1, max_abs=2.0)
    xi = xin.clone().requires_grad_(True)
    out = nb.ResizeTransform(1 / 16, 3)(xi)
    gout = _randn(out.shape, 8)
    out.backward(gout)
    _save("resize_up16", x=xin, out=out, gout=gout, gx=xi.grad)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "jacdet":
        gen_jacdet()
    elif len(sys.argv) > 1 and sys.argv[1] == "uncertainty":
        gen_uncertainty()
    elif len(sys.argv) > 1 and sys.argv[1] == "warp_img_size":
        gen_warp_img_size()
    elif len(sys.argv) > 1 and sys.argv[1] == "resize16":
        gen_resize16()
    else:
        main()
        gen_jacdet()
        gen_uncertainty()
        gen_warp_img_size()
        gen_resize16()
