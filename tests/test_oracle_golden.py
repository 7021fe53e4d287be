"""CPU: pin the oracle (plain-C restatement + torch-CPU restatement) against the golden
fixtures generated from the REAL reference (oracle/gen_golden.py).  The reference ships no
tests of its own, so these fixtures are the parity anchor."""
import numpy as np
import pytest
import torch

from conftest import FIELD_ATOL, assert_close, assert_grad_close, assert_loss_close, load_golden
from oracle import cport, torch_ref as T


@pytest.mark.parametrize("case", ["warp_c1", "warp_c3", "warp_c5_big", "warp_adversarial"])
def test_c_oracle_warp_bit_exact(case):
    g = load_golden(case)
    out = cport.warp3d_fwd(g["df"], g["img"])
    assert np.array_equal(out, g["out"])            # same bits as torch-CPU grid_sample
    gimg, gdf = cport.warp3d_bwd(g["gout"], g["df"], g["img"])
    assert_grad_close(gimg, g["gimg"], case + " gimg", rtol=1e-6)
    assert_grad_close(gdf, g["gdf"], case + " gdf", rtol=1e-6)


@pytest.mark.parametrize("case", ["vecint_small", "vecint_large_disp"])
def test_c_oracle_vecint(case):
    g = load_golden(case)
    steps = cport.vecint_fwd(g["vec"], 7)
    assert np.array_equal(steps[-1], g["out"])
    assert_grad_close(cport.vecint_bwd(g["gout"], steps), g["gvec"], case, rtol=1e-5)


def test_c_oracle_resize():
    g = load_golden("combine_up2")
    assert_close(cport.resize_up_fwd(g["lower"], 2, 2.0, g["indiv"]), g["out"], 2e-6, "combine")
    assert_grad_close(cport.resize_up_bwd(g["gout"], 2, 2.0), g["glower"], "combine bwd", rtol=1e-5)
    for f in (2, 4, 8, 16):
        g = load_golden("resize_up%d" % f)
        assert_close(cport.resize_up_fwd(g["x"], f, float(f)), g["out"], 4e-6 * max(1, f // 8), "resize %d" % f)   # values scale with f
        assert_grad_close(cport.resize_up_bwd(g["gout"], f, float(f)), g["gx"], "resize bwd %d" % f, rtol=1e-5)


def test_c_oracle_pyramids():
    g = load_golden("target_pyramid")
    for i in range(5):
        assert_close(cport.interp_size_fwd(g["y"], g["size%d" % i]), g["out%d" % i], 1e-6, "pyr %d" % i)
    g = load_golden("avgpool2")
    assert_close(cport.avgpool2_fwd(g["x_even"]), g["out_even"], 1e-7, "pool even")
    assert_close(cport.avgpool2_fwd(g["x_odd"]), g["out_odd"], 1e-7, "pool odd")


@pytest.mark.parametrize("win", [9, 7, 5, 3])
def test_c_oracle_ncc(win):
    g = load_golden("ncc")
    loss, grad = cport.ncc(g["pred"], g["target"], win, 0.05, want_grad=True)
    assert_loss_close(loss, g["loss_w%d" % win], "ncc %d" % win)
    assert_grad_close(grad, g["gpred_w%d" % win], "ncc grad %d" % win, rtol=2e-5)


def test_c_oracle_kl_and_l2():
    g = load_golden("kl_diag")
    assert_loss_close(cport.kl_diag_fwd(g["mu0"], g["sigma0"]), g["kl_std"], "kl std")
    assert_loss_close(cport.kl_diag_fwd(g["mu0"], g["sigma0"], g["mu1"], g["sigma1"]), g["kl_gen"], "kl gen")
    gm, gs = cport.kl_diag_bwd(g["mu0"], g["sigma0"], g["mu1"], g["sigma1"])
    assert_grad_close(gm, g["gmu0_gen"], "kl gmu", rtol=1e-5)
    assert_grad_close(gs, g["gsigma0_gen"], "kl gsigma", rtol=1e-5)
    g = load_golden("l2reg")
    assert_loss_close(cport.l2reg_fwd(g["f"], 0.025), g["loss"], "l2")
    assert_grad_close(cport.l2reg_bwd(g["f"], 0.025), g["gf"], "l2 grad", rtol=1e-5)


def test_torch_ref_hot_path_matches_reference_fixture():
    """oracle/torch_ref.py (what bench.py times as the CPU baseline) == the real reference."""
    g = load_golden("hot_path_3lvl")
    L, total = int(g["latent_levels"]), int(g["total_levels"])
    t = lambda k: torch.from_numpy(g[k])
    d = {l: t("df%d" % l).requires_grad_(True) for l in range(L)}
    m = {l: t("mu%d" % l).requires_grad_(True) for l in range(L)}
    s = {l: t("sigma%d" % l).requires_grad_(True) for l in range(L)}
    torch.set_num_threads(1)
    loss, parts, outs = T.hot_path_losses(t("x"), t("y"), d, m, s, total)
    loss.backward()
    assert_loss_close(loss.item(), g["total"], "total")
    assert_loss_close(parts["kl"].item(), g["kl"], "kl")
    assert_loss_close(parts["recon"].item(), g["recon"], "recon")
    assert_loss_close(parts["reg"].item(), g["reg"], "reg")
    for l in range(L):
        assert_close(outs["moved"][l].detach().numpy(), g["moved%d" % l], 1e-6, "moved")
        assert_close(outs["final"][l].detach().numpy(), g["final%d" % l], 1e-6, "final")
        assert_grad_close(d[l].grad.numpy(), g["gdf%d" % l], "gdf %d" % l, rtol=1e-5)


def test_torch_ref_combine_dfs():
    g = load_golden("combine_dfs")
    dfs = {l: torch.from_numpy(g["df%d" % l]) for l in range(2)}
    comb, fin = T.combine_dfs(dfs, [int(v) for v in g["input_size"]])
    for l in range(2):
        assert_close(comb[l].numpy(), g["combined%d" % l], 1e-6, "combined")
        assert_close(fin[l].numpy(), g["final%d" % l], 1e-6, "final")


def test_live_reference_when_present():
    """In the build container the real reference is importable: the oracle must agree with it on
    fresh seeded inputs too (skipped on the GPU box, where /root/reference does not exist)."""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    from pulpo_b200 import synthetic as syn
    nb, ls, cp, md = ref_import.load()
    shape = (9, 11, 13)
    df = syn.make_field(shape, 42, max_abs=4.0)
    img = syn.make_field(shape, 43, max_abs=1.0, channels=2)
    ref = nb.SpatialTransformer(shape)(df.clone(), img)
    assert np.array_equal(cport.warp3d_fwd(df.numpy(), img.numpy()), ref.numpy())
    ref_v = nb.VecInt(shape, 7)(df.clone())
    assert np.array_equal(cport.vecint_fwd(df.numpy(), 7)[-1], ref_v.numpy())
    x, y = syn.make_pair((14, 15, 16), 44)
    assert_loss_close(cport.ncc(x.numpy(), y.numpy(), 5, 0.05), ls.NCC_loss(x, y, win_size=5).item(), "ncc live")


def test_numpy_oracle_jacobian_det_matches_reference_fixture():
    """f-2: jacobian_det / JDetStd restatement (oracle/jacdet_ref.py) against outputs of the live reference."""
    from oracle import jacdet_ref as J
    g = load_golden("jacdet")
    for t in "ab":
        assert np.array_equal(J.jacobian_det(g["df_" + t]), g["det_" + t])
        assert np.array_equal(J.jacobian_det(g["df_" + t], normalize=False), g["det_nonorm_" + t])
        assert abs(J.jdet_std(g["df_" + t], 0.7) - float(g["jdetstd_" + t])) <= 1e-6 * abs(float(g["jdetstd_" + t]))
    assert (g["det_nonorm_b"] <= 0).any(), "fixture should contain folded voxels"


def test_uncertainty_metrics_oracle_vs_reference_golden():
    """f-3: streaming moments + squared errors + global NCC (oracle ops) against the stack-based numbers and the
    reference's own Evaluate.ncc (tests/golden/uncertainty.npz, generated by oracle/gen_golden.py)."""
    from oracle.moments_ref import TorchCpuOps, ref_global_ncc
    from pulpo_b200 import mc
    g = load_golden("uncertainty")
    moved, y = torch.from_numpy(g["all_moved"]), torch.from_numpy(g["y"])
    assert_loss_close(ref_global_ncc(g["var"], g["mse"]), float(g["ncc"]), "ncc restatement", rtol=1e-6)
    mom = mc.MCMoments(moved.shape[1:], "cpu", ops=TorchCpuOps)
    sq = mc.MCSqErr(moved.shape[1:], "cpu", ops=TorchCpuOps)
    for i in range(moved.shape[0]):
        mom.update(moved[i])
        sq.update(moved[i], y[0])
    r = mc.uncertainty_metrics(mom, sq)
    assert_close(r["var"].numpy(), g["var"], 1e-6, "var map")
    assert_close(r["mse"].numpy(), g["mse"], 1e-6, "mse map")
    assert_loss_close(float(r["ncc"]), float(g["ncc"]), "ncc(var, mse)", rtol=1e-4)   # streaming vs stacked var
    assert_loss_close(float(r["var_mean"]), float(g["var_mean"]), "var.mean()", rtol=1e-5)


def test_torch_ref_warp_with_larger_image_golden():
    """warp with a field-sized grid and a larger image (evaluate.py:198,240,246) -- torch restatement vs reference."""
    g = load_golden("warp_img_size")
    df, img = torch.from_numpy(g["df"]).requires_grad_(True), torch.from_numpy(g["img"]).requires_grad_(True)
    out = T.warp(df, img)
    assert tuple(out.shape[2:]) == tuple(df.shape[2:])
    assert_close(out.detach().numpy(), g["out"], 1e-6, "warp img-size out")
    out.backward(torch.from_numpy(g["gout"]))
    assert_grad_close(df.grad.numpy(), g["gdf"], "gdf", rtol=1e-5)
    assert_grad_close(img.grad.numpy(), g["gimg"], "gimg", rtol=1e-5)
