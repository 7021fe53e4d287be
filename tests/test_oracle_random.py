"""CPU: the plain-C oracle against the ATen ops the reference calls (through oracle/torch_ref.py, itself pinned on the
golden fixtures and on the live reference), on SEEDED RANDOM RAGGED shapes -- sizes the fixtures do not hold: axes of
2 and 3 voxels, odd / prime extents, batches, windows wider than an axis, displacements far outside the volume.
The C oracle is what the GPU parity tests compare the CUDA kernels with, so it has to be right at every shape."""
import numpy as np
import pytest
import torch

from conftest import assert_close, assert_grad_close, assert_loss_close
from oracle import cport, torch_ref as T

SHAPES = [(2, 2, 2), (2, 5, 3), (3, 4, 17), (7, 6, 5), (9, 11, 13), (16, 8, 12), (5, 19, 4)]


def _rng(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def _field(shape, seed, amp, batch=1, channels=3):
    return amp * torch.randn(batch, channels, *shape, generator=_rng(seed))


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("amp", [0.4, 3.0, 40.0])
def test_c_oracle_warp_random_shapes_bit_exact(shape, amp):
    """Warp value bits and int32 corner indices == torch-CPU grid_sample (reference normalisation, src/network_blocks.py:101-121)
    for small, moderate and far-out-of-volume displacements (border clamping on most voxels at amp = 40)."""
    seed = sum(shape) * 7 + int(amp * 10)
    B = 2 if shape[0] < 8 else 1
    df = _field(shape, seed, amp, batch=B)
    img = torch.rand(B, 2, *shape, generator=_rng(seed + 1))
    dfr, imr = df.clone().requires_grad_(True), img.clone().requires_grad_(True)
    ref = T.warp(dfr, imr)
    out, idx = cport.warp3d_fwd(df.numpy(), img.numpy(), want_idx=True)
    assert np.array_equal(out, ref.detach().numpy())
    # the indices: floor of the unnormalised, clamped sample position, recomputed with torch ops in the reference's order
    grid = T.identity_grid(shape)
    for a in range(3):
        S = shape[a]
        n = 2 * ((grid[:, a] + df[:, a]) / (S - 1) - 0.5)
        p = torch.clamp(((n + 1) * S - 1) / 2, 0, S - 1)
        want = torch.floor(p).to(torch.int32)        # the index contract: floor of the clamped position, 0 .. S-1
        assert np.array_equal(idx[:, a], want.numpy()), "axis %d" % a
    gout = torch.randn(ref.shape, generator=_rng(seed + 2))
    ref.backward(gout)
    gimg, gdf = cport.warp3d_bwd(gout.numpy(), df.numpy(), img.numpy())
    assert_grad_close(gimg, imr.grad.numpy(), "gimg", rtol=2e-6)
    assert_grad_close(gdf, dfr.grad.numpy(), "gdf", rtol=2e-6)


@pytest.mark.parametrize("shape", SHAPES)
def test_c_oracle_vecint_random_shapes(shape):
    vec = _field(shape, 100 + sum(shape), 2.5)
    vr = vec.clone().requires_grad_(True)
    ref = T.vecint(vr, 7)
    steps = cport.vecint_fwd(vec.numpy(), 7)
    assert np.array_equal(steps[-1], ref.detach().numpy())          # bit-identical integration
    gout = torch.randn(ref.shape, generator=_rng(5))
    ref.backward(gout)
    assert_grad_close(cport.vecint_bwd(gout.numpy(), steps), vr.grad.numpy(), "gvec", rtol=1e-5)


@pytest.mark.parametrize("shape", [(2, 2, 2), (3, 5, 2), (4, 7, 6), (6, 5, 9)])
@pytest.mark.parametrize("factor", [2, 4])
def test_c_oracle_resize_random_shapes(shape, factor):
    x = _field(shape, 200 + factor, 1.0, batch=2)
    xr = x.clone().requires_grad_(True)
    ref = T.resize_field(xr, float(factor))
    assert_close(cport.resize_up_fwd(x.numpy(), factor, float(factor)), ref.detach().numpy(), 4e-6 * factor, "resize")
    gout = torch.randn(ref.shape, generator=_rng(6))
    ref.backward(gout)
    assert_grad_close(cport.resize_up_bwd(gout.numpy(), factor, float(factor)), xr.grad.numpy(), "resize bwd", rtol=1e-5)
    if factor == 2:   # the pyramid combination (src/components/pulpo.py:308)
        add = _field(tuple(2 * s for s in shape), 201, 1.0, batch=2)
        assert_close(cport.resize_up_fwd(x.numpy(), 2, 2.0, add.numpy()), T.combine_level(x, add).numpy(), 8e-6, "combine")


@pytest.mark.parametrize("shape", [(2, 3, 4), (5, 5, 5), (9, 8, 7), (12, 10, 6)])
def test_c_oracle_pyramids_random_shapes(shape):
    x = torch.rand(2, 1, *shape, generator=_rng(300 + sum(shape)))
    ref = torch.nn.functional.avg_pool3d(x, kernel_size=2, stride=2, padding=0, ceil_mode=True)   # pulpo.py:174
    assert_close(cport.avgpool2_fwd(x.numpy()), ref.numpy(), 1e-7, "avgpool (ceil_mode: ragged last cells)")
    for size in [tuple(max(1, s // 2) for s in shape), tuple((s + 1) // 2 for s in shape), (3, 2, 5)]:
        assert_close(cport.interp_size_fwd(x.numpy(), size), T.target_to_size(x, size).numpy(), 1e-6, "interp %s" % (size,))


@pytest.mark.parametrize("shape,win", [((4, 4, 4), 9), ((6, 7, 8), 3), ((10, 9, 11), 5), ((12, 13, 9), 7), ((5, 16, 10), 9)])
def test_c_oracle_ncc_random_shapes(shape, win):
    """Windows wider than an axis (zero padding dominates) and ragged extents; batch mean (src/losses.py:134)."""
    g = _rng(400 + win + sum(shape))
    pred = torch.rand(2, 1, *shape, generator=g)
    target = (0.7 * pred + 0.3 * torch.rand(2, 1, *shape, generator=g)).contiguous()
    pr = pred.clone().requires_grad_(True)
    ref = T.ncc_loss(pr, target, win_size=win, gamma=0.05)
    ref.backward()
    loss, grad = cport.ncc(pred.numpy(), target.numpy(), win, 0.05, want_grad=True)
    assert_loss_close(loss, ref.item(), "ncc")
    assert_grad_close(grad, pr.grad.numpy(), "ncc grad", rtol=5e-5)


@pytest.mark.parametrize("shape", [(2, 2, 2), (3, 5, 7), (8, 6, 4)])
def test_c_oracle_kl_and_l2_random_shapes(shape):
    g = _rng(500 + sum(shape))
    mu0, mu1 = torch.randn(2, 3, *shape, generator=g), torch.randn(2, 3, *shape, generator=g)
    s0, s1 = 0.2 + torch.rand(2, 3, *shape, generator=g), 0.2 + torch.rand(2, 3, *shape, generator=g)
    m, s = mu0.clone().requires_grad_(True), s0.clone().requires_grad_(True)
    ref = T.kl_diag(m, s, mu1, s1)
    ref.backward()
    assert_loss_close(cport.kl_diag_fwd(mu0.numpy(), s0.numpy(), mu1.numpy(), s1.numpy()), ref.item(), "kl")
    gm, gs = cport.kl_diag_bwd(mu0.numpy(), s0.numpy(), mu1.numpy(), s1.numpy())
    assert_grad_close(gm, m.grad.numpy(), "kl gmu", rtol=1e-5)
    assert_grad_close(gs, s.grad.numpy(), "kl gsigma", rtol=1e-5)
    f = _field(shape, 501, 2.0, batch=2)
    fr = f.clone().requires_grad_(True)
    ref = T.l2_reg(fr, 0.025)
    ref.backward()
    assert_loss_close(cport.l2reg_fwd(f.numpy(), 0.025), ref.item(), "l2")
    assert_grad_close(cport.l2reg_bwd(f.numpy(), 0.025), fr.grad.numpy(), "l2 grad", rtol=1e-5)
