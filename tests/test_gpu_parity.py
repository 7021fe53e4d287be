"""Parity of the CUDA path (through the C ABI) against the reference: committed golden
fixtures (outputs of the real reference, tests/golden) and the plain-C oracle on seeded
inputs.  Bit-exact for sampling indices; north_star tolerances for floating point."""
import numpy as np
import pytest
import torch

from conftest import FIELD_ATOL, assert_close, assert_grad_close, assert_loss_close, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def PF():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from pulpo_b200 import functional
    return functional


def dev(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.requires_grad_(True) if grad else t


# ----------------------------------------------------------------------------- warp (a2)
@pytest.mark.parametrize("case", ["warp_c1", "warp_c3", "warp_c5_big", "warp_adversarial"])
def test_warp_golden(PF, case):
    from oracle import cport
    g = load_golden(case)
    df, img = dev(g["df"], True), dev(g["img"], True)
    out = PF.warp(df, img)
    # forward arithmetic follows the CPU grid sampler op for op -> identical bits expected
    assert np.array_equal(out.detach().cpu().numpy(), g["out"]), \
        "warp fwd not bit-identical: max-abs %.3e" % np.abs(out.detach().cpu().numpy() - g["out"]).max()
    assert_close(out.detach().cpu().numpy(), g["out"], FIELD_ATOL, case + " out")
    out.backward(dev(g["gout"]))
    assert_grad_close(df.grad.cpu().numpy(), g["gdf"], case + " gdf")
    assert_grad_close(img.grad.cpu().numpy(), g["gimg"], case + " gimg")
    # integer sampling indices: bit-exact against the oracle (itself bit-identical to torch-CPU outputs)
    _, idx = PF.warp_indices(dev(g["df"]), dev(g["img"]))
    _, idx_ref = cport.warp3d_fwd(g["df"], g["img"], want_idx=True)
    assert np.array_equal(idx.cpu().numpy(), idx_ref), "%d index mismatches" % (idx.cpu().numpy() != idx_ref).sum()


def test_warp_indices_seeded_vs_oracle(PF):
    from oracle import cport
    from pulpo_b200 import synthetic as syn
    shape = (40, 48, 56)
    for seed, amp in [(0, 3.0), (1, 40.0), (2, 0.01)]:
        df = syn.make_field(shape, seed, max_abs=amp)
        img = syn.make_field(shape, 50 + seed, max_abs=1.0, channels=2)
        out, idx = PF.warp_indices(df.cuda(), img.cuda())
        out_ref, idx_ref = cport.warp3d_fwd(df.numpy(), img.numpy(), want_idx=True)
        assert np.array_equal(idx.cpu().numpy(), idx_ref)
        assert np.array_equal(out.cpu().numpy(), out_ref)
        out4 = PF.warp(df.cuda(), img.cuda())   # vectorised kernel variant
        assert np.array_equal(out4.cpu().numpy(), out_ref)


@pytest.mark.parametrize("S", [2, 3, 5, 14, 28, 63, 64, 65, 127, 128, 129, 160, 192, 224, 255, 256, 257, 1000])
def test_exact_constant_division_sweep(PF, S):
    """The kernels divide by (S-1) with a 5-FMA correctly-rounded sequence instead of a generic
    division; indices must still match the oracle's true division for every axis size."""
    from oracle import cport
    g = torch.Generator().manual_seed(S)
    shape = (2, 4, S)                       # the swept axis is D2; D0/D1 tiny
    n = 40
    df = torch.zeros(n, 3, *shape)
    df[:, 2] = (torch.rand(n, *shape, generator=g) - 0.5) * 2.2 * S
    ints = torch.randint(-S, S, df[:, 2].shape, generator=g).float()
    pick = torch.rand(df[:, 2].shape, generator=g)
    df[:, 2] = torch.where(pick < 0.15, ints, torch.where(pick < 0.3, ints + 0.5, df[:, 2]))
    df[:, 0] = (torch.rand(n, *shape, generator=g) - 0.5) * 3
    df[:, 1] = (torch.rand(n, *shape, generator=g) - 0.5) * 5
    img = torch.rand(n, 1, *shape, generator=g)
    out, idx = PF.warp_indices(df.cuda(), img.cuda())
    out_ref, idx_ref = cport.warp3d_fwd(df.numpy(), img.numpy(), want_idx=True)
    assert np.array_equal(idx.cpu().numpy(), idx_ref), "%d index mismatches" % (idx.cpu().numpy() != idx_ref).sum()
    assert np.array_equal(out.cpu().numpy(), out_ref)


def test_warp_cuda_rcp_mode_matches_oracle_mode1(PF):
    from oracle import cport
    from pulpo_b200 import synthetic as syn
    shape = (20, 24, 28)
    df = syn.make_field(shape, 3, max_abs=5.0)
    img = syn.make_field(shape, 4, max_abs=1.0, channels=1)
    _, idx = PF.warp_indices(df.cuda(), img.cuda(), coord_mode=PF.CUDA_RCP)
    _, idx_ref = cport.warp3d_fwd(df.numpy(), img.numpy(), mode=cport.CUDA_RCP, want_idx=True)
    assert np.array_equal(idx.cpu().numpy(), idx_ref)


# ----------------------------------------------------------------------------- VecInt (a3)
@pytest.mark.parametrize("case", ["vecint_small", "vecint_large_disp"])
def test_vecint_golden(PF, case):
    """CPU_EXACT mode: bit-identical to the reference (torch-CPU) forward."""
    g = load_golden(case)
    vec = dev(g["vec"], True)
    out = PF.vecint(vec, 7, PF.CPU_EXACT)
    assert np.array_equal(out.detach().cpu().numpy(), g["out"]), \
        "vecint fwd not bit-identical: max-abs %.3e" % np.abs(out.detach().cpu().numpy() - g["out"]).max()
    out.backward(dev(g["gout"]))
    assert_grad_close(vec.grad.cpu().numpy(), g["gvec"], case + " gvec")
    with torch.no_grad():   # ping-pong (no saved states) variant
        out2 = PF.vecint(dev(g["vec"]), 7, PF.CPU_EXACT)
    assert np.array_equal(out2.cpu().numpy(), g["out"])


@pytest.mark.parametrize("case", ["vecint_small", "vecint_large_disp"])
@pytest.mark.parametrize("mode", ["FAST", "CUDA_RCP"])
def test_vecint_golden_lean_modes(PF, case, mode):
    """Default FAST mode (one-FMA sample position, FMA interpolation) and the torch-CUDA rounding:
    within the north_star field tolerance of the reference, gradients within the gradient tolerance."""
    g = load_golden(case)
    vec = dev(g["vec"], True)
    out = PF.vecint(vec, 7, getattr(PF, mode))
    assert_close(out.detach().cpu().numpy(), g["out"], FIELD_ATOL, case + " out " + mode)
    out.backward(dev(g["gout"]))
    assert_grad_close(vec.grad.cpu().numpy(), g["gvec"], case + " gvec " + mode)
    with torch.no_grad():
        out2 = PF.vecint(dev(g["vec"]), 7, getattr(PF, mode))
    assert torch.equal(out2, out.detach())


def test_vecint_scatter_variants_agree_at_level0_size(PF):
    """Full config-2 level-0 size (80x96x112): lane/plane-combined scatter vs one reduction per corner,
    and FAST vs CPU_EXACT, through the C ABI (size-independent property: same adjoint)."""
    import ctypes
    from pulpo_b200 import _lib, synthetic as syn
    L = _lib.lib()
    shape = (80, 96, 112)
    v = syn.make_field(shape, 3, max_abs=3.0).cuda()
    gout = syn.make_field(shape, 4, max_abs=1.0).cuda()
    B = 1
    res = {}
    for name, fmode, bmode in [("exact", 0, 0), ("fast", 2, 2), ("fast_naive", 2, 2 | 0x100)]:
        ws = torch.empty(L.pulpo_vecint_ws_bytes(7, 1, B, *shape) // 4, device="cuda")
        scr = torch.empty(L.pulpo_vecint_bwd_scratch_bytes(B, *shape) // 4, device="cuda")
        out, gv = torch.empty_like(v), torch.empty_like(v)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(L.pulpo_vecint_fwd(P(v), P(out), P(ws), ws.numel() * 4, 7, 1, B, *shape, fmode, st))
        _lib.check(L.pulpo_vecint_bwd(P(gout), P(ws), P(gv), P(scr), scr.numel() * 4, 7, B, *shape, bmode, st))
        torch.cuda.synchronize()
        res[name] = (out.cpu().numpy(), gv.cpu().numpy())
    assert_close(res["fast"][0], res["exact"][0], FIELD_ATOL, "fast vs exact out")
    assert_grad_close(res["fast"][1], res["exact"][1], "fast vs exact gvec")
    assert_grad_close(res["fast_naive"][1], res["fast"][1], "combined vs per-corner scatter")


def test_vecint_multi_level_launch_matches_per_level_calls(PF):
    """pulpo_vecint_multi_fwd/bwd (all pyramid levels in one cooperative launch) against one launch per
    level: same kernels, same arithmetic -> bit-identical forward, gradients within the atomics tolerance."""
    import ctypes
    from pulpo_b200 import _lib, synthetic as syn
    L = _lib.lib()
    shapes, B, n = [(20, 24, 28), (10, 12, 14), (5, 6, 7)], 2, 7
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    v = [syn.make_field(sh, 3 + i, batch=B, max_abs=3.0).cuda() for i, sh in enumerate(shapes)]
    g = [syn.make_field(sh, 9 + i, batch=B, max_abs=1.0).cuda() for i, sh in enumerate(shapes)]
    single_out = [PF.vecint(t.clone().requires_grad_(True), n) for t in v]
    ws = [torch.empty(L.pulpo_vecint_ws_bytes(n, 1, B, *sh) // 4, device="cuda") for sh in shapes]
    scr = [torch.empty(L.pulpo_vecint_bwd_scratch_bytes(B, *sh) // 4, device="cuda") for sh in shapes]
    out = [torch.empty_like(t) for t in v]
    gv = [torch.empty_like(t) for t in v]
    arr = (_lib.VecIntLevel * len(shapes))()
    for i, sh in enumerate(shapes):
        arr[i] = _lib.VecIntLevel(v[i].data_ptr(), out[i].data_ptr(), ws[i].data_ptr(), ws[i].numel() * 4,
                                  scr[i].data_ptr(), scr[i].numel() * 4, *sh)
    _lib.check(L.pulpo_vecint_multi_fwd(arr, len(shapes), n, 1, B, 0, st))
    for i in range(len(shapes)):
        arr[i].inp, arr[i].out = g[i].data_ptr(), gv[i].data_ptr()
    _lib.check(L.pulpo_vecint_multi_bwd(arr, len(shapes), n, B, 0, st))
    torch.cuda.synchronize()
    for i in range(len(shapes)):
        assert torch.equal(out[i], single_out[i].detach()), "level %d forward differs" % i
        (gref,) = torch.autograd.grad(single_out[i], single_out[i].grad_fn.next_functions[0][0].variable, g[i])
        assert_grad_close(gv[i].cpu().numpy(), gref.cpu().numpy(), "multi gvec level %d" % i)
    assert L.pulpo_vecint_multi_fwd(arr, 7, n, 1, B, 0, st) == -2     # more than 6 levels


@pytest.mark.parametrize("shapes,B", [([(16, 24, 32), (8, 12, 16), (4, 6, 8)], 2), ([(80, 96, 112), (40, 48, 56), (20, 24, 28), (10, 12, 14)], 1),
                                      ([(12, 20, 36), (6, 10, 18)], 1)])
def test_combine_vecint_multi_matches_separate_launches(PF, shapes, B):
    """pulpo_combine_vecint_multi_fwd/bwd (pyramid combination + its adjoint inside the integration launches)
    against the separate x2 up-sampling launches around pulpo_vecint_multi_*: combined fields bit-identical
    (same taps, weights and nesting order), integrated fields bit-identical, gradients within the tolerance."""
    import ctypes
    from pulpo_b200 import _lib, synthetic as syn
    L = _lib.lib()
    n, nl = 7, len(shapes)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    df = [syn.make_field(sh, 3 + i, batch=B, max_abs=3.0).cuda() for i, sh in enumerate(shapes)]
    g = [syn.make_field(sh, 9 + i, batch=B, max_abs=1.0).cuda() for i, sh in enumerate(shapes)]
    ws = [torch.empty(L.pulpo_vecint_ws_bytes(n, 1, B, *sh) // 4, device="cuda") for sh in shapes]
    scr = [torch.empty(L.pulpo_vecint_bwd_scratch_bytes(B, *sh) // 4, device="cuda") for sh in shapes]

    def run(fused):
        comb = [torch.full_like(t, float("nan")) for t in df[:-1]] + [df[-1]]
        out = [torch.empty_like(t) for t in df]
        gv = [torch.empty_like(t) for t in df]
        arr = (_lib.VecIntLevel * nl)()
        for i, sh in enumerate(shapes):
            arr[i] = _lib.VecIntLevel(comb[i].data_ptr(), out[i].data_ptr(), ws[i].data_ptr(), ws[i].numel() * 4,
                                      scr[i].data_ptr(), scr[i].numel() * 4, *sh)
        if fused:
            indiv = (ctypes.c_void_p * nl)(*[t.data_ptr() for t in df])
            _lib.check(L.pulpo_combine_vecint_multi_fwd(arr, indiv, nl, n, 1, B, 0, st))
        else:
            for l in range(nl - 2, -1, -1):
                _lib.check(L.pulpo_resize_up_fwd(vp(comb[l + 1]), vp(df[l]), vp(comb[l]), 2, 2.0, B, 3, *shapes[l + 1], st))
            _lib.check(L.pulpo_vecint_multi_fwd(arr, nl, n, 1, B, 0, st))
        for i in range(nl):
            arr[i].inp, arr[i].out = g[i].data_ptr(), gv[i].data_ptr()
        if fused:
            _lib.check(L.pulpo_combine_vecint_multi_bwd(arr, nl, n, B, 0, st))
        else:
            _lib.check(L.pulpo_vecint_multi_bwd(arr, nl, n, B, 0, st))
            for l in range(1, nl):
                _lib.check(L.pulpo_resize_up_bwd(vp(gv[l - 1]), vp(gv[l]), 2, 2.0, 1, B, 3, *shapes[l], st))
        torch.cuda.synchronize()
        return comb, out, gv

    c0, o0, g0 = run(False)
    c1, o1, g1 = run(True)
    for i in range(nl):
        assert torch.equal(c1[i], c0[i]), "combined field of level %d differs (max %.3e)" % (i, (c1[i] - c0[i]).abs().max().item())
        assert torch.equal(o1[i], o0[i]), "integrated field of level %d differs" % i
        assert_grad_close(g1[i].cpu().numpy(), g0[i].cpu().numpy(), "gradient of level %d" % i)
    # not a x2 pyramid -> rejected before any launch
    arr = (_lib.VecIntLevel * 2)()
    for i, sh in enumerate([shapes[0], (shapes[0][0] // 2, shapes[0][1] // 2, shapes[0][2] // 2 - 1)]):
        t = torch.empty(B, 3, *sh, device="cuda")
        arr[i] = _lib.VecIntLevel(t.data_ptr(), t.data_ptr(), ws[0].data_ptr(), ws[0].numel() * 4, scr[0].data_ptr(),
                                  scr[0].numel() * 4, *sh)
    indiv = (ctypes.c_void_p * 2)(df[0].data_ptr(), df[1].data_ptr())
    assert L.pulpo_combine_vecint_multi_fwd(arr, indiv, 2, n, 1, B, 0, st) == -2


def test_vecint_large_batch_runs_per_item_and_matches_single_items(PF):
    """Batches of large volumes are integrated item by item (L2-resident states; opaque per-item layout of the
    saved states): forward bit-identical to B=1 calls, backward within the gradient tolerance -- through the
    autograd wrapper (single level) and the joint multi-level launch."""
    import ctypes
    from pulpo_b200 import _lib, synthetic as syn
    L = _lib.lib()
    n, B = 7, 2
    shapes = [(80, 96, 112), (40, 48, 56)]
    v = [syn.make_field(sh, 3 + i, batch=B, max_abs=3.0).cuda() for i, sh in enumerate(shapes)]
    g = [syn.make_field(sh, 9 + i, batch=B, max_abs=1.0).cuda() for i, sh in enumerate(shapes)]
    # autograd wrapper, level 0 alone (2 x 860160 voxels -> split)
    vb = v[0].clone().requires_grad_(True)
    ob = PF.vecint(vb, n)
    ob.backward(g[0])
    for b in range(B):
        v1 = v[0][b:b + 1].clone().requires_grad_(True)
        o1 = PF.vecint(v1, n)
        o1.backward(g[0][b:b + 1])
        assert torch.equal(ob[b:b + 1].detach(), o1.detach()), "item %d forward differs" % b
        assert_grad_close(vb.grad[b:b + 1].cpu().numpy(), v1.grad.cpu().numpy(), "item %d gradient" % b)
    with torch.no_grad():
        assert torch.equal(PF.vecint(v[0], n), ob.detach())       # two-state (no saved steps) path
    # joint two-level launch
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ws = [torch.empty(L.pulpo_vecint_ws_bytes(n, 1, B, *sh) // 4, device="cuda") for sh in shapes]
    scr = [torch.empty(L.pulpo_vecint_bwd_scratch_bytes(B, *sh) // 4, device="cuda") for sh in shapes]
    out = [torch.empty_like(t) for t in v]
    gv = [torch.empty_like(t) for t in v]
    arr = (_lib.VecIntLevel * 2)()
    for i, sh in enumerate(shapes):
        arr[i] = _lib.VecIntLevel(v[i].data_ptr(), out[i].data_ptr(), ws[i].data_ptr(), ws[i].numel() * 4,
                                  scr[i].data_ptr(), scr[i].numel() * 4, *sh)
    _lib.check(L.pulpo_vecint_multi_fwd(arr, 2, n, 1, B, 0, st))
    for i in range(2):
        arr[i].inp, arr[i].out = g[i].data_ptr(), gv[i].data_ptr()
    _lib.check(L.pulpo_vecint_multi_bwd(arr, 2, n, B, 0, st))
    torch.cuda.synchronize()
    assert torch.equal(out[0], ob.detach())
    assert_grad_close(gv[0].cpu().numpy(), vb.grad.cpu().numpy(), "joint launch gradient, level 0")
    v1 = v[1].clone().requires_grad_(True)
    o1 = PF.vecint(v1, n)          # level 1 alone is below the split threshold: one joint launch
    o1.backward(g[1])
    assert torch.equal(out[1], o1.detach())
    assert_grad_close(gv[1].cpu().numpy(), v1.grad.cpu().numpy(), "joint launch gradient, level 1")
    # a workspace sized for one item only is refused
    arr[0].inp, arr[0].out = v[0].data_ptr(), out[0].data_ptr()
    arr[0].ws_bytes = ws[0].numel() * 4 // 2
    assert L.pulpo_vecint_multi_fwd(arr, 1, n, 1, B, 0, st) == -4


def test_vecint_zero_steps_and_zero_field(PF):
    v = torch.randn(1, 3, 6, 8, 10, device="cuda")
    assert torch.equal(PF.vecint(v, 0), v)
    z = torch.zeros(1, 3, 6, 8, 12, device="cuda")
    out = PF.vecint(z, 7)
    # zero velocity is NOT the identity in the reference (p = v*S/(S-1) - 0.5), but v=0 stays 0
    assert torch.equal(out, z)


# ----------------------------------------------------------------------------- resize / combine (a4, a5)
def test_combine_up2_golden(PF):
    g = load_golden("combine_up2")
    lo, ind = dev(g["lower"], True), dev(g["indiv"], True)
    out = PF.resize_up(lo, 2, 2.0, addend=ind)
    assert_close(out.detach().cpu().numpy(), g["out"], FIELD_ATOL, "combine out")
    out.backward(dev(g["gout"]))
    assert_grad_close(lo.grad.cpu().numpy(), g["glower"], "combine glower")
    assert_grad_close(ind.grad.cpu().numpy(), g["gindiv"], "combine gindiv")


@pytest.mark.parametrize("f", [2, 4, 8, 16])
def test_resize_up_golden(PF, f):
    from pulpo_b200.network_blocks import ResizeTransform
    g = load_golden("resize_up%d" % f)
    x = dev(g["x"], True)
    out = ResizeTransform(1 / f, 3)(x)
    assert_close(out.detach().cpu().numpy(), g["out"], FIELD_ATOL, "resize out")
    out.backward(dev(g["gout"]))
    assert_grad_close(x.grad.cpu().numpy(), g["gx"], "resize gx")


def test_resize_factor_one_is_noop(PF):
    from pulpo_b200.network_blocks import ResizeTransform
    x = torch.randn(1, 3, 4, 5, 6, device="cuda")
    assert ResizeTransform(1.0, 3)(x) is x


# ----------------------------------------------------------------------------- pyramids (a8, a10)
def test_target_pyramid_golden(PF):
    g = load_golden("target_pyramid")
    y = dev(g["y"])
    for i in range(5):
        out = PF.interp_to_size(y, tuple(int(s) for s in g["size%d" % i]))
        assert_close(out.cpu().numpy(), g["out%d" % i], 1e-6, "target pyramid %d" % i)


@pytest.mark.parametrize("shape,nl,B", [((160, 192, 224), 4, 1), ((64, 64, 64), 3, 2), ((16, 32, 48), 4, 1), ((8, 12, 20), 2, 3),
                                        ((8, 4, 12), 1, 1)])
def test_avgpool2_pyramid_equals_chain(PF, shape, nl, B):
    """pulpo_avgpool2_pyramid_fwd (all pooling levels in one launch, input read once) is bit-identical to the
    chain of pulpo_avgpool2_fwd launches; shapes that are not multiples of 2^levels are refused."""
    import ctypes
    from pulpo_b200 import _lib
    L = _lib.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    x = torch.rand(B, 1, *shape, device="cuda")
    chain, src, sh = [], x, shape
    for _ in range(nl):
        o = torch.empty(B, 1, *[(v + 1) // 2 for v in sh], device="cuda")
        _lib.check(L.pulpo_avgpool2_fwd(vp(src), vp(o), B, 1, *sh, st))
        chain.append(o)
        src, sh = o, tuple(o.shape[2:])
    outs = [torch.full_like(t, float("nan")) for t in chain]
    arr = (ctypes.c_void_p * nl)(*[t.data_ptr() for t in outs])
    _lib.check(L.pulpo_avgpool2_pyramid_fwd(vp(x), arr, nl, B, 1, *shape, st))
    torch.cuda.synchronize()
    for a, b in zip(outs, chain):
        assert torch.equal(a, b)
    assert torch.equal(outs[0], torch.nn.functional.avg_pool3d(x, 2, 2, ceil_mode=True))
    assert L.pulpo_avgpool2_pyramid_fwd(vp(x), arr, nl, B, 1, shape[0], shape[1], shape[2] - 2, st) == -3


def test_avgpool2_golden(PF):
    g = load_golden("avgpool2")
    assert_close(PF.avgpool2(dev(g["x_even"])).cpu().numpy(), g["out_even"], 1e-6, "avgpool even")
    assert_close(PF.avgpool2(dev(g["x_odd"])).cpu().numpy(), g["out_odd"], 1e-6, "avgpool odd (ceil_mode)")


# ----------------------------------------------------------------------------- NCC (a9)
@pytest.mark.parametrize("win", [9, 7, 5, 3])
def test_ncc_golden(PF, win):
    g = load_golden("ncc")
    pred, target = dev(g["pred"], True), dev(g["target"])
    loss = PF.ncc_loss(pred, target, win, 0.05)
    assert_loss_close(loss.item(), g["loss_w%d" % win], "ncc loss w=%d" % win)
    loss.backward()
    assert_grad_close(pred.grad.cpu().numpy(), g["gpred_w%d" % win], "ncc grad w=%d" % win)


def test_ncc_vs_oracle_midsize_ragged(PF):
    from oracle import cport
    from pulpo_b200 import synthetic as syn
    # sizes that are not multiples of the CTA tile (32 x 16) and span several z-chunks
    x, y = syn.make_pair((70, 45, 83), 5, batch=1)
    for win in (9, 3):
        p = x.cuda().requires_grad_(True)
        loss = PF.ncc_loss(p, y.cuda(), win, 0.05)
        loss.backward()
        ref, gref = cport.ncc(x.numpy(), y.numpy(), win, 0.05, want_grad=True)
        assert_loss_close(loss.item(), ref, "ncc ragged w=%d" % win)
        assert_grad_close(p.grad.cpu().numpy(), gref, "ncc ragged grad w=%d" % win)


@pytest.mark.parametrize("shape,win,batch", [((40, 48, 56), 5, 1), ((33, 50, 60), 3, 2), ((30, 41, 44), 9, 1),
                                             ((47, 70, 100), 7, 1), ((64, 64, 64), 9, 1)])
def test_ncc_tma_path_vs_oracle(PF, shape, win, batch):
    """Volumes that take the persistent TMA kernel (D2 % 4 == 0, D2 >= 44): ragged tiles in D1 and D2,
    CTAs whose z range crosses a column boundary, batch > 1, every window size."""
    from oracle import cport
    from pulpo_b200 import synthetic as syn
    x, y = syn.make_pair(shape, 7, batch=batch)
    p = x.cuda().requires_grad_(True)
    loss = PF.ncc_loss(p, y.cuda(), win, 0.05)
    loss.backward()
    ref, gref = cport.ncc(x.numpy(), y.numpy(), win, 0.05, want_grad=True)
    assert_loss_close(loss.item(), ref, "ncc tma %s w=%d" % (shape, win))
    assert_grad_close(p.grad.cpu().numpy(), gref, "ncc tma grad %s w=%d" % (shape, win))
    with torch.no_grad():   # no coefficient volumes saved
        assert_loss_close(PF.ncc_loss(x.cuda(), y.cuda(), win, 0.05).item(), ref, "ncc tma no-grad")


def test_ncc_full_size_properties(PF):
    """BASELINE config-2 size (160x192x224, win 9), where the oracle is too slow: size-independent
    properties -- symmetry in the two images, linearity in gamma, NCC(x, x) below NCC(x, y) < 0, and
    the backward equals gamma-scaled backward (closed form is linear in the upstream scale)."""
    from pulpo_b200 import synthetic as syn
    x, y = (t.cuda() for t in syn.make_pair((160, 192, 224), 1))
    p = x.clone().requires_grad_(True)
    lxy = PF.ncc_loss(p, y, 9, 0.05)
    lxy.backward()
    g1 = p.grad.clone()
    assert_loss_close(lxy.item(), PF.ncc_loss(y, x, 9, 0.05).item(), "ncc symmetry")
    p.grad = None
    l2 = PF.ncc_loss(p, y, 9, 0.1)
    l2.backward()
    assert_loss_close(l2.item(), 2.0 * lxy.item(), "ncc linear in gamma")
    assert_grad_close(p.grad.cpu().numpy(), 2.0 * g1.cpu().numpy(), "ncc grad linear in gamma")
    assert PF.ncc_loss(x, x, 9, 0.05).item() < lxy.item() < 0.0


# ----------------------------------------------------------------------------- KL (a11) / L2 (f-1)
def test_kl_golden(PF):
    g = load_golden("kl_diag")
    for tag, m1, s1 in [("std", np.zeros_like(g["mu0"]), np.ones_like(g["sigma0"])), ("gen", g["mu1"], g["sigma1"])]:
        mu, sg = dev(g["mu0"], True), dev(g["sigma0"], True)
        kl = PF.kl_diag(mu, sg, dev(m1), dev(s1))
        assert_loss_close(kl.item(), g["kl_" + tag], "kl " + tag)
        kl.backward()
        assert_grad_close(mu.grad.cpu().numpy(), g["gmu0_" + tag], "kl gmu " + tag)
        assert_grad_close(sg.grad.cpu().numpy(), g["gsigma0_" + tag], "kl gsigma " + tag)
    # N(0,1) fast path through the expanded-constant prior
    from pulpo_b200.components.pulpo import PULPoPrior
    mu, sg = dev(g["mu0"]), dev(g["sigma0"])
    pm, ps = PULPoPrior()({0: mu}, {0: sg})
    assert_loss_close(PF.kl_diag(mu, sg, pm[0], ps[0]).item(), g["kl_std"], "kl fast path")


def test_gauss_sample_kl_fused(PF):
    """f-4: gauss_sampler fused with the level's KL == the two separate reference expressions
    (src/network_blocks.py:7-8, src/losses.py:47-76) on the golden KL inputs: z bitwise for the same
    noise, KL and the joint gradient within the loss / gradient tolerances; ragged and aligned sizes."""
    from pulpo_b200 import network_blocks as NB
    g = load_golden("kl_diag")
    for tag, m1, s1 in [("std", None, None), ("gen", g["mu1"], g["sigma1"])]:
        mu, sg = dev(g["mu0"], True), dev(g["sigma0"], True)
        gen = torch.Generator(device="cuda").manual_seed(5)
        noise = torch.randn(sg.shape, dtype=torch.float32, device="cuda", generator=gen)
        gen.manual_seed(5)
        z, kl = NB.gauss_sampler_kl(mu, sg, var=1, generator=gen, prior_mu=None if m1 is None else dev(m1),
                                    prior_sigma=None if s1 is None else dev(s1))
        z_ref = mu.detach() + sg.detach() * (1 * noise)
        assert torch.equal(z, z_ref), "fused sample differs from mu + sigma * (var * noise)"
        assert_loss_close(kl.item(), g["kl_" + tag], "fused kl " + tag)
        w = torch.randn(z.shape, device="cuda", generator=gen)
        ((z * w).sum() + kl).backward()
        assert_grad_close(mu.grad.cpu().numpy(), g["gmu0_" + tag] + w.cpu().numpy(), "fused gmu " + tag)
        assert_grad_close(sg.grad.cpu().numpy(), g["gsigma0_" + tag] + (w * noise).cpu().numpy(), "fused gsigma " + tag)
    # only one of the two outputs used; ragged length (scalar path); var != 1
    for shape in [(2, 3, 5, 7, 9), (1, 3, 8, 8, 8)]:
        gen = torch.Generator(device="cuda").manual_seed(1)
        mu0 = torch.randn(shape, device="cuda", generator=gen)
        sg0 = torch.rand(shape, device="cuda", generator=gen) + 0.1
        noise = torch.randn(shape, device="cuda", generator=gen)
        mu, sg = mu0.clone().requires_grad_(True), sg0.clone().requires_grad_(True)
        z, kl = PF.gauss_sample_kl(mu, sg, noise, var=2)
        assert torch.equal(z, mu0 + sg0 * (2 * noise))
        kl_sep = PF.kl_diag(mu0, sg0, None, None)
        assert_loss_close(kl.item(), kl_sep.item(), "fused kl vs separate")
        z.sum().backward()
        assert torch.equal(mu.grad, torch.ones_like(mu0)) and torch.equal(sg.grad, 2 * noise)
        mu2, sg2 = mu0.clone().requires_grad_(True), sg0.clone().requires_grad_(True)
        mu3, sg3 = mu0.clone().requires_grad_(True), sg0.clone().requires_grad_(True)
        PF.gauss_sample_kl(mu2, sg2, noise)[1].backward()
        PF.kl_diag(mu3, sg3, None, None).backward()
        torch.testing.assert_close(mu2.grad, mu3.grad, rtol=1e-6, atol=0)
        torch.testing.assert_close(sg2.grad, sg3.grad, rtol=1e-6, atol=1e-7)
    with torch.no_grad():
        z, kl = PF.gauss_sample_kl(mu0, sg0, noise)
        assert not z.requires_grad and torch.isfinite(kl)


def test_l2reg_golden(PF):
    g = load_golden("l2reg")
    f = dev(g["f"], True)
    loss = PF.l2_reg(f, 0.025)
    assert_loss_close(loss.item(), g["loss"], "l2reg")
    loss.backward()
    assert_grad_close(f.grad.cpu().numpy(), g["gf"], "l2reg grad")


def test_warp_l2reg_fused_matches_separate_kernels(PF):
    """pulpo_warp3d_l2reg_fwd/bwd (warp + L2_reg in one pass over the field) against the
    golden L2_reg fixture and the separate warp kernels, ragged and vectorisable shapes."""
    import ctypes
    from pulpo_b200 import _lib, synthetic as syn
    L = _lib.lib()
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = load_golden("l2reg")
    cases = [(torch.from_numpy(g["f"]).cuda(), float(g["loss"]), g["gf"])]
    for shape, seed in [((12, 20, 28), 1), ((9, 11, 13), 2), ((16, 24, 32), 3)]:
        f = syn.make_field(shape, seed, batch=2, max_abs=2.5).cuda().requires_grad_(True)
        ref = PF.l2_reg(f, 0.025)
        ref.backward()
        cases.append((f.detach(), ref.item(), f.grad.cpu().numpy()))
    for f, loss_ref, gf_ref in cases:
        B, _, D0, D1, D2 = f.shape
        img = torch.rand(B, 2, D0, D1, D2, device="cuda")
        gout = torch.randn(B, 2, D0, D1, D2, device="cuda")
        out_ref = PF.warp(f, img)
        fr = f.clone().requires_grad_(True)
        PF.warp(fr, img).backward(gout)
        out, reg = torch.empty_like(img), torch.zeros((), device="cuda")
        ws = torch.zeros(L.pulpo_reduce_ws_bytes(), dtype=torch.uint8, device="cuda")
        for _ in range(2):   # twice: the workspace must reset itself
            _lib.check(L.pulpo_warp3d_l2reg_fwd(vp(img), vp(f), vp(out), 0.025, vp(reg), vp(ws), ws.numel(), B, 2, D0, D1, D2, 0, st))
        assert torch.equal(out, out_ref)
        assert_loss_close(reg.item(), loss_ref, "fused l2reg")
        gdf = torch.empty_like(f)
        _lib.check(L.pulpo_warp3d_l2reg_bwd(vp(gout), vp(img), vp(f), vp(gdf), 0.025, None, B, 2, D0, D1, D2, 0, st))
        want = fr.grad.cpu().numpy() + gf_ref
        assert_grad_close(gdf.cpu().numpy(), want, "fused warp+l2reg grad")


# ----------------------------------------------------------------------------- decoder chain + losses (a6, a7, a10, a12)
def test_hot_path_golden(PF):
    from pulpo_b200.models import RegistrationHotPath
    g = load_golden("hot_path_3lvl")
    total, latent = int(g["total_levels"]), int(g["latent_levels"])
    hp = RegistrationHotPath(list(g["x"].shape[2:]), total, latent, beta=0.1, gamma=0.05, lamb=0.025).cuda()
    dfs = {l: dev(g["df%d" % l], True) for l in range(latent)}
    mus = {l: dev(g["mu%d" % l], True) for l in range(latent)}
    sgs = {l: dev(g["sigma%d" % l], True) for l in range(latent)}
    loss, parts, outs = hp(dev(g["x"]), dev(g["y"]), dfs, mus, sgs)
    for l in range(latent):
        assert_close(outs["combined"][l].detach().cpu().numpy(), g["combined%d" % l], FIELD_ATOL, "combined %d" % l)
        assert_close(outs["final"][l].detach().cpu().numpy(), g["final%d" % l], FIELD_ATOL, "final %d" % l)
        assert_close(outs["moved"][l].detach().cpu().numpy(), g["moved%d" % l], FIELD_ATOL, "moved %d" % l)
        assert_loss_close(parts["kl_levels"][l].item(), g["kl_level%d" % l], "kl level %d" % l)  # fixture holds w*KL before beta
        assert_loss_close(parts["recon_levels"][l].item(), g["recon_level%d" % l], "recon level %d" % l)
        assert_loss_close(parts["reg_levels"][l].item(), g["reg_level%d" % l], "reg level %d" % l)
    assert_loss_close(parts["kl"].item(), g["kl"], "kl")
    assert_loss_close(parts["recon"].item(), g["recon"], "recon")
    assert_loss_close(parts["reg"].item(), g["reg"], "reg")
    assert_loss_close(loss.item(), g["total"], "total")
    loss.backward()
    for l in range(latent):
        assert_grad_close(dfs[l].grad.cpu().numpy(), g["gdf%d" % l], "gdf %d" % l)
        assert_grad_close(mus[l].grad.cpu().numpy(), g["gmu%d" % l], "gmu %d" % l)
        assert_grad_close(sgs[l].grad.cpu().numpy(), g["gsigma%d" % l], "gsigma %d" % l)


def test_combine_dfs_golden(PF):
    from pulpo_b200.models import combine_dfs
    g = load_golden("combine_dfs")
    dfs = {l: dev(g["df%d" % l]) for l in range(2)}
    with torch.no_grad():
        comb, fin = combine_dfs(dfs, [int(s) for s in g["input_size"]])
    for l in range(2):
        assert_close(comb[l].cpu().numpy(), g["combined%d" % l], FIELD_ATOL, "combined %d" % l)
        assert_close(fin[l].cpu().numpy(), g["final%d" % l], FIELD_ATOL, "final %d" % l)


def test_hot_path_vs_torch_oracle_config1(PF):
    """config 1 of BASELINE.json (64^3, 3 latent / 4 total levels) against the torch-CPU restatement."""
    from oracle import torch_ref as T
    from pulpo_b200 import synthetic as syn
    from pulpo_b200.models import RegistrationHotPath
    size, total, latent = [64, 64, 64], 4, 3
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(size, total, latent, seed=0)
    d_ref = {l: dfs[l].clone().requires_grad_(True) for l in dfs}
    m_ref = {l: mus[l].clone().requires_grad_(True) for l in dfs}
    s_ref = {l: sgs[l].clone().requires_grad_(True) for l in dfs}
    ref_total, ref_parts, ref_out = T.hot_path_losses(x, y, d_ref, m_ref, s_ref, total)
    ref_total.backward()
    hp = RegistrationHotPath(size, total, latent).cuda()
    d = {l: dfs[l].cuda().requires_grad_(True) for l in dfs}
    m = {l: mus[l].cuda().requires_grad_(True) for l in dfs}
    s = {l: sgs[l].cuda().requires_grad_(True) for l in dfs}
    loss, parts, outs = hp(x.cuda(), y.cuda(), d, m, s)
    loss.backward()
    for k in ("kl", "recon", "reg"):
        assert_loss_close(parts[k].item(), ref_parts[k].item(), k)
    assert_loss_close(loss.item(), ref_total.item(), "total")
    for l in range(latent):
        assert_close(outs["moved"][l].detach().cpu().numpy(), ref_out["moved"][l].detach().numpy(), FIELD_ATOL, "moved")
        assert_close(outs["final"][l].detach().cpu().numpy(), ref_out["final"][l].detach().numpy(), FIELD_ATOL, "final")
        assert_grad_close(d[l].grad.cpu().numpy(), d_ref[l].grad.numpy(), "gdf %d" % l)
        assert_grad_close(m[l].grad.cpu().numpy(), m_ref[l].grad.numpy(), "gmu %d" % l)
        assert_grad_close(s[l].grad.cpu().numpy(), s_ref[l].grad.numpy(), "gsigma %d" % l)


# ----------------------------------------------------------------------------- HotPathPlan (multi-stream, graph-capturable)
def _run_plan(g_or_inputs, total, latent, size, B, multi_stream=True, graph=False, **kw):
    from pulpo_b200.plan import HotPathPlan
    x, y, dfs, mus, sgs = g_or_inputs
    plan = HotPathPlan(size, total, latent, batch=B, multi_stream=multi_stream, **kw)
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            plan.run(x, y, dfs, mus, sgs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            plan.run(x, y, dfs, mus, sgs)
        plan.losses.zero_()
        gr.replay()
    else:
        plan.run(x, y, dfs, mus, sgs)
    torch.cuda.synchronize()
    return plan


@pytest.mark.parametrize("multi_stream,graph,fuse_reg,kw", [
    (False, False, True, {}), (True, False, False, {}), (True, True, True, {}),
    (True, True, True, {"fuse_combine": True, "pool_pyramid": False, "aux_early": True}),
    (False, False, True, {"fuse_combine": True}), (True, True, True, {"dpos": False}), (False, False, True, {"dpos": False}),
    (True, True, True, {"aux_after": "up2"}), (True, True, True, {"reg_coarse": True}), (False, False, True, {"reg_coarse": True}), (True, True, True, {"aux_after": "warp"}), (False, False, True, {"aux_after": "warp"})])
def test_plan_golden(PF, multi_stream, graph, fuse_reg, kw):
    g = load_golden("hot_path_3lvl")
    total, latent = int(g["total_levels"]), int(g["latent_levels"])
    size, B = list(g["x"].shape[2:]), g["x"].shape[0]
    inputs = (dev(g["x"]), dev(g["y"]), {l: dev(g["df%d" % l]) for l in range(latent)},
              {l: dev(g["mu%d" % l]) for l in range(latent)}, {l: dev(g["sigma%d" % l]) for l in range(latent)})
    plan = _run_plan(inputs, total, latent, size, B, multi_stream, graph, fuse_reg=fuse_reg, **kw)
    losses = plan.losses.cpu().numpy()
    assert_loss_close(losses[0].sum(), g["kl"], "kl")
    assert_loss_close(losses[1].sum(), g["recon"], "recon")
    assert_loss_close(losses[2].sum(), g["reg"], "reg")
    assert_loss_close(plan.total.item(), g["total"], "total")
    for l in range(latent):
        assert_loss_close(losses[1][l], g["recon_level%d" % l], "recon level %d" % l)
        assert_loss_close(losses[2][l], g["reg_level%d" % l], "reg level %d" % l)
        assert_close(plan.final[l].cpu().numpy(), g["final%d" % l], FIELD_ATOL, "final %d" % l)
        assert_close(plan.moved[l].cpu().numpy(), g["moved%d" % l], FIELD_ATOL, "moved %d" % l)
        assert_grad_close(plan.gdf[l].cpu().numpy(), g["gdf%d" % l], "gdf %d" % l)
        assert_grad_close(plan.gmu[l].cpu().numpy(), g["gmu%d" % l], "gmu %d" % l)
        assert_grad_close(plan.gsigma[l].cpu().numpy(), g["gsigma%d" % l], "gsigma %d" % l)


def test_warp_dpos_forward_and_streaming_backward_match_gather_backward(PF):
    """pulpo_warp3d_fwd_dpos: same moved image bit for bit, and dpos such that gout * dpos is the gather-form backward's
    field gradient (pulpo_warp3d_bwd); pulpo_warp3d_bwd_dpos and pulpo_l2reg_fwd_bwd (product inside the regulariser's
    pass) against the separate kernels.  Ragged and vectorisable shapes, clamped borders (large displacements)."""
    import ctypes
    from pulpo_b200 import _lib, synthetic as syn
    L = _lib.lib()
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for shape, seed, B, amp in [((12, 20, 28), 1, 2, 2.5), ((9, 11, 13), 2, 1, 6.0), ((40, 48, 56), 3, 1, 30.0), ((16, 24, 32), 4, 3, 0.3)]:
        D0, D1, D2 = shape
        f = syn.make_field(shape, seed, batch=B, max_abs=amp).cuda()
        img = torch.rand(B, 1, *shape, device="cuda")
        gout = torch.randn(B, 1, *shape, device="cuda")
        out_ref = PF.warp(f, img)
        fr = f.clone().requires_grad_(True)
        PF.warp(fr, img).backward(gout)
        out, dpos = torch.empty_like(img), torch.full((B, 3, *shape), float("nan"), device="cuda")
        _lib.check(L.pulpo_warp3d_fwd_dpos(vp(img), vp(f), vp(out), vp(dpos), B, D0, D1, D2, 0, st))
        assert torch.equal(out, out_ref)
        assert bool(torch.isfinite(dpos).all())
        assert_grad_close((gout * dpos).cpu().numpy(), fr.grad.cpu().numpy(), "gout * dpos")
        gdf = torch.full_like(f, 7.0)
        _lib.check(L.pulpo_warp3d_bwd_dpos(vp(gout), vp(dpos), vp(gdf), 0, B, D0, D1, D2, st))
        assert torch.equal(gdf, gout * dpos)
        _lib.check(L.pulpo_warp3d_bwd_dpos(vp(gout), vp(dpos), vp(gdf), 1, B, D0, D1, D2, st))
        assert_grad_close(gdf.cpu().numpy(), (2 * gout * dpos).cpu().numpy(), "accumulate")
        # regulariser value + gradient (+ product) in one pass
        freg = f.clone().requires_grad_(True)
        reg_ref = PF.l2_reg(freg, 0.025)
        reg_ref.backward()
        ws = torch.zeros(L.pulpo_reduce_ws_bytes(), dtype=torch.uint8, device="cuda")
        reg = torch.zeros((), device="cuda")
        for with_prod in (False, True, True):   # repeated: the workspace resets itself
            g2 = torch.full_like(f, 3.0)
            _lib.check(L.pulpo_l2reg_fwd_bwd(vp(f), 0.025, vp(reg), vp(gout) if with_prod else None,
                                             vp(dpos) if with_prod else None, vp(g2), 0, vp(ws), ws.numel(), B, 3, D0, D1, D2, st))
            assert_loss_close(reg.item(), reg_ref.item(), "fused l2reg value")
            want = freg.grad + (gout * dpos if with_prod else 0)
            assert_grad_close(g2.cpu().numpy(), want.cpu().numpy(), "fused l2reg grad (+ product)")
        g3 = torch.ones_like(f)
        _lib.check(L.pulpo_l2reg_fwd_bwd(vp(f), 0.025, vp(reg), vp(gout), vp(dpos), vp(g3), 1, vp(ws), ws.numel(), B, 3, D0, D1, D2, st))
        assert_grad_close(g3.cpu().numpy(), (freg.grad + gout * dpos + 1).cpu().numpy(), "fused l2reg grad, accumulate")


def test_l2reg_of_upsampled_field_closed_form_on_coarse_grid(PF):
    """pulpo_l2reg_up2_fwd_bwd: L2_reg(ResizeTransform(1/2)(v)) and its gradient w.r.t. v, evaluated on the coarse grid,
    against the composition of the separate kernels (x2 resize -> L2_reg -> autograd), and the fused resize adjoint
    pulpo_resize_up2_bwd_dpos against product + adjoint.  Ragged / minimal / vectorisable shapes, batches."""
    import ctypes
    from pulpo_b200 import _lib, synthetic as syn
    L = _lib.lib()
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for shape, B, amp in [((6, 8, 10), 2, 3.0), ((2, 2, 2), 1, 1.0), ((3, 5, 4), 1, 40.0), ((20, 24, 28), 1, 9.0), ((40, 48, 56), 1, 21.0), ((1, 3, 2), 2, 2.0)]:
        v = syn.make_field(shape, 7, batch=B, max_abs=amp).cuda()
        vr = v.clone().requires_grad_(True)
        ref = PF.l2_reg(PF.resize_up(vr, 2, 2.0), 0.025)
        ref.backward()
        ws = torch.zeros(L.pulpo_reduce_ws_bytes(), dtype=torch.uint8, device="cuda")
        scr = torch.full((L.pulpo_l2reg_up2_scratch_bytes(B, 3, *shape) // 4,), float("nan"), device="cuda")
        out = torch.zeros((), device="cuda")
        for acc in (0, 0, 1):
            g = torch.full_like(v, 2.0)
            _lib.check(L.pulpo_l2reg_up2_fwd_bwd(vp(v), 0.025, vp(out), vp(g), acc, vp(scr), scr.numel() * 4, vp(ws), ws.numel(),
                                                 B, 3, *shape, st))
            assert_loss_close(out.item(), ref.item(), "coarse-grid l2reg value %s" % (shape,))
            assert_grad_close(g.cpu().numpy(), (vr.grad + (2.0 if acc else 0.0)).cpu().numpy(), "coarse-grid l2reg grad %s" % (shape,))
        if shape[2] % 2 == 0 and shape[0] >= 2:
            full = tuple(2 * s for s in shape)
            gm, dp = torch.randn(B, 1, *full, device="cuda"), torch.randn(B, 3, *full, device="cuda")
            want = torch.empty_like(v)
            prod = (gm * dp).contiguous()
            _lib.check(L.pulpo_resize_up_bwd(vp(prod), vp(want), 2, 2.0, 0, B, 3, *shape, st))
            got = torch.full_like(v, 1.5)
            _lib.check(L.pulpo_resize_up2_bwd_dpos(vp(gm), vp(dp), vp(got), 2.0, 1, B, *shape, st))
            assert_grad_close(got.cpu().numpy(), (want + 1.5).cpu().numpy(), "fused resize adjoint %s" % (shape,))


def test_plan_without_regulariser_dpos_equals_gather_backward(PF):
    """with_reg=False: the stored-dpos backward (pulpo_warp3d_bwd_dpos) against the gather-form warp backward."""
    from pulpo_b200 import synthetic as syn
    size, total, latent = [32, 48, 64], 4, 3
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(size, total, latent, seed=5)
    dev_in = (x.cuda(), y.cuda(), {l: dfs[l].cuda() for l in dfs}, {l: mus[l].cuda() for l in dfs},
              {l: sgs[l].cuda() for l in dfs})
    a = _run_plan(dev_in, total, latent, size, 1, True, True, with_reg=False, dpos=True)
    b = _run_plan(dev_in, total, latent, size, 1, True, False, with_reg=False, dpos=False)
    assert a.dpos and not b.dpos
    assert_loss_close(a.total.item(), b.total.item(), "total")
    for l in range(latent):
        assert torch.equal(a.moved[l], b.moved[l])
        assert_grad_close(a.gdf[l].cpu().numpy(), b.gdf[l].cpu().numpy(), "gdf %d" % l)


def test_plan_full_size_config2_vs_torch_oracle(PF):
    """The benchmarked configuration itself (BASELINE config 2: 160x192x224 pair, 5 total / 4 latent levels,
    bench.py's synthetic inputs, fused regulariser, CUDA graph) against the torch-CPU restatement at FULL size:
    losses, moved images, final fields and every input gradient (a few seconds of CPU time)."""
    from oracle import torch_ref as T
    from pulpo_b200 import synthetic as syn
    size, total, latent = [160, 192, 224], 5, 4
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(size, total, latent, seed=0)
    d_ref = {l: dfs[l].clone().requires_grad_(True) for l in dfs}
    m_ref = {l: mus[l].clone().requires_grad_(True) for l in dfs}
    s_ref = {l: sgs[l].clone().requires_grad_(True) for l in dfs}
    ref_total, ref_parts, ref_out = T.hot_path_losses(x, y, d_ref, m_ref, s_ref, total)
    ref_total.backward()
    plan = _run_plan((x.cuda(), y.cuda(), {l: dfs[l].cuda() for l in dfs}, {l: mus[l].cuda() for l in dfs},
                      {l: sgs[l].cuda() for l in dfs}), total, latent, size, 1, True, True, fuse_reg=True)
    losses = plan.losses.sum(dim=1).cpu().numpy()
    for i, k in enumerate(("kl", "recon", "reg")):
        assert_loss_close(losses[i], ref_parts[k].item(), k)
    assert_loss_close(plan.total.item(), ref_total.item(), "total")
    for l in range(latent):
        assert_close(plan.moved[l].cpu().numpy(), ref_out["moved"][l].detach().numpy(), FIELD_ATOL, "moved %d" % l)
        assert_close(plan.final[l].cpu().numpy(), ref_out["final"][l].detach().numpy(), FIELD_ATOL, "final %d" % l)
        assert_grad_close(plan.gdf[l].cpu().numpy(), d_ref[l].grad.numpy(), "gdf %d" % l)
        assert_grad_close(plan.gmu[l].cpu().numpy(), m_ref[l].grad.numpy(), "gmu %d" % l)
        assert_grad_close(plan.gsigma[l].cpu().numpy(), s_ref[l].grad.numpy(), "gsigma %d" % l)


def test_plan_matches_autograd_modules_config1(PF):
    """The planned multi-stream path and the autograd drop-in modules are the same arithmetic."""
    from pulpo_b200 import synthetic as syn
    from pulpo_b200.models import RegistrationHotPath
    size, total, latent = [64, 64, 64], 4, 3
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(size, total, latent, seed=2)
    xc, yc = x.cuda(), y.cuda()
    d = {l: dfs[l].cuda().requires_grad_(True) for l in dfs}
    m = {l: mus[l].cuda().requires_grad_(True) for l in dfs}
    s = {l: sgs[l].cuda().requires_grad_(True) for l in dfs}
    loss, parts, outs = RegistrationHotPath(size, total, latent).cuda()(xc, yc, d, m, s)
    loss.backward()
    plan = _run_plan((xc, yc, {l: dfs[l].cuda() for l in dfs}, {l: mus[l].cuda() for l in dfs},
                      {l: sgs[l].cuda() for l in dfs}), total, latent, size, 1, True, True)
    assert_loss_close(plan.total.item(), loss.item(), "total")
    for l in range(latent):
        assert torch.equal(plan.moved[l], outs["moved"][l]), "moved %d differs" % l
        assert_grad_close(plan.gdf[l].cpu().numpy(), d[l].grad.cpu().numpy(), "gdf %d" % l)
        assert_grad_close(plan.gmu[l].cpu().numpy(), m[l].grad.cpu().numpy(), "gmu %d" % l)
        assert_grad_close(plan.gsigma[l].cpu().numpy(), s[l].grad.cpu().numpy(), "gsigma %d" % l)


def test_plan_full_res_mode_vs_torch_oracle_and_modules(PF):
    """df_resolution="full_res" (src/components/pulpo.py:146: every level warps / compares at the input size, output
    resize factors 2 / 4 / 8 here and 16 at config 2) through the graph-captured plan: all losses, moved images, final
    fields and input gradients against the torch-CPU restatement, and the same losses from the autograd modules."""
    from oracle import torch_ref as T
    from pulpo_b200 import synthetic as syn
    from pulpo_b200.models import RegistrationHotPath
    size, total, latent = [32, 48, 64], 4, 3
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(size, total, latent, seed=4)
    dev_in = (x.cuda(), y.cuda(), {l: dfs[l].cuda() for l in dfs}, {l: mus[l].cuda() for l in dfs},
              {l: sgs[l].cuda() for l in dfs})
    plan = _run_plan(dev_in, total, latent, size, 1, True, True, df_resolution="full_res")
    assert [plan.ofac[l] for l in range(latent)] == [2, 4, 8]
    d_ref = {l: dfs[l].clone().requires_grad_(True) for l in dfs}
    m_ref = {l: mus[l].clone().requires_grad_(True) for l in dfs}
    s_ref = {l: sgs[l].clone().requires_grad_(True) for l in dfs}
    ref, parts, ref_out = T.hot_path_losses(x, y, d_ref, m_ref, s_ref, total, full_res=True)
    ref.backward()
    assert_loss_close(plan.total.item(), ref.item(), "total (full_res)")
    for l in range(latent):
        assert tuple(plan.moved[l].shape[2:]) == tuple(size)
        assert_close(plan.moved[l].cpu().numpy(), ref_out["moved"][l].detach().numpy(), FIELD_ATOL, "moved %d" % l)
        assert_close(plan.final[l].cpu().numpy(), ref_out["final"][l].detach().numpy(), FIELD_ATOL, "final %d" % l)
        assert_grad_close(plan.gdf[l].cpu().numpy(), d_ref[l].grad.numpy(), "gdf %d" % l)
        assert_grad_close(plan.gmu[l].cpu().numpy(), m_ref[l].grad.numpy(), "gmu %d" % l)
    d = {l: dfs[l].cuda().requires_grad_(True) for l in dfs}
    loss, _, _ = RegistrationHotPath(size, total, latent, df_resolution="full_res").cuda()(dev_in[0], dev_in[1], d, dev_in[3], dev_in[4])
    assert_loss_close(loss.item(), ref.item(), "modules total (full_res)")


def test_warp_image_larger_than_field_golden(PF):
    """A level-sized field resampling a full-resolution image (grid_sample normalises with the field size and
    samples the image at its own size; evaluate.py:198,240,246): forward bit-identical, both gradients."""
    from pulpo_b200.network_blocks import SpatialTransformer
    g = load_golden("warp_img_size")
    df, img = dev(g["df"], True), dev(g["img"], True)
    out = SpatialTransformer(tuple(g["df"].shape[2:]))(df, img)
    assert tuple(out.shape) == tuple(g["out"].shape)
    assert np.array_equal(out.detach().cpu().numpy(), g["out"]), \
        "not bit-identical: max-abs %.3e" % np.abs(out.detach().cpu().numpy() - g["out"]).max()
    out.backward(dev(g["gout"]))
    assert_grad_close(df.grad.cpu().numpy(), g["gdf"], "gdf")
    assert_grad_close(img.grad.cpu().numpy(), g["gimg"], "gimg")


# ----------------------------------------------------------------------------- MC moments (f-3, config 3)
def test_moments_kernels_match_torch_std(PF):
    from pulpo_b200 import mc
    g = torch.Generator(device="cuda").manual_seed(5)
    stack = torch.randn(9, 3, 10, 12, 14, device="cuda", generator=g) * 0.7 + 2.0
    a, b = mc.MCMoments(stack.shape[1:], "cuda"), mc.MCMoments(stack.shape[1:], "cuda")
    for i in range(5):
        a.update(stack[i])
    for i in range(5, 9):
        b.update(stack[i])
    a.merge_state(b.mean, b.m2, b.count)
    assert a.count == 9
    ref = stack.double()
    assert_close(a.mean.cpu().numpy(), ref.mean(0).cpu().numpy(), 1e-5, "mc mean")
    assert_close(a.std().cpu().numpy(), ref.std(0).cpu().numpy(), 1e-5, "mc std")
    assert_close(a.std_channel_mean().cpu().numpy(), ref.std(0).mean(0).cpu().numpy(), 1e-5, "mc std channel mean")


def test_uncertainty_metrics_kernels_vs_reference_golden(PF):
    """f-3: streaming squared-error map + global NCC(var, mse) kernels against the numbers the reference's own
    formulas give on explicit sample stacks (evaluate.py:1534-1545, Evaluate.ncc :334-353; golden uncertainty.npz)."""
    from pulpo_b200 import mc
    g = load_golden("uncertainty")
    moved, y = dev(g["all_moved"]), dev(g["y"])
    mom, sq = mc.MCMoments(moved.shape[1:], "cuda"), mc.MCSqErr(moved.shape[1:], "cuda")
    for i in range(moved.shape[0]):
        mom.update(moved[i])
        sq.update(moved[i], y[0])
    r = mc.uncertainty_metrics(mom, sq)
    assert_close(r["var"].cpu().numpy(), g["var"], 1e-6, "var map")
    assert_close(r["mse"].cpu().numpy(), g["mse"], 1e-6, "mse map")
    assert_loss_close(float(r["ncc"]), float(g["ncc"]), "ncc(var, mse)", rtol=1e-4)
    assert_loss_close(float(r["var_mean"]), float(g["var_mean"]), "var.mean()", rtol=1e-5)
    # the kernel alone on the golden maps: same number as the reference method to fp32 round-off
    out = PF.global_ncc(dev(g["var"]), dev(g["mse"]))
    assert_loss_close(float(out[0]), float(g["ncc"]), "global_ncc kernel", rtol=2e-5)


def test_mc_uncertainty_hot_path_matches_stacked_reference_semantics(PF):
    """Config 3 at small size: N MC samples of (gauss_sampler -> combine -> integrate -> warp), streamed
    through the moments kernels, against torch.std over explicit stacks (evaluate.py:243-251) of the
    oracle's CPU path on the same per-sample noise."""
    from oracle import torch_ref as T
    from pulpo_b200 import mc, synthetic as syn
    from pulpo_b200.models import combine_dfs
    from pulpo_b200.network_blocks import SpatialTransformer, gauss_sampler
    size, total, latent, N = [32, 32, 32], 4, 3, 6
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(size, total, latent, seed=3)
    xc = x.cuda()
    mu = {l: dfs[l].cuda() for l in dfs}          # posterior mean of the velocity field
    sg = {l: (0.3 * sgs[l]).cuda() for l in dfs}
    st = SpatialTransformer(size)

    def sample_fn(i, gen):
        v = {l: gauss_sampler(mu[l], sg[l], generator=gen) for l in mu}
        _, final = combine_dfs(v, size)
        return {"final0": final[0][0], "moved0": st(final[0], xc)[0]}

    res = mc.mc_uncertainty(sample_fn, N, seed0=11)
    # reference semantics on CPU with the very same noise (drawn on the GPU generator, copied)
    finals, moveds = [], []
    for i in range(N):
        gen = mc.sample_generator(11, i, "cuda")
        v = {l: gauss_sampler(mu[l], sg[l], generator=gen).cpu() for l in mu}
        _, final = T.combine_dfs(v, size)
        finals.append(final[0][0])
        moveds.append(T.warp(final[0], x)[0])
    f_std = torch.stack(finals).std(dim=0).mean(dim=0)
    m_std = torch.stack(moveds).std(dim=0).mean(dim=0)
    assert_close(res["final0"].std_channel_mean().cpu().numpy(), f_std.numpy(), FIELD_ATOL, "final df std")
    assert_close(res["moved0"].variance_map().cpu().numpy(), (m_std ** 2).numpy(), FIELD_ATOL, "variance map")


def test_pipeline_matches_plan_across_slots(PF):
    """HotPathPipeline (packed pinned inputs, copy stream, graph per slot): every ticket returns the loss
    the plan computes for that step's inputs, also when slots are reused and results are read late."""
    from pulpo_b200 import synthetic as syn
    from pulpo_b200.pipeline import HotPathPipeline
    size, total, latent = [32, 32, 32], 4, 3
    pipe = HotPathPipeline(size, total, latent)
    inputs = [syn.make_hot_path_inputs(size, total, latent, seed=s) for s in range(5)]
    hbs = [pipe.host_batch().fill(*inp) for inp in inputs]
    ref = []
    for inp in inputs:
        x, y, dfs, mus, sgs = inp
        plan = _run_plan((x.cuda(), y.cuda(), {l: dfs[l].cuda() for l in dfs}, {l: mus[l].cuda() for l in dfs},
                          {l: sgs[l].cuda() for l in dfs}), total, latent, size, 1, True, False)
        ref.append((plan.total.item(), plan.losses.sum(dim=1).cpu().numpy()))
    tickets, got = [], {}
    for k, hb in enumerate(hbs):
        tickets.append(pipe.submit(hb))
        if k >= 1:
            got[k - 1] = pipe.result(tickets[k - 1])
    got[len(hbs) - 1] = pipe.result(tickets[-1])
    for k in range(len(hbs)):
        assert_loss_close(got[k][0], ref[k][0], "pipeline total %d" % k)
        for j in range(3):
            assert_loss_close(got[k][1 + j], ref[k][1][j], "pipeline part %d/%d" % (k, j))
    with pytest.raises(RuntimeError):
        pipe.result(tickets[0])       # slot long reused


# ----------------------------------------------------------------------------- BASELINE full sizes (configs 2, 4, 5)
@pytest.mark.parametrize("shape", [(160, 192, 224)])
def test_vecint_full_resolution_equals_iterated_warp_config4(PF, shape):
    """Config 4 (7-step scaling-and-squaring at full 160x192x224): the cooperative VecInt kernel against
    seven launches of the independent warp kernel (v <- v + warp(v, v), network_blocks.py:173-177).  Both
    follow the CPU sampler op by op, so the forward must agree bit for bit; the backward (VecInt's
    own/scatter ping-pong vs autograd through 7 warp backward launches with their atomics) within the
    gradient tolerance."""
    from pulpo_b200 import synthetic as syn
    v = syn.make_field(shape, 21, max_abs=3.0).cuda()
    g = syn.make_field(shape, 22, max_abs=1.0).cuda()
    a = v.clone().requires_grad_(True)
    out = PF.vecint(a, 7, PF.CPU_EXACT)
    out.backward(g)
    b = v.clone().requires_grad_(True)
    w = b * (1.0 / 128.0)
    for _ in range(7):
        w = w + PF.warp(w, w, PF.CPU_EXACT)
    assert torch.equal(out.detach(), w.detach()), "max-abs %.3e" % float((out.detach() - w.detach()).abs().max())
    w.backward(g)
    assert_grad_close(a.grad.cpu().numpy(), b.grad.cpu().numpy(), "vecint full-res gvec")


def test_warp_full_size_properties_config2(PF):
    """Config-2 size: identity up to the reference's S/(S-1) - 0.5 convention is hard to state, so use
    exact properties: a constant image is reproduced exactly for any field (weights sum to 1 only up to
    rounding -> tolerance), integer sampling indices stay inside the volume, and warping a 3-channel image
    equals warping its channels one by one (bitwise)."""
    from pulpo_b200 import synthetic as syn
    shape = (160, 192, 224)
    df = (syn.make_field(shape, 31, max_abs=40.0)).cuda()
    const = torch.full((1, 1) + shape, 0.625, device="cuda")
    out, idx = PF.warp_indices(df, const)
    assert float((out - 0.625).abs().max()) <= 1e-6
    for a, s in enumerate(shape):
        assert int(idx[:, a].min()) >= 0 and int(idx[:, a].max()) <= s - 1
    img3 = torch.rand((1, 3) + shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    w3 = PF.warp(df, img3)
    for c in range(3):
        assert torch.equal(w3[:, c:c + 1], PF.warp(df, img3[:, c:c + 1].contiguous()))


def test_plan_batch_equals_per_pair_means_config5(PF):
    """Config 5 shards a batch over ranks; every loss term is a batch mean (losses.py:64,134,222), so a
    B=3 step must equal the mean of three B=1 steps, and each pair's gradients are 1/3 of its own."""
    from pulpo_b200 import synthetic as syn
    size, total, latent, B = [32, 32, 32], 4, 3, 3
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(size, total, latent, seed=4, batch=B)
    cu = lambda t: t.cuda()
    big = _run_plan((cu(x), cu(y), {l: cu(dfs[l]) for l in dfs}, {l: cu(mus[l]) for l in dfs}, {l: cu(sgs[l]) for l in dfs}),
                    total, latent, size, B, True, False)
    tot, gd = 0.0, {l: [] for l in dfs}
    for b in range(B):
        sl = lambda t: t[b:b + 1].contiguous().cuda()
        one = _run_plan((sl(x), sl(y), {l: sl(dfs[l]) for l in dfs}, {l: sl(mus[l]) for l in dfs},
                         {l: sl(sgs[l]) for l in dfs}), total, latent, size, 1, True, False)
        tot += one.total.item() / B
        for l in dfs:
            gd[l].append(one.gdf[l].clone() / B)
    assert_loss_close(big.total.item(), tot, "batch total")
    for l in dfs:
        assert_grad_close(big.gdf[l].cpu().numpy(), torch.cat(gd[l]).cpu().numpy(), "batch gdf %d" % l)


# ----------------------------------------------------------------------------- Jacobian determinant (f-2)
@pytest.mark.parametrize("t", ["a", "b"])
def test_jacobian_det_and_jdetstd_golden(PF, t):
    from pulpo_b200 import losses as PL
    g = load_golden("jacdet")
    df = dev(g["df_" + t], True)
    det = PL.jacobian_det(df, normalize=True)
    assert_close(det.detach().cpu().numpy(), g["det_" + t], 1e-5, "jacobian_det")
    loss = PL.JDetStd(df, lamb=0.7)
    assert_loss_close(loss.item(), float(g["jdetstd_" + t]), "JDetStd")
    loss.backward()
    assert_grad_close(df.grad.cpu().numpy(), g["gdf_std_" + t], "JDetStd grad")
    d2 = dev(g["df_" + t], True)
    det2 = PL.jacobian_det(d2, normalize=False)
    assert_close(det2.detach().cpu().numpy(), g["det_nonorm_" + t], 1e-4 * max(1.0, float(np.abs(g["det_nonorm_" + t]).max())),
                 "jacobian_det no-normalize")
    det2.backward(dev(g["gout_" + t]))
    assert_grad_close(d2.grad.cpu().numpy(), g["gdf_nonorm_" + t], "jacobian_det grad")


def test_jacobian_det_full_size_vs_oracle_slices(PF):
    """Config-2 size: the CUDA determinant map against the numpy oracle (vectorised, seconds at this size) on
    three slabs, and the std against numpy in float64."""
    from oracle import jacdet_ref as J
    from pulpo_b200 import functional as F, synthetic as syn
    shape = (160, 192, 224)
    df = syn.make_field(shape, 51, max_abs=6.0)
    det = F.jacobian_det(df.cuda(), True).cpu().numpy()
    ref_full = J.jacobian_det(df.numpy())
    for z0 in (0, 75, 150):   # slabs including both z faces (replication padding)
        assert_close(det[:, z0:z0 + 10], ref_full[:, z0:z0 + 10], 1e-5, "jacdet slab %d" % z0)
    s = F.jdet_std(df.cuda(), 1.0).item()
    assert_loss_close(s, float(ref_full.astype(np.float64).std(ddof=1)), "full-size JDetStd")


# ----------------------------------------------------------------------------- encoder feedback resampling (f-4)
@pytest.mark.parametrize("B", [1, 2])
def test_feedback_resample_cat_matches_torch(PF, B):
    """src/components/pulpo.py:195-206: six coarser-level tensors resized x2 and concatenated; forward and the
    gradients to every item against F.interpolate + torch.cat on the CPU (the reference's own ops)."""
    import torch.nn.functional as F
    from pulpo_b200.components.pulpo import feedback_resample
    g = torch.Generator().manual_seed(3)
    lo, hi = (5, 6, 7), (10, 12, 14)
    chans = [4, 3, 3, 3, 3, 1]      # samples(zdim), velocity_fields, individual, combined, final dfs, transformed
    items = [torch.randn(B, c, *lo, generator=g) for c in chans]
    ref_in = [t.clone().requires_grad_(True) for t in items]
    ref = torch.cat([F.interpolate(t, size=hi, mode="trilinear", align_corners=False) for t in ref_in], dim=1)
    gout = torch.randn(ref.shape, generator=g)
    ref.backward(gout)
    ours_in = [t.cuda().requires_grad_(True) for t in items]
    out = feedback_resample(ours_in, hi)
    assert_close(out.detach().cpu().numpy(), ref.detach().numpy(), 1e-5, "feedback cat")
    out.backward(gout.cuda())
    for a, b in zip(ours_in, ref_in):
        assert_grad_close(a.grad.cpu().numpy(), b.grad.numpy(), "feedback grad")


# ----------------------------------------------------------------------------- autograd contract
def test_modules_under_no_grad_and_anomaly_mode(PF):
    """The drop-ins must work under torch.no_grad() (inference: no saved states / coefficient volumes) and
    under torch.autograd.set_detect_anomaly(True) (the reference trains with it on, src/train.py)."""
    from pulpo_b200 import synthetic as syn
    from pulpo_b200.models import RegistrationHotPath
    size, total, latent = [32, 32, 32], 4, 3
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(size, total, latent, seed=6)
    hp = RegistrationHotPath(size, total, latent).cuda()
    cu = lambda d: {l: d[l].cuda() for l in d}
    with torch.no_grad():
        l0, _, outs0 = hp(x.cuda(), y.cuda(), cu(dfs), cu(mus), cu(sgs))
    assert not l0.requires_grad
    with torch.autograd.set_detect_anomaly(True):
        d = {l: dfs[l].cuda().requires_grad_(True) for l in dfs}
        m = {l: mus[l].cuda().requires_grad_(True) for l in dfs}
        s = {l: sgs[l].cuda().requires_grad_(True) for l in dfs}
        l1, _, outs1 = hp(x.cuda(), y.cuda(), d, m, s)
        l1.backward()
    assert_loss_close(l0.item(), l1.item(), "no_grad vs grad loss")
    for l in range(latent):
        assert torch.equal(outs0["moved"][l], outs1["moved"][l].detach())
        assert torch.isfinite(d[l].grad).all() and torch.isfinite(m[l].grad).all() and torch.isfinite(s[l].grad).all()


def test_transform_segmentation_many_channels_and_image_gradient(PF):
    """PULPo.transform_segmentation warps 36-channel one-hot maps (models.py:370-388): multi-channel warp,
    and the gradient into the moving image (scatter half) against the C oracle."""
    from oracle import cport
    from pulpo_b200 import synthetic as syn
    shape, C = (12, 14, 16), 36
    df = syn.make_field(shape, 61, max_abs=4.0)
    seg = (torch.rand((1, C) + shape, generator=torch.Generator().manual_seed(2)) > 0.7).float()
    ref, _ = cport.warp3d_fwd(df.numpy(), seg.numpy(), want_idx=True)
    img = seg.cuda().requires_grad_(True)
    out = PF.warp(df.cuda(), img)
    assert np.array_equal(out.detach().cpu().numpy(), ref)
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    out.backward(gout.cuda())
    gimg_ref, gdf_ref = cport.warp3d_bwd(gout.numpy(), df.numpy(), seg.numpy())
    assert_grad_close(img.grad.cpu().numpy(), gimg_ref, "segmentation gimg")


# ----------------------------------------------------------------------------- NaN / Inf propagation
# The reference stops training when a loss turns NaN (src/models.py:188-194): the kernels must let non-finite values
# through exactly where torch does -- no silent clamping -- and must not invent them where torch stays finite
# (grid_sample sends a NaN / +-Inf sample position to a border voxel: the output there is finite).
def _same_nonfinite(a, b, what, atol=FIELD_ATOL, superset=False):
    """``superset``: the kernels keep every trilinear footprint 2x2x2 in-bounds -- at a sample position of exactly
    S-1 they read corner S-2 with weight 0 where torch skips the out-of-range corner S -- so a non-finite value at
    S-2 gives 0 * Inf = NaN here and a finite number in torch.  Never fewer non-finite values than torch, and equal
    wherever both are finite; a loss that is NaN in torch is NaN here."""
    a, b = a.detach().cpu().numpy(), b.detach().cpu().numpy()
    if superset:
        assert not np.any(~np.isfinite(b) & np.isfinite(a)), "%s: finite where torch is not" % what
        fin = np.isfinite(a) & np.isfinite(b)
    else:
        assert np.array_equal(np.isnan(a), np.isnan(b)), "%s: NaN pattern differs (%d vs %d NaNs)" % (what, np.isnan(a).sum(), np.isnan(b).sum())
        assert np.array_equal(np.isposinf(a), np.isposinf(b)) and np.array_equal(np.isneginf(a), np.isneginf(b)), what + ": Inf pattern"
        fin = np.isfinite(b)
    assert_close(a[fin], b[fin], atol, what + " (finite part)")


def test_nonfinite_warp_matches_torch(PF):
    from oracle import torch_ref as T
    from pulpo_b200 import synthetic as syn
    shape = (8, 9, 12)
    df = syn.make_field(shape, 3, max_abs=2.0)
    img = syn.make_field(shape, 4, max_abs=1.0, channels=2)
    df[0, 0, 2, 3, 4] = float("nan"); df[0, 1, 5, 5, 5] = float("inf"); df[0, 2, 1, 1, 1] = float("-inf")
    df[0, 2, 6, 2, 7] = 1e30; df[0, 0, 3, 3, 3] = -1e30
    img[0, 0, 4, 4, 6] = float("nan"); img[0, 1, 2, 2, 2] = float("inf")
    ref = T.warp(df, img)
    out = PF.warp(df.cuda(), img.cuda())
    _same_nonfinite(out, ref, "warp", atol=1e-6)
    # sampling indices stay bit-exact for the non-finite positions too (NaN, +Inf -> S-1; -Inf -> 0)
    from oracle import cport
    _, idx = PF.warp_indices(df.cuda(), img.cuda())
    _, idx_ref = cport.warp3d_fwd(df.numpy(), img.numpy(), want_idx=True)
    assert np.array_equal(idx.cpu().numpy(), idx_ref)


def test_nonfinite_vecint_matches_torch(PF):
    from oracle import torch_ref as T
    from pulpo_b200 import synthetic as syn
    vec = syn.make_field((8, 10, 12), 5, max_abs=3.0)
    vec[0, 1, 3, 4, 5] = float("nan")
    _same_nonfinite(PF.vecint(vec.cuda(), 7), T.vecint(vec, 7), "vecint (NaN)", atol=1e-5)
    vec[0, 0, 6, 2, 9] = float("inf")       # sends samples to the border: the zero-weight-corner caveat applies
    _same_nonfinite(PF.vecint(vec.cuda(), 7), T.vecint(vec, 7), "vecint (NaN + Inf)", atol=1e-5, superset=True)


@pytest.mark.parametrize("bad", [float("nan"), float("inf")])
def test_nonfinite_losses_match_torch(PF, bad):
    from oracle import torch_ref as T
    from pulpo_b200 import synthetic as syn
    shape = (12, 16, 48)
    x, y = syn.make_pair(shape, 1)
    xb = x.clone(); xb[0, 0, 5, 6, 7] = bad
    for win in (9, 3):      # TMA-staged kernel (D2 >= 44) and its small-window variant
        got, ref = PF.ncc_loss(xb.cuda(), y.cuda(), win, 0.05), T.ncc_loss(xb, y, win, 0.05)
        assert torch.isnan(ref) and torch.isnan(got.cpu()), "NCC(win %d) must turn NaN like torch (got %r, ref %r)" % (win, got, ref)
    xs = x[:, :, :6, :7, :9].contiguous(); ys = y[:, :, :6, :7, :9].contiguous()   # generic kernel
    xs[0, 0, 1, 2, 3] = bad
    assert torch.isnan(T.ncc_loss(xs, ys, 5, 0.05)) and torch.isnan(PF.ncc_loss(xs.cuda(), ys.cuda(), 5, 0.05).cpu())
    f = syn.make_field((6, 8, 10), 2, max_abs=2.0); f[0, 2, 3, 4, 5] = bad
    got, ref = PF.l2_reg(f.cuda(), 0.025), T.l2_reg(f, 0.025)
    _same_nonfinite(got, ref, "L2_reg")
    mu = syn.make_field((6, 8, 10), 6, max_abs=1.0); sg = mu.abs() + 0.1
    mub = mu.clone(); mub[0, 0, 1, 1, 1] = bad
    z, o = torch.zeros_like(mu), torch.ones_like(mu)
    _same_nonfinite(PF.kl_diag(mub.cuda(), sg.cuda(), z.cuda(), o.cuda()), T.kl_diag(mub, sg, z, o), "KL (mu)")
    sgb = sg.clone(); sgb[0, 1, 2, 2, 2] = bad
    _same_nonfinite(PF.kl_diag(mu.cuda(), sgb.cuda(), z.cuda(), o.cuda()), T.kl_diag(mu, sgb, z, o), "KL (sigma)")
    # and the gradient carries the NaN back to the offending element (anomaly mode would flag it there)
    with torch.autograd.set_detect_anomaly(False):      # (the reference's PULPo.__init__ switches it on process-wide)
        m = mub.cuda().requires_grad_(True)
        PF.kl_diag(m, sg.cuda(), z.cuda(), o.cuda()).backward()
        mr = mub.clone().requires_grad_(True)
        T.kl_diag(mr, sg, z, o).backward()
    assert np.array_equal(np.isnan(m.grad.cpu().numpy()), np.isnan(mr.grad.numpy()))


# ----------------------------------------------------------------------------- VecInt step synchronisation
@pytest.mark.parametrize("shape,B,amp", [((80, 96, 112), 1, 45.0), ((20, 24, 28), 2, 12.0), ((10, 12, 14), 1, 6.0),
                                         ((16, 24, 32), 2, 200.0), ((9, 11, 13), 1, 3.0)])
def test_vecint_dataflow_sync_equals_grid_barriers(PF, shape, B, amp, monkeypatch):
    """The item-to-item (row counter) synchronisation of the integration steps against the grid-barrier path it
    replaces: forward bit-identical, gradients equal up to the order of the scatter's atomic sums.  Covers the
    level-0 size, ragged rows (no cache-line ownership -> acquire fence), batches and a field whose reach exceeds
    the volume (every item waits for every row)."""
    from pulpo_b200 import synthetic as syn
    vec = syn.make_field(shape, 31, batch=B, max_abs=amp).cuda()
    gout = syn.make_field(shape, 32, batch=B, max_abs=1.0).cuda()

    def run():
        v = vec.clone().requires_grad_(True)      # requires_grad -> saved states -> dataflow path when enabled
        o = PF.vecint(v, 7)
        o.backward(gout)
        torch.cuda.synchronize()
        return o.detach(), v.grad.detach()
    monkeypatch.setenv("PULPO_VI_DATAFLOW", "0")
    o_bar, g_bar = run()
    monkeypatch.setenv("PULPO_VI_DATAFLOW", "1")
    for _ in range(3):                            # several launches: the counters are re-armed by every launch
        o_flow, g_flow = run()
        assert torch.equal(o_flow, o_bar), "forward differs: max-abs %.3e" % float((o_flow - o_bar).abs().max())
        assert_grad_close(g_flow.cpu().numpy(), g_bar.cpu().numpy(), "vecint grad (dataflow vs barriers)", rtol=1e-5)
    with torch.no_grad():
        assert torch.equal(PF.vecint(vec, 7), o_bar)          # forward-only (two ping-pong states, barriers) agrees too


def test_vecint_dataflow_multi_level_plan_equals_barriers(PF, monkeypatch):
    """All pyramid levels in one cooperative launch (HotPathPlan) with dataflow vs grid barriers."""
    from pulpo_b200 import synthetic as syn
    size, total, latent = [64, 96, 112], 4, 3
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(size, total, latent, seed=5)
    dev_in = (x.cuda(), y.cuda(), {l: dfs[l].cuda() for l in dfs}, {l: mus[l].cuda() for l in dfs},
              {l: sgs[l].cuda() for l in dfs})
    monkeypatch.setenv("PULPO_VI_DATAFLOW", "0")
    ref = _run_plan(dev_in, total, latent, size, 1, True, False)
    monkeypatch.setenv("PULPO_VI_DATAFLOW", "1")
    new = _run_plan(dev_in, total, latent, size, 1, True, True)
    assert_loss_close(new.total.item(), ref.total.item(), "total")
    for l in range(latent):
        assert torch.equal(new.integ[l], ref.integ[l]), "integrated field %d differs" % l
        assert_grad_close(new.gdf[l].cpu().numpy(), ref.gdf[l].cpu().numpy(), "gdf %d" % l, rtol=1e-5)


# ----------------------------------------------------------------------------- MC sampling in one launch (config 3)
def test_philox_sampler_is_sharding_invariant_and_standard_normal(PF):
    """gauss_sampler (src/network_blocks.py:7-8) for all levels of an MC sample in one graph-capturable launch:
    z = mu + sigma * eps exactly; eps depends on (seed, sample id, level, element) only -- the same sample id gives
    the same bits whether it is addressed explicitly or through the device-side counter of another 'rank' -- and is
    standard normal."""
    from pulpo_b200 import mc
    g = torch.Generator(device="cuda").manual_seed(1)
    mu = {l: torch.randn(1, 3, 20 >> l, 24 >> l, 28 >> l, device="cuda", generator=g) for l in range(3)}
    sg = {l: torch.rand(1, 3, 20 >> l, 24 >> l, 28 >> l, device="cuda", generator=g) + 0.1 for l in range(3)}
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    a = mc.PhiloxSampler(mu, sg, seed=9, first_id=1, id_stride=4, count_dev=cnt, dump_noise=True)   # "rank 1 of 4"
    b = mc.PhiloxSampler(mu, sg, seed=9, dump_noise=True)
    cnt.fill_(2)                                     # third sample of that rank -> sample id 1 + 4 * 2 = 9
    za = {l: t.clone() for l, t in a.draw().items()}
    zb = {l: t.clone() for l, t in b.draw(9).items()}      # draw() returns the sampler's static buffers
    for l in mu:
        assert torch.equal(za[l], zb[l]), "level %d: sample 9 differs between counter and explicit addressing" % l
        assert torch.equal(zb[l], mu[l] + sg[l] * b.eps[l])
    z8 = {l: t.clone() for l, t in b.draw(8).items()}
    assert not torch.equal(z8[0], zb[0])
    assert not torch.equal(mc.PhiloxSampler(mu, sg, seed=10).draw(9)[0], zb[0])
    big = {0: torch.zeros(1, 3, 64, 64, 64, device="cuda")}
    s = mc.PhiloxSampler(big, {0: torch.ones_like(big[0])}, seed=3)
    e = s.draw(0)[0].double()
    n = e.numel()
    assert abs(float(e.mean())) < 5 / n ** 0.5 and abs(float(e.var()) - 1.0) < 5 * (2.0 / n) ** 0.5
    assert abs(float((e ** 3).mean())) < 0.02 and abs(float((e ** 4).mean()) - 3.0) < 0.05
    assert float(e.abs().max()) < 6.5
    e2 = s.draw(1)[0].double()
    assert abs(float((e * e2).mean())) < 5 / n ** 0.5          # different samples are uncorrelated


def test_streaming_stats_graph_matches_mcmoments(PF):
    """StreamingStats (all maps + squared errors in one launch, device-side sample count, CUDA-graph replayed)
    against the per-map MCMoments / MCSqErr updates."""
    from pulpo_b200 import mc
    g = torch.Generator(device="cuda").manual_seed(2)
    bufs = {"a": torch.empty(3, 6, 7, 8, device="cuda"), "b": torch.empty(1, 9, 5, 4, device="cuda")}
    tgt = torch.rand(1, 9, 5, 4, device="cuda", generator=g)
    st = mc.StreamingStats(bufs, targets={"b": tgt})
    ref = {k: mc.MCMoments(v.shape, "cuda") for k, v in bufs.items()}
    sq = mc.MCSqErr(bufs["b"].shape, "cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        st.update()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        st.update()
    st.reset()
    for i in range(7):
        for k in bufs:
            bufs[k].copy_(torch.randn(bufs[k].shape, device="cuda", generator=g) * (i + 1))
            ref[k].update(bufs[k])
        sq.update(bufs["b"], tgt)
        gr.replay()
        st.count += 1
    out = st.states()
    for k in bufs:
        assert out[k].count == 7
        assert torch.equal(out[k].mean, ref[k].mean) and torch.equal(out[k].m2, ref[k].m2)
    assert torch.equal(out["b:sqerr"].acc, sq.acc)
    one = mc.sliced_uncertainty(out)
    assert torch.equal(one["a"], ref["a"].std_channel_mean()) and torch.equal(one["b:mse"], sq.mse())
    flat = st.reduce_to_maps([7])                 # the flat (multi-GPU) reduction path at world size 1
    torch.testing.assert_close(flat["a"], ref["a"].std_channel_mean(), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(flat["b:mse"], sq.mse()[0], rtol=1e-6, atol=1e-9)


def test_moments_merge_std_kernel_equals_sequential_chan_merges(PF):
    """The one-pass W-way merge + std of the multi-GPU MC reduction == W - 1 pairwise Chan merges then std
    (same arithmetic, same order), including a part that saw no samples."""
    from pulpo_b200 import mc
    g = torch.Generator(device="cuda").manual_seed(8)
    chunk, counts = 5003, [3, 0, 4, 2]
    parts = []
    for c in counts:
        st = mc.MCMoments((chunk,), "cuda")
        for _ in range(c):
            st.update(torch.randn(chunk, device="cuda", generator=g) * 2.0 + 1.0)
        parts.append(st)
    mean_parts = torch.cat([p.mean for p in parts]); m2_parts = torch.cat([p.m2 for p in parts])
    out = torch.empty(chunk, device="cuda")
    PF.moments_merge_std(mean_parts, m2_parts, counts, chunk, out)
    ref = mc.MCMoments((chunk,), "cuda")
    for p in parts:
        ref.merge_state(p.mean, p.m2, p.count)
    assert torch.equal(out, ref.std())
