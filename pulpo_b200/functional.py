"""torch.autograd wrappers over the C ABI (include/pulpo_b200.h).

Every function takes/returns CUDA fp32 tensors laid out [B,C,D0,D1,D2]; tensors are made
contiguous here, all memory (outputs, saved states, workspaces) is allocated by torch and
passed down as raw pointers, and kernels are enqueued on torch's current stream.  There is no
CPU path: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import CPU_EXACT, CUDA_RCP, FAST, check  # noqa: F401

_vp = ctypes.c_void_p


def _ptr(t):
    return None if t is None else _vp(t.data_ptr())


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _prep(t, name):
    if not t.is_cuda:
        raise RuntimeError("pulpo_b200: %s must be a CUDA tensor (there is no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("pulpo_b200: %s must be float32, got %s" % (name, t.dtype))
    return t.contiguous()


def _dims5(t, name):
    if t.dim() != 5:
        raise NotImplementedError("pulpo_b200: %s must be [B,C,D0,D1,D2] (only ndims == 3 is implemented)" % name)
    return tuple(int(s) for s in t.shape)


# reduction workspaces: tiny, zero-initialised once, self-resetting; one per (device, stream)
_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# ----------------------------------------------------------------------------- warp (a2)
class _Warp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, df, img, coord_mode):
        df, img = _prep(df, "df"), _prep(img, "moving_image")
        B, C, I0, I1, I2 = _dims5(img, "moving_image")
        Bd, Cd, D0, D1, D2 = _dims5(df, "df")
        if Bd != B or Cd != 3:
            raise RuntimeError("pulpo_b200: df must be [B,3,*size] with moving_image's batch size, got %s vs %s"
                               % (tuple(df.shape), tuple(img.shape)))
        # grid_sample semantics: the output has the field's spatial size; the image may have another one (a
        # level-sized field resampling a full-resolution image, evaluate.py:198,240,246)
        out = torch.empty((B, C, D0, D1, D2), dtype=torch.float32, device=img.device)
        if (I0, I1, I2) == (D0, D1, D2):
            check(_lib.lib().pulpo_warp3d_fwd(_ptr(img), _ptr(df), _ptr(out), None, B, C, D0, D1, D2, coord_mode,
                                              _stream()), "warp3d_fwd")
        else:
            check(_lib.lib().pulpo_warp3d_fwd_img(_ptr(img), _ptr(df), _ptr(out), None, B, C, D0, D1, D2, I0, I1, I2,
                                                  coord_mode, _stream()), "warp3d_fwd_img")
        ctx.save_for_backward(df, img)
        ctx.coord_mode = coord_mode
        return out

    @staticmethod
    def backward(ctx, gout):
        df, img = ctx.saved_tensors
        gout = _prep(gout, "grad_output")
        B, C, I0, I1, I2 = img.shape
        D0, D1, D2 = df.shape[2:]
        need_df, need_img = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gdf = torch.empty_like(df) if need_df else None
        gimg = torch.zeros_like(img) if need_img else None
        if need_df or need_img:
            if (I0, I1, I2) == (D0, D1, D2):
                check(_lib.lib().pulpo_warp3d_bwd(_ptr(gout), _ptr(img), _ptr(df), _ptr(gimg), _ptr(gdf), B, C, D0, D1,
                                                  D2, ctx.coord_mode, _stream()), "warp3d_bwd")
            else:
                check(_lib.lib().pulpo_warp3d_bwd_img(_ptr(gout), _ptr(img), _ptr(df), _ptr(gimg), _ptr(gdf), B, C, D0,
                                                      D1, D2, I0, I1, I2, ctx.coord_mode, _stream()), "warp3d_bwd_img")
        return gdf, gimg, None


def warp(df, moving_image, coord_mode=CPU_EXACT):
    """SpatialTransformer.forward (src/network_blocks.py:101-121)."""
    return _Warp.apply(df, moving_image, coord_mode)


def warp_indices(df, moving_image, coord_mode=CPU_EXACT):
    """Warp + the int32 floor(p) corner indices [B,3,D0,D1,D2] (bit-exact parity contract)."""
    df, img = _prep(df, "df"), _prep(moving_image, "moving_image")
    B, C, D0, D1, D2 = _dims5(img, "moving_image")
    out = torch.empty_like(img)
    idx = torch.empty((B, 3, D0, D1, D2), dtype=torch.int32, device=img.device)
    check(_lib.lib().pulpo_warp3d_fwd(_ptr(img), _ptr(df), _ptr(out), _ptr(idx), B, C, D0, D1, D2, coord_mode,
                                      _stream()), "warp3d_fwd")
    return out, idx


# ----------------------------------------------------------------------------- VecInt (a3)
class _VecInt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec, nsteps, coord_mode):
        vec = _prep(vec, "vec")
        B, C, D0, D1, D2 = _dims5(vec, "vec")
        if C != 3:
            raise RuntimeError("pulpo_b200: VecInt expects a 3-channel field, got C=%d" % C)
        save = 1 if ctx.needs_input_grad[0] else 0
        L = _lib.lib()
        nbytes = L.pulpo_vecint_ws_bytes(nsteps, save, B, D0, D1, D2)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=vec.device)
        out = torch.empty_like(vec)
        check(L.pulpo_vecint_fwd(_ptr(vec), _ptr(out), _ptr(ws), nbytes, nsteps, save, B, D0, D1, D2, coord_mode,
                                 _stream()), "vecint_fwd")
        ctx.nsteps, ctx.coord_mode, ctx.dims = nsteps, coord_mode, (B, D0, D1, D2)
        if save:
            ctx.save_for_backward(ws)
        return out

    @staticmethod
    def backward(ctx, gout):
        (ws,) = ctx.saved_tensors
        gout = _prep(gout, "grad_output")
        B, D0, D1, D2 = ctx.dims
        L = _lib.lib()
        nbytes = L.pulpo_vecint_bwd_scratch_bytes(B, D0, D1, D2)
        scratch = torch.empty(nbytes // 4, dtype=torch.float32, device=gout.device)
        gvec = torch.empty_like(gout)
        check(L.pulpo_vecint_bwd(_ptr(gout), _ptr(ws), _ptr(gvec), _ptr(scratch), nbytes, ctx.nsteps, B, D0, D1, D2,
                                 ctx.coord_mode, _stream()), "vecint_bwd")
        return gvec, None, None


def vecint(vec, nsteps=7, coord_mode=CPU_EXACT):
    """VecInt.forward (src/network_blocks.py:173-177).  ``CPU_EXACT`` (default) reproduces torch-CPU bit
    for bit; ``FAST`` computes the sample position in one FMA and interpolates with FMAs (fields within
    ~3e-5 of the reference -- inside the 1e-4 contract, but NCC gradients downstream amplify it to ~1e-3
    relative, and the kernel is bound by the L1 pipe, not by instruction issue, so it buys little)."""
    return _VecInt.apply(vec, nsteps, coord_mode)


# ----------------------------------------------------------------------------- resize (a4, a5)
class _ResizeUp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, addend, factor, scale):
        x = _prep(x, "x")
        B, C, d0, d1, d2 = _dims5(x, "x")
        oshape = (B, C, factor * d0, factor * d1, factor * d2)
        if addend is not None:
            addend = _prep(addend, "addend")
            if tuple(addend.shape) != oshape:
                raise RuntimeError("pulpo_b200: cannot add fields of shape %s and %s" % (oshape, tuple(addend.shape)))
        out = torch.empty(oshape, dtype=torch.float32, device=x.device)
        check(_lib.lib().pulpo_resize_up_fwd(_ptr(x), _ptr(addend), _ptr(out), factor, float(scale), B, C, d0, d1,
                                             d2, _stream()), "resize_up_fwd")
        ctx.factor, ctx.scale, ctx.idims = factor, float(scale), (B, C, d0, d1, d2)
        return out

    @staticmethod
    def backward(ctx, gout):
        gout = _prep(gout, "grad_output")
        gx = None
        if ctx.needs_input_grad[0]:
            B, C, d0, d1, d2 = ctx.idims
            gx = torch.empty(ctx.idims, dtype=torch.float32, device=gout.device)
            check(_lib.lib().pulpo_resize_up_bwd(_ptr(gout), _ptr(gx), ctx.factor, ctx.scale, 0, B, C, d0, d1, d2,
                                                 _stream()), "resize_up_bwd")
        return gx, (gout if ctx.needs_input_grad[1] else None), None, None


def resize_up(x, factor, scale=None, addend=None):
    """ResizeTransform.forward for integer factor>1 (value scale + trilinear up-sampling),
    optionally fused with DFAdder (src/network_blocks.py:138-158)."""
    return _ResizeUp.apply(x, addend, int(factor), float(factor if scale is None else scale))


def interp_to_size(x, size):
    """F.interpolate(x, size=size, mode='trilinear', align_corners=False) -- src/losses.py:313.
    Used on the fixed image (no gradient)."""
    x = _prep(x.detach(), "x")
    B, C, i0, i1, i2 = _dims5(x, "x")
    o0, o1, o2 = (int(s) for s in size)
    if (o0, o1, o2) == (i0, i1, i2):
        return x
    out = torch.empty((B, C, o0, o1, o2), dtype=torch.float32, device=x.device)
    check(_lib.lib().pulpo_interp_size_fwd(_ptr(x), _ptr(out), B, C, i0, i1, i2, o0, o1, o2, _stream()),
          "interp_size_fwd")
    return out


def avgpool2(x):
    """avg_pool3d(x, 2, 2, 0, ceil_mode=True) -- src/components/pulpo.py:174,177 (no gradient)."""
    x = _prep(x.detach(), "x")
    B, C, D0, D1, D2 = _dims5(x, "x")
    out = torch.empty((B, C, (D0 + 1) // 2, (D1 + 1) // 2, (D2 + 1) // 2), dtype=torch.float32, device=x.device)
    check(_lib.lib().pulpo_avgpool2_fwd(_ptr(x), _ptr(out), B, C, D0, D1, D2, _stream()), "avgpool2_fwd")
    return out


# ----------------------------------------------------------------------------- encoder feedback (f-4)
class _ResampleCat(torch.autograd.Function):
    """cat([F.interpolate(t, size, trilinear, align_corners=False) for t in items], dim=1) without the
    per-item temporaries and the concat pass: every item is resampled straight into its channel slice of
    the output (one launch per item and batch element); the backward is the exact adjoint in gather form."""

    @staticmethod
    def forward(ctx, size, *items):
        items = [_prep(t, "feedback item") for t in items]
        B = int(items[0].shape[0])
        o0, o1, o2 = (int(v) for v in size)
        So = o0 * o1 * o2
        ctot = sum(int(t.shape[1]) for t in items)
        out = torch.empty((B, ctot, o0, o1, o2), dtype=torch.float32, device=items[0].device)
        L, st = _lib.lib(), _stream()
        c0, meta = 0, []
        for t in items:
            Bt, C, i0, i1, i2 = _dims5(t, "feedback item")
            if Bt != B:
                raise RuntimeError("pulpo_b200: feedback items differ in batch size")
            for b in range(B):
                dst = _vp(out.data_ptr() + 4 * ((b * ctot + c0) * So))
                src = _vp(t.data_ptr() + 4 * b * C * i0 * i1 * i2)
                if (i0, i1, i2) == (o0, o1, o2):
                    out[b, c0:c0 + C].copy_(t[b])
                else:
                    check(L.pulpo_interp_size_fwd(src, dst, 1, C, i0, i1, i2, o0, o1, o2, st), "interp_size_fwd")
            meta.append((c0, C, i0, i1, i2))
            c0 += C
        ctx.meta, ctx.osize, ctx.ctot, ctx.B = meta, (o0, o1, o2), ctot, B
        return out

    @staticmethod
    def backward(ctx, gout):
        gout = _prep(gout, "grad_output")
        o0, o1, o2 = ctx.osize
        So, B, ctot = o0 * o1 * o2, ctx.B, ctx.ctot
        L, st = _lib.lib(), _stream()
        grads = []
        for k, (c0, C, i0, i1, i2) in enumerate(ctx.meta):
            if not ctx.needs_input_grad[1 + k]:
                grads.append(None)
                continue
            if (i0, i1, i2) == (o0, o1, o2):
                grads.append(gout[:, c0:c0 + C].contiguous())
                continue
            f = o0 // i0
            if f < 2 or (f * i0, f * i1, f * i2) != (o0, o1, o2):
                raise NotImplementedError("pulpo_b200: feedback resampling backward needs an integer size ratio "
                                          "(got %s -> %s)" % ((i0, i1, i2), (o0, o1, o2)))
            g = torch.empty((B, C, i0, i1, i2), dtype=torch.float32, device=gout.device)
            for b in range(B):
                src = _vp(gout.data_ptr() + 4 * ((b * ctot + c0) * So))
                dst = _vp(g.data_ptr() + 4 * b * C * i0 * i1 * i2)
                check(L.pulpo_resize_up_bwd(src, dst, f, 1.0, 0, 1, C, i0, i1, i2, st), "resize_up_bwd")
            grads.append(g)
        return (None, *grads)


def resample_cat(items, size):
    """The encoder's feedback tensor (src/components/pulpo.py:195-206): every item of the coarser level
    trilinearly resized to ``size`` and concatenated along channels."""
    return _ResampleCat.apply(tuple(int(v) for v in size), *items)


# ----------------------------------------------------------------------------- NCC (a9)
class _NCC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, win, gamma):
        pred, target = _prep(pred, "y_pred"), _prep(target, "y_true")
        B, C, D0, D1, D2 = _dims5(pred, "y_pred")
        if target.shape != pred.shape:
            raise RuntimeError("pulpo_b200: NCC inputs differ in shape: %s vs %s" % (tuple(pred.shape), tuple(target.shape)))
        L = _lib.lib()
        need = bool(ctx.needs_input_grad[0])
        abc = torch.empty((3,) + tuple(pred.shape), dtype=torch.float32, device=pred.device) if need else None
        nbytes = L.pulpo_ncc_ws_bytes(B, C, D0, D1, D2)
        ws = _workspace(nbytes, pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        check(L.pulpo_ncc_fwd(_ptr(pred), _ptr(target), _ptr(loss), _ptr(abc), _ptr(ws), ws.numel(), int(win),
                              float(gamma), B, C, D0, D1, D2, _stream()), "ncc_fwd")
        ctx.win, ctx.gamma = int(win), float(gamma)
        if need:
            ctx.save_for_backward(abc, pred, target)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        abc, pred, target = ctx.saved_tensors
        B, C, D0, D1, D2 = pred.shape
        gloss = gloss.to(torch.float32).contiguous()
        gpred = torch.empty_like(pred)
        check(_lib.lib().pulpo_ncc_bwd(_ptr(abc), _ptr(pred), _ptr(target), _ptr(gloss), _ptr(gpred), ctx.win,
                                       ctx.gamma, B, C, D0, D1, D2, _stream()), "ncc_bwd")
        return gpred, None, None, None


def ncc_loss(y_pred, y_true, win_size=9, gamma=0.05):
    """NCC_loss (src/losses.py:85-135).  Differentiable w.r.t. y_pred only (the fixed image
    never needs a gradient on this path)."""
    if y_true.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("pulpo_b200: NCC gradient w.r.t. y_true is not implemented (never needed: "
                                  "y_true is the fixed image)")
    return _NCC.apply(y_pred, y_true, win_size, gamma)


# ----------------------------------------------------------------------------- KL (a11)
def _const_value(t):
    """Value of a tensor tagged as an expanded constant (see components.pulpo.PULPoPrior), else
    None.  Only the Python-side tag is read: no device access, so this is CUDA-graph safe."""
    if t is None:
        return None
    tag = getattr(t, "_pulpo_const", None)
    if tag is not None and t.numel() > 0 and all(s == 0 for s in t.stride()):
        return float(tag)
    return None


class _KL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu0, sigma0, mu1, sigma1, eps):
        mu0, sigma0 = _prep(mu0, "mu0"), _prep(sigma0, "sigma0")
        mu1 = None if mu1 is None else _prep(mu1, "mu1")
        sigma1 = None if sigma1 is None else _prep(sigma1, "sigma1")
        B = int(mu0.shape[0])
        n = mu0.numel() // B
        L = _lib.lib()
        ws = _workspace(L.pulpo_reduce_ws_bytes(), mu0.device)
        out = torch.empty((), dtype=torch.float32, device=mu0.device)
        check(L.pulpo_kl_diag_fwd(_ptr(mu0), _ptr(sigma0), _ptr(mu1), _ptr(sigma1), float(eps), 1.0, _ptr(out), _ptr(ws),
                                  ws.numel(), B, n, _stream()), "kl_diag_fwd")
        ctx.eps, ctx.B, ctx.n = float(eps), B, n
        ctx.have1 = (mu1 is not None, sigma1 is not None)
        ctx.save_for_backward(mu0, sigma0, *[t for t in (mu1, sigma1) if t is not None])
        return out

    @staticmethod
    def backward(ctx, gloss):
        saved = list(ctx.saved_tensors)
        mu0, sigma0 = saved[0], saved[1]
        rest = saved[2:]
        mu1 = rest.pop(0) if ctx.have1[0] else None
        sigma1 = rest.pop(0) if ctx.have1[1] else None
        gloss = gloss.to(torch.float32).contiguous()
        gmu, gsg = torch.empty_like(mu0), torch.empty_like(sigma0)
        check(_lib.lib().pulpo_kl_diag_bwd(_ptr(gloss), _ptr(mu0), _ptr(sigma0), _ptr(mu1), _ptr(sigma1), ctx.eps,
                                           1.0, _ptr(gmu), _ptr(gsg), ctx.B, ctx.n, _stream()), "kl_diag_bwd")
        return gmu, gsg, None, None, None


def kl_diag(mu0, sigma0, mu1, sigma1, eps=1e-10):
    """KL_two_gauss_with_diag_cov (src/losses.py:47-76): KL[p0 || p1], differentiable w.r.t. p0.
    Expanded constant priors (mu1 == 0, sigma1 == 1 with stride 0) take the N(0,1) fast path."""
    for t in (mu1, sigma1):
        if t is not None and t.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("pulpo_b200: KL gradient w.r.t. the second distribution is not implemented "
                                      "(the prior is constant on this path, src/components/pulpo.py:337-339)")
    if _const_value(mu1) == 0.0:
        mu1 = None
    if _const_value(sigma1) == 1.0:
        sigma1 = None
    return _KL.apply(mu0, sigma0, mu1, sigma1, eps)


class _GaussSampleKL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, sigma, noise, mu1, sigma1, var, eps):
        mu, sigma, noise = _prep(mu, "mu"), _prep(sigma, "sigma"), _prep(noise, "noise")
        if sigma.shape != mu.shape or noise.shape != mu.shape:
            raise RuntimeError("pulpo_b200.gauss_sample_kl: mu, sigma and noise must have one shape")
        mu1 = None if mu1 is None else _prep(mu1, "mu1")
        sigma1 = None if sigma1 is None else _prep(sigma1, "sigma1")
        B = int(mu.shape[0])
        n = mu.numel() // B
        L = _lib.lib()
        ws = _workspace(L.pulpo_reduce_ws_bytes(), mu.device)
        z = torch.empty_like(mu)
        out = torch.empty((), dtype=torch.float32, device=mu.device)
        check(L.pulpo_gauss_sample_kl_fwd(_ptr(mu), _ptr(sigma), _ptr(noise), _ptr(mu1), _ptr(sigma1), float(var),
                                          float(eps), 1.0, _ptr(z), _ptr(out), _ptr(ws), ws.numel(), B, n, _stream()),
              "gauss_sample_kl_fwd")
        ctx.var, ctx.eps, ctx.B, ctx.n = float(var), float(eps), B, n
        ctx.have1 = (mu1 is not None, sigma1 is not None)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(mu, sigma, noise, *[t for t in (mu1, sigma1) if t is not None])
        return z, out

    @staticmethod
    def backward(ctx, gz, gloss):
        saved = list(ctx.saved_tensors)
        mu, sigma, noise = saved[:3]
        rest = saved[3:]
        mu1 = rest.pop(0) if ctx.have1[0] else None
        sigma1 = rest.pop(0) if ctx.have1[1] else None
        gz = None if gz is None else gz.to(torch.float32).contiguous()
        have_kl = gloss is not None
        gloss = None if gloss is None else gloss.to(torch.float32).contiguous()
        gmu, gsg = torch.empty_like(mu), torch.empty_like(sigma)
        check(_lib.lib().pulpo_gauss_sample_kl_bwd(_ptr(gz), _ptr(gloss), int(have_kl), _ptr(mu), _ptr(sigma),
                                                   _ptr(noise), _ptr(mu1), _ptr(sigma1), ctx.var, ctx.eps, 1.0,
                                                   _ptr(gmu), _ptr(gsg), ctx.B, ctx.n, _stream()),
              "gauss_sample_kl_bwd")
        return gmu, gsg, None, None, None, None, None


def gauss_sample_kl(mu, sigma, noise, mu1=None, sigma1=None, var=1, eps=1e-10):
    """gauss_sampler (src/network_blocks.py:7-8) and KL_two_gauss_with_diag_cov(mu, sigma, mu1, sigma1)
    (src/losses.py:47-76) in one pass over mu and sigma (SURVEY f-4).  Returns (z, kl); both are
    differentiable w.r.t. mu and sigma, and one fused backward serves whichever of them is used."""
    for t in (mu1, sigma1):
        if t is not None and t.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("pulpo_b200: KL gradient w.r.t. the second distribution is not implemented")
    if _const_value(mu1) == 0.0:
        mu1 = None
    if _const_value(sigma1) == 1.0:
        sigma1 = None
    return _GaussSampleKL.apply(mu, sigma, noise, mu1, sigma1, var, eps)


# ----------------------------------------------------------------------------- L2 reg (f-1)
class _L2Reg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, lamb):
        f = _prep(f, "deformation_field")
        B, C, D0, D1, D2 = _dims5(f, "deformation_field")
        L = _lib.lib()
        ws = _workspace(L.pulpo_reduce_ws_bytes(), f.device)
        out = torch.empty((), dtype=torch.float32, device=f.device)
        check(L.pulpo_l2reg_fwd(_ptr(f), float(lamb), _ptr(out), _ptr(ws), ws.numel(), B, C, D0, D1, D2, _stream()),
              "l2reg_fwd")
        ctx.lamb = float(lamb)
        ctx.save_for_backward(f)
        return out

    @staticmethod
    def backward(ctx, gloss):
        (f,) = ctx.saved_tensors
        B, C, D0, D1, D2 = f.shape
        gloss = gloss.to(torch.float32).contiguous()
        gf = torch.empty_like(f)
        check(_lib.lib().pulpo_l2reg_bwd(_ptr(gloss), _ptr(f), ctx.lamb, _ptr(gf), 0, B, C, D0, D1, D2, _stream()),
              "l2reg_bwd")
        return gf, None


def l2_reg(deformation_field, lamb=0.0):
    """L2_reg (src/losses.py:208-222), 3-D branch."""
    return _L2Reg.apply(deformation_field, lamb)


# ----------------------------------------------------------------------------- Jacobian determinant (f-2)
class _JacDet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, df, normalize):
        df = _prep(df, "deformation_field")
        B, C, D0, D1, D2 = _dims5(df, "deformation_field")
        if C != 3:
            raise RuntimeError("pulpo_b200: jacobian_det expects a 3-channel field, got C=%d" % C)
        det = torch.empty((B, D0, D1, D2), dtype=torch.float32, device=df.device)
        check(_lib.lib().pulpo_jacdet_fwd(_ptr(df), _ptr(det), int(bool(normalize)), B, D0, D1, D2, _stream()),
              "jacdet_fwd")
        ctx.normalize = int(bool(normalize))
        ctx.save_for_backward(df)
        return det

    @staticmethod
    def backward(ctx, gdet):
        (df,) = ctx.saved_tensors
        gdet = _prep(gdet, "grad_output")
        B, C, D0, D1, D2 = df.shape
        L = _lib.lib()
        nbytes = L.pulpo_jacdet_bwd_ws_bytes(B, D0, D1, D2)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=df.device)
        gdf = torch.empty_like(df)
        check(L.pulpo_jacdet_bwd(_ptr(gdet), _ptr(df), _ptr(gdf), _ptr(ws), nbytes, ctx.normalize, B, D0, D1, D2,
                                 _stream()), "jacdet_bwd")
        return gdf, None


def jacobian_det(deformation_field, normalize=True):
    """jacobian_det (src/losses.py:147-199), 3-D branch: [B,3,D,H,W] -> [B,D,H,W]."""
    return _JacDet.apply(deformation_field, normalize)


class _Std(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lamb):
        x = _prep(x, "x")
        L = _lib.lib()
        ws = torch.zeros(L.pulpo_std_ws_bytes(), dtype=torch.uint8, device=x.device)   # holds mean/std for backward
        out = torch.empty((), dtype=torch.float32, device=x.device)
        check(L.pulpo_std_fwd(_ptr(x), float(lamb), _ptr(out), _ptr(ws), ws.numel(), x.numel(), _stream()), "std_fwd")
        ctx.lamb = float(lamb)
        ctx.save_for_backward(x, ws)
        return out

    @staticmethod
    def backward(ctx, gloss):
        x, ws = ctx.saved_tensors
        gloss = gloss.to(torch.float32).contiguous()
        gx = torch.empty_like(x)
        check(_lib.lib().pulpo_std_bwd(_ptr(gloss), _ptr(x), _ptr(ws), ctx.lamb, _ptr(gx), x.numel(), _stream()),
              "std_bwd")
        return gx, None


def jdet_std(deformation_field, lamb=0.0, normalize=True):
    """JDetStd (src/losses.py:202-204): lamb * jacobian_det(field).std() (unbiased, all elements)."""
    return _Std.apply(jacobian_det(deformation_field, normalize), lamb)


# ----------------------------------------------------------------------------- MC moments (f-3)
def moments_update(x, mean, m2, count):
    """Welford update of per-voxel (mean, M2) with sample x; count includes x."""
    x = _prep(x, "x")
    check(_lib.lib().pulpo_moments_update(_ptr(x), _ptr(mean), _ptr(m2), int(count), x.numel(), _stream()),
          "moments_update")


def moments_merge(mean_a, m2_a, count_a, mean_b, m2_b, count_b):
    """Chan merge of two partial states into (mean_a, m2_a)."""
    check(_lib.lib().pulpo_moments_merge(_ptr(mean_a), _ptr(m2_a), int(count_a), _ptr(mean_b), _ptr(m2_b),
                                         int(count_b), mean_a.numel(), _stream()), "moments_merge")


def moments_std(m2, count):
    """Unbiased per-voxel std (torch.std(axis=0) of evaluate.py:243-251)."""
    out = torch.empty_like(m2)
    check(_lib.lib().pulpo_moments_std(_ptr(m2), _ptr(out), int(count), m2.numel(), _stream()), "moments_std")
    return out


def sqerr_update(x, y, acc, first):
    """acc (+)= (x - y)**2 per voxel: the streamed numerator of the MSE map (evaluate.py:1538)."""
    x, y = _prep(x, "x"), _prep(y, "y")
    if x.numel() != y.numel() or x.numel() != acc.numel():
        raise RuntimeError("sqerr_update: x %s, y %s, acc %s" % (tuple(x.shape), tuple(y.shape), tuple(acc.shape)))
    check(_lib.lib().pulpo_sqerr_update(_ptr(x), _ptr(y), _ptr(acc), int(bool(first)), x.numel(), _stream()), "sqerr_update")


def global_ncc(a, v, scale_a=1.0, scale_v=1.0, square_a=False):
    """Evaluate.ncc(a', v') of evaluate.py:334-353 (zero_norm=True) with a' = scale_a * (a**2 if square_a else a),
    v' = scale_v * v.  Returns a 2-element CUDA tensor: (ncc, mean(a'))."""
    a, v = _prep(a, "a"), _prep(v, "v")
    if a.numel() != v.numel():
        raise RuntimeError("global_ncc: %d vs %d elements" % (a.numel(), v.numel()))
    lib = _lib.lib()
    ws = torch.zeros(lib.pulpo_global_ncc_ws_bytes(), dtype=torch.uint8, device=a.device)
    out = torch.empty(2, dtype=torch.float32, device=a.device)
    check(lib.pulpo_global_ncc(_ptr(a), _ptr(v), float(scale_a), float(scale_v), int(bool(square_a)), a.numel(), _ptr(out),
                               _ptr(ws), ws.numel(), _stream()), "global_ncc")
    return out


def moments_merge_std(mean_parts, m2_parts, counts, chunk, out):
    """Chan merge of len(counts) partial (mean, M2) slices of ``chunk`` elements (part r at r * chunk) in part order
    and the unbiased std of the union, written to ``out`` -- one launch (multi-GPU MC reduction)."""
    import ctypes
    arr = (ctypes.c_int * len(counts))(*[int(c) for c in counts])
    check(_lib.lib().pulpo_moments_merge_std(_ptr(mean_parts), _ptr(m2_parts), arr, len(counts), int(chunk), _ptr(out), _stream()),
          "moments_merge_std")
    return out
