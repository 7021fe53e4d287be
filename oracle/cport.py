"""ctypes binding of the plain-C oracle (oracle/pulpo_oracle.c).  TEST INFRASTRUCTURE ONLY.

Arrays are contiguous float32 numpy arrays shaped [B, C, D0, D1, D2].
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpulpo_oracle.so")
_lib = None

CPU_EXACT = 0
CUDA_RCP = 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pulpo_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_ncc.restype = ctypes.c_double
        _lib.orc_kl_diag_fwd.restype = ctypes.c_double
        _lib.orc_l2reg_fwd.restype = ctypes.c_double
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _dims(a):
    B, C, D0, D1, D2 = a.shape
    return [ctypes.c_int(int(v)) for v in (B, C, D0, D1, D2)]


def warp3d_fwd(df, img, mode=CPU_EXACT, want_idx=False):
    df, img = _f(df), _f(img)
    B, C, D0, D1, D2 = img.shape
    out = np.empty_like(img)
    idx = np.empty((B, 3, D0, D1, D2), np.int32) if want_idx else None
    lib().orc_warp3d_fwd(_p(img), _p(df), _p(out), _p(idx), *_dims(img), ctypes.c_int(mode))
    return (out, idx) if want_idx else out


def warp3d_bwd(gout, df, img, mode=CPU_EXACT, need_gimg=True):
    gout, df, img = _f(gout), _f(df), _f(img)
    gimg = np.zeros_like(img) if need_gimg else None
    gdf = np.empty_like(df)
    lib().orc_warp3d_bwd(_p(gout), _p(img), _p(df), _p(gimg), _p(gdf), *_dims(img), ctypes.c_int(mode))
    return gimg, gdf


def vecint_fwd(vec, nsteps=7, mode=CPU_EXACT):
    vec = _f(vec)
    B, C, D0, D1, D2 = vec.shape
    steps = np.empty((nsteps + 1,) + vec.shape, np.float32)
    lib().orc_vecint_fwd(_p(vec), _p(steps), ctypes.c_int(nsteps), ctypes.c_int(B), ctypes.c_int(D0),
                         ctypes.c_int(D1), ctypes.c_int(D2), ctypes.c_int(mode))
    return steps


def vecint_bwd(gout, steps, mode=CPU_EXACT):
    gout, steps = _f(gout), _f(steps)
    nsteps = steps.shape[0] - 1
    B, C, D0, D1, D2 = gout.shape
    gvec = np.empty_like(gout)
    lib().orc_vecint_bwd(_p(gout), _p(steps), _p(gvec), ctypes.c_int(nsteps), ctypes.c_int(B),
                         ctypes.c_int(D0), ctypes.c_int(D1), ctypes.c_int(D2), ctypes.c_int(mode))
    return gvec


def resize_up_fwd(x, factor, scale, addend=None):
    x = _f(x)
    B, C, d0, d1, d2 = x.shape
    out = np.empty((B, C, factor * d0, factor * d1, factor * d2), np.float32)
    addend = None if addend is None else _f(addend)
    lib().orc_resize_up_fwd(_p(x), _p(addend), _p(out), ctypes.c_int(factor), ctypes.c_float(scale), *_dims(x))
    return out


def resize_up_bwd(gout, factor, scale):
    gout = _f(gout)
    B, C, o0, o1, o2 = gout.shape
    gx = np.empty((B, C, o0 // factor, o1 // factor, o2 // factor), np.float32)
    lib().orc_resize_up_bwd(_p(gout), _p(gx), ctypes.c_int(factor), ctypes.c_float(scale), *_dims(gx))
    return gx


def interp_size_fwd(x, size):
    x = _f(x)
    B, C = x.shape[:2]
    out = np.empty((B, C) + tuple(int(s) for s in size), np.float32)
    lib().orc_interp_size_fwd(_p(x), _p(out), *_dims(x), *[ctypes.c_int(int(s)) for s in size])
    return out


def avgpool2_fwd(x):
    x = _f(x)
    B, C, D0, D1, D2 = x.shape
    out = np.empty((B, C, (D0 + 1) // 2, (D1 + 1) // 2, (D2 + 1) // 2), np.float32)
    lib().orc_avgpool2_fwd(_p(x), _p(out), *_dims(x))
    return out


def ncc(pred, target, win=9, gamma=0.05, want_grad=False):
    pred, target = _f(pred), _f(target)
    g = np.empty_like(pred) if want_grad else None
    loss = lib().orc_ncc(_p(pred), _p(target), _p(g), ctypes.c_int(win), ctypes.c_float(gamma), *_dims(pred))
    return (loss, g) if want_grad else loss


def kl_diag_fwd(mu0, sg0, mu1=None, sg1=None, eps=1e-10):
    mu0, sg0 = _f(mu0), _f(sg0)
    mu1 = None if mu1 is None else _f(mu1)
    sg1 = None if sg1 is None else _f(sg1)
    B = mu0.shape[0]
    n = mu0.size // B
    return lib().orc_kl_diag_fwd(_p(mu0), _p(sg0), _p(mu1), _p(sg1), ctypes.c_float(eps),
                                 ctypes.c_int(B), ctypes.c_longlong(n))


def kl_diag_bwd(mu0, sg0, mu1=None, sg1=None, eps=1e-10, gscale=1.0):
    mu0, sg0 = _f(mu0), _f(sg0)
    mu1 = None if mu1 is None else _f(mu1)
    sg1 = None if sg1 is None else _f(sg1)
    B = mu0.shape[0]
    n = mu0.size // B
    gm, gs = np.empty_like(mu0), np.empty_like(sg0)
    lib().orc_kl_diag_bwd(_p(mu0), _p(sg0), _p(mu1), _p(sg1), ctypes.c_float(eps), ctypes.c_float(gscale),
                          _p(gm), _p(gs), ctypes.c_int(B), ctypes.c_longlong(n))
    return gm, gs


def l2reg_fwd(f, lamb):
    f = _f(f)
    return lib().orc_l2reg_fwd(_p(f), ctypes.c_float(lamb), *_dims(f))


def l2reg_bwd(f, lamb, gscale=1.0):
    f = _f(f)
    g = np.empty_like(f)
    lib().orc_l2reg_bwd(_p(f), ctypes.c_float(lamb), ctypes.c_float(gscale), _p(g), *_dims(f))
    return g
