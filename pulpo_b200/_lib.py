"""ctypes binding of libpulpo_b200.so (the C ABI declared in include/pulpo_b200.h).

There is NO fallback: if the CUDA library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PULPO_B200_LIB selects a tuning variant built with `python -m pulpo_b200.build --tag=...`
LIB_PATH = os.environ.get("PULPO_B200_LIB") or os.path.join(_HERE, "lib", "libpulpo_b200.so")

CPU_EXACT = 0   # PULPO_COORD_CPU_EXACT
CUDA_RCP = 1    # PULPO_COORD_CUDA_RCP
FAST = 2        # PULPO_COORD_FAST (VecInt only)

_vp, _i, _f, _sz, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_longlong



class VecIntLevel(ctypes.Structure):
    """pulpo_vecint_level (include/pulpo_b200.h)."""
    _fields_ = [("inp", _vp), ("out", _vp), ("ws", _vp), ("ws_bytes", _sz), ("scratch", _vp), ("scratch_bytes", _sz),
                ("D0", _i), ("D1", _i), ("D2", _i)]


class KlLevel(ctypes.Structure):
    """pulpo_kl_level (include/pulpo_b200.h)."""
    _fields_ = [("mu", _vp), ("sigma", _vp), ("gmu", _vp), ("gsigma", _vp), ("out", _vp), ("n", _ll), ("weight", _f)]


class GaussLevel(ctypes.Structure):
    """pulpo_gauss_level (include/pulpo_b200.h)."""
    _fields_ = [("mu", _vp), ("sigma", _vp), ("z", _vp), ("eps_out", _vp), ("n", _ll)]


class MomentsMap(ctypes.Structure):
    """pulpo_moments_map (include/pulpo_b200.h)."""
    _fields_ = [("x", _vp), ("mean", _vp), ("m2", _vp), ("target", _vp), ("sqerr_acc", _vp), ("n", _ll)]


# name -> (restype, argtypes); mirrors include/pulpo_b200.h one to one
SIGNATURES = {
    "pulpo_version": (_i, []),
    "pulpo_strerror": (ctypes.c_char_p, [_i]),
    "pulpo_warp3d_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_warp3d_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_warp3d_fwd_img": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_warp3d_bwd_img": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_warp3d_fwd_dpos": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "pulpo_warp3d_bwd_dpos": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "pulpo_warp3d_l2reg_fwd": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_warp3d_l2reg_bwd": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_vecint_ws_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "pulpo_vecint_fwd": (_i, [_vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_vecint_bwd_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "pulpo_vecint_bwd": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_vecint_multi_fwd": (_i, [ctypes.POINTER(VecIntLevel), _i, _i, _i, _i, _i, _vp]),
    "pulpo_vecint_multi_bwd": (_i, [ctypes.POINTER(VecIntLevel), _i, _i, _i, _i, _vp]),
    "pulpo_combine_vecint_multi_fwd": (_i, [ctypes.POINTER(VecIntLevel), ctypes.POINTER(ctypes.c_void_p), _i, _i, _i, _i, _i, _vp]),
    "pulpo_combine_vecint_multi_bwd": (_i, [ctypes.POINTER(VecIntLevel), _i, _i, _i, _i, _vp]),
    "pulpo_resize_up_fwd": (_i, [_vp, _vp, _vp, _i, _f, _i, _i, _i, _i, _i, _vp]),
    "pulpo_resize_up_bwd": (_i, [_vp, _vp, _i, _f, _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_resize_up2_bwd_dpos": (_i, [_vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _vp]),
    "pulpo_interp_size_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_avgpool2_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "pulpo_avgpool2_pyramid_fwd": (_i, [_vp, ctypes.POINTER(ctypes.c_void_p), _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_ncc_ws_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "pulpo_ncc_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _i, _f, _i, _i, _i, _i, _i, _vp]),
    "pulpo_ncc_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _f, _i, _i, _i, _i, _i, _vp]),
    "pulpo_reduce_ws_bytes": (_sz, []),
    "pulpo_kl_diag_fwd": (_i, [_vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _sz, _i, _ll, _vp]),
    "pulpo_kl_diag_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _i, _ll, _vp]),
    "pulpo_gauss_sample_kl_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _sz, _i, _ll, _vp]),
    "pulpo_gauss_sample_kl_bwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _i, _ll, _vp]),
    "pulpo_kl_multi_ws_bytes": (_sz, []),
    "pulpo_kl_n01_multi": (_i, [ctypes.POINTER(KlLevel), _i, _f, _i, _vp, _sz, _vp]),
    "pulpo_l2reg_fwd": (_i, [_vp, _f, _vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "pulpo_l2reg_bwd": (_i, [_vp, _vp, _f, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pulpo_l2reg_fwd_bwd": (_i, [_vp, _f, _vp, _vp, _vp, _vp, _i, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "pulpo_l2reg_up2_scratch_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "pulpo_l2reg_up2_fwd_bwd": (_i, [_vp, _f, _vp, _vp, _i, _vp, _sz, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "pulpo_jacdet_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "pulpo_jacdet_bwd_ws_bytes": (_sz, [_i, _i, _i, _i]),
    "pulpo_jacdet_bwd": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "pulpo_std_ws_bytes": (_sz, []),
    "pulpo_std_fwd": (_i, [_vp, _f, _vp, _vp, _sz, _ll, _vp]),
    "pulpo_std_bwd": (_i, [_vp, _vp, _vp, _f, _vp, _ll, _vp]),
    "pulpo_moments_update": (_i, [_vp, _vp, _vp, _i, _ll, _vp]),
    "pulpo_moments_merge": (_i, [_vp, _vp, _i, _vp, _vp, _i, _ll, _vp]),
    "pulpo_moments_std": (_i, [_vp, _vp, _i, _ll, _vp]),
    "pulpo_gauss_sample_multi": (_i, [ctypes.POINTER(GaussLevel), _i, ctypes.c_ulonglong, _vp, _i, _i, _f, _vp]),
    "pulpo_moments_merge_std": (_i, [_vp, _vp, ctypes.POINTER(ctypes.c_int), _i, _ll, _vp, _vp]),
    "pulpo_moments_update_multi": (_i, [ctypes.POINTER(MomentsMap), _i, _vp, _vp]),
    "pulpo_counter_add": (_i, [_vp, _i, _i, _vp]),
    "pulpo_loss_total": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp]),
    "pulpo_sqerr_update": (_i, [_vp, _vp, _vp, _i, _ll, _vp]),
    "pulpo_global_ncc_ws_bytes": (_sz, []),
    "pulpo_global_ncc": (_i, [_vp, _vp, _f, _f, _i, _ll, _vp, _vp, _sz, _vp]),
}

_lib = None


class KernelProfiler:
    """Optional per-call instrumentation used by bench.py: counts C-ABI launches and, when
    ``timing`` is on, brackets every call with CUDA events on the current stream."""

    def __init__(self):
        self.enabled = False
        self.timing = False
        self.records = []      # (name, args, start_event, end_event)
        self.launches = 0

    def reset(self):
        self.records, self.launches = [], 0


profiler = KernelProfiler()
_LAUNCHING = None  # names of entry points that enqueue a kernel


class _Proxy:
    def __init__(self, handle):
        self._h = handle

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if not profiler.enabled or name.endswith("_bytes") or name in ("pulpo_version", "pulpo_strerror"):
            return fn

        def wrapped(*args):
            profiler.launches += 1
            if not profiler.timing:
                return fn(*args)
            import torch
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*args)
            e.record()
            profiler.records.append((name, args, s, e))
            return r
        return wrapped


def lib():
    """Load the library once.  Raises RuntimeError (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libpulpo_b200.so not found at %s -- build it with `python -m pulpo_b200.build` "
                "(nvcc, sm_100a). pulpo_b200 has no CPU or PyTorch fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = _Proxy(handle)
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().pulpo_strerror(status).decode()
        raise RuntimeError("libpulpo_b200 %s failed: %s (status %d)" % (what, msg, status))
