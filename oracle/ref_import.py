"""Import the REAL reference: /root/reference (read-only, build container) or, on the GPU box, the
unmodified copy of its `src` package staged under oracle/_ref by oracle/stage_ref.py.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py (fixture generation), by CPU tests that skip when
no reference tree is present, and by the whole-model drop-in test on the GPU (tests/test_gpu_dropin.py).

``src.models`` needs ``pytorch_lightning`` and ``torchvision`` symbols that are not
installed / not needed; a minimal stub is injected into ``sys.modules``.
"""
from __future__ import annotations

import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/stage_ref.py (GPU box)
REF_ROOT = os.environ.get("PULPO_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "src")) and os.path.isdir(os.path.join(_STAGED, "src")):
    REF_ROOT = _STAGED


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src"))


def _stub_lightning():
    if "pytorch_lightning" in sys.modules:
        return
    import torch.nn as nn

    pl = types.ModuleType("pytorch_lightning")

    class _HP(dict):
        __getattr__ = dict.__getitem__

    class LightningModule(nn.Module):
        def save_hyperparameters(self):
            import inspect
            frame = inspect.currentframe().f_back
            args = inspect.getargvalues(frame)
            self.hparams = _HP({k: args.locals[k] for k in args.args if k != "self"})

        def log_dict(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    sys.modules["pytorch_lightning"] = pl
    try:
        import torchvision.utils  # noqa: F401
    except Exception:
        tv = types.ModuleType("torchvision")
        tvu = types.ModuleType("torchvision.utils")
        tvu.make_grid = lambda *a, **k: None
        tvu.flow_to_image = lambda *a, **k: None
        tv.utils = tvu
        sys.modules["torchvision"] = tv
        sys.modules["torchvision.utils"] = tvu


def load():
    """Returns (network_blocks, losses, components.pulpo, models) of the real reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _stub_lightning()
    import importlib
    nb = importlib.import_module("src.network_blocks")
    ls = importlib.import_module("src.losses")
    cp = importlib.import_module("src.components.pulpo")
    md = importlib.import_module("src.models")
    return nb, ls, cp, md
