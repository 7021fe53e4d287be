import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# tolerances stated by BASELINE.json:north_star
FIELD_ATOL = 1e-4      # max-abs on warped intensities and fields
LOSS_RTOL = 1e-5       # relative error on losses
GRAD_RTOL = 1e-4       # gradients: max-abs <= GRAD_RTOL * max|g_ref| (atomics reorder fp32 sums)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def assert_close(a, b, atol, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    err = float(np.max(np.abs(a - b))) if a.size else 0.0
    assert err <= atol, "%s: max-abs error %.3e > %.1e" % (what, err, atol)


def assert_grad_close(g, ref, what="", rtol=GRAD_RTOL):
    g, ref = np.asarray(g, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert g.shape == ref.shape, "%s: shape %s vs %s" % (what, g.shape, ref.shape)
    scale = max(float(np.max(np.abs(ref))), 1e-30)
    err = float(np.max(np.abs(g - ref)))
    assert err <= rtol * scale, "%s: grad max-abs error %.3e > %.1e * %.3e" % (what, err, rtol, scale)
    den = float(np.linalg.norm(g.ravel()) * np.linalg.norm(ref.ravel()))
    if den > 0:
        cos = float(np.dot(g.ravel(), ref.ravel())) / den
        assert cos >= 1 - 1e-6, "%s: cosine %.9f" % (what, cos)


def assert_loss_close(v, ref, what="", rtol=LOSS_RTOL):
    v, ref = float(v), float(ref)
    assert abs(v - ref) <= rtol * max(abs(ref), 1e-30), "%s: %.9g vs %.9g (rel %.3e > %.1e)" % (
        what, v, ref, abs(v - ref) / max(abs(ref), 1e-30), rtol)
