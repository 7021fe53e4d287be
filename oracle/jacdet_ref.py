"""numpy restatement of jacobian_det / JDetStd (src/losses.py:147-204, 3-D branch).  TEST INFRASTRUCTURE ONLY.

fp32 arithmetic in the reference's op order: normalise by 2/shape per channel (:176-179), flip the
channels and scale flipped channel j by (shape[j]-1-1)/2 (:194), replication-padded central
differences (-0.5, 0, 0.5) along z, y, x (:181-197), add the identity, 3x3 determinant as written (:199).
Pinned against the live reference by tests/golden/jacdet.npz (oracle/gen_golden.py)."""
from __future__ import annotations

import numpy as np


def jacobian_det(df, normalize=True):
    df = np.asarray(df, dtype=np.float32)
    B, C, D, H, W = df.shape
    shape = np.array([D, H, W], dtype=np.float32)
    f = df
    if normalize:
        f = np.stack([(f[:, c] * np.float32(2.0)) / shape[c] for c in range(3)], axis=1)
    scale = (shape - np.float32(1.0) - np.float32(1.0)).reshape(1, 3, 1, 1, 1)
    vox = (f[:, ::-1] * scale / np.float32(2.0)).astype(np.float32)

    def grad(axis):
        pad = [(0, 0)] * 5
        pad[axis] = (1, 1)
        p = np.pad(vox, pad, mode="edge")
        lo = np.take(p, np.arange(0, p.shape[axis] - 2), axis=axis)
        hi = np.take(p, np.arange(2, p.shape[axis]), axis=axis)
        return (np.float32(-0.5) * lo + np.float32(0.5) * hi).astype(np.float32)

    J = np.stack([grad(2), grad(3), grad(4)], axis=1)   # [B, axis, channel, D, H, W]
    for a in range(3):
        J[:, a, a] += np.float32(1.0)
    det = (J[:, 0, 0] * (J[:, 1, 1] * J[:, 2, 2] - J[:, 2, 1] * J[:, 1, 2])
           - J[:, 0, 1] * (J[:, 1, 0] * J[:, 2, 2] - J[:, 2, 0] * J[:, 1, 2])
           + J[:, 0, 2] * (J[:, 1, 0] * J[:, 2, 1] - J[:, 2, 0] * J[:, 1, 1]))
    return det.astype(np.float32)


def jdet_std(df, lamb, normalize=True):
    return float(lamb) * float(jacobian_det(df, normalize).astype(np.float64).std(ddof=1))
