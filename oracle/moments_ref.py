"""CPU restatement of the per-voxel MC moments (TEST INFRASTRUCTURE ONLY).

Reference semantics: ``torch.std(stack, axis=0)`` over the sample stacks of evaluate.py:243-251.
``TorchCpuOps`` is the same Welford update / Chan merge / unbiased std as csrc/losses.cu in plain
torch-CPU fp32 arithmetic; tests inject it into ``pulpo_b200.mc`` to exercise the multi-rank host
logic over ``gloo`` and compare against ``torch.std`` of an explicit stack.
"""
from __future__ import annotations

import torch


class TorchCpuOps:
    @staticmethod
    def update(x, mean, m2, count):
        if count == 1:
            mean.copy_(x)
            m2.zero_()
            return
        d = x - mean
        mean.add_(d * torch.tensor(1.0 / count, dtype=torch.float32))
        m2.add_(d * (x - mean))

    @staticmethod
    def merge(mean_a, m2_a, count_a, mean_b, m2_b, count_b):
        n = count_a + count_b
        d = mean_b - mean_a
        mean_a.add_(d * torch.tensor(count_b / n, dtype=torch.float32))
        m2_a.add_(m2_b + d * d * torch.tensor(count_a * count_b / n, dtype=torch.float32))

    @staticmethod
    def std(m2, count):
        return torch.sqrt(m2 * torch.tensor(1.0 / (count - 1), dtype=torch.float32))

    @staticmethod
    def sqerr(x, y, acc, first):
        d = (x - y.reshape(x.shape)) ** 2
        if first:
            acc.copy_(d)
        else:
            acc.add_(d)

    @staticmethod
    def global_ncc(a, v, scale_a, scale_v, square_a):
        import numpy as np
        a = a * a if square_a else a
        r = ref_global_ncc((a * torch.tensor(scale_a, dtype=torch.float32)).numpy(),
                           (v * torch.tensor(scale_v, dtype=torch.float32)).numpy())
        return torch.tensor([r, float(np.mean(a.numpy() * np.float32(scale_a)))], dtype=torch.float32)


def ref_global_ncc(a, v):
    """Restatement of Evaluate.ncc (evaluate.py:334-353, zero_norm=True) on numpy arrays, fp32 like the reference:
    both maps flattened and centred; the first divided by (population std x length + 1e-15), the second by
    (population std + 1e-15); the result is their dot product (np.correlate in 'valid' mode of equal lengths)."""
    import numpy as np
    a = np.asarray(a).reshape(-1)
    v = np.asarray(v).reshape(-1)
    eps = 1e-15
    an = (a - np.mean(a)) / (np.std(a) * len(a) + eps)
    vn = (v - np.mean(v)) / (np.std(v) + eps)
    return float(np.dot(an, vn))
