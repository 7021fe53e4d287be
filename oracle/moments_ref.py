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
