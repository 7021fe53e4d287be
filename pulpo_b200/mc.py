"""Monte-Carlo uncertainty maps, sharded by deformation sample (BASELINE.json config 3).

Reference semantics (``Evaluate.predict``, evaluate.py:205-251, ndims == 3 branch): draw N samples
one by one, stack ``outputs[l]`` / ``final_dfs[l]`` / ``individual_dfs[l]`` into ``[N, C, *S]``
tensors (~14 GB at N = 128), then ``torch.mean(torch.std(stack, axis=0), axis=0)`` per level;
``uncertainty`` (evaluate.py:1534-1545) squares the std into a variance map and compares it with the
per-voxel MSE ``mean((all_moved - y)**2, axis=0)``.

Here no stack exists.  Every (pair, sample) is an independent hot-path instance, so the N samples
are dealt to the ranks; each rank streams its samples through per-voxel Welford states
``(count, mean, M2)`` (``pulpo_moments_update``), the partial states are all-gathered
(NCCL over NVLink: 8 B/voxel/channel per rank) and Chan-merged in rank order
(``pulpo_moments_merge``), and ``std = sqrt(M2 / (N - 1))`` (``pulpo_moments_std``), channel mean
and square follow.  Sample *i* always draws its noise from ``seed0 + i`` (see ``sample_generator``),
so any sharding -- 1, 2, 4 or 8 ranks -- sees the same N samples and the merged maps agree up to
fp32 merge order.

Reference quirk, flagged not reproduced: evaluate.py:238 averages ``individual_dfs`` of the *last*
sample only (it indexes the loop variable instead of ``all_individual_dfs``); ``MCMoments.mean`` is
the true mean over all samples.

The numerical steps go through libpulpo_b200 (CUDA only, no fallback).  ``ops`` exists so the
host logic (sharding, gather, merge order) can be tested on CPU with ``gloo``: tests inject
the oracle's implementation there.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional

import torch

from . import functional as PF


class _KernelOps:
    """Default numerical back end: the C-ABI moments kernels (raise on non-CUDA tensors)."""

    @staticmethod
    def update(x, mean, m2, count):
        PF.moments_update(x, mean, m2, count)

    @staticmethod
    def merge(mean_a, m2_a, count_a, mean_b, m2_b, count_b):
        PF.moments_merge(mean_a, m2_a, count_a, mean_b, m2_b, count_b)

    @staticmethod
    def std(m2, count):
        return PF.moments_std(m2, count)


def shard_samples(num_samples: int, rank: int, world: int) -> List[int]:
    """Sample ids of ``rank``: round-robin, so every rank gets floor/ceil(N / world) samples."""
    if num_samples < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("shard_samples: bad arguments (N=%r, rank=%r, world=%r)" % (num_samples, rank, world))
    return list(range(rank, num_samples, world))


def sample_generator(seed0: int, sample_id: int, device) -> torch.Generator:
    """Philox stream of MC sample ``sample_id``: independent of which rank draws it."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed0) + int(sample_id))
    return g


class MCMoments:
    """Streaming per-voxel moments of one map ``[C, *S]`` (any shape; treated as flat)."""

    def __init__(self, shape, device, ops=None):
        self.ops = ops if ops is not None else _KernelOps
        self.count = 0
        self.mean = torch.zeros(tuple(shape), dtype=torch.float32, device=device)
        self.m2 = torch.zeros(tuple(shape), dtype=torch.float32, device=device)

    def update(self, x: torch.Tensor) -> None:
        if tuple(x.shape) != tuple(self.mean.shape):
            raise RuntimeError("MCMoments.update: sample of shape %s, state of shape %s"
                               % (tuple(x.shape), tuple(self.mean.shape)))
        self.count += 1
        self.ops.update(x.detach().contiguous(), self.mean, self.m2, self.count)

    def merge_state(self, mean_b: torch.Tensor, m2_b: torch.Tensor, count_b: int) -> None:
        if count_b == 0:
            return
        if self.count == 0:
            self.mean.copy_(mean_b)
            self.m2.copy_(m2_b)
        else:
            self.ops.merge(self.mean, self.m2, self.count, mean_b.contiguous(), m2_b.contiguous(), int(count_b))
        self.count += int(count_b)

    def std(self) -> torch.Tensor:
        """Unbiased per-voxel std over the samples seen (torch.std(axis=0) of evaluate.py:243)."""
        if self.count < 2:
            raise RuntimeError("MCMoments.std needs at least 2 samples (got %d)" % self.count)
        return self.ops.std(self.m2, self.count)

    def std_channel_mean(self) -> torch.Tensor:
        """``torch.mean(torch.std(stack, axis=0), axis=0)`` -- evaluate.py:243-251."""
        return self.std().mean(dim=0)

    def variance_map(self) -> torch.Tensor:
        """``moved_std ** 2`` of evaluate.py:1540."""
        return self.std_channel_mean() ** 2


def merge_across_ranks(local: Dict[str, MCMoments], group=None, dst: Optional[int] = 0) -> Dict[str, MCMoments]:
    """All-gather every rank's ``(count, mean, M2)`` and Chan-merge them in rank order (deterministic
    for a given world size).  With ``dst`` set only that rank merges (the others return their
    local states untouched); ``dst=None`` merges everywhere."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    names = sorted(local.keys())
    dev = local[names[0]].mean.device
    counts = torch.tensor([local[n].count for n in names], dtype=torch.int64, device=dev)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    merged = {}
    for j, n in enumerate(names):
        st = local[n]
        packed = torch.stack([st.mean, st.m2])                       # one collective per map
        parts = [torch.empty_like(packed) for _ in range(world)]
        dist.all_gather(parts, packed, group=group)
        if dst is not None and rank != dst:
            merged[n] = st
            continue
        out = MCMoments(st.mean.shape, dev, ops=st.ops)
        for r in range(world):                                        # fixed order: rank 0, 1, 2, ...
            out.merge_state(parts[r][0], parts[r][1], int(all_counts[r][j].item()))
        merged[n] = out
    return merged


def mc_uncertainty(sample_fn: Callable[[int, torch.Generator], Dict[str, torch.Tensor]], num_samples: int,
                   seed0: int = 0, device=None, group=None, dst: Optional[int] = 0, ops=None,
                   sample_ids: Optional[Iterable[int]] = None) -> Dict[str, MCMoments]:
    """Run this rank's share of ``num_samples`` MC samples and merge the per-voxel moments.

    ``sample_fn(sample_id, generator)`` runs one deformation sample (``gauss_sampler(mu, sigma,
    generator=generator)`` -> decode -> warp) under ``torch.no_grad()`` and returns the maps to
    track, e.g. ``{"moved0": outputs[0][0], "final0": final_dfs[0][0]}`` (each ``[C, *S]``).
    Returns ``{name: MCMoments}`` holding all ``num_samples`` samples on rank ``dst``.
    """
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    if world > 1 and num_samples < world:
        raise ValueError("mc_uncertainty: %d samples cannot be dealt to %d ranks (every rank needs at least one)"
                         % (num_samples, world))
    ids = list(sample_ids) if sample_ids is not None else shard_samples(num_samples, rank, world)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    states: Dict[str, MCMoments] = {}
    with torch.no_grad():
        for i in ids:
            maps = sample_fn(i, sample_generator(seed0, i, device))
            for name, t in maps.items():
                if name not in states:
                    states[name] = MCMoments(t.shape, t.device, ops=ops)
                states[name].update(t)
    return merge_across_ranks(states, group=group, dst=dst) if world > 1 else states
