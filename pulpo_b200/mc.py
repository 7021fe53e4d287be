"""Monte-Carlo uncertainty maps, sharded by deformation sample (BASELINE.json config 3).

Reference semantics (``Evaluate.predict``, evaluate.py:205-251, ndims == 3 branch): draw N samples
one by one, stack ``outputs[l]`` / ``final_dfs[l]`` / ``individual_dfs[l]`` into ``[N, C, *S]``
tensors (~14 GB at N = 128), then ``torch.mean(torch.std(stack, axis=0), axis=0)`` per level;
``uncertainty`` (evaluate.py:1534-1545) squares the std into a variance map and compares it with the
per-voxel MSE ``mean((all_moved - y)**2, axis=0)``.

Here no stack exists.  Every (pair, sample) is an independent hot-path instance, so the N samples
are dealt to the ranks; each rank streams its samples through per-voxel Welford states
``(count, mean, M2)`` (``pulpo_moments_update``), the partial states are reduced to rank 0 over a binomial tree
(NCCL send/recv over NVLink: every rank sends its 8 B/voxel/channel once; a Chan merge,
``pulpo_moments_merge``, at every hop), and ``std = sqrt(M2 / (N - 1))`` (``pulpo_moments_std``), channel mean
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

    @staticmethod
    def sqerr(x, y, acc, first):
        PF.sqerr_update(x, y, acc, first)

    @staticmethod
    def merge_std(mean_parts, m2_parts, counts, chunk, out):
        return PF.moments_merge_std(mean_parts, m2_parts, counts, chunk, out)

    @staticmethod
    def global_ncc(a, v, scale_a, scale_v, square_a):
        return PF.global_ncc(a, v, scale_a, scale_v, square_a)


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


class MCSqErr:
    """Streaming per-voxel sum of squared errors of the MC samples against the fixed image ``y``:
    ``mse() == torch.mean((all_moved - y)**2, axis=0)`` of evaluate.py:1538 without the sample stack."""

    def __init__(self, shape, device, ops=None):
        self.ops = ops if ops is not None else _KernelOps
        self.count = 0
        self.acc = torch.zeros(tuple(shape), dtype=torch.float32, device=device)

    def update(self, x: torch.Tensor, y: torch.Tensor) -> None:
        if tuple(x.shape) != tuple(self.acc.shape) or x.numel() != y.numel():
            raise RuntimeError("MCSqErr.update: sample %s, target %s, state %s"
                               % (tuple(x.shape), tuple(y.shape), tuple(self.acc.shape)))
        self.ops.sqerr(x.detach().contiguous(), y.detach().contiguous(), self.acc, self.count == 0)
        self.count += 1

    def mse(self) -> torch.Tensor:
        if self.count < 1:
            raise RuntimeError("MCSqErr.mse needs at least 1 sample")
        return self.acc / float(self.count)


def uncertainty_metrics(moved: MCMoments, sqerr: MCSqErr) -> Dict[str, object]:
    """The per-pair numbers of ``Evaluate.uncertainty`` (evaluate.py:1534-1545): variance map
    ``moved_std ** 2``, MSE map, ``ncc(var, mse)`` (evaluate.py:334-353) and ``var.mean()``.
    One extra pass over the two maps: squaring the std and scaling the error sums ride in the NCC kernel."""
    if moved.count != sqerr.count:
        raise RuntimeError("uncertainty_metrics: %d samples in the moments, %d in the squared errors"
                           % (moved.count, sqerr.count))
    std = moved.std_channel_mean()                       # [*S]
    r = moved.ops.global_ncc(std, sqerr.acc.reshape(std.shape), 1.0, 1.0 / float(sqerr.count), True)
    return {"var": std ** 2, "mse": sqerr.mse().reshape(std.shape), "ncc": r[0], "var_mean": r[1]}


class StreamingStats:
    """Statistics of ALL tracked maps of an MC loop, updated by ONE graph-capturable launch per sample.

    ``maps``: {name: tensor} -- the STATIC buffers one sample's results land in (e.g. ``plan.moved[0][0]``);
    ``targets``: {name: fixed image} for the maps whose squared error is streamed too.  ``update()`` enqueues
    ``pulpo_moments_update_multi`` + ``pulpo_counter_add`` on the current stream (the sample count lives on the
    device, so the pair can be captured in the same CUDA graph as the sample's forward pass); the host only
    tallies how many times it ran.  ``states()`` hands the result out as ``MCMoments`` / ``MCSqErr`` objects that
    ``merge_across_ranks`` and ``uncertainty_metrics`` take."""

    def __init__(self, maps: Dict[str, torch.Tensor], targets: Optional[Dict[str, torch.Tensor]] = None):
        import ctypes
        from . import _lib
        targets = targets or {}
        self.names = list(maps.keys())
        self.maps = {n: maps[n] for n in self.names}
        dev = next(iter(maps.values())).device
        self.dev = dev
        # one flat buffer per statistic (padded so that any world size up to 16 splits it evenly): the multi-GPU
        # reduction exchanges contiguous slices of these without any packing pass
        self.offsets, off = {}, 0
        for n in self.names:
            self.offsets[n] = (off, maps[n].numel())
            off += maps[n].numel()
        self.total = off
        pad = -(-off // 5040) * 5040
        self.flat = torch.zeros(2, pad, dtype=torch.float32, device=dev)           # [mean | M2]
        self.mean = {n: self.flat[0, o:o + k].view(maps[n].shape) for n, (o, k) in self.offsets.items()}
        self.m2 = {n: self.flat[1, o:o + k].view(maps[n].shape) for n, (o, k) in self.offsets.items()}
        self.acc_offsets, off = {}, 0
        for n in self.names:
            if n in targets:
                self.acc_offsets[n] = (off, maps[n].numel())
                off += maps[n].numel()
        self.acc_flat = torch.zeros(max(-(-off // 5040) * 5040, 5040), dtype=torch.float32, device=dev)
        self.acc = {n: self.acc_flat[o:o + k].view(maps[n].shape) for n, (o, k) in self.acc_offsets.items()}
        self.targets = {n: targets[n].contiguous() for n in self.acc}
        self.count_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.count = 0
        arr = (_lib.MomentsMap * len(self.names))()
        for k, n in enumerate(self.names):
            t = maps[n]
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise RuntimeError("StreamingStats: map %r must be a contiguous fp32 CUDA tensor" % n)
            if n in self.acc and self.targets[n].numel() != t.numel():
                raise RuntimeError("StreamingStats: target of %r has %d elements, the map %d" % (n, self.targets[n].numel(), t.numel()))
            arr[k] = _lib.MomentsMap(t.data_ptr(), self.mean[n].data_ptr(), self.m2[n].data_ptr(),
                                     self.targets[n].data_ptr() if n in self.acc else None,
                                     self.acc[n].data_ptr() if n in self.acc else None, t.numel())
        self._arr, self._vp = arr, ctypes.c_void_p

    def reset(self):
        from . import _lib
        _lib.check(_lib.lib().pulpo_counter_add(self._vp(self.count_dev.data_ptr()), 0, 1,
                                                self._vp(torch.cuda.current_stream(self.dev).cuda_stream)), "counter reset")
        self.count = 0

    def update(self):
        """Enqueue the update for the sample currently in the map buffers (graph-capturable; does not touch ``count``)."""
        from . import _lib
        lib, st = _lib.lib(), self._vp(torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(lib.pulpo_moments_update_multi(self._arr, len(self.names), self._vp(self.count_dev.data_ptr()), st), "moments_update_multi")
        _lib.check(lib.pulpo_counter_add(self._vp(self.count_dev.data_ptr()), 1, 0, st), "counter_add")

    def reduce_to_maps(self, counts: List[int], group=None, dst: int = 0):
        """See ``reduce_flat_stats``: ``{name: std_channel_mean}`` + ``{name + ":mse"}`` on rank ``dst``."""
        shapes = {n: tuple(self.maps[n].shape) for n in self.names}
        if getattr(self, "_recv", None) is None:
            self._recv = torch.empty_like(self.flat)       # receive buffer of the all_to_all, kept across jobs
        return reduce_flat_stats(self.flat, self.acc_flat, self.offsets, self.acc_offsets, shapes, counts, _KernelOps,
                                 group=group, dst=dst, recv=self._recv)

    def states(self) -> Dict[str, "MCMoments | MCSqErr"]:
        out: Dict[str, object] = {}
        for n in self.names:
            st = MCMoments.__new__(MCMoments)
            st.ops, st.count, st.mean, st.m2 = _KernelOps, self.count, self.mean[n], self.m2[n]
            out[n] = st
            if n in self.acc:
                sq = MCSqErr.__new__(MCSqErr)
                sq.ops, sq.count, sq.acc = _KernelOps, self.count, self.acc[n]
                out[n + ":sqerr"] = sq
        return out


class PhiloxSampler:
    """gauss_sampler (src/network_blocks.py:7-8) of all levels of an MC sample as ONE graph-capturable launch:
    ``z[l] = mu[l] + sigma[l] * eps`` with eps from a counter-based Philox stream keyed by ``(seed, sample id)``, so
    sample *i* is the same noise on whatever rank draws it.  The sample id is ``first_id + id_stride * count`` where
    ``count`` is the device-side sample counter of a ``StreamingStats`` (or 0): captured together with the forward pass
    and the statistics update, a single graph replay runs a whole sample without any host-side RNG call."""

    def __init__(self, mu: Dict[int, torch.Tensor], sigma: Dict[int, torch.Tensor], seed: int = 0, first_id: int = 0,
                 id_stride: int = 1, count_dev: Optional[torch.Tensor] = None, dump_noise: bool = False):
        import ctypes
        from . import _lib
        self.levels = sorted(mu.keys())
        self.mu = {l: mu[l].contiguous() for l in self.levels}
        self.sigma = {l: sigma[l].contiguous() for l in self.levels}
        self.z = {l: torch.empty_like(self.mu[l]) for l in self.levels}
        self.eps = {l: torch.empty_like(self.mu[l]) for l in self.levels} if dump_noise else None
        self.dev = self.mu[self.levels[0]].device
        self.seed, self.first_id, self.id_stride = int(seed), int(first_id), int(id_stride)
        self.count_dev = count_dev
        arr = (_lib.GaussLevel * len(self.levels))()
        for k, l in enumerate(self.levels):
            if not (self.mu[l].is_cuda and self.mu[l].dtype == torch.float32 and self.sigma[l].shape == self.mu[l].shape):
                raise RuntimeError("PhiloxSampler: mu / sigma of level %d must be fp32 CUDA tensors of one shape" % l)
            arr[k] = _lib.GaussLevel(self.mu[l].data_ptr(), self.sigma[l].data_ptr(), self.z[l].data_ptr(),
                                     self.eps[l].data_ptr() if dump_noise else None, self.mu[l].numel())
        self._arr, self._vp = arr, ctypes.c_void_p

    def draw(self, sample_id: Optional[int] = None):
        """Enqueue the sampling of the next sample (device counter) or of an explicit ``sample_id`` (not capturable)."""
        from . import _lib
        st = self._vp(torch.cuda.current_stream(self.dev).cuda_stream)
        if sample_id is None:
            cnt = self._vp(self.count_dev.data_ptr()) if self.count_dev is not None else None
            first, stride = self.first_id, self.id_stride
        else:
            cnt, first, stride = None, int(sample_id), 0
        _lib.check(_lib.lib().pulpo_gauss_sample_multi(self._arr, len(self.levels), self.seed, cnt, first, stride, 1.0, st),
                   "gauss_sample_multi")
        return self.z


def reduce_flat_stats(flat, acc_flat, offsets, acc_offsets, shapes, counts, ops, group=None, dst: int = 0, recv=None):
    """Multi-GPU reduction of flat MC statistics straight to the reported maps (evaluate.py:243-251, 1538), sized for
    NVSwitch.  ``flat``: [2, T] (means | M2s) of all tracked maps back to back, ``acc_flat``: [TA] squared-error sums,
    both padded so that the world size divides them; ``counts[r]`` = samples rank r ran (known on the host).

    Every rank owns 1/W of the flat state: two ``all_to_all`` (means, M2s) hand rank r slice r of every rank's state --
    each rank sends its state once, all links busy at once, no packing pass -- and a ``reduce_scatter`` sums the
    squared errors; the rank Chan-merges its W partial slices in rank order (deterministic), turns them into the
    unbiased std / the MSE, and ONE ``gather`` brings those result slices to ``dst``, which forms the channel means."""
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    W = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    N = int(sum(counts))
    T, TA = flat.shape[1], acc_flat.numel()
    dev = flat.device
    if T % W or TA % W:
        raise RuntimeError("reduce_flat_stats: world size %d does not divide the padded state (%d, %d)" % (W, T, TA))
    ch, cha = T // W, TA // W
    if multi:
        if recv is None:
            recv = torch.empty_like(flat)
        dist.all_to_all_single(recv[0], flat[0], group=group)
        dist.all_to_all_single(recv[1], flat[1], group=group)
        acc = torch.empty(cha, dtype=torch.float32, device=dev)
        try:
            dist.reduce_scatter_tensor(acc, acc_flat, group=group)
        except (RuntimeError, NotImplementedError):          # gloo (CPU tests): no reduce_scatter
            tmp = acc_flat.clone()
            dist.all_reduce(tmp, group=group)
            acc = tmp[rank * cha:(rank + 1) * cha]
        mine = torch.empty(ch + cha, dtype=torch.float32, device=dev)
        if hasattr(ops, "merge_std"):          # one pass: Chan merge of the W partial slices in rank order + unbiased std
            ops.merge_std(recv[0], recv[1], [int(c) for c in counts], ch, mine[:ch])
        else:
            mean, m2, seen = recv[0, :ch], recv[1, :ch], int(counts[0])     # rank 0's partial slice, merged into in place
            for r in range(1, W):
                if counts[r] > 0:
                    if seen == 0:
                        mean.copy_(recv[0, r * ch:(r + 1) * ch]); m2.copy_(recv[1, r * ch:(r + 1) * ch])
                    else:
                        ops.merge(mean, m2, seen, recv[0, r * ch:(r + 1) * ch], recv[1, r * ch:(r + 1) * ch], int(counts[r]))
                    seen += int(counts[r])
            mine[:ch] = ops.std(m2.contiguous(), N)
        torch.mul(acc, 1.0 / N, out=mine[ch:])
    else:
        mine = torch.cat([ops.std(flat[1].contiguous(), N), acc_flat * (1.0 / N)])
    if multi:
        gathered = torch.empty(W, ch + cha, dtype=torch.float32, device=dev) if rank == dst else None
        dist.gather(mine, list(gathered.unbind(0)) if rank == dst else None,
                    dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
        if rank != dst:
            return None
        std_flat = gathered[:, :ch].reshape(-1)
        mse_flat = gathered[:, ch:].reshape(-1)
    else:
        std_flat, mse_flat = mine[:T], mine[T:]
    out = {}
    for n, (o, k) in offsets.items():
        out[n] = std_flat[o:o + k].view(shapes[n]).mean(dim=0)
    for n, (o, k) in acc_offsets.items():
        out[n + ":mse"] = mse_flat[o:o + k].view(shapes[n])[0]
    return out


def sliced_uncertainty(states: Dict[str, "MCMoments | MCSqErr"], group=None, dst: int = 0, device=None):
    """Multi-GPU reduction of MC statistics straight to the maps ``Evaluate.predict / uncertainty`` report
    (evaluate.py:243-251, 1534-1545), sized for NVSwitch: every rank owns 1/W of the voxels.

    1. one ``all_to_all``: rank r receives voxel slice r of every rank's ``(mean, M2)`` (and squared-error sums) --
       each rank sends its state once, all links busy at the same time (a tree reduce to one rank serialises
       log2(W) hops of the full state);
    2. each rank Chan-merges the W partial slices in rank order (deterministic) and turns them into what is reported:
       per-voxel unbiased std averaged over channels, and the MSE;
    3. one ``gather`` of those result slices (1 float per voxel per map, 6x less than (mean, M2) of a 3-channel field)
       to rank ``dst``.
    Returns on ``dst``: ``{name: std_channel_mean [*S]}`` plus ``{name + ":mse": [*S]}`` for the squared-error states;
    ``None`` on the other ranks."""
    import torch.distributed as dist
    names = sorted(n for n in states if not isinstance(states[n], MCSqErr))
    sq_names = sorted(n for n in states if isinstance(states[n], MCSqErr))
    if device is None:
        device = states[names[0]].mean.device
    ops = states[names[0]].ops
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    W = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    if not multi:
        out = {n: states[n].std_channel_mean() for n in names}
        out.update({n.replace(":sqerr", ":mse"): states[n].mse() for n in sq_names})
        return out
    # layout of one rank's packed state: for every map [C, nvox] -> W slices of `chunk` voxels (zero padded)
    meta, total = [], 0
    for n in names:
        C, nvox = states[n].mean.shape[0], states[n].mean[0].numel()
        chunk = (nvox + W - 1) // W
        meta.append((n, C, nvox, chunk, total, 2 * C * chunk))
        total += 2 * C * chunk
    for n in sq_names:
        C, nvox = states[n].acc.shape[0], states[n].acc[0].numel()
        chunk = (nvox + W - 1) // W
        meta.append((n, C, nvox, chunk, total, C * chunk))
        total += C * chunk
    send = torch.zeros(W, total, dtype=torch.float32, device=device)
    for n, C, nvox, chunk, off, sz in meta:
        st = states[n]
        parts = [st.acc.reshape(C, nvox)] if isinstance(st, MCSqErr) else [st.mean.reshape(C, nvox), st.m2.reshape(C, nvox)]
        packed = torch.zeros(len(parts), C, W * chunk, dtype=torch.float32, device=device)
        for k, p in enumerate(parts):
            packed[k, :, :nvox] = p
        # [parts, C, W, chunk] -> [W, parts * C * chunk]
        send[:, off:off + sz] = packed.reshape(len(parts), C, W, chunk).permute(2, 0, 1, 3).reshape(W, sz)
    counts = torch.tensor([states[names[0]].count], dtype=torch.int64, device=device)
    all_counts = [torch.empty_like(counts) for _ in range(W)]
    dist.all_gather(all_counts, counts, group=group)
    all_counts = [int(c.item()) for c in all_counts]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    # merge this rank's slice of every map over the W partial states, in rank order
    res_parts, res_meta = [], []
    for n, C, nvox, chunk, off, sz in meta:
        if n in sq_names:
            acc = recv[:, off:off + sz].sum(dim=0) / float(sum(all_counts))
            res = acc.reshape(C, chunk).mean(dim=0) if C > 1 else acc.reshape(chunk)
        else:
            m = MCMoments((C, chunk), device, ops=ops)
            for r in range(W):
                blk = recv[r, off:off + sz].reshape(2, C, chunk)
                m.merge_state(blk[0], blk[1], all_counts[r])
            res = m.std_channel_mean()
        res_parts.append(res.reshape(-1))
        res_meta.append((n, nvox, chunk))
    mine = torch.cat(res_parts)
    gathered = [torch.empty_like(mine) for _ in range(W)] if rank == dst else None
    dist.gather(mine, gathered, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
    if rank != dst:
        return None
    out, pos = {}, 0
    for n, nvox, chunk in res_meta:
        full = torch.cat([g[pos:pos + chunk] for g in gathered])[:nvox]
        st = states[n]
        shape = tuple((st.acc if isinstance(st, MCSqErr) else st.mean).shape[1:])
        out[n.replace(":sqerr", ":mse")] = full.reshape(shape)
        pos += chunk
    return out


def _check_same_maps(local_meta, group):
    """Every rank must hold the same maps with the same shapes; a rank that drew no sample (empty explicit
    ``sample_ids``) may hold none.  Decided on ALL ranks before any tensor collective, so a mismatch raises
    everywhere instead of leaving the other ranks blocked in a collective."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    metas = [None] * world
    dist.all_gather_object(metas, local_meta, group=group)
    shapes = {}
    for r, m in enumerate(metas):
        for name, (shape, _count) in m.items():
            if shapes.setdefault(name, tuple(shape)) != tuple(shape):
                raise RuntimeError("merge_across_ranks: map %r has shape %s on rank %d but %s elsewhere"
                                   % (name, tuple(shape), r, shapes[name]))
    return shapes, metas


def merge_across_ranks(local: Dict[str, "MCMoments | MCSqErr"], group=None, dst: Optional[int] = 0, device=None,
                       ops=None) -> Dict[str, "MCMoments | MCSqErr"]:
    """Reduce every rank's partial state to rank ``dst`` (``dst=None``: rank 0, then broadcast to all).

    ``MCMoments``: binomial-tree reduction of ``(count, mean, M2)`` with a Chan merge at every hop -- each rank
    sends its state ONCE (8 B/voxel/channel over NVLink, log2(world) hops deep) and needs one receive buffer,
    instead of all-gathering world x (mean, M2) to every rank; the merge order is fixed by the world size, so the
    result is deterministic.  ``MCSqErr``: one ``reduce(SUM)`` of the error sums.  Ranks other than ``dst``
    return their local (partial) states."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    root = 0 if dst is None else int(dst)
    meta = {n: (tuple(st.acc.shape if isinstance(st, MCSqErr) else st.mean.shape), st.count) for n, st in local.items()}
    kinds = {n: isinstance(st, MCSqErr) for n, st in local.items()}
    shapes, metas = _check_same_maps({n: (meta[n][0], meta[n][1], ) for n in meta}, group)
    all_kinds = [None] * world
    dist.all_gather_object(all_kinds, kinds, group=group)
    for k in all_kinds:
        for n, v in k.items():
            kinds.setdefault(n, v)
    if device is None:
        if not local:
            raise RuntimeError("merge_across_ranks: a rank without local maps must pass `device`")
        any_st = next(iter(local.values()))
        device = (any_st.acc if isinstance(any_st, MCSqErr) else any_st.mean).device
    if ops is None and local:
        ops = next(iter(local.values())).ops
    rel = (rank - root) % world
    merged = {}
    for n in sorted(shapes):
        st = local.get(n)
        if st is None:                                    # this rank drew no sample: zero-count state
            st = MCSqErr(shapes[n], device, ops=ops) if kinds[n] else MCMoments(shapes[n], device, ops=ops)
        if kinds[n]:
            total = torch.tensor([st.count], dtype=torch.int64, device=device)
            out = MCSqErr(shapes[n], device, ops=st.ops)
            out.acc.copy_(st.acc)
            if dst is None:
                dist.all_reduce(out.acc, group=group)
                dist.all_reduce(total, group=group)
            else:
                dist.reduce(out.acc, dst=dist.get_global_rank(group, root) if group is not None else root, group=group)
                dist.all_reduce(total, group=group)
            out.count = int(total.item())
            merged[n] = out if (dst is None or rank == root) else st
            continue
        acc = MCMoments(shapes[n], device, ops=st.ops)
        acc.merge_state(st.mean, st.m2, st.count)
        buf = torch.empty((2,) + tuple(shapes[n]), dtype=torch.float32, device=device)
        cnt = torch.zeros(1, dtype=torch.int64, device=device)
        step = 1
        while step < world:
            if rel % (2 * step) == step:                  # sender: hand the partial state down the tree, then done
                peer = (rel - step + root) % world
                peer = dist.get_global_rank(group, peer) if group is not None else peer
                dist.send(torch.tensor([acc.count], dtype=torch.int64, device=device), dst=peer, group=group)
                dist.send(torch.stack([acc.mean, acc.m2]), dst=peer, group=group)
                break
            if rel % (2 * step) == 0 and rel + step < world:
                peer = (rel + step + root) % world
                peer = dist.get_global_rank(group, peer) if group is not None else peer
                dist.recv(cnt, src=peer, group=group)
                dist.recv(buf, src=peer, group=group)
                acc.merge_state(buf[0], buf[1], int(cnt.item()))
            step *= 2
        if dst is None:
            packed = torch.stack([acc.mean, acc.m2])
            total = torch.tensor([acc.count], dtype=torch.int64, device=device)
            src = dist.get_global_rank(group, root) if group is not None else root
            dist.broadcast(packed, src=src, group=group)
            dist.broadcast(total, src=src, group=group)
            acc.mean.copy_(packed[0]); acc.m2.copy_(packed[1]); acc.count = int(total.item())
            merged[n] = acc
        else:
            merged[n] = acc if rank == root else st
    return merged


def mc_uncertainty(sample_fn: Callable[[int, torch.Generator], Dict[str, torch.Tensor]], num_samples: int,
                   seed0: int = 0, device=None, group=None, dst: Optional[int] = 0, ops=None,
                   sample_ids: Optional[Iterable[int]] = None,
                   targets: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, "MCMoments | MCSqErr"]:
    """Run this rank's share of ``num_samples`` MC samples and merge the per-voxel statistics.

    ``sample_fn(sample_id, generator)`` runs one deformation sample (``gauss_sampler(mu, sigma,
    generator=generator)`` -> decode -> warp) under ``torch.no_grad()`` and returns the maps to
    track, e.g. ``{"moved0": outputs[0][0], "final0": final_dfs[0][0]}`` (each ``[C, *S]``).
    ``targets`` maps a map name to the fixed image it is compared with: for those maps the squared errors are
    streamed too and returned under ``name + ":sqerr"`` (the MSE map of evaluate.py:1538).
    Returns ``{name: MCMoments | MCSqErr}`` holding all ``num_samples`` samples on rank ``dst``.
    """
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    if world > 1 and num_samples < world and sample_ids is None:
        raise ValueError("mc_uncertainty: %d samples cannot be dealt to %d ranks (every rank needs at least one)"
                         % (num_samples, world))
    ids = list(sample_ids) if sample_ids is not None else shard_samples(num_samples, rank, world)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    states: Dict[str, object] = {}
    with torch.no_grad():
        for i in ids:
            maps = sample_fn(i, sample_generator(seed0, i, device))
            for name, t in maps.items():
                if name not in states:
                    states[name] = MCMoments(t.shape, t.device, ops=ops)
                states[name].update(t)
                if targets is not None and name in targets:
                    key = name + ":sqerr"
                    if key not in states:
                        states[key] = MCSqErr(t.shape, t.device, ops=ops)
                    states[key].update(t, targets[name])
    return merge_across_ranks(states, group=group, dst=dst, device=device, ops=ops) if world > 1 else states
