"""World-size-2 `gloo` tests (CPU) of the multi-rank host logic:
MC-sample sharding + all-gather + Chan merge of the per-voxel moments (config 3), and the
pair-sharded loss all-reduce (config 5: every loss term is a batch mean, so averaging equal
shards reproduces the global-batch loss).  The numerical kernels are CUDA-only; here the
oracle's CPU implementation is injected through the `ops` seam of pulpo_b200.mc."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.moments_ref import TorchCpuOps
from pulpo_b200 import mc

SHAPE = (3, 6, 5, 4)
N_SAMPLES = 11


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sample(i, gen):
    # deterministic in (seed0 + i) only: any rank that draws sample i gets the same tensor
    base = torch.linspace(-1.0, 1.0, steps=int(torch.tensor(SHAPE).prod())).reshape(SHAPE)
    noise = torch.randn(SHAPE, generator=gen)
    return {"final0": base + 0.5 * noise, "moved0": (0.1 * noise[:1]).abs()}


def _target():
    return torch.linspace(0.0, 0.3, steps=int(torch.tensor(SHAPE[1:]).prod())).reshape((1,) + SHAPE[1:])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = mc.mc_uncertainty(_sample, N_SAMPLES, seed0=123, device=torch.device("cpu"), ops=TorchCpuOps, dst=0,
                                targets={"moved0": _target()})
        # config 5: per-rank batch-mean losses averaged over equal shards == global batch mean
        g = torch.Generator().manual_seed(7)
        per_pair = torch.rand(4, generator=g)              # 4 pairs, 2 per rank
        local = per_pair[rank * 2:(rank + 1) * 2].mean()
        t = local.clone()
        dist.all_reduce(t)
        t /= world
        if rank == 0:
            torch.save({"count": res["final0"].count, "std": res["final0"].std(), "mean": res["final0"].mean,
                        "var_map": res["moved0"].variance_map(), "loss": t, "loss_ref": per_pair.mean(),
                        "mse": res["moved0:sqerr"].mse(), "sq_count": res["moved0:sqerr"].count,
                        "metrics": mc.uncertainty_metrics(res["moved0"], res["moved0:sqerr"])},
                       os.path.join(out_dir, "r0.pt"))
        else:
            assert res["final0"].count == len(mc.shard_samples(N_SAMPLES, rank, world))   # untouched local state
    finally:
        dist.destroy_process_group()


def test_shard_samples_partition():
    for n, w in [(128, 8), (11, 2), (5, 4), (3, 3)]:
        ids = [mc.shard_samples(n, r, w) for r in range(w)]
        assert sorted(i for part in ids for i in part) == list(range(n))
        assert max(len(p) for p in ids) - min(len(p) for p in ids) <= 1
    with pytest.raises(ValueError):
        mc.shard_samples(4, 2, 2)


def test_mc_moments_single_process_matches_torch_std():
    states = mc.mc_uncertainty(_sample, N_SAMPLES, seed0=123, device=torch.device("cpu"), ops=TorchCpuOps)
    stack = torch.stack([_sample(i, mc.sample_generator(123, i, "cpu"))["final0"] for i in range(N_SAMPLES)])
    assert states["final0"].count == N_SAMPLES
    torch.testing.assert_close(states["final0"].std(), stack.std(dim=0), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(states["final0"].mean, stack.mean(dim=0), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(states["final0"].std_channel_mean(), stack.std(dim=0).mean(dim=0), rtol=1e-5, atol=1e-6)


def test_product_ops_refuse_cpu_tensors():
    st = mc.MCMoments(SHAPE, "cpu")          # default ops = the CUDA kernels
    with pytest.raises(RuntimeError):
        st.update(torch.zeros(SHAPE))


@pytest.mark.timeout(180)
def test_mc_sharded_over_two_gloo_ranks_matches_unsharded(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "r0.pt"))
    stack = torch.stack([_sample(i, mc.sample_generator(123, i, "cpu"))["final0"] for i in range(N_SAMPLES)])
    moved = torch.stack([_sample(i, mc.sample_generator(123, i, "cpu"))["moved0"] for i in range(N_SAMPLES)])
    assert got["count"] == N_SAMPLES
    torch.testing.assert_close(got["std"], stack.std(dim=0), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["mean"], stack.mean(dim=0), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["var_map"], moved.std(dim=0).mean(dim=0) ** 2, rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(got["loss"], got["loss_ref"], rtol=1e-6, atol=0)
    # f-3: MSE map and global NCC(var, mse) from the merged streaming states == the stack-based reference formulas
    from oracle.moments_ref import ref_global_ncc
    mse_ref = ((moved - _target()) ** 2).mean(dim=0)
    var_ref = moved.std(dim=0).mean(dim=0) ** 2
    assert got["sq_count"] == N_SAMPLES
    torch.testing.assert_close(got["mse"], mse_ref, rtol=1e-5, atol=1e-8)
    ncc_ref = ref_global_ncc(var_ref.numpy(), mse_ref.squeeze(0).numpy())
    assert abs(float(got["metrics"]["ncc"]) - ncc_ref) <= 1e-4 * abs(ncc_ref) + 1e-6


def _worker_tree(rank, world, port, out_dir):
    """3 ranks (not a power of two) with an explicit sample deal that leaves rank 2 without samples: the empty rank
    still takes part in every collective with zero-count states, and dst=None leaves the merged state on all ranks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = {0: [0, 1, 2, 3, 4, 5], 1: [6, 7, 8, 9, 10], 2: []}[rank]
        res = mc.mc_uncertainty(_sample, N_SAMPLES, seed0=123, device=torch.device("cpu"), ops=TorchCpuOps, dst=None,
                                sample_ids=ids, targets={"moved0": _target()})
        torch.save({"count": res["final0"].count, "std": res["final0"].std(), "mse": res["moved0:sqerr"].mse()},
                   os.path.join(out_dir, "r%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_mc_tree_merge_three_ranks_one_empty(tmp_path):
    port = _free_port()
    mp.spawn(_worker_tree, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    stack = torch.stack([_sample(i, mc.sample_generator(123, i, "cpu"))["final0"] for i in range(N_SAMPLES)])
    moved = torch.stack([_sample(i, mc.sample_generator(123, i, "cpu"))["moved0"] for i in range(N_SAMPLES)])
    for r in range(3):
        got = torch.load(os.path.join(str(tmp_path), "r%d.pt" % r))
        assert got["count"] == N_SAMPLES
        torch.testing.assert_close(got["std"], stack.std(dim=0), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(got["mse"], ((moved - _target()) ** 2).mean(dim=0), rtol=1e-5, atol=1e-8)


def _worker_mismatch(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape = SHAPE if rank == 0 else (3, 6, 5, 5)
        st = mc.MCMoments(shape, "cpu", ops=TorchCpuOps)
        st.update(torch.zeros(shape))
        try:
            mc.merge_across_ranks({"m": st}, dst=0)
            ok = False
        except RuntimeError:
            ok = True                      # raised on EVERY rank, before any tensor collective: nobody hangs
        open(os.path.join(out_dir, "ok%d" % rank), "w").write(str(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_merge_shape_mismatch_raises_on_all_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_worker_mismatch, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(open(os.path.join(str(tmp_path), "ok%d" % r)).read() == "True" for r in range(2))



def _worker_sliced(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        states = mc.mc_uncertainty(_sample, N_SAMPLES, seed0=123, device=torch.device("cpu"), ops=TorchCpuOps, dst=0,
                                   sample_ids=mc.shard_samples(N_SAMPLES, rank, world), targets={"moved0": _target()})
        # mc_uncertainty already merged onto rank 0; for the sliced reduction start again from the per-rank partial states
        local = {}
        for i in mc.shard_samples(N_SAMPLES, rank, world):
            maps = _sample(i, mc.sample_generator(123, i, "cpu"))
            for n, t in maps.items():
                local.setdefault(n, mc.MCMoments(t.shape, "cpu", ops=TorchCpuOps)).update(t)
            local.setdefault("moved0:sqerr", mc.MCSqErr(maps["moved0"].shape, "cpu", ops=TorchCpuOps)).update(maps["moved0"], _target())
        res = mc.sliced_uncertainty(local, dst=0, device=torch.device("cpu"))
        if rank == 0:
            torch.save(res, os.path.join(out_dir, "sliced.pt"))
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sliced_all_to_all_reduction_three_ranks(tmp_path):
    """all_to_all of voxel slices + per-slice Chan merge + gather of the std maps == stack-based std / MSE
    (3 ranks: the slice size does not divide the voxel count)."""
    port = _free_port()
    mp.spawn(_worker_sliced, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    got = torch.load(os.path.join(str(tmp_path), "sliced.pt"))
    stack = torch.stack([_sample(i, mc.sample_generator(123, i, "cpu"))["final0"] for i in range(N_SAMPLES)])
    moved = torch.stack([_sample(i, mc.sample_generator(123, i, "cpu"))["moved0"] for i in range(N_SAMPLES)])
    torch.testing.assert_close(got["final0"], stack.std(dim=0).mean(dim=0), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["moved0"], moved.std(dim=0).mean(dim=0), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(got["moved0:mse"], ((moved - _target()) ** 2).mean(dim=0)[0], rtol=1e-5, atol=1e-8)



def _worker_flat(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # flat per-rank statistics as StreamingStats lays them out: [mean | M2] of both maps back to back, padded
        shapes = {"final0": SHAPE, "moved0": (1,) + SHAPE[1:]}
        offsets, off = {}, 0
        for n, sh in shapes.items():
            k = int(torch.tensor(sh).prod())
            offsets[n] = (off, k)
            off += k
        pad = -(-off // 5040) * 5040
        flat = torch.zeros(2, pad)
        acc_flat = torch.zeros(5040)
        acc_offsets = {"moved0": (0, offsets["moved0"][1])}
        ids = mc.shard_samples(N_SAMPLES, rank, world)
        for c, i in enumerate(ids):
            maps = _sample(i, mc.sample_generator(123, i, "cpu"))
            for n, (o, k) in offsets.items():
                TorchCpuOps.update(maps[n].reshape(-1), flat[0, o:o + k], flat[1, o:o + k], c + 1)
            TorchCpuOps.sqerr(maps["moved0"].reshape(-1), _target().reshape(-1), acc_flat[:acc_offsets["moved0"][1]], c == 0)
        counts = [len(mc.shard_samples(N_SAMPLES, r, world)) for r in range(world)]
        res = mc.reduce_flat_stats(flat, acc_flat, offsets, acc_offsets, shapes, counts, TorchCpuOps, dst=0)
        if rank == 0:
            torch.save(res, os.path.join(out_dir, "flat.pt"))
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.timeout(180)
def test_flat_all_to_all_reduction(tmp_path, world):
    """reduce_flat_stats (what StreamingStats.reduce_to_maps runs on the GPUs): all_to_all of 1/W slices, Chan merge
    per slice, gather of std / MSE -- against the stack-based numbers of evaluate.py:243-251, 1538."""
    port = _free_port()
    mp.spawn(_worker_flat, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(str(tmp_path), "flat.pt"))
    stack = torch.stack([_sample(i, mc.sample_generator(123, i, "cpu"))["final0"] for i in range(N_SAMPLES)])
    moved = torch.stack([_sample(i, mc.sample_generator(123, i, "cpu"))["moved0"] for i in range(N_SAMPLES)])
    torch.testing.assert_close(got["final0"], stack.std(dim=0).mean(dim=0), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["moved0"], moved.std(dim=0).mean(dim=0), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(got["moved0:mse"], ((moved - _target()) ** 2).mean(dim=0)[0], rtol=1e-5, atol=1e-8)
