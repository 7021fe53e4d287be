"""Two-GPU NCCL run of the MC-sample sharding (config 3) with the real kernels: the merged per-voxel
moments of 2 ranks x 4 samples equal the single-rank run over the same 8 samples (same per-sample
Philox seeds).  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

SIZE, TOTAL, LATENT, N = [32, 32, 32], 4, 3, 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sample_fn_factory(dev):
    from pulpo_b200 import synthetic as syn
    from pulpo_b200.models import combine_dfs
    from pulpo_b200.network_blocks import SpatialTransformer, gauss_sampler
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(SIZE, TOTAL, LATENT, seed=3)
    xc = x.to(dev)
    mu = {l: dfs[l].to(dev) for l in dfs}
    sg = {l: (0.3 * sgs[l]).to(dev) for l in dfs}
    st = SpatialTransformer(SIZE)

    def sample_fn(i, gen):
        v = {l: gauss_sampler(mu[l], sg[l], generator=gen) for l in mu}
        _, final = combine_dfs(v, SIZE)
        return {"final0": final[0][0], "moved0": st(final[0], xc)[0]}
    return sample_fn, y.to(dev)[0]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from pulpo_b200 import mc
        fn, y = _sample_fn_factory(dev)
        res = mc.mc_uncertainty(fn, N, seed0=11, device=dev, dst=0, targets={"moved0": y})
        if rank == 0:
            m = mc.uncertainty_metrics(res["moved0"], res["moved0:sqerr"])
            torch.save({"count": res["final0"].count, "std": res["final0"].std().cpu(), "var": res["moved0"].variance_map().cpu(),
                        "mse": m["mse"].cpu(), "ncc": float(m["ncc"])}, os.path.join(out_dir, "r0.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_mc_uncertainty_two_gpus_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pulpo_b200 import mc
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "r0.pt"))
    dev = torch.device("cuda", 0)
    fn, y = _sample_fn_factory(dev)
    one = mc.mc_uncertainty(fn, N, seed0=11, device=dev, targets={"moved0": y})
    m = mc.uncertainty_metrics(one["moved0"], one["moved0:sqerr"])
    assert got["count"] == N
    torch.testing.assert_close(got["mse"], m["mse"].cpu(), rtol=1e-4, atol=1e-7)
    assert abs(got["ncc"] - float(m["ncc"])) <= 1e-3 * abs(float(m["ncc"])) + 1e-6
    torch.testing.assert_close(got["std"], one["final0"].std().cpu(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got["var"], one["moved0"].variance_map().cpu(), rtol=1e-4, atol=1e-7)



def _graph_job(dev, rank, world, n_samples):
    """bench.py's mc128 job at small size: Philox sampler + forward plan + StreamingStats in one CUDA graph per sample,
    flat all-to-all reduction to rank 0."""
    from pulpo_b200 import mc, synthetic as syn
    from pulpo_b200.plan import HotPathPlan
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(SIZE, TOTAL, LATENT, seed=3)
    x, y = x.to(dev), y.to(dev)
    mu = {l: dfs[l].to(dev) for l in dfs}
    sg = {l: (0.3 * sgs[l]).to(dev) for l in dfs}
    plan = HotPathPlan(SIZE, TOTAL, LATENT, batch=1, device=dev, with_reg=False)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    sampler = mc.PhiloxSampler(mu, sg, seed=5, first_id=rank, id_stride=world, count_dev=cnt)
    plan.run_forward(x, sampler.z)
    bufs = {"moved0": plan.moved[0][0], "final0": plan.final[0][0], "final1": plan.final[1][0]}
    stats = mc.StreamingStats(bufs, targets={"moved0": y[0]})
    stats.count_dev = cnt
    stats.reset()
    for _ in mc.shard_samples(n_samples, rank, world):
        sampler.draw()
        plan.run_forward(x, sampler.z)
        stats.update()
        stats.count += 1
    return stats.reduce_to_maps([len(mc.shard_samples(n_samples, r, world)) for r in range(world)], dst=0)


def _worker_graph(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        res = _graph_job(dev, rank, world, 7)
        if rank == 0:
            torch.save({k: v.cpu() for k, v in res.items()}, os.path.join(out_dir, "g0.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_streaming_stats_flat_reduction_two_gpus_nccl(tmp_path):
    """7 Philox samples dealt to 2 GPUs (4 + 3) and reduced with all_to_all / reduce_scatter / gather == the same
    7 samples on one GPU (the noise of sample i does not depend on the rank that draws it)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker_graph, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "g0.pt"))
    one = _graph_job(torch.device("cuda", 0), 0, 1, 7)
    for k in one:
        torch.testing.assert_close(got[k], one[k].cpu(), rtol=1e-4, atol=1e-6)
