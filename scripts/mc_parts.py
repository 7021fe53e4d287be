"""Where does an MC sample's time go?  CUDA-event and host timers around the parts of bench.py's mc128 job."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulpo_b200 import mc, synthetic as syn
from pulpo_b200.plan import HotPathPlan

dev = torch.device("cuda", 0)
size, total, latent = [160, 192, 224], 5, 4
x_h, y_h, d_h, m_h, s_h = syn.make_hot_path_inputs(size, total, latent, seed=0)
x, y = x_h.to(dev), y_h.to(dev)
mu = {l: d_h[l].to(dev) for l in range(latent)}
sg = {l: (0.3 * s_h[l]).to(dev) for l in range(latent)}
plan = HotPathPlan(size, total, latent, batch=1, device=dev, with_reg=False)
z = {l: torch.empty_like(mu[l]) for l in range(latent)}
plan.run_forward(x, z)
bufs = {}
for l in range(latent):
    bufs["moved%d" % l], bufs["final%d" % l], bufs["indiv%d" % l] = plan.moved[l][0], plan.final[l][0], z[l][0]
stats = mc.StreamingStats(bufs, targets={"moved0": y[0]})
gen = torch.Generator(device=dev)


def graph_of(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


g_fwd = graph_of(lambda: plan.run_forward(x, z))
g_stats = graph_of(stats.update)
g_all = graph_of(lambda: (plan.run_forward(x, z), stats.update()))


def noise(i):
    gen.manual_seed(i)
    for l in range(latent):
        torch.normal(mu[l], sg[l], generator=gen, out=z[l])


def timeit(name, fn, n=16):
    fn(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(n):
        fn(i)
    e1.record(); t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("%-28s gpu %.1f us/sample   host-enqueue %.1f us/sample" % (name, e0.elapsed_time(e1) * 1e3 / n, t_host * 1e6 / n))


timeit("noise (4 x torch.normal)", noise)
timeit("forward graph", lambda i: g_fwd.replay())
timeit("statistics graph", lambda i: g_stats.replay())
timeit("forward+statistics graph", lambda i: g_all.replay())
timeit("whole sample", lambda i: (noise(i), g_all.replay()))
