"""Timeline of one CUDA-graph replay of the benchmarked step (config 2 by default): every kernel with its start
offset, duration and stream, from CUPTI through torch.profiler (no ncu serialisation: this is the step as it runs).
    python scripts/step_timeline.py [--replays 5] [--out gpurun_out/timeline.md] [plan kwargs as k=v ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulpo_b200 import synthetic as syn  # noqa: E402
from pulpo_b200.plan import HotPathPlan  # noqa: E402


def main():
    argv = sys.argv[1:]
    replays = int(argv[argv.index("--replays") + 1]) if "--replays" in argv else 5
    out = argv[argv.index("--out") + 1] if "--out" in argv else None
    kw = {}
    for a in argv:
        if "=" in a and not a.startswith("--"):
            k, v = a.split("=", 1)
            kw[k] = (v == "True") if v in ("True", "False") else (int(v) if v.lstrip("-").isdigit() else v)
    size, total, latent = [160, 192, 224], 5, 4
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(size, total, latent, seed=0)
    x, y = x.cuda(), y.cuda()
    dfs = {l: dfs[l].cuda() for l in dfs}
    mus = {l: mus[l].cuda() for l in mus}
    sgs = {l: sgs[l].cuda() for l in sgs}
    plan = HotPathPlan(size, total, latent, batch=1, **kw)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            plan.run(x, y, dfs, mus, sgs)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        plan.run(x, y, dfs, mus, sgs)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(replays):
            g.replay()
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()
           and "memset" not in e.name.lower()]
    evs.sort(key=lambda e: e.time_range.start)
    # split into replays by the large idle gaps (host synchronise between replays)
    groups, cur = [], []
    for e in evs:
        if cur and e.time_range.start - max(c.time_range.end for c in cur) > 50:
            groups.append(cur)
            cur = []
        cur.append(e)
    if cur:
        groups.append(cur)
    grp = groups[-1]
    t0 = grp[0].time_range.start
    lines = ["# Timeline of one graph replay of the step (CUPTI via torch.profiler; %d kernels, %.1f us wall)" %
             (len(grp), max(e.time_range.end for e in grp) - t0), "",
             "| start us | dur us | end us | stream | kernel |", "|---:|---:|---:|---:|---|"]
    import re
    for e in grp:
        name = re.sub(r"\(.*$", "", e.name).replace("void ", "").replace("pulpo::", "")
        lines.append("| %.1f | %.1f | %.1f | %s | `%s` |" % (e.time_range.start - t0, e.time_range.end - e.time_range.start,
                                                           e.time_range.end - t0, getattr(e, "device_resource_id", "?"), name[:70]))
    walls = [max(e.time_range.end for e in gg) - gg[0].time_range.start for gg in groups]
    lines += ["", "replay wall times (us): " + ", ".join("%.1f" % w for w in walls)]
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
