"""Run selected hot-path ops a few times at given shapes; print CUDA-event times.  Used alone for
timing and under `ncu -k regex:...` for a full capture of one kernel.
    python scripts/prof_one.py vecint 80 96 112 [--reps 5]
    python scripts/prof_one.py warp|ncc|up2|l2 160 192 224
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulpo_b200 import _lib, functional as PF, synthetic as syn  # noqa: E402


def main():
    op = sys.argv[1]
    shape = tuple(int(v) for v in sys.argv[2:5])
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 5
    win = int(sys.argv[sys.argv.index("--win") + 1]) if "--win" in sys.argv else 9
    nsteps = int(sys.argv[sys.argv.index("--nsteps") + 1]) if "--nsteps" in sys.argv else 7
    mode = int(sys.argv[sys.argv.index("--mode") + 1], 0) if "--mode" in sys.argv else 0
    x, y = (t.cuda() for t in syn.make_pair(shape, 0))
    f = syn.make_field(shape, 1, max_abs=3.0).cuda()
    flush = torch.empty(160 * 1024 * 1024 // 4, device="cuda")

    def run():
        if op == "vecint":
            v = f.clone().requires_grad_(True)
            o = PF.vecint(v, nsteps, mode)
        elif op == "warp":
            v = f.clone().requires_grad_(True)
            o = PF.warp(v, x, mode & 1)
        elif op == "ncc":
            v = x.clone().requires_grad_(True)
            o = PF.ncc_loss(v, y, win, 0.05)
        elif op == "up2":
            v = f.clone().requires_grad_(True)
            o = PF.resize_up(v, 2, 2.0)
        elif op == "l2":
            v = f.clone().requires_grad_(True)
            o = PF.l2_reg(v, 0.025)
        flush.zero_()
        o.backward(torch.ones_like(o))

    run()
    torch.cuda.synchronize()
    _lib.profiler.enabled = _lib.profiler.timing = True
    _lib.profiler.reset()
    for _ in range(reps):
        run()
    torch.cuda.synchronize()
    agg = {}
    for name, args, s, e in _lib.profiler.records:
        agg.setdefault(name, []).append(s.elapsed_time(e) * 1e3)
    for name, ts in agg.items():
        ts.sort()
        print("%-26s %s  min %8.1f us  median %8.1f us" % (name, "x".join(map(str, shape)), ts[0], ts[len(ts) // 2]))


if __name__ == "__main__":
    main()
