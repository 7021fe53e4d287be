"""Integration forward / backward time against the step count (fixed costs = prologue + epilogue + launch):
    python scripts/vi_steps.py [D0 D1 D2]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulpo_b200 import _lib, functional as PF, synthetic as syn  # noqa: E402

shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (80, 96, 112)
f = syn.make_field(shape, 1, max_abs=3.0).cuda()
flush = torch.empty(160 * 1024 * 1024 // 4, device="cuda")

for n in (0, 1, 2, 3, 5, 7):
    def run():
        v = f.clone().requires_grad_(True)
        o = PF.vecint(v, n, 0)
        flush.zero_()
        o.backward(torch.ones_like(o))
    run(); run()
    torch.cuda.synchronize()
    _lib.profiler.enabled = _lib.profiler.timing = True
    _lib.profiler.reset()
    for _ in range(8):
        run()
    torch.cuda.synchronize()
    _lib.profiler.enabled = _lib.profiler.timing = False
    agg = {}
    for name, args, s, e in _lib.profiler.records:
        agg.setdefault(name, []).append(s.elapsed_time(e) * 1e3)
    print("nsteps %d  " % n + "  ".join("%s min %.1f med %.1f us" % (k.replace("pulpo_", ""), min(t), sorted(t)[len(t) // 2]) for k, t in agg.items()))
