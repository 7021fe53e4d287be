"""Run every hot-path kernel once at its level-0 (OASIS) shape between cudaProfilerStart/Stop,
for `ncu --profile-from-start off`.  Also prints CUDA-event times (no profiler) with --time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulpo_b200 import functional as PF, synthetic as syn  # noqa: E402


def main():
    timing = "--time" in sys.argv
    full, half, quarter = (160, 192, 224), (80, 96, 112), (40, 48, 56)
    x, y = (t.cuda() for t in syn.make_pair(full, 0))
    df_full = syn.make_field(full, 1, max_abs=3.0).cuda()
    v_half = syn.make_field(half, 2, max_abs=3.0).cuda()
    lo = syn.make_field(quarter, 3, max_abs=3.0).cuda()
    mu = syn.make_field(half, 4, max_abs=1.5).cuda()
    sg = (0.2 + 0.8 * torch.rand(1, 3, *half)).cuda()

    def ops():
        out = {}
        d = df_full.clone().requires_grad_(True)
        moved = PF.warp(d, x)
        out["warp"] = (moved, d)
        v = v_half.clone().requires_grad_(True)
        integ = PF.vecint(v, 7)
        out["vecint"] = (integ, v)
        xi = v_half.clone().requires_grad_(True)
        up = PF.resize_up(xi, 2, 2.0)
        out["up2"] = (up, xi)
        l = lo.clone().requires_grad_(True)
        comb = PF.resize_up(l, 2, 2.0, addend=v_half)
        out["combine"] = (comb, l)
        p = x.clone().requires_grad_(True)
        out["ncc"] = (PF.ncc_loss(p, y, 9, 0.05), p)
        f = df_full.clone().requires_grad_(True)
        out["l2"] = (PF.l2_reg(f, 0.025), f)
        m, s = mu.clone().requires_grad_(True), sg.clone().requires_grad_(True)
        out["kl"] = (PF.kl_diag(m, s, None, None), m)
        PF.interp_to_size(y, half)
        PF.avgpool2(x)
        for k, (o, leaf) in out.items():
            o.backward(torch.ones_like(o))
        return out

    ops()
    torch.cuda.synchronize()
    if timing:
        from pulpo_b200 import _lib
        _lib.profiler.enabled = _lib.profiler.timing = True
        for _ in range(3):
            _lib.profiler.reset()
            ops()
            torch.cuda.synchronize()
        for name, args, s, e in _lib.profiler.records:
            print("%-24s %8.1f us" % (name, s.elapsed_time(e) * 1e3))
        return
    torch.cuda.profiler.start()
    ops()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
