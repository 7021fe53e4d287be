"""Stand-alone CUDA-event times of the level-0 links of the hot-path chain (config 2 shapes by default), each
launched alone after an L2 flush, through the C ABI.
    python scripts/chain_time.py [--reps 7] [D0 D1 D2 (half-res)]
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulpo_b200 import _lib, synthetic as syn  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 7
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
    if "--reps" in sys.argv:
        args.remove(str(reps))
    if only:
        args.remove(",".join(only))
    half = tuple(int(v) for v in args[:3]) if len(args) >= 3 else (80, 96, 112)
    full = tuple(2 * v for v in half)
    L = _lib.lib()
    vp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    x, y = (t.cuda() for t in syn.make_pair(full, 0))
    integ = syn.make_field(half, 1, max_abs=3.0).cuda()
    final = torch.empty(1, 3, *full, device="cuda")
    moved, gmoved = torch.empty_like(x), torch.empty_like(x)
    dpos, gfinal = torch.empty_like(final), torch.empty_like(final)
    ginteg = torch.empty_like(integ)
    abc = torch.empty(3, 1, 1, *full, device="cuda")
    reg, ncc = torch.zeros((), device="cuda"), torch.zeros((), device="cuda")
    ws = torch.zeros(L.pulpo_reduce_ws_bytes(), dtype=torch.uint8, device="cuda")
    wsn = torch.zeros(L.pulpo_ncc_ws_bytes(1, 1, *full), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    rscr = torch.empty(L.pulpo_l2reg_up2_scratch_bytes(1, 3, *half) // 4, device="cuda")
    ops = [
        ("l2reg_up2 (coarse)", lambda: L.pulpo_l2reg_up2_fwd_bwd(vp(integ), 0.025, vp(reg), vp(ginteg), 0, vp(rscr), rscr.numel() * 4, vp(ws), ws.numel(), 1, 3, *half, st)),
        ("resize_up_fwd x2", lambda: L.pulpo_resize_up_fwd(vp(integ), None, vp(final), 2, 2.0, 1, 3, *half, st)),
        ("warp3d_l2reg_fwd", lambda: L.pulpo_warp3d_l2reg_fwd(vp(x), vp(final), vp(moved), 0.025, vp(reg), vp(ws), ws.numel(), 1, 1, *full, 0, st)),
        ("warp3d_fwd", lambda: L.pulpo_warp3d_fwd(vp(x), vp(final), vp(moved), None, 1, 1, *full, 0, st)),
        ("warp3d_fwd_dpos", lambda: L.pulpo_warp3d_fwd_dpos(vp(x), vp(final), vp(moved), vp(dpos), 1, *full, 0, st)),
        ("ncc_fwd", lambda: L.pulpo_ncc_fwd(vp(moved), vp(y), vp(ncc), vp(abc), vp(wsn), wsn.numel(), 9, 0.05, 1, 1, *full, st)),
        ("ncc_bwd", lambda: L.pulpo_ncc_bwd(vp(abc), vp(moved), vp(y), None, vp(gmoved), 9, 0.05, 1, 1, *full, st)),
        ("warp3d_l2reg_bwd", lambda: L.pulpo_warp3d_l2reg_bwd(vp(gmoved), vp(x), vp(final), vp(gfinal), 0.025, None, 1, 1, *full, 0, st)),
        ("l2reg_fwd_bwd+prod", lambda: L.pulpo_l2reg_fwd_bwd(vp(final), 0.025, vp(reg), vp(gmoved), vp(dpos), vp(gfinal), 0, vp(ws), ws.numel(), 1, 3, *full, st)),
        ("l2reg_fwd_bwd", lambda: L.pulpo_l2reg_fwd_bwd(vp(final), 0.025, vp(reg), None, None, vp(gfinal), 0, vp(ws), ws.numel(), 1, 3, *full, st)),
        ("warp3d_bwd_dpos", lambda: L.pulpo_warp3d_bwd_dpos(vp(gmoved), vp(dpos), vp(gfinal), 0, 1, *full, st)),
        ("resize_up_bwd x2", lambda: L.pulpo_resize_up_bwd(vp(gfinal), vp(ginteg), 2, 2.0, 0, 1, 3, *half, st)),
        ("resize_up2_bwd_dpos", lambda: L.pulpo_resize_up2_bwd_dpos(vp(gmoved), vp(dpos), vp(ginteg), 2.0, 1, 1, *half, st)),
    ]
    for name, fn in ops:
        if only and not any(o in name for o in only):
            _lib.check(fn(), name)     # keep the data flow (later ops read this one's outputs)
            continue
        ts = []
        for _ in range(reps + 1):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            _lib.check(fn(), name)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts = sorted(ts[1:])
        print("%-22s %s  min %7.1f us  median %7.1f us" % (name, "x".join(map(str, full)), ts[0], ts[len(ts) // 2]))


if __name__ == "__main__":
    main()
