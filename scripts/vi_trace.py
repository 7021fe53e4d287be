"""Tuning: per-CTA work time per backward step of the multi-level VecInt (needs a -DPULPO_VI_TRACE build).
    python -m pulpo_b200.build --tag=trace -DPULPO_VI_TRACE
    PULPO_B200_LIB=.../libpulpo_b200_trace.so python scripts/vi_trace.py
"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulpo_b200 import _lib, synthetic as syn
L = _lib.lib()
raw = ctypes.CDLL(_lib.LIB_PATH)
shapes, B, n = [(80, 96, 112), (40, 48, 56), (20, 24, 28), (10, 12, 14)], 1, 7
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
amps = [45.0, 21.0, 9.0, 3.0]   # magnitudes of the combined fields in the bench workload (3 voxels per level, summed coarse to fine)
v = [syn.make_field(sh, 3 + i, max_abs=amps[i]).cuda() for i, sh in enumerate(shapes)]
g = [syn.make_field(sh, 9 + i, max_abs=1.0).cuda() for i, sh in enumerate(shapes)]
ws = [torch.empty(L.pulpo_vecint_ws_bytes(n, 1, B, *sh) // 4, device="cuda") for sh in shapes]
scr = [torch.empty(L.pulpo_vecint_bwd_scratch_bytes(B, *sh) // 4, device="cuda") for sh in shapes]
out = [torch.empty_like(t) for t in v]
gv = [torch.empty_like(t) for t in v]
arr = (_lib.VecIntLevel * len(shapes))()
flush = torch.empty(160 * 1024 * 1024 // 4, device="cuda")
for rep in range(3):
    for i, sh in enumerate(shapes):
        arr[i] = _lib.VecIntLevel(v[i].data_ptr(), out[i].data_ptr(), ws[i].data_ptr(), ws[i].numel() * 4,
                                  scr[i].data_ptr(), scr[i].numel() * 4, *sh)
    _lib.check(L.pulpo_vecint_multi_fwd(arr, len(shapes), n, 1, B, 0, st))
    flush.zero_()
    for i in range(len(shapes)):
        arr[i].inp, arr[i].out = g[i].data_ptr(), gv[i].data_ptr()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.pulpo_vecint_multi_bwd(arr, len(shapes), n, B, 0, st))
    e1.record()
    torch.cuda.synchronize()
print("bwd total %.1f us" % (e0.elapsed_time(e1) * 1e3))
buf = (ctypes.c_ulonglong * (148 * 64))()
assert raw.pulpo_debug_vi_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(148, 32, 2)[:, :n, :]
t0 = t[:, :, 0].min()
for k in range(n - 1, -1, -1):
    start, end = t[:, k, 0] - t0, t[:, k, 1] - t0
    work = end - start
    print("step %d: start spread %5d ns | work min %6d median %6d p90 %6d max %6d ns | step wall %6d ns" % (
        k, start.max() - start.min(), work.min(), np.median(work), np.percentile(work, 90), work.max(), end.max() - start.min()))
k = n - 1
work = (t[:, k, 1] - t[:, k, 0])
print("slowest CTAs at step %d:" % k, np.argsort(work)[-8:], np.sort(work)[-8:])
print("fastest CTAs:", np.argsort(work)[:8], np.sort(work)[:8])
