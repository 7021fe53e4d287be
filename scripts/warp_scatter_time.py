"""Time the warp backward with the image-gradient scatter (segmentation path, C channels)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulpo_b200 import _lib, functional as PF, synthetic as syn
shape, C = (160, 192, 224), int(sys.argv[1]) if len(sys.argv) > 1 else 4
f = syn.make_field(shape, 1, max_abs=3.0).cuda()
img = torch.rand(1, C, *shape, device="cuda")
for _ in range(2):
    i = img.clone().requires_grad_(True)
    PF.warp(f, i).backward(torch.ones(1, C, *shape, device="cuda"))
torch.cuda.synchronize()
_lib.profiler.enabled = _lib.profiler.timing = True
_lib.profiler.reset()
for _ in range(4):
    i = img.clone().requires_grad_(True)
    PF.warp(f, i).backward(torch.ones(1, C, *shape, device="cuda"))
torch.cuda.synchronize()
agg = {}
for name, args, s, e in _lib.profiler.records:
    agg.setdefault(name, []).append(s.elapsed_time(e) * 1e3)
for k, t in agg.items():
    print("%s C=%d: min %.1f us" % (k, C, min(t)))
