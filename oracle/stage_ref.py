"""Stage the reference's Python package for the GPU box (TEST INFRASTRUCTURE ONLY).

    python -m oracle.stage_ref        # /root/reference/src/**/*.py  ->  oracle/_ref/src/

The whole-model drop-in test (tests/test_gpu_dropin.py) runs the UNMODIFIED reference model
(`src.models.PULPo.training_step`, /root/reference/src/models.py:134-196) next to the same model with
pulpo_b200's modules patched in, on the GPU.  /root/reference does not exist on the GPU box, so the reference's
own `src` package is copied -- unmodified, .py files only -- into oracle/_ref/, which is git-ignored (never enters
the history) but not gpurun-ignored (it travels with the snapshot, like the built .so files).  Run by
`__graft_entry__.build()` whenever /root/reference is present.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def stage(ref_root: str = "/root/reference") -> str | None:
    src = os.path.join(ref_root, "src")
    if not os.path.isdir(src):
        return None
    dst = os.path.join(DST, "src")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    for dirpath, dirnames, filenames in os.walk(src):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, src)
        out = os.path.join(dst, rel) if rel != "." else dst
        os.makedirs(out, exist_ok=True)
        for f in filenames:
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(dirpath, f), os.path.join(out, f))
    return dst


if __name__ == "__main__":
    print(stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
