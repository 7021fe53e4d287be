"""Build libpulpo_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pulpo_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpulpo_b200.so")
SOURCES = ["warp3d.cu", "vecint.cu", "resize.cu", "ncc.cu", "ncc_tma.cu", "losses.cu", "jacdet.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--use_fast_math=false",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "pulpo_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defs=(), tag: str = "") -> str:
    """``defs``/``tag`` build a tuning variant (``-DNAME=VALUE`` ...) as lib/libpulpo_b200_<tag>.so,
    selected at run time with PULPO_B200_LIB (see _lib.py); the default build takes neither."""
    lib = LIB if not tag else os.path.join(LIBDIR, "libpulpo_b200_%s.so" % tag)
    if not force and not tag and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", (".%s.o" % tag) if tag else ".o"))
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + ["-D" + d for d in defs] + \
              (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write("== %s ==\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libpulpo_b200")
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
    if tag:
        for o in objs:
            os.remove(o)
    return lib


if __name__ == "__main__":
    _defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    _tag = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--tag=")), "")
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defs=_defs, tag=_tag))
