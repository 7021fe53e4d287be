"""Host-side placement for the end-to-end path: one process per GPU, bound to the CPU cores (and so, by
first touch, the host memory) of the NUMA node its GPU hangs off.

The end-to-end step copies 90 MB of pinned host memory to the GPU (HotPathPipeline); with several ranks on
one node the copies of ranks whose pinned buffers sit on the other socket cross the inter-socket link and
share it.  ``bind_to_gpu_numa`` must run before the pinned buffers are allocated.  It never widens the
process's CPU set (a container cpuset is respected) and is a no-op when NVML or the affinity information
is not available.
"""
from __future__ import annotations

import os


def _physical_index(device_index: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis and all(v.strip().isdigit() for v in vis.split(",")):
        return int(vis.split(",")[device_index])
    return device_index


def gpu_cpu_affinity(device_index: int):
    """CPU ids NVML reports as local to the GPU, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(_physical_index(device_index))
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa(device_index: int, min_cpus: int = 2) -> dict:
    """Restrict this process to the allowed CPUs local to ``device_index``'s GPU.  Returns what was done."""
    info = {"bound": False, "cpus": None}
    if not hasattr(os, "sched_setaffinity"):
        return info
    local = gpu_cpu_affinity(device_index)
    if not local:
        return info
    allowed = os.sched_getaffinity(0)
    target = allowed & local
    if len(target) < min_cpus or target == allowed:
        info["cpus"] = len(allowed)
        return info
    try:
        os.sched_setaffinity(0, target)
    except OSError:
        return info
    info.update(bound=True, cpus=len(target))
    return info
