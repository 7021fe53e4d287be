#!/usr/bin/env python
"""bench.py -- PULPo registration hot path on B200 (BASELINE.json metric:
"warp+NCC fwd/bwd Gvoxels/s at 1/2/4/8 B200; % of HBM roofline").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # reference PyTorch CPU path (oracle port)

A "step" = one pass of the hot path (per level: combine -> 7-step integration -> output resize
-> warp; hierarchical NCC + beta*KL + L2 losses; backward to the velocity fields / mu / sigma)
over one synthetic OASIS-shaped (160x192x224) pair per GPU -- configs[1] of BASELINE.json.
Multi-GPU: one process per GPU (torchrun), pairs sharded across ranks with no data-path
collective (weak scaling); the three loss scalars are all-reduced over NCCL each step.

value   = voxels of all ranks / max-over-ranks device time, inputs resident in HBM, the whole
          step replayed as one CUDA graph.
e2e     = same metric through the public module API with pinned HOST inputs: H2D copies of
          x, y, velocity fields, mu, sigma and the D2H read of the loss are inside the timed region.
roofline= dominant kernel: algorithmic bytes / its CUDA-event time, vs MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (input_size, total_levels, latent_levels)
    "oasis_160x192x224_5tot_4lat": ([160, 192, 224], 5, 4),      # BASELINE.json configs[1] (and [4] with --batch / --grad-allreduce)
    "cube64_4tot_3lat": ([64, 64, 64], 4, 3),                    # configs[0]
    "oasis_4tot_4lat": ([160, 192, 224], 4, 4),                  # configs[3]: total = latent levels -> level 0 integrates at full resolution
    "vecint_fullres": ([160, 192, 224], 1, 1),                   # configs[3]: stand-alone VecInt((160,192,224), 7) fwd + bwd
    "mc128": ([160, 192, 224], 5, 4),                            # configs[2]: 128 MC deformation samples per pair, sharded over the ranks
}
CPU_WORKLOADS = ("oasis_160x192x224_5tot_4lat", "cube64_4tot_3lat")
ALGO_BYTES_PER_VOXEL = 220.9   # SURVEY.md 8d, config 2, fwd+bwd incl. L2_reg
SURVEY_BYTES_PER_VOXEL = {"oasis_160x192x224_5tot_4lat": ALGO_BYTES_PER_VOXEL, "cube64_4tot_3lat": 220.7, "oasis_4tot_4lat": 739.0}
METRIC = "hot-path fwd+bwd throughput (warp + integration + pyramid + NCC/KL/L2)"
UNIT = "Gvoxel/s"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The region is ~20 ms long (20 steps of
    ~1 ms), far below nvidia-smi's 100 ms polling grain, so NVML is polled directly from a thread (~1 ms
    period; the main thread sits in a CUDA synchronize and has released the GIL).  Falls back to
    `nvidia-smi -lms` when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []
        self.nvml, self.handle, self.samples, self.stop_flag, self.thread = None, None, [], False, None
        self.inert = gpu_index is None       # ranks other than 0: no polling thread fighting the launch thread for the GIL
        if self.inert:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append((time.perf_counter(), n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM),
                                     n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)))
            except Exception:
                self.errors = getattr(self, "errors", 0) + 1      # transient NVML error: keep polling
            time.sleep(0.001)

    window = "timed region"

    def count_inside(self):
        if self.inert:
            return 99
        t1 = self.t_end if self.t_end is not None else time.perf_counter()
        return sum(1 for t, _, _ in self.samples if self.t_begin <= t <= t1) if self.nvml is not None else 99

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def start(self):
        """Start polling (call it before the warm-up so the thread is up and running when the timed region
        starts); ``mark_begin`` / ``mark_end`` bracket the region on the host clock."""
        self.t_begin, self.t_end = time.perf_counter(), None
        if self.inert:
            return
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.inert:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["not sampled on this rank"]}
        if self.nvml is not None:
            n = self.nvml
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            t1 = self.t_end if self.t_end is not None else time.perf_counter()
            inside = [(c, r) for t, c, r in self.samples if self.t_begin <= t <= t1]
            sm = sorted(c for c, _ in inside)
            mask = 0
            for _, r in inside:
                mask |= int(r)
            names = (("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap))
            try:
                smax = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            except Exception:
                smax = None
            return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": smax,
                    "reasons": [k for k, bit in names if mask & bit], "samples": len(sm), "source": "nvml, 1 ms polling",
                    "window": self.window}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            p = [t.strip() for t in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = [v for v in sm if v > 0.5 * (max(sm) if sm else 1)]
        med = busy[len(busy) // 2] if busy else (sm[len(sm) // 2] if sm else None)
        return {"sm_mhz": med, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------ reference arm
def _cpu_hot_path_step(T, torch, inputs, total_levels):
    x, y, dfs, mus, sgs = inputs
    d = {l: dfs[l].clone().requires_grad_(True) for l in dfs}
    m = {l: mus[l].clone().requires_grad_(True) for l in dfs}
    s = {l: sgs[l].clone().requires_grad_(True) for l in dfs}
    t0 = time.perf_counter()
    loss, _, _ = T.hot_path_losses(x, y, d, m, s, total_levels)
    loss.backward()
    return time.perf_counter() - t0, float(loss.detach())


def cpu_baseline(budget_s, steps=1, warmup=0, want_workload=None, allow_smaller=True):
    """Reference PyTorch CPU path (oracle/torch_ref.py: the same ATen ops the reference calls)
    on the host cores, on a bounded sample of the workload, calibrated on a 32^3 pass.
    ``allow_smaller=False`` (the --impl reference arm): always the requested workload itself; when
    (steps + warmup) passes do not fit the budget, fewer passes are run (never a smaller volume) and
    the number actually run is reported.  ``allow_smaller=True`` (the in-line cpu_baseline of our own
    arm, ~25 s): the largest shape whose passes fit the budget, named in ``sample``."""
    import torch
    from oracle import torch_ref as T
    from pulpo_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cal = syn.make_hot_path_inputs([32, 32, 32], 4, 3, seed=0)
    _cpu_hot_path_step(T, torch, cal, 4)
    t_cal, _ = _cpu_hot_path_step(T, torch, cal, 4)
    per_voxel = t_cal / 32 ** 3
    candidates = [("oasis_160x192x224_5tot_4lat", [160, 192, 224], 5, 4), ("half_80x96x112_4tot_3lat", [80, 96, 112], 4, 3),
                  ("cube64_4tot_3lat", [64, 64, 64], 4, 3), ("cube32_4tot_3lat", [32, 32, 32], 4, 3)]
    if want_workload:
        candidates = [c for c in candidates if c[0] == want_workload] + candidates
    if not allow_smaller:
        chosen = candidates[0]
        est = per_voxel * chosen[1][0] * chosen[1][1] * chosen[1][2] * 1.3
        fit = max(1, int(budget_s / max(est, 1e-9)))
        if steps + warmup > fit:          # fewer passes of the SAME workload, never another one
            warmup = min(warmup, 1 if fit > 1 else 0)
            steps = max(1, fit - warmup)
    else:
        chosen = candidates[-1]
        for c in candidates:
            n = c[1][0] * c[1][1] * c[1][2]
            if per_voxel * n * 1.3 * (steps + warmup) <= budget_s:
                chosen = c
                break
    name, size, total, latent = chosen
    inputs = syn.make_hot_path_inputs(size, total, latent, seed=0)
    for _ in range(warmup):
        _cpu_hot_path_step(T, torch, inputs, total)
    times = [_cpu_hot_path_step(T, torch, inputs, total)[0] for _ in range(steps)]
    nvox = size[0] * size[1] * size[2]
    tot = sum(times)
    return {"value": nvox * steps / tot / 1e9, "unit": UNIT, "cores": cores, "kind": "port", "workload": name,
            "steps_run": steps, "warmup_run": warmup,
            "sample": "%d step(s) (+%d warm-up) of %s (B=1) fwd+bwd via oracle/torch_ref.py (torch %s CPU, %d threads), %.2f s/step"
                      % (steps, warmup, name, torch.__version__, cores, tot / steps)}, tot / steps, name


def torch_cuda_baseline(dev, inputs, total_levels, steps=3):
    """Context row: the same reference restatement (the ATen ops the reference calls) on THIS GPU through
    stock PyTorch eager CUDA kernels, TF32 off -- "what you get by moving the reference to a B200"."""
    import torch
    from oracle import torch_ref as T
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        x, y, dfs, mus, sgs = inputs
        x, y = x.to(dev), y.to(dev)

        def one():
            d = {l: dfs[l].to(dev).requires_grad_(True) for l in dfs}
            m = {l: mus[l].to(dev).requires_grad_(True) for l in dfs}
            s = {l: sgs[l].to(dev).requires_grad_(True) for l in dfs}
            loss, _, _ = T.hot_path_losses(x, y, d, m, s, total_levels)
            loss.backward()
        one()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle/torch_ref.py, the same ATen ops
    the reference calls) on all host cores, ALWAYS on the workload named in config.workload (the same one our arm
    runs); if steps + warmup passes would exceed the time budget, fewer passes of that same workload are run
    (cpu_baseline.steps_run says how many)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload not in CPU_WORKLOADS:
        _emit({"impl": "reference", "unavailable": "workload %s has no CPU reference arm" % args.workload})
        return
    cb, ms, name = cpu_baseline(budget_s=args.ref_budget_s, steps=args.steps, warmup=args.warmup,
                                want_workload=args.workload, allow_smaller=False)
    assert name == args.workload
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "same_config": True, "batch_per_gpu": 1, "steps_run": cb["steps_run"],
                       "note": "reference PyTorch CPU path on host cores; rank 0 only"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def kernel_breakdown(prof_step, prof_steps, peak, peak_src):
    """Run ``prof_step`` a few times with CUDA events around every C-ABI call; returns the per-entry-point table, the
    roofline object of the launch shape with the largest share of the step, and the algorithmic bytes of one step
    (sum over its calls, pulpo_b200/roofline.py = SURVEY.md 8d)."""
    import torch
    from pulpo_b200 import _lib
    from pulpo_b200.roofline import algo_bytes
    _lib.profiler.enabled, _lib.profiler.timing = True, True
    _lib.profiler.reset()
    for _ in range(prof_steps):
        prof_step()
    torch.cuda.synchronize()
    _lib.profiler.enabled = _lib.profiler.timing = False
    agg, shapes = {}, {}
    for name, cargs, s_ev, e_ev in _lib.profiler.records:
        t_ms, nb = s_ev.elapsed_time(e_ev), algo_bytes(name, cargs)
        a = agg.setdefault(name, {"ms": 0.0, "bytes": 0, "calls": 0})
        a["ms"] += t_ms; a["bytes"] += nb; a["calls"] += 1
        g = shapes.setdefault((name, nb), {"ms": 0.0, "n": 0})   # one entry per distinct launch shape
        g["ms"] += t_ms; g["n"] += 1
    breakdown = {k: {"ms_per_step": v["ms"] / prof_steps, "calls_per_step": v["calls"] / prof_steps,
                     "algo_MB_per_step": v["bytes"] / prof_steps / 1e6,
                     "GBps": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None}
                 for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    (dname, dbytes), dg = max(shapes.items(), key=lambda kv: kv[1]["ms"])
    d_ms = dg["ms"] / dg["n"]
    achieved = dbytes / (d_ms * 1e-3) / 1e9
    # DRAM traffic of that kernel per launch from the committed `ncu --set full` capture of this same command
    # (profiles/<tag>_traffic.json, written by scripts/summarize_ncu_full.py); None when no capture is committed
    traffic, traffic_src = None, None
    try:
        import glob
        for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
            tj = json.load(open(path))
            for kname, rec in tj.items():
                if rec.get("entry_point") == dname or KERNEL_OF_ENTRY.get(dname) == kname:
                    traffic, traffic_src = rec["dram_bytes_per_launch"], os.path.relpath(path, ROOT)
    except Exception:
        pass
    roofline = {"kernel": dname, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "bytes_per_launch": dbytes, "us_per_launch": d_ms * 1e3, "launches_timed": dg["n"],
                "share_of_step": dg["ms"] / sum(v["ms"] for v in agg.values()),
                "note": "launch shape with the largest share of the step; algorithmic bytes (SURVEY.md 8d) / mean CUDA-event duration"}
    return breakdown, roofline, sum(v["bytes"] for v in agg.values()) / prof_steps


KERNEL_OF_ENTRY = {"pulpo_vecint_multi_bwd": "vecint_bwd_kernel<0, 2>", "pulpo_vecint_multi_fwd": "vecint_fwd_kernel<0>",
                   "pulpo_warp3d_l2reg_bwd": "warp3d_bwd_kernel<0, 0, 1>", "pulpo_warp3d_l2reg_fwd": "warp3d_fwd_kernel<0, 0, 1>",
                   "pulpo_ncc_fwd": "ncc_tma_kernel<9, 1>", "pulpo_ncc_bwd": "ncc_tma_kernel<9, 0>"}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pulpo_b200 import _lib, synthetic as syn
    from pulpo_b200.models import RegistrationHotPath
    from pulpo_b200.roofline import algo_bytes, measured_peaks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- pulpo_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = {"bound": False}
    if world > 1:
        # one process per GPU: keep this rank's threads and pinned host buffers on its GPU's NUMA node
        from pulpo_b200.hostmem import bind_to_gpu_numa
        numa = bind_to_gpu_numa(local)
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()   # fail loudly if the CUDA library is missing

    size, total, latent = WORKLOADS[args.workload]
    nvox = size[0] * size[1] * size[2]
    B = args.batch
    plan_kw = dict(fuse_reg=bool(args.fuse_reg), fuse_combine=bool(args.fuse_combine), pool_pyramid=bool(args.pool_pyramid),
                   aux_early=bool(args.aux_early), dpos=bool(args.dpos), aux_after=args.aux_after, reg_coarse=bool(args.reg_coarse))
    x_h, y_h, d_h, m_h, s_h = syn.make_hot_path_inputs(size, total, latent, seed=rank, batch=B,
                                                       field_sigma_vox=args.field_sigma_vox, max_abs=args.field_max_abs)
    host = [x_h, y_h] + [d_h[l] for l in range(latent)] + [m_h[l] for l in range(latent)] + [s_h[l] for l in range(latent)]
    host = [t.pin_memory() for t in host]
    h2d_bytes = sum(t.numel() * 4 for t in host)
    hp = RegistrationHotPath(size, total, latent).to(dev)

    def to_dev(non_blocking=True):
        ts = [t.to(dev, non_blocking=non_blocking) for t in host]
        x, y = ts[0], ts[1]
        d = {l: ts[2 + l].requires_grad_(True) for l in range(latent)}
        m = {l: ts[2 + latent + l].requires_grad_(True) for l in range(latent)}
        s = {l: ts[2 + 2 * latent + l].requires_grad_(True) for l in range(latent)}
        return x, y, d, m, s

    x, y, d, m, s = to_dev(False)
    leaves = list(d.values()) + list(m.values()) + list(s.values())
    scal = torch.zeros(3, device=dev)

    plan = None
    if args.engine == "plan":
        from pulpo_b200.plan import HotPathPlan
        plan = HotPathPlan(size, total, latent, batch=B, device=dev, **plan_kw)
        dd = {l: d[l].detach() for l in d}
        mm = {l: m[l].detach() for l in m}
        ss = {l: s[l].detach() for l in s}

    def exchange(parts3):
        # the only exchange the path itself has: the three loss scalars (gradients of the convs live upstream)
        if world > 1:
            scal.copy_(parts3)
            dist.all_reduce(scal)

    # config 5 (BASELINE.json configs[4]): data-parallel training also all-reduces the conv encoder/decoder gradients
    # (14 667 012 fp32 = 58.7 MB at n0=32, 5/4 levels).  The convs are out of scope, so a buffer of that
    # size stands for them; its all-reduce is issued asynchronously at the start of a step and waited for at its
    # end, i.e. it overlaps the hot path of the same step on NCCL's own stream.
    gradbuf = torch.zeros(14667012, device=dev) if (args.grad_allreduce and world > 1) else None

    def compute():
        """One hot-path forward+backward (the part that is captured as a CUDA graph)."""
        if plan is not None:     # pre-planned multi-stream launch sequence (same kernels, no autograd)
            return plan.run(x, y, dd, mm, ss)
        for t in leaves:
            t.grad = None
        loss, parts, _ = hp(x, y, d, m, s)
        loss.backward()
        step.parts = torch.stack([parts["kl"], parts["recon"], parts["reg"]]).detach()
        return loss

    def step():
        loss = compute()
        if plan is None:
            exchange(step.parts)
        return loss

    # ---- warm-up (eager), then capture the step as ONE CUDA graph
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    launch_mode = "cuda_graph"
    graph = None
    if not args.no_graph and plan is not None:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    compute()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = compute()
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
        except Exception as e:  # pragma: no cover
            sys.stderr.write("bench.py: CUDA graph capture failed (%r); timing eager launches\n" % (e,))
            graph, launch_mode = None, "eager"
    else:
        launch_mode = "eager"

    def run_step():
        work = dist.all_reduce(gradbuf, async_op=True) if gradbuf is not None else None
        if graph is not None:
            graph.replay()     # the per-term loss sums accumulate on the device (plan.running): no per-step collective
        else:
            step()
        if work is not None:
            work.wait()        # stream-side wait: the next step's kernels queue behind the all-reduce

    def sync_losses():
        # ONE all-reduce of the logged loss scalars per timed region instead of one per step
        if world > 1 and plan is not None:
            scal.copy_(plan.running)
            dist.all_reduce(scal)

    # count our kernel launches per step (C-ABI calls; each enqueues exactly one kernel)
    _lib.profiler.enabled, _lib.profiler.timing = True, False
    _lib.profiler.reset()
    step()
    launches_per_step = _lib.profiler.launches
    _lib.profiler.enabled = False
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: EXACTLY K steps, CUDA events, max over ranks
    clocks = ClockSampler(local if rank == 0 else None)    # NVML is polled on rank 0 only
    clocks.start()
    for _ in range(args.warmup):
        run_step()
    sync_losses()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.mark_begin()
    e0.record()
    for _ in range(args.steps):
        run_step()
    sync_losses()
    e1.record()
    barrier()
    clocks.mark_end()
    # An NVML query occasionally blocks for longer than the whole (~20 ms) timed region.  If fewer than 3 samples
    # fell inside it, keep the GPU under the very same load (untimed replays of the same step) for a short
    # continuation and sample there too; the window is reported.
    if clocks.count_inside() < 3:
        t_stop = time.perf_counter() + 0.25
        while time.perf_counter() < t_stop:
            for _ in range(10):     # rank-local replays only: no collective (ranks may differ in whether they continue)
                graph.replay() if graph is not None else compute()
            torch.cuda.synchronize()
        clocks.mark_end()
        clocks.window = "timed region + 0.25 s continuation at the same load (untimed)"
    clk = clocks.stop()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B * nvox / (ms_step * 1e-3) / 1e9

    # ---- e2e: public API, pinned host inputs, H2D + D2H inside the timed region
    e2e_steps = max(4, min(args.steps, 20))
    if plan is not None:
        # HotPathPipeline: one packed pinned buffer per step -> one H2D copy on a copy stream, double-buffered
        # against the graph-replayed compute; the loss scalars are read back (D2H) for every step
        from pulpo_b200.pipeline import HotPathPipeline
        pipe = HotPathPipeline(size, total, latent, batch=B, device=dev, **plan_kw)
        hbs = [pipe.host_batch().fill(x_h, y_h, d_h, m_h, s_h) for _ in range(2)]
        h2d_bytes, d2h_bytes = pipe.h2d_bytes, 16
        e2e_api = "pulpo_b200.pipeline.HotPathPipeline.submit/result (packed pinned batch, copy/compute overlap)"

        def e2e_run(n):
            last, acc = None, [0.0, 0.0, 0.0]
            for k in range(n):
                t = pipe.submit(hbs[k % 2])
                if last is not None:
                    r = pipe.result(last)        # D2H read of the previous step's result
                    acc = [a + b for a, b in zip(acc, r[1:])]
                last = t
            r = pipe.result(last)
            if world > 1:                        # the logged scalars: one all-reduce per region, not per step
                exchange(torch.tensor([a + b for a, b in zip(acc, r[1:])], device=dev))
            return r
    else:
        d2h_bytes = 4
        e2e_api = "pulpo_b200.models.RegistrationHotPath forward + backward (autograd modules)"

        def e2e_run(n):
            out = None
            for _ in range(n):
                xx, yy, dd_, mm_, ss_ = to_dev(True)
                loss, _, _ = hp(xx, yy, dd_, mm_, ss_)
                loss.backward()
                out = float(loss.item())     # D2H read of the step's result
            return out
    e2e_run(3)
    barrier()
    e0.record()
    e2e_run(e2e_steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * B * nvox / (ms_e2e / e2e_steps * 1e-3) / 1e9

    # ---- per-kernel breakdown with CUDA events around every C-ABI call (eager, same stream)
    peak, peak_src = measured_peaks(ROOT)
    # (single-stream plan: the multi-stream plan overlaps levels, so per-call events would not add up)
    prof_step = step
    if plan is not None:
        from pulpo_b200.plan import HotPathPlan
        plan1 = HotPathPlan(size, total, latent, batch=B, device=dev, multi_stream=False, **plan_kw)
        prof_step = lambda: plan1.run(x, y, dd, mm, ss)
        prof_step()
        torch.cuda.synchronize()
    breakdown, roofline, step_algo_bytes = kernel_breakdown(prof_step, max(2, min(args.steps, 5)), peak, peak_src)
    # SURVEY.md 8d totals where it states them (they count the regulariser's own read of the field, which the fused
    # warp kernels do not pay again); otherwise the sum of the per-call algorithmic bytes of one step
    algo_bpv = SURVEY_BYTES_PER_VOXEL.get(args.workload, step_algo_bytes / float(B * nvox))

    selfcheck = None
    if world > 1 and not args.no_selfcheck:
        try:
            selfcheck = multi_gpu_selfcheck(dev, rank, world)
        except Exception as e:  # pragma: no cover
            selfcheck = {"ok": False, "error": repr(e)}

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb, _, _ = cpu_baseline(budget_s=25.0, steps=1, warmup=0, want_workload=args.workload)
            except Exception as e:  # pragma: no cover
                cb = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "pairs_per_gpu": B, "voxels_per_pair": nvox,
                       "levels": "%d total / %d latent, level_res, 7 integration steps" % (total, latent),
                       "launch": launch_mode, "engine": args.engine,
                       "parallelism": "pairs sharded over %d GPU(s), no data-path collective; loss scalars all-reduced once per timed region%s"
                                      % (world, "; 58.7 MB gradient stand-in all-reduced every step (overlapped)" if gradbuf is not None else ""),
                       "l2": "per-step working set ~1.5 GB >> 126 MB L2; no explicit flush",
                       "velocity_fields": "smooth N(0,1), max |v| = %g voxels per level, sigma = %s" % (args.field_max_abs,
                           "%g voxels of each level's grid" % args.field_sigma_vox if args.field_sigma_vox is not None
                           else "8 full-res voxels at every level")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, "api": e2e_api,
                    "host_numa_binding": numa},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clk,
            "roofline": roofline,
            "path_roofline": {"algo_bytes_per_voxel": algo_bpv,
                              "achieved": value / world * algo_bpv, "peak": peak, "unit": "GB/s",
                              "frac": value / world * algo_bpv / peak, "peak_source": peak_src},
            "kernels": breakdown,
        }
        if cb is not None:
            line["cpu_baseline"] = cb
        if selfcheck is not None:
            line["selfcheck"] = selfcheck
        if world == 1 and not args.no_cpu_baseline:
            try:
                ms_t = torch_cuda_baseline(dev, (x_h, y_h, d_h, m_h, s_h), total)
                line["torch_cuda_baseline"] = {"value": B * nvox / (ms_t * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_t,
                                               "what": "reference restatement (oracle/torch_ref.py) on stock PyTorch %s CUDA "
                                                       "kernels on this GPU, fp32, TF32 off; context only" % torch.__version__}
            except Exception as e:  # pragma: no cover
                line["torch_cuda_baseline"] = {"value": None, "what": "failed: %r" % (e,)}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------ multi-GPU self-check
def multi_gpu_selfcheck(dev, rank, world):
    """N > 1 only, outside every timed region: the MC-sample sharding of config 3 at 32^3 with the real kernels over
    NCCL -- Philox sampler + forward plan + streaming statistics per sample, samples dealt round-robin to the ranks
    (ragged: 2 W + 1 of them), flat all_to_all / reduce_scatter / gather reduction to rank 0 -- against the same samples
    run on rank 0 alone (the noise of sample i does not depend on the rank that draws it).  This is the content of
    tests/test_gpu_multirank.py, which a 1-GPU test box has to skip.  Every rank executes the same collectives in the
    same order; the comparison itself is rank-local."""
    import torch
    import torch.distributed as dist
    from pulpo_b200 import mc, synthetic as syn
    from pulpo_b200.plan import HotPathPlan

    size, total, latent, n_samples = [32, 32, 32], 4, 3, 2 * world + 1
    solo = [dist.new_group([r]) for r in range(world)][rank]      # collective: every rank creates every group
    x, y, dfs, mus, sgs = syn.make_hot_path_inputs(size, total, latent, seed=3)
    x, y = x.to(dev), y.to(dev)
    mu = {l: dfs[l].to(dev) for l in dfs}
    sg = {l: (0.3 * sgs[l]).to(dev) for l in dfs}

    def job(r, w, group):
        plan = HotPathPlan(size, total, latent, batch=1, device=dev, with_reg=False)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        sampler = mc.PhiloxSampler(mu, sg, seed=5, first_id=r, id_stride=w, count_dev=cnt)
        plan.run_forward(x, sampler.z)
        bufs = {"moved0": plan.moved[0][0], "final0": plan.final[0][0], "final1": plan.final[1][0]}
        stats = mc.StreamingStats(bufs, targets={"moved0": y[0]})
        stats.count_dev = cnt
        stats.reset()
        for _ in mc.shard_samples(n_samples, r, w):
            sampler.draw()
            plan.run_forward(x, sampler.z)
            stats.update()
            stats.count += 1
        return stats.reduce_to_maps([len(mc.shard_samples(n_samples, q, w)) for q in range(w)], group=group, dst=0)

    got = job(rank, world, None)
    torch.cuda.synchronize()
    out = {"what": "config-3 sharding at 32^3: %d Philox samples over %d ranks, NCCL all_to_all + reduce_scatter + gather "
                   "to rank 0, vs the same samples on rank 0 alone" % (n_samples, world), "ranks": world}
    if rank == 0:
        one = job(0, 1, solo)
        torch.cuda.synchronize()
        worst = 0.0
        for k in one:
            a, b = got[k].double(), one[k].double()
            worst = max(worst, float(((a - b).abs() / (1e-6 + 1e-4 * b.abs())).max()))
        out.update(ok=bool(worst <= 1.0), maps=sorted(one.keys()), worst_err_over_tol=worst, tol="rtol 1e-4, atol 1e-6")
    return out    # no collective after the rank-local comparison: a failure on rank 0 cannot leave the others waiting


# ------------------------------------------------------------------------------------------ shared rank plumbing
def _rank_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- pulpo_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from pulpo_b200 import _lib
    _lib.lib()   # fail loudly if the CUDA library is missing
    return torch, dist, world, rank, local, dev


def _timed(torch, dist, world, dev, local, rank, warmup, steps, fn, after=None):
    """W warm-up calls, then exactly K calls of ``fn`` between CUDA events, barrier + synchronize on both sides,
    max over ranks; clocks sampled during the region on rank 0."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    clocks = ClockSampler(local if rank == 0 else None)
    clocks.start()
    for _ in range(warmup):
        fn()
    if after:
        after()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.mark_begin()
    e0.record()
    for _ in range(steps):
        fn()
    if after:
        after()
    e1.record()
    barrier()
    clocks.mark_end()
    if clocks.count_inside() < 3:
        t_stop = time.perf_counter() + 0.25
        while time.perf_counter() < t_stop:
            fn()
            torch.cuda.synchronize()
        clocks.mark_end()
        clocks.window = "timed region + 0.25 s continuation at the same load (untimed)"
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, clk


# ------------------------------------------------------------------------------------------ config 4 (i): stand-alone VecInt
def run_vecint_fullres(args):
    """BASELINE.json configs[3]: 7-step scaling and squaring at full 160x192x224 resolution -- stand-alone
    VecInt((160,192,224), 7) forward + backward through the C ABI (pulpo_vecint_fwd / pulpo_vecint_bwd), one field per
    GPU.  Algorithmic bytes: 7 * 24 B (fwd) + 7 * 36 B (bwd) = 420 B per voxel."""
    import ctypes
    torch, dist, world, rank, local, dev = _rank_setup()
    from pulpo_b200 import _lib, synthetic as syn
    from pulpo_b200.roofline import measured_peaks
    lib = _lib.lib()
    size = WORKLOADS[args.workload][0]
    nvox, B, nsteps = size[0] * size[1] * size[2], args.batch, 7
    vp = ctypes.c_void_p
    vec_h = syn.make_field(size, 100 + rank, batch=B, max_abs=args.field_max_abs).pin_memory()
    gout_h = syn.make_field(size, 200 + rank, batch=B, max_abs=1.0).pin_memory()
    vec, gout = vec_h.to(dev), gout_h.to(dev)
    out, gvec = torch.empty_like(vec), torch.empty_like(vec)
    ws = torch.empty(lib.pulpo_vecint_ws_bytes(nsteps, 1, B, *size) // 4, device=dev)
    scr = torch.empty(lib.pulpo_vecint_bwd_scratch_bytes(B, *size) // 4, device=dev)

    def compute():
        st = vp(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.pulpo_vecint_fwd(vp(vec.data_ptr()), vp(out.data_ptr()), vp(ws.data_ptr()), ws.numel() * 4, nsteps, 1, B,
                                        *size, 0, st), "vecint_fwd")
        _lib.check(lib.pulpo_vecint_bwd(vp(gout.data_ptr()), vp(ws.data_ptr()), vp(gvec.data_ptr()), vp(scr.data_ptr()),
                                        scr.numel() * 4, nsteps, B, *size, 0, st), "vecint_bwd")
    for _ in range(3):
        compute()
    torch.cuda.synchronize()
    graph, launch_mode = None, "eager"
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            compute()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            compute()
        launch_mode = "cuda_graph"
    step = graph.replay if graph is not None else compute
    ms_total, clk = _timed(torch, dist, world, dev, local, rank, args.warmup, args.steps, step)
    ms_step = ms_total / args.steps
    value = world * B * nvox / (ms_step * 1e-3) / 1e9
    # e2e: pinned host fields in, a 4-byte checksum of the gradient out, every step
    chk_h = torch.zeros(1).pin_memory()

    def e2e_step():
        vec.copy_(vec_h, non_blocking=True)
        gout.copy_(gout_h, non_blocking=True)
        step()
        chk_h.copy_(gvec.view(-1)[:1], non_blocking=True)       # D2H read of (one element of) the step's result
        torch.cuda.current_stream().synchronize()
    e2e_steps = max(4, min(args.steps, 10))
    ms_e2e, _ = _timed(torch, dist, world, dev, local, rank, 2, e2e_steps, e2e_step)
    peak, peak_src = measured_peaks(ROOT)
    breakdown, roofline, step_bytes = kernel_breakdown(compute, max(2, min(args.steps, 5)), peak, peak_src)
    if rank == 0:
        bpv = step_bytes / float(B * nvox)
        _emit({"metric": "scaling-and-squaring fwd+bwd throughput (7 steps, full resolution)", "value": value, "unit": UNIT,
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": args.workload, "fields_per_gpu": B, "voxels_per_field": nvox, "nsteps": nsteps,
                          "launch": launch_mode, "l2": "7 saved states x 110 MB + 3 gradient states >> 126 MB L2; no explicit flush",
                          "velocity_field": "smooth N(0,1), max |v| = %g voxels" % args.field_max_abs},
               "e2e": {"value": world * B * nvox / (ms_e2e / e2e_steps * 1e-3) / 1e9, "unit": UNIT,
                       "h2d_bytes_per_step": 2 * B * 3 * nvox * 4, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / e2e_steps,
                       "steps": e2e_steps, "api": "pulpo_vecint_fwd / pulpo_vecint_bwd (C ABI) with pinned host vec / grad_output"},
               "gpu_launches": 2 * args.steps * (B if B > 1 else 1), "gpu_launches_per_step": 2 * (B if B > 1 else 1), "clocks": clk,
               "roofline": roofline,
               "path_roofline": {"algo_bytes_per_voxel": bpv, "achieved": value / world * bpv, "peak": peak, "unit": "GB/s",
                                 "frac": value / world * bpv / peak, "peak_source": peak_src},
               "kernels": breakdown})
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ config 3: MC uncertainty
def run_mc128(args):
    """BASELINE.json configs[2]: uncertainty inference -- N = 128 MC deformation samples of one OASIS-shaped pair,
    dealt round-robin to the ranks (16 per GPU on 8), per-voxel statistics streamed on every rank (Welford moments of
    moved / final / individual fields of every level + squared errors of the level-0 moved image), partial states
    reduced to rank 0 over a binomial tree (NCCL send/recv + Chan merge) and the variance map, MSE map and
    NCC(var, mse) of Evaluate.uncertainty (evaluate.py:1534-1545) formed there -- all inside the timed region.
    A "step" = the whole N-sample job for one pair.  Sample i draws its noise from Philox seed seed0 + i on whatever
    rank runs it, so any rank count sees the same 128 samples.  Reference loop: evaluate.py:205-251."""
    torch, dist, world, rank, local, dev = _rank_setup()
    from pulpo_b200 import _lib, mc, synthetic as syn
    from pulpo_b200.plan import HotPathPlan
    from pulpo_b200.roofline import measured_peaks
    size, total, latent = WORKLOADS[args.workload]
    nvox, N = size[0] * size[1] * size[2], args.samples
    x_h, y_h, d_h, m_h, s_h = syn.make_hot_path_inputs(size, total, latent, seed=0, max_abs=args.field_max_abs)   # same pair on every rank
    x, y = x_h.to(dev), y_h.to(dev)
    mu = {l: d_h[l].to(dev) for l in range(latent)}              # posterior mean of the velocity field (stands for the conv output)
    sg = {l: (0.3 * s_h[l]).to(dev) for l in range(latent)}
    plan = HotPathPlan(size, total, latent, batch=1, device=dev, with_reg=False)
    from pulpo_b200 import functional as PF
    ids = mc.shard_samples(N, rank, world)       # round-robin: sample ids rank, rank + world, ...
    tracked = ["moved%d" % l for l in range(latent)] + ["final%d" % l for l in range(latent)] + ["indiv%d" % l for l in range(latent)]
    count_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    sampler = mc.PhiloxSampler(mu, sg, seed=args.seed0, first_id=rank, id_stride=world, count_dev=count_dev)
    z = sampler.z                                # static sample buffers the graph reads
    for _ in range(2):
        plan.run_forward(x, z)
    torch.cuda.synchronize()
    bufs = {}
    for l in range(latent):
        bufs["moved%d" % l], bufs["final%d" % l], bufs["indiv%d" % l] = plan.moved[l][0], plan.final[l][0], z[l][0]
    stats = mc.StreamingStats(bufs, targets={"moved0": y[0]})
    stats.count_dev = count_dev                  # sampler and statistics share the device-side sample counter
    sampler.count_dev = count_dev

    def one_sample():            # a whole MC sample -- noise, forward pass, all statistics -- as one CUDA graph
        sampler.draw()
        plan.run_forward(x, z)
        stats.update()
    graph, launch_mode = None, "eager"
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            one_sample()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one_sample()
        launch_mode = "cuda_graph per sample (Philox noise + forward + statistics)"
    launches_fwd = plan.launches + 3
    result = {}

    trace = os.environ.get("PULPO_MC_TRACE") == "1"
    counts = [len(mc.shard_samples(N, r, world)) for r in range(world)]

    def job():
        if trace:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        stats.reset()
        for _ in ids:
            graph.replay() if graph is not None else one_sample()
            stats.count += 1
        if trace:
            ev[1].record()
        res = stats.reduce_to_maps(counts, dst=0)
        if trace:
            ev[2].record()
            torch.cuda.synchronize()
            sys.stderr.write("[mc trace] rank %d: samples %.3f ms, reduce %.3f ms\n" % (rank, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
        if rank == 0:
            r = PF.global_ncc(res["moved0"], res["moved0:mse"], 1.0, 1.0, True)      # ncc(var = std^2, mse), var.mean()
            result.update(var=res["moved0"] ** 2, mse=res["moved0:mse"], ncc=r[0], var_mean=r[1], std={k: res[k] for k in tracked})
    # >= 3 warm-up jobs: NCCL sets up the all_to_all / reduce_scatter / gather connections lazily over the first jobs
    mc_warmup = max(3, args.warmup)
    ms_total, clk = _timed(torch, dist, world, dev, local, rank, mc_warmup, args.steps, job)
    ms_step = ms_total / args.steps
    value = N * nvox / (ms_step * 1e-3) / 1e9
    # e2e: the pair comes from pinned host memory and the variance map goes back to the host, every job
    xp, yp = x_h.pin_memory(), y_h.pin_memory()
    var_host = torch.empty(size, dtype=torch.float32).pin_memory()

    def e2e_job():
        x.copy_(xp, non_blocking=True)
        y.copy_(yp, non_blocking=True)
        job()
        if rank == 0:
            var_host.copy_(result["var"].reshape(size), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e, _ = _timed(torch, dist, world, dev, local, rank, 2, e2e_steps, e2e_job)
    peak, peak_src = measured_peaks(ROOT)
    state = {"m": None}

    def prof_one():      # one sample + its moment updates, eager, single stream
        sampler.draw(0)
        plan1.run_forward(x, z)
        if state["m"] is None:
            state["m"] = {k: mc.MCMoments(v.shape, dev) for k, v in (("moved0", plan1.moved[0][0]), ("final0", plan1.final[0][0]))}
        state["m"]["moved0"].update(plan1.moved[0][0])
        state["m"]["final0"].update(plan1.final[0][0])
    plan1 = HotPathPlan(size, total, latent, batch=1, device=dev, with_reg=False, multi_stream=False)
    prof_one()
    breakdown, roofline, _ = kernel_breakdown(prof_one, 3, peak, peak_src)
    if rank == 0:
        per_sample_launches = launches_fwd
        _emit({"metric": "MC uncertainty inference throughput (128 deformation samples per pair, variance map on rank 0)",
               "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": mc_warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "samples_per_s": N / (ms_step * 1e-3),
               "config": {"workload": args.workload, "samples": N, "samples_per_gpu": len(ids), "voxels_per_sample": nvox,
                          "levels": "%d total / %d latent, level_res, 7 integration steps" % (total, latent), "launch": launch_mode,
                          "tracked_maps": tracked + ["moved0:sqerr"],
                          "parallelism": "samples dealt round-robin to %d GPU(s); all_to_all of 1/W slices of (mean, M2) + reduce_scatter of the "
                                         "squared errors, per-slice Chan merge, one gather of the std / MSE slices to rank 0 (NCCL)" % world,
                          "l2": "per-sample working set ~0.9 GB >> 126 MB L2; no explicit flush"},
               "e2e": {"value": N * nvox / (ms_e2e / e2e_steps * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": 2 * nvox * 4,
                       "d2h_bytes_per_step": nvox * 4, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                       "api": "pulpo_b200.mc.mc_uncertainty + uncertainty_metrics (pinned host pair in, variance map out)"},
               "gpu_launches": per_sample_launches * len(ids) * args.steps, "gpu_launches_per_step": per_sample_launches * len(ids),
               "clocks": clk, "roofline": roofline, "kernels": breakdown,
               "uncertainty": {"ncc_var_mse": float(result["ncc"]), "var_mean": float(result["var_mean"])}})
    if world > 1:
        dist.destroy_process_group()


_RESULT_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write banners there too (NCCL prints its version
    on the first communicator when NCCL_DEBUG is set), so keep the real stdout for the result line and
    point fd 1 at stderr for everything else."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="oasis_160x192x224_5tot_4lat", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=1, help="pairs per GPU")
    ap.add_argument("--engine", default="plan", choices=["plan", "autograd"],
                    help="plan: pre-planned multi-stream C-ABI sequence; autograd: the drop-in nn.Modules")
    ap.add_argument("--fuse-reg", type=int, default=1, help="plan engine: L2_reg fused into the warp kernels (1) or separate kernels (0)")
    ap.add_argument("--dpos", type=int, default=1, help="plan engine: the warp forward stores d out / d df and the backward is a streaming product inside the regulariser's pass (1) or the gather-form warp backward (0)")
    ap.add_argument("--fuse-combine", type=int, default=0, help="plan engine: pyramid combination inside the integration launches (1) or separate launches (0)")
    ap.add_argument("--pool-pyramid", type=int, default=1, help="plan engine: moving-image pyramid in one launch (1) or one launch per level (0)")
    ap.add_argument("--aux-early", type=int, default=0, help="plan engine: pyramid + KL start with the step (1) or after the integration (0)")
    ap.add_argument("--reg-coarse", type=int, default=0, help="plan engine: regulariser of the x2-resized level in closed form on the coarse field (aux stream), backward = resize adjoint of gmoved * dpos formed on the fly (1) or a full-resolution regulariser + product pass (0)")
    ap.add_argument("--aux-after", default="up2", choices=["integ", "up2", "warp"],
                    help="plan engine: the aux stream (moving-image pyramid, KL) starts after the integration, after level 0's output resize or after level 0's warp")
    ap.add_argument("--field-sigma-vox", type=float, default=None,
                    help="smoothing of the synthetic velocity fields in each level's own voxels (default: the same physical "
                         "length scale at every level, 8 full-resolution voxels, which makes the combined field fold)")
    ap.add_argument("--field-max-abs", type=float, default=3.0,
                    help="max |v| of every level's synthetic velocity field in that level's voxels (SURVEY 8d: 3; the "
                         "coarse-to-fine sum then reaches ~45 level-0 voxels)")
    ap.add_argument("--ref-budget-s", type=float, default=600.0,
                    help="--impl reference: wall-clock budget; passes of the SAME workload are dropped (never the volume shrunk) to fit")
    ap.add_argument("--grad-allreduce", type=int, default=0,
                    help="N > 1: also all-reduce a 58.7 MB stand-in for the conv gradients every step, overlapped with the hot path (config 5)")
    ap.add_argument("--seed0", type=int, default=0, help="mc128: Philox seed of sample 0 (sample i uses seed0 + i)")
    ap.add_argument("--samples", type=int, default=128, help="mc128: MC deformation samples per pair (dealt to the ranks)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true", help="N > 1: skip the multi-GPU MC-sharding self-check (untimed)")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "vecint_fullres":
        run_vecint_fullres(args)
    elif args.workload == "mc128":
        run_mc128(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
