"""HotPathPipeline -- end-to-end streaming of hot-path steps from HOST memory.

The step itself (HotPathPlan, one CUDA graph) takes ~1 ms at config 2, but its inputs are 90 MB
(x, y: 27.5 MB each; velocity fields, mu, sigma of four levels) and live on the host in the
end-to-end setting (data loader / upstream stage).  Copying them tensor by tensor and then
computing serialises ~2.5 ms of PCIe time in front of every step.  Here:

  * a step's inputs travel as ONE packed pinned buffer -> one cudaMemcpyAsync (full PCIe rate,
    no per-tensor launch overhead); ``host_batch()`` hands out the pinned buffer with named views
    to fill in place;
  * device input slots are double-buffered and fed from a dedicated copy stream, so the H2D copy of
    step i+1 overlaps the compute of step i (events order copy -> compute -> slot reuse);
  * the compute is a CUDA-graph replay per slot; the loss scalars are copied D2H into pinned
    memory on the compute stream and read when the caller asks for the result.

Gradients stay on the device (``plan.gdf / gmu / gsigma``) for the upstream (PyTorch) backward.
Reference call structure: PULPo.training_step (src/models.py:134-164) minus the convolutions.
"""
from __future__ import annotations

import torch

from .plan import HotPathPlan
from .synthetic import level_sizes


class HostBatch:
    """One step's inputs in a single pinned fp32 buffer, with views in the plan's input order."""

    def __init__(self, layout, nfloats):
        self.buf = torch.empty(nfloats, dtype=torch.float32).pin_memory()
        self.views = {name: self.buf[o:o + n].view(shape) for name, (o, n, shape) in layout.items()}

    def __getitem__(self, name):
        return self.views[name]

    def fill(self, x, y, dfs, mus, sigmas):
        self.views["x"].copy_(x)
        self.views["y"].copy_(y)
        for l in dfs:
            self.views["df%d" % l].copy_(dfs[l])
            self.views["mu%d" % l].copy_(mus[l])
            self.views["sigma%d" % l].copy_(sigmas[l])
        return self


class HotPathPipeline:
    def __init__(self, input_size, total_levels, latent_levels, batch=1, depth=2, device=None, graph=True, **plan_kw):
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.L, self.B, self.depth = latent_levels, batch, depth
        self.plan = HotPathPlan(input_size, total_levels, latent_levels, batch=batch, device=self.dev, **plan_kw)
        sizes = level_sizes(list(input_size), total_levels)
        lk = total_levels - latent_levels
        self.layout, off = {}, 0

        def add(name, shape):
            nonlocal off
            n = 1
            for s in shape:
                n *= int(s)
            self.layout[name] = (off, n, tuple(int(s) for s in shape))
            off += (n + 3) // 4 * 4            # keep every view 16-byte aligned

        add("x", (batch, 1, *input_size))
        add("y", (batch, 1, *input_size))
        for l in range(latent_levels):
            for name in ("df", "mu", "sigma"):
                add("%s%d" % (name, l), (batch, 3, *sizes[lk + l]))
        self.nfloats = off
        self.h2d_bytes = 4 * off
        self.slots = []
        for _ in range(depth):
            buf = torch.empty(off, dtype=torch.float32, device=self.dev)
            v = {name: buf[o:o + n].view(shape) for name, (o, n, shape) in self.layout.items()}
            self.slots.append({"buf": buf, "v": v, "graph": None, "free": torch.cuda.Event(), "ready": torch.cuda.Event()})
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.results = [torch.zeros(4, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.dres = torch.zeros(4, dtype=torch.float32, device=self.dev)
        self.use_graph = graph
        self.n = 0

    def host_batch(self) -> HostBatch:
        return HostBatch(self.layout, self.nfloats)

    def _compute(self, slot):
        v = slot["v"]
        L = self.L
        total = self.plan.run(v["x"], v["y"], {l: v["df%d" % l] for l in range(L)}, {l: v["mu%d" % l] for l in range(L)},
                              {l: v["sigma%d" % l] for l in range(L)})
        # (total, kl, recon, reg) for the D2H read, by the library's own one-block kernel (no ATen op in the graph)
        from ._lib import check
        import ctypes
        check(self.plan.lib.pulpo_loss_total(ctypes.c_void_p(self.plan.losses.data_ptr()), 3, L,
                                             ctypes.c_void_p(self.dres.data_ptr()), ctypes.c_void_p(self.dres.data_ptr() + 4), 0,
                                             ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)), "loss_total")

    def _capture(self, slot):
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            slot["buf"].fill_(0.5)             # benign values for the warm-up passes
            for _ in range(2):
                self._compute(slot)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._compute(slot)
        slot["graph"] = g

    def submit(self, host: HostBatch) -> int:
        """Enqueue one step: H2D of ``host`` (copy stream), compute (current stream), D2H of the loss
        scalars.  Returns a ticket for ``result``.  The host buffer may be refilled once
        ``result(ticket)`` -- or any later ticket -- has returned."""
        i = self.n % self.depth
        slot = self.slots[i]
        cur = torch.cuda.current_stream(self.dev)
        if self.use_graph and slot["graph"] is None:
            self._capture(slot)            # first use of this slot: capture its graph
        with torch.cuda.stream(self.copy_stream):
            if self.n >= self.depth:
                self.copy_stream.wait_event(slot["free"])    # the compute that last read this slot is done
            slot["buf"].copy_(host.buf, non_blocking=True)
            slot["ready"].record(self.copy_stream)
        cur.wait_event(slot["ready"])
        if self.use_graph:
            slot["graph"].replay()
        else:
            self._compute(slot)
        slot["free"].record(cur)
        self.results[i].copy_(self.dres, non_blocking=True)
        self.done[i].record(cur)
        self.n += 1
        return self.n - 1

    def result(self, ticket: int):
        """Block until step ``ticket`` is done; returns (total, kl, recon, reg) as Python floats."""
        if ticket < self.n - self.depth or ticket >= self.n:
            raise RuntimeError("HotPathPipeline.result: ticket %d is no longer (or not yet) in flight" % ticket)
        i = ticket % self.depth
        self.done[i].synchronize()
        r = self.results[i]
        return float(r[0]), float(r[1]), float(r[2]), float(r[3])
