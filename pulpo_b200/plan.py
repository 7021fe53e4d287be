"""HotPathPlan -- the whole registration hot path (forward + backward) as one pre-planned,
multi-stream, CUDA-graph-capturable sequence of C-ABI launches.

Same arithmetic and the same kernels as the autograd drop-in modules (``RegistrationHotPath``),
but orchestrated B200-first:
  * every buffer (outputs, saved integration states, NCC coefficient volumes, gradients,
    reduction workspaces) is allocated once, so a step is pure kernel launches and can be
    replayed as a CUDA graph;
  * the latent levels are independent once the coarse-to-fine field combination is done, so
    each level runs on its own stream: the three small levels (<= 107 k voxels, pure launch /
    grid-sync latency) hide behind the full-resolution level instead of serialising with it;
  * the loss weights (level weight, beta, gamma, lambda; reference src/models.py:104-123,157)
    are folded into the kernels' scale arguments, so the backward of every loss term starts
    right after its forward -- no scalar graph, no autograd bookkeeping kernels;
  * gradient sums that autograd would do in extra passes are folded into producers:
    L2-reg's gradient accumulates into the warp's field gradient, the Laplacian-pyramid
    up-sampling adjoint accumulates into the coarser level's gradient (``fuse_combine=True``
    moves the combination and its adjoint into the integration launches: fewer launches,
    measured slower).

Reference call structure reproduced: SVFDecoder.forward (src/components/pulpo.py:301-319) per
level, Autoencoder's moving pyramid (:168-179), HierarchicalReconstructionLoss / KLLoss /
Regularization (src/losses.py:225-355) with PULPo's weights (src/models.py:104-123).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import CPU_EXACT, FAST, check
from .models import loss_config
from .synthetic import level_sizes

_vp = ctypes.c_void_p


def _p(t, byte_offset=0):
    return None if t is None else _vp(t.data_ptr() + byte_offset)


class HotPathPlan:
    def __init__(self, input_size, total_levels, latent_levels, batch=1, beta=0.1, gamma=0.05, lamb=0.025,
                 with_reg=True, nsteps=7, coord_mode=CPU_EXACT, device=None, multi_stream=True, fuse_reg=True,
                 vecint_mode=CPU_EXACT, fuse_combine=False, pool_pyramid=True, aux_early=False, df_resolution="level_res",
                 dpos=True, aux_after="up2", reg_coarse=False):
        self.L = L = latent_levels
        self.B = B = batch
        self.lk = lk = total_levels - latent_levels
        self.nsteps, self.mode, self.with_reg = nsteps, coord_mode, with_reg
        self.vi_mode = vecint_mode   # CPU_EXACT: bit-identical fields; FAST is within 1e-4 but buys little (L1-pipe bound)
        self.fuse_reg = bool(fuse_reg and with_reg)   # L2_reg rides in the warp kernels (same field, same step)
        # dpos: the level warps' forward also stores d out / d df (the interpolant's spatial gradient at the sample
        # point, masked and scaled), so their backward is the streaming product gmoved * dpos -- formed inside the
        # regulariser's value + gradient pass over the same field when there is one.  False: the gather-form warp
        # backward (second position chain, second round of corner gathers), regulariser inside the warp kernels.
        self.dpos = bool(dpos) and (self.fuse_reg or not with_reg)
        # pyramid combination (and its adjoint) inside the integration launches: fewer launches (25 instead of 31) but
        # measured slower (0.70 vs 0.68 ms at config 2, re-measured with the 1.4 us barriers): the in-kernel phases run on
        # the cooperative grid's 113 k / 75 k threads and are latency-bound
        self.fuse_combine = bool(fuse_combine)
        self.dev = dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lib = _lib.lib()
        self.full = tuple(int(s) for s in input_size)
        sizes = level_sizes(self.full, total_levels)
        if df_resolution not in ("level_res", "full_res"):
            raise ValueError("df_resolution must be 'level_res' or 'full_res' (src/components/pulpo.py:146)")
        self.df_resolution = df_resolution
        full_res = df_resolution == "full_res"
        self.insz = {l: tuple(sizes[lk + l]) for l in range(L)}
        # outsize rule of Autoencoder.__init__ (src/components/pulpo.py:146): level 0 and, in full_res mode, every level
        # warp / compare at the input size; the integration always runs at the latent size (:297)
        self.outsz = {l: (self.full if (l == 0 or full_res) else self.insz[l]) for l in range(L)}
        self.ofac = {}
        for l in range(L):
            f = [o // i for o, i in zip(self.outsz[l], self.insz[l])]
            if any(i * f[0] != o for o, i in zip(self.outsz[l], self.insz[l])) or f[0] < 1 or f[0] > 64:
                raise NotImplementedError("HotPathPlan: the output resize of level %d must be one integer factor <= 64 "
                                          "(got %s -> %s)" % (l, self.insz[l], self.outsz[l]))
            self.ofac[l] = f[0]
            if l + 1 < L and tuple(2 * s for s in self.insz[l + 1]) != self.insz[l]:
                raise ValueError("level sizes must halve exactly (reference DFAdder would shape-mismatch)")
        win, kl_w, rec_w, reg_w = loss_config(L, lk, 3, df_resolution)
        self.win = win
        self.kl_weight = {l: float(beta * kl_w[l]) for l in range(L)}
        self.gamma_eff = {l: float(gamma * rec_w[l]) for l in range(L)}
        self.lamb_eff = {l: float(lamb * reg_w[l]) for l in range(L)}

        def buf(*shape):
            return torch.empty(shape, dtype=torch.float32, device=dev)

        lib = self.lib
        # moving pyramid: pooled[i] = x pooled (i+1) times; level l >= 1 uses pooled[lk + l - 1]
        self.pooled, s = [], self.full
        for _ in range((lk + L - 1) if (L > 1 and not full_res) else 0):   # full_res: every level warps x itself
            s = tuple((v + 1) // 2 for v in s)
            self.pooled.append(buf(B, 1, *s))
        self.yt = {l: buf(B, 1, *self.outsz[l]) for l in range(L) if self.outsz[l] != self.full}
        self.comb = {l: buf(B, 3, *self.insz[l]) for l in range(L - 1)}          # coarsest level aliases its input
        self.integ = {l: buf(B, 3, *self.insz[l]) for l in range(L)}
        self.final = {l: (buf(B, 3, *self.outsz[l]) if self.outsz[l] != self.insz[l] else self.integ[l]) for l in range(L)}
        self.moved = {l: buf(B, 1, *self.outsz[l]) for l in range(L)}
        self.vi_ws, self.vi_scr = {}, {}
        for l in range(L):
            d = self.insz[l]
            self.vi_ws[l] = buf(lib.pulpo_vecint_ws_bytes(nsteps, 1, B, *d) // 4)
            self.vi_scr[l] = buf(lib.pulpo_vecint_bwd_scratch_bytes(B, *d) // 4)
        self.abc = {l: buf(3, B, 1, *self.outsz[l]) for l in range(L)}
        self.gmoved = {l: buf(B, 1, *self.outsz[l]) for l in range(L)}
        self.dposb = {l: buf(B, 3, *self.outsz[l]) for l in range(L)} if self.dpos else {}
        # reg_coarse (opt-in): on levels with a x2 output resize the regulariser of the final field is evaluated in closed
        # form on the integrated (coarse) field -- value and gradient, three 1-D passes over 10 MB arrays on a side stream
        # as soon as the integration is done (pulpo_l2reg_up2_fwd_bwd) -- and the level's backward is ONE launch: the
        # resize adjoint of gmoved * dpos, formed on the fly and accumulated on top (pulpo_resize_up2_bwd_dpos), so
        # neither the full-resolution field gradient nor a full-resolution regulariser pass exists.  Measured at config 2:
        # the level-0 chain gets 12 us shorter, but the side-stream kernels cannot overlap the register-saturated
        # level-0 kernels (they wait for CTA slots, 62 us of work stretch to 170 us and delay the coarse levels):
        # 0.731 vs 0.683 ms per step.  Default: the regulariser's value + gradient pass over the final field carries the
        # product (pulpo_l2reg_fwd_bwd), then the resize adjoint.
        self.reg_coarse = {l: bool(reg_coarse) and self.dpos and self.with_reg and self.ofac[l] == 2 and self.insz[l][2] % 2 == 0
                           and self.insz[l][0] >= 2 for l in range(L)}
        self.reg_scr = {l: buf(lib.pulpo_l2reg_up2_scratch_bytes(B, 3, *self.insz[l]) // 4) for l in range(L) if self.reg_coarse[l]}
        self.gfinal = {l: buf(B, 3, *self.outsz[l]) for l in range(L)}
        self.ginteg = {l: (buf(B, 3, *self.insz[l]) if self.outsz[l] != self.insz[l] else self.gfinal[l]) for l in range(L)}
        self.gdf = {l: buf(B, 3, *self.insz[l]) for l in range(L)}
        self.gmu = {l: buf(B, 3, *self.insz[l]) for l in range(L)}
        self.gsigma = {l: buf(B, 3, *self.insz[l]) for l in range(L)}
        self.losses = torch.zeros(3, L, dtype=torch.float32, device=dev)        # rows: kl, recon, reg
        self.total = torch.zeros((), dtype=torch.float32, device=dev)
        self.running = torch.zeros(3, dtype=torch.float32, device=dev)          # per-term sums over the steps run so far
        # reduction workspaces (ticket counters must start at zero; kernels reset them)
        rbytes = lib.pulpo_reduce_ws_bytes()
        self.ws_klm = torch.zeros(lib.pulpo_kl_multi_ws_bytes(), dtype=torch.uint8, device=dev)
        self.ws_l2 = [torch.zeros(rbytes, dtype=torch.uint8, device=dev) for _ in range(L)]
        self.ws_ncc = [torch.zeros(lib.pulpo_ncc_ws_bytes(B, 1, *self.outsz[l]), dtype=torch.uint8, device=dev)
                       for l in range(L)]
        npool = len(self.pooled)
        edge = 4 if npool == 1 else (1 << npool)
        self.pool_pyramid = bool(pool_pyramid) and 1 <= npool <= 4 and all(s % edge == 0 for s in self.full)
        self.aux_early = bool(aux_early)
        # when the aux stream's work (moving-image pyramid, KL) may start if not with the step: after the integration
        # ("integ"), after level 0's output resize ("up2") or after level 0's warp ("warp")
        if aux_after not in ("integ", "up2", "warp"):
            raise ValueError("aux_after must be 'integ', 'up2' or 'warp'")
        self.aux_after = aux_after
        self.multi_stream = multi_stream
        # level 0 carries the critical path: its stream gets the highest priority, so that its persistent kernels are not
        # kept waiting for CTA slots by the small kernels of the coarser levels / the aux stream (PULPO_PLAN_PRIO=0: off)
        import os
        prio = os.environ.get("PULPO_PLAN_PRIO", "1") != "0"
        # streams: one per level, the aux stream (moving-image pyramid, KL), the coarse-grid regulariser's stream
        self.streams = [torch.cuda.Stream(device=dev, priority=(-1 if (prio and l == 0) else 0)) for l in range(L + 2)] \
            if multi_stream else None
        self.launches = 0

    # ------------------------------------------------------------------------------------------
    def _loss_ptr(self, row, l):
        return _p(self.losses, 4 * (row * self.L + l))

    def run(self, x, y, dfs, mus, sigmas):
        """Enqueue one forward+backward.  Inputs: contiguous fp32 CUDA tensors (x, y: [B,1,*full];
        dfs/mus/sigmas: {level: [B,3,*level_size]}).  Returns the 0-d total-loss tensor; per-term
        losses are in ``self.losses`` and gradients in ``self.gdf / gmu / gsigma``."""
        lib, L, B, mode = self.lib, self.L, self.B, self.mode
        cur = torch.cuda.current_stream(self.dev)
        ms = self.multi_stream
        lv = [self.streams[l] if ms else cur for l in range(L)]
        aux = self.streams[L] if ms else cur
        regs = self.streams[L + 1] if ms else cur
        n = [0]

        def call(fn, *a):
            n[0] += 1
            check(fn(*a), fn.__name__ if hasattr(fn, "__name__") else "")

        def H(s):
            return _vp(s.cuda_stream)

        start = torch.cuda.Event()
        start.record(cur)
        # ---- moving-image pyramid (pulpo.py:168-179) and the KL terms on the aux stream.  Both are independent of the
        #      fields; `aux_early` decides whether they start with the step (next to the combination / integration
        #      launch, which as a cooperative kernel needs every SM to itself) or right after the integration
        #      (next to the level-0 resize / warp kernels).
        ev_lx = {}

        def aux_work(after):
            if ms:
                aux.wait_event(after)
            if not self.pooled:
                pass
            elif self.pool_pyramid:
                outs = (ctypes.c_void_p * len(self.pooled))(*[t.data_ptr() for t in self.pooled])
                call(lib.pulpo_avgpool2_pyramid_fwd, _p(x), outs, len(self.pooled), B, 1, *self.full, H(aux))
                ev = torch.cuda.Event()
                ev.record(aux)
                for l in range(1, L):
                    ev_lx[l] = ev
            else:
                src, shape = x, self.full
                for i, dst in enumerate(self.pooled):
                    call(lib.pulpo_avgpool2_fwd, _p(src), _p(dst), B, 1, *shape, H(aux))
                    src, shape = dst, tuple(dst.shape[2:])
                    l = i - self.lk + 1
                    if l >= 1:
                        ev_lx[l] = torch.cuda.Event()
                        ev_lx[l].record(aux)
            # KL of every level, value and gradients, in one launch (losses.py:47-76 with the N(0,1) prior;
            # weight = level weight * beta)
            kl_arr = (_lib.KlLevel * L)()
            for l in range(L):
                nlat = 3 * self.insz[l][0] * self.insz[l][1] * self.insz[l][2]
                kl_arr[l] = _lib.KlLevel(mus[l].data_ptr(), sigmas[l].data_ptr(), self.gmu[l].data_ptr(),
                                         self.gsigma[l].data_ptr(), self.losses.data_ptr() + 4 * (0 * L + l), nlat,
                                         self.kl_weight[l])
            call(lib.pulpo_kl_n01_multi, kl_arr, L, 1e-10, B, _p(self.ws_klm), self.ws_klm.numel(), H(aux))
            ev = torch.cuda.Event()
            ev.record(aux)
            return ev

        ev_kl = aux_work(start) if self.aux_early else None
        lx = {0: x}
        for l in range(1, L):
            lx[l] = x if self.df_resolution == "full_res" else self.pooled[self.lk + l - 1]

        # ---- coarse-to-fine field combination (pulpo.py:308) on the current stream (or, fuse_combine, inside the
        #      integration launch)
        comb = {L - 1: dfs[L - 1]}
        for l in range(L - 2, -1, -1):
            if not self.fuse_combine:
                d = self.insz[l + 1]
                call(lib.pulpo_resize_up_fwd, _p(comb[l + 1]), _p(dfs[l]), _p(self.comb[l]), 2, 2.0, B, 3, *d, H(cur))
            comb[l] = self.comb[l]

        # ---- integrate every level in ONE cooperative launch (a cooperative kernel owns all SMs, so
        #      per-level launches would serialise; pulpo.py:311 for each decoder)
        lv_arr = (_lib.VecIntLevel * L)()
        for l in range(L):
            ws, scr = self.vi_ws[l], self.vi_scr[l]
            lv_arr[l] = _lib.VecIntLevel(comb[l].data_ptr(), self.integ[l].data_ptr(), ws.data_ptr(), ws.numel() * 4,
                                         scr.data_ptr(), scr.numel() * 4, *self.insz[l])
        if self.fuse_combine:
            indiv = (ctypes.c_void_p * L)(*[dfs[l].data_ptr() for l in range(L)])
            call(lib.pulpo_combine_vecint_multi_fwd, lv_arr, indiv, L, self.nsteps, 1, B, self.vi_mode, H(cur))
        else:
            call(lib.pulpo_vecint_multi_fwd, lv_arr, L, self.nsteps, 1, B, self.vi_mode, H(cur))
        ev_int = torch.cuda.Event()
        ev_int.record(cur)
        if not self.aux_early and self.aux_after == "integ":
            ev_kl = aux_work(ev_int)

        # ---- per level: resize, warp, losses and their backward, on the level's stream
        ev_done = {}
        late_aux = (not self.aux_early) and self.aux_after != "integ"
        order = ([0] + list(range(L - 1, 0, -1))) if late_aux else list(range(L - 1, -1, -1))   # level 0 first: the aux work hangs off it

        def aux_after_level0(s):
            ev = torch.cuda.Event()
            ev.record(s)
            return aux_work(ev)

        for l in order:
            s = lv[l]
            if ms:
                s.wait_event(start)
            din, dout = self.insz[l], self.outsz[l]
            hs = H(s)
            if ms:
                s.wait_event(ev_int)
            # resize the integrated field to the output size
            if dout != din:
                call(lib.pulpo_resize_up_fwd, _p(self.integ[l]), None, _p(self.final[l]), self.ofac[l], float(self.ofac[l]),
                     B, 3, *din, hs)
            ev_reg = None
            if self.reg_coarse[l]:
                if ms:
                    regs.wait_event(ev_int)
                call(lib.pulpo_l2reg_up2_fwd_bwd, _p(self.integ[l]), self.lamb_eff[l], self._loss_ptr(2, l), _p(self.ginteg[l]), 0,
                     _p(self.reg_scr[l]), self.reg_scr[l].numel() * 4, _p(self.ws_l2[l]), self.ws_l2[l].numel(), B, 3, *din, H(regs))
                ev_reg = torch.cuda.Event()
                ev_reg.record(regs)
            if late_aux and l == 0 and self.aux_after == "up2":
                ev_kl = aux_after_level0(s)
            # warp the (pooled) moving image
            if ms and l in ev_lx:
                s.wait_event(ev_lx[l])
            if self.dpos:
                call(lib.pulpo_warp3d_fwd_dpos, _p(lx[l]), _p(self.final[l]), _p(self.moved[l]), _p(self.dposb[l]), B, *dout,
                     mode, hs)
            elif self.fuse_reg:
                call(lib.pulpo_warp3d_l2reg_fwd, _p(lx[l]), _p(self.final[l]), _p(self.moved[l]), self.lamb_eff[l],
                     self._loss_ptr(2, l), _p(self.ws_l2[l]), self.ws_l2[l].numel(), B, 1, *dout, mode, hs)
            else:
                call(lib.pulpo_warp3d_fwd, _p(lx[l]), _p(self.final[l]), _p(self.moved[l]), None, B, 1, *dout, mode, hs)
            if late_aux and l == 0 and self.aux_after == "warp":
                ev_kl = aux_after_level0(s)
            # NCC against the resized fixed image (losses.py:313-318), backward immediately
            if dout != self.full:
                call(lib.pulpo_interp_size_fwd, _p(y), _p(self.yt[l]), B, 1, *self.full, *dout, hs)
                yt = self.yt[l]
            else:
                yt = y
            wsn = self.ws_ncc[l]
            call(lib.pulpo_ncc_fwd, _p(self.moved[l]), _p(yt), self._loss_ptr(1, l), _p(self.abc[l]), _p(wsn),
                 wsn.numel(), self.win[l], self.gamma_eff[l], B, 1, *dout, hs)
            call(lib.pulpo_ncc_bwd, _p(self.abc[l]), _p(self.moved[l]), _p(yt), None, _p(self.gmoved[l]),
                 self.win[l], self.gamma_eff[l], B, 1, *dout, hs)
            if self.reg_coarse[l]:
                if ms:
                    s.wait_event(ev_reg)
                call(lib.pulpo_resize_up2_bwd_dpos, _p(self.gmoved[l]), _p(self.dposb[l]), _p(self.ginteg[l]), 2.0, 1, B, *din, hs)
            elif self.dpos and self.with_reg:
                # L2_reg value + gradient and the warp's backward (gmoved * dpos) in one pass over the final field
                call(lib.pulpo_l2reg_fwd_bwd, _p(self.final[l]), self.lamb_eff[l], self._loss_ptr(2, l), _p(self.gmoved[l]),
                     _p(self.dposb[l]), _p(self.gfinal[l]), 0, _p(self.ws_l2[l]), self.ws_l2[l].numel(), B, 3, *dout, hs)
            elif self.dpos:
                call(lib.pulpo_warp3d_bwd_dpos, _p(self.gmoved[l]), _p(self.dposb[l]), _p(self.gfinal[l]), 0, B, *dout, hs)
            elif self.fuse_reg:
                call(lib.pulpo_warp3d_l2reg_bwd, _p(self.gmoved[l]), _p(lx[l]), _p(self.final[l]), _p(self.gfinal[l]),
                     self.lamb_eff[l], None, B, 1, *dout, mode, hs)
            else:
                call(lib.pulpo_warp3d_bwd, _p(self.gmoved[l]), _p(lx[l]), _p(self.final[l]), None, _p(self.gfinal[l]),
                     B, 1, *dout, mode, hs)
            if self.with_reg and not self.fuse_reg:   # L2_reg on the final field; its gradient accumulates into the warp's
                call(lib.pulpo_l2reg_fwd, _p(self.final[l]), self.lamb_eff[l], self._loss_ptr(2, l),
                     _p(self.ws_l2[l]), self.ws_l2[l].numel(), B, 3, *dout, hs)
                call(lib.pulpo_l2reg_bwd, None, _p(self.final[l]), self.lamb_eff[l], _p(self.gfinal[l]), 1, B, 3,
                     *dout, hs)
            if dout != din and not self.reg_coarse[l]:
                call(lib.pulpo_resize_up_bwd, _p(self.gfinal[l]), _p(self.ginteg[l]), self.ofac[l], float(self.ofac[l]), 0,
                     B, 3, *din, hs)
            ev_done[l] = torch.cuda.Event()
            ev_done[l].record(s)

        # ---- total loss on the aux stream (every term is known once the levels' forward kernels ran), next to
        #      the integration backward instead of behind it
        if ms:
            for l in range(L):
                aux.wait_event(ev_done[l])
            call(lib.pulpo_loss_total, _p(self.losses), 3, L, _p(self.total), _p(self.running), 1, H(aux))
            ev_total = torch.cuda.Event()
            ev_total.record(aux)

        # ---- backward of the integration, again one launch for all levels
        if ms:
            cur.wait_event(ev_kl)
            for l in range(L):
                cur.wait_event(ev_done[l])
        for l in range(L):
            lv_arr[l].inp, lv_arr[l].out = self.ginteg[l].data_ptr(), self.gdf[l].data_ptr()
        if self.fuse_combine:
            # ... and the adjoint of the combination (fine to coarse) inside the same launch
            call(lib.pulpo_combine_vecint_multi_bwd, lv_arr, L, self.nsteps, B, self.vi_mode, H(cur))
        else:
            call(lib.pulpo_vecint_multi_bwd, lv_arr, L, self.nsteps, B, self.vi_mode, H(cur))
            # ---- fine-to-coarse: the adjoint of the combination accumulates into the coarser gradient
            for l in range(1, L):
                call(lib.pulpo_resize_up_bwd, _p(self.gdf[l - 1]), _p(self.gdf[l]), 2, 2.0, 1, B, 3, *self.insz[l], H(cur))
        if ms:
            cur.wait_event(ev_total)
        else:
            call(lib.pulpo_loss_total, _p(self.losses), 3, L, _p(self.total), _p(self.running), 1, H(cur))
        self.launches = n[0]
        return self.total

    # ------------------------------------------------------------------------------------------
    def run_forward(self, x, dfs, levels=None):
        """Inference half of ``run`` (MC uncertainty sampling, evaluate.py:205-251 -> models.py:312-331,349-368):
        combine -> integrate -> output resize -> warp for every level, no losses, no backward.  ``dfs[l]`` are the
        sampled velocity fields.  Results in ``self.comb / self.integ / self.final / self.moved``; graph-capturable."""
        lib, L, B, mode = self.lib, self.L, self.B, self.mode
        cur = torch.cuda.current_stream(self.dev)
        ms = self.multi_stream
        lv = [self.streams[l] if ms else cur for l in range(L)]
        aux = self.streams[L] if ms else cur
        n = [0]

        def call(fn, *a):
            n[0] += 1
            check(fn(*a), fn.__name__ if hasattr(fn, "__name__") else "")

        def H(s):
            return _vp(s.cuda_stream)

        start = torch.cuda.Event()
        start.record(cur)
        ev_lx = None
        if self.pooled:
            if ms:
                aux.wait_event(start)
            if self.pool_pyramid:
                outs = (ctypes.c_void_p * len(self.pooled))(*[t.data_ptr() for t in self.pooled])
                call(lib.pulpo_avgpool2_pyramid_fwd, _p(x), outs, len(self.pooled), B, 1, *self.full, H(aux))
            else:
                src, shape = x, self.full
                for dst in self.pooled:
                    call(lib.pulpo_avgpool2_fwd, _p(src), _p(dst), B, 1, *shape, H(aux))
                    src, shape = dst, tuple(dst.shape[2:])
            ev_lx = torch.cuda.Event()
            ev_lx.record(aux)
        lx = {0: x}
        for l in range(1, L):
            lx[l] = x if self.df_resolution == "full_res" else self.pooled[self.lk + l - 1]
        comb = {L - 1: dfs[L - 1]}
        for l in range(L - 2, -1, -1):
            call(lib.pulpo_resize_up_fwd, _p(comb[l + 1]), _p(dfs[l]), _p(self.comb[l]), 2, 2.0, B, 3, *self.insz[l + 1], H(cur))
            comb[l] = self.comb[l]
        lv_arr = (_lib.VecIntLevel * L)()
        for l in range(L):
            ws, scr = self.vi_ws[l], self.vi_scr[l]
            lv_arr[l] = _lib.VecIntLevel(comb[l].data_ptr(), self.integ[l].data_ptr(), ws.data_ptr(), ws.numel() * 4,
                                         scr.data_ptr(), scr.numel() * 4, *self.insz[l])
        call(lib.pulpo_vecint_multi_fwd, lv_arr, L, self.nsteps, 1, B, self.vi_mode, H(cur))
        ev_int = torch.cuda.Event()
        ev_int.record(cur)
        for l in range(L - 1, -1, -1):
            s = lv[l]
            if ms:
                s.wait_event(ev_int)
                if ev_lx is not None and l > 0:
                    s.wait_event(ev_lx)
            din, dout = self.insz[l], self.outsz[l]
            if dout != din:
                call(lib.pulpo_resize_up_fwd, _p(self.integ[l]), None, _p(self.final[l]), self.ofac[l], float(self.ofac[l]),
                     B, 3, *din, H(s))
            call(lib.pulpo_warp3d_fwd, _p(lx[l]), _p(self.final[l]), _p(self.moved[l]), None, B, 1, *dout, mode, H(s))
            if ms:
                ev = torch.cuda.Event()
                ev.record(s)
                cur.wait_event(ev)
        self.launches = n[0]
        return self.moved

    # convenience -----------------------------------------------------------------------------
    def outputs(self):
        return {"combined": dict(self.comb), "final": dict(self.final), "moved": dict(self.moved)}
