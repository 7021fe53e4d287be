"""Whole-model drop-in test: the UNMODIFIED reference model
`src.models.PULPo` (/root/reference/src/models.py:24-196; staged for the GPU box by oracle/stage_ref.py) runs one
`training_step` on CUDA twice -- once as shipped, once with pulpo_b200's modules and losses bound to the names the
reference imports (`src/models.py:11-20`, `src/components/pulpo.py:7`; exactly the edit INTEGRATION.md describes) --
with the same weights and the same noise.  Loss, per-level loss terms, warped images / fields and the gradients of
the conv encoder/decoder parameters (which receive their upstream gradient through the hot path) must agree.
Skipped when no reference tree is available."""
import contextlib

import pytest
import torch

pytestmark = pytest.mark.gpu

FEEDBACK = ["samples", "velocity_fields", "individual_dfs", "combined_dfs", "final_dfs", "transformed"]


@contextlib.contextmanager
def patched_reference(cp, md):
    """Bind pulpo_b200's drop-ins to the names the reference's modules imported."""
    from pulpo_b200 import losses as PL, network_blocks as PN
    blocks = {n: getattr(PN, n) for n in ("SpatialTransformer", "ResizeTransform", "DFAdder", "VecInt")}
    model_names = dict(blocks, gauss_sampler=PN.gauss_sampler,
                       **{n: getattr(PL, n) for n in ("HierarchicalKLLoss", "HierarchicalReconstructionLoss",
                                                      "HierarchicalRegularization", "L2_reg", "JDetStd",
                                                      "KL_two_gauss_with_diag_cov")})
    saved = []
    try:
        for mod, names in ((cp, blocks), (md, model_names)):
            for n, v in names.items():
                saved.append((mod, n, getattr(mod, n)))
                setattr(mod, n, v)
        yield
    finally:
        for mod, n, v in saved:
            setattr(mod, n, v)


def _build(md, size, total, latent, df_resolution, dev):
    torch.manual_seed(1234)
    m = md.PULPo(total_levels=total, latent_levels=latent, beta=0.1, input_size=list(size), feedback=list(FEEDBACK),
                 df_resolution=df_resolution, n0=4, cp_depth=3)
    return m.to(dev).train()


def _step(model, batch, seed):
    torch.manual_seed(seed)                    # same Philox stream for gauss_sampler's randn_like in both models
    torch.cuda.manual_seed(seed)
    for p in model.parameters():
        p.grad = None
    x, y = batch[0], batch[1]
    down = model.downpath(x, y)
    torch.manual_seed(seed); torch.cuda.manual_seed(seed)
    outs = model.autoencoder(x, down)          # (mus, sigmas, samples, velocity_fields, individual, combined, final, y_hat)
    torch.manual_seed(seed); torch.cuda.manual_seed(seed)
    loss = model.training_step(batch, 0)
    loss.backward()
    return loss.detach(), outs, {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("df_resolution", ["level_res", "full_res"])
@pytest.mark.timeout(600)
def test_reference_training_step_with_pulpo_b200_modules(df_resolution):
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("no reference tree (run `python -m oracle.stage_ref` where /root/reference exists)")
    from pulpo_b200 import synthetic as syn
    nb, ls, cp, md = ref_import.load()
    dev = torch.device("cuda", 0)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.set_float32_matmul_precision("highest")
    try:
        size, total, latent = (32, 32, 32), 4, 3
        x, y = (t.to(dev) for t in syn.make_pair(size, 0))
        empty = torch.empty(0, device=dev)
        batch = (x, y, empty, empty, empty, empty, empty, empty)
        ref_model = _build(md, size, total, latent, df_resolution, dev)
        with patched_reference(cp, md):
            new_model = _build(md, size, total, latent, df_resolution, dev)
        from pulpo_b200 import network_blocks as PN
        assert isinstance(new_model.autoencoder.decoders[0].spatial_transform, PN.SpatialTransformer)
        assert isinstance(new_model.autoencoder.decoders[0].integrate, PN.VecInt)
        assert not isinstance(ref_model.autoencoder.decoders[0].integrate, PN.VecInt)
        # same weights: the reference's checkpoint (incl. its persistent `grid` buffers) loads strictly
        new_model.load_state_dict(ref_model.state_dict(), strict=True)

        l_ref, o_ref, g_ref = _step(ref_model, batch, 77)
        with patched_reference(cp, md):     # free functions (L2_reg, KL) are looked up at call time through the wrappers
            l_new, o_new, g_new = _step(new_model, batch, 77)

        rel = abs(float(l_new) - float(l_ref)) / abs(float(l_ref))
        assert rel <= 1e-5, "total loss %.8g vs %.8g (rel %.2e)" % (float(l_new), float(l_ref), rel)
        names = ("mu", "sigma", "sample", "velocity", "individual_df", "combined_df", "final_df", "y_hat")
        for k, name in enumerate(names):
            for l in o_ref[k]:
                err = float((o_new[k][l].detach() - o_ref[k][l].detach()).abs().max())
                assert err <= 1e-4, "%s[%d]: max-abs %.3e" % (name, l, err)
        assert set(g_new) == set(g_ref) and len(g_ref) > 20
        # conv biases in front of a BatchNorm have a mathematically zero gradient (pure cancellation noise in both
        # models), so every tensor is judged against its own scale plus a small share of the largest gradient
        gmax = max(float(g.abs().max()) for g in g_ref.values())
        worst, dot, na, nb_ = 0.0, 0.0, 0.0, 0.0
        for n in g_ref:
            scale = float(g_ref[n].abs().max())
            err = float((g_new[n] - g_ref[n]).abs().max())
            worst = max(worst, err / gmax)
            assert err <= 2e-3 * scale + 1e-4 * gmax, "grad of %s: max-abs %.3e vs scale %.3e (largest %.3e)" % (n, err, scale, gmax)
            dot += float((g_new[n].double() * g_ref[n].double()).sum())
            na += float((g_new[n].double() ** 2).sum()); nb_ += float((g_ref[n].double() ** 2).sum())
        cos = dot / (na ** 0.5 * nb_ ** 0.5)
        assert cos >= 1 - 1e-6, "cosine of the full parameter gradient %.9f" % cos
        print("drop-in %s: loss rel %.2e, worst parameter-gradient error %.2e of the largest gradient, cosine %.9f, %d tensors"
              % (df_resolution, rel, worst, cos, len(g_ref)))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic = old
        torch.autograd.set_detect_anomaly(False)     # PULPo.__init__ (src/models.py:50) switched it on process-wide
