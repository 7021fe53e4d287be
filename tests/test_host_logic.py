"""CPU: host-side logic of the drop-in layer (no kernels run)."""
import math
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_loss_config_matches_reference_formulas():
    from oracle import torch_ref as T
    from pulpo_b200.models import loss_config
    for L, lk in [(4, 1), (3, 1), (1, 0), (4, 0), (2, 1)]:
        assert loss_config(L, lk) == T.loss_weights(L, lk)
    win, kl, rec, reg = loss_config(4, 1)
    assert win == {0: 9, 1: 7, 2: 5, 3: 3}
    assert rec == {0: 0.5, 1: 8.0, 2: 64.0, 3: 512.0} and reg[0] == 0.125 and kl[3] == 512.0


def test_loss_config_against_live_reference_when_present():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    from pulpo_b200.models import loss_config
    nb, ls, cp, md = ref_import.load()
    fb = ["samples", "velocity_fields", "individual_dfs", "combined_dfs", "final_dfs", "transformed"]
    model = md.PULPo(total_levels=4, latent_levels=3, beta=0.1, input_size=[16, 16, 16], feedback=fb, n0=2, cp_depth=0)
    torch.autograd.set_detect_anomaly(False)
    win, kl, rec, reg = loss_config(3, 1)
    assert model.hierarchical_recon_loss.window_size == win
    assert model.hierarchical_kl_loss.weight_dict == kl
    assert model.hierarchical_recon_loss.weight_dict == rec
    assert model.hierarchical_regularization.weight_dict == reg


def test_level_sizes_and_synthetic_inputs():
    from pulpo_b200 import synthetic as syn
    assert syn.level_sizes([160, 192, 224], 5) == {0: [160, 192, 224], 1: [80, 96, 112], 2: [40, 48, 56],
                                                   3: [20, 24, 28], 4: [10, 12, 14]}
    assert syn.level_sizes([5, 7, 9], 3)[2] == [2, 2, 3]          # ceil
    x, y, d, m, s = syn.make_hot_path_inputs([16, 16, 16], 3, 2, seed=3)
    x2, y2, d2, _, _ = syn.make_hot_path_inputs([16, 16, 16], 3, 2, seed=3)
    assert torch.equal(x, x2) and torch.equal(d[1], d2[1])        # deterministic
    assert x.min() == 0 and float((x == 0).float().mean()) > 0.2  # exact-zero background
    assert abs(float(d[0].abs().max()) - 3.0) < 1e-4 and float(s[0].min()) >= 0.2


def test_modules_fail_loudly_without_cuda_tensors():
    from pulpo_b200.network_blocks import ResizeTransform, SpatialTransformer, VecInt
    from pulpo_b200.losses import NCC_loss, HierarchicalReconstructionLoss
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SpatialTransformer((4, 4, 4))(torch.zeros(1, 3, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VecInt((4, 4, 4), 7)(torch.zeros(1, 3, 4, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        NCC_loss(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4))
    with pytest.raises(NotImplementedError):
        SpatialTransformer((4, 4))
    with pytest.raises(NotImplementedError):
        ResizeTransform(2.0, 3)(torch.zeros(1, 3, 4, 4, 4))      # down-sampling is not on the hot path
    with pytest.raises(NotImplementedError):
        HierarchicalReconstructionLoss(["mse"], {0: 1.0}, False, 3, {0: 9})
    x = torch.zeros(1, 3, 4, 4, 4)
    assert ResizeTransform(1.0, 3)(x) is x                        # factor 1 is a strict no-op


def test_reference_checkpoints_load_strict():
    """Reference checkpoints carry persistent identity-grid buffers (src/network_blocks.py:99)."""
    from pulpo_b200.components.pulpo import SVFDecoder
    dec = SVFDecoder(3, [4, 6, 8], [8, 12, 16], "level_res", n0=2, cp_depth=0)
    sd = {"spatial_transform.grid": torch.zeros(1, 3, 8, 12, 16), "integrate.transformer.grid": torch.zeros(1, 3, 4, 6, 8)}
    res = dec.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for attr in ("spatial_transform", "integrate", "resizer_level", "resizer_output", "velocity_field"):
        assert hasattr(dec, attr)
    assert dec.resizer_output.factor == 2.0 and dec.resizer_level.factor == 2.0
    dec3 = SVFDecoder(3, [4, 6, 8], [4, 6, 8], "level_res", n0=4, cp_depth=3)
    keys = list(dec3.state_dict().keys())
    assert "velocity_field._op.0._op.0.weight" in keys and "velocity_field._op.2.weight" in keys


def test_algorithmic_bytes_match_survey_totals():
    """Byte accounting used for roofline.achieved reproduces SURVEY.md 8(d): 1520.1 MB per
    config-2 pair (fwd+bwd incl. L2_reg).  Pure host arithmetic over the plan's launch list."""
    from pulpo_b200.roofline import algo_bytes
    full, L, lk, B = (160, 192, 224), 4, 1, 1
    sizes = {l: tuple(v // 2 ** (l + lk) for v in full) for l in range(L)}
    N0 = full[0] * full[1] * full[2]
    total = 0
    P = None
    for l in range(L):
        d = sizes[l]
        out = full if l == 0 else d
        n_lat = 3 * d[0] * d[1] * d[2]
        if l + 1 < L:
            c = sizes[l + 1]
            total += algo_bytes("pulpo_resize_up_fwd", (P, 1, P, 2, 2.0, B, 3) + c)
            total += algo_bytes("pulpo_resize_up_bwd", (P, P, 2, 2.0, 0, B, 3) + c)
        total += algo_bytes("pulpo_vecint_fwd", (P, P, P, 0, 7, 1, B) + d)
        total += algo_bytes("pulpo_vecint_bwd", (P, P, P, P, 0, 7, B) + d)
        if l == 0:
            total += algo_bytes("pulpo_resize_up_fwd", (P, None, P, 2, 2.0, B, 3) + d)
            total += algo_bytes("pulpo_resize_up_bwd", (P, P, 2, 2.0, 0, B, 3) + d)
        total += algo_bytes("pulpo_warp3d_fwd", (P, P, P, None, B, 1) + out)
        total += algo_bytes("pulpo_warp3d_bwd", (P, P, P, None, 1, B, 1) + out)
        total += algo_bytes("pulpo_ncc_fwd", (P, P, P, P, P, 0, 9, 0.05, B, 1) + out)
        total += algo_bytes("pulpo_ncc_bwd", (P, P, P, None, P, 9, 0.05, B, 1) + out)
        total += algo_bytes("pulpo_kl_diag_fwd", (P, P, None, None, 1e-10, 1.0, P, P, 0, B, n_lat))
        total += algo_bytes("pulpo_kl_diag_bwd", (None, P, P, None, None, 1e-10, 1.0, P, P, B, n_lat))
        total += algo_bytes("pulpo_l2reg_fwd", (P, 0.025, P, P, 0, B, 3) + out)
        total += algo_bytes("pulpo_l2reg_bwd", (None, P, 0.025, P, 0, B, 3) + out)
    # pyramids: 8*N0 reads once + writes of the lower levels
    per_voxel = total / N0
    assert 170 < per_voxel < 240, per_voxel      # SURVEY: 220.9 B/voxel (its count adds the pyramids)
