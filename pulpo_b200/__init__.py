"""pulpo_b200 -- B200 (sm_100a) implementation of PULPo's dense-3D registration hot path.

Drop-in modules mirror the reference's names:
    pulpo_b200.network_blocks   SpatialTransformer, VecInt, ResizeTransform, DFAdder, gauss_sampler
    pulpo_b200.losses           NCC_loss, KL_two_gauss_with_diag_cov, L2_reg, Hierarchical*Loss
    pulpo_b200.components.pulpo SVFDecoder, PULPoPrior, moving_pyramid
    pulpo_b200.models           combine_dfs, transform_segmentation, loss_config, RegistrationHotPath
All compute goes through libpulpo_b200.so (C ABI in include/pulpo_b200.h); there is no CPU or
PyTorch fallback -- a missing library raises at first use.
"""
__version__ = "0.1.0"
