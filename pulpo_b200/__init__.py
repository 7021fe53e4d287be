"""pulpo_b200 -- B200 (sm_100a) implementation of PULPo's dense-3D registration hot path.

Drop-in modules mirror the reference's names:
    pulpo_b200.network_blocks   SpatialTransformer, VecInt, ResizeTransform, DFAdder, gauss_sampler(_kl)
    pulpo_b200.losses           NCC_loss, KL_two_gauss_with_diag_cov, L2_reg, JDetStd, jacobian_det, Hierarchical*Loss
    pulpo_b200.components.pulpo SVFDecoder, PULPoPrior, moving_pyramid
    pulpo_b200.models           combine_dfs, transform_segmentation, loss_config, RegistrationHotPath
Beyond the reference's call structure:
    pulpo_b200.plan             HotPathPlan: one forward+backward as a multi-stream, CUDA-graph-capturable launch sequence
    pulpo_b200.pipeline         HotPathPipeline: steps streamed from pinned host memory (copy / compute overlap)
    pulpo_b200.mc               MC-sample sharding and streaming per-voxel moments (Evaluate.predict / uncertainty)
    pulpo_b200.hostmem          NUMA-local host buffers for one-process-per-GPU runs
All compute goes through libpulpo_b200.so (C ABI in include/pulpo_b200.h); there is no CPU or
PyTorch fallback -- a missing library raises at first use.
"""
__version__ = "0.1.0"
