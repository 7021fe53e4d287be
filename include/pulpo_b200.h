/*
 * pulpo_b200.h -- C ABI of libpulpo_b200.so: the B200 (sm_100a) implementation of PULPo's
 * dense-3D registration hot path (warp, scaling-and-squaring, Laplacian-pyramid field
 * combination, local NCC + KL (+ L2) losses, MC moments).
 *
 * The reference (leonardsiegert/PULPo) is pure Python and has no FFI; its "operator API"
 * for this path is the constructor/forward signatures of a handful of nn.Modules and loss
 * functions.  Each entry point below names the reference interface it replaces
 * (path:line in the upstream repo).  INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - all tensors: contiguous fp32, layout [B, C, D0, D1, D2], D2 innermost; field channel a
 *     displaces along spatial axis a, in voxels (src/network_blocks.py:94-103)
 *   - every pointer is a DEVICE pointer owned by the caller (outputs and workspaces too);
 *     the library never allocates device memory, never synchronises, keeps no mutable
 *     global state and enqueues only on `stream` (a cudaStream_t passed as void*)
 *   - return value: PULPO_OK (0) or a negative pulpo_status; no C++ exception crosses
 *   - `*_ws_bytes` helpers are pure host arithmetic
 */
#ifndef PULPO_B200_H
#define PULPO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PULPO_B200_VERSION 100 /* 0.1.0 */

typedef void *pulpo_stream_t; /* cudaStream_t */

typedef enum pulpo_status {
    PULPO_OK = 0,
    PULPO_ERR_NULL_POINTER = -1,
    PULPO_ERR_INVALID_SHAPE = -2,
    PULPO_ERR_UNSUPPORTED = -3,
    PULPO_ERR_WORKSPACE = -4,
    PULPO_ERR_CUDA = -5
} pulpo_status;

/* How the sample position is rounded.  Integer corner indices are
 * bit-exact against torch-CPU in mode 0 and against torch-CUDA in mode 1. */
#define PULPO_COORD_CPU_EXACT 0 /* loc/(S-1) true division; ((n+1)*S-1)/2 rounded op by op */
#define PULPO_COORD_CUDA_RCP 1  /* loc*(1/(S-1)); fma(n+1, S, -1)/2 */
#define PULPO_COORD_FAST 2      /* VecInt only: p = fma(loc, S/(S-1), -0.5), FMA interpolation; within
                                 * ~1e-5 voxel of either torch path (no index contract on this op) */

int pulpo_version(void);
const char *pulpo_strerror(int status);

/* ---- a2: SpatialTransformer.forward(df, moving_image)   src/network_blocks.py:101-121 ----
 * out[b,c,v] = trilinear sample of img[b,c] at p(v) = clamp(((2*((v+df)/(S-1)-.5)+1)*S-1)/2).
 * idx_dbg (nullable): int32 [B,3,D0,D1,D2], floor(p) per axis (the bit-exact contract). */
int pulpo_warp3d_fwd(const float *img, const float *df, float *out, int32_t *idx_dbg,
                     int B, int C, int D0, int D1, int D2, int coord_mode, pulpo_stream_t stream);

/* grid_sampler_3d_backward + the autograd chain of network_blocks.py:103-117.
 * gimg (nullable): ACCUMULATED into with red.global.add.f32 -> caller zeroes it.  The contributions of
 *   neighbouring lanes whose footprints share a corner column are combined by warp shuffle first (about 4
 *   instead of 8 reductions per voxel and channel, no intra-warp address collisions for a smooth field).
 * gdf (nullable): overwritten; zero where the border clamp is active. */
int pulpo_warp3d_bwd(const float *gout, const float *img, const float *df, float *gimg, float *gdf,
                     int B, int C, int D0, int D1, int D2, int coord_mode, pulpo_stream_t stream);

/* Same two entry points for an image whose spatial size [I0,I1,I2] differs from the field's [D0,D1,D2]:
 * grid_sample normalises with the FIELD size (the transformer's `size`, src/network_blocks.py:106-107) and
 * unnormalises, clamps and gathers with the IMAGE size; the output has the field's size.  The reference does
 * this when a level-sized field resamples a full-resolution image or segmentation in level_res mode
 * (evaluate.py:198, 240, 246).  img / gimg: [B,C,I0,I1,I2]; df / gdf / out / gout: field-sized. */
int pulpo_warp3d_fwd_img(const float *img, const float *df, float *out, int32_t *idx_dbg, int B, int C,
                         int D0, int D1, int D2, int I0, int I1, int I2, int coord_mode,
                         pulpo_stream_t stream);
int pulpo_warp3d_bwd_img(const float *gout, const float *img, const float *df, float *gimg,
                         float *gdf, int B, int C, int D0, int D1, int D2, int I0, int I1, int I2,
                         int coord_mode, pulpo_stream_t stream);

/* a2 with the backward's gather half prepared by the forward (one image channel, the hot path's level warps,
 * src/components/pulpo.py:317).  dpos [B,3,D0,D1,D2] = d out / d df: the spatial gradient of the interpolant at the
 * sample point, masked where the border clamp is active and scaled by the autograd chain of
 * network_blocks.py:103-117 -- exactly what pulpo_warp3d_bwd multiplies gout with.  The backward is then the
 * streaming product gdf (+)= gout * dpos (pulpo_warp3d_bwd_dpos, gout [B,1,...], gdf [B,3,...]), or the same product
 * inside the regulariser's pass over the field (pulpo_l2reg_fwd_bwd): no second position chain and no second round
 * of corner gathers in the backward. */
int pulpo_warp3d_fwd_dpos(const float *img, const float *df, float *out, float *dpos, int B, int D0, int D1,
                          int D2, int coord_mode, pulpo_stream_t stream);
int pulpo_warp3d_bwd_dpos(const float *gout, const float *dpos, float *gdf, int accumulate, int B, int D0,
                          int D1, int D2, pulpo_stream_t stream);

/* a2 fused with f-1 (L2_reg, src/losses.py:208-222): the warp and the regulariser read the same
 * full-resolution field in the same step (src/models.py:160-162).  reg_out (device scalar) =
 * L2_reg(df, lamb); ws: pulpo_reduce_ws_bytes() bytes, zeroed once by the caller. */
int pulpo_warp3d_l2reg_fwd(const float *img, const float *df, float *out, float lamb, float *reg_out,
                           void *ws, size_t ws_bytes, int B, int C, int D0, int D1, int D2,
                           int coord_mode, pulpo_stream_t stream);
/* gdf = d/d df [ <gout, warp(df, img)> + reg_gloss * L2_reg(df, lamb) ];  reg_gloss: device
 * scalar, nullable = 1. */
int pulpo_warp3d_l2reg_bwd(const float *gout, const float *img, const float *df, float *gdf, float lamb,
                           const float *reg_gloss, int B, int C, int D0, int D1, int D2, int coord_mode,
                           pulpo_stream_t stream);

/* ---- a3: VecInt.forward(vec)   src/network_blocks.py:173-177 ------------------------------
 * vec,out: [B,3,D0,D1,D2].  ws holds the integration states as [*,B,S] float4 (xyz + pad):
 * save_steps=1 keeps v_0..v_{nsteps-1} for the backward (nsteps states), save_steps=0 needs
 * two ping-pong states.  One cooperative launch runs all steps. */
size_t pulpo_vecint_ws_bytes(int nsteps, int save_steps, int B, int D0, int D1, int D2);
int pulpo_vecint_fwd(const float *vec, float *out, void *ws, size_t ws_bytes, int nsteps,
                     int save_steps, int B, int D0, int D1, int D2, int coord_mode,
                     pulpo_stream_t stream);
/* gvec = d loss/d vec.  Bits 8-9 of coord_mode select the scatter strategy (tuning; same result up to
 * summation order): 0 = z-carried (default), 0x100 = one reduction per corner, 0x200 = lane + plane combining.
 * `saved` = the ws of a save_steps=1 forward; scratch: 3 rotating
 * gradient states (pulpo_vecint_bwd_scratch_bytes). */
size_t pulpo_vecint_bwd_scratch_bytes(int B, int D0, int D1, int D2);
int pulpo_vecint_bwd(const float *gout, const void *saved, float *gvec, void *scratch,
                     size_t scratch_bytes, int nsteps, int B, int D0, int D1, int D2,
                     int coord_mode, pulpo_stream_t stream);

/* Several pyramid levels in ONE cooperative launch (SVFDecoder.integrate of every level,
 * src/components/pulpo.py:311; PULPo.combine_dfs loop, src/models.py:362-367).  A cooperative kernel owns
 * every SM, so per-level launches serialise and each pays its own grid barriers; the levels are
 * independent fields with the same step count.  `levels` is a HOST array (read during the call);
 * at most 6 levels. */
typedef struct pulpo_vecint_level {
    const float *in;      /* fwd: vec [B,3,D0,D1,D2]        bwd: gout */
    float *out;           /* fwd: integrated field          bwd: gvec */
    void *ws;             /* fwd: states (pulpo_vecint_ws_bytes)   bwd: the saved states of the forward (opaque layout) */
    size_t ws_bytes;      /* fwd only */
    void *scratch;        /* bwd only (pulpo_vecint_bwd_scratch_bytes) */
    size_t scratch_bytes; /* bwd only */
    int D0, D1, D2;
} pulpo_vecint_level;
/* Batches of large volumes (>= 2^19 voxels per item summed over the levels of the call) run one item after the
 * other so that one item's states stay L2-resident; small volumes share one launch.  The layout of the saved states
 * inside `ws` follows that decision ([B][steps][S] vs [steps][B][S]) and is opaque: the backward MUST be called with
 * the same (B, nsteps, level set) as the forward that filled `ws` -- pulpo_vecint_fwd with pulpo_vecint_bwd,
 * pulpo_vecint_multi_fwd with pulpo_vecint_multi_bwd over the same levels[] (what _VecInt and HotPathPlan do).
 * Mixing them for B > 1 reads the states of the wrong item. */
int pulpo_vecint_multi_fwd(const pulpo_vecint_level *levels, int nlevels, int nsteps, int save_steps,
                           int B, int coord_mode, pulpo_stream_t stream);
int pulpo_vecint_multi_bwd(const pulpo_vecint_level *levels, int nlevels, int nsteps, int B,
                           int coord_mode, pulpo_stream_t stream);

/* a6/a7 + a3 in one launch: the coarse-to-fine combination of the Laplacian pyramid
 * (SVFDecoder.forward, src/components/pulpo.py:308; PULPo.combine_dfs, src/models.py:356-367)
 * folded into the integration launch.  levels[] is a x2 pyramid ordered fine to coarse (level l exactly
 * twice the size of level l+1); indiv: HOST array of nlevels device pointers, the levels' individual fields.
 * Forward: combined_l = 2 * up2(combined_{l+1}) + indiv[l] is WRITTEN to levels[l].in for l < nlevels-1
 * (the coarsest combined field is indiv[nlevels-1] itself; its levels[].in is not touched) and integrated
 * into levels[l].out.  Backward: levels[l].in = gradient w.r.t. the integrated field, levels[l].out =
 * gradient w.r.t. indiv[l] (= w.r.t. combined_l), including 2 * up2^T of the finer level's.
 * Six fewer launches per step; at config 2 the separate launches are nevertheless faster (they overlap
 * with the other streams' work, the in-kernel phases are latency-bound on the cooperative grid), so
 * HotPathPlan uses this only on request (fuse_combine=True). */
int pulpo_combine_vecint_multi_fwd(const pulpo_vecint_level *levels, const float *const *indiv,
                                   int nlevels, int nsteps, int save_steps, int B, int coord_mode,
                                   pulpo_stream_t stream);
int pulpo_combine_vecint_multi_bwd(const pulpo_vecint_level *levels, int nlevels, int nsteps, int B,
                                   int coord_mode, pulpo_stream_t stream);

/* ---- a4 + a5: ResizeTransform.forward (factor>1) fused with DFAdder.forward ---------------
 * src/network_blocks.py:138-150, :152-158; used at src/components/pulpo.py:308,314 and
 * src/models.py:356-367.   out = trilinear_up_f(scale * x) (+ addend),  x: [B,C,d0,d1,d2],
 * out/addend: [B,C,f*d0,f*d1,f*d2].  factor: integer >= 2.  addend nullable. */
int pulpo_resize_up_fwd(const float *x, const float *addend, float *out, int factor, float scale,
                        int B, int C, int d0, int d1, int d2, pulpo_stream_t stream);
/* exact adjoint w.r.t. x in gather form (no atomics); gout: [B,C,f*d0,f*d1,f*d2].
 * accumulate != 0: gx += ... (folds the Laplacian-pyramid gradient sum into this pass). */
int pulpo_resize_up_bwd(const float *gout, float *gx, int factor, float scale, int accumulate,
                        int B, int C, int d0, int d1, int d2, pulpo_stream_t stream);

/* The x2 output resize's adjoint (src/components/pulpo.py:314, autograd) with its input gradient formed on the fly
 * from the warp's stored dpos (pulpo_warp3d_fwd_dpos):  gx (+)= scale * up2^T( gout * dpos ),  gout [B,1,2d0,2d1,2d2]
 * (upstream gradient of the moved image), dpos [B,3,2d0,2d1,2d2], gx [B,3,d0,d1,d2] -- the full-resolution field
 * gradient never exists in HBM.  Needs d2 even, d0 >= 2, 16-byte aligned pointers (PULPO_ERR_UNSUPPORTED otherwise:
 * use pulpo_warp3d_bwd_dpos + pulpo_resize_up_bwd). */
int pulpo_resize_up2_bwd_dpos(const float *gout, const float *dpos, float *gx, float scale, int accumulate,
                              int B, int d0, int d1, int d2, pulpo_stream_t stream);

/* ---- a10: F.interpolate(y, size=..., trilinear, align_corners=False)  src/losses.py:313 --- */
int pulpo_interp_size_fwd(const float *x, float *out, int B, int C, int i0, int i1, int i2,
                          int o0, int o1, int o2, pulpo_stream_t stream);

/* ---- a8: avg_pool3d(2, 2, ceil_mode=True)  src/components/pulpo.py:171-179 ---------------- */
int pulpo_avgpool2_fwd(const float *x, float *out, int B, int C, int D0, int D1, int D2,
                       pulpo_stream_t stream);
/* The whole moving-image pyramid in one launch: outs[i] (HOST array of nlevels <= 4 device pointers) =
 * avg_pool3d(2,2) applied i+1 times; the input is read once.  Needs D0, D1, D2 divisible by 2^nlevels
 * (by 4 for nlevels == 1) and a 16-byte aligned x, else PULPO_ERR_UNSUPPORTED: use pulpo_avgpool2_fwd per
 * level.  Bit-identical to the chain of pulpo_avgpool2_fwd calls. */
int pulpo_avgpool2_pyramid_fwd(const float *x, float *const *outs, int nlevels, int B, int C, int D0,
                               int D1, int D2, pulpo_stream_t stream);

/* ---- a9: NCC_loss(y_pred, y_true, win_size, gamma)   src/losses.py:85-135 ------------------
 * loss (device scalar) = -gamma/B * sum cc.  abc (nullable): [3][B,C,S] coefficient volumes
 * the backward box-filters.  win odd, 3..11.  ws: per-CTA partial sums. */
size_t pulpo_ncc_ws_bytes(int B, int C, int D0, int D1, int D2);
int pulpo_ncc_fwd(const float *pred, const float *target, float *loss, float *abc, void *ws,
                  size_t ws_bytes, int win, float gamma, int B, int C, int D0, int D1, int D2,
                  pulpo_stream_t stream);
/* gpred = gloss * (-gamma/B) * (I*Box(a) + Box(b) + 2*J*Box(c)); gloss: device scalar (nullable = 1) */
int pulpo_ncc_bwd(const float *abc, const float *pred, const float *target, const float *gloss,
                  float *gpred, int win, float gamma, int B, int C, int D0, int D1, int D2,
                  pulpo_stream_t stream);

/* ---- a11: KL_two_gauss_with_diag_cov(mu0, sigma0, mu1, sigma1, eps)  src/losses.py:47-76 ---
 * mu1/sigma1 nullable = the N(0,1) prior of src/components/pulpo.py:337-339.  n = C*D0*D1*D2. */
size_t pulpo_reduce_ws_bytes(void);
/* out = weight * KL (weight carries the level weight and beta of src/losses.py:268-274,
 * src/models.py:157); gloss: device scalar, nullable = 1. */
int pulpo_kl_diag_fwd(const float *mu0, const float *sigma0, const float *mu1, const float *sigma1,
                      float eps, float weight, float *out, void *ws, size_t ws_bytes, int B,
                      long long n, pulpo_stream_t stream);
int pulpo_kl_diag_bwd(const float *gloss, const float *mu0, const float *sigma0, const float *mu1,
                      const float *sigma1, float eps, float weight, float *gmu0, float *gsigma0,
                      int B, long long n, pulpo_stream_t stream);

/* All levels of HierarchicalKLLoss (src/losses.py:262-276) with the N(0,1) prior, value AND gradients, in
 * one launch: out = weight * KL(N(mu, sigma) || N(0,1)) per level (batch mean), gmu/gsigma = its gradient.
 * `levels` is a HOST array (at most 6); ws: pulpo_kl_multi_ws_bytes(), zeroed once by the caller. */
typedef struct pulpo_kl_level {
    const float *mu, *sigma;   /* [B, n] */
    float *gmu, *gsigma;       /* [B, n] */
    float *out;                /* device scalar */
    long long n;
    float weight;              /* level weight * beta */
} pulpo_kl_level;
size_t pulpo_kl_multi_ws_bytes(void);
int pulpo_kl_n01_multi(const pulpo_kl_level *levels, int nlevels, float eps, int B, void *ws,
                       size_t ws_bytes, pulpo_stream_t stream);

/* ---- f-4: gauss_sampler (src/network_blocks.py:7-8) fused with the level's KL term (src/losses.py:47-76):
 * z = mu + sigma * (var * noise) and out = weight * KL[N(mu, sigma) || N(mu1, sigma1)] (batch mean; mu1 / sigma1
 * nullable = N(0,1)) from one read of mu and sigma.  `noise` is the caller's N(0,1) draw (torch.randn), so the
 * samples are the reference's for the same generator state.  ws: pulpo_reduce_ws_bytes(), zeroed once. */
int pulpo_gauss_sample_kl_fwd(const float *mu, const float *sigma, const float *noise, const float *mu1,
                              const float *sigma1, float var, float eps, float weight, float *z,
                              float *out, void *ws, size_t ws_bytes, int B, long long n,
                              pulpo_stream_t stream);
/* gmu = gz + gloss * dKL/dmu, gsigma = gz * var * noise + gloss * dKL/dsigma.  gz nullable (= 0);
 * gloss: device scalar, nullable = 1; have_kl == 0 drops the KL part (the loss was not used). */
int pulpo_gauss_sample_kl_bwd(const float *gz, const float *gloss, int have_kl, const float *mu,
                              const float *sigma, const float *noise, const float *mu1,
                              const float *sigma1, float var, float eps, float weight, float *gmu,
                              float *gsigma, int B, long long n, pulpo_stream_t stream);

/* ---- f-1: L2_reg(deformation_field, lamb)   src/losses.py:208-222 (3-D branch) ------------- */
int pulpo_l2reg_fwd(const float *f, float lamb, float *out, void *ws, size_t ws_bytes,
                    int B, int C, int D0, int D1, int D2, pulpo_stream_t stream);
/* accumulate != 0: gf += ... (lets the caller fold this gradient into the warp's gdf) */
int pulpo_l2reg_bwd(const float *gloss, const float *f, float lamb, float *gf, int accumulate,
                    int B, int C, int D0, int D1, int D2, pulpo_stream_t stream);

/* value and gradient in one pass (the plan's form: weights folded, no upstream scalar): out = L2_reg(f, lamb),
 * gf (+)= d out / d f  [+ gout * dpos].  gout [B,1,D0,D1,D2] and dpos [B,C,D0,D1,D2] (both or neither): the warp's
 * gather-half backward (pulpo_warp3d_fwd_dpos) rides in the same pass, so the field gradient is written once.
 * ws: pulpo_reduce_ws_bytes() bytes, zeroed once by the caller. */
int pulpo_l2reg_fwd_bwd(const float *f, float lamb, float *out, const float *gout, const float *dpos,
                        float *gf, int accumulate, void *ws, size_t ws_bytes, int B, int C, int D0, int D1,
                        int D2, pulpo_stream_t stream);

/* L2_reg of the x2 up-sampled field, evaluated on the coarse grid: out = L2_reg(ResizeTransform(1/2)(v), lamb) with
 * ResizeTransform(1/2)(v) = trilinear_up_2(2 v) (src/network_blocks.py:138-150; the hot path regularises exactly this
 * field at level 0, src/components/pulpo.py:314 + src/models.py:162), gv (+)= d out / d v.  v, gv: [B,C,d0,d1,d2].
 * Closed form (see csrc/losses.cu): the fine forward differences are fixed combinations of the coarse ones, so neither
 * the fine field nor its gradient is touched; three 1-D passes over coarse arrays.  scratch: caller-owned,
 * pulpo_l2reg_up2_scratch_bytes() bytes (no initialisation needed); ws: pulpo_reduce_ws_bytes() bytes, zeroed once. */
size_t pulpo_l2reg_up2_scratch_bytes(int B, int C, int d0, int d1, int d2);
int pulpo_l2reg_up2_fwd_bwd(const float *v, float lamb, float *out, float *gv, int accumulate, void *scratch,
                            size_t scratch_bytes, void *ws, size_t ws_bytes, int B, int C, int d0, int d1, int d2,
                            pulpo_stream_t stream);

/* ---- f-2: jacobian_det(deformation_field, normalize) / JDetStd   src/losses.py:147-204 (3-D branch) ----
 * det: [B,D0,D1,D2] (the reference returns jacobian[:,0,0]-shaped maps).  Replication-padded central
 * differences of the channel-flipped field scaled as the reference scales it (see csrc/jacdet.cu). */
int pulpo_jacdet_fwd(const float *df, float *det, int normalize, int B, int D0, int D1, int D2,
                     pulpo_stream_t stream);
/* gdf = (d det / d df)^T gdet; ws: 9 upstream-weighted cofactor planes. */
size_t pulpo_jacdet_bwd_ws_bytes(int B, int D0, int D1, int D2);
int pulpo_jacdet_bwd(const float *gdet, const float *df, float *gdf, void *ws, size_t ws_bytes,
                     int normalize, int B, int D0, int D1, int D2, pulpo_stream_t stream);
/* out = lamb * x.std() (unbiased, over all n elements; JDetStd, src/losses.py:202-204).  ws
 * (pulpo_std_ws_bytes, zeroed once by the caller) keeps mean and std for the backward:
 * gx = gloss * lamb * (x - mean) / ((n-1) * std);  gloss: device scalar, nullable = 1. */
size_t pulpo_std_ws_bytes(void);
int pulpo_std_fwd(const float *x, float lamb, float *out, void *ws, size_t ws_bytes, long long n,
                  pulpo_stream_t stream);
int pulpo_std_bwd(const float *gloss, const float *x, const void *ws, float lamb, float *gx,
                  long long n, pulpo_stream_t stream);

/* ---- f-3: per-voxel MC moments   evaluate.py:243-251 (std over samples) -------------------
 * Streaming Welford update of (mean, M2) with one new sample x (count = samples so far,
 * including this one), Chan merge of two partial states, and the unbiased std. */
int pulpo_moments_update(const float *x, float *mean, float *m2, int count, long long n,
                         pulpo_stream_t stream);
int pulpo_moments_merge(float *mean_a, float *m2_a, int count_a, const float *mean_b,
                        const float *m2_b, int count_b, long long n, pulpo_stream_t stream);
int pulpo_moments_std(const float *m2, float *std_out, int count, long long n,
                      pulpo_stream_t stream);

/* ---- f-3 (rest): MSE map and global NCC(var, mse) of Evaluate.uncertainty, evaluate.py:1534-1545 ----
 * pulpo_sqerr_update: acc (+)= (x - y)^2 per voxel, streamed over the MC samples (first != 0 overwrites);
 *   the MSE map is acc / N (evaluate.py:1538: torch.mean((all_moved - y)**2, axis=0)).
 * pulpo_global_ncc: Evaluate.ncc (evaluate.py:334-353, zero_norm=True) of two maps after per-map scaling,
 *   a = scale_a * (square_a ? a_in^2 : a_in), v = scale_v * v_in (so the variance map std^2 and the mean of the
 *   squared-error sums need no pass of their own):
 *   out2[0] = sum((a - mean a) / (std a * n + 1e-15) * (v - mean v) / (std v + 1e-15)), population stds;
 *   out2[1] = mean(a) (evaluate.py:1541 var.mean()).  ws: pulpo_global_ncc_ws_bytes(), zeroed once. */
int pulpo_sqerr_update(const float *x, const float *y, float *acc, int first, long long n,
                       pulpo_stream_t stream);
size_t pulpo_global_ncc_ws_bytes(void);
int pulpo_global_ncc(const float *a, const float *v, float scale_a, float scale_v, int square_a,
                     long long n, float *out2, void *ws, size_t ws_bytes, pulpo_stream_t stream);

/* ---- PULPo.training_step's total (src/models.py:164): kl*beta + recon + reg over all levels ----
 * losses: [rows, cols] per-term x per-level scalars the loss kernels wrote (weights already folded in);
 * total (nullable) = their sum in a fixed order; running (nullable, [rows]) = the per-term sums, added to
 * what is there when accumulate != 0 (so a multi-GPU job can all-reduce the logged scalars once per K steps
 * instead of once per step) or overwritten. */
int pulpo_loss_total(const float *losses, int rows, int cols, float *total, float *running,
                     int accumulate, pulpo_stream_t stream);

/* The same statistics for ALL tracked maps of one MC sample in one launch that can sit in a CUDA graph: the number
 * of samples seen so far lives in device memory (*count_dev; this sample is number *count_dev + 1) and is bumped by
 * pulpo_counter_add afterwards, so one captured launch serves every sample of the loop at evaluate.py:227-235.
 * target / sqerr_acc nullable (both or neither): sqerr_acc (+)= (x - target)^2.  `maps` is a HOST array, <= 32. */
typedef struct pulpo_moments_map {
    const float *x;
    float *mean, *m2;
    const float *target;
    float *sqerr_acc;
    long long n;
} pulpo_moments_map;
int pulpo_moments_update_multi(const pulpo_moments_map *maps, int nmaps, const int *count_dev,
                               pulpo_stream_t stream);
int pulpo_counter_add(int *counter_dev, int value, int reset, pulpo_stream_t stream);

/* ---- a13 for the MC loop: gauss_sampler (src/network_blocks.py:7-8) of ALL levels of one deformation sample ----
 * z = mu + sigma * (var * eps), eps ~ N(0,1) from Philox4x32-10 + Box-Muller keyed by (seed, sample id) and indexed by
 * (level, element): sample i has the same noise on whatever rank draws it (the sharding contract of config 3).
 * sample id = first_id + id_stride * (*count_dev) (count_dev nullable = 0): with the running sample count on the
 * device one captured launch serves the whole loop of evaluate.py:227-235.  eps_out nullable (noise dump for tests).
 * `levels` is a HOST array, <= 8. */
typedef struct pulpo_gauss_level {
    const float *mu, *sigma;
    float *z, *eps_out;
    long long n;
} pulpo_gauss_level;
int pulpo_gauss_sample_multi(const pulpo_gauss_level *levels, int nlevels, unsigned long long seed,
                             const int *count_dev, int first_id, int id_stride, float var,
                             pulpo_stream_t stream);

/* Multi-GPU reduction of the MC statistics (config 3): after the all_to_all a rank holds `nparts` partial
 * (mean, M2) slices of `chunk` elements back to back (part r at r * chunk, counts[r] samples, HOST array,
 * nparts <= 16); one pass Chan-merges them in part order and writes the unbiased std of the union. */
int pulpo_moments_merge_std(const float *mean_parts, const float *m2_parts, const int *counts,
                            int nparts, long long chunk, float *std_out, pulpo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PULPO_B200_H */
