/*
 * nerfb200.h -- C ABI of the B200-native NeRF volume-rendering hot path.
 *
 * One shared library (libnerfb200.so, built from nerf_rep_for_test_b200/csrc by
 * nvcc for sm_100a) exports these entry points.  They replace, kernel by kernel,
 * the PyTorch ops that the reference's Renderer.render(batch) issues
 * (/root/reference/src/models/nerf/renderer/volume_renderer.py:109-216); the
 * reference has no C interface for this path (its only native interface is the
 * never-loaded pybind11 module cuda/pybind.cu:11-39), so this header IS the
 * binding a maintainer would add -- see INTEGRATION.md for the ctypes stub.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *  - every pointer is a DEVICE pointer unless the name ends in _host.
 *  - the caller allocates every buffer (outputs, packed weights, workspace);
 *    the library allocates no device memory, frees nothing and keeps no user
 *    pointer after the call returns.  Its only process-wide state: the
 *    thread-local last-error string, the launch counter, the optional profiling
 *    events (nerfb200_profile_*: a diagnostic, one host thread), one-time
 *    PER-DEVICE kernel attributes (a process may drive several devices), and the
 *    per-device copy stream + event that nerfb200_render_image_host creates on
 *    first use (concurrent calls on one device are serialised by a mutex).
 *    Every call works on the CURRENT device: the caller selects the device that
 *    owns the buffers (the Python host side does, around every call).
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*), no
 *    host synchronisation inside unless stated.
 *  - return 0 on success, non-zero on failure with a message available from
 *    nerfb200_get_last_error_string(); the library never calls exit()
 *    (contrast cuda/utils.cu:9-20 of the reference).
 *  - rows of the MLP are sample points: row m = ray * n_samples + sample.
 */
#ifndef NERFB200_H_
#define NERFB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERFB200_ABI_VERSION 4

#if defined(__GNUC__)
#define NERFB200_API __attribute__((visibility("default")))
#else
#define NERFB200_API
#endif

/* arithmetic modes of the MLP query (SURVEY 8a3 / BASELINE north_star tolerances) */
#define NERFB200_MODE_FP32 0 /* CUDA-core FFMA, full-range sinf/cosf: 1e-5-relative parity mode */
#define NERFB200_MODE_BF16 1 /* tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM): performance mode */
/* fp32-accurate tensor-core mode: every operand split into two fp16 numbers (22 significand bits), three
 * tcgen05.mma kind::f16 per K step, fp32 accumulate in TMEM, full-range sincosf -- meets the same 1e-5 gates as
 * NERFB200_MODE_FP32 (requires |activation| < 65504).  feature_linear (no activation) is folded into views_linears.0
 * in fp32 at pack time. */
#define NERFB200_MODE_FP32_TC 2
/* single-pass tcgen05.mma kind::f16 with FP16 operands (11 significand bits instead of bf16's 8, saturating at
 * 65504): the kernel, the speed and the packed size of NERFB200_MODE_BF16, roughly a tenth of its error on networks
 * whose activations stay inside the fp16 range.  Inference only (the training path keeps bf16 operands). */
#define NERFB200_MODE_FP16 3
/* nerfb200_render_params.mode only: bits 0-7 = mode of the FINE pass; bits 8-15, when non-zero, = 1 + mode of the
 * COARSE pass (each model packed for its own mode).  NERFB200_MODE_BF16 | NERFB200_MODE_COARSE(NERFB200_MODE_FP32_TC)
 * puts the importance samples where the fp32 reference puts them and spends bf16 on the 75 % of the rows that only
 * feed the final colour; the coarse compositor then runs without NERFB200_COMPOSITE_FAST_MATH. */
#define NERFB200_MODE_COARSE(m) (((m) + 1) << 8)

/* compositing variants */
#define NERFB200_COMPOSITE_PLAIN 0     /* _raw2outputs, volume_renderer.py:286-357 (T uses 1-alpha+1e-10) */
#define NERFB200_COMPOSITE_ERT 1       /* _raw2outputs_with_ert intended semantics: zero weights from first T<thr */
#define NERFB200_COMPOSITE_ERT_COMPAT 2 /* ... literal :1115-1123 incl. the chunk-wide argmax quirk (2048-ray chunks) */
/* OR-ed into a variant: fast exp / sigmoid (MUFU) and fp32 prefix product / sums instead of the fp64-exact
 * arithmetic that reproduces the CPU reference bit for bit; differences <= 1e-5 of the map's scale (measured
 * 1.2e-6).  The bf16 mode uses it, the fp32 parity mode does not.  With ERT_COMPAT it takes effect in the
 * whole-pass entries only (nerfb200_composite_forward's literal kernel is always exact). */
#define NERFB200_COMPOSITE_FAST_MATH 0x10

/* network.py:22-47 -- the 24 fp32 tensors of ONE NeRF model in nn.Linear layout
 * ([out,in] row-major weight, [out] bias); state_dict names in comments. */
typedef struct nerfb200_mlp_weights {
  const float* pts_w[8];  /* pts_linears.{0..7}.weight  (256,63) (256,256)x4 (256,319) (256,256)x2 */
  const float* pts_b[8];  /* pts_linears.{0..7}.bias    (256)                                      */
  const float* views_w;   /* views_linears.0.weight     (128,283): 256 feature cols then 27 dir-PE */
  const float* views_b;   /* views_linears.0.bias       (128)                                      */
  const float* feature_w; /* feature_linear.weight      (256,256)                                  */
  const float* feature_b; /* feature_linear.bias        (256)                                      */
  const float* alpha_w;   /* alpha_linear.weight        (1,256)                                    */
  const float* alpha_b;   /* alpha_linear.bias          (1)                                        */
  const float* rgb_w;     /* rgb_linear.weight          (3,128)                                    */
  const float* rgb_b;     /* rgb_linear.bias            (3)                                        */
} nerfb200_mlp_weights;

/* gradients of the same 24 tensors (same shapes), OVERWRITTEN by nerfb200_mlp_backward */
typedef struct nerfb200_mlp_grads {
  float* pts_w[8];
  float* pts_b[8];
  float* views_w;
  float* views_b;
  float* feature_w;
  float* feature_b;
  float* alpha_w;
  float* alpha_b;
  float* rgb_w;
  float* rgb_b;
} nerfb200_mlp_grads;

/* ---- library ---------------------------------------------------------------------------- */
NERFB200_API int nerfb200_abi_version(void);
NERFB200_API const char* nerfb200_get_last_error_string(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
NERFB200_API uint64_t nerfb200_launch_count(void);

/* Optional timing of the MLP kernel for the roofline figure: while enabled, every mlp_forward launch
 * (also those inside render_rays / render_image_host) is bracketed by CUDA events on its stream.
 * profile_enable(on) resets the counters; profile_read synchronises on the recorded events and
 * returns the summed kernel time, the number of launches and the MLP rows they processed. */
NERFB200_API int nerfb200_profile_enable(int on);
NERFB200_API int nerfb200_profile_read(double* mlp_ms, uint64_t* mlp_launches, double* mlp_rows);

/* ---- a1: ray generation (volume_renderer.py:115-147) ------------------------------------- */
/* pose: [4,4] c2w row-major, intrinsics: [3,3]; rays_o/rays_d: [H*W,3], ray id = y*W+x,
 * rays_d normalised (:140). */
NERFB200_API int nerfb200_raygen(const float* pose, const float* intrinsics, int H, int W,
                    float* rays_o, float* rays_d, void* stream);

/* ---- a2: stratified coarse sampling (volume_renderer.py:218-237) ------------------------- */
/* z_table: [n_samples] = near*(1-t)+far*t evaluated by the caller with torch.linspace so the
 * table is bit-identical to the reference's.  perturb==0: broadcast.  perturb!=0: stratified
 * jitter (:228-235) with a counter-based RNG keyed on (seed, ray, sample). */
NERFB200_API int nerfb200_sample_coarse(const float* z_table, int n_rays, int n_samples, int perturb,
                           uint64_t seed, float* z_vals, void* stream);

/* ---- a3: positional encoding fused into the NeRF MLP (freq.py:3-32, network.py:49-74) ---- */
NERFB200_API size_t nerfb200_packed_weights_bytes(int mode);
/* repack one model's weights into the layout the `mode` kernel streams; re-run after every
 * optimizer step.  packed must be 1024-byte aligned. */
NERFB200_API int nerfb200_pack_weights(const nerfb200_mlp_weights* w, int mode, void* packed, void* stream);
/* raw[m] = (rgb_raw[3], sigma_raw) for row m = ray*n_samples+s at point o + d*z (evaluated as
 * fadd(o, fmul(d,z)), no FMA contraction, :165); view dir = rays_d (:143). */
NERFB200_API int nerfb200_mlp_forward(const void* packed, int mode, const float* rays_o, const float* rays_d,
                         const float* z_vals, int n_rays, int n_samples, float* raw, void* stream);

/* ---- a7: training twin of mlp_forward and the MLP backward (autograd through network.py:49-74) ---- */
/* Buffer sizes for n_rows = n_rays*n_samples MLP rows.  acts / dacts are "tile images": per 128-row tile a
 * sequence of 16 KB blocks, each the [128 rows][64 bf16] cut of one stage output in the shared-memory
 * layout of tcgen05.mma (SWIZZLE_128B); layout in csrc/train_layout.cuh.  masks = relu sign bits. */
NERFB200_API size_t nerfb200_train_acts_bytes(long long n_rows);
NERFB200_API size_t nerfb200_train_masks_bytes(long long n_rows);
NERFB200_API size_t nerfb200_mlp_backward_workspace_bytes(long long n_rows);
/* BF16 mode forward (fused tail, like inference) that additionally keeps, for the backward pass, the bf16
 * inputs/outputs of its nine stages (PE, dir-PE, relu(h0..h7), relu(views)) in `acts` (128-byte aligned) and the
 * relu sign bits in `masks` (16-byte aligned). */
NERFB200_API int nerfb200_mlp_forward_train(const void* packed, int mode, const float* rays_o,
                               const float* rays_d, const float* z_vals, int n_rays, int n_samples,
                               float* raw, void* acts, void* masks, void* stream);
/* W^T image for the backward dgrad chain; re-run after every optimizer step (1024-byte aligned). */
NERFB200_API size_t nerfb200_packed_bwd_bytes(void);
NERFB200_API int nerfb200_pack_weights_bwd(const nerfb200_mlp_weights* w, void* packed_bwd, void* stream);
/* Same; packed_fwd (may be NULL): the image nerfb200_pack_weights(NERFB200_MODE_BF16) produced from the SAME weights
 * earlier on the same stream -- the fused-tail product views_linears.0[:, :256] x feature_linear it carries is reused
 * instead of being recomputed (the training step re-packs both images after every optimizer step). */
NERFB200_API int nerfb200_pack_weights_bwd2(const nerfb200_mlp_weights* w, const void* packed_fwd, void* packed_bwd,
                               void* stream);
/* Gradients of the 24 tensors of one model given g_raw = dL/d raw [n_rows,4] (fp32) and the acts/masks of
 * nerfb200_mlp_forward_train on the same rows.  Two tcgen05 kernels: the activation-gradient chain
 * (keeps every dL/d pre-activation in the workspace) and nine split-K weight-gradient GEMMs with the bias
 * and head gradients riding along; bf16 operands, fp32 accumulation, fp32 results.  The tail is fused like
 * the forward (views_linears.0 o feature_linear = one linear map); `weights` (the same fp32 tensors that
 * were packed) is read to split its gradient back into views_linears.0 / feature_linear by the chain rule.
 * MLP inputs receive no gradient (the hierarchical sampler is detached).  workspace: 128-byte aligned. */
NERFB200_API int nerfb200_mlp_backward(const void* packed_bwd, const nerfb200_mlp_weights* weights, const float* g_raw,
                          const void* acts, const void* masks, long long n_rows, void* workspace,
                          size_t workspace_bytes, const nerfb200_mlp_grads* grads, void* stream);

/* fp32-ACCURATE training twin (true fp32 FFMA arithmetic, layer by layer, csrc/mlp_fp32_train.cu): what the reference
 * computes under autograd (network.py:49-74, trainer.py:56-60).  Reads the fp32 nn.Linear tensors directly (no packed
 * image).  acts: nerfb200_train_fp32_acts_bytes(n_rows) bytes, 16-byte aligned -- row-major fp32 planes
 * PE[.,64] | dirPE[.,32] | relu(h0..h7)[.,256] x 8 | feature[.,256] | relu(views)[.,128]. */
NERFB200_API size_t nerfb200_train_fp32_acts_bytes(long long n_rows);
NERFB200_API size_t nerfb200_train_fp32_workspace_bytes(long long n_rows);
NERFB200_API int nerfb200_mlp_forward_train_fp32(const nerfb200_mlp_weights* weights, const float* rays_o,
                                    const float* rays_d, const float* z_vals, int n_rays, int n_samples,
                                    float* raw, void* acts, void* stream);
/* Gradients of the 24 tensors (OVERWRITTEN) given g_raw = dL/d raw [n_rows,4] and the acts of the call above.
 * g_z (may be NULL): dL/d z_vals [n_rays,n_samples] through the MLP input x = o + d z and its positional encoding
 * (freq.py:23-26) -- the path by which the reference's fine loss reaches the coarse network (its sampler is not
 * detached, volume_renderer.py:181-183); the caller adds the compositor's own dL/dz (nerfb200_composite_backward_z). */
NERFB200_API int nerfb200_mlp_backward_fp32(const nerfb200_mlp_weights* weights, const float* g_raw, const void* acts,
                               const float* rays_d, int n_rays, int n_samples, void* workspace,
                               size_t workspace_bytes, const nerfb200_mlp_grads* grads, float* g_z, void* stream);

/* bf16 training path, reference graph (sampler not detached, volume_renderer.py:181-183): dL/d z_vals [n_rays,n_samples]
 * through the MLP input x = o + d z, computed from the activation-gradient planes nerfb200_mlp_backward left in
 * `workspace` (call it after nerfb200_mlp_backward on the same rows, with the same workspace) and the fp32
 * pts_linears.0 / pts_linears.5 weights: g_pe = dpre0 W0 + dpre5 W5[:, :63] on the tensor cores (bf16 operands,
 * fp32 accumulate), then the positional-encoding backward (freq.py:23-26) in fp32. */
NERFB200_API int nerfb200_mlp_backward_input(const nerfb200_mlp_weights* weights, const void* workspace,
                                const float* rays_o, const float* rays_d, const float* z_vals, int n_rays,
                                int n_samples, float* g_z, void* stream);

/* The small end of the training step, one kernel each instead of torch's ten / its 19-block multi-tensor launch.
 * mse_pair_grad: loss of trainers/nerf.py:52-65, loss[0] = mse(rgb_map_0, t) + mse(rgb_map, t) (device float, overwritten)
 * and its gradients g_rgb0 = 2 (rgb_map_0 - t) / (3 n_rays), g_rgb likewise (all [n_rays,3]).
 * adam_clip_step: clip_grad_value_(clip_value) (trainer.py:59; 0 = no clipping) + torch.optim.Adam (optimizer.py:8-28:
 * amsgrad off, weight_decay 0) over flat fp32 buffers of n elements, in place; `step` counts from 1; grads are first
 * multiplied by grad_scale (gradient averaging) and are written back clamped.  The hyper-parameters are doubles: torch forms
 * 1 - beta and the bias corrections in double before rounding to fp32, and so does this entry. */
NERFB200_API int nerfb200_mse_pair_grad(const float* rgb0, const float* rgb, const float* target, int n_rays,
                           float* g_rgb0, float* g_rgb, float* loss, void* stream);
NERFB200_API int nerfb200_adam_clip_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                            double lr, double beta1, double beta2, double eps, long long step, float clip_value,
                            float grad_scale, void* stream);

/* diagnostic twin of mlp_forward (BF16 and FP32_TC modes): additionally writes the fp32 post-activation output
 * of each of the ten stages (mlp_layout.cuh) for rows 0..127 into stage_dump [10][128][256];
 * used by the stage-level parity tests.  FP32_TC folds feature_linear into views_linears.0: its plane 8 is left
 * untouched and plane 9 holds relu(views). */
NERFB200_API int nerfb200_mlp_forward_stages(const void* packed, int mode, const float* rays_o,
                                const float* rays_d, const float* z_vals, int n_rays, int n_samples,
                                float* raw, float* stage_dump, void* stream);

/* ---- a5/a6: alpha compositing (volume_renderer.py:286-357, :1089-1157) ------------------- */
/* raw [n_rays,S,4], z_vals [n_rays,S], rays_d [n_rays,3] -> rgb_map [n_rays,3], disp/acc/depth
 * [n_rays], weights [n_rays,S] (may be NULL).  One warp per ray, shuffle prefix product.
 * variant: NERFB200_COMPOSITE_* (PLAIN / ERT optionally OR-ed with NERFB200_COMPOSITE_FAST_MATH); ERT_COMPAT groups
 * rays in chunks of compat_chunk (2048). */
NERFB200_API int nerfb200_composite_forward(const float* raw, const float* z_vals, const float* rays_d,
                               int n_rays, int n_samples, int variant, float ert_threshold,
                               int white_bkgd, int compat_chunk, float* rgb_map, float* disp_map,
                               float* acc_map, float* depth_map, float* weights, void* stream);

/* Same, for the sparse empty-space-skipping launch: keep_bits (may be NULL = every row) holds one bit per row
 * m = ray*n_samples+s, as written by nerfb200_ess_compact.  A cleared bit means zero density and the row's raw entry
 * is NOT read, so raw needs no zero fill.  PLAIN / ERT variants only. */
NERFB200_API int nerfb200_composite_forward_masked(const float* raw, const float* z_vals, const float* rays_d,
                                      const uint32_t* keep_bits, int n_rays, int n_samples, int variant,
                                      float ert_threshold, int white_bkgd, int compat_chunk, float* rgb_map,
                                      float* disp_map, float* acc_map, float* depth_map, float* weights, void* stream);
/* raw_noise_std (volume_renderer.py:310-314, :1099-1103): raw[m,3] += N(0,1) * std for m < n_rows, in place, before
 * compositing (forward and backward then see the same noisy density).  Counter-based generator keyed on
 * (seed, m); the reference's torch.randn stream is not reproduced, only its distribution. */
NERFB200_API int nerfb200_sigma_noise(float* raw, long long n_rows, float std, uint64_t seed, void* stream);

/* a7: analytic backward of NERFB200_COMPOSITE_PLAIN.  g_* are dL/d(map) (any may be NULL);
 * writes g_raw [n_rays,S,4] (overwrite). z_vals/rays_d receive no gradient here. */
NERFB200_API int nerfb200_composite_backward(const float* raw, const float* z_vals, const float* rays_d,
                                int n_rays, int n_samples, int white_bkgd, const float* g_rgb_map,
                                const float* g_acc_map, const float* g_depth_map,
                                const float* g_weights, float* g_raw, void* stream);

/* Same, additionally g_z [n_rays,S] (may be NULL): dL/d z_vals through the interval lengths
 * dists = z[i+1]-z[i] (:295-297) and through depth_map = sum w z (:339). */
NERFB200_API int nerfb200_composite_backward_z(const float* raw, const float* z_vals, const float* rays_d,
                                  int n_rays, int n_samples, int white_bkgd, const float* g_rgb_map,
                                  const float* g_acc_map, const float* g_depth_map,
                                  const float* g_weights, float* g_raw, float* g_z, void* stream);

/* ---- a4: sample_pdf + merge (volume_renderer.py:239-268, :181-183) ----------------------- */
/* kernel-level inverse-CDF lookup: cdf,bins [n_rays,n_bins]; u is [n_u] (u_per_ray==0, the
 * eval-mode linspace table) or [n_rays,n_u]; inds = searchsorted(cdf,u,right=True) (int32). */
NERFB200_API int nerfb200_sample_from_cdf(const float* cdf, const float* bins, const float* u, int u_per_ray,
                             int n_rays, int n_bins, int n_u, float* samples, int32_t* inds,
                             void* stream);
/* whole a4: weights [n_rays,S] (coarse, the kernel uses [1:-1]), z_coarse [n_rays,S] ->
 * z_all [n_rays,S+n_u] sorted; optional z_samples [n_rays,n_u], inds [n_rays,n_u], cdf
 * [n_rays,S-1] (NULL to skip). */
NERFB200_API int nerfb200_sample_pdf_merge(const float* z_coarse, const float* weights, const float* u,
                              int u_per_ray, int n_rays, int n_samples, int n_u, float* z_all,
                              float* z_samples, int32_t* inds, float* cdf, void* stream);

/* a7 through a4: the reference does not detach the importance samples (:181-183), so dL/d z_all reaches the coarse
 * weights.  Given g_z_all [n_rays,S+n_u] (gradient of the merged, sorted depths) writes g_weights [n_rays,S] (the
 * gradient of the coarse weights; entries 0 and S-1 are 0, :181 uses weights[...,1:-1]).  The forward is recomputed
 * inside (same arithmetic as nerfb200_sample_pdf_merge); z_coarse, weights, u as passed to the forward. */
NERFB200_API int nerfb200_sample_pdf_backward(const float* z_coarse, const float* weights, const float* u,
                                 int u_per_ray, int n_rays, int n_samples, int n_u, const float* g_z_all,
                                 float* g_weights, void* stream);

/* ---- a8: occupancy grid / empty-space skipping (volume_renderer.py:830-873, :963-1087) --- */
/* grid: uint8 [R,R,R] (1 = occupied), bbox [-2,2]^3.  Per ray: if more than half of the
 * coarse samples fall in empty cells, keep the occupied z's and refill with
 * linspace(min_occ,max_occ,S-n_keep), sorted (intended per-ray semantics of :1037-1077). */
NERFB200_API int nerfb200_ess_resample(const uint8_t* grid, int res, const float* rays_o, const float* rays_d,
                          int n_rays, int n_samples, float* z_vals, int32_t* n_empty, void* stream);
/* The reference's literal ESS (:1009-1077 with the stride-0 aliasing of :1020,1077): within every block of `chunk`
 * consecutive rays (the reference calls the sampler once per 2048-ray chunk, :147) each highly-empty ray, in order,
 * rewrites the ONE shared row -- keep the entries its own occupancy mask (taken at the z_table depths) marks occupied,
 * refill with linspace(min, max, n_add), sort -- and all rays of the block get the final row.  z_table [n_samples]
 * (the unperturbed depths), z_vals [n_rays,n_samples] out, scratch: n_rays uint64 (8-byte aligned).  n_samples <= 64. */
NERFB200_API int nerfb200_ess_resample_compat(const uint8_t* grid, int res, const float* rays_o, const float* rays_d,
                                 int n_rays, int n_samples, int chunk, const float* z_table, float* z_vals,
                                 uint64_t* scratch, void* stream);
/* :1079-1085: stratified jitter of already placed per-ray depths, from each row's own mid-points (the reference
 * jitters AFTER the ESS resampling); in place, counter-based RNG keyed on (seed, ray, sample). */
NERFB200_API int nerfb200_jitter_rows(float* z_vals, int n_rays, int n_samples, uint64_t seed, void* stream);
/* :963-985 as called from :1147-1155: mark cells of samples with weight>1e-4 and
 * relu(sigma_raw)>0.01; points are rays_d*z (ray origin omitted, as in the reference) unless
 * use_origin!=0. */
NERFB200_API int nerfb200_ess_update(uint8_t* grid, int res, const float* rays_o, const float* rays_d,
                        const float* z_vals, const float* raw, const float* weights, int n_rays,
                        int n_samples, int use_origin, void* stream);

/* Empty-space skipping proper (BASELINE.json configs[4]): list the rows m = ray*n_samples+s whose sample
 * lies in an occupied cell (and, when z_term != NULL, has z <= z_term[ray]) into row_ids (order
 * unspecified) and write their number to *n_active (device int32).  keep_bits (may be NULL): uint32
 * [ceil(n_rays*n_samples/32)], bit (m & 31) of word (m >> 5) = row m is listed. */
NERFB200_API int nerfb200_ess_compact(const uint8_t* grid, int res, const float* rays_o, const float* rays_d,
                         const float* z_vals, const float* z_term, int n_rays, int n_samples,
                         int32_t* row_ids, int32_t* n_active, uint32_t* keep_bits, void* stream);
/* MLP on the listed rows only (BF16 mode); every other row of raw is set to 0 (zero density).  The
 * row count is read on the device: no host synchronisation. */
NERFB200_API int nerfb200_mlp_forward_sparse(const void* packed, int mode, const float* rays_o,
                                const float* rays_d, const float* z_vals, int n_rays, int n_samples,
                                const int32_t* row_ids, const int32_t* n_active, float* raw, void* stream);
/* ray_active[r] = 1 when the segment o + d*z, z in [z_table[0], z_table[n_samples-1]] (widened by 1e-3), meets
 * the axis-aligned box [box_lo, box_hi] (3 floats each, HOST memory, +-INFINITY allowed), else 0. */
NERFB200_API int nerfb200_ray_cull(const float* rays_o, const float* rays_d, int n_rays, const float* z_table,
                      int n_samples, const float* box_lo, const float* box_hi, uint8_t* ray_active, void* stream);
/* depth at which the transmittance implied by `weights` (T_i = 1 - sum_{j<i} w_j) first drops below
 * thr; +inf when it never does.  Used to cut the fine pass behind opaque surfaces (ERT). */
NERFB200_API int nerfb200_ert_depth(const float* weights, const float* z_vals, int n_rays, int n_samples,
                       float thr, float* z_term, void* stream);
/* totals[0..1] += counts[0..1] (device side; statistics of the sparse passes) */
NERFB200_API int nerfb200_accumulate_counts(const int32_t* counts, int64_t* totals, void* stream);

/* ---- a9: KiloNeRF-style path (BASELINE configs[4] ii): occupancy-grid march, 32-wide micro-MLPs, -------
 * integration with early ray termination.  Reference kernels (never built / run by the reference):
 * cuda/generate_inputs.cu:11-35, :60-126, cuda/network_eval.cu:24-254, cuda/integrate.cu:9-57, :84-97. */
typedef struct nerfb200_kilo_camera {
  int H, W;
  float cx, cy, fx, fy;
  float c2w[9];    /* rotation part of the pose, row-major (generate_inputs.cu: c_c2w) */
  float origin[3]; /* ray origin (generate_inputs.cu: c_origin) */
} nerfb200_kilo_camera;
typedef struct nerfb200_kilo_grid {
  int res[3];      /* occupancy grid int16 [res0][res1][res2]: network id, -1 = empty space */
  float gmin[3], gmax[3]; /* global domain */
} nerfb200_kilo_grid;
typedef struct nerfb200_kilo_march_params {
  float distance_between_points; /* in units of the UNNORMALISED ray direction */
  int max_samples_per_ray;       /* compacted queries per ray and pass */
  int max_depth_index;
  float min_distance;
  float transmittance_threshold; /* early ray termination */
  int white_bkgd;
  int max_passes;                /* fixed pass count of kilo_render; passes with nothing left to do return at once */
} nerfb200_kilo_march_params;
/* floats per micro-MLP (network_eval.cu:48-52): 6212, per layer [bias(out) | W(in-major, out fastest)] */
NERFB200_API int nerfb200_kilo_param_size(void);
/* generate_inputs.cu:11-35: dirs [H*W,3], NOT normalised */
NERFB200_API int nerfb200_kilo_rays_d(const nerfb200_kilo_camera* cam, float* dirs, void* stream);
/* generate_inputs.cu:60-126: one marching pass.  query_indices [n_rays,S] (= ray*max_depth+depth), assigned
 * [n_rays,S] (-1 = unfilled; query is then -1 too), active_ray_mask / depth_indices [n_rays] are the
 * resumable state (read unless is_initial_query, always written). */
NERFB200_API int nerfb200_kilo_march(const nerfb200_kilo_grid* g, const float* origin, const float* dirs, const int16_t* grid,
                        int n_rays, float distance_between_points, int max_samples_per_ray, int max_depth_index,
                        float min_distance, int is_initial_query, int32_t* query_indices, int16_t* assigned_networks,
                        uint8_t* active_ray_mask, int32_t* depth_indices, void* stream);
NERFB200_API size_t nerfb200_kilo_workspace_bytes(int n_rays, int max_samples_per_ray, int num_networks);
/* network_eval.cu:24-254 on every filled slot: groups the queries by network on the device (counting sort,
 * replaces cuda/reorder.cu), evaluates, and writes (sigmoid rgb, relu sigma) back to the slot; unfilled slots
 * get 0.  params [num_networks,6212], domain_mins/maxs [num_networks,3].  Arithmetic: fp32-accurate on tensor
 * cores (every operand carried as two fp16 numbers, three MMAs per product, fp32 accumulate; a network's weights
 * are pre-scaled by a power of two, so their magnitude is free) -- hidden activations must stay below 65504 in
 * magnitude (saturating conversion beyond). */
NERFB200_API int nerfb200_kilo_network_eval(const nerfb200_kilo_camera* cam, const nerfb200_kilo_march_params* mp,
                               const int32_t* query_indices, const int16_t* assigned_networks, int n_rays,
                               const float* params, const float* domain_mins, const float* domain_maxs, int num_networks,
                               void* workspace, size_t workspace_bytes, float* rgb_sigma, void* stream);
/* integrate.cu:9-57 over the filled slots of one pass (dists [n_rays] = step length per ray) */
NERFB200_API int nerfb200_kilo_integrate(const float* rgb_sigma, const int16_t* assigned_networks, const float* dists, int n_rays,
                            int samples_per_ray, float transmittance_threshold, int is_initial_query, float* rgb_map,
                            float* acc_map, float* transmittance, uint8_t* active_ray_mask, void* stream);
/* whole frame: rays, max_passes x (march, sort, eval, integrate), background; no host synchronisation.
 * stats (device int64[2], may be NULL): += evaluated samples, += passes that had work. */
NERFB200_API int nerfb200_kilo_render(const nerfb200_kilo_camera* cam, const nerfb200_kilo_grid* g, const nerfb200_kilo_march_params* mp,
                         const int16_t* occupancy_grid, const float* params, const float* domain_mins,
                         const float* domain_maxs, int num_networks, void* workspace, size_t workspace_bytes,
                         float* rgb_map, float* acc_map, int64_t* stats, void* stream);

/* ---- whole pass --------------------------------------------------------------------------- */
typedef struct nerfb200_render_params {
  int n_samples;      /* 64  */
  int n_importance;   /* 128 */
  int mode;           /* NERFB200_MODE_* */
  int variant;        /* NERFB200_COMPOSITE_* */
  int white_bkgd;
  int perturb;        /* 0 for parity */
  int u_per_ray;      /* 0: u is the [n_importance] table; 1: u is [n_rays,n_importance] */
  int compat_chunk;   /* 2048 */
  float ert_threshold;
  float raw_noise_std; /* lego.yaml:23 uses 0; > 0: N(0,1)*std added to sigma_raw before the relu (:310-314) */
  uint64_t seed;
  /* a8: when non-NULL the coarse z's of every ray are passed through nerfb200_ess_resample (unperturbed depths are
   * tested and resampled first, the stratified jitter is applied afterwards from each row's mid-points, :1079-1085) */
  const uint8_t* occupancy_grid; /* uint8 [grid_res]^3, device */
  int grid_res;
  /* ess_skip = 0: reference semantics (resample highly-empty rays, nerfb200_ess_resample).
   * ess_skip = 1: empty-space SKIPPING -- samples in empty cells (and, with an ERT variant, fine
   * samples behind the coarse termination depth) are not evaluated by the MLP at all; their density
   * is 0.  eval_counts (device int64[2], may be NULL) accumulates the evaluated coarse / fine rows. */
  int ess_skip;
  int64_t* eval_counts;
  /* ess_skip only.  cull_rays != 0: rays whose segment [z_table[0], z_table[n_samples-1]] misses the box
   * [cull_lo, cull_hi] (world units) are culled before any per-sample work (nerfb200_ray_cull).  The caller must pass
   * a box that contains every point whose grid lookup can be non-empty -- the occupied cells' bounds, +-INFINITY on a
   * side whose boundary cell is occupied (lookups clamp), plus a margin; Renderer derives it from the grid. */
  int cull_rays;
  float cull_lo[3];
  float cull_hi[3];
  /* ess_skip = 0 only.  0: per-ray resampling (the intended semantics of :1037-1077).  1: the reference's LITERAL
   * behaviour -- z_vals is a stride-0 expand()ed view (:1020), so every highly-empty ray rewrites the one row all rays
   * of a compat_chunk-ray call share (nerfb200_ess_resample_compat); needs n_samples <= 64. */
  int ess_ref_compat;
} nerfb200_render_params;

/* maps for one pass: rgb [n,3], disp/acc/depth [n] */
typedef struct nerfb200_maps {
  float* rgb;
  float* disp;
  float* acc;
  float* depth;
} nerfb200_maps;

NERFB200_API size_t nerfb200_render_workspace_bytes(int n_rays, const nerfb200_render_params* p);
/* coarse sample -> MLP -> composite -> sample_pdf+merge -> MLP(fine) -> composite, all on
 * `stream`, no host sync.  packed_coarse/packed_fine from nerfb200_pack_weights(mode).
 * maps_fine members may be NULL when n_importance==0. */
NERFB200_API int nerfb200_render_rays(const void* packed_coarse, const void* packed_fine, const float* rays_o,
                         const float* rays_d, int n_rays, const float* z_table, const float* u,
                         const nerfb200_render_params* p, void* workspace, size_t workspace_bytes,
                         const nerfb200_maps* maps_coarse, const nerfb200_maps* maps_fine,
                         void* stream);

/* Host-buffer entry (the e2e number): pose_host [16], intrinsics_host [9] and the eight output
 * maps live in (pinned) HOST memory; device scratch comes from `workspace`.  Copies in, renders
 * H*W rays, copies the maps out and synchronises the stream before returning. */
NERFB200_API size_t nerfb200_render_image_workspace_bytes(int H, int W, const nerfb200_render_params* p);
NERFB200_API int nerfb200_render_image_host(const void* packed_coarse, const void* packed_fine,
                               const float* pose_host, const float* intrinsics_host, int H, int W,
                               const float* z_table, const float* u, const nerfb200_render_params* p,
                               void* workspace, size_t workspace_bytes,
                               const nerfb200_maps* maps_coarse_host,
                               const nerfb200_maps* maps_fine_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFB200_H_ */
