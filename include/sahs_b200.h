/*
 * sahs_b200.h -- C ABI of libsahs_b200.so, the B200-native (sm_100a) per-ray render/train hot path of
 * SAHS-Deformable-Nerf.
 *
 * The reference has no FFI layer: its boundary is the Python surface of nerf-pytorch/nerf
 * (SURVEY.md section 8b).  Each entry point below names the reference function (file:line, relative to
 * /root/reference/nerf-pytorch) whose arithmetic it replaces.  The Python host package `sahs_b200`
 * binds these with ctypes and re-exposes the reference's own signatures.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless the name ends in _host;
 *   - returns 0 on success, a negative SAHS_E* code otherwise; no exception crosses the ABI;
 *     sahs_last_error() returns a thread-local message for the last failure;
 *   - allocates nothing: outputs and workspaces are caller-provided (sizes via the *_bytes queries);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and re-entrant per stream;
 *   - fp32 tensors are dense row-major unless a stride argument says otherwise.
 */
#ifndef SAHS_B200_H
#define SAHS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAHS_OK 0
#define SAHS_EINVAL (-1)      /* bad argument / unsupported configuration */
#define SAHS_ECUDA (-2)       /* a CUDA runtime call or kernel launch failed */
#define SAHS_EUNSUPPORTED (-3)

#define SAHS_RAW_CH 16        /* rgb3 | seg12 | sigma1, ref: nerf/modules.py:295 */
#define SAHS_MAP_CH 15        /* rgb3 | seg12, ref: nerf/train_utils.py:205-206 */
#define SAHS_DRIVING_DIM 76
#define SAHS_POSE_CODE_DIM 36
#define SAHS_GRID_CH 32
#define SAHS_GRID_RES 32

int sahs_abi_version(void);
/* 16-bit operand format of trunk and heads on the render path: 0 = fp16 (default build: 11-bit significand, conversions
 * saturate at +-65504), 1 = bf16 (`make BF16=1` -> lib/libsahs_b200_bf16.so).  Deformation phase and training: fp16. */
int sahs_operand_format(void);
const char* sahs_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches evidence) */
uint64_t sahs_launch_count(void);

/* ---- architecture description: the dimensions NeRFaceModel.__init__ derives from the YAML
 *      (ref: nerf/models.py:189-299, nerf/modules.py:168-252, :323-369, :401-442) ------------------- */
typedef struct sahs_model_spec {
  int32_t xyz_L, xyz_inc;            /* models.coarse.num_encoding_fn_xyz / include_input_xyz        */
  int32_t dir_L, dir_inc;            /* num_encoding_fn_dir / include_input_dir                       */
  int32_t use_ambient, amb_dim, amb_L, amb_inc;   /* models.hyper.*                                   */
  int32_t use_warp, warp_layers, warp_hidden, warp_skip;
  int32_t hyper_layers, hyper_hidden, hyper_skip;
  int32_t trunk_layers, trunk_hidden, trunk_skip; /* trunk_skip is the NeRFMLP default 3              */
  int32_t trunk_driving, trunk_pose; /* include_driving (76 cols) / use_pose (36 cols) in the trunk   */
  int32_t use_grid;                  /* use_spatial_embeddings                                        */
} sahs_model_spec;

/* Canonical order of the fp32 parameter pointers handed to sahs_pack_params / sahs_fold_frame
 * (weights are nn.Linear [out,in] row-major, ref state_dict names in SURVEY.md Appendix A):
 *   [0]                 spatial_embeddings [1,32,32,32,32]            (NULL if !use_grid)
 *   then for i in 0..warp_layers-1:  warp_field_mlp.layers_xyz.i.{weight,bias}; fc_final.{weight,bias}
 *   then for i in 0..hyper_layers-1: hyper_sheep_mlp.layers_ambient.i.{weight,bias}; fc_ambient.{weight,bias}
 *   then, for the requested level:   layers_xyz.i.{weight,bias} (trunk_layers), fc_feat, fc_alpha,
 *                                    layers_dir.0-3, fc_rgb, layers_seg.0-3, fc_seg   ({weight,bias} each)
 * sahs_param_count(spec) returns the length of that list. */
int sahs_param_count(const sahs_model_spec* spec);

/* ---- (1) ray generation, stratified depths, positional encoding --------------------------------- */
/* get_ray_bundle, ref: nerf/nerf_helpers.py:178-233.  c2w: device [3,4] row-major (ld = 4).
 * ro, rd: [H*W,3] (pixel (row j, col i) at j*W+i).  Bit-exact with the reference's fp32 op order; cx, cy are
 * doubles because the reference forms width*cx in double before rounding to fp32. */
int sahs_get_ray_bundle(int height, int width, float fx, float fy, double cx, double cy, const float* c2w,
                        float* ro, float* rd, void* stream);
/* coarse depths, ref: nerf/train_utils.py:93-113.  t_vals: device [S] = linspace(0,1,S) (host-made so it is
 * bit-identical to torch.linspace); t_rand: device [R,S] uniforms or NULL for no perturbation. */
int sahs_coarse_z(int num_rays, int num_samples, float near_, float far_, int lindisp, const float* t_vals,
                  const float* t_rand, float* z_out, void* stream);
/* positional_encoding (log sampling), ref: nerf/nerf_helpers.py:305-349.  x [n,d] -> out [n, d*(inc+2L)]. */
int sahs_positional_encoding(const float* x, int64_t n, int d, int num_freqs, int include_input, float* out,
                             void* stream);

/* ---- (2) fused deformation + hyper-sheet + grid gather + radiance MLP (tcgen05/TMEM, TMA-fed) ---- */
/* Bytes of the packed 16-bit (fp16 operands) weight image of one level (warp+hyper+trunk+heads, UMMA 128B-swizzled stage
 * images in consumption order) and of the per-frame constant block (fp32 biases with the frame-constant
 * driving/pose columns folded in + the small fp32 head weights). */
int sahs_field_sizes(const sahs_model_spec* spec, size_t* packed_bytes, size_t* frame_const_bytes,
                     size_t* grid_bytes);
/* fp32 state_dict -> packed fp16 image for `level` (0 coarse, 1 fine); also re-lays the embedding grid
 * channel-last (grid_out may be NULL when !use_grid).  Replaces nothing in the reference (new layout step);
 * must be re-run after every optimizer step. */
int sahs_pack_params(const sahs_model_spec* spec, int level, const float* const* params_host_array,
                     void* packed_out, float* grid_out, void* stream);
/* Per-frame constant folding: driving [76] and pose_code [36] are identical for every point of a frame
 * (ref: nerf/models.py:518-521), so W[:, const cols] @ (driving|pose) becomes a bias. */
int sahs_fold_frame(const sahs_model_spec* spec, int level, const float* const* params_host_array,
                    const float* driving, const float* pose_code, float* frame_const_out, void* stream);
/* raw[R,S,16] = model(level, ro + rd*z, rd, driving, pose), ref: nerf/train_utils.py:9-50 (run_network),
 * nerf/models.py:514-528 / :367-380 (forward), :301-365, nerf/modules.py:254-295, :371-390, :444-462.
 * debug: optional device buffer [128*256] fp32 + debug_pass id (see SAHS_DBG_*), else NULL/-1. */
int sahs_field_fwd(const sahs_model_spec* spec, int level, const void* packed, const float* frame_const,
                   const float* grid_cl, const float* ro, const float* rd, const float* z, int num_rays,
                   int num_samples, float* raw_out, float* debug, int debug_pass, void* stream);

/* ---- (2b) training: forward with tapes, activation-gradient chain -------------------------------- */
/* Layout of the training tapes for this spec (40 ints: tx_e0, e0_k, tx_wh, whh, w_layers, tx_e1, e1_k, tx_th, th,
 * t_layers, tx_feat, tx_xtra, tx_hh, tx_total, td_wh, td_final, td_th, td_feat, td_hh, td_out, td_total,
 * n_mask_layers, e0_dim, e1_dim, xtra_dim, wh, hh, w_skip, t_skip, ct_off, ct_len, use_w, hd, packed_t_bytes,
 * bwd_stages, fc_total, packed_train_bytes, 0, 0, 0).
 * Both tapes are tile-major chunk images: tape[ceil(P/128)][total/64][128 points x 64 columns, fp16, 128B-swizzled], i.e.
 * ceil(P/128)*128 x total fp16 values (tx_total / td_total columns, every item starts at a multiple of 64);
 * masks are [n_mask_layers][P][2] x 16 bytes; saves [P,8]. */
int sahs_train_layout(const sahs_model_spec* spec, int32_t* out, int max_out);
/* Packed forward image used by training (always the merged fp16 deformation phase) and the transposed fp16 image
 * consumed by sahs_field_bwd. */
int sahs_pack_params_train(const sahs_model_spec* spec, int level, const float* const* params_host_array,
                           void* packed_out, void* stream);
int sahs_pack_params_bwd(const sahs_model_spec* spec, int level, const float* const* params_host_array,
                         void* packed_t_out, void* stream);
/* sahs_field_fwd that also records the activation tape, the activation sign masks and the warped point. */
int sahs_field_fwd_train(const sahs_model_spec* spec, int level, const void* packed_train, const float* frame_const,
                         const float* grid_cl, const float* ro, const float* rd, const float* z, int num_rays,
                         int num_samples, float* raw_out, void* tape_x, void* masks, float* saves, void* stream);
/* d raw[R,S,16] -> activation gradients of every layer (gradient tape) + scatter into the embedding-grid gradient
 * (channel-last fp32, accumulated atomically; may be NULL).  `scale` (device scalar) multiplies d_raw so that the fp16
 * gradient chain stays in range; the gradient tape and grid_grad carry that factor (the caller divides it out).  What autograd does for the reference modules
 * (nerf/modules.py:254-295, :371-390, :444-462, nerf/models.py:301-365); weight gradients are dY^T X over the tapes. */
int sahs_field_bwd(const sahs_model_spec* spec, int level, const void* packed_t, const float* frame_const,
                   const float* grid_cl, const float* ro, const float* rd, const float* z, int num_rays,
                   int num_samples, const float* d_raw, const float* scale, const void* masks, const float* saves,
                   void* tape_d, float* grid_grad, void* stream);

/* dW_l = dY_l^T X_{l-1} and db_l = sum_p dY_l for every layer, accumulated (atomically) into the caller's
 * zero-initialised fp32 gradient buffers `grads_host_array` (device pointers in the canonical parameter order; entry 0,
 * the embedding grid, is not touched).  Reads the tape chunks with TMA bulk loads.  The frame-constant input columns
 * of folded layers are left untouched (their gradient is the rank-1 product db x cvec).  `units_workspace`: >= 256 KB
 * of device memory that receives the kernel's work-unit list.  `upload_token` (host memory owned by the caller, one
 * per workspace, set to 0 whenever the workspace is (re)allocated or may have been overwritten) lets steady-state
 * steps skip the upload: the library stores a fingerprint of the list it last uploaded INTO THIS WORKSPACE there and
 * uploads (synchronously) only when the fingerprint changes.  NULL: upload on every call.  Results carry the `scale`
 * factor of sahs_field_bwd. */
int sahs_field_wgrad(const sahs_model_spec* spec, int level, float* const* grads_host_array, const void* tape_x,
                     const void* tape_d, int num_points, void* units_workspace, size_t workspace_bytes,
                     unsigned long long* upload_token, void* stream);

/* ---- in-kernel random draws (stochastic mode: perturb, radiance_field_noise_std, random u) ------------------------- */
/* Philox-4x32-10 keyed by `seed` (+ *counter_dev when given: read on the device, so a CUDA graph replays fresh draws),
 * counter = (element index, stream): element (ray, sample) of draw `stream` always gets the same value, whichever kernel
 * asks -- the compositing forward and backward regenerate identical noise instead of storing it.  Streams used by
 * predict_and_render_radiance: 0 t_rand (ref: nerf/train_utils.py:112), 1 coarse noise, 2 u (nerf_helpers.py:473),
 * 3 fine noise (volume_rendering_utils.py:47).  Uniforms are k * 2^-24, k in [0, 2^24) like torch.rand; normals are
 * Box-Muller over two 24-bit uniforms.  The reference draws from torch's generator, so parity in this mode is in
 * distribution; sahs_rng_fill materialises the same values for tests. */
typedef struct sahs_rng {
  unsigned long long seed;
  const unsigned long long* counter_dev; /* may be NULL */
  unsigned int stream;
  unsigned int reserved;
} sahs_rng;
int sahs_rng_fill(float* out, int64_t n, const sahs_rng* rng, int normal, float scale, void* stream);
/* sahs_coarse_z with t_rand drawn in the kernel (rng != NULL: perturbed depths). */
int sahs_coarse_z_rng(int num_rays, int num_samples, float near_, float far_, int lindisp, const float* t_vals,
                      const sahs_rng* rng, float* z_out, void* stream);

/* ---- (3) alpha compositing ---------------------------------------------------------------------- */
/* volume_render_radiance_field, ref: nerf/volume_rendering_utils.py:7-78 (+ cumprod_exclusive,
 * nerf/nerf_helpers.py:99-120) fused with the background overwrite raw[:, -1, :-1] = background_prior
 * (ref: nerf/train_utils.py:135-136; applied on the fly when apply_bg_overwrite != 0, raw is not modified).
 * raw [R,S,16]; z [R,S]; rd [R,3]; noise [R,S] (already scaled) or NULL; bg [R,bg_ch] or NULL (bg_ch 15 => seg
 * softmax branch).  Outputs rgb_map [R,15], disp/acc/depth [R], weights [R,S]. */
int sahs_composite_fwd(const float* raw, const float* z, const float* rd, const float* noise, const float* bg,
                       int bg_ch, int apply_bg_overwrite, int num_rays, int num_samples, int white_background,
                       float* rgb_map, float* disp, float* acc, float* weights, float* depth, void* stream);
/* backward of the above w.r.t. raw: inputs d_rgb_map [R,15], d_disp/d_acc/d_depth [R] (NULL = zero),
 * d_weights [R,S] (NULL = zero); output d_raw [R,S,16]. */
int sahs_composite_bwd(const float* raw, const float* z, const float* rd, const float* noise, const float* bg,
                       int bg_ch, int apply_bg_overwrite, int num_rays, int num_samples, int white_background,
                       const float* d_rgb_map, const float* d_disp, const float* d_acc, const float* d_weights,
                       const float* d_depth, float* d_raw, void* stream);
/* The same two kernels with the density noise drawn in the kernel: noise = noise_std * N(0,1) from `rng` (identical in
 * forward and backward); no [R,S] noise tensor exists. */
int sahs_composite_fwd_rng(const float* raw, const float* z, const float* rd, float noise_std, const sahs_rng* rng,
                           const float* bg, int bg_ch, int apply_bg_overwrite, int num_rays, int num_samples,
                           int white_background, float* rgb_map, float* disp, float* acc, float* weights, float* depth,
                           void* stream);
int sahs_composite_bwd_rng(const float* raw, const float* z, const float* rd, float noise_std, const sahs_rng* rng,
                           const float* bg, int bg_ch, int apply_bg_overwrite, int num_rays, int num_samples,
                           int white_background, const float* d_rgb_map, const float* d_disp, const float* d_acc,
                           const float* d_weights, const float* d_depth, float* d_raw, void* stream);

/* ---- (4) hierarchical importance resampling + merge ---------------------------------------------- */
/* sample_pdf_2 on bins = mid(z), weights[...,1:-1] (ref: nerf/nerf_helpers.py:454-497, call site
 * nerf/train_utils.py:157-164) followed by sort(cat(z, z_samples)) (ref: nerf/train_utils.py:166).
 * z [R,S], weights [R,S] (full compositing weights), u: device [n_fine] shared (det) when u_per_ray == 0 or
 * [R,n_fine] when u_per_ray != 0.  Outputs z_samples [R,n_fine], z_merged [R,S+n_fine], inds [R,n_fine] int64
 * (inds may be NULL).  Reproduces ATen's CPU summation orders so indices are bit-exact vs the reference on CPU. */
int sahs_sample_pdf_merge(const float* z, const float* weights, const float* u, int u_per_ray, int num_rays,
                          int num_samples, int num_fine, float* z_samples, float* z_merged, int64_t* inds,
                          void* stream);
/* sahs_sample_pdf_merge with u ~ U[0,1) drawn in the kernel (det = False, ref: nerf/nerf_helpers.py:473). */
int sahs_sample_pdf_merge_rng(const float* z, const float* weights, const sahs_rng* rng, int num_rays, int num_samples,
                              int num_fine, float* z_samples, float* z_merged, int64_t* inds, void* stream);

/* sample_pdf_2(bins [R,nb], weights [R,nb-1], num_fine) alone, the reference's public helper signature
 * (ref: nerf/nerf_helpers.py:454-497); same arithmetic as above without the merge. */
int sahs_sample_pdf(const float* bins, const float* weights, const float* u, int u_per_ray, int num_rays,
                    int num_bins, int num_fine, float* samples, int64_t* inds, void* stream);

/* ---- frame post-processing (the step after the path: what the eval script writes / Stage II consumes) ----------- */
/* rgb_map[R,15] -> uint8 rgb [R,3] (clamp, *255, truncate; ref: eval_stage_rays.py:221-227), argmax semantic label [R]
 * and its palette colour [R,3] in the reference's reversed channel order (ref: nerf/utils.py:112-140).  label_u8 and
 * seg_color_u8 may be NULL. */
int sahs_frame_postprocess(const float* map15, int64_t num_rays, uint8_t* rgb_u8, uint8_t* label_u8,
                           uint8_t* seg_color_u8, void* stream);

/* Depth map -> normal map, `torch_normal_map` of the eval script (ref: eval_stage_rays.py:116-151).  depth [size, size]
 * (the reference only broadcasts for square maps), intrinsics [fx, fy, cx, cy] with cx, cy relative; weights [size, size]
 * = the fine pass's background weight for the clean-up step (NULL: `clean=False` / no weights); central_difference
 * != 0 uses a stride of 2.  normals_out: [size - k, size - k, 3] fp32 in [0, 255], k = 1 or 2. */
int sahs_normal_map(const float* depth, int size, float fx, float fy, float cx_rel, float cy_rel, const float* weights,
                    int central_difference, float* normals_out, void* stream);

/* Semantic-weighted ray batch on the device (replaces the host-side np.random.choice(H*W, n, replace=False, p=probs)
 * of train_stage_rays_auto.py:390-420): weight_i = sum_c class_prob[c] * mask[i, c] (mask: int32 [num_pixels,
 * num_classes], one-hot in the reference), num_select distinct pixel indices drawn without replacement with
 * probability proportional to the weights (exponential-clock keys from Philox-4x32-10(seed, pixel), radix select).
 * The result is reproducible as a set for a given seed; the order of out_indices is unspecified.  At least num_select
 * pixels must have a positive weight (the reference raises otherwise).  workspace: 2048 + 4 * num_pixels bytes. */
int sahs_weighted_sample(const int32_t* mask, const float* class_prob, int64_t num_pixels, int num_classes,
                         int num_select, uint64_t seed, int64_t* out_indices, void* workspace, size_t workspace_bytes,
                         void* stream);
/* Same draw with seed + *seed_counter_dev as the seed (read on the device: capturable in a CUDA graph). */
int sahs_weighted_sample_dev(const int32_t* mask, const float* class_prob, int64_t num_pixels, int num_classes,
                             int num_select, uint64_t seed, const unsigned long long* seed_counter_dev,
                             int64_t* out_indices, void* workspace, size_t workspace_bytes, void* stream);

/* Stage-I training loss and its gradient in two small launches (replaces ~75 elementwise/reduction launches per step).
 * ref: nerf/nerf_helpers.py:14-62 (MaskCrossEntropyLoss, MaskMSELoss), assembled as in train_stage_rays_auto.py:455-468:
 * per level  l2 + ce_weight * CE + mouth_weight * sum_{k in [mouth_lo, mouth_hi)} (masked_l2[k] + masked_CE[k])
 * (0.02, 0.005 and classes 7..8 in the script), summed over the coarse and fine maps; the cross entropy's target is
 * the mask, as at the call site.  map_coarse / map_fine [R,15] (rgb + 12 class probabilities; map_fine may be NULL),
 * target_rgb [R,3], mask [R,12] float.  Outputs: stats[53] = {loss, l2_c, ce_c, l2_f, ce_f, masked_l2_c[12],
 * masked_ce_c[12], masked_l2_f[12], masked_ce_f[12]}, sample_prob[12] (the dynamic sampling weights, :466-468) and
 * d loss / d map for both levels [R,15].  workspace: 4096 floats (per-CTA partial sums).  Two launches; deterministic
 * (fixed reduction order). */
int sahs_stage1_loss(const float* map_coarse, const float* map_fine, const float* target_rgb, const float* mask,
                     int num_rays, int num_classes, float ce_weight, float mouth_weight, int mouth_lo, int mouth_hi,
                     float* stats, float* sample_prob, float* d_map_coarse, float* d_map_fine, float* workspace,
                     void* stream);

/* One Adam step on flat fp32 buffers (parameters, gradients, first and second moments, n elements each, 16-byte
 * aligned).  ref: train_stage_rays_auto.py:200-210 (torch.optim.Adam, no weight decay / amsgrad) with the decayed
 * learning rate of :503-509 passed as `lr`.  step counts from 1 (bias correction); grads are multiplied by grad_scale
 * first (1 / world size after a sum all-reduce).  Same arithmetic as torch.optim.Adam's single-tensor path
 * (hyper-parameters are doubles because 1 - beta2 must be formed in double, as Python does). */
int sahs_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                   double beta1, double beta2, double eps, int step, float grad_scale, void* stream);

/* The same step with the step count and the exponential learning-rate schedule (ref: train_stage_rays_auto.py:503-509)
 * in DEVICE memory, so that a whole training step can be recorded in a CUDA graph and replayed without host code:
 * hyper_dev (8 doubles): [0] lr0, [1] decay factor, [2] decay steps (cfg.scheduler.lr_decay * 1000), [3] beta1,
 * [4] beta2, [5] eps, [6] t = the step being applied (>= 1), [7] i = schedule position, lr = lr0 * factor^(i / steps).
 * sahs_adam_advance adds 1 to [6] and [7] (one launch, after the step); sahs_counter_add adds `inc` to a device
 * counter (the per-step part of the ray sampler's seed, see sahs_weighted_sample_dev). */
int sahs_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const double* hyper_dev, float grad_scale, void* stream);
int sahs_adam_advance(double* hyper_dev, void* stream);
int sahs_counter_add(unsigned long long* counter_dev, unsigned long long inc, void* stream);

/* ---- Stage II: SPADE generator ------------------------------------------------------------ */
/* Every 3x3 convolution of Generator / Generator_audio (ref: nerf/_init_spade.py:114-139 SPADELayer, :183-199 IdEncoder,
 * :235-282 SPADEBlock, :286-315 RefineNetwork) as one implicit-GEMM tcgen05 kernel on NHWC fp16 activations.
 *   mode 0: nn.Conv2d(3x3, padding 1) on the input resized (nearest) by 2^up_shift / 2^down_shift to the output's size
 *           (F.interpolate(fid, size=x.size()) of :132 and nn.Upsample(scale_factor=2) of :254 are never materialised);
 *   mode 1: the same with stride 2 (residual_downsample :250, ResBlock2d downsample :17-18);
 *   mode 2: nn.ConvTranspose2d(3x3, stride 2, padding 1, output_padding 1) (residual_upsample :255);
 *   mode 3: 3-channel input stored as [H, W, 4] fp16: all nine taps in one K chunk (layer1 of both networks);
 *   mode 4: one output-parity class (t2_class = 2 py + px) of the mode-2 transposed conv as a plain conv over the input grid
 *           with the class's 1, 2, 2 or 4 taps, written to output pixels (2i + py, 2j + px): four launches per layer, no
 *           multiplications by zero (mode 2 in one launch spends 9 taps per output where 2.25 contribute).  packed_w then
 *           holds only the class's taps, [ky ascending][chunk][kx ascending].
 * packed_w: [ntiles][9 * cin / 64 chunks (mode 3: 1)][ntile rows x 64 columns fp16, 128B-swizzled K-major], chunk order
 * [ky][64-channel chunk][kx]; bias: [ntiles * ntile] fp32.  Built by sahs_b200/spade.py from the state_dict (spectral
 * norm and eval-mode BatchNorm folded in).
 * epilogue flags: 1 ReLU, 2 + aux (residual add, same geometry as the output), 8 fp32 output; or 4 = SPADE modulation:
 * each N tile is [gamma(64) | beta(64)] of 64 channels and the kernel writes
 *   leaky_relu((aux - mean) * rstd * (1 + gamma) + beta, 0.2)      (ref: :128-137 followed by the block's LeakyReLU(0.2))
 * with aux = the tensor being normalised (read at (oy >> aux_shift, ox >> aux_shift): a nearest-upsampled x1 stays at
 * its stored resolution) and mean / rstd from sahs_instnorm_stats. */
typedef struct sahs_conv_desc {
  const void* in;          /* fp16 NHWC [in_h, in_w, in_cs] */
  int in_h, in_w, in_cs, cin;
  int out_h, out_w;
  int mode, up_shift, down_shift;
  const void* packed_w;
  const float* bias;
  int ntile, ntiles;       /* output columns per CTA pass (16, 64 or 128) and number of such tiles */
  int epilogue;
  const void* aux;
  int aux_cs, aux_shift;
  const float* mean;
  const float* rstd;
  void* out;               /* fp16 NHWC [out_h, out_w, out_cs] (fp32 with flag 8) */
  int out_cs, cout;
  int t2_class;            /* mode 4 only */
} sahs_conv_desc;
int sahs_spade_conv(const sahs_conv_desc* desc, void* stream);
/* bounded-wait diagnostic of the conv kernel (0 = healthy), as sahs_field_status */
int sahs_spade_conv_status(int* out4_host);
/* Per-channel mean and 1 / sqrt(biased variance + eps) over num_pixels of an NHWC fp16 tensor (nn.InstanceNorm2d,
 * affine=False; ref: nerf/_init_spade.py:118).  Deterministic (no atomics).  sums_workspace: 512 * channels doubles. */
int sahs_instnorm_stats(const void* x, int64_t num_pixels, int channels, int channel_stride, float eps,
                        double* sums_workspace, float* mean, float* rstd, void* stream);
/* nn.AvgPool2d(2, stride=2) on NHWC fp16 (ref: :240-249, :292) */
int sahs_avgpool2(const void* x, int height, int width, int channels, void* y, void* stream);

/* Diagnostic word written by the field kernel when a bounded mbarrier wait times out (0 = healthy):
 * out4_host[0] code (+100 dgrad kernel, +200 wgrad kernel), [1] tag, [2] block, [3] thread.  The words live in mapped
 * host memory, so this works (and issues no CUDA call) after a kernel trapped. */
int sahs_field_status(int* out4_host);

/* Host-only plan introspection for the CPU test-suite (no device work); layouts documented in field_host.cu
 * (14 ints per stage, 8 per fold section, 3 per copy section, 21 dims). */
int sahs_debug_plan(const sahs_model_spec* spec, int param_count, int32_t* stages, int max_stages, int32_t* folds,
                    int max_folds, int32_t* copies, int max_copies, int32_t* dims_out);

/* debug_pass ids for sahs_field_fwd (value after bias+activation of that pass, fp32, tile 0, [128,256]) */
#define SAHS_DBG_NONE (-1)
#define SAHS_DBG_WARP(i) (i)            /* merged warp|hyper layer i: cols 0..wh-1 warp, wh.. hyper     */
#define SAHS_DBG_MAPPED 16              /* cols 0-2 mapped xyz, 3.. ambient, 8..39 grid feature         */
#define SAHS_DBG_TRUNK(i) (32 + (i))    /* trunk layer i; i == trunk_layers is fc_feat                  */
#define SAHS_DBG_HEAD(i) (64 + (i))     /* cols 0-127 dir hidden i, 128-255 seg hidden i                */
#define SAHS_DBG_PROF 99                /* debug buffer receives (tag, clock64) event pairs of one tile */
#define SAHS_DBG_PROF_LIGHT 98          /* same, per-pass events only */
#define SAHS_DBG_PROF_DUO 96            /* two-tile render kernel: (tag, clock64) events of both sets' 4th tile + the issuer */
#define SAHS_DBG_PROF_PROD 97           /* per-pass worker events from the PRODUCTION instantiation; needs a -DSAHS_PROF_PROD=1 build */

#ifdef __cplusplus
}
#endif
#endif /* SAHS_B200_H */
