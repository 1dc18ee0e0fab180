/*
 * rnerf_b200.h -- C ABI of the B200-native Robust-NeRF render-and-train hot path.
 *
 * The reference (ShawnnnLiu/Robust-NeRF) is pure Python/PyTorch and has no FFI; the
 * drop-in boundary is therefore its Python API.  This header is the
 * boundary UNDER that API: one `extern "C"` entry point per kernel family, plain device
 * pointers + sizes + a CUDA stream, no torch types.  Each entry cites the reference code
 * (path:line relative to the reference repo) whose aten-op sequence it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - every function enqueues work on `stream` and returns immediately (no sync, no alloc);
 *   - return value: RN_OK (0) or an rn_status error code; never throws;
 *   - all float tensors are contiguous fp32 row-major unless stated; indices are int64
 *     like the reference's (`image_indices`, `searchsorted` output).
 */
#ifndef RNERF_B200_H_
#define RNERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rn_stream_t; /* cudaStream_t */

typedef enum {
  RN_OK = 0,
  RN_ERR_INVALID_ARG = 1,    /* null pointer / bad size / unsupported configuration   */
  RN_ERR_CUDA = 2,           /* a CUDA runtime call or launch failed (see rn_last_cuda_error) */
  RN_ERR_UNSUPPORTED_ARCH = 3, /* device is not sm_100 (tcgen05/TMEM/TMA required)     */
  RN_ERR_DRIVER = 4          /* cuTensorMapEncodeTiled unavailable / failed            */
} rn_status;

int rn_version(void);
const char* rn_status_string(int status);
int rn_last_cuda_error(void);            /* cudaError_t of the last RN_ERR_CUDA, else 0 */
int rn_device_sm_count(int* sm_count_host);
/* number of kernels this library has launched so far in this process (bench.py: gpu_launches) */
unsigned long long rn_launch_count(void);
/* measurement hooks: when enabled, every tcgen05 GEMM launch is bracketed by CUDA events on its
 * stream; rn_prof_collect synchronises and returns, per mode (0 NT fwd, 1 NN dgrad, 2 TN wgrad,
 * 3 = NN chain and TN stream running side by side, timed as one span), the summed kernel
 * milliseconds, executed FLOPs (2*M*N*K incl. padding) and launch counts: arrays of FOUR. */
/* Kernel-variant flags (A/B timing and cross-checks; every variant is parity-tested against the others):
 *   0: MLP forward as one launch per layer (0) or as the CTA-pair chain (2, default)
 *   3: data gradients as one launch per layer (0) or as the CTA-pair chain (1, default)
 *   4: inference encodes the points inside the forward chain (1, default) or in a separate kernel (0)
 *   5: programmatic dependent launch for the GEMM-family kernels (default 0)
 *   6: L2 cache hints, bit mask: 1 chains, 2 split-K loads evict_first (default 0), 8 = no policies on the hand-off
 *      between the data-gradient chain and the weight-gradient stream (they are on by default)
 *   9: SMs given to the weight-gradient stream that runs beside the data-gradient chain (default: 84 on a 148-SM
 *      part; 0 = off: one split-K launch per layer after the chain)
 *   7, 8, 10, 11: measurement only (SM limit of the data-gradient chain / of the split-K kernels; bit mask of
 *             timing options for the overlapped backward, 32 = record hand-off lags for rn_debug_stream_lag;
 *             microseconds over which the chain staggers the start of its clusters).  Flag 1 exists only in
 *             RN_EXPERIMENTS builds.
 * rn_get_flag reads a flag back.  Unknown flags return RN_ERR_INVALID_ARG. */
int rn_set_flag(int flag, int value);
int rn_get_flag(int flag, int* value_host);
int rn_prof_enable(int on);
/* measurement hook: mean / max time in microseconds between the data-gradient chain publishing a 128-point block and
 * the weight-gradient stream issuing its load, over the last stream launch made with rn_set_flag(10, 32) */
int rn_debug_stream_lag(double* mean_us_host, double* max_us_host, int* ctas_host);
/* ... and the microseconds each CTA pair's leader (entries 0, 2, 4, ...) spent on its chunks in that launch */
int rn_debug_stream_busy(unsigned int* us_host, int n);
/* host logic only: splits per GEMM (splits_out[<= 11]) the stream would use on `sms` SMs; RN_ERR_INVALID_ARG if too few */
int rn_debug_stream_plan(int sms, int* splits_out, int* n_problems_out);
int rn_prof_collect(double* ms4_host, double* flops4_host, int* launches4_host);

/* ------------------------------------------------------------------------------------------
 * Network geometry (fixed: the reference's default ModelConfig, noisy_src/config.py:10-24;
 * other values are rejected in Python with NotImplementedError -- no fallback path).
 * ---------------------------------------------------------------------------------------- */
#define RN_NUM_PARAM_TENSORS 24      /* state_dict order: pts_linears.{0..7}.{weight,bias},
                                        sigma_linear, feature_linear, dir_linear, rgb_linear */
#define RN_NUM_PARAMS 595844         /* outputs/lego_clean_20251206_210328/summary.json:46 */
#define RN_MLP_FLOP_PER_POINT 1186816 /* SURVEY.md section 8(a) row A1 */

/* ---- SE(3) pose parameters: CameraPoseParameters.get_poses, train_pose_opt.py:122-226 ---- */
int rn_se3_poses_fwd(const float* initial_poses /*[n_total,4,4]*/, const float* rot_deltas /*[n_total,3]*/,
                     const float* trans_deltas /*[n_total,3]*/, const int64_t* indices /*[n] or NULL = arange*/,
                     int n, int n_total, int learn_rotation, int learn_translation,
                     float* poses_out /*[n,4,4]*/, rn_stream_t stream);
/* d_rot / d_trans are ACCUMULATED into (caller zero-fills); either may be NULL. */
int rn_se3_poses_bwd(const float* initial_poses, const float* rot_deltas, const int64_t* indices,
                     int n, int n_total, const float* g_poses /*[n,4,4]*/,
                     float* d_rot /*[n_total,3]*/, float* d_trans /*[n_total,3]*/, rn_stream_t stream);

/* ---- ray generation: rays.py:17-99, data_pose_opt.py:83-148,200-223 ---- */
int rn_ray_directions(int H, int W, float focal, float cx, float cy, float* dirs /*[H,W,3]*/, rn_stream_t stream);
int rn_get_rays(const float* directions /*[n,3]*/, const float* c2w /*[4,4]*/, int64_t n,
                float* rays_o /*[n,3]*/, float* rays_d /*[n,3]*/, rn_stream_t stream);
int rn_get_rays_bwd(const float* directions, const float* c2w, int64_t n, const float* g_o, const float* g_d,
                    float* g_c2w /*[4,4], overwritten*/, float* g_directions /*[n,3] or NULL*/, rn_stream_t stream);
/* pixel batch -> rays, poses[image_idx] gathered on the fly (net effect of get_rays_for_batch) */
int rn_raygen_fwd(const int64_t* image_idx /*[B]*/, const float* pixel_uv /*[B,2] (u,v) as float*/, int64_t B,
                  const float* poses /*[n_poses,4,4]*/, int n_poses, int H, int W, float focal, float cx, float cy,
                  float* rays_o, float* rays_d, rn_stream_t stream);
int rn_raygen_bwd(const int64_t* image_idx, const float* pixel_uv, int64_t B, const float* poses, int n_poses,
                  int H, int W, float focal, float cx, float cy, const float* g_o, const float* g_d,
                  float* g_poses /*[n_poses,4,4], overwritten; deterministic per-image reduction*/, rn_stream_t stream);
/* fused: exp-map pose update + ray generation (north_star subsystem 1) and its backward into
 * the rotation / translation parameters (train_pose_opt.py:340-341 in one launch each way). */
int rn_raygen_se3_fwd(const int64_t* image_idx, const float* pixel_uv, int64_t B, const float* initial_poses,
                      const float* rot_deltas, const float* trans_deltas, int n_poses, int learn_rotation,
                      int learn_translation, int H, int W, float focal, float cx, float cy,
                      float* rays_o, float* rays_d, rn_stream_t stream);
int rn_raygen_se3_bwd(const int64_t* image_idx, const float* pixel_uv, int64_t B, const float* initial_poses,
                      const float* rot_deltas, int n_poses, int learn_rotation, int H, int W, float focal,
                      float cx, float cy, const float* g_o, const float* g_d,
                      float* d_rot /*[n_poses,3] overwritten*/, float* d_trans /*[n_poses,3] overwritten*/,
                      rn_stream_t stream);
/* PixelSampler.sample_batch bookkeeping (data_pose_opt.py:56-76,188-198) without the tables:
 * image = idx / (H*W), v = (idx % (H*W)) / W, u = idx % W; rgb gathered from images[N,H,W,3]. */
int rn_pixel_gather(const int64_t* flat_idx /*[B]*/, int64_t B, int H, int W, const float* images /*or NULL*/,
                    int64_t* image_idx_out, float* pixel_uv_out, float* target_rgb_out /*or NULL*/, rn_stream_t stream);
/* Same with the training images kept as uint8 [N,H,W,3] (a quarter of the fp32 table: 192 MB instead of 768 MB at
 * 100 x 800^2).  Lossless for the reference's data: data.py:118-135 quantises every image to uint8 and divides by
 * 255.0 in fp32, which is exactly what this gather does (rgb = float(k) / 255.0f, correctly rounded). */
int rn_pixel_gather_u8(const int64_t* flat_idx /*[B]*/, int64_t B, int H, int W, const uint8_t* images_u8,
                       int64_t* image_idx_out, float* pixel_uv_out, float* target_rgb_out, rn_stream_t stream);

/* ---- evaluation metrics: metrics.py:15-116 ---- */
/* Per image of a batch pred/target [N,H,W,3] (fp32 in [0,1]): mean squared error (compute_mse; PSNR = 20 log10(max)
 * - 10 log10(mse) is left to the caller) and the mean Gaussian-window SSIM (compute_ssim: 11 x 11 window, sigma 1.5,
 * zero padding).  scratch: rn_image_metrics_scratch_bytes(N,H,W) bytes.  Deterministic (no atomics). */
size_t rn_image_metrics_scratch_bytes(int N, int H, int W);
int rn_image_metrics(const float* pred, const float* target, int N, int H, int W, float C1, float C2, float* scratch,
                     float* mse_out /*[N]*/, float* ssim_out /*[N]*/, rn_stream_t stream);

/* ---- pose-noise initialisation and pose-error tracking: noise.py:71-268, train_pose_opt.py:232-271
 *       ---- */
/* noisy_out[i] = add_noise_to_pose(poses[i]) for n camera-to-world matrices [n,4,4], given the raw standard-normal
 * draws of the reference's generator calls: g_angle [n] and g_axis [n,3] (random_rotation_matrix: angle = g * std_rad
 * about the normalised axis, Rodrigues, left-multiplied onto R; both NULL = no rotation noise) and g_trans [n,3]
 * (random_translation; NULL = none).  Translation std = |camera position| * trans_pct / 100 when trans_pct > 0
 * (NoiseConfig.get_translation_std), else trans_std_abs.  info_out [n,2] (may be NULL): actual_rotation_deg,
 * actual_translation_norm of add_noise_to_pose's noise_info. */
int rn_pose_noise(const float* poses, int n, const float* g_angle, const float* g_axis, const float* g_trans,
                  float rot_std_rad, float trans_std_abs, double trans_pct, float* noisy_out, float* info_out,
                  rn_stream_t stream);
/* err_out[i] = {rotation_error_deg, translation_error} of compute_pose_error(gt_poses[i], cur_poses[i]) (noise.py:237-268). */
int rn_pose_errors(const float* gt_poses, const float* cur_poses, int n, float* err_out /*[n,2]*/, rn_stream_t stream);

/* ---- sampling: rays.py:145-333 ---- */
/* z = lower + (upper-lower)*t_rand over the base depths z_base (linspace built by the caller,
 * rays.py:185-195); t_rand NULL = no perturbation.  pts may be NULL. */
int rn_stratified_fwd(const float* rays_o, const float* rays_d, int64_t B, const float* z_base /*[Nc]*/, int Nc,
                      const float* t_rand /*[B,Nc] or NULL*/, float* z_out /*[B,Nc]*/, float* pts_out /*[B,Nc,3] or NULL*/,
                      rn_stream_t stream);
/* pts = o + d*z (rays.py:208,331) and its backward into o, d */
int rn_points_fwd(const float* rays_o, const float* rays_d, const float* z /*[B,S]*/, int64_t B, int S,
                  float* pts /*[B,S,3]*/, rn_stream_t stream);
int rn_points_bwd(const float* g_pts /*[B,S,3]*/, const float* z, int64_t B, int S,
                  float* g_o /*[B,3]*/, float* g_d /*[B,3]*/, rn_stream_t stream);
/* sample_pdf (rays.py:213-279): warp-per-ray sequential-order cdf + binary search (right=True).
 * u has row stride u_stride (0 = one shared row, the det linspace). inds_out may be NULL. */
int rn_sample_pdf_fwd(const float* bins /*[B,nb]*/, const float* weights /*[B,nb-1]*/, int64_t B, int nb,
                      const float* u, int64_t u_stride, int Nf, float* samples /*[B,Nf]*/,
                      int64_t* inds_out /*[B,Nf] or NULL*/, rn_stream_t stream);
/* sample_hierarchical (rays.py:282-333): mid-point bins, interior weights, inverse cdf, sorted
 * merge with the coarse depths; z_all sorted ascending; pts_out may be NULL. */
int rn_sample_hierarchical_fwd(const float* rays_o, const float* rays_d, const float* z_coarse /*[B,Nc]*/,
                               const float* weights /*[B,Nc]*/, int64_t B, int Nc, const float* u, int64_t u_stride,
                               int Nf, float* z_all /*[B,Nc+Nf]*/, float* pts_out /*[B,Nc+Nf,3] or NULL*/,
                               int64_t* inds_out /*[B,Nf] or NULL*/, rn_stream_t stream);

/* ---- alpha compositing: rendering.py:20-116 ---- */
/* Inputs either (rgb[B,S,3], sigma[B,S]) post-activation like the reference, or raw4[B,S,4]
 * = (rgb pre-sigmoid x3, sigma pre-ReLU) straight from rn_mlp_fwd (rgb = sigma = NULL then).
 * noise: already scaled randn draw [B,S] or NULL.  t_min > 0 enables early termination
 * (samples whose incoming transmittance < t_min get weight 0); 0 = exact reference semantics. */
int rn_composite_fwd(const float* rgb, const float* sigma, const float* raw4, const float* z /*[B,S]*/,
                     const float* rays_d /*[B,3]*/, const float* noise, int64_t B, int S, int white_background,
                     float t_min, float* rgb_map /*[B,3]*/, float* depth_map /*[B]*/, float* acc_map /*[B]*/,
                     float* weights /*[B,S]*/, rn_stream_t stream);
/* g_depth / g_acc / g_weights / d_rays_d may be NULL. Outputs d_rgb[B,S,3] + d_sigma[B,S]
 * (post-activation mode) or d_raw4[B,S,4] (raw mode). */
int rn_composite_bwd(const float* rgb, const float* sigma, const float* raw4, const float* z, const float* rays_d,
                     const float* noise, int64_t B, int S, int white_background, const float* g_rgb_map,
                     const float* g_depth, const float* g_acc, const float* g_weights, float* d_rgb, float* d_sigma,
                     float* d_raw4, float* d_rays_d /*[B,3]*/, rn_stream_t stream);
/* MSE loss of train.py:89-99 on a rendered batch: loss_out[0] += mean((rgb_map-target)^2),
 * g_rgb_map = 2*(rgb_map-target)/(3B) * loss_scale.  loss_out is accumulated (caller zero-fills). */
int rn_mse_loss_fwd_bwd(const float* rgb_map, const float* target, int64_t B, float loss_scale,
                        float* loss_out /*[1]*/, float* g_rgb_map /*[B,3]*/, rn_stream_t stream);
/* The whole training loss of train.py:88-99 / train_pose_opt.py:355-375 in one launch: loss_out[0] = mse(coarse) +
 * mse(fine), [1] = coarse, [2] = fine (written, not accumulated); g_* = 2*(rgb-target)/(3B).  rgb_fine / g_fine may both
 * be NULL (no fine network). */
int rn_mse2_loss_fwd_bwd(const float* rgb_coarse, const float* rgb_fine, const float* target, int64_t B,
                         float* loss_out /*[3]*/, float* g_coarse /*[B,3]*/, float* g_fine /*[B,3] or NULL*/, rn_stream_t stream);

/* ---- NeRF MLP: model.py:20-196 (PE fused in; bf16 tcgen05 GEMMs, fp32 accumulate) ---- */
size_t rn_mlp_packed_weight_bytes(void);
/* params_host: 24 DEVICE pointers in state_dict order, array itself in HOST memory */
int rn_mlp_pack_weights(const float* const* params_host, void* packed, rn_stream_t stream);
size_t rn_mlp_workspace_bytes(int64_t M, int training);
/* raw_out[M,4] = (rgb pre-sigmoid x3, sigma pre-ReLU).  dirs[M/dir_group,3]: one view direction
 * per dir_group consecutive points (dir_group = samples per ray in render_rays; 1 for the generic
 * NeRF.forward(x, d)).  training=1 keeps every activation in `workspace` for rn_mlp_bwd. */
int rn_mlp_fwd(const void* packed, const float* pts /*[M,3]*/, const float* dirs, int64_t M, int dir_group,
               void* workspace, int training, float* raw_out, rn_stream_t stream);
/* grad_flat[RN_NUM_PARAMS] in state_dict order (overwritten); g_pts / g_dirs may be NULL
 * (clean-pose training needs neither). workspace is the one rn_mlp_fwd(training=1) filled. */
int rn_mlp_bwd(const void* packed, const float* pts, const float* dirs, int64_t M, int dir_group, void* workspace,
               const float* g_raw /*[M,4]*/, float* grad_flat, float* g_pts /*[M,3]*/, float* g_dirs /*[M/dir_group,3]*/,
               rn_stream_t stream);
/* head activations of model.py:181,194: rgb = sigmoid(raw[:, :3]), sigma = relu(raw[:, 3]) and backward */
int rn_head_act_fwd(const float* raw4, int64_t M, float* rgb /*[M,3]*/, float* sigma /*[M,1]*/, rn_stream_t stream);
int rn_head_act_bwd(const float* raw4, int64_t M, const float* g_rgb, const float* g_sigma, float* g_raw4, rn_stream_t stream);
/* positional encoding alone (model.py:58-80), fp32 in/out, for PositionalEncoding.forward */
int rn_posenc_fwd(const float* x /*[n,C]*/, int64_t n, int C, int num_freqs, float* out /*[n,C*(1+2L)]*/, rn_stream_t stream);
int rn_posenc_bwd(const float* x, int64_t n, int C, int num_freqs, const float* g_out, float* g_x, rn_stream_t stream);

/* ---- building block exposed for unit tests / profiling: one bf16 tcgen05 GEMM ----
 * mode 0 (NT): D[M,N] = act(A[M,K] * B[N,K]^T + bias)            forward layer
 * mode 1 (NN): D[M,N] = (A[M,K] * B[K,N]) (.) mask                data gradient
 * mode 2 (TN): D[Mo,N] (fp32) = A[K,Mo]^T * B[K,N]                weight gradient (split-K inside)
 * A, B, D(mode 0/1) are bf16 with leading dimensions in ELEMENTS; ld % 8 == 0.
 * ReLU masks are packed: [M][N/32] uint32 words, bit j of word c <=> element (m, 32c+j) > 0.
 * mode 0 optionally WRITES the mask of its (bf16-rounded) output to mask_out; mode 1 optionally
 * APPLIES mask_bits to its output. */
int rn_gemm_bf16(int mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd,
                 int64_t M, int N, int64_t K, const float* bias, int relu, const uint32_t* mask_bits, uint32_t* mask_out,
                 float* colsum_out /*mode 2: sum_k A[k,:] (bias gradient) or NULL*/, void* scratch, size_t scratch_bytes,
                 rn_stream_t stream);
size_t rn_gemm_scratch_bytes(void);

/* ---- one-call evaluation render (rendering.py:119-323 with is_train=False; train.py:122-160, inference.py:76-105) ---- */
/* Renders the rays [ray_begin, ray_end) of one view in tiles of `tile_rays`, dealing tiles round-robin: this call renders
 * tiles tile_first, tile_first + tile_step, ... (rank r of n: tile_first = r, tile_step = n; no collective).  Rays come
 * from the camera (`pose`: DEVICE 4x4 c2w; pixel (u, v) = (i % W, i / W), rays.py:17-99) or are given (rays_o / rays_d
 * [ray_end - ray_begin, 3], pose NULL).  z_base [Nc] and u_det [Nf] are the reference's linspace rows (rays.py:185-192,
 * 252), built once by the caller.  packed_fine NULL or Nf = 0: coarse only.  Outputs are indexed by ray - ray_begin; rays
 * of tiles this call does not own are left untouched.  workspace: rn_render_workspace_bytes(tile_rays, Nc, Nf).
 * *rays_rendered_host (optional, HOST) receives the number of rays rendered.  Nine launches per tile, no synchronisation. */
/* viewdirs = rays_d / |rays_d| (rendering.py:165), fp32, unfused multiply-adds */
int rn_view_dirs(const float* rays_d /*[B,3]*/, int64_t B, float* viewdirs /*[B,3]*/, rn_stream_t stream);
size_t rn_mlp_infer_workspace_bytes(int64_t M, int dir_group);
size_t rn_render_workspace_bytes(int64_t tile_rays, int Nc, int Nf);
int rn_render_view(const void* packed_coarse, const void* packed_fine, const float* pose, const float* rays_o, const float* rays_d,
                   int H, int W, float focal, float cx, float cy, int64_t ray_begin, int64_t ray_end, int64_t tile_rays,
                   int tile_first, int tile_step, const float* z_base, int Nc, const float* u_det, int Nf, int white_background,
                   void* workspace, float* rgb_out /*[n,3]*/, float* depth_out /*[n] or NULL*/, float* acc_out /*[n] or NULL*/,
                   int64_t* rays_rendered_host, rn_stream_t stream);

/* ---- optimiser tail (train.py:115-117, train_pose_opt.py:398-409): clip_grad_norm_ + Adam ---- */
/* One launch pair over a flat fp32 parameter/gradient/moment buffer.  `norm_groups` partitions the
 * buffer into ranges clipped independently (joint clip = 1 group; pose-opt = one group per net). */
int rn_clip_adam_step(float* params, float* grads /*scaled in place by the clip*/, float* exp_avg, float* exp_avg_sq, int64_t n,
                      const int64_t* group_offsets_host /*[n_groups+1]*/, const float* group_max_norm_host, int n_groups,
                      float lr, float beta1, float beta2, float eps, int step,
                      float* norms_out /*scratch+output, 8 + 8*64 floats; [0..n_groups) = gradient norms*/,
                      const float* hyper_dev /*NULL, or device [lr, 1-beta1^step, sqrt(1-beta2^step)] overriding lr/step
                                               (so a captured CUDA graph can be replayed across steps)*/,
                      float grad_scale /*applied to grads on load: 1 / world size when the buffer holds an all-reduced SUM
                                         (the mean is taken here instead of in a separate launch), else 1*/,
                      rn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RNERF_B200_H_ */
