/*
 * scade_b200 -- C ABI of the B200-native (sm_100a) SCADE per-ray renderer.
 *
 * The reference (mikacuy/scade @ 23139b1) has no FFI: its hot path is a set of Python callables
 * (SURVEY.md §8(b)).  This header is the boundary a binding of that path would target: every
 * entry point names the reference function it replaces (RS = run_scade_scannet.py,
 * H = model/run_nerf_helpers.py).  The Python mirror in scade_b200/ binds these with ctypes
 * (see INTEGRATION.md); nothing here depends on torch.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major fp32 unless the comment says HOST;
 *   - the caller allocates all outputs and workspaces; no function allocates or frees device memory;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work on it (no host sync);
 *   - return value 0 = ok, otherwise a scade_status; scade_last_error_string() describes the last
 *     failure on the calling thread.  No C++ exceptions cross the boundary.
 */
#ifndef SCADE_B200_H
#define SCADE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCADE_B200_VERSION 103

typedef enum {
  SCADE_OK = 0,
  SCADE_ERR_INVALID_ARGUMENT = 1,
  SCADE_ERR_UNSUPPORTED = 2,      /* shape outside what the selected kernel family handles */
  SCADE_ERR_WORKSPACE = 3,        /* workspace too small */
  SCADE_ERR_CUDA = 4              /* a CUDA runtime call / launch failed */
} scade_status;

/* Arithmetic used for the MLP's wide layers. */
typedef enum {
  SCADE_PREC_FP32 = 0,   /* fp32 FFMA GEMMs: the reference's own arithmetic (cuBLAS SGEMM / MKL) */
  SCADE_PREC_TC_F16 = 1, /* tcgen05 tensor cores, fp16 operands, fp32 accumulate in TMEM (fast mode; forward + backward) */
  SCADE_PREC_TC_F16X3 = 2 /* tcgen05 tensor cores at an fp32-level tolerance: every operand an fp16 (hi, lo) pair, three MMA
                             passes per product (hi*hi + lo*hi + hi*lo) into one fp32 accumulator (tight mode; forward only) */
} scade_precision;

int scade_version(void);
const char* scade_last_error_string(void);
/* Number of CUDA kernels this library has launched in the calling process (diagnostic; bench.py reports it). */
uint64_t scade_kernel_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Field network: NeRF(D, W, input_ch, input_ch_views, skips=[skip], use_viewdirs=True)  (H:193-247)
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t D;               /* number of pts_linears (reference default 8) */
  int32_t W;               /* layer width (256) */
  int32_t multires;        /* positional-encoding octaves for points (9 -> input_ch = 57), H:174-189 */
  int32_t multires_views;  /* octaves for view directions (0 -> 3 channels, identity) */
  int32_t skip;            /* i such that input_pts is re-concatenated after layer i (4); -1 = none, H:229-230 */
} scade_net_desc;

/* Number of parameter tensors = 2*D + 8, in the reference's state_dict order (H:206-219):
 *   pts_linears.0.weight, pts_linears.0.bias, ..., pts_linears.{D-1}.weight, .bias,
 *   views_linears.0.weight, .bias, feature_linear.weight, .bias, alpha_linear.weight, .bias,
 *   rgb_linear.weight, .bias.   Weights are (out, in) row-major exactly as nn.Linear stores them. */
#define SCADE_MAX_PARAM_TENSORS 40

typedef struct {
  scade_net_desc desc;
  const float* params[SCADE_MAX_PARAM_TENSORS]; /* HOST array of DEVICE pointers (fp32 master weights) */
  const void* packed_f16;                       /* DEVICE buffer written by scade_mlp_pack_f16 / scade_mlp_pack for the
                                                   precision the net is used with, or NULL */
} scade_net;

/* Bytes of the fp16 tile image used by SCADE_PREC_TC_F16 (0 if the shape is unsupported there). */
size_t scade_mlp_packed_bytes(const scade_net_desc* desc);
/* Re-pack fp32 master weights into the swizzled fp16 K-major tile stream the tcgen05 kernel
 * bulk-copies (call after every optimizer step).  Replaces nothing in the reference: torch keeps
 * fp32 weights and cuBLAS reads them directly (H:131, H:227). */
int scade_mlp_pack_f16(const scade_net* net, void* packed_out, void* stream);

/* The same two calls for a given tensor-core precision: SCADE_PREC_TC_F16 (identical to the two above) or
 * SCADE_PREC_TC_F16X3, whose stream interleaves a W_hi and a W_lo = fp16(w - W_hi) stage per K stage.  A scade_net used with
 * SCADE_PREC_TC_F16X3 points packed_f16 at a buffer packed for that precision. */
size_t scade_mlp_packed_bytes_for(const scade_net_desc* desc, int precision);
int scade_mlp_pack(const scade_net* net, int precision, void* packed_out, void* stream);

/* Workspace bytes for scade_mlp_forward* on P points; save_for_backward adds the activation stash. */
size_t scade_mlp_workspace_bytes(const scade_net_desc* desc, int64_t P, int precision, int save_for_backward);

/* run_network (RS:48-63) fused with the sampling of points along rays (RS:657):
 *   pts = o + d*z ; x = (pts - bb_center)*bb_scale ; embed (H:142-172) ; NeRF.forward (H:223-247).
 * rays: [N, ray_stride] with o at 0..2, d at 3..5, viewdirs at 8..10 (RS:628-632); z: [N,S];
 * raw_out: [N,S,4] = (rgb_raw3, softplus_beta10(alpha)).  bb_center: HOST float[3]. */
int scade_mlp_forward_rays(const scade_net* net, int precision, const float* rays, int ray_stride,
                           const float* z_vals, int64_t N, int S, const float* bb_center_host,
                           float bb_scale, float* raw_out, void* workspace, size_t workspace_bytes,
                           int save_for_backward, void* stream);

/* scade_mlp_forward_rays followed by scade_raw2outputs (RS:659-660 / RS:718-720) as ONE kernel: the alpha compositing
 * (compute_weights' exclusive transmittance product + the weighted sums of raw2outputs, RS:511-562) runs inside the network
 * kernel on the (rgb_raw, sigma) values its last layer just produced -- a compositor warp per CTA takes them over through a
 * 4 KB cache-resident slot of `workspace` -- so `raw` never travels through device memory (raw_out may be NULL; pass a buffer
 * for retraw).  Bit-identical to the two calls.  SCADE_PREC_TC_F16 only, S a multiple of 32 (64 + 128 = 192 included);
 * otherwise SCADE_ERR_UNSUPPORTED (scade_mlp_forward_rays_composite_supported tells beforehand).  weights [N,S] is required;
 * the maps are nullable.  workspace: scade_mlp_workspace_bytes(desc, N*S, SCADE_PREC_TC_F16, 0) bytes, 16-byte aligned. */
int scade_mlp_forward_rays_composite_supported(const scade_net_desc* desc, int precision, int S);
int scade_mlp_forward_rays_composite(const scade_net* net, int precision, const float* rays, int ray_stride,
                                     const float* z_vals, int64_t N, int S, const float* bb_center_host, float bb_scale,
                                     float* raw_out, float* weights, float* rgb_map, float* disp_map, float* acc_map,
                                     float* depth_map, void* workspace, size_t workspace_bytes, void* stream);

/* NeRF.forward (H:223-247) on an already embedded input x [P, input_ch + input_ch_views]. */
int scade_mlp_forward_embedded(const scade_net* net, int precision, const float* x, int64_t P,
                               float* out, void* workspace, size_t workspace_bytes,
                               int save_for_backward, void* stream);

/* Backward of either forward above (autograd of H:223-247, RS:985): d_out [P,4] -> gradients of all
 * parameter tensors, ACCUMULATED into grads[i] (HOST array of DEVICE pointers, same order/shape as
 * params).  Uses the stash a forward call with save_for_backward=1 left in `workspace` (same precision).
 * SCADE_PREC_TC_F16: dgrad chain and weight gradients on tcgen05 (fp16 operands, gradients scaled by a power of
 * two taken from the largest gradient entering the chain, fp32 accumulation, fp32 gradient tensors).
 * No gradient w.r.t. the inputs is produced (z samples are detached, RS:711). */
int scade_mlp_backward(const scade_net* net, int precision, const float* d_out, int64_t P,
                       float* const* grads_host, void* workspace, size_t workspace_bytes, void* stream);

/* Diagnostic: byte offsets of the SCADE_PREC_TC_F16 training stash inside the workspace of a forward call with
 * save_for_backward=1 (tests decode the stashed fp16 activations / gradients through it).  Fills out[0..n) with
 * T (128-point tiles), D, total, emb, feat, hv, dzv, dzf, maskv, alpha, gs, h[8], dz[8], maskh[8]; returns the count.
 * Every activation region is [T][chunks][128 rows][128 B]: the K-major SWIZZLE_128B image of a [128 points x 64
 * features] fp16 tile (16-byte piece j of row r sits at r*128 + ((j ^ (r & 7)) << 4)). */
int scade_mlp_tc_stash_layout(const scade_net_desc* desc, int64_t P, int64_t* out, int n);

/* Diagnostic (host arithmetic, no GPU): how scade_mlp_forward_rays_composite splits N rays of S samples over the SM pairs of a
 * device with n_sms SMs.  *clusters = SM pairs launched (2 CTAs each).  *chain_iters = 0: 512-point steps strided over the
 * clusters, every CTA's 256 points per step are whole rays (S in {32, 64, 128, 256}).  *chain_iters > 0: CTA c walks the
 * contiguous points [c * chain_iters * 256, (c + 1) * chain_iters * 256), a range that starts and ends on a ray boundary. */
int scade_mlp_composite_plan(int S, int64_t N, int n_sms, int* clusters, int* chain_iters);

/* Embedder.embed (H:171-172): x [P,3] -> [P, 3 + 6*multires]. */
int scade_embed(const float* x, int64_t P, int multires, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Rays and sample placement
 * --------------------------------------------------------------------------------------------- */
/* get_rays (H:285-305) for the pixel rectangle rows [0,H) x cols [col0, col0+ncols) (the with_5_9
 * crop of RS:109-116 is a column window).  intrinsic: HOST (fx,fy,cx,cy); c2w: HOST 3x4 row-major.
 * rays_o / rays_d: [H, ncols, 3]. */
int scade_get_rays(int H, int W, const float* intrinsic_host, const float* c2w_host, int col0, int ncols,
                   float* rays_o, float* rays_d, void* stream);

/* render()'s batch assembly (RS:123-141): [N,11] = (o, d, near, far, d/|d|). */
int scade_make_ray_batch(const float* rays_o, const float* rays_d, int64_t N, float near, float far,
                         float* ray_batch, void* stream);

/* get_rays + batch assembly in one pass for pixels [pix0, pix0+N) of the (cropped) image, row-major. */
int scade_camera_ray_batch(int H, int W, const float* intrinsic_host, const float* c2w_host, int col0,
                           int ncols, int64_t pix0, int64_t N, float near, float far, float* ray_batch,
                           void* stream);

/* Coarse sample placement (RS:640-655): t = linspace(0,1,Nc); z = near(1-t)+far*t (or lindisp, RS:651);
 * if t_rand != NULL apply perturb_z_vals (RS:564-579) with those uniforms [N,Nc]. */
int scade_coarse_z_vals(const float* rays, int ray_stride, int64_t N, int Nc, int lindisp,
                        const float* t_rand, float* z_vals, void* stream);

/* perturb_z_vals (RS:564-579) on given z [N,S] with explicit t_rand [N,S]. */
int scade_perturb_z_vals(const float* z_in, const float* t_rand, int64_t N, int S, float* z_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Alpha compositing: compute_weights + raw2outputs (RS:511-562)
 * --------------------------------------------------------------------------------------------- */
/* rays_d: [N, d_stride] (first 3 floats of each row used).  noise: [N,S] sigma noise already scaled by
 * raw_noise_std, or NULL (RS:544-552).  Any output pointer may be NULL. */
int scade_raw2outputs(const float* raw, const float* z_vals, const float* rays_d, int d_stride,
                      const float* noise, int64_t N, int S, float* rgb_map, float* disp_map,
                      float* acc_map, float* weights, float* depth_map, void* stream);

/* Autograd of raw2outputs w.r.t. raw.  Any d_* input may be NULL (= zero).  d_raw: [N,S,4]. */
int scade_raw2outputs_backward(const float* raw, const float* z_vals, const float* rays_d, int d_stride,
                               const float* noise, int64_t N, int S, const float* d_rgb_map,
                               const float* d_disp_map, const float* d_acc_map, const float* d_weights,
                               const float* d_depth_map, float* d_raw, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Hierarchical resampling: sample_pdf family (H:337-538) and the sort-merge of RS:713
 * --------------------------------------------------------------------------------------------- */
/* sample_pdf / sample_pdf_return_u / *_joint*: bins [N,B], weights [N,B-1], samples_out [N,n_samples].
 * u: explicit uniforms [N,n_samples] (load_u, H:393), or [n_samples] when u_is_joint (one row shared
 * by all rays, H:452-453), or NULL for det=True -> linspace(0,1,n_samples) (H:347).
 * u_out (nullable): the [N,n_samples] uniforms actually used (the `u` return of H:436). */
int scade_sample_pdf(const float* bins, const float* weights, int64_t N, int B, int n_samples,
                     const float* u, int u_is_joint, float* samples_out, float* u_out, void* stream);

/* Autograd of sample_pdf w.r.t. weights (bins carry no gradient on the reference path, RS:711).
 * u: [N,n_samples] as returned in u_out.  d_weights: [N,B-1]. */
int scade_sample_pdf_backward(const float* bins, const float* weights, const float* u, int64_t N, int B,
                              int n_samples, const float* d_samples, float* d_weights, void* stream);

/* The render_rays form (RS:702-713 / RS:723-726): bins are the mid-points of z_vals [N,S] and the
 * weights are weights[:,1:-1], both formed on the fly.  If z_merged != NULL also writes
 * sort(cat(z_vals, samples)) [N, S+n_samples] (RS:713).  z_std (nullable) [N]: std of the samples
 * (RS:744, unbiased=False). */
int scade_resample_from_z(const float* z_vals, const float* weights_full, int64_t N, int S, int n_samples,
                          const float* u, int u_is_joint, float* samples_out, float* u_out,
                          float* z_merged, float* z_std, void* stream);

/* Compositing + resampling of every ray in one launch: scade_raw2outputs (RS:530-562, no sigma noise) followed by
 * scade_resample_from_z on the weights it produced (RS:660 + RS:702-713 for the coarse pass, RS:720 + RS:723-730 for the fine
 * pass); the weights travel through shared memory.  Same results as the two calls.  weights [N,S] is required; the map
 * outputs and u_out / z_merged / z_std are nullable as in the two calls. */
int scade_composite_resample(const float* raw, const float* z_vals, const float* rays_d, int d_stride, int64_t N, int S,
                             float* rgb_map, float* disp_map, float* acc_map, float* weights, float* depth_map,
                             int n_samples, const float* u, int u_is_joint, float* samples_out, float* u_out,
                             float* z_merged, float* z_std, void* stream);

/* Backward of the above w.r.t. weights_full [N,S] (zero on the first and last column). */
int scade_resample_from_z_backward(const float* z_vals, const float* weights_full, const float* u, int64_t N,
                                   int S, int n_samples, const float* d_samples, float* d_weights_full,
                                   int accumulate, void* stream);

/* torch.sort(torch.cat([a, b], -1), -1) values (RS:713): a [N,Na], b [N,Nb] -> out [N,Na+Nb]. */
int scade_sort_merge(const float* a, int Na, const float* b, int Nb, int64_t N, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Losses: compute_space_carving_loss (H:93-128) and img2mse (H:11)
 * --------------------------------------------------------------------------------------------- */
/* pred [N,P]; hyp [K,N,1] (hyp_full=0) or [K,N,P] (hyp_full=1); mask [N] or NULL; loss_out: 1 float.
 * d_pred / d_hyp (nullable, same shapes as pred / hyp) receive the gradient of grad_scale * loss.
 * workspace: scade_space_carving_workspace_bytes(K,N,P) bytes. */
size_t scade_space_carving_workspace_bytes(int K, int64_t N, int P);
int scade_space_carving_loss(const float* pred, const float* hyp, int hyp_full, const float* mask, int K,
                             int64_t N, int P, int is_joint, float threshold, float grad_scale,
                             float* loss_out, float* d_pred, float* d_hyp, void* workspace,
                             size_t workspace_bytes, void* stream);

/* The default branch with the loss glue of RS:954 fused in: hyp_raw [K,N,1] are the RAW hypotheses, the kernel forms
 * h = hyp_raw * (*scale_dev) + (*shift_dev) itself (scale / shift: one DEVICE float each, the step's DEPTH_SCALES[img_i] /
 * DEPTH_SHIFTS[img_i]) and returns grad_scale * d loss / d scale and d loss / d shift in *d_scale, *d_shift (both nullable)
 * instead of a [K,N] hypothesis gradient -- the train step loses the elementwise mul / add and the two reductions autograd
 * ran for them.  accumulate != 0: the two values are ADDED to what *d_scale / *d_shift hold (the parameters' .grad slots).
 * `denominator` overrides N in the mean over rays when > 0 (ray-sharded training divides by the GLOBAL count). */
int scade_space_carving_loss_affine(const float* pred, const float* hyp_raw, const float* scale_dev,
                                    const float* shift_dev, const float* mask, int K, int64_t N, int P,
                                    float threshold, float grad_scale, int64_t denominator, float* loss_out,
                                    float* d_pred, float* d_scale, float* d_shift, int accumulate, void* stream);

/* The joint branch (H:115-119: mean over rays BEFORE the min over k) for a ray-sharded step (SURVEY 8(e) "Exception"), in two
 * halves around one all-reduce of K*P floats:
 *   accumulate: qsum_out[k,p] = sum over this rank's N rays of dist[k,n,p]  (zeroed first);
 *   (caller: sum qsum over the ranks)
 *   finish    : per p, k*(p) = first arg-min of qsum[k,p] / N_global; loss_out = mean_p min_k  (the GLOBAL loss, identical on every
 *               rank); d_pred / d_hyp (nullable) = gradient of grad_scale * loss w.r.t. this rank's N rays.
 * kstar_workspace: P int32.  With N_global == N and no all-reduce this is scade_space_carving_loss(is_joint=1). */
int scade_space_carving_joint_accumulate(const float* pred, const float* hyp, int hyp_full, const float* mask, int K,
                                         int64_t N, int P, float threshold, float* qsum_out, void* stream);
int scade_space_carving_joint_finish(const float* pred, const float* hyp, int hyp_full, const float* mask,
                                     const float* qsum, int K, int64_t N, int64_t N_global, int P, float threshold,
                                     float grad_scale, float* loss_out, float* d_pred, float* d_hyp,
                                     int32_t* kstar_workspace, void* stream);

/* img2mse (H:11): loss_out = mean((x-y)^2) over n elements; d_x (nullable) = grad_scale * d loss/d x.
 * `denominator` overrides n in the mean when > 0 (ray-sharded training divides by the GLOBAL count). */
int scade_img2mse(const float* x, const float* y, int64_t n, int64_t denominator, float grad_scale,
                  float* loss_out, float* d_x, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training ray-batch construction: get_ray_batch_from_one_image_hypothesis_idx (RS:772-827)
 * --------------------------------------------------------------------------------------------- */
/* For the N selected pixels (select_inds: DEVICE int64 flat indices row*W+col, the np.random.choice of H:281) of ONE
 * training image: get_rays (H:285-305) at those pixels, the [N,11] ray batch of render() (RS:123-141), and the gathers of
 * RS:788-821.  image [H,W,3]; depth [H,W,depth_channels] or NULL; valid_depth [H,W] (bool bytes) or NULL; hypotheses
 * [K,H,W] or NULL (all_hypothesis[img_i] with its trailing singleton axis dropped); cached_u [H,W,n_u] or NULL.
 * Outputs (NULL = skip, except target_s): ray_batch [N,11]; rays_o_d [2,N,3] (= batch_rays, RS:824); target_s [N,3];
 * target_d [N,depth_channels]; target_vd [N]; target_h [K,N]; mask [N] (1, or 0 inside the four 20x20 corners when
 * mask_corners, RS:810-821); u_out [N,n_u].  intrinsic_host: HOST (fx,fy,cx,cy); c2w_host: HOST 3x4 row-major. */
int scade_gather_train_batch(int H, int W, const float* intrinsic_host, const float* c2w_host,
                             const int64_t* select_inds, int64_t N, float near, float far, const float* image,
                             const float* depth, int depth_channels, const uint8_t* valid_depth,
                             const float* hypotheses, int K, const float* cached_u, int n_u, int mask_corners,
                             float* ray_batch, float* rays_o_d, float* target_s, float* target_d,
                             uint8_t* target_vd, float* target_h, float* mask, float* u_out, void* stream);

/* The same with the K hypotheses read from the resident fp16 store ("<img>_<k>.npy hypotheses -> pinned /
 * GPU-resident fp16 store"): hypotheses_f16 [K,H,W] IEEE half; target_h receives the exact fp32 value of each half. */
int scade_gather_train_batch_h16(int H, int W, const float* intrinsic_host, const float* c2w_host,
                                 const int64_t* select_inds, int64_t N, float near, float far, const float* image,
                                 const float* depth, int depth_channels, const uint8_t* valid_depth,
                                 const uint16_t* hypotheses_f16, int K, const float* cached_u, int n_u, int mask_corners,
                                 float* ray_batch, float* rays_o_d, float* target_s, float* target_d,
                                 uint8_t* target_vd, float* target_h, float* mask, float* u_out, void* stream);

/* Builds the store: out[i] = half(clip(hyp[i], near, far)) (the clip of data/load_scene.py:348), n elements. */
int scade_pack_hypotheses_f16(const float* hyp, int64_t n, float near, float far, uint16_t* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer: torch.optim.Adam(grad_vars, lr, betas=(0.9, 0.999)).step()  (RS:469, RS:993)
 * --------------------------------------------------------------------------------------------- */
/* One Adam step (no amsgrad / weight decay) over a flat fp32 range of n parameters, in place: param, exp_avg, exp_avg_sq
 * are updated from grad.  `step` is the 1-based step count AFTER this step (torch's state['step']); lr is the value
 * update_learning_rate (train_utils/hyperparameter_update.py:3-5) put into param_groups.  All four buffers are
 * 16-byte aligned device pointers.  Hyper-parameters are doubles (as in Python): 1 - beta and the bias corrections are
 * formed in double and rounded to fp32 once, like torch does. */
int scade_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                    double beta1, double beta2, double eps, int64_t step, void* stream);

/* The same step with the step count and the learning rate in DEVICE memory, so that the launch can be captured in a CUDA graph
 * and replayed every iteration: *step_dev is the number of steps taken so far (int64, incremented by the call), *lr_dev the
 * current learning rate (double).  Bias corrections are formed on the device in double with the same expressions. */
int scade_adam_step_graph(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                          const double* lr_dev, double beta1, double beta2, double eps, int64_t* step_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Video / eval post-processing: render_video (RS:236-262), write_images_with_metrics (RS:395-405)
 * --------------------------------------------------------------------------------------------- */
/* One frame: rgb [H,W,3], depth_map [H,W], z_vals / weights [H,W,S] (render() outputs) ->
 *   frame_out [H, 3W, 3] uint8 BGR (nullable) = [ to8b(rgb) | LUT_depth[to8b(depth / depth_scale)] | LUT_std[to8b(std)] ]
 *   depth_std_out [H,W] (nullable) = sqrt(clamp(sum_s (z - depth)^2 w, 0, 1))                      (RS:257-258)
 *   depth16_out [H,W] uint16 (nullable) = to16b(depth)                                             (H:14, RS:403)
 * to8b / to16b truncate like numpy's astype (H:13-14).  lut_depth / lut_std: DEVICE [256,3] BGR tables (cv2.applyColorMap's
 * COLORMAP_TURBO / COLORMAP_VIRIDIS, RS:255,259) or NULL for grey. */
int scade_video_frame(const float* rgb, const float* depth_map, const float* z_vals, const float* weights, int H,
                      int W, int S, float depth_scale, const uint8_t* lut_depth, const uint8_t* lut_std,
                      uint8_t* frame_out, float* depth_std_out, uint16_t* depth16_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * render_rays (RS:581-751), N_importance > 0 branch, forward only, one call, one stream
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t N_samples;      /* coarse samples per ray (RS:585) */
  int32_t N_importance;   /* importance samples per ray (RS:591), must be > 0 */
  int32_t lindisp;        /* RS:589 */
  int32_t precision;      /* scade_precision */
  int32_t is_joint;       /* RS:596: u rows shared by all rays */
  int32_t ray_stride;     /* floats per ray_batch row (11, or 14 with depth_range, RS:633) */
  float bb_center[3];     /* RS:52 */
  float bb_scale;
} scade_render_cfg;

typedef struct {          /* RS:733-744; any pointer may be NULL to skip that output */
  float* rgb_map;   /* [N,3] */
  float* disp_map;  /* [N] */
  float* acc_map;   /* [N] */
  float* depth_map; /* [N] */
  float* z_vals;    /* [N,Nc+Nf] */
  float* weights;   /* [N,Nc+Nf] */
  float* pred_hyp;  /* [N,Nf] */
  float* u;         /* [N,Nf] */
  float* raw;       /* [N,Nc+Nf,4] (retraw) */
  float* rgb0;      /* [N,3] */
  float* disp0;     /* [N] */
  float* acc0;      /* [N] */
  float* depth0;    /* [N] */
  float* z_vals0;   /* [N,Nc] */
  float* weights0;  /* [N,Nc] */
  float* z_std;     /* [N] */
} scade_render_out;

size_t scade_render_rays_workspace_bytes(const scade_render_cfg* cfg, const scade_net_desc* coarse,
                                         const scade_net_desc* fine, int64_t N);

/* t_rand [N,Nc] (NULL = no perturbation, i.e. perturb == 0 -> det=True in both sample_pdf calls,
 * RS:705,726); u_coarse / u_fine [N,Nf] or [Nf] if is_joint (required iff t_rand != NULL; u_fine is
 * the reference's cached_u, RS:597). */
int scade_render_rays_forward(const scade_render_cfg* cfg, const float* ray_batch, int64_t N,
                              const scade_net* coarse, const scade_net* fine, const float* t_rand,
                              const float* u_coarse, const float* u_fine, const scade_render_out* out,
                              void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCADE_B200_H */
