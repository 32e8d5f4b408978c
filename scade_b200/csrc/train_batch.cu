// Training ray-batch construction on the device.
//
//   reference: get_ray_batch_from_one_image_hypothesis_idx (run_scade_scannet.py:772-827) -- every step it builds a full-image
//   meshgrid (RS:773), runs get_rays over all H*W pixels (RS:784), then gathers rays, colour targets, depths, the K depth
//   hypotheses (RS:791) and optionally cached uniforms / the corner mask (RS:805-821) at N_rand pixels with seven advanced-
//   indexing launches; render() then re-assembles the [N,11] ray batch (RS:123-141).
//   Here ONE kernel does all of it for the selected pixels only: one warp per selected pixel; lanes 0..10 write the ray-batch
//   row (same arithmetic as pixel_ray / write_batch_row in rays.cu, bit-identical to get_rays), the whole warp strides over the
//   K hypotheses and the cached uniforms of that pixel.  The pixel choice itself (np.random.choice, H:281) stays on the host so
//   that the reference's RNG stream is preserved; its indices are the only per-step host->device traffic (8 B per ray).
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace scade {

struct BatchCamera {
  float fx, fy, cx, cy;
  float r[3][3];
  float t[3];
};

struct GatherArgs {
  BatchCamera cam;
  int H, W;
  const int64_t* select;       // [N] flat pixel indices (row * W + col)
  int64_t N;
  float near, far;
  const float* image;          // [H,W,3]
  const float* depth; int Cd;  // [H,W,Cd] or null
  const uint8_t* valid;        // [H,W] bool or null
  const void* hyp; int K;      // [K,H,W] fp32 (or fp16 when hyp_f16: the resident hypothesis store) or null
  int hyp_f16;
  const float* cached_u; int Nu;   // [H,W,Nu] or null
  int mask_corners;
  float* ray_batch;            // [N,11] or null
  float* rays_od;              // [2,N,3] or null
  float* target_s;             // [N,3]
  float* target_d;             // [N,Cd]
  uint8_t* target_vd;          // [N]
  float* target_h;             // [K,N]
  float* mask;                 // [N]
  float* u_out;                // [N,Nu]
};

__global__ void __launch_bounds__(256) gather_train_batch_kernel(const __grid_constant__ GatherArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= a.N) return;
  const int64_t pix = a.select[n];
  const int row = (int)(pix / a.W), col = (int)(pix % a.W);
  // H:295-304 -- identical operation order to rays.cu::pixel_ray
  const float i = (float)col, j = (float)row;
  const float d0 = __fdiv_rn(__fsub_rn(__fadd_rn(i, 0.5f), a.cam.cx), a.cam.fx);
  const float d1 = __fdiv_rn(__fsub_rn(__fsub_rn((float)a.H, __fadd_rn(j, 0.5f)), a.cam.cy), a.cam.fy);
  float d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    d[c] = __fadd_rn(__fadd_rn(__fmul_rn(d0, a.cam.r[c][0]), __fmul_rn(d1, a.cam.r[c][1])), __fmul_rn(-1.0f, a.cam.r[c][2]));
  if (lane < 11 && a.ray_batch) {                                            // RS:123-141
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    float v;
    if (lane < 3) v = a.cam.t[lane];
    else if (lane < 6) v = d[lane - 3];
    else if (lane == 6) v = a.near;
    else if (lane == 7) v = a.far;
    else v = __fdiv_rn(d[lane - 8], nrm);
    a.ray_batch[n * 11 + lane] = v;
  }
  if (lane < 3) {
    if (a.rays_od) {                                                         // batch_rays = stack([rays_o, rays_d])  RS:824
      a.rays_od[n * 3 + lane] = a.cam.t[lane];
      a.rays_od[(a.N + n) * 3 + lane] = d[lane];
    }
    a.target_s[n * 3 + lane] = a.image[pix * 3 + lane];                      // RS:788
  }
  if (a.depth && lane < a.Cd) a.target_d[n * a.Cd + lane] = a.depth[pix * a.Cd + lane];         // RS:789
  if (lane == 0) {
    if (a.valid) a.target_vd[n] = a.valid[pix];                              // RS:790
    if (a.mask) {                                                            // RS:810-821: 20 x 20 pixel corners are masked out
      const bool edge_r = row < 20 || row >= a.H - 20, edge_c = col < 20 || col >= a.W - 20;
      a.mask[n] = (a.mask_corners && edge_r && edge_c) ? 0.0f : 1.0f;
    }
  }
  if (a.hyp) {
    const int64_t plane = (int64_t)a.H * a.W;
    if (a.hyp_f16) {
      const __half* h16 = reinterpret_cast<const __half*>(a.hyp);
      for (int k = lane; k < a.K; k += 32) a.target_h[(int64_t)k * a.N + n] = __half2float(h16[k * plane + pix]);
    } else {
      const float* h32 = reinterpret_cast<const float*>(a.hyp);
      for (int k = lane; k < a.K; k += 32) a.target_h[(int64_t)k * a.N + n] = h32[k * plane + pix];   // RS:791
    }
  }
  if (a.cached_u)
    for (int u = lane; u < a.Nu; u += 32) a.u_out[n * a.Nu + u] = a.cached_u[pix * a.Nu + u];         // RS:805-806
}

// np.clip(d, near, far) (data/load_scene.py:348) + fp32 -> fp16, for the resident hypothesis store
__global__ void pack_hypotheses_f16_kernel(const float* __restrict__ in, int64_t n, float lo, float hi, __half* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(fminf(fmaxf(in[i], lo), hi));
}

}  // namespace scade

using namespace scade;

static int gather_train_batch_impl(int H, int W, const float* intrinsic_host, const float* c2w_host,
                                        const int64_t* select_inds, int64_t N, float near, float far, const float* image,
                                        const float* depth, int depth_channels, const uint8_t* valid_depth,
                                        const void* hypotheses, int hyp_f16, int K, const float* cached_u, int n_u, int mask_corners,
                                        float* ray_batch, float* rays_o_d, float* target_s, float* target_d,
                                        uint8_t* target_vd, float* target_h, float* mask, float* u_out, void* stream) {
  SCADE_CHECK_ARG(H > 0 && W > 0 && N >= 0 && intrinsic_host && c2w_host, "gather_train_batch: bad image / camera");
  SCADE_CHECK_ARG(N == 0 || (select_inds && image && target_s), "gather_train_batch: null pointer");
  SCADE_CHECK_ARG(!depth || (target_d && depth_channels >= 1 && depth_channels <= 32), "gather_train_batch: depth needs target_d and 1..32 channels");
  SCADE_CHECK_ARG(!valid_depth || target_vd, "gather_train_batch: valid_depth needs target_vd");
  SCADE_CHECK_ARG(!hypotheses || (target_h && K > 0), "gather_train_batch: hypotheses need target_h and K > 0");
  SCADE_CHECK_ARG(!cached_u || (u_out && n_u > 0), "gather_train_batch: cached_u needs u_out and n_u > 0");
  if (N == 0) return SCADE_OK;
  GatherArgs a{};
  a.cam.fx = intrinsic_host[0]; a.cam.fy = intrinsic_host[1]; a.cam.cx = intrinsic_host[2]; a.cam.cy = intrinsic_host[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) a.cam.r[r][c] = c2w_host[r * 4 + c];
    a.cam.t[r] = c2w_host[r * 4 + 3];
  }
  a.H = H; a.W = W; a.select = select_inds; a.N = N; a.near = near; a.far = far;
  a.image = image; a.depth = depth; a.Cd = depth_channels; a.valid = valid_depth; a.hyp = hypotheses; a.hyp_f16 = hyp_f16; a.K = K;
  a.cached_u = cached_u; a.Nu = n_u; a.mask_corners = mask_corners;
  a.ray_batch = ray_batch; a.rays_od = rays_o_d; a.target_s = target_s; a.target_d = target_d; a.target_vd = target_vd;
  a.target_h = target_h; a.mask = mask; a.u_out = u_out;
  gather_train_batch_kernel<<<(unsigned)ceil_div<int64_t>(N, 8), 256, 0, as_stream(stream)>>>(a);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

extern "C" int scade_gather_train_batch(int H, int W, const float* intrinsic_host, const float* c2w_host,
                                        const int64_t* select_inds, int64_t N, float near, float far, const float* image,
                                        const float* depth, int depth_channels, const uint8_t* valid_depth,
                                        const float* hypotheses, int K, const float* cached_u, int n_u, int mask_corners,
                                        float* ray_batch, float* rays_o_d, float* target_s, float* target_d,
                                        uint8_t* target_vd, float* target_h, float* mask, float* u_out, void* stream) {
  return gather_train_batch_impl(H, W, intrinsic_host, c2w_host, select_inds, N, near, far, image, depth, depth_channels, valid_depth,
                                 hypotheses, 0, K, cached_u, n_u, mask_corners, ray_batch, rays_o_d, target_s, target_d, target_vd,
                                 target_h, mask, u_out, stream);
}

extern "C" int scade_gather_train_batch_h16(int H, int W, const float* intrinsic_host, const float* c2w_host,
                                            const int64_t* select_inds, int64_t N, float near, float far, const float* image,
                                            const float* depth, int depth_channels, const uint8_t* valid_depth,
                                            const uint16_t* hypotheses_f16, int K, const float* cached_u, int n_u, int mask_corners,
                                            float* ray_batch, float* rays_o_d, float* target_s, float* target_d,
                                            uint8_t* target_vd, float* target_h, float* mask, float* u_out, void* stream) {
  return gather_train_batch_impl(H, W, intrinsic_host, c2w_host, select_inds, N, near, far, image, depth, depth_channels, valid_depth,
                                 hypotheses_f16, 1, K, cached_u, n_u, mask_corners, ray_batch, rays_o_d, target_s, target_d, target_vd,
                                 target_h, mask, u_out, stream);
}

extern "C" int scade_pack_hypotheses_f16(const float* hyp, int64_t n, float near, float far, uint16_t* out, void* stream) {
  SCADE_CHECK_ARG(hyp && out && n >= 0, "pack_hypotheses_f16: bad arguments");
  if (n == 0) return SCADE_OK;
  const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256), 8 * num_sms());
  pack_hypotheses_f16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(hyp, n, near, far, reinterpret_cast<__half*>(out));
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}
