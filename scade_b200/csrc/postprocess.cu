// Video / eval post-processing on the device.
//
//   reference: render_video (run_scade_scannet.py:236-262) pulls rgb, depth_map, z_vals and weights of every frame to the host,
//   then per frame: to8b(rgb) (H:13), to8b(depth / far) + COLORMAP_TURBO, the depth standard deviation
//   sqrt(clamp(sum((z - depth)^2 w), 0, 1)) (RS:257-258) + COLORMAP_VIRIDIS, concatenated side by side, BGR for cv2.imwrite.
//   write_images_with_metrics (RS:395-405) stores to16b(depth) (H:14).
//   Here one kernel builds the finished [H, 3W, 3] uint8 BGR frame in HBM: one warp per pixel for the weighted variance over
//   the S samples (the only non-trivial traffic: 8 B per sample, HBM-bound), the 8-bit quantisation with numpy's
//   truncation semantics and the two 256-entry colour look-ups (tables supplied by the caller, cv2's own LUTs).
#include "common.cuh"

namespace scade {

// numpy: (255 * clip(x, 0, 1)).astype(uint8) -- the product is formed in float32 (x is float32), then truncated
__device__ __forceinline__ int quant8(float x) {
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  return (int)__fmul_rn(255.0f, x);
}

__global__ void __launch_bounds__(256) video_frame_kernel(const float* __restrict__ rgb, const float* __restrict__ depth,
                                                          const float* __restrict__ z_vals, const float* __restrict__ weights,
                                                          int64_t P, int S, int W, float depth_scale,
                                                          const uint8_t* __restrict__ lut_depth, const uint8_t* __restrict__ lut_std,
                                                          uint8_t* __restrict__ frame, float* __restrict__ depth_std_out,
                                                          uint16_t* __restrict__ depth16_out) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const float dm = depth[p];
  float var = 0.f;
  for (int s = lane; s < S; s += 32) {                                         // RS:257
    const float dz = __fsub_rn(z_vals[p * S + s], dm);
    var = __fadd_rn(var, __fmul_rn(__fmul_rn(dz, dz), weights[p * S + s]));
  }
  var = warp_sum(var);
  const float sd = sqrtf(fminf(fmaxf(var, 0.0f), 1.0f));                      // RS:258
  if (lane == 0) {
    if (depth_std_out) depth_std_out[p] = sd;
    if (depth16_out) depth16_out[p] = (uint16_t)(int)__fmul_rn(65535.0f, fminf(fmaxf(dm, 0.0f), 1.0f));      // H:14
  }
  if (frame == nullptr) return;
  const int64_t row = p / W, col = p % W;
  uint8_t* out = frame + (row * 3 * W) * 3;
  if (lane < 3) {
    out[col * 3 + lane] = (uint8_t)quant8(rgb[p * 3 + (2 - lane)]);           // to8b + COLOR_RGB2BGR  (RS:252-253)
  } else if (lane < 6) {
    const int c = lane - 3;
    const int q = quant8(__fdiv_rn(dm, depth_scale));                          // RS:255
    out[(W + col) * 3 + c] = lut_depth ? lut_depth[q * 3 + c] : (uint8_t)q;
  } else if (lane < 9) {
    const int c = lane - 6;
    const int q = quant8(sd);                                                  // RS:259
    out[(2 * W + col) * 3 + c] = lut_std ? lut_std[q * 3 + c] : (uint8_t)q;
  }
}

}  // namespace scade

using namespace scade;

extern "C" int scade_video_frame(const float* rgb, const float* depth_map, const float* z_vals, const float* weights, int H,
                                 int W, int S, float depth_scale, const uint8_t* lut_depth, const uint8_t* lut_std,
                                 uint8_t* frame_out, float* depth_std_out, uint16_t* depth16_out, void* stream) {
  SCADE_CHECK_ARG(H >= 0 && W > 0 && S > 0 && depth_map && z_vals && weights, "video_frame: bad arguments");
  SCADE_CHECK_ARG(frame_out == nullptr || rgb != nullptr, "video_frame: a frame needs rgb");
  SCADE_CHECK_ARG(depth_scale > 0.f, "video_frame: depth_scale must be positive");
  const int64_t P = (int64_t)H * W;
  if (P == 0) return SCADE_OK;
  video_frame_kernel<<<(unsigned)ceil_div<int64_t>(P, 8), 256, 0, as_stream(stream)>>>(rgb, depth_map, z_vals, weights, P, S, W,
                                                                                      depth_scale, lut_depth, lut_std, frame_out,
                                                                                      depth_std_out, depth16_out);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}
