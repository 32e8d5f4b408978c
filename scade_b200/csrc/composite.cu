// Alpha compositing: compute_weights + raw2outputs (run_scade_scannet.py:511-562), forward and backward.
//
// One warp per ray.  Samples are walked in chunks of 32 (lane = sample): the transmittance
// T_i = prod_{j<i}(1 - alpha_j + 1e-10) (RS:520) is an exclusive product scan done with warp shuffles
// plus a carry between chunks; rgb/depth/acc are shuffle reductions.  HBM-bound: 20 B/sample in
// (raw float4 + z), 4 B/sample out (weights), 28 B/ray out.  Loads are 512-B coalesced per warp.
#include "composite.cuh"

namespace scade {

constexpr int COMP_WARPS = 4;

__global__ void __launch_bounds__(COMP_WARPS * 32)
raw2outputs_fwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                       int d_stride, const float* __restrict__ noise, int64_t N, int S, float* __restrict__ rgb_map,
                       float* __restrict__ disp_map, float* __restrict__ acc_map, float* __restrict__ weights,
                       float* __restrict__ depth_map) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
  if (r >= N) return;
  composite_ray_fwd<false>(raw, z, rays_d, d_stride, noise, r, S, lane, rgb_map, disp_map, acc_map, weights, depth_map, nullptr,
                           nullptr);
}

// Backward.  Pass 1 re-runs the forward scan and keeps T_i in shared memory;
// pass 2 walks the chunks in reverse with a suffix-sum carry:
//   g_i = dL/dw_i (direct + through rgb/depth/acc/disp);  dL/dalpha_i = g_i T_i - (sum_{j>i} g_j w_j)/(1-alpha_i+1e-10)
__global__ void __launch_bounds__(COMP_WARPS * 32)
raw2outputs_bwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                       int d_stride, const float* __restrict__ noise, int64_t N, int S,
                       const float* __restrict__ d_rgb_map, const float* __restrict__ d_disp,
                       const float* __restrict__ d_acc, const float* __restrict__ d_w,
                       const float* __restrict__ d_depth, float4* __restrict__ d_raw) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * COMP_WARPS + wid;
  if (r >= N) return;
  float* sT = smem + (size_t)wid * S;
  const float dx = rays_d[r * d_stride], dy = rays_d[r * d_stride + 1], dz = rays_d[r * d_stride + 2];
  const float norm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float4* raw_r = raw + r * S;
  const float* z_r = z + r * S;
  float carry = 1.0f, sdepth = 0.f, sacc = 0.f;
  for (int base = 0; base < S; base += 32) {
    const int i = base + lane;
    const bool valid = i < S;
    float sig_raw = valid ? raw_r[i].w : 0.f;
    float zi = valid ? z_r[i] : 0.f;
    float zn = (i + 1 < S) ? z_r[i + 1] : zi;
    float nz = (noise != nullptr && valid) ? noise[r * S + i] : 0.f;
    SampleTerms t = sample_terms(sig_raw, nz, zi, zn, i == S - 1, norm);
    float tf = valid ? t.tfac : 1.0f;
    float incl = warp_scan_prod(tf, lane);
    float excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 1.0f;
    float T = carry * excl;
    carry *= __shfl_sync(FULL, incl, 31);
    if (valid) sT[i] = T;
    float w = valid ? t.alpha * T : 0.f;
    sdepth += w * zi;
    sacc += w;
  }
  sdepth = warp_sum(sdepth);
  sacc = warp_sum(sacc);
  __syncwarp();
  const float gr = d_rgb_map ? d_rgb_map[r * 3] : 0.f, gg = d_rgb_map ? d_rgb_map[r * 3 + 1] : 0.f,
              gb = d_rgb_map ? d_rgb_map[r * 3 + 2] : 0.f;
  float g_depth = d_depth ? d_depth[r] : 0.f;
  float g_acc = d_acc ? d_acc[r] : 0.f;
  if (d_disp != nullptr) {
    // disp = 1 / max(1e-10, depth/acc): gradient only on the live branch of the max
    float q = sdepth / sacc;
    if (q > 1e-10f) {
      float dq = -d_disp[r] / (q * q);
      g_depth += dq / sacc;
      g_acc += -dq * sdepth / (sacc * sacc);
    }
  }
  float suffix_carry = 0.f;
  const int nchunks = (S + 31) / 32;
  for (int c = nchunks - 1; c >= 0; --c) {
    const int i = c * 32 + lane;
    const bool valid = i < S;
    float4 rw = valid ? raw_r[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float zi = valid ? z_r[i] : 0.f;
    float zn = (i + 1 < S) ? z_r[i + 1] : zi;
    float nz = (noise != nullptr && valid) ? noise[r * S + i] : 0.f;
    SampleTerms t = sample_terms(rw.w, nz, zi, zn, i == S - 1, norm);
    float T = valid ? sT[i] : 0.f;
    float w = t.alpha * T;
    float cr = sigmoidf_(rw.x), cg = sigmoidf_(rw.y), cb = sigmoidf_(rw.z);
    float g = (d_w && valid ? d_w[r * S + i] : 0.f) + gr * cr + gg * cg + gb * cb + g_depth * zi + g_acc;
    float gw = valid ? g * w : 0.f;
    float incl = warp_rscan_sum(gw, lane);
    float suffix = suffix_carry + (incl - gw);
    suffix_carry += __shfl_sync(FULL, incl, 0);
    float d_alpha = g * T - suffix / t.tfac;
    float d_sig = (t.pre > 0.f) ? d_alpha * t.dist * t.e : 0.f;
    if (valid) d_raw[r * S + i] = make_float4(w * cr * (1.f - cr) * gr, w * cg * (1.f - cg) * gg, w * cb * (1.f - cb) * gb, d_sig);
  }
}

}  // namespace scade

using namespace scade;

extern "C" int scade_raw2outputs(const float* raw, const float* z_vals, const float* rays_d, int d_stride,
                                 const float* noise, int64_t N, int S, float* rgb_map, float* disp_map,
                                 float* acc_map, float* weights, float* depth_map, void* stream) {
  SCADE_CHECK_ARG(raw && z_vals && rays_d && N >= 0 && S > 0 && d_stride >= 3, "raw2outputs: bad arguments");
  SCADE_CHECK_ARG((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "raw2outputs: raw must be 16-byte aligned");
  if (N == 0) return SCADE_OK;
  raw2outputs_fwd_kernel<<<(unsigned)ceil_div<int64_t>(N, COMP_WARPS), COMP_WARPS * 32, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(raw), z_vals, rays_d, d_stride, noise, N, S, rgb_map, disp_map, acc_map, weights,
      depth_map);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

extern "C" int scade_raw2outputs_backward(const float* raw, const float* z_vals, const float* rays_d, int d_stride,
                                          const float* noise, int64_t N, int S, const float* d_rgb_map,
                                          const float* d_disp_map, const float* d_acc_map, const float* d_weights,
                                          const float* d_depth_map, float* d_raw, void* stream) {
  SCADE_CHECK_ARG(raw && z_vals && rays_d && d_raw && N >= 0 && S > 0 && d_stride >= 3, "raw2outputs_backward: bad arguments");
  SCADE_CHECK_ARG(((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(d_raw)) & 15) == 0,
                  "raw2outputs_backward: raw/d_raw must be 16-byte aligned");
  size_t smem = (size_t)COMP_WARPS * S * sizeof(float);
  SCADE_CHECK_ARG(smem <= 200 * 1024, "raw2outputs_backward: S too large");
  if (N == 0) return SCADE_OK;
  if (smem > 48 * 1024)
    SCADE_CUDA(cudaFuncSetAttribute(raw2outputs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  raw2outputs_bwd_kernel<<<(unsigned)ceil_div<int64_t>(N, COMP_WARPS), COMP_WARPS * 32, smem, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(raw), z_vals, rays_d, d_stride, noise, N, S, d_rgb_map, d_disp_map, d_acc_map,
      d_weights, d_depth_map, reinterpret_cast<float4*>(d_raw));
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}
