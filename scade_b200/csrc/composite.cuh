// Forward compositing of ONE ray by one warp (compute_weights + raw2outputs, run_scade_scannet.py:511-562), shared by
// raw2outputs_fwd_kernel and the fused compositing + resampling kernel (sample_pdf.cu).
#pragma once
#include "common.cuh"

namespace scade {

struct SampleTerms {
  float alpha, e, dist, tfac, pre;
};

__device__ __forceinline__ SampleTerms sample_terms(float sigma_raw, float noise, float z_i, float z_next, bool last,
                                                    float norm) {
  SampleTerms s;
  float d = last ? 1e10f : (z_next - z_i);            // RS:514-515
  s.dist = d * norm;                                   // RS:516
  s.pre = sigma_raw + noise;                           // RS:518
  float sig = fmaxf(s.pre, 0.0f);                      // act_fn = relu, RS:512
  s.e = expf(-sig * s.dist);
  s.alpha = 1.0f - s.e;
  s.tfac = 1.0f - s.alpha + 1e-10f;                    // RS:520
  return s;
}

// kStage: also leave the ray's z values and weights in shared memory (s_z[S], s_w[S]) for a resampling step that follows in
// the same kernel.
template <bool kStage>
__device__ __forceinline__ void composite_ray_fwd(const float4* __restrict__ raw, const float* __restrict__ z,
                                                  const float* __restrict__ rays_d, int d_stride, const float* __restrict__ noise,
                                                  int64_t r, int S, int lane, float* __restrict__ rgb_map,
                                                  float* __restrict__ disp_map, float* __restrict__ acc_map,
                                                  float* __restrict__ weights, float* __restrict__ depth_map, float* s_z, float* s_w) {
  const float dx = rays_d[r * d_stride], dy = rays_d[r * d_stride + 1], dz = rays_d[r * d_stride + 2];
  const float norm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float4* raw_r = raw + r * S;
  const float* z_r = z + r * S;
  float carry = 1.0f, sr = 0.f, sg = 0.f, sb = 0.f, sdepth = 0.f, sacc = 0.f;
  // Samples are walked 8 chunks (256 samples) at a time: all of a group's loads are issued before its first use, so a ray
  // of up to 256 samples pays one DRAM latency instead of one per chunk; z_{i+1} comes from the neighbouring lane.
  constexpr int G = 8;
  for (int base0 = 0; base0 < S; base0 += 32 * G) {
    float4 rw[G];
    float zz[G];
#pragma unroll
    for (int c = 0; c < G; ++c) {
      const int i = base0 + 32 * c + lane;
      const bool valid = i < S;
      rw[c] = valid ? raw_r[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      zz[c] = valid ? z_r[i] : 0.f;
    }
    const int i_after = base0 + 32 * G;                 // first sample of the next group (only when S > 256)
    const float z_after = (i_after < S) ? z_r[i_after] : 0.f;
#pragma unroll
    for (int c = 0; c < G; ++c) {
      const int base = base0 + 32 * c;
      if (base >= S) break;
      const int i = base + lane;
      const bool valid = i < S;
      const float zi = zz[c];
      const float z_next_chunk = (c + 1 < G) ? __shfl_sync(FULL, zz[c + 1 < G ? c + 1 : c], 0) : z_after;
      float zn = __shfl_down_sync(FULL, zi, 1);
      if (lane == 31) zn = z_next_chunk;
      if (!(i + 1 < S)) zn = zi;
      float nz = (noise != nullptr && valid) ? noise[r * S + i] : 0.f;
      SampleTerms t = sample_terms(rw[c].w, nz, zi, zn, i == S - 1, norm);
      float tf = valid ? t.tfac : 1.0f;
      float incl = warp_scan_prod(tf, lane);
      float excl = __shfl_up_sync(FULL, incl, 1);
      if (lane == 0) excl = 1.0f;
      float w = valid ? t.alpha * (carry * excl) : 0.f;
      carry *= __shfl_sync(FULL, incl, 31);
      if (valid && weights != nullptr) weights[r * S + i] = w;
      if (kStage && valid) { s_z[i] = zi; s_w[i] = w; }
      // explicit fmaf: the compositing fused into the tensor-core MLP epilogue (mlp_tc.cu, kComp) repeats exactly this
      // arithmetic in exactly this order and must produce the same bits
      sr = fmaf(w, sigmoidf_(rw[c].x), sr);              // RS:543, RS:556
      sg = fmaf(w, sigmoidf_(rw[c].y), sg);
      sb = fmaf(w, sigmoidf_(rw[c].z), sb);
      sdepth = fmaf(w, zi, sdepth);                      // RS:558
      sacc += w;                                         // RS:560
    }
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sdepth = warp_sum(sdepth); sacc = warp_sum(sacc);
  if (lane == 0) {
    if (rgb_map) { rgb_map[r * 3] = sr; rgb_map[r * 3 + 1] = sg; rgb_map[r * 3 + 2] = sb; }
    if (depth_map) depth_map[r] = sdepth;
    if (acc_map) acc_map[r] = sacc;
    if (disp_map) {
      float q = sdepth / sacc;                          // RS:559; torch.max propagates the nan of 0/0
      disp_map[r] = (q != q) ? q : 1.0f / fmaxf(1e-10f, q);
    }
  }
}

}  // namespace scade
