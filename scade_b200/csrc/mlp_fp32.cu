// SCADE_PREC_FP32 field network: Embedder (model/run_nerf_helpers.py:142-172), run_network's point
// normalisation (run_scade_scannet.py:52), NeRF.forward (H:223-247) and its backward, as a sequence of
// fp32 GEMM launches (sgemm.cuh) with fused bias/ReLU/concat/mask epilogues.  Handles any D, W,
// multires, skip position.  This is the arithmetic the reference itself uses (fp32 SGEMM), kept as the
// exactness mode; the throughput path is mlp_tc.cu.
#include <algorithm>

#include "mlp_common.cuh"
#include "sgemm.cuh"

namespace scade {

// ---- encoding ----------------------------------------------------------------------------------
// One thread per (point, 3-channel group): group 0 = identity, 1+2k = sin octave k, 2+2k = cos octave k
// for the position; then the same for the view direction.  Writes are contiguous along the row.
__global__ void encode_rays_kernel(const float* __restrict__ rays, int ray_stride, const float* __restrict__ z,
                                   int64_t N, int S, float cx, float cy, float cz, float bb_scale, int multires,
                                   int multires_views, float* __restrict__ x0, int ldx) {
  const int gp = 1 + 2 * multires, gv = 1 + 2 * multires_views, groups = gp + gv;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * S * groups) return;
  int grp = (int)(idx % groups);
  int64_t pt = idx / groups;
  int64_t r = pt / S;
  const float* ray = rays + r * ray_stride;
  float v[3];
  float* out;
  int g;
  if (grp < gp) {
    float zz = z[pt];
    const float c[3] = {cx, cy, cz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float p = __fadd_rn(ray[a], __fmul_rn(ray[3 + a], zz));       // RS:657  pts = o + d*z
      v[a] = __fmul_rn(__fsub_rn(p, c[a]), bb_scale);               // RS:52
    }
    g = grp;
    out = x0 + pt * ldx + 3 * grp;
  } else {
    v[0] = ray[8]; v[1] = ray[9]; v[2] = ray[10];                   // RS:632 viewdirs
    g = grp - gp;
    out = x0 + pt * ldx + 3 * gp + 3 * g;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) out[a] = embed_channel(v[a], g);
}

__global__ void embed_kernel(const float* __restrict__ x, int64_t P, int multires, float* __restrict__ out) {
  const int groups = 1 + 2 * multires;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * groups) return;
  int grp = (int)(idx % groups);
  int64_t pt = idx / groups;
#pragma unroll
  for (int a = 0; a < 3; ++a) out[pt * 3 * groups + 3 * grp + a] = embed_channel(x[pt * 3 + a], grp);
}

// ---- heads -------------------------------------------------------------------------------------
// alpha_linear (W -> 1) on h_last and rgb_linear (W/2 -> 3) on hv, one warp per point, then
// outputs = cat([rgb, softplus(alpha, beta=10)])  (H:233, H:241-242).
__global__ void heads_fwd_kernel(const float* __restrict__ h_last, int ldh, int Kh, const float* __restrict__ hv,
                                 int Kv, const float* __restrict__ w_alpha, const float* __restrict__ b_alpha,
                                 const float* __restrict__ w_rgb, const float* __restrict__ b_rgb, int64_t P,
                                 float4* __restrict__ out, float* __restrict__ alpha_pre) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= P) return;
  float a = 0.f, r = 0.f, g = 0.f, b = 0.f;
  for (int k = lane; k < Kh; k += 32) a = fmaf(h_last[m * ldh + k], w_alpha[k], a);
  for (int k = lane; k < Kv; k += 32) {
    float v = hv[m * Kv + k];
    r = fmaf(v, w_rgb[k], r);
    g = fmaf(v, w_rgb[Kv + k], g);
    b = fmaf(v, w_rgb[2 * Kv + k], b);
  }
  a = warp_sum(a); r = warp_sum(r); g = warp_sum(g); b = warp_sum(b);
  if (lane == 0) {
    a += b_alpha[0];
    out[m] = make_float4(r + b_rgb[0], g + b_rgb[1], b + b_rgb[2], softplus_beta10(a));
    if (alpha_pre) alpha_pre[m] = a;
  }
}

// d_alpha_pre = d_sigma * softplus'(alpha_pre);  d_zv[m,k] = (sum_j d_rgb[m,j] W_rgb[j,k]) * [hv > 0]
__global__ void heads_bwd_kernel(const float4* __restrict__ d_out, const float* __restrict__ alpha_pre,
                                 const float* __restrict__ hv, int Kv, const float* __restrict__ w_rgb, int64_t P,
                                 float* __restrict__ d_alpha, float* __restrict__ d_zv) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * Kv) return;
  int64_t m = idx / Kv;
  int k = (int)(idx % Kv);
  float4 d = d_out[m];
  float v = d.x * w_rgb[k] + d.y * w_rgb[Kv + k] + d.z * w_rgb[2 * Kv + k];
  d_zv[idx] = hv[idx] > 0.f ? v : 0.f;
  if (k == 0) {
    float ap = alpha_pre[m];
    float bx = ap * 10.0f;
    d_alpha[m] = d.w * (bx > 20.0f ? 1.0f : sigmoidf_(bx));
  }
}

// Gradients of the two skinny heads: gW_rgb[j,k] += sum_m d_rgb[m,j] hv[m,k], gW_alpha[k] += sum_m d_alpha[m] h[m,k]
// and their biases.  Block = slab of rows; thread = column k; one atomicAdd per (thread, output).
__global__ void heads_wgrad_kernel(const float4* __restrict__ d_out, const float* __restrict__ d_alpha,
                                   const float* __restrict__ h_last, int ldh, int Kh, const float* __restrict__ hv,
                                   int Kv, int64_t P, int64_t rows_per_block, float* __restrict__ gw_alpha,
                                   float* __restrict__ gb_alpha, float* __restrict__ gw_rgb,
                                   float* __restrict__ gb_rgb) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(P, r0 + rows_per_block);
  for (int k = threadIdx.x; k < Kh; k += blockDim.x) {
    float acc = 0.f;
    for (int64_t m = r0; m < r1; ++m) acc = fmaf(d_alpha[m], h_last[m * ldh + k], acc);
    atomicAdd(&gw_alpha[k], acc);
  }
  for (int k = threadIdx.x; k < Kv; k += blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int64_t m = r0; m < r1; ++m) {
      float4 d = d_out[m];
      float v = hv[m * Kv + k];
      a0 = fmaf(d.x, v, a0); a1 = fmaf(d.y, v, a1); a2 = fmaf(d.z, v, a2);
    }
    atomicAdd(&gw_rgb[k], a0); atomicAdd(&gw_rgb[Kv + k], a1); atomicAdd(&gw_rgb[2 * Kv + k], a2);
  }
  if (threadIdx.x < 4) {
    float acc = 0.f;
    for (int64_t m = r0; m < r1; ++m) {
      float4 d = d_out[m];
      acc += threadIdx.x == 0 ? d.x : threadIdx.x == 1 ? d.y : threadIdx.x == 2 ? d.z : d_alpha[m];
    }
    if (threadIdx.x < 3) atomicAdd(&gb_rgb[threadIdx.x], acc); else atomicAdd(gb_alpha, acc);
  }
}

// bias gradient: gb[n] += sum_m dz[m,n]
__global__ void colsum_kernel(const float* __restrict__ dz, int64_t ld, int64_t P, int n, int64_t rows_per_block,
                              float* __restrict__ gb) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(P, r0 + rows_per_block);
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    float acc = 0.f;
    for (int64_t m = r0; m < r1; ++m) acc += dz[m * ld + k];
    atomicAdd(&gb[k], acc);
  }
}

// ---- workspace layout ----------------------------------------------------------------------------
struct Fp32Layout {
  size_t x0, feature, hv, alpha_pre, h[32], dz_a, dz_b, dzv, dalpha, total;
  int n_h;
};

static Fp32Layout fp32_layout(const scade_net_desc& d, int64_t P, int save) {
  Fp32Layout L{};
  NetDims nd(d);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  L.x0 = take((size_t)P * nd.in_all * 4);
  L.feature = take((size_t)P * d.W * 4);
  L.hv = take((size_t)P * (d.W / 2) * 4);
  L.alpha_pre = take((size_t)P * 4);
  L.n_h = save ? d.D : 2;
  for (int i = 0; i < L.n_h; ++i) L.h[i] = take((size_t)P * d.W * 4);
  if (save) {
    L.dz_a = take((size_t)P * d.W * 4);
    L.dz_b = take((size_t)P * d.W * 4);
    L.dzv = take((size_t)P * (d.W / 2) * 4);
    L.dalpha = take((size_t)P * 4);
  }
  L.total = off;
  return L;
}

size_t mlp_fp32_workspace_bytes(const scade_net_desc& d, int64_t P, int save) { return fp32_layout(d, P, save).total; }

static int fp32_forward_from_x0(const scade_net& net, int64_t P, char* ws, const Fp32Layout& L, float* out, int save,
                                cudaStream_t st) {
  const scade_net_desc& d = net.desc;
  NetDims nd(d);
  const float* x0 = reinterpret_cast<const float*>(ws + L.x0);
  const float* h_in = x0;
  int64_t ld_in = nd.in_all;
  float* h_out = nullptr;
  for (int i = 0; i < d.D; ++i) {
    const float* Wt = net.params[2 * i];
    const float* b = net.params[2 * i + 1];
    h_out = reinterpret_cast<float*>(ws + L.h[save ? i : (i & 1)]);
    GemmEpilogue ep;
    ep.bias = b; ep.relu = 1;
    if (i == 0) {
      SCADE_TRY(launch_sgemm(0, x0, nd.in_all, Wt, nd.in_ch, h_out, d.W, P, d.W, nd.in_ch, ep, 1, st));
    } else if (i - 1 == d.skip) {
      // h = cat([input_pts, h]) (H:230): two K segments, the second applies bias + ReLU
      int fan_in = nd.in_ch + d.W;
      GemmEpilogue e1;
      SCADE_TRY(launch_sgemm(0, x0, nd.in_all, Wt, fan_in, h_out, d.W, P, d.W, nd.in_ch, e1, 1, st));
      ep.accumulate = 1;
      SCADE_TRY(launch_sgemm(0, h_in, ld_in, Wt + nd.in_ch, fan_in, h_out, d.W, P, d.W, d.W, ep, 1, st));
    } else {
      SCADE_TRY(launch_sgemm(0, h_in, ld_in, Wt, d.W, h_out, d.W, P, d.W, d.W, ep, 1, st));
    }
    h_in = h_out;
    ld_in = d.W;
  }
  const int pv = 2 * d.D;       // views_linears.0, feature_linear, alpha_linear, rgb_linear
  float* feature = reinterpret_cast<float*>(ws + L.feature);
  float* hv = reinterpret_cast<float*>(ws + L.hv);
  int fan_last = d.W;
  const float* h_last = h_in;
  int64_t ld_last = ld_in;
  if (d.skip == d.D - 1) {
    // reference would feed cat([input_pts, h]) to the heads; not reachable with skips=[4], D=8
    set_error("skip after the last layer is not supported");
    return SCADE_ERR_UNSUPPORTED;
  }
  {
    GemmEpilogue ep;
    ep.bias = net.params[pv + 3];
    SCADE_TRY(launch_sgemm(0, h_last, ld_last, net.params[pv + 2], fan_last, feature, d.W, P, d.W, fan_last, ep, 1, st));
  }
  {
    // views_linears.0 on cat([feature, input_views]) (H:235-239)
    int fan_in = d.W + nd.in_views;
    GemmEpilogue e1;
    SCADE_TRY(launch_sgemm(0, feature, d.W, net.params[pv], fan_in, hv, d.W / 2, P, d.W / 2, d.W, e1, 1, st));
    GemmEpilogue ep;
    ep.bias = net.params[pv + 1]; ep.relu = 1; ep.accumulate = 1;
    SCADE_TRY(launch_sgemm(0, x0 + nd.in_ch, nd.in_all, net.params[pv] + d.W, fan_in, hv, d.W / 2, P, d.W / 2,
                           nd.in_views, ep, 1, st));
  }
  float* alpha_pre = save ? reinterpret_cast<float*>(ws + L.alpha_pre) : nullptr;
  heads_fwd_kernel<<<(unsigned)ceil_div<int64_t>(P, 8), 256, 0, st>>>(
      h_last, (int)ld_last, d.W, hv, d.W / 2, net.params[pv + 4], net.params[pv + 5], net.params[pv + 6],
      net.params[pv + 7], P, reinterpret_cast<float4*>(out), alpha_pre);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

int mlp_fp32_forward_rays(const scade_net& net, const float* rays, int ray_stride, const float* z, int64_t N, int S,
                          const float* bb_center, float bb_scale, float* raw_out, void* workspace, size_t ws_bytes,
                          int save, cudaStream_t st) {
  const int64_t P = N * S;
  Fp32Layout L = fp32_layout(net.desc, P, save);
  if (ws_bytes < L.total) {
    set_error("mlp_forward_rays(fp32): workspace %zu < %zu bytes", ws_bytes, L.total);
    return SCADE_ERR_WORKSPACE;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  NetDims nd(net.desc);
  int groups = (1 + 2 * net.desc.multires) + (1 + 2 * net.desc.multires_views);
  int64_t total = P * groups;
  encode_rays_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, st>>>(
      rays, ray_stride, z, N, S, bb_center[0], bb_center[1], bb_center[2], bb_scale, net.desc.multires,
      net.desc.multires_views, reinterpret_cast<float*>(ws + L.x0), nd.in_all);
  SCADE_LAUNCH_CHECK();
  return fp32_forward_from_x0(net, P, ws, L, raw_out, save, st);
}

int mlp_fp32_forward_embedded(const scade_net& net, const float* x, int64_t P, float* out, void* workspace,
                              size_t ws_bytes, int save, cudaStream_t st) {
  Fp32Layout L = fp32_layout(net.desc, P, save);
  if (ws_bytes < L.total) {
    set_error("mlp_forward_embedded(fp32): workspace %zu < %zu bytes", ws_bytes, L.total);
    return SCADE_ERR_WORKSPACE;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  NetDims nd(net.desc);
  SCADE_CUDA(cudaMemcpyAsync(ws + L.x0, x, (size_t)P * nd.in_all * 4, cudaMemcpyDeviceToDevice, st));
  return fp32_forward_from_x0(net, P, ws, L, out, save, st);
}

int mlp_fp32_backward(const scade_net& net, const float* d_out, int64_t P, float* const* grads, void* workspace,
                      size_t ws_bytes, cudaStream_t st) {
  const scade_net_desc& d = net.desc;
  NetDims nd(d);
  Fp32Layout L = fp32_layout(d, P, 1);
  if (ws_bytes < L.total) {
    set_error("mlp_backward(fp32): workspace %zu < %zu bytes", ws_bytes, L.total);
    return SCADE_ERR_WORKSPACE;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  const float* x0 = reinterpret_cast<const float*>(ws + L.x0);
  const float* feature = reinterpret_cast<const float*>(ws + L.feature);
  const float* hv = reinterpret_cast<const float*>(ws + L.hv);
  const float* alpha_pre = reinterpret_cast<const float*>(ws + L.alpha_pre);
  auto H = [&](int i) { return reinterpret_cast<const float*>(ws + L.h[i]); };   // output of pts layer i
  float* dz_a = reinterpret_cast<float*>(ws + L.dz_a);
  float* dz_b = reinterpret_cast<float*>(ws + L.dz_b);
  float* dzv = reinterpret_cast<float*>(ws + L.dzv);
  float* dalpha = reinterpret_cast<float*>(ws + L.dalpha);
  const int pv = 2 * d.D;
  const int Wv = d.W / 2;
  const float* h_last = H(d.D - 1);
  // split-K factor for weight gradients: enough CTAs to fill the machine
  auto splits_for = [&](int M, int N) {
    int tiles = (int)(ceil_div(M, GBM) * ceil_div(N, GBN));
    int s = std::max(1, 2 * num_sms() / std::max(1, tiles));
    int64_t max_s = std::max<int64_t>(1, P / 256);
    return (int)std::min<int64_t>(s, max_s);
  };
  const int64_t rpb = 512;
  const unsigned slabs = (unsigned)ceil_div<int64_t>(P, rpb);
  GemmEpilogue none;
  none.accumulate = 1;   // weight gradients ACCUMULATE into grads[] (atomics when split-K)
  GemmEpilogue plain;

  // heads
  heads_bwd_kernel<<<(unsigned)ceil_div<int64_t>(P * Wv, 256), 256, 0, st>>>(
      reinterpret_cast<const float4*>(d_out), alpha_pre, hv, Wv, net.params[pv + 6], P, dalpha, dzv);
  SCADE_LAUNCH_CHECK();
  heads_wgrad_kernel<<<slabs, 256, 0, st>>>(reinterpret_cast<const float4*>(d_out), dalpha, h_last, d.W, d.W, hv, Wv, P,
                                            rpb, grads[pv + 4], grads[pv + 5], grads[pv + 6], grads[pv + 7]);
  SCADE_LAUNCH_CHECK();
  // views_linears.0: gW += dzv^T [feature, views];  gb += colsum(dzv)
  {
    int fan_in = d.W + nd.in_views;
    SCADE_TRY(launch_sgemm(2, dzv, Wv, feature, d.W, grads[pv], fan_in, Wv, d.W, P, none, splits_for(Wv, d.W), st));
    SCADE_TRY(launch_sgemm(2, dzv, Wv, x0 + nd.in_ch, nd.in_all, grads[pv] + d.W, fan_in, Wv, nd.in_views, P, none,
                           splits_for(Wv, nd.in_views), st));
    colsum_kernel<<<slabs, 256, 0, st>>>(dzv, Wv, P, Wv, rpb, grads[pv + 1]);
    SCADE_LAUNCH_CHECK();
    // d_feature = dzv . Wv[:, :W]
    SCADE_TRY(launch_sgemm(1, dzv, Wv, net.params[pv], fan_in, dz_a, d.W, P, d.W, Wv, plain, 1, st));
  }
  // feature_linear: gW += d_feature^T h_last; gb += colsum;  d_h = d_feature . Wf + d_alpha (x) w_alpha, masked by h_last > 0
  SCADE_TRY(launch_sgemm(2, dz_a, d.W, h_last, d.W, grads[pv + 2], d.W, d.W, d.W, P, none, splits_for(d.W, d.W), st));
  colsum_kernel<<<slabs, 256, 0, st>>>(dz_a, d.W, P, d.W, rpb, grads[pv + 3]);
  SCADE_LAUNCH_CHECK();
  {
    GemmEpilogue ep;
    ep.r1_col = dalpha; ep.r1_row = net.params[pv + 4];
    ep.mask = h_last; ep.ldmask = d.W;
    SCADE_TRY(launch_sgemm(1, dz_a, d.W, net.params[pv + 2], d.W, dz_b, d.W, P, d.W, d.W, ep, 1, st));
  }
  float* dz = dz_b;       // d pre-activation of pts layer i
  float* other = dz_a;
  for (int i = d.D - 1; i >= 0; --i) {
    const float* Wt = net.params[2 * i];
    float* gW = grads[2 * i];
    colsum_kernel<<<slabs, 256, 0, st>>>(dz, d.W, P, d.W, rpb, grads[2 * i + 1]);
    SCADE_LAUNCH_CHECK();
    if (i == 0) {
      SCADE_TRY(launch_sgemm(2, dz, d.W, x0, nd.in_all, gW, nd.in_ch, d.W, nd.in_ch, P, none, splits_for(d.W, nd.in_ch), st));
      break;
    }
    const float* h_prev = H(i - 1);
    if (i - 1 == d.skip) {
      int fan_in = nd.in_ch + d.W;
      SCADE_TRY(launch_sgemm(2, dz, d.W, x0, nd.in_all, gW, fan_in, d.W, nd.in_ch, P, none, splits_for(d.W, nd.in_ch), st));
      SCADE_TRY(launch_sgemm(2, dz, d.W, h_prev, d.W, gW + nd.in_ch, fan_in, d.W, d.W, P, none, splits_for(d.W, d.W), st));
      GemmEpilogue ep;
      ep.mask = h_prev; ep.ldmask = d.W;
      SCADE_TRY(launch_sgemm(1, dz, d.W, Wt + nd.in_ch, fan_in, other, d.W, P, d.W, d.W, ep, 1, st));
    } else {
      SCADE_TRY(launch_sgemm(2, dz, d.W, h_prev, d.W, gW, d.W, d.W, d.W, P, none, splits_for(d.W, d.W), st));
      GemmEpilogue ep;
      ep.mask = h_prev; ep.ldmask = d.W;
      SCADE_TRY(launch_sgemm(1, dz, d.W, Wt, d.W, other, d.W, P, d.W, d.W, ep, 1, st));
    }
    std::swap(dz, other);
  }
  return SCADE_OK;
}

int embed_launch(const float* x, int64_t P, int multires, float* out, cudaStream_t st) {
  int64_t total = P * (1 + 2 * multires);
  embed_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, st>>>(x, P, multires, out);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

}  // namespace scade
