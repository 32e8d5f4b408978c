// Losses of the SCADE train step.
//   compute_space_carving_loss   model/run_nerf_helpers.py:93-128  (forward + backward in one pass)
//   img2mse                      model/run_nerf_helpers.py:11
//
// Space carving, default branch (H:122-126): loss = mean_n mean_p min_k m_n |pred[n,p] - hyp[k,n]|.
// One warp per ray: lanes walk the P samples, the K hypotheses of the ray sit in shared memory, the
// min over K runs in registers, the sum over p is a shuffle reduction and the per-block partial goes to
// one atomicAdd.  The gradient is written in the same pass (sign of the arg-min term; first k on ties,
// zero at exact equality, like torch.min(dim)/abs).  The reference materialises [K,N,P] distances
// (10.5 MB at 20x4096x128); here traffic is 4*(P + K) B/ray in, 4*(P + K) out.
// Joint branch (H:115-119): mean over rays first -> needs a [K,P] reduction across rays (atomics into
// the workspace), then min over k per p, then a second pass for the gradient.
#include "common.cuh"

namespace scade {

constexpr int SC_WARPS = 8;
// rays per block: 32 (hyp[k, r0 .. r0+31] is one 128-byte line per hypothesis) for large batches; 8 (one ray per warp) when
// the batch is too small to fill the machine with 32-ray blocks (a 512-ray shard of an 8-GPU step would be 16 blocks)

__device__ __forceinline__ float sc_dist(float pred, float h, float m, float thr, float& sgn) {
  float diff = pred - h;
  float d = fabsf(diff) * m;                            // H:106, H:110
  sgn = (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) * m;
  if (thr > 0.f && d < thr) { d = 0.f; sgn = 0.f; }     // H:112-113
  return d;
}

// Block = 8 warps x 32 consecutive rays (4 rays per warp, one after the other).  The [K x 32 rays] tile of hypotheses is
// staged in shared memory by loads that run along N (hyp is [K, N, 1]: ray r of hypothesis k sits at k*N + r, so a per-ray
// walk over k is K scattered 4-byte loads -- 4.6% of the HBM roofline in round 1), and the hypothesis gradients leave the same way.
// Per ray: lanes walk the P samples, min over K in registers, gradient in the same pass, no [K,N,P] tensor.
template <int SC_RAYS>
__global__ void __launch_bounds__(SC_WARPS * 32)
space_carving_ray_kernel(const float* __restrict__ pred, const float* __restrict__ hyp, int hyp_full,
                         const float* __restrict__ mask, int K, int64_t N, int P, float thr, float gscale, float inv_n,
                         float* __restrict__ loss_out, float* __restrict__ d_pred, float* __restrict__ d_hyp,
                         const float* __restrict__ scale_dev, const float* __restrict__ shift_dev, float* __restrict__ d_scale,
                         float* __restrict__ d_shift) {
  extern __shared__ float smem[];   // s_h[K][32 rays] | s_dh[K][32 rays] | per warp: K*32 lane-private gradient accumulators
  __shared__ float s_part[SC_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * SC_RAYS;
  float* s_h = smem;
  float* s_dh = s_h + K * SC_RAYS;
  float* s_g = s_dh + K * SC_RAYS + (size_t)wid * K * 32;
  // affine form (RS:954 fused in): hyp holds the RAW hypotheses, h = hyp * scale + shift is formed here (mul then add, like
  // torch's two ops) and d loss / d scale, d loss / d shift leave as two atomics per block (d_ss[0], d_ss[1])
  const bool affine = scale_dev != nullptr;
  const float a_scale = affine ? scale_dev[0] : 1.0f, a_shift = affine ? shift_dev[0] : 0.0f;
  const bool want_dh = d_hyp != nullptr || d_scale != nullptr;
  if (!hyp_full) {
    for (int k = wid; k < K; k += SC_WARPS) {
      const int64_t r = r0 + lane;
      if (lane < SC_RAYS) {
        float h = r < N ? hyp[(int64_t)k * N + r] : 0.f;
        if (affine) h = __fadd_rn(__fmul_rn(h, a_scale), a_shift);
        s_h[k * SC_RAYS + lane] = h;
      }
    }
  }
  __syncthreads();
  const float gval = gscale * inv_n / (float)P;
  float warp_total = 0.f;
  for (int j = 0; j < SC_RAYS / SC_WARPS; ++j) {
    const int lr = wid * (SC_RAYS / SC_WARPS) + j;
    const int64_t r = r0 + lr;
    if (r >= N) break;
    const float m = mask ? mask[r] : 1.0f;
    if (want_dh && !hyp_full) for (int i = lane; i < K * 32; i += 32) s_g[i] = 0.f;
    __syncwarp();
    float ray_sum = 0.f;
    for (int p = lane; p < P; p += 32) {
      float pr = pred[r * P + p];
      // min over k of m |pred - h_k| (thresholded): m >= 0 is the ray's constant and the threshold map is monotone, so the
      // arg-min is the NEAREST hypothesis -- the loop tracks |pred - h| alone (4 instructions per hypothesis: the kernel is
      // issue-bound at K * P = 2560 distances per ray, not HBM-bound) and mask / threshold / sign are applied once to the
      // winner.  Strict '<' keeps the first k on ties like torch.min(dim) (H:124); where the winner's distance is zeroed
      // (m == 0 or below the threshold) its gradient is zero whichever k is named, so the tie order there is unobservable.
      float best_ad = 0.f, best_h = 0.f;
      int bk = 0;
      if (hyp_full) {
        for (int k = 0; k < K; ++k) {
          const float h = hyp[((int64_t)k * N + r) * P + p];
          const float ad = fabsf(pr - h);
          if (k == 0 || ad < best_ad) { best_ad = ad; best_h = h; bk = k; }
        }
      } else {
        const float* hk = s_h + lr;
        best_h = hk[0];
        best_ad = fabsf(pr - best_h);
#pragma unroll 4
        for (int k = 1; k < K; ++k) {
          const float h = hk[k * SC_RAYS];
          const float ad = fabsf(pr - h);
          if (ad < best_ad) { best_ad = ad; best_h = h; bk = k; }
        }
      }
      float bsgn;
      const float best = sc_dist(pr, best_h, m, thr, bsgn);
      ray_sum += best;
      float g = bsgn * gval;
      if (d_pred) d_pred[r * P + p] = g;
      if (want_dh) {
        if (hyp_full) {
          for (int k = 0; k < K; ++k) d_hyp[((int64_t)k * N + r) * P + p] = (k == bk) ? -g : 0.f;
        } else {
          s_g[bk * 32 + lane] -= g;
        }
      }
    }
    if (want_dh && !hyp_full) {
      __syncwarp();
      for (int k = 0; k < K; ++k) {
        float v = warp_sum(s_g[k * 32 + lane]);
        if (lane == 0) s_dh[k * SC_RAYS + lr] = v;
      }
      __syncwarp();
    }
    warp_total += warp_sum(ray_sum) / (float)P;          // H:125
  }
  if (lane == 0) s_part[wid] = warp_total;
  __syncthreads();
  if (want_dh && !hyp_full) {
    float ps = 0.f, pt = 0.f;                            // d scale = sum d_h * h_raw, d shift = sum d_h  (RS:954)
    for (int k = wid; k < K; k += SC_WARPS) {
      const int64_t r = r0 + lane;
      if (lane < SC_RAYS && r < N) {
        const float g = s_dh[k * SC_RAYS + lane];        // rows of rays beyond N were never written: guarded by r < N
        if (d_hyp) d_hyp[(int64_t)k * N + r] = g;
        if (d_scale) { ps = fmaf(g, hyp[(int64_t)k * N + r], ps); pt += g; }
      }
    }
    if (d_scale) {
      ps = warp_sum(ps);
      pt = warp_sum(pt);
      if (lane == 0) { atomicAdd(d_scale, ps); atomicAdd(d_shift, pt); }
    }
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < SC_WARPS; ++w) s += s_part[w];
    atomicAdd(loss_out, s * inv_n);                      // H:126
  }
}

// ---- joint branch -------------------------------------------------------------------------------
__global__ void sc_joint_accumulate_kernel(const float* __restrict__ pred, const float* __restrict__ hyp, int hyp_full,
                                           const float* __restrict__ mask, int K, int64_t N, int P, float thr,
                                           int64_t rays_per_block, float* __restrict__ qsum /*[K,P]*/) {
  // block handles a slab of rays; thread handles (k, p) pairs -> one atomicAdd per pair per block
  const int64_t r0 = (int64_t)blockIdx.x * rays_per_block;
  const int64_t r1 = min(N, r0 + rays_per_block);
  for (int kp = threadIdx.x; kp < K * P; kp += blockDim.x) {
    int k = kp / P, p = kp % P;
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      float h = hyp_full ? hyp[((int64_t)k * N + r) * P + p] : hyp[(int64_t)k * N + r];
      float sg;
      acc += sc_dist(pred[r * P + p], h, mask ? mask[r] : 1.0f, thr, sg);
    }
    atomicAdd(&qsum[kp], acc);
  }
}

__global__ void sc_joint_select_kernel(const float* __restrict__ qsum, int K, float inv_n, int P, int* __restrict__ kstar,
                                       float* __restrict__ loss_out) {
  // single block; per p: min over k of qsum/N (H:117-118), then mean over p (H:119)
  __shared__ float s_red[32];
  float part = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float best = 0.f;
    int bk = 0;
    for (int k = 0; k < K; ++k) {
      float v = qsum[k * P + p] * inv_n;
      if (k == 0 || v < best) { best = v; bk = k; }
    }
    kstar[p] = bk;
    part += best;
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) s += s_red[w];
    *loss_out = s / (float)P;
  }
}

__global__ void sc_joint_grad_kernel(const float* __restrict__ pred, const float* __restrict__ hyp, int hyp_full,
                                     const float* __restrict__ mask, const int* __restrict__ kstar, int K, int64_t N,
                                     int P, float thr, float gscale, float inv_n, float* __restrict__ d_pred,
                                     float* __restrict__ d_hyp) {
  // one warp per ray (d_hyp [K,N,1] needs a reduction over p)
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= N) return;
  const float m = mask ? mask[r] : 1.0f;
  const float gval = gscale * inv_n / (float)P;
  if (d_hyp && hyp_full)
    for (int k = 0; k < K; ++k)
      for (int p = lane; p < P; p += 32) d_hyp[((int64_t)k * N + r) * P + p] = 0.f;
  if (d_hyp && !hyp_full)
    for (int k = lane; k < K; k += 32) d_hyp[(int64_t)k * N + r] = 0.f;
  __syncwarp();
  for (int p0 = 0; p0 < P; p0 += 32) {
    int p = p0 + lane;
    float g = 0.f;
    int k = 0;
    if (p < P) {
      k = kstar[p];
      float h = hyp_full ? hyp[((int64_t)k * N + r) * P + p] : hyp[(int64_t)k * N + r];
      float sg;
      sc_dist(pred[r * P + p], h, m, thr, sg);
      g = sg * gval;
      if (d_pred) d_pred[r * P + p] = g;
      if (d_hyp && hyp_full) d_hyp[((int64_t)k * N + r) * P + p] = -g;
    }
    if (d_hyp && !hyp_full && p < P) atomicAdd(&d_hyp[(int64_t)k * N + r], -g);
  }
}

// ---- img2mse ------------------------------------------------------------------------------------
__global__ void mse_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, float inv_den,
                           float gscale, float* __restrict__ loss_out, float* __restrict__ d_x) {
  __shared__ float s_red[32];
  float part = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d = x[i] - y[i];
    part += d * d;
    if (d_x) d_x[i] = 2.0f * d * inv_den * gscale;
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) s += s_red[w];
    atomicAdd(loss_out, s * inv_den);
  }
}

}  // namespace scade

using namespace scade;

extern "C" size_t scade_space_carving_workspace_bytes(int K, int64_t N, int P) {
  (void)N;
  return align_up((size_t)K * P * sizeof(float)) + align_up((size_t)P * sizeof(int));
}

template <int SC_RAYS>
static int launch_space_carving_t(const float* pred, const float* hyp, int hyp_full, const float* mask, int K, int64_t N, int P,
                                  float thr, float gscale, float inv_n, float* loss_out, float* d_pred, float* d_hyp,
                                  const float* scale_dev, const float* shift_dev, float* d_scale, float* d_shift, cudaStream_t st) {
  size_t smem = ((size_t)2 * K * SC_RAYS + (size_t)SC_WARPS * K * 32) * sizeof(float);
  SCADE_CHECK_ARG(smem <= 160 * 1024, "space_carving_loss: K=%d too large", K);
  if (smem > 48 * 1024)
    SCADE_CUDA(cudaFuncSetAttribute(space_carving_ray_kernel<SC_RAYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  space_carving_ray_kernel<SC_RAYS><<<(unsigned)ceil_div<int64_t>(N, SC_RAYS), SC_WARPS * 32, smem, st>>>(
      pred, hyp, hyp_full, mask, K, N, P, thr, gscale, inv_n, loss_out, d_pred, d_hyp, scale_dev, shift_dev, d_scale, d_shift);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

static int launch_space_carving(const float* pred, const float* hyp, int hyp_full, const float* mask, int K, int64_t N, int P,
                                float thr, float gscale, float inv_n, float* loss_out, float* d_pred, float* d_hyp,
                                const float* scale_dev, const float* shift_dev, float* d_scale, float* d_shift, cudaStream_t st) {
  if (N >= 8192)
    return launch_space_carving_t<32>(pred, hyp, hyp_full, mask, K, N, P, thr, gscale, inv_n, loss_out, d_pred, d_hyp, scale_dev,
                                      shift_dev, d_scale, d_shift, st);
  return launch_space_carving_t<8>(pred, hyp, hyp_full, mask, K, N, P, thr, gscale, inv_n, loss_out, d_pred, d_hyp, scale_dev,
                                   shift_dev, d_scale, d_shift, st);
}

static int sc_joint_accumulate(const float* pred, const float* hyp, int hyp_full, const float* mask, int K, int64_t N, int P,
                               float threshold, float* qsum, cudaStream_t st) {
  SCADE_CUDA(cudaMemsetAsync(qsum, 0, (size_t)K * P * sizeof(float), st));
  if (N == 0) return SCADE_OK;
  int64_t rays_per_block = 64;
  sc_joint_accumulate_kernel<<<(unsigned)ceil_div<int64_t>(N, rays_per_block), 256, 0, st>>>(
      pred, hyp, hyp_full, mask, K, N, P, threshold, rays_per_block, qsum);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

static int sc_joint_finish(const float* pred, const float* hyp, int hyp_full, const float* mask, const float* qsum, int* kstar,
                           int K, int64_t N, int64_t n_global, int P, float threshold, float grad_scale, float* loss_out,
                           float* d_pred, float* d_hyp, cudaStream_t st) {
  const float inv_n = 1.0f / (float)n_global;
  sc_joint_select_kernel<<<1, 256, 0, st>>>(qsum, K, inv_n, P, kstar, loss_out);
  SCADE_LAUNCH_CHECK();
  if ((d_pred || d_hyp) && N > 0) {
    sc_joint_grad_kernel<<<(unsigned)ceil_div<int64_t>(N, 4), 128, 0, st>>>(pred, hyp, hyp_full, mask, kstar, K, N, P,
                                                                           threshold, grad_scale, inv_n, d_pred, d_hyp);
    SCADE_LAUNCH_CHECK();
  }
  return SCADE_OK;
}

extern "C" int scade_space_carving_loss(const float* pred, const float* hyp, int hyp_full, const float* mask, int K,
                                        int64_t N, int P, int is_joint, float threshold, float grad_scale,
                                        float* loss_out, float* d_pred, float* d_hyp, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  SCADE_CHECK_ARG(pred && hyp && loss_out && K > 0 && N > 0 && P > 0, "space_carving_loss: bad arguments");
  cudaStream_t st = as_stream(stream);
  SCADE_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  if (!is_joint)
    return launch_space_carving(pred, hyp, hyp_full, mask, K, N, P, threshold, grad_scale, 1.0f / (float)N, loss_out, d_pred, d_hyp,
                                nullptr, nullptr, nullptr, nullptr, st);
  size_t need = scade_space_carving_workspace_bytes(K, N, P);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("space_carving_loss(is_joint): workspace %zu < %zu bytes", workspace_bytes, need);
    return SCADE_ERR_WORKSPACE;
  }
  float* qsum = reinterpret_cast<float*>(workspace);
  int* kstar = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + align_up((size_t)K * P * sizeof(float)));
  SCADE_TRY(sc_joint_accumulate(pred, hyp, hyp_full, mask, K, N, P, threshold, qsum, st));
  return sc_joint_finish(pred, hyp, hyp_full, mask, qsum, kstar, K, N, N, P, threshold, grad_scale, loss_out, d_pred, d_hyp, st);
}

extern "C" int scade_space_carving_loss_affine(const float* pred, const float* hyp_raw, const float* scale_dev,
                                               const float* shift_dev, const float* mask, int K, int64_t N, int P,
                                               float threshold, float grad_scale, int64_t denominator, float* loss_out,
                                               float* d_pred, float* d_scale, float* d_shift, int accumulate, void* stream) {
  SCADE_CHECK_ARG(pred && hyp_raw && scale_dev && shift_dev && loss_out && K > 0 && N > 0 && P > 0 && ((d_scale == nullptr) == (d_shift == nullptr)),
                  "space_carving_loss_affine: bad arguments");
  cudaStream_t st = as_stream(stream);
  SCADE_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  if (d_scale && !accumulate) {
    SCADE_CUDA(cudaMemsetAsync(d_scale, 0, sizeof(float), st));
    SCADE_CUDA(cudaMemsetAsync(d_shift, 0, sizeof(float), st));
  }
  const float inv_n = 1.0f / (float)(denominator > 0 ? denominator : N);
  return launch_space_carving(pred, hyp_raw, 0, mask, K, N, P, threshold, grad_scale, inv_n, loss_out, d_pred, nullptr, scale_dev,
                              shift_dev, d_scale, d_shift, st);
}

extern "C" int scade_space_carving_joint_accumulate(const float* pred, const float* hyp, int hyp_full, const float* mask, int K,
                                                    int64_t N, int P, float threshold, float* qsum_out, void* stream) {
  SCADE_CHECK_ARG(pred && hyp && qsum_out && K > 0 && N >= 0 && P > 0, "space_carving_joint_accumulate: bad arguments");
  return sc_joint_accumulate(pred, hyp, hyp_full, mask, K, N, P, threshold, qsum_out, as_stream(stream));
}

extern "C" int scade_space_carving_joint_finish(const float* pred, const float* hyp, int hyp_full, const float* mask,
                                                const float* qsum, int K, int64_t N, int64_t N_global, int P, float threshold,
                                                float grad_scale, float* loss_out, float* d_pred, float* d_hyp,
                                                int32_t* kstar_workspace, void* stream) {
  SCADE_CHECK_ARG(pred && hyp && qsum && loss_out && kstar_workspace && K > 0 && N >= 0 && N_global >= N && N_global > 0 && P > 0,
                  "space_carving_joint_finish: bad arguments");
  cudaStream_t st = as_stream(stream);
  SCADE_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  return sc_joint_finish(pred, hyp, hyp_full, mask, qsum, kstar_workspace, K, N, N_global, P, threshold, grad_scale, loss_out,
                         d_pred, d_hyp, st);
}

extern "C" int scade_img2mse(const float* x, const float* y, int64_t n, int64_t denominator, float grad_scale,
                             float* loss_out, float* d_x, void* stream) {
  SCADE_CHECK_ARG(x && y && loss_out && n > 0, "img2mse: bad arguments");
  cudaStream_t st = as_stream(stream);
  SCADE_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  float inv_den = 1.0f / (float)(denominator > 0 ? denominator : n);
  int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256), 2 * num_sms());
  mse_kernel<<<blocks, 256, 0, st>>>(x, y, n, inv_den, grad_scale, loss_out, d_x);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}
