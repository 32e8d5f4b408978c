// Hierarchical resampling: sample_pdf / sample_pdf_return_u / sample_pdf_joint(_return_u)
// (model/run_nerf_helpers.py:337-538), its backward w.r.t. the weights, and the sort-merge of the
// coarse and importance samples (run_scade_scannet.py:713).
//
// One warp per ray; the ray's cdf and bins live in shared memory.
//   pdf/cdf   : (w + 1e-5)/sum -> inclusive warp-shuffle scan with a carry between 32-wide chunks
//   inversion : per lane, binary search for #{cdf <= u} (searchsorted right=True, H:366), then the
//               below/above gather and the linear interpolation with the denom<1e-5 -> 1 rule (H:368-381)
//   merge     : bitonic sort of the S coarse + n importance values in shared memory (u may be unsorted)
// HBM traffic is 4*(2S-3+2n) B/ray; everything else stays on chip.
#include "composite.cuh"

namespace scade {

constexpr int SP_WARPS = 4;

struct RayPdfSource {
  // explicit form: bins [N,B], weights [N,B-1];  from-z form: z [N,S] (B = S-1), weights_full [N,S]
  const float* bins;
  const float* weights;
  int from_z;
  int B;
};

__device__ __forceinline__ float src_bin(const RayPdfSource& s, int64_t r, int i) {
  if (s.from_z) {
    const float* z = s.bins + r * (s.B + 1);
    return 0.5f * (z[i + 1] + z[i]);                    // RS:702 / RS:723
  }
  return s.bins[r * s.B + i];
}
__device__ __forceinline__ float src_weight(const RayPdfSource& s, int64_t r, int i) {
  if (s.from_z) return s.weights[r * (s.B + 1) + i + 1];   // weights[..., 1:-1]  (RS:705)
  return s.weights[r * (s.B - 1) + i];
}

// Fills s_cdf[0..B) and s_bins[0..B); returns sum(w + 1e-5).  Every global load of the ray (weights, bin edges) is issued
// before the first reduction, and the scan pass re-reads the staged w + 1e-5 from shared memory: one DRAM latency per ray.
__device__ __forceinline__ float build_cdf(const RayPdfSource& src, int64_t r, float* s_cdf, float* s_bins, int lane) {
  const int B = src.B, nw = B - 1;
  float part = 0.f;
  for (int i = lane; i < nw; i += 32) {
    const float w = src_weight(src, r, i) + 1e-5f;                               // H:339
    s_cdf[i + 1] = w;
    part += w;
  }
  for (int i = lane; i < B; i += 32) s_bins[i] = src_bin(src, r, i);
  const float total = warp_sum(part);                                            // H:340
  float carry = 0.f;
  if (lane == 0) s_cdf[0] = 0.f;                                                 // H:343
  for (int base = 0; base < nw; base += 32) {
    int i = base + lane;
    float p = i < nw ? s_cdf[i + 1] / total : 0.f;                               // own write, same lane
    float incl = warp_scan_sum(p, lane) + carry;                                 // H:342
    if (i < nw) s_cdf[i + 1] = incl;
    carry = __shfl_sync(FULL, incl, 31);
  }
  __syncwarp();
  return total;
}

__device__ __forceinline__ int search_right(const float* s_cdf, int B, float u) {
  int lo = 0, hi = B;                                   // first index with cdf > u == #{cdf <= u}
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (s_cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// G independent searches per lane, branch-free and in lock step: #{a[0..len) <= key} (kStrict: #{a < key}) for a sorted
// array in shared memory.  Every level issues its G shared-memory loads back to back, so a lane pays one load latency per
// level instead of one per level and key (the per-key while loop was the sampler's critical path: 8 dependent loads per key).
template <int G, bool kStrict>
__device__ __forceinline__ void count_below(const float* a, int len, const float (&key)[G], int (&pos)[G]) {
#pragma unroll
  for (int g = 0; g < G; ++g) pos[g] = 0;
  int step = 1;
  while ((step << 1) <= len) step <<= 1;
  for (; step > 0; step >>= 1) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int t = pos[g] + step;
      if (t <= len) {
        const float v = a[t - 1];
        if (kStrict ? (v < key[g]) : (v <= key[g])) pos[g] = t;
      }
    }
  }
}

__device__ __forceinline__ float fetch_u(const float* u, int u_is_joint, int64_t r, int n, int j) {
  if (u == nullptr) return torch_linspace(0.f, 1.f, n, j);       // det, H:347
  return u_is_joint ? u[j] : u[r * n + j];
}

__device__ __forceinline__ void bitonic_sort_warp(float* s, int n_pow2, int lane) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < n_pow2 / 2; t += 32) {
        // t-th compare-exchange of this stage: partner indices differ in bit j
        int lo_idx = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int hi_idx = lo_idx | j;
        bool up = (lo_idx & k) == 0;
        float a = s[lo_idx], b = s[hi_idx];
        if ((a > b) == up) { s[lo_idx] = b; s[hi_idx] = a; }
      }
      __syncwarp();
    }
  }
}

// Resampling of ONE ray by one warp: the body of sample_pdf_kernel.  `r_src` indexes the source arrays (0 when they are this
// warp's shared-memory staging of the ray), `r` the outputs.  s_base: 2 B + 2 sort_pow2 floats of shared memory.
__device__ __forceinline__ void resample_ray(const RayPdfSource& src, int64_t r_src, int64_t r, int n, const float* __restrict__ u,
                                             int u_is_joint, float* __restrict__ samples_out, float* __restrict__ u_out,
                                             float* __restrict__ z_merged, float* __restrict__ z_std, int sort_pow2, float* s_base,
                                             int lane) {
  const int B = src.B;
  float* s_cdf = s_base;                                              // cdf | bins | sort area | merge output
  float* s_bins = s_cdf + B;
  float* s_sort = s_bins + B;
  build_cdf(src, r_src, s_cdf, s_bins, lane);
  constexpr int G = 4;
  float ssum = 0.f;
  const bool one_group = n <= 32 * G;                   // (n = 128: every lane holds all of its samples in registers)
  float kept[G];
  for (int j0 = 0; j0 < n; j0 += 32 * G) {
    float uj[G];
    int idx[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int j = j0 + 32 * g + lane;
      uj[g] = j < n ? fetch_u(u, u_is_joint, r, n, j) : 0.f;
    }
    count_below<G, false>(s_cdf, B, uj, idx);           // searchsorted(cdf, u, right=True), H:366
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int j = j0 + 32 * g + lane;
      kept[g] = 0.f;
      if (j >= n) continue;
      int below = max(0, idx[g] - 1);                   // H:368
      int above = min(B - 1, idx[g]);                   // H:369
      float c0 = s_cdf[below], c1 = s_cdf[above];
      float b0 = s_bins[below], b1 = s_bins[above];
      float den = c1 - c0;                              // H:378
      if (den < 1e-5f) den = 1.0f;                      // H:379
      float t = (uj[g] - c0) / den;                     // H:380
      float smp = b0 + t * (b1 - b0);                   // H:381
      samples_out[r * n + j] = smp;
      if (u_out) u_out[r * n + j] = uj[g];
      if (z_merged) s_sort[(B + 1) + j] = smp;
      kept[g] = smp;
      ssum += smp;
    }
  }
  if (z_std != nullptr) {                               // RS:744  torch.std(z_samples, unbiased=False)
    float mean = warp_sum(ssum) / (float)n;
    float v = 0.f;
    if (one_group) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int j = 32 * g + lane;
        if (j < n) { float d = kept[g] - mean; v += d * d; }
      }
    } else {
      for (int j = lane; j < n; j += 32) {
        float d = samples_out[r * n + j] - mean;        // own writes: visible to the writing thread
        v += d * d;
      }
    }
    v = warp_sum(v) / (float)n;
    if (lane == 0) z_std[r] = sqrtf(v);
  }
  if (z_merged != nullptr) {                            // RS:713  sort(cat([z_vals, z_samples]))
    const int S = B + 1, tot = S + n;
    const float* z = src.bins + r_src * S;              // from_z form only
    for (int i = lane; i < S; i += 32) s_sort[i] = z[i];
    __syncwarp();
    // Both lists are normally already sorted (the stratified z_vals always, the importance samples whenever u is -- det
    // sampling): then the sorted concatenation is a merge, and every element's output slot is its own index plus its rank
    // in the other list (one binary search each).  Checked per ray; anything else takes the bitonic network.
    const float* za = s_sort;
    const float* zb = s_sort + S;
    bool sorted = true;
    for (int i = lane; i < S - 1; i += 32) sorted &= za[i] <= za[i + 1];
    for (int j = lane; j < n - 1; j += 32) sorted &= zb[j] <= zb[j + 1];
    if (__all_sync(FULL, sorted)) {
      float* s_out = s_sort + sort_pow2;
      for (int i0 = 0; i0 < S; i0 += 32 * G) {          // #{b < a}: equal values keep the coarse sample first
        float key[G];
        int rank[G];
#pragma unroll
        for (int g = 0; g < G; ++g) { const int i = i0 + 32 * g + lane; key[g] = i < S ? za[i] : 0.f; }
        count_below<G, true>(zb, n, key, rank);
#pragma unroll
        for (int g = 0; g < G; ++g) { const int i = i0 + 32 * g + lane; if (i < S) s_out[i + rank[g]] = key[g]; }
      }
      for (int j0 = 0; j0 < n; j0 += 32 * G) {          // #{a <= b}
        float key[G];
        int rank[G];
#pragma unroll
        for (int g = 0; g < G; ++g) { const int j = j0 + 32 * g + lane; key[g] = j < n ? zb[j] : 0.f; }
        count_below<G, false>(za, S, key, rank);
#pragma unroll
        for (int g = 0; g < G; ++g) { const int j = j0 + 32 * g + lane; if (j < n) s_out[j + rank[g]] = key[g]; }
      }
      __syncwarp();
      for (int i = lane; i < tot; i += 32) z_merged[r * tot + i] = s_out[i];
    } else {
      for (int i = tot + lane; i < sort_pow2; i += 32) s_sort[i] = __int_as_float(0x7f800000);
      __syncwarp();
      bitonic_sort_warp(s_sort, sort_pow2, lane);
      for (int i = lane; i < tot; i += 32) z_merged[r * tot + i] = s_sort[i];
    }
  }
}

__global__ void __launch_bounds__(SP_WARPS * 32)
sample_pdf_kernel(RayPdfSource src, int64_t N, int n, const float* __restrict__ u, int u_is_joint,
                  float* __restrict__ samples_out, float* __restrict__ u_out, float* __restrict__ z_merged,
                  float* __restrict__ z_std, int sort_pow2) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * SP_WARPS + wid;
  if (r >= N) return;
  resample_ray(src, r, r, n, u, u_is_joint, samples_out, u_out, z_merged, z_std, sort_pow2,
               smem + (size_t)wid * (2 * src.B + 2 * sort_pow2), lane);
}

// Compositing + resampling of a ray in ONE kernel (RS:660 + RS:702-713, or RS:720 + RS:723-730): the weights and z values go
// from the compositing scan to the inverse-CDF sampler through shared memory instead of a second kernel re-reading them.
__global__ void __launch_bounds__(SP_WARPS * 32)
composite_resample_kernel(const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int d_stride,
                          int64_t N, int S, float* __restrict__ rgb_map, float* __restrict__ disp_map, float* __restrict__ acc_map,
                          float* __restrict__ weights, float* __restrict__ depth_map, int n, const float* __restrict__ u,
                          int u_is_joint, float* __restrict__ samples_out, float* __restrict__ u_out, float* __restrict__ z_merged,
                          float* __restrict__ z_std, int sort_pow2) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * SP_WARPS + wid;
  if (r >= N) return;
  float* s_z = smem + (size_t)wid * (2 * S + 2 * (S - 1) + 2 * sort_pow2);
  float* s_w = s_z + S;
  composite_ray_fwd<true>(raw, z, rays_d, d_stride, nullptr, r, S, lane, rgb_map, disp_map, acc_map, weights, depth_map, s_z, s_w);
  __syncwarp();
  RayPdfSource src{s_z, s_w, 1, S - 1};
  resample_ray(src, 0, r, n, u, u_is_joint, samples_out, u_out, z_merged, z_std, sort_pow2, s_w + S, lane);
}

// d weights from d samples.  s = b_lo + (u - C_lo)/den * (b_hi - b_lo):
//   den >= 1e-5 : ds/dC_lo = db (u - C_hi)/den^2,  ds/dC_hi = -db (u - C_lo)/den^2
//   den <  1e-5 : den := 1 ->  ds/dC_lo = -db,     ds/dC_hi = 0
// scatter into dC, suffix-sum to d pdf, then d w_j = (d pdf_j - sum_m d pdf_m pdf_m) / sum(w + 1e-5).
__global__ void __launch_bounds__(SP_WARPS * 32)
sample_pdf_bwd_kernel(RayPdfSource src, int64_t N, int n, const float* __restrict__ u,
                      const float* __restrict__ d_samples, float* __restrict__ d_weights, int accumulate) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * SP_WARPS + wid;
  if (r >= N) return;
  const int B = src.B, nw = B - 1;
  float* s_cdf = smem + (size_t)wid * (3 * B);
  float* s_bins = s_cdf + B;
  float* s_dc = s_bins + B;
  const float total = build_cdf(src, r, s_cdf, s_bins, lane);
  for (int i = lane; i < B; i += 32) s_dc[i] = 0.f;
  __syncwarp();
  for (int j = lane; j < n; j += 32) {
    float uj = u[r * n + j];
    float g = d_samples[r * n + j];
    int idx = search_right(s_cdf, B, uj);
    int below = max(0, idx - 1), above = min(B - 1, idx);
    float c0 = s_cdf[below], c1 = s_cdf[above];
    float db = s_bins[above] - s_bins[below];
    float den = c1 - c0;
    float g_lo, g_hi;
    if (den < 1e-5f) { g_lo = -db * g; g_hi = 0.f; }
    else { float inv = 1.0f / (den * den); g_lo = db * (uj - c1) * inv * g; g_hi = -db * (uj - c0) * inv * g; }
    atomicAdd(&s_dc[below], g_lo);
    atomicAdd(&s_dc[above], g_hi);
  }
  __syncwarp();
  // d pdf[m] = sum_{j > m} dC[j]  (m = 0..nw-1): reverse inclusive scan of dC[1..B)
  float dot = 0.f;
  float carry = 0.f;
  const int nchunks = (nw + 31) / 32;
  for (int c = nchunks - 1; c >= 0; --c) {
    int m = c * 32 + lane;
    float v = m < nw ? s_dc[m + 1] : 0.f;
    float incl = warp_rscan_sum(v, lane) + carry;
    carry = __shfl_sync(FULL, incl, 0);
    __syncwarp();
    if (m < nw) {
      s_bins[m] = incl;                                 // reuse: d pdf
      float p = (src_weight(src, r, m) + 1e-5f) / total;
      dot += incl * p;
    }
  }
  dot = warp_sum(dot);
  __syncwarp();
  if (src.from_z) {
    float* out = d_weights + r * (B + 1);
    for (int i = lane; i < B + 1; i += 32) {
      float v = (i >= 1 && i <= nw) ? (s_bins[i - 1] - dot) / total : 0.f;
      out[i] = accumulate ? out[i] + v : v;
    }
  } else {
    float* out = d_weights + r * nw;
    for (int i = lane; i < nw; i += 32) {
      float v = (s_bins[i] - dot) / total;
      out[i] = accumulate ? out[i] + v : v;
    }
  }
}

__global__ void __launch_bounds__(SP_WARPS * 32)
sort_merge_kernel(const float* __restrict__ a, int Na, const float* __restrict__ b, int Nb, int64_t N,
                  float* __restrict__ out, int sort_pow2) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * SP_WARPS + wid;
  if (r >= N) return;
  float* s = smem + (size_t)wid * sort_pow2;
  const int tot = Na + Nb;
  for (int i = lane; i < Na; i += 32) s[i] = a[r * Na + i];
  for (int i = lane; i < Nb; i += 32) s[Na + i] = b[r * Nb + i];
  for (int i = tot + lane; i < sort_pow2; i += 32) s[i] = __int_as_float(0x7f800000);
  __syncwarp();
  bitonic_sort_warp(s, sort_pow2, lane);
  for (int i = lane; i < tot; i += 32) out[r * tot + i] = s[i];
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static int launch_sample(RayPdfSource src, int64_t N, int n, const float* u, int u_is_joint, float* samples_out,
                         float* u_out, float* z_merged, float* z_std, void* stream) {
  int sort_pow2 = z_merged ? next_pow2(src.B + 1 + n) : 0;
  size_t smem = (size_t)SP_WARPS * (2 * src.B + 2 * sort_pow2) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("sample_pdf: %d bins / %d samples exceed the shared-memory budget", src.B, n);
    return SCADE_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    SCADE_CUDA(cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sample_pdf_kernel<<<(unsigned)ceil_div<int64_t>(N, SP_WARPS), SP_WARPS * 32, smem, as_stream(stream)>>>(
      src, N, n, u, u_is_joint, samples_out, u_out, z_merged, z_std, sort_pow2);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

static int launch_composite_resample(const float* raw, const float* z, const float* rays_d, int d_stride, int64_t N, int S,
                                     float* rgb_map, float* disp_map, float* acc_map, float* weights, float* depth_map, int n,
                                     const float* u, int u_is_joint, float* samples_out, float* u_out, float* z_merged,
                                     float* z_std, void* stream) {
  int sort_pow2 = z_merged ? next_pow2(S + n) : 0;
  size_t smem = (size_t)SP_WARPS * (2 * S + 2 * (S - 1) + 2 * sort_pow2) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("composite_resample: %d samples / %d importance samples exceed the shared-memory budget", S, n);
    return SCADE_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    SCADE_CUDA(cudaFuncSetAttribute(composite_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  composite_resample_kernel<<<(unsigned)ceil_div<int64_t>(N, SP_WARPS), SP_WARPS * 32, smem, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(raw), z, rays_d, d_stride, N, S, rgb_map, disp_map, acc_map, weights, depth_map, n, u,
      u_is_joint, samples_out, u_out, z_merged, z_std, sort_pow2);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

static int launch_sample_bwd(RayPdfSource src, int64_t N, int n, const float* u, const float* d_samples,
                             float* d_weights, int accumulate, void* stream) {
  size_t smem = (size_t)SP_WARPS * 3 * src.B * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("sample_pdf_backward: %d bins exceed the shared-memory budget", src.B);
    return SCADE_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    SCADE_CUDA(cudaFuncSetAttribute(sample_pdf_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sample_pdf_bwd_kernel<<<(unsigned)ceil_div<int64_t>(N, SP_WARPS), SP_WARPS * 32, smem, as_stream(stream)>>>(
      src, N, n, u, d_samples, d_weights, accumulate);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

}  // namespace scade

using namespace scade;

extern "C" int scade_sample_pdf(const float* bins, const float* weights, int64_t N, int B, int n_samples,
                                const float* u, int u_is_joint, float* samples_out, float* u_out, void* stream) {
  SCADE_CHECK_ARG(bins && weights && samples_out && N >= 0 && B >= 2 && n_samples > 0, "sample_pdf: bad arguments");
  if (N == 0) return SCADE_OK;
  RayPdfSource src{bins, weights, 0, B};
  return launch_sample(src, N, n_samples, u, u_is_joint, samples_out, u_out, nullptr, nullptr, stream);
}

extern "C" int scade_sample_pdf_backward(const float* bins, const float* weights, const float* u, int64_t N, int B,
                                         int n_samples, const float* d_samples, float* d_weights, void* stream) {
  SCADE_CHECK_ARG(bins && weights && u && d_samples && d_weights && N >= 0 && B >= 2 && n_samples > 0,
                  "sample_pdf_backward: bad arguments");
  if (N == 0) return SCADE_OK;
  RayPdfSource src{bins, weights, 0, B};
  return launch_sample_bwd(src, N, n_samples, u, d_samples, d_weights, 0, stream);
}

extern "C" int scade_resample_from_z(const float* z_vals, const float* weights_full, int64_t N, int S, int n_samples,
                                     const float* u, int u_is_joint, float* samples_out, float* u_out,
                                     float* z_merged, float* z_std, void* stream) {
  SCADE_CHECK_ARG(z_vals && weights_full && samples_out && N >= 0 && S >= 3 && n_samples > 0,
                  "resample_from_z: bad arguments");
  if (N == 0) return SCADE_OK;
  RayPdfSource src{z_vals, weights_full, 1, S - 1};
  return launch_sample(src, N, n_samples, u, u_is_joint, samples_out, u_out, z_merged, z_std, stream);
}

extern "C" int scade_composite_resample(const float* raw, const float* z_vals, const float* rays_d, int d_stride, int64_t N, int S,
                                        float* rgb_map, float* disp_map, float* acc_map, float* weights, float* depth_map,
                                        int n_samples, const float* u, int u_is_joint, float* samples_out, float* u_out,
                                        float* z_merged, float* z_std, void* stream) {
  SCADE_CHECK_ARG(raw && z_vals && rays_d && weights && samples_out && N >= 0 && S >= 3 && n_samples > 0 && d_stride >= 3,
                  "composite_resample: bad arguments");
  SCADE_CHECK_ARG((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "composite_resample: raw must be 16-byte aligned");
  if (N == 0) return SCADE_OK;
  return launch_composite_resample(raw, z_vals, rays_d, d_stride, N, S, rgb_map, disp_map, acc_map, weights, depth_map, n_samples,
                                   u, u_is_joint, samples_out, u_out, z_merged, z_std, stream);
}

extern "C" int scade_resample_from_z_backward(const float* z_vals, const float* weights_full, const float* u,
                                              int64_t N, int S, int n_samples, const float* d_samples,
                                              float* d_weights_full, int accumulate, void* stream) {
  SCADE_CHECK_ARG(z_vals && weights_full && u && d_samples && d_weights_full && N >= 0 && S >= 3 && n_samples > 0,
                  "resample_from_z_backward: bad arguments");
  if (N == 0) return SCADE_OK;
  RayPdfSource src{z_vals, weights_full, 1, S - 1};
  return launch_sample_bwd(src, N, n_samples, u, d_samples, d_weights_full, accumulate, stream);
}

extern "C" int scade_sort_merge(const float* a, int Na, const float* b, int Nb, int64_t N, float* out, void* stream) {
  SCADE_CHECK_ARG(a && b && out && N >= 0 && Na >= 0 && Nb >= 0 && Na + Nb > 0, "sort_merge: bad arguments");
  if (N == 0) return SCADE_OK;
  int p2 = next_pow2(Na + Nb);
  size_t smem = (size_t)SP_WARPS * p2 * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("sort_merge: %d values per ray exceed the shared-memory budget", Na + Nb);
    return SCADE_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    SCADE_CUDA(cudaFuncSetAttribute(sort_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sort_merge_kernel<<<(unsigned)ceil_div<int64_t>(N, SP_WARPS), SP_WARPS * 32, smem, as_stream(stream)>>>(a, Na, b, Nb,
                                                                                                         N, out, p2);
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}
