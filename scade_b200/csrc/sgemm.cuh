// fp32 FFMA GEMM used by the SCADE_PREC_FP32 MLP path (forward NT, dgrad NN, wgrad TN with split-K).
// This is the "reference arithmetic" mode: the reference runs its nn.Linear layers as true fp32
// SGEMMs, so parity against it can be stated to fp32 round-off.  The tensor-core
// path (mlp_tc.cu) is the fast one; this kernel trades speed for exactness and shape generality.
//
//   C[m,n] (+)= sum_k A(m,k) * B(k,n)      A(m,k) = A[m*a_rs + k*a_cs],  B(k,n) = B[k*b_rs + n*b_cs]
// 128x128x8 tiles, 256 threads, 8x8 outputs per thread, register-staged double buffering.
#pragma once
#include "common.cuh"

namespace scade {

struct GemmEpilogue {
  const float* bias = nullptr;      // [N] added to every row
  int accumulate = 0;               // C += result (reads C)
  int relu = 0;                     // max(., 0)
  const float* mask = nullptr;      // result *= (mask[m*ldmask + n] > 0)   (ReLU backward)
  int64_t ldmask = 0;
  const float* r1_col = nullptr;    // rank-1 term added before the mask: r1_col[m] * r1_row[n]
  const float* r1_row = nullptr;
  int atomic = 0;                   // atomicAdd into C (split-K)
};

struct GemmArgs {
  const float* A; int64_t a_rs, a_cs;
  const float* B; int64_t b_rs, b_cs;
  float* C; int64_t ldc;
  int64_t M; int N; int64_t K;
  int64_t k_per_split;
  GemmEpilogue ep;
};

constexpr int GBM = 128, GBN = 128, GBK = 8, GPAD = 4;

template <bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  __shared__ float As[2][GBK][GBM + GPAD];
  __shared__ float Bs[2][GBK][GBN + GPAD];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * GBM;
  const int n0 = blockIdx.y * GBN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_per_split;
  const int64_t kend = min(g.K, kbeg + g.k_per_split);
  const int ty = tid / 16, tx = tid % 16;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int mm, kk;
      if (A_KCONTIG) { mm = tid / 2; kk = (tid % 2) * 4 + j; } else { kk = tid / 32; mm = (tid % 32) + 32 * j; }
      int64_t m = m0 + mm, k = k0 + kk;
      ra[j] = (m < g.M && k < kend) ? g.A[m * g.a_rs + k * g.a_cs] : 0.f;
      int nn, kb;
      if (B_KCONTIG) { nn = tid / 2; kb = (tid % 2) * 4 + j; } else { kb = tid / 32; nn = (tid % 32) + 32 * j; }
      int n = n0 + nn;
      int64_t k2 = k0 + kb;
      rb[j] = (n < g.N && k2 < kend) ? g.B[k2 * g.b_rs + (int64_t)n * g.b_cs] : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int mm, kk;
      if (A_KCONTIG) { mm = tid / 2; kk = (tid % 2) * 4 + j; } else { kk = tid / 32; mm = (tid % 32) + 32 * j; }
      As[buf][kk][mm] = ra[j];
      int nn, kb;
      if (B_KCONTIG) { nn = tid / 2; kb = (tid % 2) * 4 + j; } else { kb = tid / 32; nn = (tid % 32) + 32 * j; }
      Bs[buf][kb][nn] = rb[j];
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    load_tiles(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += GBK) {
    const bool more = k0 + GBK < kend;
    if (more) load_tiles(k0 + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  const GemmEpilogue& ep = g.ep;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      float v = acc[i][j];
      float* c = g.C + m * g.ldc + n;
      if (ep.atomic) {
        atomicAdd(c, v);
        continue;
      }
      if (ep.accumulate) v += *c;
      if (ep.bias) v += ep.bias[n];
      if (ep.r1_col) v += ep.r1_col[m] * ep.r1_row[n];
      if (ep.relu) v = fmaxf(v, 0.f);
      if (ep.mask) v = ep.mask[m * ep.ldmask + n] > 0.f ? v : 0.f;
      *c = v;
    }
  }
}

// layout: 0 = NT (A[m,k], B[n,k]),  1 = NN (A[m,k], B[k,n]),  2 = TN (A[k,m], B[k,n])
inline int launch_sgemm(int layout, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                        int64_t M, int N, int64_t K, const GemmEpilogue& ep, int splits, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return SCADE_OK;
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.ep = ep;
  if (splits < 1) splits = 1;
  g.k_per_split = ceil_div<int64_t>(ceil_div<int64_t>(K, splits), GBK) * GBK;
  splits = (int)ceil_div<int64_t>(K, g.k_per_split);
  if (splits > 1) g.ep.atomic = 1;
  dim3 grid((unsigned)ceil_div<int64_t>(M, GBM), (unsigned)ceil_div(N, GBN), (unsigned)splits);
  if (layout == 0) {
    g.a_rs = lda; g.a_cs = 1; g.b_rs = 1; g.b_cs = ldb;
    sgemm_kernel<true, true><<<grid, 256, 0, st>>>(g);
  } else if (layout == 1) {
    g.a_rs = lda; g.a_cs = 1; g.b_rs = ldb; g.b_cs = 1;
    sgemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
  } else {
    g.a_rs = 1; g.a_cs = lda; g.b_rs = ldb; g.b_cs = 1;
    sgemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
  }
  SCADE_LAUNCH_CHECK();
  return SCADE_OK;
}

}  // namespace scade
